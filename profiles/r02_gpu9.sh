#!/bin/bash
# Round 2, GPU call 9: z slabs inside the fused kernel with 15-row tiles (16 warps: 128 registers per thread, no spills) against 16-row tiles.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $O/r02_pytest_v9.log 2>&1
tail -5 $O/r02_pytest_v9.log
( ALISB_FROWS=15 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused" ) > $O/r02_pytest_v9_rows15.log 2>&1
tail -3 $O/r02_pytest_v9_rows15.log
rm -f $O/r02_pmz_ab_v9.txt
for cfg in "0 16 64" "0 15 64" "1 15 64" "1 15 128" "1 16 64"; do
	set -- $cfg
	echo "ALISB_FUSE_PMZ=$1 ALISB_FROWS=$2 ALISB_FCHUNK=$3" >> $O/r02_pmz_ab_v9.txt
	ALISB_FUSE_PMZ=$1 ALISB_FROWS=$2 ALISB_FCHUNK=$3 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>> $O/r02_pmz_ab_v9.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('  value %.0f Mcells/s  ms/step %.3f  dominant avg %.3f ms  frac_moved %.3f  fused_launches %s launches %s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac_of_moved_bytes'], d.get('fused_launches'), d.get('gpu_launches')))
" >> $O/r02_pmz_ab_v9.txt
done
cat $O/r02_pmz_ab_v9.txt
ALISB_FUSE_PMZ=1 ALISB_FROWS=15 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_bench_n1_v9.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_launches_bench_n1_v9.csv')))
hdr=None
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); print(d['Kernel Name'][:50], d['Grid Size'], d['Block Size'], d['Metric Value'])
PY
