#!/bin/bash
# Round 2, GPU call 8: z CPML slabs inside the fused kernel (yee_fused<..,PMZ>), boundary copies as three concurrent chains.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $O/r02_pytest_v8.log 2>&1
tail -15 $O/r02_pytest_v8.log
rm -f $O/r02_pmz_ab_v8.txt
for cfg in "1 1" "0 1" "1 0"; do
	set -- $cfg
	echo "ALISB_FUSE_PMZ=$1 ALISB_CONC=$2" >> $O/r02_pmz_ab_v8.txt
	ALISB_FUSE_PMZ=$1 ALISB_CONC=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>> $O/r02_pmz_ab_v8.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('  value %.0f Mcells/s  ms/step %.3f  dominant avg %.3f ms  frac_moved %.3f  fused_launches %s launches %s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac_of_moved_bytes'], d.get('fused_launches'), d.get('gpu_launches')))
" >> $O/r02_pmz_ab_v8.txt
done
cat $O/r02_pmz_ab_v8.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_bench_n1_v8.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
tail -30 $O/r02_launches_bench_n1_v8.csv | cut -d, -f5,8,9,15 | cut -c1-150
