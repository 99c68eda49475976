#!/bin/bash
# Round 2, GPU call 2: the in-place wavefront fused kernel (yee_wave): parity, A/B timing against yee_fused, ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or fusion" ) > $O/r02_pytest_wave_v2.log 2>&1
tail -15 $O/r02_pytest_wave_v2.log
for cfg in "0 16 64" "1 16 64" "1 16 32" "1 16 113" "1 12 64" "1 8 64"; do
	set -- $cfg
	echo "ALISB_WAVE=$1 ALISB_FROWS=$2 ALISB_FCHUNK=$3" >> $O/r02_wave_ab_v2.txt
	ALISB_WAVE=$1 ALISB_FROWS=$2 ALISB_FCHUNK=$3 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>> $O/r02_wave_ab_v2.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('  value %.0f Mcells/s  ms/step %.3f  dominant avg %.3f ms  frac_moved %.3f  fused_launches %s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac_of_moved_bytes'], d.get('fused_launches')))
" >> $O/r02_wave_ab_v2.txt
done
cat $O/r02_wave_ab_v2.txt
for w in 0 1; do
	ALISB_WAVE=$w timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -k regex:"yee_fused|yee_wave" \
		-o $O/r02_ncu_fusedkernel_wave$w python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_ncu_fusedkernel_wave$w.log 2>&1
	ncu -i $O/r02_ncu_fusedkernel_wave$w.ncu-rep --page details > $O/r02_ncu_fusedkernel_wave$w.txt 2>&1
done
ls -la $O | tail -12
