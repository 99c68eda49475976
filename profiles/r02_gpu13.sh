#!/bin/bash
# Round 2, GPU call 13: fused launches on a high-priority stream (ALISB_FUSE_PRIO) against equal priorities.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or fusion or slabs" ) > $O/r02_pytest_v13.log 2>&1
tail -4 $O/r02_pytest_v13.log
rm -f $O/r02_prio_ab_v13.txt
for cfg in "1 1" "0 1" "1 0" "1 1" "0 1"; do
	set -- $cfg
	echo "ALISB_FUSE_PRIO=$1 ALISB_CONC=$2" >> $O/r02_prio_ab_v13.txt
	ALISB_FUSE_PRIO=$1 ALISB_CONC=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>> $O/r02_prio_ab_v13.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('  value %.0f Mcells/s  ms/step %.3f  dominant avg %.3f ms  frac_moved %.3f  fused_launches %s launches %s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac_of_moved_bytes'], d.get('fused_launches'), d.get('gpu_launches')))
" >> $O/r02_prio_ab_v13.txt
done
cat $O/r02_prio_ab_v13.txt
