#!/bin/bash
# Round 2, GPU call 6: yee_wave lag sweep; remaining stalls of yee_fused; the launch sequence of one PhC step; flux forecast.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or fusion" ) > $O/r02_pytest_v6.log 2>&1
tail -5 $O/r02_pytest_v6.log | head -3
rm -f $O/r02_wave_ab_v6.txt
for cfg in "0 16 64 0" "1 16 64 0" "1 16 64 2" "1 16 64 4" "1 16 64 8" "1 16 113 4" "1 16 226 4"; do
	set -- $cfg
	echo "ALISB_WAVE=$1 ALISB_FROWS=$2 ALISB_FCHUNK=$3 ALISB_WAVE_LAG=$4" >> $O/r02_wave_ab_v6.txt
	ALISB_WAVE=$1 ALISB_FROWS=$2 ALISB_FCHUNK=$3 ALISB_WAVE_LAG=$4 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>> $O/r02_wave_ab_v6.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('  value %.0f Mcells/s  ms/step %.3f  dominant avg %.3f ms  frac_moved %.3f  fused_launches %s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac_of_moved_bytes'], d.get('fused_launches')))
" >> $O/r02_wave_ab_v6.txt
done
cat $O/r02_wave_ab_v6.txt
ALISB_WAVE=0 timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -k regex:"yee_fused" \
	-o $O/r02_ncu_fused_v6 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_ncu_fused_v6.log 2>&1
mkdir -p /tmp/phc && cp tests/bin/phc_band_diagram_cuda /tmp/phc/ && ( cd /tmp/phc && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30000 -c 80 --csv --log-file $O/r02_launches_phc_v6.csv ./phc_band_diagram_cuda 1.2 1500 > /dev/null 2>&1 )
tail -40 $O/r02_launches_phc_v6.csv | cut -d, -f5,8,9,15 | head -40
rm -f $O/r02_named_profiles_v6.txt
for ex in Nanoblock_Scattering PhC_Band_Diagram; do
	d=/tmp/prof_$ex; rm -rf $d; mkdir -p $d; cp tests/bin/refex_${ex}_cuda $d/$ex
	( cd $d; s=$(date +%s.%N); ALIS_PROFILE=1 ./$ex > out.txt 2> prof.txt; e=$(date +%s.%N); echo "== $ex wall $(python -c "print(round($e-$s,3))") s"; cat prof.txt ) >> $O/r02_named_profiles_v6.txt 2>&1
done
cat $O/r02_named_profiles_v6.txt
