#!/bin/bash
# Round 2, GPU call 5: yee_wave with batched helpers; ADE box cut; device-resident slice sums; start-up trace.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_monitor_eval.py -m gpu -q -x ) > $O/r02_pytest_v5.log 2>&1
tail -5 $O/r02_pytest_v5.log
rm -f $O/r02_wave_ab_v5.txt
for cfg in "0 16 64" "1 16 64" "1 16 113" "1 12 64"; do
	set -- $cfg
	echo "ALISB_WAVE=$1 ALISB_FROWS=$2 ALISB_FCHUNK=$3" >> $O/r02_wave_ab_v5.txt
	ALISB_WAVE=$1 ALISB_FROWS=$2 ALISB_FCHUNK=$3 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>> $O/r02_wave_ab_v5.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('  value %.0f Mcells/s  ms/step %.3f  dominant avg %.3f ms  frac_moved %.3f  fused_launches %s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac_of_moved_bytes'], d.get('fused_launches')))
" >> $O/r02_wave_ab_v5.txt
done
cat $O/r02_wave_ab_v5.txt
ALISB_WAVE=1 timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -k regex:"yee_wave" \
	-o $O/r02_ncu_wave_v5 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_ncu_wave_v5.log 2>&1
B=tests/bin/bench_cube_mine_cuda
rm -f $O/r02_ade_bench_v5.txt
for v in "" "au"; do for fuse in 1 0; do
	[ -z "$v" ] && [ $fuse = 0 ] && continue
	echo "variant=[$v] ALISB_ADE_FUSE=$fuse $(ALISB_ADE_FUSE=$fuse $B 512 512 512 40 5 0 $v | tail -1)" >> $O/r02_ade_bench_v5.txt
done; done
cat $O/r02_ade_bench_v5.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file $O/r02_launches_au512_v5.csv $B 512 512 512 3 1 0 au > /dev/null 2>&1
rm -f $O/r02_named_profiles_v5.txt
for ex in Dipole_PML Nanoblock_Scattering PhC_Band_Diagram Metasurface_Farfield Dipole_Band; do
	d=/tmp/prof_$ex; rm -rf $d; mkdir -p $d; cp tests/bin/refex_${ex}_cuda $d/$ex
	( cd $d; s=$(date +%s.%N); ALIS_PROFILE=1 ALISB_DEBUG=2 ./$ex > out.txt 2> prof.txt; e=$(date +%s.%N); echo "== $ex wall $(python -c "print(round($e-$s,3))") s"; grep -v "^\[alisb\] set\|^\[alisb\] fused" prof.txt | head -60 ) >> $O/r02_named_profiles_v5.txt 2>&1
done
cat $O/r02_named_profiles_v5.txt
( time timeout 900 python -m pytest tests/test_refex.py -m gpu -q -x -k "Dipole_PML or Nanoblock or PhC_Band or Metasurface or Dipole_Band or Nanowire" ) > $O/r02_pytest_refex_v5.log 2>&1
tail -4 $O/r02_pytest_refex_v5.log
