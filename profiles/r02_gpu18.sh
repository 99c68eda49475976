#!/bin/bash
# Round 2, GPU call 18: boundary-copy lists resolved into one gather (copy_resolved), mailbox results (box_gather_mail,
# flux_mail), TFSF auxiliary worlds on a side stream: parity first, then the step of the small named configurations,
# launch lists, and one full ncu capture of the volume kernels of Nanoblock and PhC.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out; T=${1:-v18}
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 900 python -m pytest tests/test_copy_resolve.py tests/test_gpu_dropin.py tests/test_gpu_examples.py tests/test_zzz_fuzz_reference.py tests/test_monitor_eval.py -m gpu -q -x -k "not dipole_band" ) > $O/r02_pytest_$T.log 2>&1
tail -5 $O/r02_pytest_$T.log
( time timeout 900 python -m pytest tests/test_refex.py tests/test_shards.py -m gpu -q -x -k "Dipole_PML or Nanoblock or PhC_Band or Metasurface or Dipole_Band or nanoblock or metasurface or phc" ) > $O/r02_pytest_refex_$T.log 2>&1
tail -5 $O/r02_pytest_refex_$T.log
cp gpurun_out/refex_timing.json $O/r02_refex_timing_$T.json 2>/dev/null
rm -f $O/r02_small_trace_$T.txt
for ex in Nanoblock_Scattering PhC_Band_Diagram Metasurface_Farfield Dipole_Band Dipole_PML; do
	d=/tmp/prof_$ex; rm -rf $d; mkdir -p $d; cp tests/bin/refex_${ex}_cuda $d/$ex
	( cd $d; s=$(date +%s.%N); ALIS_PROFILE=1 ./$ex > out.txt 2> prof.txt; e=$(date +%s.%N); echo "== $ex wall $(python -c "print(round($e-$s,3))") s"; cat prof.txt | tail -16 ) >> $O/r02_small_trace_$T.txt 2>&1
done
# the same with the three changes off, one by one (A/B on the two device-bound examples)
for knob in ALISB_RESOLVE ALISB_MAILBOX ALISB_AUX_ASIDE; do
	for ex in Nanoblock_Scattering PhC_Band_Diagram; do
		d=/tmp/prof_$ex
		( cd $d; s=$(date +%s.%N); env $knob=0 ./$ex > out.txt 2> prof.txt; e=$(date +%s.%N); echo "== $ex with $knob=0 wall $(python -c "print(round($e-$s,3))") s" ) >> $O/r02_small_trace_$T.txt 2>&1
	done
done
cat $O/r02_small_trace_$T.txt
for ex in Nanoblock_Scattering PhC_Band_Diagram; do
	d=/tmp/ncu_$ex; rm -rf $d; mkdir -p $d; cp tests/bin/refex_${ex}_cuda $d/$ex
	( cd $d; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20000 -c 40 --csv --log-file $O/r02_launches_${ex}_$T.csv ./$ex > /dev/null 2>&1 )
	python - $O/r02_launches_${ex}_$T.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; n=0
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); n+=1
        if n<=24: print(d['Kernel Name'][:44], d['Grid Size'], d['Block Size'], d['Metric Value'])
PY
	( cd $d; timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 20000 -c 4 -k regex:"yee_volume" -o $O/r02_ncu_${ex}_vol_$T ./$ex > /dev/null 2>&1 )
done
ls -la $O | tail -12
