#!/bin/bash
# Round 2, GPU calls 16, 17: list kernels walked as per-array chains, ADE current loads batched, real / imaginary volume launches
# side by side, flux_finish with parallel loads: parity, then the step of the small named configurations again.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 1500 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_examples.py tests/test_zzz_fuzz_reference.py -m gpu -q -x -k "not dipole_band" ) > $O/r02_pytest_v17.log 2>&1
tail -5 $O/r02_pytest_v17.log
( time timeout 1500 python -m pytest tests/test_refex.py tests/test_shards.py -m gpu -q -x -k "Dipole_PML or Nanoblock or PhC_Band or Metasurface or Dipole_Band or nanoblock or metasurface or phc" ) > $O/r02_pytest_refex_v17.log 2>&1
tail -5 $O/r02_pytest_refex_v17.log
grep -h REFEX_TIMING $O/r02_pytest_refex_v17.log | head
cp gpurun_out/refex_timing.json $O/r02_refex_timing_v17.json 2>/dev/null
rm -f $O/r02_small_trace_v17.txt
for ex in Nanoblock_Scattering PhC_Band_Diagram Metasurface_Farfield Dipole_Band; do
	d=/tmp/prof_$ex; rm -rf $d; mkdir -p $d; cp tests/bin/refex_${ex}_cuda $d/$ex
	( cd $d; s=$(date +%s.%N); ALIS_PROFILE=1 ./$ex > out.txt 2> prof.txt; e=$(date +%s.%N); echo "== $ex wall $(python -c "print(round($e-$s,3))") s"; cat prof.txt | tail -16 ) >> $O/r02_small_trace_v17.txt 2>&1
done
cat $O/r02_small_trace_v17.txt
for ex in Nanoblock_Scattering PhC_Band_Diagram; do
	d=/tmp/ncu_$ex; rm -rf $d; mkdir -p $d; cp tests/bin/refex_${ex}_cuda $d/$ex
	( cd $d; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20000 -c 40 --csv --log-file $O/r02_launches_${ex}_v17.csv ./$ex > /dev/null 2>&1 )
	python - $O/r02_launches_${ex}_v17.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; n=0
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); n+=1
        if n<=24: print(d['Kernel Name'][:44], d['Grid Size'], d['Block Size'], d['Metric Value'])
PY
done
