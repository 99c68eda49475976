#!/bin/bash
# Round 2, GPU call 25: writeSpectrum's running DFT on a worker thread: the two drivers that call it against the reference, and the
# loop time of PhC_Band_Diagram with and without the worker.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out; T=${1:-v25}
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 300 python -m pytest tests/test_refex.py -m gpu -q -x -k "PhC_Band or Dipole_Band" ) > $O/r02_pytest_refex_$T.log 2>&1
tail -4 $O/r02_pytest_refex_$T.log
rm -f $O/r02_small_ab_$T.txt
run() {   # example, label, env...
	ex=$1; label=$2; shift 2
	d=/tmp/ab_$$_$RANDOM; rm -rf $d; mkdir -p $d; cp tests/bin/refex_${ex}_cuda $d/$ex
	( cd $d; env ALIS_PROFILE=1 "$@" ./$ex > out.txt 2> prof.txt
	  python - "$ex" "$label" <<'PY'
import sys,re
tot=init=0.0; rows={}
for l in open('prof.txt'):
    m=re.match(r'\[alis-b200 profile\] (.+?)\s+(\d+)?\s+([\d.]+)\s*$', l)
    if not m: continue
    name=m.group(1).strip(); v=float(m.group(3))
    if name.startswith('engine create'): init=v
    elif name.startswith('since first'): tot=v
    else: rows[name]=v
print("%-22s %-22s loop %.3f s (start-up %.3f)  get %.3f writeSpectrum %.3f sliceFreqDom %.3f" % (sys.argv[1], sys.argv[2], tot-init, init, rows.get('get',0), rows.get('writeSpectrum',0), rows.get('sliceFreqDom',0)))
PY
	) >> $O/r02_small_ab_$T.txt 2>&1
	rm -rf $d
}
run PhC_Band_Diagram default
run PhC_Band_Diagram spectrum-thread=0 ALIS_SPECTRUM_THREAD=0
run PhC_Band_Diagram default-again
run Dipole_Band default
cat $O/r02_small_ab_$T.txt
