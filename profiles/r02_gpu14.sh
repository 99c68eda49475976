#!/bin/bash
# Round 2, GPU call 14: where a step of the small named configurations goes (host time per pass issue, kernel list).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
rm -f $O/r02_small_trace_v15.txt
for ex in Nanoblock_Scattering PhC_Band_Diagram Dipole_PML; do
	d=/tmp/prof_$ex; rm -rf $d; mkdir -p $d; cp tests/bin/refex_${ex}_cuda $d/$ex
	( cd $d; s=$(date +%s.%N); ALIS_PROFILE=1 ALISB_DEBUG=2 ./$ex > out.txt 2> prof.txt; e=$(date +%s.%N); echo "== $ex wall $(python -c "print(round($e-$s,3))") s"; grep -v "^\[alisb\] set\|^\[alisb\] fused\|create:\|cells \|rows \|plan" prof.txt | tail -45 ) >> $O/r02_small_trace_v15.txt 2>&1
done
cat $O/r02_small_trace_v15.txt
