#!/bin/bash
# Round 2, GPU call 23 (call 22 ran the previous build: its engine had failed to compile here and the old library travelled):
# leaner mailbox -- a probe value and its sequence number in one 8-byte store, flux sums carried to the host by the last plane
# with one system-scope fence.  Parity of what reads through it, loop times, bench line, ncu capture for these kernel sources.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out; T=${1:-v23}
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
strings alis_b200/libalis_b200_engine.so | grep -c "torn probe word" | sed 's/^/engine has the new mailbox: /'
( time timeout 600 python -m pytest tests/test_copy_resolve.py tests/test_gpu_dropin.py tests/test_monitor_eval.py -m gpu -q -x ) > $O/r02_pytest_$T.log 2>&1
tail -4 $O/r02_pytest_$T.log
( time timeout 600 python -m pytest tests/test_refex.py -m gpu -q -x -k "Dipole_PML or Nanoblock or PhC_Band or Metasurface or Dipole_Band" ) > $O/r02_pytest_refex_$T.log 2>&1
tail -4 $O/r02_pytest_refex_$T.log
cp gpurun_out/refex_timing.json $O/r02_refex_timing_$T.json 2>/dev/null
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
print("%-22s %-22s loop %.3f s (start-up %.3f)  get %.3f poynting %.3f sliceSnap %.3f" % (sys.argv[1], sys.argv[2], tot-init, init, rows.get('get',0), rows.get('poyntingOut',0), rows.get('sliceSnap',0)))
PY
	) >> $O/r02_small_ab_$T.txt 2>&1
	rm -rf $d
}
for ex in Nanoblock_Scattering PhC_Band_Diagram; do
	run $ex default
	run $ex default-again
done
cat $O/r02_small_ab_$T.txt
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1_$T.json 2> $O/r02_bench_n1_$T.err; tail -c 400 $O/r02_bench_n1_$T.json
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 2 -k regex:"yee_fused" \
	-o $O/r02_ncu_fused_$T python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_ncu_fused_$T.log 2>&1
ncu -i $O/r02_ncu_fused_$T.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size > $O/r02_ncu_fused_${T}_raw.csv 2>&1
cat $O/r02_ncu_fused_${T}_raw.csv | cut -d, -f5,9,12- | tail -3
