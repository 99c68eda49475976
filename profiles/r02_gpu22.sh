#!/bin/bash
# Round 2, GPU call 22: leaner mailbox (a probe value and its sequence number in one 8-byte store, flux sums carried to the host by
# the last plane with one system-scope fence): parity of everything that reads through it, loop times, then the bench line and
# the ncu capture of the dominant kernel for the final kernel sources.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out; T=${1:-v22}
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 900 python -m pytest tests/test_copy_resolve.py tests/test_gpu_dropin.py tests/test_gpu_examples.py tests/test_monitor_eval.py tests/test_gpu_parity.py -m gpu -q -x -k "not dipole_band and not at_scale" ) > $O/r02_pytest_$T.log 2>&1
tail -5 $O/r02_pytest_$T.log
( time timeout 900 python -m pytest tests/test_refex.py tests/test_shards.py -m gpu -q -x -k "Dipole_PML or Nanoblock or PhC_Band or Metasurface or Dipole_Band or nanoblock or metasurface or phc" ) > $O/r02_pytest_refex_$T.log 2>&1
tail -5 $O/r02_pytest_refex_$T.log
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
	run $ex mailbox=0 ALISB_MAILBOX=0
done
for ex in Metasurface_Farfield Dipole_Band Dipole_PML; do run $ex default; done
cat $O/r02_small_ab_$T.txt
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1_$T.json 2> $O/r02_bench_n1_$T.err; tail -c 600 $O/r02_bench_n1_$T.json
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 2 -k regex:"yee_fused" \
	-o $O/r02_ncu_fused_$T python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_ncu_fused_$T.log 2>&1
ncu -i $O/r02_ncu_fused_$T.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size > $O/r02_ncu_fused_${T}_raw.csv 2>&1
cat $O/r02_ncu_fused_${T}_raw.csv | cut -d, -f5,9,12- | tail -3
for ex in PhC_Band_Diagram Nanoblock_Scattering; do
	d=/tmp/ncu_$ex; rm -rf $d; mkdir -p $d; cp tests/bin/refex_${ex}_cuda $d/$ex
	( cd $d; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20000 -c 24 --csv --log-file $O/r02_launches_${ex}_$T.csv $d/$ex > /dev/null 2>&1 )
	python - $O/r02_launches_${ex}_$T.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; n=0
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); n+=1
        if n<=16: print(d['Kernel Name'][:44], d['Grid Size'], d['Block Size'], d['Metric Value'])
PY
done
