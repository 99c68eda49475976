#!/bin/bash
# Round 2, final validation on one GPU: whole GPU suite, smoke, both bench arms, launch list and the ncu capture the bench line's
# `traffic` is taken from (profiles/dominant_kernel_traffic.json carries the sha of the kernel sources it was captured from).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out; T=${1:-final}
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
( time timeout 2400 python -m pytest tests -m gpu -q -x --durations=8 ) > $O/r02_pytest_gpu_$T.log 2>&1
tail -16 $O/r02_pytest_gpu_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke_$T.log 2>&1; tail -2 $O/r02_smoke_$T.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1_$T.json 2> $O/r02_bench_n1_$T.err; tail -c 2600 $O/r02_bench_n1_$T.json
python bench.py --impl reference --steps 10 --warmup 3 > $O/r02_bench_ref_$T.json 2> $O/r02_bench_ref_$T.err; tail -c 700 $O/r02_bench_ref_$T.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_n1_$T.csv \
	python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 2 -k regex:"yee_fused" \
	-o $O/r02_ncu_fused_$T python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_ncu_fused_$T.log 2>&1
ncu -i $O/r02_ncu_fused_$T.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size > $O/r02_ncu_fused_${T}_raw.csv 2>&1
cat $O/r02_ncu_fused_${T}_raw.csv | cut -d, -f5,9,12- | tail -3
ls -la $O | tail -15
