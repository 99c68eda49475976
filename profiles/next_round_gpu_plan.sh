#!/bin/bash
# Measurement plan for the next round's first GPU calls (this round's budget ended before the
# device-side monitor / far-field / geometry kernels could be profiled).  Each block is one gpurun
# call; everything lands in gpurun_out/.
#
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/next_round_gpu_plan.sh tests'
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash profiles/next_round_gpu_plan.sh time'
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/next_round_gpu_plan.sh ncu'
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/next_round_gpu_plan.sh sanitize'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
case "$1" in
tests)      # whole GPU suite (incl. the phasor-guard test that calls CUDA from a signal handler) + smoke + bench
	timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
	python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
	python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
	tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_n1.json ;;
time)       # host-call timings of the monitor / far-field / geometry entry points + putObjects + examples
	python profiles/monitors_probe.py 512 > gpurun_out/monitors_probe_512.json 2>&1
	bash profiles/geometry_r01.sh
	python -m pytest tests/test_gpu_examples.py -m gpu -q -s 2>&1 | grep -E "EXAMPLE_TIMING|passed|failed" > gpurun_out/examples.log
	# Dipole_Band's 6400-step loop: default phasor path against one thread per (cell, frequency) and/or a ring of
	# 8 pinned staging slots (default: each updatePhaser waits for the previous step's copy, host and device alternate)
	export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH; BIN=$PWD/tests/bin/dipole_band_cuda; LOG=$PWD/gpurun_out/examples.log
	for par in 0 1; do for ring in 1 8; do
		d=/tmp/db_${par}_${ring}; mkdir -p $d
		( cd $d; cp $BIN .; s=$(date +%s.%N); ALISB_PHASER_PAR=$par ALISB_PHASER_RING=$ring ALIS_FF_TIMING=1 ./dipole_band_cuda > log.txt 2>&1
		  echo "dipole_band ALISB_PHASER_PAR=$par ALISB_PHASER_RING=$ring $(python -c "import time; print(round(time.time()-$s,2))") s" >> $LOG )
	done; done
	for f in /tmp/db_1_8/dipole_band_cuda.out/*.txt; do cmp -s $f /tmp/db_0_1/dipole_band_cuda.out/$(basename $f) || echo "DIFFERS $(basename $f)" >> $LOG; done
	tail -5 gpurun_out/monitors_probe_512.json gpurun_out/examples.log ;;
ncu)        # one full capture per new kernel (first launch each), DRAM bytes and durations
	ncu --set full --clock-control none --import-source on --launch-count 1 \
		-k regex:"eval_box|reduce_absmax|column_sums|absorption_columns|surface_sums|geometry_fill|geometry_bounds|phaser_face" \
		-o gpurun_out/monitors_full python profiles/monitors_probe.py 256 > gpurun_out/ncu_monitors.log 2>&1
	ncu -i gpurun_out/monitors_full.ncu-rep --page raw --csv \
		--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread \
		> gpurun_out/monitors_full_raw.csv 2>&1
	tail -12 gpurun_out/monitors_full_raw.csv ;;
sanitize)   # memcheck + racecheck of every kernel on small grids (the reference has no sanitizer runs)
	for tool in memcheck racecheck; do
		timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest \
			"tests/test_gpu_parity.py::test_box_world_seeded_bit_exact" tests/test_monitor_eval.py tests/test_farfield_sums.py \
			"tests/test_zz_geometry.py::test_cuda_rasteriser_bit_exact" -m gpu -q -x 2>&1 | tail -15 > gpurun_out/sanitize_$tool.log
		tail -4 gpurun_out/sanitize_$tool.log
	done ;;
*) echo "usage: $0 tests|time|ncu|sanitize" ;;
esac
