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
	# Dipole_Band's 6400-step loop: default phaser kernel against one thread per (cell, frequency)
	mkdir -p /tmp/db0 /tmp/db1; export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
	for par in 0 1; do ( cd /tmp/db$par; cp $OLDPWD/tests/bin/dipole_band_cuda .; s=$(date +%s.%N); ALISB_PHASER_PAR=$par ALIS_FF_TIMING=1 ./dipole_band_cuda > log.txt 2>&1; echo "dipole_band ALISB_PHASER_PAR=$par $(python -c "import time; print(round(time.time()-$s,2))") s" ) >> gpurun_out/examples.log; done
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
