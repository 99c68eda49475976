#!/bin/bash
# Round 1, last GPU call (6 GPU-minutes were left): the device-side monitors (rows f1, f2) first,
# then as much of the full GPU suite as the remaining time allows.  Logs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/v8_gpu.txt 2>&1
timeout 150 python -m pytest tests/test_monitor_eval.py tests/test_farfield_sums.py \
	"tests/test_gpu_dropin.py::test_dropin_monitor_values" "tests/test_gpu_dropin.py::test_device_monitors_equal_host_view_monitors" \
	-m gpu -q -rA 2>&1 | tail -60 > gpurun_out/v8_pytest_new.log
timeout 120 python -m pytest tests/test_gpu_examples.py -m gpu -q -s -k "metasurface or dipole_band or nanoblock" 2>&1 | grep -E "EXAMPLE_TIMING|passed|failed|Error|error" > gpurun_out/v8_examples.log
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/v8_pytest_all.log
tail -5 gpurun_out/v8_pytest_new.log gpurun_out/v8_examples.log gpurun_out/v8_pytest_all.log
