timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python profiles/small_grid_probe.py 2>&1 | tail -6
ALISB_MINCHUNK=4 timeout 120 python profiles/small_grid_probe.py 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-200
