timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused" 2>&1 | tail -3
timeout 280 python profiles/eps_probe.py 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('vacuum', d['ms_per_step'], d['value'], d['roofline']['avg_launch_ms'])"
