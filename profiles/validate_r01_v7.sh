set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_v7.log; cat gpurun_out/pytest_gpu_v7.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_n1_v7.json 2> gpurun_out/bench_n1_v7.err; cat gpurun_out/bench_n1_v7.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v7.json 2>> gpurun_out/bench_n1_v7.err
timeout 600 python bench.py --steps 30 --warmup 3 --fuse 0 --no-cpu-baseline > gpurun_out/bench_n1_v7_twopass.json 2>> gpurun_out/bench_n1_v7.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01_v7.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_v7.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:yee_fused -s 2 -c 2 -o gpurun_out/prof_r01_v7 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_v7.log 2>&1
timeout 300 python profiles/small_grid_probe.py 256 512 1024 2>&1 | tail -3 > gpurun_out/cube_sweep_v7.txt; cat gpurun_out/cube_sweep_v7.txt
