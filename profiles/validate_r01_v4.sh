set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_v4.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_n1_v4.json 2> gpurun_out/bench_n1_v4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v4.json 2>> gpurun_out/bench_n1_v4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_v4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_v4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:yee_volume -s 40 -c 24 -o gpurun_out/prof_r01_v4 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_v4.log 2>&1
cat gpurun_out/pytest_gpu_v4.log; cat gpurun_out/bench_n1_v4.json
