timeout 600 ncu --set full --clock-control none --import-source on -k regex:yee_fused -s 2 -c 2 -o gpurun_out/prof_r01_fused_v2 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --fuse 1 > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01_v6.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_v6.log 2>&1
