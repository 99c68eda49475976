timeout 600 ncu --set full --clock-control none --import-source on -k regex:yee_fused -s 2 -c 2 -o gpurun_out/prof_r01_fused -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --fuse 1 > gpurun_out/ncu_fused.log 2>&1
tail -3 gpurun_out/ncu_fused.log
for fc in 8 32 64; do ALISB_FCHUNK=$fc timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --fuse 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fchunk $fc', d['ms_per_step'], d['roofline']['avg_launch_ms'])"; done
