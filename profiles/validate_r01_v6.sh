timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_v6.log; cat gpurun_out/pytest_gpu_v6.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_n1_v6.json 2> gpurun_out/bench_n1_v6.err; cat gpurun_out/bench_n1_v6.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_n2_v6.json 2> gpurun_out/bench_n2_v6.err; cat gpurun_out/bench_n2_v6.json | cut -c1-300
