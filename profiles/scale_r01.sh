# usage: bash profiles/scale_r01.sh N   (run under gpurun --gpus N)
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n${N}_v6.json 2> gpurun_out/bench_n${N}_v6.err
cat gpurun_out/bench_n${N}_v6.json
ALISB_NO_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n${N}_v6_nooverlap.json 2>> gpurun_out/bench_n${N}_v6.err
cat gpurun_out/bench_n${N}_v6_nooverlap.json
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
