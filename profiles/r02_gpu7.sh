#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): NCCL parity tests; strong-scaling point N=2 of the 2048^3 cube (1024 planes per GPU:
# two-pass, a second set of field arrays does not fit) and the same with in-place fused stepping (yee_wave); weak N=2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/r02_v7_gpu.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -rs ) > $O/r02_pytest_multi_n2_v7.log 2>&1
tail -4 $O/r02_pytest_multi_n2_v7.log
run() {  # name, env..., then bench args
	name=$1; shift
	( env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline $BARGS ) > $O/$name.json 2> $O/$name.err
	python - $O/$name.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1].split('/')[-1], 'value %.0f per-GPU %.0f ms/step %.2f e2e %s fused %s' % (d['value'], d['per_gpu_mcells_per_s'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d.get('fused_launches')))
PY
}
BARGS="--scaling strong" run r02_bench_strong_n2_v7 A=1
BARGS="--scaling strong" run r02_bench_strong_n2_wave_v7 ALISB_WAVE=1
BARGS="--scaling strong" run r02_bench_strong_n2_nooverlap_v7 ALISB_NO_OVERLAP=1
BARGS="" run r02_bench_weak_n2_v7 A=1
