#!/bin/bash
# Round 2: strong-scaling points of the 2048^3 cube on N GPUs (bench.py --scaling strong), overlap on/off, plus the weak point.
# usage: bash profiles/r02_gpu_strong.sh N
N=$1
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
run() {  # name, env..., then bench args in BARGS
	name=$1; shift
	if [ $N = 1 ]; then ( env "$@" timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline $BARGS ) > $O/$name.json 2> $O/$name.err
	else ( env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline $BARGS ) > $O/$name.json 2> $O/$name.err; fi
	python - $O/$name.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1].split('/')[-1], 'value %.0f per-GPU %.0f ms/step %.2f e2e %s fused %s' % (d['value'], d['per_gpu_mcells_per_s'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d.get('fused_launches')))
PY
}
BARGS="--scaling strong" run r02_bench_strong_n${N} A=1
if [ $N -gt 1 ]; then BARGS="--scaling strong" run r02_bench_strong_n${N}_nooverlap ALISB_NO_OVERLAP=1; fi
if [ $N -le 2 ]; then BARGS="--scaling strong" run r02_bench_strong_n${N}_inplace ALISB_WAVE=1; fi
# the same cube through the public C API in ONE process (ALIS_DEVICES: host/shard.c), updateE / updateH / get() per step
if [ $N -ge 4 ]; then
	B=tests/bin/bench_cube_mine_cuda
	DEV=$(seq -s, 0 $((N-1)))
	for nz in $((256*N)); do
		echo "ALIS_DEVICES=$DEV nz=$nz $(ALIS_DEVICES=$DEV timeout 900 $B 2048 2048 $nz 10 3 1 | tail -1)" >> $O/r02_capi_shards_n${N}.txt
	done
	cat $O/r02_capi_shards_n${N}.txt
fi
