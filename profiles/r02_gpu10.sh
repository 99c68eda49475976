#!/bin/bash
# Round 2, GPU calls 10, 11: worlds cut into z slabs behind the public C API (tests/test_shards.py) with the CUDA engine.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=$PWD/gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/r02_v10_gpu.txt
( time timeout 1500 python -m pytest tests/test_shards.py tests/test_gpu_multi.py -m gpu -q -x -rs ) > $O/r02_pytest_shards_v11.log 2>&1
tail -8 $O/r02_pytest_shards_v11.log
B=tests/bin/bench_cube_mine_cuda
NG=$(nvidia-smi -L | wc -l)
rm -f $O/r02_capi_shards_v11.txt
for dev in "" "0,0" "0,1" "0,1,2,3"; do
	nd=$(echo "$dev" | tr ',' '\n' | sort -u | grep -c .)
	[ "$nd" -gt "$NG" ] && continue
	for fuse in 1 0; do
		echo "ALIS_DEVICES=[$dev] ALISB_FUSE=$fuse $(ALIS_DEVICES=$dev ALISB_FUSE=$fuse timeout 600 $B 2048 2048 512 10 3 1 | tail -1)" >> $O/r02_capi_shards_v11.txt
	done
done
cat $O/r02_capi_shards_v11.txt
