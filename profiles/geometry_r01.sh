#!/bin/bash
# putObjects (geometry rasterisation, row f3) on a perforated-slab structure at 256^3 cells, default
# sub-pixel sampling (subN = 12): unmodified reference on the host cores, this library's host
# rasteriser, and the device rasteriser (ALIS_DEVICE_GEOMETRY=1).  -> gpurun_out/v9_putobjects.txt
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/alis_b200:$LD_LIBRARY_PATH
{
echo "host_threads $(nproc)"
for n in 256; do
	echo "reference $(timeout 60 oracle/_ref/bin/putobjects_bench_ref $n 2>&1 | tail -1)"
	echo "host_path $(timeout 60 tests/bin/putobjects_bench_mine_cuda $n 2>&1 | tail -1)"
	echo "device_path $(ALIS_DEVICE_GEOMETRY=1 timeout 60 tests/bin/putobjects_bench_mine_cuda $n 2>&1 | tail -1)"
done
} | tee gpurun_out/v9_putobjects.txt
