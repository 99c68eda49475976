#!/bin/sh
# Builds the CUDA engine IN-TREE for sm_100a: alis_b200/libalis_b200_engine.so
# -fmad=false: no FMA contraction, required for bit parity with the reference.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
	-Xcompiler -fPIC,-fopenmp,-O2 -shared -I"$HERE/../include" -I"$HERE/csrc" \
	-o "$HERE/libalis_b200_engine.so" "$HERE/csrc/engine.cu" -lgomp "$@"
echo "build_engine: ok -> $HERE/libalis_b200_engine.so"
