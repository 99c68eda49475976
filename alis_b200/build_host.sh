#!/bin/sh
# Builds the host C layer (the ALiS-compatible public API) IN-TREE:
#   alis_b200/libalis.so        host layer + CUDA engine   (the drop-in product library)
#   tests/bin/libalis_oracle.so host layer + CPU oracle    (TEST INFRASTRUCTURE: lets the host layer be
#                                                            validated against the reference without a GPU)
#   tests/bin/stepdump_mine_*   harness drivers linked against either
# Same flags as the reference build (gcc -O3 -std=c99, no fast-math) so float/complex expressions
# compile to the same arithmetic.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
CC=${ALIS_HOST_CC:-gcc}
OBJ=$HERE/host/obj
mkdir -p "$OBJ" "$ROOT/tests/bin"
CFLAGS="-std=c99 -O3 -fopenmp -fPIC -w -I$ROOT/include -I$HERE/host"
SRCS="world object shapes source update view fields output slices phaser farfld engine_bridge shard h5lite"
OBJS=""
for f in $SRCS; do
	$CC $CFLAGS -c "$HERE/host/$f.c" -o "$OBJ/$f.o"
	OBJS="$OBJS $OBJ/$f.o"
done
if [ -f "$HERE/libalis_b200_engine.so" ]; then
	$CC -shared -fopenmp -o "$HERE/libalis.so" $OBJS -L"$HERE" -lalis_b200_engine -Wl,-rpath,'$ORIGIN' -lm
fi
if [ -f "$ROOT/oracle/liboracle_alis.so" ]; then
	$CC -shared -fopenmp -o "$ROOT/tests/bin/libalis_oracle.so" $OBJS -L"$ROOT/oracle" -loracle_alis -Wl,-rpath,'$ORIGIN/../../oracle' -lm
fi
HARNESS=$ROOT/oracle/harness
for h in stepdump bench_cube monitors putobjects_bench api_misc graded_slices; do
	[ -f "$HARNESS/$h.c" ] || continue
	if [ -f "$HERE/libalis.so" ]; then
		$CC -std=c99 -O2 -fopenmp -w -DHARNESS_MODE_MINE -I"$ROOT/include" -I"$HARNESS" "$HARNESS/$h.c" -L"$HERE" -lalis -Wl,-rpath,'$ORIGIN/../../alis_b200' -lm -o "$ROOT/tests/bin/${h}_mine_cuda"
	fi
	if [ -f "$ROOT/tests/bin/libalis_oracle.so" ]; then
		$CC -std=c99 -O2 -fopenmp -w -DHARNESS_MODE_MINE -I"$ROOT/include" -I"$HARNESS" "$HARNESS/$h.c" -L"$ROOT/tests/bin" -lalis_oracle -Wl,-rpath,'$ORIGIN' -lm -o "$ROOT/tests/bin/${h}_mine_oracle"
	fi
done
for ex in "$ROOT"/examples/*.c; do
	[ -f "$ex" ] || continue
	b=$(basename "$ex" .c)
	if [ -f "$HERE/libalis.so" ]; then
		$CC -std=c99 -O2 -fopenmp -w -I"$ROOT/include" "$ex" -L"$HERE" -lalis -Wl,-rpath,'$ORIGIN/../../alis_b200' -lm -o "$ROOT/tests/bin/${b}_cuda"
	fi
	if [ -f "$ROOT/tests/bin/libalis_oracle.so" ]; then
		$CC -std=c99 -O2 -fopenmp -w -I"$ROOT/include" "$ex" -L"$ROOT/tests/bin" -lalis_oracle -Wl,-rpath,'$ORIGIN' -lm -o "$ROOT/tests/bin/${b}_oracle"
	fi
done
# the reference's OWN example drivers (reference examples/*.c, all 16, unmodified), compiled where they lie
# against include/alis.h and linked with the drop-in library: tests/bin/refex_<Name>_cuda run on the GPU box
# next to oracle/_ref/bin/refex_<Name>_ref (tests/test_gpu_refex.py).  Only where the reference tree is mounted;
# the binaries travel to the GPU box, no source is copied.
REF=${ALIS_REFERENCE:-/root/reference}
for ex in "$REF"/examples/*.c; do
	[ -f "$ex" ] || continue
	b=$(basename "$ex" .c)
	if [ -f "$HERE/libalis.so" ]; then
		$CC -std=c99 -O2 -fopenmp -w -I"$ROOT/include" "$ex" -L"$HERE" -lalis -Wl,-rpath,'$ORIGIN/../../alis_b200' -lm -o "$ROOT/tests/bin/refex_${b}_cuda"
	fi
	if [ -f "$ROOT/tests/bin/libalis_oracle.so" ]; then
		$CC -std=c99 -O2 -fopenmp -w -I"$ROOT/include" "$ex" -L"$ROOT/tests/bin" -lalis_oracle -Wl,-rpath,'$ORIGIN' -lm -o "$ROOT/tests/bin/refex_${b}_oracle"
	fi
done
# CPU check of the symbolic walk behind kernel copy_resolved (tests/test_copy_resolve.py)
${CXX:-g++} -O1 -std=c++17 -ffp-contract=off -o "$ROOT/tests/bin/rc_check" "$ROOT/tests/native/rc_check.cpp"
echo "build_host: ok"
