#!/bin/sh
# Builds the UNMODIFIED reference (hseelab/alis, /root/reference/src/*.c) into
# oracle/_ref/ with the reference's own flags (reference src/Makefile:2-4:
# gcc -fopenmp -O3 -std=c99 -fpic), using two shim headers for the libraries that
# are absent from this image (hdf5.h -> this repo's minimal HDF5 reader/writer, png.h -> no-ops).  Test infrastructure only: the
# products (libalis_ref.so/.a, abench, harness binaries) are used as the parity
# checker and as the CPU baseline; nothing under alis_b200/ links or loads them.
# Sources are compiled where they lie; nothing is copied into the repo.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${ALIS_REFERENCE:-/root/reference}
OUT=$HERE/_ref
if [ ! -d "$REF/src" ]; then
	echo "build_ref: $REF/src not present (GPU box?) - keeping prebuilt $OUT" >&2
	exit 0
fi
mkdir -p "$OUT/obj" "$OUT/bin"
CFLAGS="-fopenmp -O3 -std=c99 -fpic -w -I$HERE/shim -I$REF/src"
for f in world object shapes source update output slices fields phaser farfld; do
	if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/src/$f.c" -nt "$OUT/obj/$f.o" ]; then
		gcc $CFLAGS -c "$REF/src/$f.c" -o "$OUT/obj/$f.o"
	fi
done
# the hdf5.h shim maps the reference's H5* calls onto this repo's minimal HDF5 reader/writer
gcc -std=c99 -O2 -fpic -w -c "$HERE/../alis_b200/host/h5lite.c" -o "$OUT/obj/h5lite.o"
# world.o must pick up a changed shim
if [ "$HERE/shim/hdf5.h" -nt "$OUT/obj/world.o" ]; then gcc $CFLAGS -c "$REF/src/world.c" -o "$OUT/obj/world.o"; fi
# ... and slices.o a changed png.h shim (a minimal real PNG writer, so that slice images can be compared)
if [ "$HERE/shim/png.h" -nt "$OUT/obj/slices.o" ]; then gcc $CFLAGS -c "$REF/src/slices.c" -o "$OUT/obj/slices.o"; fi
rm -f "$OUT/libalis_ref.a"
ar rcs "$OUT/libalis_ref.a" "$OUT"/obj/*.o
gcc -shared -fopenmp -o "$OUT/libalis_ref.so" "$OUT"/obj/*.o -lm
gcc -std=c99 -fopenmp -O3 -w -o "$OUT/bin/abench" "$REF/bin/abench.c"
# harness drivers (ours, oracle/harness/*.c) linked against the reference library
HOSTDIR=$HERE/../alis_b200/host
INC="-I$HERE/shim -I$REF/src -I$HERE/harness -I$HERE/../include -I$HOSTDIR"
# (1) unmodified reference does everything
gcc -std=c99 -fopenmp -O3 -w -DHARNESS_MODE_REF $INC "$HERE/harness/stepdump.c" "$OUT/libalis_ref.a" -lm -o "$OUT/bin/stepdump_ref"
gcc -std=c99 -fopenmp -O3 -w $INC "$HERE/harness/bench_cube.c" "$OUT/libalis_ref.a" -lm -o "$OUT/bin/bench_cube_ref"
gcc -std=c99 -fopenmp -O3 -w $INC "$HERE/harness/monitors.c" "$OUT/libalis_ref.a" -lm -o "$OUT/bin/monitors_ref"
gcc -std=c99 -fopenmp -O3 -w $INC "$HERE/harness/api_misc.c" "$OUT/libalis_ref.a" -lm -o "$OUT/bin/api_misc_ref"
gcc -std=c99 -fopenmp -O3 -w $INC "$HERE/harness/putobjects_bench.c" "$OUT/libalis_ref.a" -lm -o "$OUT/bin/putobjects_bench_ref"
gcc -std=c99 -fopenmp -O3 -w $INC "$HERE/harness/graded_slices.c" "$OUT/libalis_ref.a" -lm -o "$OUT/bin/graded_slices_ref"
# the example drivers (ours, examples/*.c) linked against the reference: the CPU twins used for
# output and wall-time comparison on the GPU box
for ex in "$HERE"/../examples/*.c; do
	[ -f "$ex" ] || continue
	gcc -std=c99 -fopenmp -O3 -w $INC "$ex" "$OUT/libalis_ref.a" -lm -o "$OUT/bin/$(basename "$ex" .c)_ref"
done
# the reference's OWN example drivers, unmodified, compiled where they lie (nothing is copied): the CPU twins
# of tests/bin/refex_*_cuda (tests/test_gpu_refex.py)
for ex in "$REF"/examples/*.c; do
	[ -f "$ex" ] || continue
	gcc -std=c99 -fopenmp -O3 -w -I"$HERE/shim" -I"$REF/src" "$ex" "$OUT/libalis_ref.a" -lm -o "$OUT/bin/refex_$(basename "$ex" .c)_ref"
done
# (2) reference setup + CPU oracle engine through the bridge
if [ -f "$HERE/liboracle_alis.so" ]; then
	gcc -std=c99 -fopenmp -O3 -w -DHARNESS_MODE_BRIDGE $INC "$HERE/harness/stepdump.c" "$HOSTDIR/engine_bridge.c" "$HOSTDIR/shard.c" \
		"$OUT/libalis_ref.a" -L"$HERE" -loracle_alis -Wl,-rpath,'$ORIGIN/../..' -lm -o "$OUT/bin/stepdump_bridge_oracle"
fi
# (3) reference setup + CUDA engine through the bridge (runs on the GPU box only)
ENGINE_DIR=$HERE/../alis_b200
if [ -f "$ENGINE_DIR/libalis_b200_engine.so" ]; then
	gcc -std=c99 -fopenmp -O3 -w -DHARNESS_MODE_BRIDGE $INC "$HERE/harness/stepdump.c" "$HOSTDIR/engine_bridge.c" "$HOSTDIR/shard.c" \
		"$OUT/libalis_ref.a" -L"$ENGINE_DIR" -lalis_b200_engine -Wl,-rpath,'$ORIGIN/../../../alis_b200' -lm -o "$OUT/bin/stepdump_bridge_cuda"
fi
echo "build_ref: ok -> $OUT"
