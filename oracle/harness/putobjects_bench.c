/*
 * putobjects_bench.c -- times putObjects (geometry rasterisation, reference object.c:339-405) on a
 * photonic-crystal-like structure: a dielectric slab perforated by a triangular lattice of holes
 * on a substrate, plus a dielectric sphere, default sub-pixel sampling.  The same source links against
 * the unmodified reference, and against this library (host rasteriser, or the device rasteriser
 * with ALIS_DEVICE_GEOMETRY=1).  Prints one JSON line; a checksum of the inverse-permittivity
 * arrays (through EE of a unit field is not available before stepping, so RI slices are summed)
 * lets the runs be compared.
 *
 *   putobjects_bench <cells per axis> [subN]
 */
#include <alis.h>
#include <sys/time.h>

static double now(void) { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6*tv.tv_usec; }

int main(int argc, char **argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 160;
	const int subN = argc > 2 ? atoi(argv[2]) : 0;
	const float half = 5.0f*n;                                    /* dx = 10 nm */
	dom Dom = {{-half, half}, {-half, half}, {-half, half}};
	res Res = {5, {10}, {1240, subN}};
	sur Sur = {{PML, PML}, {PML, PML}, {PML, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	object Slab = {Difference, {2}, objects {
		{Box, {{-INF, INF}, {-INF, INF}, {-110, 110}}},
		{DLattice, {{420}, {420*sqrtf(3)}}, objects {{RodZ, {{-126, 126}, {-126, 126}, {-INF, INF}}}}}}};
	object Substrate = {Box, {{-INF, INF}, {-INF, INF}, {-INF, -300}}};
	object Sphere = {Ball, {{-90, 90}, {-90, 90}, {200, 380}}};
	const double t0 = now();
	putObjects(W, n(2.0), Sphere, n(3.4), Slab, n(1.45), Substrate, Air);
	const double t1 = now();
	slice XY = createSliceXY(W, 0), XZ = createSliceXZ(W, 0), YZ = createSliceYZ(W, 30);
	const double chk = (double)sliceSum(W, RI, XY) + (double)sliceSum(W, RI, XZ) + (double)sliceSum(W, RI, YZ) + (double)sliceSum(W, eEff, XZ);
	printf("\n{\"cells\": %ld, \"subN\": %d, \"putObjects_seconds\": %.3f, \"checksum\": %.9g}\n", (long)W->iNUM*W->jNUM*W->kNUM, W->subN, t1-t0, chk);
	return 0;
}
