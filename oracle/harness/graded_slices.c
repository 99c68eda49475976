/* Test driver (test infrastructure): slices of a graded (non-uniform) grid are resampled onto the
 * uniform raster before they are written (reference slices.c:3-36, host/slices.c resample_along).
 * Writes 1-D and 2-D text slices and a 3-D HDF5 slice of a dipole field on a world graded along all three axes;
 * tests/test_reference_examples.py builds it against the reference and against the host layer and
 * compares the files byte for byte. */
#include <alis.h>

int main(void)
{
	dom Dom = {{600}, {500}, {400}};
	res Res = {5, {40, 30, 25}, {0}};
	Res.nug.x.m = 3, Res.nug.x.Min = 120, Res.nug.x.Max = 120;
	Res.nug.y.m = 2, Res.nug.y.Min = 100, Res.nug.y.Max = 100;
	Res.nug.z.m = 2, Res.nug.z.Min = 80, Res.nug.z.Max = 80;
	sur Sur = {{PML}, {PML}, {PML}};
	world W = createWorld(Dom, Res, Sur, "graded_slices.out");
	object Block = {Box, {{-150, 90}, {-90, 70}, {-70, 40}}};
	putObjects(W, n(1.8), Block, Air);
	pointDipole(W, Ex, 35, -20, 15, Pulse, 500, 150);

	slice XY = createSliceXY(W, 10), XZ = createSliceXZ(W, -30), YZ = createSliceYZ(W, 50);
	slice Line = createSlice(W, -INF, INF, 20, 20, 10, 10), Vol = createSlice(W, -200, 200, -150, 150, -100, 100);
	for (int n=1; n<=120; n++) {
		updateH(W);
		updateE(W);
		if (n == 60 || n == 120) {
			sliceSnap(W, Ex, XY, 0, txt, "/xy-Ex-%d", n);
			sliceSnap(W, Hy, XZ, 0, txt, "/xz-Hy-%d", n);
			sliceSnap(W, EE, YZ, 0, txt, "/yz-EE-%d", n);
			sliceSnap(W, Ez, Line, 0, txt, "/line-Ez-%d", n);
			sliceSnap(W, LogEE, Vol, 0, h5, "/vol-LogEE-%d", n);
			sliceSnap(W, RI, XZ, txt, "/xz-RI-%d", n);          /* structure functions take no frame count */
		}
	}
	return 0;
}
