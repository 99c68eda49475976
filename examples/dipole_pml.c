/* Point dipole radiating in vacuum inside a CPML box (one octant, mirror symmetry on the three
 * coordinate planes).  Same physics as the reference's examples/Dipole_PML.c; written against the
 * public API in include/alis.h and buildable against either library (see examples/README.md).
 * Output: <ID>/Ez-*.png snapshots (real PNGs with this library) and <ID>/probe.txt. */
#include <alis.h>

int main(int argc, char **argv)
{
	dom Dom = {{1000}, {1000}, {1000}};
	res Res = {10, {20}, {1240}};
	sur Sur = {{SYM, PML}, {SYM, PML}, {SYM, PML}, {24}};
	world W = createWorld(Dom, Res, Sur, "%s.out", argv[0]);

	pointDipole(W, Ez, 0, 0, 0, Pulse, 400, 800, 10000000);

	slice XY = createSlice(W, W->xMIN, W->xMAX, W->yMIN, W->yMAX, 0, 0);
	slice XZ = createSlice(W, W->xMIN, W->xMAX, 0, 0, W->zMIN, W->zMAX);

	for (int n=1, N=3000/W->dt; timer(n, N); n++) {
		updateE(W);
		updateH(W);
		sliceSnap(W, Ez, XY, 10, png(dkbr,1), "/%%/");
		sliceSnap(W, Ez, XZ, 10, png(dkbr,1), "/%%/");
		writeRow(W, "/probe", W->t, get(W, Ez, 300, 0, 0), get(W, Hy, 300, 0, 0), get(W, Ez, 0, 0, 500));
	}
	return 0;
}
