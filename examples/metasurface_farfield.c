/* Diffraction from a periodic dielectric metasurface on a substrate: oblique TFSF plane wave,
 * Bloch-periodic in x, mirror in y, CPML in z; one phaser (running DFT at 500 nm) on the z faces,
 * near-to-far-field transform with 4x4 periodic replicas.  Same physics as the reference's
 * examples/Metasurface_Farfield.c.
 * Output: <ID>/Far-field-theta.txt and <ID>/Log-*.png far-field maps. */
#include <alis.h>

int main(int argc, char **argv)
{
	dom DTF = {{500}, {500}, {300}};
	dom DSF = {{500}, {500}, {500}};
	res Res = {10, {20}, {1240}};
	sur Sur = {{BBC}, {SYM}, {PML}};
	world W = createWorld(DSF, Res, Sur, "%s.out", argv[0]);

	object Nanoblock = {Lattice, {{500}, {500}}, objects {{Box, {{0, 200}, {-100, 100}, {-100, 100}}}}};
	object Substrate = {Box, {{-INF, INF}, {-INF, INF}, {-INF,-100}}};
	putObjects(W, n(4), Nanoblock, n(1.5), Substrate, Air);
	planewave(W, DTF, Ppol, 55, 0, Sine, 530, 5);

	phaser P = createPhaser(W, 500,-2000, 2000,-2000, 2000);
	for (int n=1, N=10*W->T; timer(n, N); n++) {
		updateE(W);
		updateH(W);
		if (N-n < W->T) updatePhaser(W, P);
	}
	farFieldTheta(W, P, 0, "/Far-field-theta");
	farFieldProfile(W, P, AzimuthalMap, 360, 180, png(jet,0), "/Log-");
	return 0;
}
