/* One k-point of the band diagram of a triangular-lattice photonic-crystal slab: Bloch boundary
 * in x, hexagonal Bloch boundary in y (complex fields), Band dipole excitation, spectrum of a probe.
 * Same physics as the reference's examples/PhC_Band_Diagram.c.
 * Output: <ID>.txt (wavelength, |DFT|, phase/pi). */
#include <alis.h>

int main(int argc, char **argv)
{
	float kx, ky, k = argc > 1 ? atof(argv[1]) : 1.2;
	if (k<=0 || k>=3) kx = 0, ky = 0;
	else if (k < 1) kx =   0, ky =   k;   /* Gamma - M */
	else if (k < 2) kx = k-1, ky =   1;   /* M - K     */
	else            kx = 3-k, ky = 3-k;   /* K - Gamma */

	dom Dom = {{250}, {0, 250*sqrtf(3)}, {500}};
	res Res = {5, {10, 5*sqrtf(3), 10}, {1240}};
	sur Sur = {{BBC, 500*3/kx}, {HBC, 500*sqrtf(3)/ky}, {SYM, PML}};
	world W = createWorld(Dom, Res, Sur, "%s.out-%03.1f", argv[0], k);
	object Slab = {Difference, {2}, objects {
						{Box, {{-INF, INF}, {-INF, INF}, {-100, 100}}},
						{DLattice, {{500}, {500*sqrtf(3)}}, objects {
							{RodZ, {{-200, 200}, {-200, 200}, {-INF, INF}}},
						}},
					}};
	putObjects(W, n(3.5), Slab, Air);
	pointDipole(W, iHz, 200, 100*sqrtf(3), 0, Band, 750, 2000);

	int steps = argc > 2 ? atoi(argv[2]) : 750*2000/5/W->dt;
	for (int n=1, N=steps; timer(n, W->N+N); n++) {
		updateE(W);
		updateH(W);
		if (n > W->N) writeSpectrum(W, N, 750, 2000, "", get(W, Hz, 200, 0, 0));
	}
	return 0;
}
