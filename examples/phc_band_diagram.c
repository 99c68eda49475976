/* One k-point of the band diagram of a triangular-lattice photonic-crystal slab: Bloch boundary
 * in x, hexagonal Bloch boundary in y (complex fields), Band dipole excitation, spectrum of a probe.
 * Same physics as the reference's examples/PhC_Band_Diagram.c.
 *   phc_band_diagram [k along Gamma-M-K-Gamma, 0..3, default 1.2] [steps after the source has ended]
 * Output: <ID>.txt (wavelength, |DFT|, phase/pi). */
#include <alis.h>

static const float PITCH = 500;           /* lattice constant, nm */

/* position k on the path Gamma (0) - M (1) - K (2) - Gamma (3) -> reduced wave vector */
static void path_point(float k, float *kx, float *ky)
{
	*kx = *ky = 0;
	if (k <= 0 || k >= 3) return;
	if (k < 1)      { *ky = k; }
	else if (k < 2) { *kx = k-1; *ky = 1; }
	else            { *kx = 3-k; *ky = 3-k; }
}

int main(int argc, char **argv)
{
	const float k = argc > 1 ? atof(argv[1]) : 1.2;
	float kx, ky;
	path_point(k, &kx, &ky);

	const float row = PITCH/2*sqrtf(3);                            /* distance between hole rows */
	dom cell = {{PITCH/2}, {0, row}, {500}};
	res grid = {5, {10, 5*sqrtf(3), 10}, {1240}};
	sur walls = {{BBC, PITCH*3/kx}, {HBC, PITCH*sqrtf(3)/ky}, {SYM, PML}};
	world sim = createWorld(cell, grid, walls, "%s.out-%03.1f", argv[0], k);

	object hole = {RodZ, {{-200, 200}, {-200, 200}, {-INF, INF}}};
	object holes = {DLattice, {{PITCH}, {PITCH*sqrtf(3)}}, &hole};
	object parts[2] = {{Box, {{-INF, INF}, {-INF, INF}, {-100, 100}}}, holes};
	object slab = {Difference, {2}, parts};
	putObjects(sim, n(3.5), slab, Air);
	pointDipole(sim, iHz, 200, 100*sqrtf(3), 0, Band, 750, 2000);

	const long listen = argc > 2 ? atoi(argv[2]) : 750*2000/5/sim->dt;
	for (long step = 1; timer(step, sim->N + listen); step++) {
		updateE(sim);
		updateH(sim);
		if (step <= (long)sim->N) continue;                        /* the source is still on */
		writeSpectrum(sim, listen, 750, 2000, "", get(sim, Hz, 200, 0, 0));
	}
	return 0;
}
