/* Diffraction from a periodic dielectric metasurface on a substrate: oblique TFSF plane wave,
 * Bloch-periodic in x, mirror in y, CPML in z; one phaser (running DFT at 500 nm) on the z faces,
 * near-to-far-field transform with 4x4 periodic replicas.  Same physics as the reference's
 * examples/Metasurface_Farfield.c.
 * Output: <ID>/Far-field-theta.txt and <ID>/Log-*.png far-field maps. */
#include <alis.h>

static const float PERIOD = 500;          /* nm, unit cell in x and y */

int main(int argc, char **argv)
{
	dom unit_cell = {{PERIOD}, {PERIOD}, {500}};
	dom illuminated = {{PERIOD}, {PERIOD}, {300}};                 /* total-field region of the plane wave */
	res grid = {10, {20}, {1240}};
	sur walls = {{BBC}, {SYM}, {PML}};
	world sim = createWorld(unit_cell, grid, walls, "%s.out", argv[0]);

	object pillar = {Box, {{0, 200}, {-100, 100}, {-100, 100}}};
	object pillars = {Lattice, {{PERIOD}, {PERIOD}}, &pillar};
	object substrate = {Box, {{-INF, INF}, {-INF, INF}, {-INF,-100}}};
	putObjects(sim, n(4), pillars, n(1.5), substrate, Air);
	planewave(sim, illuminated, Ppol, 55, 0, Sine, 530, 5);

	/* the monitor box is four periods wide: the far field is that of 4 x 4 replicas */
	phaser dft = createPhaser(sim, 500, -2000, 2000, -2000, 2000);
	const long periods = 10, period = sim->T;
	for (long step = 1; timer(step, periods*period); step++) {
		updateE(sim);
		updateH(sim);
		if (periods*period - step < period) updatePhaser(sim, dft);   /* the last period only */
	}
	farFieldTheta(sim, dft, 0, "/Far-field-theta");
	farFieldProfile(sim, dft, AzimuthalMap, 360, 180, png(jet,0), "/Log-");
	return 0;
}
