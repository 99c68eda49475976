/* Broadband dipole in vacuum on a graded (non-uniform) grid with mirror symmetry: 81 phasers
 * (400-800 nm) on the box faces every step, then radiated power, far-field spectrum and pattern.
 * The three printed numbers are 1 by construction (unit-power dipole, phasers normalised by the
 * source spectrum).  Same physics as the reference's examples/Dipole_Band.c.
 * Output: <ID>/Power.txt, Far-field.txt, Far-field-theta.txt, Near-field.txt. */
#include <alis.h>

/* Ex on the z axis at height z, scaled so that a unit-power dipole gives 1 in the far zone */
static float scaled_probe(world sim, float z)
{
	return get(sim, Ex, 0, 0, z) * z * sqrtf(8*PI/3);
}

int main(int argc, char **argv)
{
	dom octant = {{500}, {500}, {500}};
	res grid = {10, {20}, {1240}, {{2, 100}, {2, 100}, {2, 100}}};   /* cells double outside +-100 nm */
	sur walls = {{SYM, PML}, {SYM, PML}, {SYM, PML}};
	world sim = createWorld(octant, grid, walls, "%s.out", argv[0]);

	pointDipole(sim, Ex, 0, 0, 0, Band, 400, 800);
	phaser list = createPhasers(sim, 400, 800, 5);
	const long steps = list->N;                                   /* one beat period of the frequency comb */

	for (long step = 1; timer(step, steps); step++) {
		updateE(sim);
		updateH(sim);
		updatePhaser(sim, list);
		const float at_250 = scaled_probe(sim, 250), at_500 = scaled_probe(sim, 500);
		writeSpectrum(sim, steps, 400, 800, "/Near-field", at_250, at_500);
	}
	writePoyntingSpectrum(sim, "/Power", list);
	writeFarFieldSpectrum(sim, Unpol, 0, 0, "/Far-field", list);

	phaser green = choosePhaser(list, 500);
	farFieldTheta(sim, green, 0, "/Far-field-theta");
	printf("Total radiated power: %f\n", phaserPoyntingOut(green));
	printf("Intensity to (0,0) / (3/(8*PI)): %f\n", farFieldI(green, Unpol, 0, 0) * 8*PI/3);
	printf("Radiated power to hemisphere * 2: %f\n", farFieldFlux(green, Unpol, 0, 0, 90) * 2);
	return 0;
}
