/* Point dipole radiating in vacuum inside a CPML box (one octant, mirror symmetry on the three
 * coordinate planes).  Same physics as the reference's examples/Dipole_PML.c; written against the
 * public API in include/alis.h and buildable against either library (see examples/README.md).
 * Output: PNG frames of two cuts under <ID>/ and <ID>/probe.txt (t, Ez and Hy at x = 300, Ez at z = 500). */
#include <alis.h>

static const float HALF_SIZE = 1000;      /* nm, each axis (the mirror planes double it) */
static const float CELL = 20, TIME_STEP = 10;
static const float RUN_TIME = 3000;

int main(int argc, char **argv)
{
	dom box = {{HALF_SIZE}, {HALF_SIZE}, {HALF_SIZE}};
	res grid = {TIME_STEP, {CELL}, {1240}};
	sur walls = {{SYM, PML}, {SYM, PML}, {SYM, PML}, {24}};          /* 24 cells: true CPML */
	world sim = createWorld(box, grid, walls, "%s.out", argv[0]);

	pointDipole(sim, Ez, 0, 0, 0, Pulse, 400, 800, 10000000);

	slice horizontal = createSlice(sim, sim->xMIN, sim->xMAX, sim->yMIN, sim->yMAX, 0, 0);
	slice vertical   = createSlice(sim, sim->xMIN, sim->xMAX, 0, 0, sim->zMIN, sim->zMAX);

	const int steps = RUN_TIME/sim->dt;
	int step = 1;
	while (timer(step, steps)) {
		updateE(sim);
		updateH(sim);
		sliceSnap(sim, Ez, horizontal, 10, png(dkbr,1), "/%%/");
		sliceSnap(sim, Ez, vertical, 10, png(dkbr,1), "/%%/");
		const float e_near = get(sim, Ez, 300, 0, 0), h_near = get(sim, Hy, 300, 0, 0), e_axis = get(sim, Ez, 0, 0, 500);
		writeRow(sim, "/probe", sim->t, e_near, h_near, e_axis);
		step++;
	}
	return 0;
}
