/*
 * bench_cube.c -- times updateE+updateH of an all-PML vacuum box through the PUBLIC API
 * (createWorld + pointDipole + the run loop), i.e. the synthetic configuration of SURVEY 8d.
 *
 *   bench_cube <nx> <ny> <nz> <steps> <warmup> [probe_every_step [au]]
 *
 * prints one line:  cells=<n> steps=<k> seconds=<s> mcells_per_s=<v> threads=<t> probe=<Ez>
 * With probe_every_step = 1 the loop also reads get(W, Ez, ...) after every step, i.e. the way a
 * driver with a point monitor runs (one 4-byte device->host read per step in our library).
 * With 2 it additionally evaluates sliceMax(W, EE, XY) on the z=0 plane every 4th step, the monitor
 * cadence of examples/Dipole_PML.c:21-25 without the file output (in our library: a device->host
 * read of the plane's E components, ~100 MB at 2048^2).
 * With a 7th argument "au" a gold block (alis_m.h Au: one Drude pole + two critical-point pairs) covering
 * the central half of every axis, i.e. 1/8 of the interior, is placed first (putObjects): the dispersive
 * variant of the benchmark (SURVEY 8d additive terms); the line then also reports ade_cells.
 * Linked against the unmodified reference it is the CPU baseline (`--impl reference` and
 * cpu_baseline.kind "reference" in bench.py).
 */
#include <alis.h>
#include <sys/time.h>
#include <omp.h>

static double now(void) { struct timeval t; gettimeofday(&t, 0); return t.tv_sec + 1e-6*t.tv_usec; }

int main(int argc, char **argv)
{
	if (argc < 6) { printf("usage: %s nx ny nz steps warmup\n", argv[0]); return 2; }
	int n[3] = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3])}, steps = atoi(argv[4]), warm = atoi(argv[5]);
	int every = argc > 6 ? atoi(argv[6]) : 0;
	double acc = 0;
	float lo[3], hi[3];
	for (int d=0; d<3; d++) {            /* interior = n-26 cells, centred like alis_b200/synthetic.py */
		int interior = n[d] - 26;
		int Min = -(interior/2), Max = Min + interior - 1;
		lo[d] = 10.f*Min; hi[d] = 10.f*Max;
	}
	dom Dom = {{lo[0], hi[0]}, {lo[1], hi[1]}, {lo[2], hi[2]}};
	res Res = {5, {10}, {1240}};
	sur Sur = {{PML, PML}, {PML, PML}, {PML, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	if (W->iNUM != n[0] || W->jNUM != n[1] || W->kNUM != n[2]) { printf("grid mismatch %d %d %d\n", W->iNUM, W->jNUM, W->kNUM); return 1; }
	double ade_cells = 0;
	if (argc > 7 && !strcmp(argv[7], "au")) {
		object Block = {Box, {{lo[0]/2, hi[0]/2}, {lo[1]/2, hi[1]/2}, {lo[2]/2, hi[2]/2}}};
		putObjects(W, Au, Block, Air);
		for (int q=0; q<W->CE->N; q++)
			ade_cells += (double)(W->CE->iMax[q]-W->CE->iMin[q]+1)*(W->CE->jMax[q]-W->CE->jMin[q]+1)*(W->CE->kMax[q]-W->CE->kMin[q]+1);
	}
	pointDipole(W, Ez, 0, 0, 0, Pulse, 400, 800, 1);
	slice XY = createSliceXY(W, 0);
	double smax = 0;
	for (int s=0; s<warm; s++) { updateE(W); updateH(W); if (every) acc += get(W, Ez, 0, 0, 50); if (every == 2 && s % 4 == 3) smax += sliceMax(W, EE, XY); }
	acc = get(W, Ez, 0, 0, 50);                       /* (also drains the device queue before the clock starts) */
	double t0 = now();
	for (int s=0; s<steps; s++) { updateE(W); updateH(W); if (every) acc += get(W, Ez, 0, 0, 50); if (every == 2 && s % 4 == 3) smax += sliceMax(W, EE, XY); }
	acc += get(W, Ez, 0, 0, 50);                      /* device work is complete when this returns */
	double dt = now() - t0;
	double cells = (double)n[0]*n[1]*n[2];
	printf("cells=%.0f steps=%d seconds=%.6f mcells_per_s=%.3f threads=%d probe=%.9g slicemax=%.9g ade_cells=%.0f\n", cells, steps, dt, cells*steps/dt*1e-6,
		omp_get_max_threads(), acc, smax, ade_cells);
	return 0;
}
