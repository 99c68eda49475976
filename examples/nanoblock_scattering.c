/* Absorption and scattering cross sections of a gold nanoblock on a glass substrate, illuminated
 * by an oblique (55 deg, P-polarised) plane wave injected through a TFSF box; Drude + 2
 * critical-point gold, CPML.  Same physics as the reference's examples/Nanoblock_Scattering.c.
 * Output: <ID>/Cross-section.txt (t, absorbed, scattered, extinction per period). */
#include <alis.h>

/* net power into the box around the particle (absorption) and out of the larger box in the
 * scattered-field region (scattering), both taken after each half step */
static void add_fluxes(world sim, float *absorbed, float *scattered)
{
	*absorbed  += poyntingIn (sim, -250, 250, -250, 250, -250, 250);
	*scattered += poyntingOut(sim, -350, 350, -350, 350, -350, 350);
}

int main(int argc, char **argv)
{
	dom total_field = {{300}, {300}, {300}};
	dom whole = {{400}, {400}, {400}};
	res grid = {5, {10}, {1240}};
	sur walls = {{PML}, {SYM, PML}, {PML}};
	world sim = createWorld(whole, grid, walls, "%s.out", argv[0]);

	object block = {Box, {{-200, 200}, {-200, 200}, {-200, 200}}};
	object glass = {Box, {{-INF, INF}, {-INF, INF}, {-INF,-200}}};
	putObjects(sim, Au, block, n(1.5), glass, Air);
	planewave(sim, total_field, Ppol, 55, 0, Sine, 500, 20);

	const long periods = 40, period = sim->T;
	float absorbed = 0, scattered = 0;
	for (long step = 1; timer(step, periods*period); step++) {
		updateH(sim);
		add_fluxes(sim, &absorbed, &scattered);
		updateE(sim);
		add_fluxes(sim, &absorbed, &scattered);
		if (step % period) continue;
		writeRow(sim, "/Cross-section", sim->t, absorbed/period, scattered/period, (absorbed+scattered)/period);
		absorbed = scattered = 0;
	}
	return 0;
}
