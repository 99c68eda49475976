/* Absorption and scattering cross sections of a gold nanoblock on a glass substrate, illuminated
 * by an oblique (55 deg, P-polarised) plane wave injected through a TFSF box; Drude + 2
 * critical-point gold, CPML.  Same physics as the reference's examples/Nanoblock_Scattering.c.
 * Output: <ID>/Cross-section.txt (t, absorbed, scattered, extinction per period). */
#include <alis.h>

int main(int argc, char **argv)
{
	dom DTF = {{300}, {300}, {300}};          /* total-field region     */
	dom DSF = {{400}, {400}, {400}};          /* scattered-field region */
	res Res = {5, {10}, {1240}};
	sur Sur = {{PML}, {SYM, PML}, {PML}};
	world W = createWorld(DSF, Res, Sur, "%s.out", argv[0]);

	object Nanoblock = {Box, {{-200, 200}, {-200, 200}, {-200, 200}}};
	object Substrate = {Box, {{-INF, INF}, {-INF, INF}, {-INF,-200}}};
	putObjects(W, Au, Nanoblock, n(1.5), Substrate, Air);
	planewave(W, DTF, Ppol, 55, 0, Sine, 500, 20);

	float absorbed = 0, scattered = 0;
	for (int n=1, N=40*W->T; timer(n, N); n++) {
		updateH(W);
		absorbed  += poyntingIn (W,-250, 250,-250, 250,-250, 250);
		scattered += poyntingOut(W,-350, 350,-350, 350,-350, 350);
		updateE(W);
		absorbed  += poyntingIn (W,-250, 250,-250, 250,-250, 250);
		scattered += poyntingOut(W,-350, 350,-350, 350,-350, 350);
		if (n%W->T == 0) {
			writeRow(W, "/Cross-section", W->t, absorbed/W->T, scattered/W->T, (absorbed+scattered)/W->T);
			absorbed = scattered = 0;
		}
	}
	return 0;
}
