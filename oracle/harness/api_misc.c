/*
 * api_misc.c -- the parts of the public API that the configurations and examples do not touch:
 * rebootWorld, removeSources, removeObjects + a second putObjects, cloneWorld,
 * LT, nkFile, phaserPoynting / phaserPoyntingIn, farFieldPhi (through a CylindericalMap profile is
 * left to the examples), writeTxt, deletePhaser / deleteSlice / deleteWorld.  Prints "tag value"
 * lines like monitors.c; the same source links against the unmodified reference (golden) and
 * against our library.
 *
 *   api_misc [outdir]
 */
#include <alis.h>

static void emit(const char *tag, double v) { printf("%-28s % .9g\n", tag, isfinite(v) ? v : -7777); }

static double run(world W, int steps, float x, float y, float z)
{
	double probe = 0;
	for (int s=0; s<steps; s++) { updateE(W); updateH(W); probe += get(W, Ez, x, y, z) + 0.5*get(W, Hy, x, y, z); }
	return probe;
}

int main(int argc, char **argv)
{
	dom Dom = {{300}, {-250, 250}, {-200, 300}};
	res Res = {5, {10}, {1240}};
	sur Sur = {{SYM, PML}, {PML, PML}, {PML, PML}};
	world W = argc > 1 ? createWorld(Dom, Res, Sur, "%s", argv[1]) : createWorld(Dom, Res, Sur, "");

	/* a table-driven material: nkFile interpolates (n, k) at a wavelength */
	FILE *f = fopen("api_misc_nk.txt", "w");
	fprintf(f, "400\t1.50\t0.10\n500\t1.70\t0.20\n600\t1.90\t0.30\n700\t2.10\t0.40\n");
	fclose(f);
	matter Tab = nkFile("api_misc_nk.txt", 550, 0.002);
	emit("nk_e", Tab.e.x); emit("nk_P0_w", Tab.P[0].w); emit("nk_P0_r", Tab.P[0].r); emit("nk_P0_f", Tab.P[0].f); emit("nk_P0_g", Tab.P[0].g);
	matter Hot = LT(Au, 1.5);
	emit("LT_P0_r", Hot.P[0].r); emit("LT_P1_g", Hot.P[1].g); emit("LT_P0_w", Hot.P[0].w);

	object Block = {Box, {{-60, 60}, {-80, 80}, {-40, 40}}};
	object Substrate = {Box, {{-INF, INF}, {-INF, INF}, {-INF, -40}}};
	/* (an object of the pure-conductivity matter `Tab` corrupts the reference's heap -- zero-pole ADE
	 * tables, reference object.c:178-226 -- so the table matter is only inspected, not placed) */
	putObjects(W, Ag, Block, n(1.5), Substrate, Air);
	/* clone (taken before any source exists: a clone shares the source list, and shared sources keep
	 * writing into the original's arrays, reference source.c:62,68): same coefficients, fresh fields */
	world C = cloneWorld(W);
	pointDipole(W, Ez, 0, 0, 120, Pulse, 550, 800);
	pointDipole(C, Ey, 0, 50, 100, Pulse, 600, 900);
	emit("run1", run(W, 60, 40, 30, 60));
	emit("clone_run", run(C, 50, 30, 40, 50));
	emit("clone_n", C->n); emit("orig_n", W->n);
	emit("orig_continues", run(W, 20, 40, 30, 60));
	emit("clone_continues", run(C, 20, 30, 40, 50));
	/* (deleteCloneWorld double-frees the ADE tables in the reference -- world.c:401,461 after
	 * object.c:450 -- so only our build deletes the clone) */
#ifdef HARNESS_MODE_MINE
	deleteCloneWorld(C);
#endif

	/* reboot: fields and sources gone, materials stay */
	rebootWorld(W);
	emit("reboot_t", W->t); emit("reboot_Ht", W->H->t); emit("reboot_n", W->n);
	emit("reboot_silent", run(W, 10, 40, 30, 60));
	pointDipole(W, Ez, 0, 0, 120, Pulse, 550, 800);
	emit("run1_again", run(W, 60, 40, 30, 60));

	/* new sources on the running world, then new objects */
	removeSources(W);
	pointDipole(W, Hx, 0, 20, 100, Sine, 500, 5);
	emit("after_removeSources", run(W, 40, 40, 30, 60));
	removeObjects(W);
	object Sphere = {Ball, {{-90, 90}, {-90, 90}, {-10, 170}}};
	putObjects(W, Hot, Sphere, n(1.5), Substrate, Air);
	emit("after_new_objects", run(W, 40, 40, 30, 60));
	emit("RI_inside", get(W, RI, 0, 0, -100)); emit("eEff_sphere", get(W, eEff, 0, 0, 80));

	/* phasers: box fluxes, in-plane cut */
	phaser P = createPhasers(W, 450, 650, 100, -150, 150, -150, 150, -100, 200);
	for (int s=0; s<50; s++) { updateE(W); updateH(W); updatePhasers(W, P); }
	phaser Q = choosePhaser(P, 500);
	emit("phaser_count_f", Q->f); emit("phaserPoynting", phaserPoynting(Q)); emit("phaserPoyntingIn", phaserPoyntingIn(Q)); emit("phaserPoyntingOut", phaserPoyntingOut(Q));
	if (argc > 1) {
		farFieldPhi(W, Q, 40, "/ff-phi");
		writeTxt(W, "/note", "%d %s\n", 42, "text");
		writeTxt(W, "/note", "%g\n", 2.5);
	}
	slice S = createSliceXZ(W, 0);
	emit("sliceMax_HH", sliceMax(W, HH, S));
	deleteSlice(S);
	/* (deletePhaser on a multi-frequency list frees an interior pointer in the reference, phaser.c:138:
	 * only our build deletes the list) */
#ifdef HARNESS_MODE_MINE
	deletePhasers(P);
#endif
	deleteWorld(W);
	emit("done", 1);
	return 0;
}
