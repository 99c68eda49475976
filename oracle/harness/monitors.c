/*
 * monitors.c -- exercises the monitor API (probes, flux boxes, absorption, world integrals,
 * slices, phasers, phasor flux, far field) on the named configurations and prints every value
 * with %.9g, one per line.  The same source is linked against the unmodified reference
 * (golden) and against our library; the two outputs are compared value by value.
 *
 *   monitors <config> <steps> [outdir]
 */
#include <alis.h>
#include <omp.h>
#include "configs.h"

static void emit(const char *tag, double v) { printf("%-28s % .9g\n", tag, v); }

/* user code indexing the phasor arrays directly, as reference examples/Cross_Polarization_T.c:3-19
 * does: no library call in between that could refresh them */
static double direct_phasor_sum_serial(phaser Q)
{
	double s = 0;
	for (int i=Q->iMin; i<=Q->iMax; i++) for (int j=Q->jMin; j<=Q->jMax; j++) {
		const float complex e = Q->zMax.Ex[i][j][Q->kMax], h = Q->zMax.Hy[i][j][Q->kMax];
		s += crealf(e*conjf(h)) + 0.5*cabsf(Q->zMax.Ey[i][j][Q->kMax]);
	}
	return s;
}

static double direct_phasor_sum(phaser Q)
{
#ifdef HARNESS_MODE_MINE
	/* HARNESS_DIRECT_PHASOR=1: rely on the page-protection guard of the phasor views (the reference's
	 * semantics); otherwise refresh them explicitly first.  Same values either way.
	 * HARNESS_DIRECT_PHASOR=omp: every thread of an OpenMP team walks the arrays at once, as
	 * Cross_Polarization_T.c does -- all of them must see the complete, fresh data */
	const char *mode = getenv("HARNESS_DIRECT_PHASOR");
	if (!mode) alisSyncPhaser(Q);
	if (mode && !strcmp(mode, "omp")) {
		double part[64];
		int nt = 0;
		#pragma omp parallel num_threads(8)
		{
			#pragma omp single
			nt = omp_get_num_threads();
			part[omp_get_thread_num()] = direct_phasor_sum_serial(Q);
		}
		for (int t=1; t<nt; t++) if (part[t] != part[0]) return NAN;
		return part[0];
	}
#endif
	return direct_phasor_sum_serial(Q);
}

/* a slice "writer" that prints checksums of the sampled slice instead of a file: user-supplied
 * methods are part of the API (sliceSnap only special-cases png) */
static void checksum(world W, slice S, float Max, char *F, char *filename, unsigned long int *cmap)
{
	double s1 = 0, s2 = 0; long bad = 0;
	char tag[64];
	(void)W; (void)filename; (void)cmap;
	for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++) {
		const float v = S->F[i][j][k];
		if (isfinite(v)) s1 += v, s2 += (double)v*v; else bad++;
	}
	sprintf(tag, "snap_%s_sum", F); emit(tag, s1);
	sprintf(tag, "snap_%s_sq", F); emit(tag, s2);
	sprintf(tag, "snap_%s_bad", F); emit(tag, bad);
	sprintf(tag, "snap_%s_max", F); emit(tag, isfinite(Max) ? Max : -1);
}

int main(int argc, char **argv)
{
	if (argc < 3) { printf("usage: %s <config> <steps> [outdir]\n", argv[0]); return 2; }
	const harness_config *cfg = harness_find(argv[1]);
	if (!cfg) { printf("unknown config %s\n", argv[1]); return 2; }
	int steps = atoi(argv[2]);
	if (steps <= 0) steps = cfg->default_steps;
	world W = cfg->make();
	if (argc > 3) strcpy(W->ID, argv[3]);

	const int is_nano = !strcmp(cfg->name, "nanoblock"), is_meta = !strcmp(cfg->name, "metasurface");
	const int is_band = !strcmp(cfg->name, "dipole_band"), is_phc = !strcmp(cfg->name, "phc");
	phaser P = 0;
	if (is_meta) P = createPhaser(W, 500,-2000, 2000,-2000, 2000);
	if (is_band) P = createPhasers(W, 400, 800, 50);
	/* random worlds: a two-frequency list on whatever absorbing faces the world has */
	const int is_fuzz = !strcmp(cfg->name, "fuzz");
	if (is_fuzz && (W->xMaxSur==PML || W->yMaxSur==PML || W->zMaxSur==PML)) P = createPhasers(W, 400, 600, 200);
	slice XY = createSliceXY(W, 0), XZ = createSliceXZ(W, 0);

	double flux_in = 0, flux_out = 0, probe = 0, absn = 0, direct_mid = 0;
	object Nanoblock = {Box, {{-200, 200}, {-200, 200}, {-200, 200}}};
	for (int s=1; s<=steps; s++) {
		if (cfg->order) updateH(W); else updateE(W);
		if (is_nano) { flux_in += poyntingIn(W,-250, 250,-250, 250,-250, 250); flux_out += poyntingOut(W,-350, 350,-350, 350,-350, 350); }
		if (cfg->order) updateE(W); else updateH(W);
		if (is_nano) { flux_in += poyntingIn(W,-250, 250,-250, 250,-250, 250); flux_out += poyntingOut(W,-350, 350,-350, 350,-350, 350); absn += objectAbsorption(W, Nanoblock); }
		if (P && steps - s < 40) updatePhaser(W, P);
		if (P && steps - s == 25 && (!is_fuzz || W->zMaxSur==PML)) direct_mid = direct_phasor_sum(is_band ? choosePhaser(P, 500) : P);
		probe += get(W, is_phc ? Hz : Ez, 20, 10, 30);
	}
	emit("t", W->t); emit("n", W->n);
	emit("probe_sum", probe);
	emit("get_Ex", get(W, Ex, 30, 20, 10)); emit("get_Ey", get(W, Ey, 30, 20, 10)); emit("get_Ez", get(W, Ez, 30, 20, 10));
	emit("get_Hx", get(W, Hx, 30, 20, 10)); emit("get_Hy", get(W, Hy, 30, 20, 10)); emit("get_Hz", get(W, Hz, 30, 20, 10));
	emit("get_Ex_mirror", get(W, Ex, -30, -20, -10)); emit("get_Hy_mirror", get(W, Hy, -30, -20, -10));
	emit("get_iEz", get(W, iEz, 30, 20, 10)); emit("get_iHx", get(W, iHx, 30, 20, 10));
	emit("get_EE", get(W, EE, 10, 20, 30)); emit("get_HH", get(W, HH, 10, 20, 30)); emit("get_U", get(W, U, 10, 20, 30));
	emit("get_Sx", get(W, Sx, 10, 20, 30)); emit("get_Sy", get(W, Sy, 10, 20, 30)); emit("get_Sz", get(W, Sz, 10, 20, 30));
	emit("get_Txx", get(W, Txx, 10, 20, 30)); emit("get_Txy", get(W, Txy, 10, 20, 30)); emit("get_Tyz", get(W, Tyz, 10, 20, 30)); emit("get_Tzx", get(W, Tzx, 10, 20, 30));
	emit("get_DivE", get(W, DivE, 10, 20, 30)); emit("get_RI", get(W, RI, 10, 20, 30)); emit("get_eEff", get(W, eEff, 10, 20, -150));
	emit("get_Jx", get(W, Jx, 10, 20, 30)); emit("get_JE", get(W, JE, 10, 20, 30)); emit("get_Fz", get(W, Fz, 10, 20, 30));
	emit("get_ScEx", get(W, ScEx, 10, 20, 30)); emit("get_ScEE", get(W, ScEE, 10, 20, 30)); emit("get_SoEz", get(W, SoEz, 10, 20, 30)); emit("get_SoEE", get(W, SoEE, 10, 20, 30));
	/* every remaining public field function at two points (one of them a mirror image on SYM worlds);
	 * non-finite values (log of zero, 0/0 of a dispersive model at W->f = 0) are printed as class codes */
	{
		struct { const char *name; field *F; } rest[] = {
			{"Contour", Contour}, {"ContourRI", ContourRI}, {"EyHz", EyHz}, {"EzHy", EzHy}, {"Fx", Fx}, {"Fy", Fy},
			{"JxHy", JxHy}, {"JxHz", JxHz}, {"Jy", Jy}, {"JyHx", JyHx}, {"JyHz", JyHz}, {"Jz", Jz}, {"JzHx", JzHx}, {"JzHy", JzHy},
			{"LogEE", LogEE}, {"LogHH", LogHH}, {"LogJE", LogJE}, {"LogRI", LogRI}, {"LogU", LogU}, {"LogUE", LogUE}, {"LogUH", LogUH},
			{"ScEy", ScEy}, {"ScEz", ScEz}, {"SoEx", SoEx}, {"SoEy", SoEy}, {"Txz", Txz}, {"Tyx", Tyx}, {"Tzy", Tzy}, {"UE", UE}, {"UH", UH},
			{"iEy", iEy}, {"iHy", iHy}, {"rawEx", rawEx}, {"rawEy", rawEy}, {"rawEz", rawEz}, {"rawHx", rawHx}, {"rawHy", rawHy}, {"rawHz", rawHz},
			{"rawJx", rawJx}, {"rawJy", rawJy}, {"rawJz", rawJz}, {"raweEffx", raweEffx}, {"raweEffy", raweEffy}, {"raweEffz", raweEffz},
			{"ExHy", ExHy}, {"ExHz", ExHz}, {"EyHx", EyHx}, {"EzHx", EzHx}, {"Tyy", Tyy}, {"Tzz", Tzz}, {"Jx", Jx}, {"eEff", eEff}, {"RI", RI}
		};
		char tag[64];
		for (unsigned q=0; q<sizeof rest/sizeof rest[0]; q++) for (int pt=0; pt<2; pt++) {
			/* second point: mirrored only along axes that have a mirror plane (elsewhere it would lie outside the domain) */
			const float v = pt ? get(W, rest[q].F, W->xMinSur==SYM ? -30 : 30, W->yMinSur==SYM ? -20 : 20, W->zMinSur==SYM ? -10 : 10) : get(W, rest[q].F, 10, 20, 30);
			sprintf(tag, "at%d_%s", pt, rest[q].name);
			emit(tag, isfinite(v) ? v : isnan(v) ? -7777 : v > 0 ? 8888 : -8888);
		}
	}
	emit("worldMax_Ez", worldMax(W, Ez)); emit("worldSum_U", worldSum(W, U));
	emit("sliceMax_XY_EE", sliceMax(W, EE, XY)); emit("sliceSum_XZ_U", sliceSum(W, U, XZ));
	emit("poyntingX", poyntingX(W, 100)); emit("poyntingY", poyntingY(W, 50, -100, 100, -80, 120)); emit("poyntingZ", poyntingZ(W, 100, -150, 150));
	emit("poyntingOut", poyntingOut(W, -120, 120, 20, 120, -100, 140));
	/* slices and world reductions of the field functions (evaluated on the device by alis-b200) */
	{
		struct { const char *name; field *F; } fn[] = {
			{"rawEx", rawEx}, {"rawEz", rawEz}, {"rawHy", rawHy}, {"Ex", Ex}, {"Ey", Ey}, {"Ez", Ez}, {"Hx", Hx}, {"Hy", Hy}, {"Hz", Hz},
			{"iEx", iEx}, {"iHz", iHz}, {"EE", EE}, {"HH", HH}, {"UH", UH}, {"ExHy", ExHy}, {"EyHx", EyHx}, {"EzHx", EzHx}, {"ExHz", ExHz},
			{"Sx", Sx}, {"Sy", Sy}, {"Sz", Sz}, {"Txx", Txx}, {"Tyy", Tyy}, {"Tzz", Tzz}, {"Txy", Txy}, {"Tyz", Tyz}, {"Tzx", Tzx},
			{"DivE", DivE}, {"U", U}, {"ScEx", ScEx}
		};
		slice YZ = createSliceYZ(W, 0);
		char tag[64];
		for (unsigned q=0; q<sizeof fn/sizeof fn[0]; q++) {
			sprintf(tag, "worldMax_%s", fn[q].name); emit(tag, worldMax(W, fn[q].F));
			sprintf(tag, "worldSum_%s", fn[q].name); emit(tag, worldSum(W, fn[q].F));
			sprintf(tag, "sliceMax_XZ_%s", fn[q].name); emit(tag, sliceMax(W, fn[q].F, XZ));
			sprintf(tag, "sliceSum_XY_%s", fn[q].name); emit(tag, sliceSum(W, fn[q].F, XY));
			sprintf(tag, "sliceSum_YZ_%s", fn[q].name); emit(tag, sliceSum(W, fn[q].F, YZ));
		}
		sliceSnap(W, Ez, XY, 0, checksum, "chk-%%");
		sliceSnap(W, Hx, YZ, 0, checksum, "chk-%%");
		sliceSnap(W, rawHz, XZ, 0, checksum, "chk-%%");
		sliceSnap(W, EE, XZ, 0, checksum, "chk-%%");
		sliceSnap(W, LogEE, XY, 0, checksum, "chk-%%");
		sliceSnap(W, LogHH, YZ, 0, checksum, "chk-%%");
		sliceSnap(W, Sz, XY, 0, checksum, "chk-%%");
		sliceSnap(W, Txy, XZ, 0, checksum, "chk-%%");
		sliceSnap(W, DivE, YZ, 0, checksum, "chk-%%");
		sliceSnap(W, iEy, XY, 0, checksum, "chk-%%");
		sliceSnap(W, U, XY, 0, checksum, "chk-%%");
		for (int r=0; r<3; r++) { sliceTimeAvg(W, HH, XY, 3, checksum, "avg-%%"); sliceFreqDom(W, Ex, XZ, 3, 500, checksum, "dft-%%"); }
	}
	if (is_nano) { emit("flux_in", flux_in); emit("flux_out", flux_out); emit("absorption", absn); }
	if (P) {
		phaser Q = is_band ? choosePhaser(P, 500) : P;
		if (!is_fuzz || W->zMaxSur==PML) { emit("phasor_direct_mid", direct_mid); emit("phasor_direct_end", direct_phasor_sum(Q)); }
		emit("phaser_f", Q->f); emit("phaser_A", Q->A); emit("phaser_N", Q->N);
		emit("phaserPoyntingOut", phaserPoyntingOut(Q));
		float complex a = farField(Q, Ppol, 20, 30), b = farField(Q, Spol, 140, 250);
		emit("ff_P_20_30_re", crealf(a)); emit("ff_P_20_30_im", cimagf(a));
		emit("ff_S_140_250_re", crealf(b)); emit("ff_S_140_250_im", cimagf(b));
		emit("ffI_unpol_0_0", farFieldI(Q, Unpol, 0, 0)); emit("ffI_P_55_0", farFieldI(Q, Ppol, 55, 0)); emit("ffI_S_89.5_10", farFieldI(Q, Spol, 89.5, 10));
		emit("ffFlux_30", farFieldFlux(Q, Unpol, 0, 0, 30, 0, 5));
		emit("ffFlux_P_tilted", farFieldFlux(Q, Ppol, 10, 20, 25, 0, 5)); emit("ffFlux_S_ring", farFieldFlux(Q, Spol, 170, 45, 10, 30, 5));
		double acc = 0;
		for (int x=0; x<24; x++) for (int y=0; y<12; y++) { float ang[2]; AzimuthalMap(x, y, 24, 12, ang); acc += farFieldI(Q, Unpol, ang[0], ang[1]); }
		emit("ff_map_sum", acc);
	}
#ifdef HARNESS_MODE_MINE
	fprintf(stderr, "device_monitor_calls %lu\n", alisDeviceMonitorCalls());
#endif
	if (argc > 3) {
		sliceSnap(W, Ez, XY, 0, txt, "/snap-%%");
		sliceSnap(W, EE, XZ, 0, png(hot,0), "/snap-%%");
		writeRow(W, "/row", W->t, probe, 2.5);
		if (P) farFieldTheta(W, is_band ? choosePhaser(P, 500) : P, 0, "/ff-theta");
	}
	return 0;
}
