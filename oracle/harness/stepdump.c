/*
 * stepdump.c -- run a named configuration for K steps and dump the complete field state.
 *
 *   stepdump <config> <steps> <out.bin> [order]
 *
 * order 0: updateE then updateH per step (most examples); 1: updateH then updateE
 * (examples/Nanoblock_Scattering.c:23-27).
 *
 * Build modes (see oracle/build_ref.sh and alis_b200/build.py):
 *   -DHARNESS_MODE_REF     unmodified reference does everything                 (golden)
 *   -DHARNESS_MODE_BRIDGE  reference setup, stepping by our engine via the bridge
 *   -DHARNESS_MODE_MINE    our host layer + our engine (the drop-in library)
 *
 * Output: "ALISDUMP" nx ny nz ncomplex(int32 x4) then for m, for set in (E,H), for comp:
 * (nx+2)(ny+2)(nz+2) floats in the reference's layout (z fastest, ghost layer included).
 */
#include <alis.h>
#include "configs.h"

#if defined(HARNESS_MODE_BRIDGE)
#include "engine_bridge.h"
int alis_source_target(world W, source S, int set, int *comp, int *m)
{
	vfield *F = set == 0 ? W->E : W->H;
	for (int q=0; q<=W->complexField; q++) {
		if (S->F == F[q].x) { *comp = 0; *m = q; return 0; }
		if (S->F == F[q].y) { *comp = 1; *m = q; return 0; }
		if (S->F == F[q].z) { *comp = 2; *m = q; return 0; }
	}
	return -1;
}
int alis_bridge_device(void) { return 0; }
#endif

int main(int argc, char **argv)
{
	if (argc < 4) { printf("usage: %s <config> <steps> <out.bin> [order]\n", argv[0]); return 2; }
	const harness_config *cfg = harness_find(argv[1]);
	if (!cfg) { printf("unknown config %s\n", argv[1]); return 2; }
	int steps = atoi(argv[2]);
	if (steps <= 0) steps = cfg->default_steps;
	int order = argc > 4 ? atoi(argv[4]) : cfg->order;
	world W = cfg->make();
	world P1 = cfg->pump1 ? cfg->pump1() : 0, P2 = cfg->pump2 ? cfg->pump2() : 0;
	const float alt = getenv("HARNESS_ALT") ? (float)atof(getenv("HARNESS_ALT")) : cfg->alt;   /* 0 switches the mixing off */
	world all[3] = {P1, P2, W};                       /* pumps first, the signal world last */

#if defined(HARNESS_MODE_BRIDGE)
	alis_bridge *Bs[3] = {P1 ? alis_bridge_create(P1) : 0, P2 ? alis_bridge_create(P2) : 0, alis_bridge_create(W)};
	alis_bridge *B = Bs[2];
	#define STEP_E(q) alis_bridge_step_E(Bs[q])
	#define STEP_H(q) alis_bridge_step_H(Bs[q])
	#define MIX()     alis_bridge_chi2(Bs[2], cfg->alias ? Bs[2] : Bs[0], cfg->alias ? Bs[0] : (P2 ? Bs[1] : Bs[0]), alt)
#else
	#define STEP_E(q) updateE(all[q])
	#define STEP_H(q) updateH(all[q])
	#define MIX()     do { if (cfg->alias) nonLinearTWM(W, W, P1, alt); else if (P2) nonLinearTWM(W, P1, P2, alt); \
	                       else nonLinearSHG(W, P1, alt); } while (0)
#endif
	for (int s=0; s<steps; s++) {
		if (order) {
			for (int q=0; q<3; q++) if (all[q]) STEP_H(q);
			for (int q=0; q<3; q++) if (all[q]) STEP_E(q);
			if (P1) MIX();
		} else {
			for (int q=0; q<3; q++) if (all[q]) STEP_E(q);
			if (P1) MIX();
			for (int q=0; q<3; q++) if (all[q]) STEP_H(q);
		}
	}

	FILE *f = fopen(argv[3], "wb");
	if (!f) { printf("cannot open %s\n", argv[3]); return 1; }
	int hdr[4] = {W->iNUM, W->jNUM, W->kNUM, W->complexField};
	fwrite("ALISDUMP", 1, 8, f);
	fwrite(hdr, sizeof(int), 4, f);
	size_t cnt = (size_t)(W->iNUM+2)*(W->jNUM+2)*(W->kNUM+2);
	float *buf = malloc(cnt*sizeof(float));
	for (int m=0; m<=W->complexField; m++) for (int set=0; set<2; set++) for (int comp=0; comp<3; comp++) {
#if defined(HARNESS_MODE_BRIDGE)
		alis_bridge_read_field(B, set, comp, m, buf);
#elif defined(HARNESS_MODE_MINE)
		alisReadField(W, set, comp, m, buf);
#else
		vfield *V = set == 0 ? &W->E[m] : &W->H[m];
		float ***A = comp == 0 ? V->x : comp == 1 ? V->y : V->z;
		memcpy(buf, &A[W->iMIN-1][W->jMIN-1][W->kMIN-1], cnt*sizeof(float));
#endif
		fwrite(buf, sizeof(float), cnt, f);
	}
	fclose(f);
	printf("\n%s: %d steps, grid %dx%dx%d complex=%d n=%lu t=%g tE=%g tH=%g\n", cfg->name, steps, W->iNUM, W->jNUM, W->kNUM,
		W->complexField, (unsigned long)W->n, W->t, W->E->t, W->H->t);
#if defined(HARNESS_MODE_MINE)
	fprintf(stderr, "device_geometry_objects %lu\n", alisDeviceGeometryObjects());
#endif
#if defined(HARNESS_MODE_BRIDGE)
	printf("fused_launches=%lu\n", (unsigned long)alisb_fused_launch_count(alis_bridge_engine(B)));
#endif
	return 0;
}
