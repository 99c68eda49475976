/*
 * update.c -- the run loop entry points.  updateE / updateH hand the step to the B200 engine
 * (reference update.c:275-341 ran the loops on the CPU); timer() is the progress / loop-condition
 * helper of reference update.c:14-45.
 */
#include <sys/time.h>
#include <string.h>
#include "alis_c.h"
#include "host_internal.h"

int alis_bridge_device(void)
{
	const char *s = getenv("ALIS_DEVICE");
	return s ? atoi(s) : 0;
}

/* which device array a source adds to (the reference compares S->F with the field pointers,
 * source.c:59-69; here createSource records the target explicitly) */
int alis_source_target(world W, source S, int set, int *comp, int *m)
{
	(void)W;
	if (S->b200_set != set) return -1;
	*comp = S->b200_comp; *m = S->b200_m;
	return 0;
}

void alis_host_attach(world W) { W->b200_bridge = 0; W->b200_view = 0; }

alis_bridge *alis_host_bridge(world W)
{
	if (!W->b200_bridge) { ALIS_PROF("engine create (CUDA init)"); W->b200_bridge = alis_bridge_create(W); }
	return (alis_bridge*)W->b200_bridge;
}

void alis_host_detach(world W)
{
	alis_flux_cache_drop(W);
	alis_view_free(W);
	if (W->b200_bridge) alis_bridge_destroy((alis_bridge*)W->b200_bridge);
	W->b200_bridge = 0;
}

void alis_host_dirty(world W, int mask) { if (W->b200_bridge) alis_bridge_mark(W->b200_bridge, mask); }
void alis_host_invalidate(world W) { alis_host_dirty(W, ALIS_DIRTY_ALL); }

void alis_host_reset_fields(world W)
{
	if (!W->b200_bridge) return;                /* never stepped: device state is still zero */
	if (alis_bridge_reset_fields(W->b200_bridge) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
	alis_host_touch(W);
}

/* ---- ALIS_PROFILE=1 ------------------------------------------------------------------------------------- */
#include <sys/time.h>
static struct { const char *name; double s; long n; } prof_tab[64];
static int prof_n, prof_mode = -1;
static double prof_start;
double alis_prof_now(void) { struct timeval t; gettimeofday(&t, 0); return t.tv_sec + 1e-6*t.tv_usec; }
static void prof_report(void)
{
	fprintf(stderr, "[alis-b200 profile] %-28s %10s %10s\n", "entry point (inclusive)", "calls", "seconds");
	for (int q=0; q<prof_n; q++) fprintf(stderr, "[alis-b200 profile] %-28s %10ld %10.4f\n", prof_tab[q].name, prof_tab[q].n, prof_tab[q].s);
	fprintf(stderr, "[alis-b200 profile] %-28s %10s %10.4f\n", "since first profiled call", "", alis_prof_now() - prof_start);
}
int alis_prof_enabled(void)
{
	if (prof_mode < 0) {
		const char *e = getenv("ALIS_PROFILE");
		prof_mode = e && atoi(e) ? 1 : 0;
		if (prof_mode) { prof_start = alis_prof_now(); atexit(prof_report); }
	}
	return prof_mode;
}
void alis_prof_add(const char *name, double seconds)
{
	int q = 0;
	while (q < prof_n && prof_tab[q].name != name && strcmp(prof_tab[q].name, name)) q++;
	if (q == prof_n) { if (prof_n == 64) return; prof_tab[prof_n++].name = name; }
	prof_tab[q].s += seconds; prof_tab[q].n++;
}

void updateE(world W)
{
	ALIS_PROF("updateE");
	alis_bridge_step_E(alis_host_bridge(W));
	alis_host_touch(W);
}

void updateH(world W)
{
	ALIS_PROF("updateH");
	alis_bridge_step_H(alis_host_bridge(W));
	alis_host_touch(W);
}

/* Public in the reference (alis_c.h:181-182) and applied there to host arrays.  Fields are device
 * resident here: the copy is applied to the host views passed in and is NOT written back -- the
 * engine performs all boundary copies of a step itself (update.c:284-297,311-339). */
void (copyBoundary)(float ***F, float ***G, float cosk, float sink, int iMin, int iMax, int jMin, int jMax, int kMin, int kMax, int iAlt, int jAlt, int kAlt)
{
	if (!F) return;
	for (int i=iMin; i<=iMax; i++) for (int j=jMin; j<=jMax; j++) for (int k=kMin; k<=kMax; k++) {
		if (G) {
			F[i][j][k] = cosk * F[i+iAlt][j+jAlt][k+kAlt] + sink * G[i+iAlt][j+jAlt][k+kAlt];
			G[i][j][k] = cosk * G[i+iAlt][j+jAlt][k+kAlt] - sink * F[i+iAlt][j+jAlt][k+kAlt];
		}
		else F[i][j][k] = cosk * F[i+iAlt][j+jAlt][k+kAlt];
	}
}

/* chi(2) second-harmonic / three-wave mixing (reference update.c:354-411): the signal world's E
 * receives 2*Alt*d*inv_eps*fill*(E1..*E2..) inside every ADE object box.  All worlds are device
 * resident; the coupling is one engine call per object and field set (kernel chi2_couple).  The
 * worlds must share one grid (the reference indexes all three with the same i,j,k).         */
static void same_grid(world A, world B, const char *who)
{
	if (A->iMIN != B->iMIN || A->iMAX != B->iMAX || A->jMIN != B->jMIN || A->jMAX != B->jMAX || A->kMIN != B->kMIN || A->kMAX != B->kMAX
	    || B->complexField < A->complexField) {
		printf("%s: the coupled worlds must have the same grid.\n", who);
		exit(0);
	}
}

void nonLinearSHG(world WS, world WP, float Alt)
{
	same_grid(WS, WP, "nonLinearSHG");
	alis_bridge_chi2(alis_host_bridge(WS), alis_host_bridge(WP), alis_host_bridge(WP), Alt);
	alis_host_touch(WS);
}

void nonLinearTWM(world WS, world WP1, world WP2, float Alt)
{
	same_grid(WS, WP1, "nonLinearTWM");
	same_grid(WS, WP2, "nonLinearTWM");
	alis_bridge_chi2(alis_host_bridge(WS), alis_host_bridge(WP1), alis_host_bridge(WP2), Alt);
	alis_host_touch(WS);
}

/* progress line at most once per second; returns n<=N so it can be the loop condition
 * (reference update.c:14-45) */
int (timer)(long int n, long int N, char *str, ...)
{
	static long int firstStep, lastStep;
	static unsigned long long firstTime, lastTime;
	struct timeval tv;
	gettimeofday(&tv, 0);
	const unsigned long long now = tv.tv_sec * 1000000LL + tv.tv_usec;

	if (!n || !firstStep) {
		firstStep = lastStep = n;
		firstTime = lastTime = now;
	}
	else if (n>N || now/1000000LL > lastTime/1000000LL) {
		unsigned long long elapsed = now - firstTime;
		unsigned long long total = elapsed + (N-n+1) * (now-lastTime) / (n-lastStep);
		if (total - elapsed > 3600000000LL) total -= total%60000000LL;
		lastStep = n;
		lastTime = now;
		if (str[0]) printf("\r%s ", str);
		else printf("\r");
		printf("[%02lld:%02lld:%02lld/%02lld:%02lld:%02lld] [%ld/%ld: %.1f%%]",
			elapsed/3600000000LL, elapsed%3600000000LL/60000000LL, elapsed%60000000LL/1000000LL,
			total/3600000000LL, total%3600000000LL/60000000LL, total%60000000LL/1000000LL,
			(n-1), N, 100.0*(n-1)/N);
		fflush(stdout);
		if (n>N) printf("\n");
	}
	return n<=N;
}
