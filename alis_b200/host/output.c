/*
 * output.c -- scalar monitors and text writers: point probes, world maxima / integrals,
 * time-domain Poynting flux through planes and boxes, per-object absorption, tab-separated text
 * files and the running-DFT spectrum writer.
 *
 * Field data is device resident; every monitor prefetches just the region it integrates over
 * into the host views (view.c) and then evaluates the reference's sums on the host (round 1; the
 * device-side reductions are SURVEY.md 8 row f1).  Definitions follow reference src/output.c
 * (cited per function); sums run in index order, i.e. they equal the reference's result with one
 * OpenMP thread (its `omp reduction` makes the last digits thread-count dependent).
 */
#define _GNU_SOURCE
#include <sys/time.h>
#include <sys/stat.h>
#include <string.h>
#include <pthread.h>
#include <sched.h>
#include "alis_c.h"
#include "host_internal.h"

static char *dup_string(const char *s) { char *d = malloc(strlen(s)+1); strcpy(d, s); return d; }   /* (strdup is not ISO C99) */

/* ---- per (world, name) registry shared by the writers (reference output.c:14-23,107-128) --------- */
typedef struct entry {
	world W;
	char *name;
	FILE *file;
	unsigned long calls;
	unsigned long long stamp;
	float complex **spectrum;
	struct entry *next;
} entry;

/* finds or creates the entry; rewrites `name` into the full path <W->ID><name>; creates the output
 * directory when the name starts with '/' and the leaf directory when it ends with '/' */
static entry *lookup(world W, char *name, entry **list)
{
	entry *e = *list;
	while (e && (e->W != W || strcmp(name, e->name))) e = e->next;
	if (!e) {
		e = calloc(1, sizeof *e);
		e->W = W;
		e->name = dup_string(name);
		e->next = *list;
		*list = e;
		char path[1024];
		strcat(strcpy(path, W->ID), e->name);
		if (e->name[0]=='/') mkdir(W->ID, 0755);
		if (path[0] && path[strlen(path)-1]=='/') mkdir(path, 0755);
	}
	char path[1024];
	strcat(strcpy(path, W->ID), e->name);
	strcpy(name, path);
	e->calls++;
	return e;
}

/* ---- probes ------------------------------------------------------------------------------------------ */
float get(world W, field F, float x, float y, float z)
{
	ALIS_PROF("get");
	return F(W, W->xtoi(W, x), W->ytoj(W, y), W->ztok(W, z));
}

/* reference output.c:56-74 */
float worldMax(world W, field F)
{
	ALIS_PROF("worldMax");
	float best = 0, f;
	const int box[6] = { W->iMin, W->iMax-1, W->jMin, W->jMax-1, W->kMin, W->kMax-1 };
	if (alis_device_absmax(W, F, box, &best)) return best;              /* reduced on the device: one number comes back */
	alis_view_need_for(W, F, W->iMin, W->iMax, W->jMin, W->jMax, W->kMin, W->kMax);
	for (int i=W->iMin; i<W->iMax; i++) for (int j=W->jMin; j<W->jMax; j++) for (int k=W->kMin; k<W->kMax; k++)
		if (best < (f=fabsf(F(W, i, j, k)))) best = f;
	return best;
}

/* trapezoidal volume integral over the interior (reference output.c:77-103) */
static float column_sum(world W, field F, int i, int j)
{
	float s = 0;
	for (int k=W->kMin+1; k<W->kMax; k++) s += W->CE->dz[k] * F(W, i, j, k);
	s += 0.5 * W->CE->dz[W->kMin] * F(W, i, j, W->kMin);
	s += 0.5 * W->CE->dz[W->kMax] * F(W, i, j, W->kMax);
	return s;
}

static float sheet_sum(world W, field F, int i)
{
	float s = 0;
	for (int j=W->jMin+1; j<W->jMax; j++) s += W->CE->dy[j] * column_sum(W, F, i, j);
	s += 0.5 * W->CE->dy[W->jMin] * column_sum(W, F, i, W->jMin);
	s += 0.5 * W->CE->dy[W->jMax] * column_sum(W, F, i, W->jMax);
	return s;
}

/* the same with the z column sums already formed on the device: cols[(i-iMin)*nj + (j-jMin)] */
static float sheet_sum_cols(world W, const float *cols, int i)
{
	const int nj = W->jMax-W->jMin+1;
	const float *c = cols + (size_t)(i-W->iMin)*nj - W->jMin;
	float s = 0;
	for (int j=W->jMin+1; j<W->jMax; j++) s += W->CE->dy[j] * c[j];
	s += 0.5 * W->CE->dy[W->jMin] * c[W->jMin];
	s += 0.5 * W->CE->dy[W->jMax] * c[W->jMax];
	return s;
}

float worldSum(world W, field F)
{
	ALIS_PROF("worldSum");
	float Sum = 0;
	const int box[6] = { W->iMin, W->iMax, W->jMin, W->jMax, W->kMin, W->kMax };
	float *cols = xmalloc(sizeof(float)*(size_t)(W->iMax-W->iMin+1)*(W->jMax-W->jMin+1));
	if (cols && alis_device_column_sums(W, F, box, cols)) {              /* one float per z column comes back, not the fields */
		for (int i=W->iMin+1; i<W->iMax; i++) Sum += W->CE->dx[i] * sheet_sum_cols(W, cols, i);
		Sum += 0.5 * W->CE->dx[W->iMin] * sheet_sum_cols(W, cols, W->iMin);
		Sum += 0.5 * W->CE->dx[W->iMax] * sheet_sum_cols(W, cols, W->iMax);
		free(cols);
	}
	else {
	free(cols);
	alis_view_need_for(W, F, W->iMin, W->iMax, W->jMin, W->jMax, W->kMin, W->kMax);
	for (int i=W->iMin+1; i<W->iMax; i++) Sum += W->CE->dx[i] * sheet_sum(W, F, i);
	Sum += 0.5 * W->CE->dx[W->iMin] * sheet_sum(W, F, W->iMin);
	Sum += 0.5 * W->CE->dx[W->iMax] * sheet_sum(W, F, W->iMax);
	}
	if (W->xMinSur==SYM) Sum *= 2;
	if (W->yMinSur==SYM) Sum *= 2;
	if (W->zMinSur==SYM) Sum *= 2;
	return Sum;
}

/* ---- text writers (reference output.c:131-174) ------------------------------------------------------ */
void (writeTxt)(world W, char *format, ...)
{
	static entry *files;
	char str[1024], filename[1024];
	va_list ap;
	va_start(ap, format);
	vsprintf(filename, format, ap);
	format = va_arg(ap, char*);
	vsprintf(str, format, ap);
	va_end(ap);
	entry *e = lookup(W, filename, &files);
	struct timeval tv;
	gettimeofday(&tv, 0);
	unsigned long long now = tv.tv_sec * 1000000LL + tv.tv_usec;

	if (!e->stamp) {
		e->file = fopen(strcat(filename, ".txt"), "w");
		e->stamp = now;
		if (!e->file) { printf("alis-b200: cannot open %s\n", filename); exit(0); }
	}
	fputs(str, e->file);
	if (now - e->stamp >= 300*1000000LL) {       /* flush every five minutes, as the reference */
		fflush(e->file);
		e->stamp = now;
	}
}

/* one row of tab-separated %g values; the macro appends a NaN-boxed sentinel */
void (writeRow)(world W, char *format, ...)
{
	char str[1024], filename[1024];
	va_list ap;
	va_start(ap, format);
	vsprintf(filename, format, ap);
	double f = va_arg(ap, double);
	if (!NAN2INT(f)) {
		int n = sprintf(str, "%g", f);
		f = va_arg(ap, double);
		while (!NAN2INT(f)) {
			n += sprintf(str+n, "\t%g", f);
			f = va_arg(ap, double);
		}
		sprintf(str+n, "\n");
		(writeTxt)(W, "%s", filename, "%s", str, "");
	}
	va_end(ap);
}

/* ---- the running DFT of writeSpectrum beside the time loop ------------------------------------------------------
 * A writeSpectrum call adds one sample to every bin: `bins` complex exponentials, 9 us of host time per call in
 * PhC_Band_Diagram -- on the critical path of a launch-bound loop, because the sample comes from a `get` that has just
 * waited for the step and the next step is only enqueued afterwards.  The additions do not have to happen now: the
 * bins are read once, when the record is complete.  A call therefore only queues (row, sample, phase step); one worker
 * thread performs the additions, in call order (one consumer, FIFO: every bin receives the same float operations in the
 * same order as the in-line loop, so the spectrum is bit-identical), while the driver's thread is already waiting for
 * the next step.  The call that completes a record drains the queue first.  ALIS_SPECTRUM_THREAD=0 / 1: in line / worker
 * (default: worker with the CUDA engine only). */
typedef struct { float complex *row; int bins, nMax; float t; double d; } spec_job;
#define SPEC_RING 4096
static spec_job spec_ring[SPEC_RING];
static unsigned long spec_head, spec_tail;          /* jobs queued / jobs finished (monotonic) */
static pthread_t spec_thread;
static int spec_mode = -1;                          /* -1 undecided, 0 in line, 1 worker running */
static pthread_mutex_t spec_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t spec_wake = PTHREAD_COND_INITIALIZER;
static int spec_asleep;

static void spec_accumulate(const spec_job *j)
{
	for (int n=0; n<j->bins; n++) j->row[n] += (float) j->d * cexpf(2*PI*I*j->t*(j->nMax-n));      /* reference output.c:214 */
}

static void *spec_worker(void *arg)
{
	(void)arg;
	for (;;) {
		unsigned long idle = 0;
		unsigned long long since = 0;
		while (__atomic_load_n(&spec_head, __ATOMIC_ACQUIRE) == spec_tail) {
			/* stays awake for 300 us after its last job (a time loop that calls every 50..200 us never pays for a wake-up),
			 * then sleeps on the condition variable so that it does not take a core from anybody */
			if ((++idle & 31) == 1) {
				struct timeval tv; gettimeofday(&tv, 0);
				const unsigned long long now = tv.tv_sec*1000000ULL + tv.tv_usec;
				if (!since) since = now;
				if (now - since < 300) { sched_yield(); continue; }
			} else { sched_yield(); continue; }
			pthread_mutex_lock(&spec_lock);
			__atomic_store_n(&spec_asleep, 1, __ATOMIC_SEQ_CST);
			while (__atomic_load_n(&spec_head, __ATOMIC_SEQ_CST) == spec_tail) pthread_cond_wait(&spec_wake, &spec_lock);
			__atomic_store_n(&spec_asleep, 0, __ATOMIC_SEQ_CST);
			pthread_mutex_unlock(&spec_lock);
			idle = 0; since = 0;
		}
		spec_accumulate(&spec_ring[spec_tail % SPEC_RING]);
		__atomic_store_n(&spec_tail, spec_tail + 1, __ATOMIC_RELEASE);
	}
	return 0;
}

static void spec_drain(void)
{
	while (__atomic_load_n(&spec_tail, __ATOMIC_ACQUIRE) != spec_head) sched_yield();
}

static void spec_submit(const spec_job *j)
{
	if (spec_mode < 0) {
		/* default: on with the CUDA engine (the host cores are idle while the device steps), off with a CPU engine (the
		 * oracle's OpenMP team owns every core: a spinning worker measured 17 % slower there) */
		const char *s = getenv("ALIS_SPECTRUM_THREAD");
		const int want = s ? atoi(s) : !strncmp(alisb_backend(), "cuda", 4);
		spec_mode = want && pthread_create(&spec_thread, 0, spec_worker, 0) == 0 ? 1 : 0;
		if (spec_mode) pthread_detach(spec_thread);
	}
	if (!spec_mode) { spec_accumulate(j); return; }
	while (spec_head - __atomic_load_n(&spec_tail, __ATOMIC_ACQUIRE) >= SPEC_RING) sched_yield();     /* ring full: the worker is behind */
	spec_ring[spec_head % SPEC_RING] = *j;
	__atomic_store_n(&spec_head, spec_head + 1, __ATOMIC_SEQ_CST);
	if (__atomic_load_n(&spec_asleep, __ATOMIC_SEQ_CST)) {
		pthread_mutex_lock(&spec_lock);
		pthread_cond_signal(&spec_wake);
		pthread_mutex_unlock(&spec_lock);
	}
}

/* Running DFT of scalar samples between wavelengths LMin..LMax with the resolution of an N-step
 * record; normalised by the spectrum of the driving source; written on the N-th call
 * (reference output.c:177-230). */
void (writeSpectrum)(world W, int N, float LMin, float LMax, char *format, ...)
{
	ALIS_PROF("writeSpectrum");
	static entry *spectra;
	char filename[1024];
	double d;
	float dL = LMin * LMax / (N * W->dt);
	int M = 0, nMin = floor(LMin/dL), nMax = ceil(LMax/dL);
	va_list ap;

	va_start(ap, format);
	vsprintf(filename, format, ap);
	do d = va_arg(ap, double), M++;
	while (!NAN2INT(d));
	va_end(ap);
	va_start(ap, format);
	vsprintf(filename, format, ap);
	entry *e = lookup(W, filename, &spectra);
	const int bins = 1-nMin+nMax;

	if (!e->spectrum) {
		/* the source that defines the record length: of this world or of its TFSF auxiliary world */
		source S = W->SE && W->N==W->SE->N ? W->SE : W->SH && W->N==W->SH->N ? W->SH : 0;
		for (world w = W; !S && w->S && w->S->W; S = w->SE && W->N==w->SE->N ? w->SE : w->SH && W->N==w->SH->N ? w->SH : 0) w = W->S->W;

		MALLOC(e->spectrum, 0, M);
		MALLOC(e->spectrum[0], 0, M*bins);
		for (int m=1; m<M; m++) e->spectrum[m] = e->spectrum[0]+m*bins;

		if (S && W->N) {
			for (int n=0; n<bins; n++) {
				float f = (nMax-n) / (dL * nMin * nMax);
				for (int m=1; m<W->N; m++) e->spectrum[0][n] += S->waveForm(W, S, W->dt*m, S->phase) * cexpf(2*PI*I*f*W->dt*m);
				if (cabsf(e->spectrum[0][n]) < 1E-3) e->spectrum[0][n] = 0;
				else if (W->sourceType == 2) e->spectrum[0][n] *= sqrtf(f * S->wavelength);
				else if (W->sourceType == 3) e->spectrum[0][n] *= f * S->wavelength;
			}
		}
	}

	for (int m=1; m<M; m++) {
		float t = W->t / (dL * nMin * nMax);
		d = va_arg(ap, double);
		const spec_job job = { e->spectrum[m], bins, nMax, t, d };
		spec_submit(&job);
	}
	va_end(ap);

	if (e->calls == (unsigned long)N) {
		if (spec_mode > 0) spec_drain();
		e->file = fopen(strcat(filename, ".txt"), "w");
		if (!e->file) { printf("alis-b200: cannot open %s\n", filename); exit(0); }
		for (int n=0; n<bins; n++) {
			fprintf(e->file, "%g", (dL * nMin * nMax) / (nMax-n));
			for (int m=1; m<M; m++) fprintf(e->file, "\t%g", e->spectrum[0][n] ? cabsf(e->spectrum[m][n]/e->spectrum[0][n]) : 0);
			for (int m=1; m<M; m++) fprintf(e->file, "\t%g", e->spectrum[0][n] ? cargf(e->spectrum[m][n]/e->spectrum[0][n])/PI : 0);
			fprintf(e->file, "\n");
		}
		fclose(e->file);
	}
}

/* ---- time-domain Poynting flux (reference output.c:233-362) ------------------------------------------
 * Flux through the plane normal to axis a at index `at`, over [lo_b,hi_b] x [lo_c,hi_c] with
 * b = a+1, c = a+2 (cyclic):   S_a = E_b H_c - E_c H_b,  H averaged over the two planes that
 * flank the E plane, trapezoidal weights in both tangential directions.  The double sums run ON
 * THE DEVICE (alisb_plane_flux, kernel flux_partial): only two numbers per plane come back, and a
 * Poynting box is a single call with one synchronisation. */
typedef struct {
	int a, at, bMin, bMax, cMin, cMax, mirrored;
	float bLo, bHi, cLo, cHi;
} plane_req;

/* index range of a tangential extent: a degenerate extent becomes the two cells around it */
static void tangential_range(world W, int axis, float lo, float hi, int *iLo, int *iHi)
{
	int (*toi)(world, float) = axis==0 ? W->xtoi : axis==1 ? W->ytoj : W->ztok;
	const int Min = axis==0 ? W->iMin : axis==1 ? W->jMin : W->kMin, Max = axis==0 ? W->iMax : axis==1 ? W->jMax : W->kMax;
	*iLo = lo==hi ? MAX(Min, toi(W, lo)-1) : MAX(Min, toi(W, lo));
	*iHi = lo==hi ? MIN(Max, toi(W, hi)+1) : MIN(Max, toi(W, hi));
}

static plane_req plane_request(world W, int a, float pos, float bLo, float bHi, float cLo, float cHi)
{
	const int b = (a+1)%3, c = (a+2)%3;
	const int sur[3] = { W->xMinSur, W->yMinSur, W->zMinSur };
	int (*toi)(world, float) = a==0 ? W->xtoi : a==1 ? W->ytoj : W->ztok;
	plane_req r = { a, 0, 0, 0, 0, 0, pos<0 && sur[a]==SYM, bLo, bHi, cLo, cHi };
	r.at = toi(W, r.mirrored ? -pos : pos);
	tangential_range(W, b, bLo, bHi, &r.bMin, &r.bMax);
	tangential_range(W, c, cLo, cHi, &r.cMin, &r.cMax);
	return r;
}

/* The two partial sums of each requested plane, from the device.
 *
 * A driver's loop asks for the same flux boxes after every pass (Nanoblock_Scattering.c:24-29: poyntingIn,
 * poyntingOut after updateH and again after updateE), and every device request costs a synchronisation.
 * The planes asked for since the previous pass are therefore remembered: the first request after a pass
 * evaluates them too, in the same launch, and the following requests are answered from that result without
 * touching the device.  Results are tied to the device-state version (bumped by every pass and every field
 * write), so a cached sum is always the sum the kernel would return now -- same kernel, same numbers. */
#define FLUX_SLOTS 24
typedef struct flux_cache {
	world W; unsigned long version;
	int n; alisb_flux_plane P[FLUX_SLOTS]; double s[FLUX_SLOTS][2]; char asked[FLUX_SLOTS];   /* results for `version` */
	int nprev; alisb_flux_plane prev[FLUX_SLOTS];                                             /* planes asked for under the previous version */
	struct flux_cache *next;
} flux_cache;
static flux_cache *flux_caches;

static int same_plane(const alisb_flux_plane *a, const alisb_flux_plane *b)
{
	return a->axis==b->axis && a->at==b->at && a->b0==b->b0 && a->b1==b->b1 && a->c0==b->c0 && a->c1==b->c1;
}

void alis_flux_cache_drop(world W)
{
	for (flux_cache **c = &flux_caches; *c; ) {
		if ((*c)->W == W) { flux_cache *dead = *c; *c = dead->next; free(dead); }
		else c = &(*c)->next;
	}
}

/* the same two sums from the host views: a world cut into z slabs over several devices (ALIS_DEVICES), where a
 * plane may cross the cuts.  Terms and weights of kernel flux_partial (reference output.c:233-332), added in double. */
static void plane_sums_host(world W, const plane_req *r, int n, double *sums)
{
	for (int q=0; q<n; q++) {
		const int a = r[q].a, b = (a+1)%3, c = (a+2)%3;
		int lo[3], hi[3];
		lo[a] = r[q].at; hi[a] = r[q].at + 1; lo[b] = r[q].bMin; hi[b] = r[q].bMax; lo[c] = r[q].cMin; hi[c] = r[q].cMax;
		alis_view_need(W, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);
		float ***E[3] = { W->E->x, W->E->y, W->E->z }, ***H[3] = { W->H->x, W->H->y, W->H->z };
		const float *dE[3] = { W->CE->dx, W->CE->dy, W->CE->dz }, *dH[3] = { W->CH->dx, W->CH->dy, W->CH->dz };
		double s1 = 0, s2 = 0;
		for (int ib=r[q].bMin; ib<=r[q].bMax; ib++) for (int ic=r[q].cMin; ic<=r[q].cMax; ic++) {
			int x[3], y[3];
			x[a] = r[q].at; x[b] = ib; x[c] = ic;
			y[a] = r[q].at + 1; y[b] = ib; y[c] = ic;
			if (ib > r[q].bMin) {
				const float wc = (ic == r[q].cMin || ic == r[q].cMax) ? 0.5f : 1.0f;
				s1 += (double)dH[b][ib] * ((double)wc * (double)dE[c][ic] * (double)E[b][x[0]][x[1]][x[2]] * (double)(H[c][x[0]][x[1]][x[2]] + H[c][y[0]][y[1]][y[2]]));
			}
			if (ic > r[q].cMin) {
				const float wb = (ib == r[q].bMin || ib == r[q].bMax) ? 0.5f : 1.0f;
				s2 -= (double)dH[c][ic] * ((double)wb * (double)dE[b][ib] * (double)E[c][x[0]][x[1]][x[2]] * (double)(H[b][x[0]][x[1]][x[2]] + H[b][y[0]][y[1]][y[2]]));
			}
		}
		sums[2*q] = s1; sums[2*q+1] = s2;
	}
}

static void plane_sums(world W, const plane_req *r, int n, double *sums)
{
	alis_bridge *B = alis_host_bridge(W);
	alis_bridge_commit(B);
	if (alis_bridge_engines(B) > 1) { plane_sums_host(W, r, n, sums); return; }
	alisb_flux_plane P[6];
	const int org[3] = { W->iMIN-1, W->jMIN-1, W->kMIN-1 };
	for (int q=0; q<n; q++) {
		const int a = r[q].a, b = (a+1)%3, c = (a+2)%3;
		P[q].axis = a; P[q].at = r[q].at - org[a];
		P[q].b0 = r[q].bMin - org[b]; P[q].b1 = r[q].bMax - org[b];
		P[q].c0 = r[q].cMin - org[c]; P[q].c1 = r[q].cMax - org[c];
	}
	flux_cache *C = flux_caches;
	while (C && C->W != W) C = C->next;
	if (!C) { C = xcalloc(1, sizeof *C); C->W = W; C->version = alis_view_version(W); C->next = flux_caches; flux_caches = C; }
	if (C->version != alis_view_version(W)) {          /* a pass has run: what was asked for becomes the forecast */
		C->nprev = 0;
		for (int q=0; q<C->n; q++) if (C->asked[q]) C->prev[C->nprev++] = C->P[q];
		C->n = 0; C->version = alis_view_version(W);
	}
	int missing = 0;
	for (int q=0; q<n; q++) {
		int k = 0;
		while (k < C->n && !same_plane(&C->P[k], &P[q])) k++;
		if (k == C->n) missing++;
	}
	if (missing) {
		/* one launch: the planes of this request that are not there yet, then the forecast */
		alisb_flux_plane batch[FLUX_SLOTS];
		int nb = 0;
		for (int pass=0; pass<2; pass++) {
			const alisb_flux_plane *src = pass == 0 ? P : C->prev;
			const int ns = pass == 0 ? n : C->nprev;
			for (int q=0; q<ns; q++) {
				int k = 0, dup = 0;
				while (k < C->n && !same_plane(&C->P[k], &src[q])) k++;
				for (int t=0; t<nb; t++) if (same_plane(&batch[t], &src[q])) dup = 1;
				if (k == C->n && !dup && C->n + nb < FLUX_SLOTS) batch[nb++] = src[q];
			}
		}
		double out[2*FLUX_SLOTS];
		if (alisb_plane_flux(alis_bridge_engine(B), batch, nb, out) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
		for (int t=0; t<nb; t++) { C->P[C->n] = batch[t]; C->s[C->n][0] = out[2*t]; C->s[C->n][1] = out[2*t+1]; C->asked[C->n] = 0; C->n++; }
	}
	for (int q=0; q<n; q++) {
		int k = 0;
		while (k < C->n && !same_plane(&C->P[k], &P[q])) k++;
		if (k == C->n) {                                  /* table full (more than FLUX_SLOTS distinct planes between two passes): ask directly */
			if (alisb_plane_flux(alis_bridge_engine(B), &P[q], 1, &sums[2*q]) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
			continue;
		}
		C->asked[k] = 1;
		sums[2*q] = C->s[k][0]; sums[2*q+1] = C->s[k][1];
	}
}

static float plane_finish(world W, const plane_req *r, double s1, double s2)
{
	const int a = r->a, b = (a+1)%3, c = (a+2)%3;
	const int sur[3] = { W->xMinSur, W->yMinSur, W->zMinSur };
	const int flat[3] = { W->Dom.x.Min==W->Dom.x.Max, W->Dom.y.Min==W->Dom.y.Max, W->Dom.z.Min==W->Dom.z.Max };
	const float *dHb = b==0 ? W->CH->dx : b==1 ? W->CH->dy : W->CH->dz, *dHc = c==0 ? W->CH->dx : c==1 ? W->CH->dy : W->CH->dz;
	/* the reference reduces each term in its own OpenMP loop and then adds the two partial sums */
	float Sum = (float)s1, Sum2 = (float)s2;
	Sum += Sum2;
	/* per-unit-length normalisation of collapsed (2-D) extents; mirror symmetry doubles the flux.
	 * Order of the two axes as in the reference: y before z, x before z, x before y. */
	const int first = MIN(b, c), second = MAX(b, c);
	for (int pass=0; pass<2; pass++) {
		const int ax = pass==0 ? first : second;
		const float lo = ax==b ? r->bLo : r->cLo, hi = ax==b ? r->bHi : r->cHi;
		const float *dH = ax==b ? dHb : dHc;
		const int top = ax==b ? r->bMax : r->cMax;
		if (lo==hi || flat[ax]) Sum /= dH[top]+dH[top-1];
	}
	if (sur[first]==SYM) Sum *= 2;
	if (sur[second]==SYM) Sum *= 2;
	return (r->mirrored ? -0.5 : 0.5) * Sum;
}

static float poynting_plane(world W, int a, float pos, float bLo, float bHi, float cLo, float cHi)
{
	plane_req r = plane_request(W, a, pos, bLo, bHi, cLo, cHi);
	double s[2];
	plane_sums(W, &r, 1, s);
	return plane_finish(W, &r, s[0], s[1]);
}

float (poyntingX)(world W, float x, float yMin, float yMax, float zMin, float zMax, ...) { return poynting_plane(W, 0, x, yMin, yMax, zMin, zMax); }
float (poyntingY)(world W, float y, float xMin, float xMax, float zMin, float zMax, ...) { return poynting_plane(W, 1, y, zMin, zMax, xMin, xMax); }
float (poyntingZ)(world W, float z, float xMin, float xMax, float yMin, float yMax, ...) { return poynting_plane(W, 2, z, xMin, xMax, yMin, yMax); }

/* net outward flux through the faces of a box; faces on Bloch-periodic boundaries cancel
 * (reference output.c:335-362).  All faces go to the device in one request. */
float poyntingOut(world W, float xMin, float xMax, float yMin, float yMax, float zMin, float zMax)
{
	ALIS_PROF("poyntingOut");
	const float lo[3] = { xMin, yMin, zMin }, hi[3] = { xMax, yMax, zMax };
	const float wlo[3] = { W->xMin, W->yMin, W->zMin }, whi[3] = { W->xMax, W->yMax, W->zMax };
	const int sur[3] = { W->xMinSur, W->yMinSur, W->zMinSur };
	plane_req req[6];
	int n = 0, first[3] = { -1, -1, -1 }, second[3] = { -1, -1, -1 };
	for (int a=0; a<3; a++) {
		if (!(lo[a] > wlo[a] || hi[a] < whi[a] || sur[a] != BBC)) continue;
		const int b = (a+1)%3, c = (a+2)%3;
		first[a] = n; req[n++] = plane_request(W, a, hi[a], lo[b], hi[b], lo[c], hi[c]);
		if (!(lo[a] < 0 && sur[a]==SYM)) { second[a] = n; req[n++] = plane_request(W, a, lo[a], lo[b], hi[b], lo[c], hi[c]); }
	}
	double sums[12];
	if (n) plane_sums(W, req, n, sums);
	float f[3] = { 0, 0, 0 };
	for (int a=0; a<3; a++) if (first[a] >= 0) {
		f[a] = plane_finish(W, &req[first[a]], sums[2*first[a]], sums[2*first[a]+1]);
		if (second[a] < 0) f[a] *= 2;
		else f[a] -= plane_finish(W, &req[second[a]], sums[2*second[a]], sums[2*second[a]+1]);
	}
	return f[0] + f[1] + f[2];
}

/* ---- absorption inside a dispersive object (reference output.c:23-47, 365-435) ---------------------------
 * integral of E.J over the object's bounding box with trapezoidal end corrections, J = conduction
 * + polarisation currents scaled back by eps */
static int same_object(object A, object B)
{
	if (A.S != B.S || A.p != B.p) return 0;
	for (int a=0; a<3; a++) for (int b=0; b<3; b++) if (A.f[a][b] != B.f[a][b]) return 0;
	return 1;
}

static int needs_aux(matter M)
{
	for (int a=0; a<3; a++) for (int b=0; b<6; b++) if (M.d[a][b]) return 2;
	if (M.P[0].w || M.P[0].r) return 1;
	return 0;
}

static float object_current(world W, int c, int n, int i, int j, int k)
{
	coeffs *C = W->CE;
	float a = c==0 ? C->ax[n] : c==1 ? C->ay[n] : C->az[n];
	float ***fill = c==0 ? C->Jx[n] : c==1 ? C->Jy[n] : C->Jz[n];
	float ***E = c==0 ? W->E->x : c==1 ? W->E->y : W->E->z;
	float ***inv = c==0 ? C->x : c==1 ? C->y : C->z;
	float ****J = c==0 ? W->E->Jx[n] : c==1 ? W->E->Jy[n] : W->E->Jz[n];
	float f = 0;
	if (a) f += 2 * fill[i][j][k] * E[i][j][k] / (1 + 1 / a);
	for (int p=0; p<C->NK[n]; p++) if (J[p]) f += J[p][i][j][k];
	return f / inv[i][j][k];
}

/* the three z sums of E.J of one column (reference output.c:375-404, before the x,y weights) */
static void column_sums_host(world W, int n, int i, int j, float out[3])
{
	coeffs *C = W->CE;
	float xs = 0, ys = 0, zs = 0;
	const int k0 = MAX(W->kMIN, C->kMin[n]), k1 = MIN(W->kMAX, C->kMax[n]);

	for (int k=k0; k<=k1; k++) {
		if (C->x[i][j][k]) xs += C->dz[k] * W->E->x[i][j][k] * object_current(W, 0, n, i, j, k);
		if (C->y[i][j][k]) ys += C->dz[k] * W->E->y[i][j][k] * object_current(W, 1, n, i, j, k);
		if (C->z[i][j][k]) zs += W->CH->dz[k] * W->E->z[i][j][k] * object_current(W, 2, n, i, j, k);
	}
	if (k0 == W->kMIN) {
		const int k = W->kMIN;
		if (C->x[i][j][k]) xs -= C->dz[k] * W->E->x[i][j][k] * object_current(W, 0, n, i, j, k) * 0.5;
		if (C->y[i][j][k]) ys -= C->dz[k] * W->E->y[i][j][k] * object_current(W, 1, n, i, j, k) * 0.5;
		if (C->z[i][j][k]) zs -= W->CH->dz[k] * W->E->z[i][j][k] * object_current(W, 2, n, i, j, k);
	}
	if (k1 == W->kMAX) {
		const int k = W->kMAX;
		if (C->x[i][j][k]) xs -= C->dz[k] * W->E->x[i][j][k] * object_current(W, 0, n, i, j, k) * 0.5;
		if (C->y[i][j][k]) ys -= C->dz[k] * W->E->y[i][j][k] * object_current(W, 1, n, i, j, k) * 0.5;
	}
	out[0] = xs; out[1] = ys; out[2] = zs;
}

/* columns of the object's box: from the device (alisb_absorption_columns: three floats per column
 * come back instead of the E, J and material arrays), or from the host views */
typedef struct { world W; int n, i0, j0, nj; float *dev; } abs_cols;

static float column_absorption(const abs_cols *A, int i, int j, float x, float y, float z)
{
	world W = A->W;
	coeffs *C = W->CE;
	float s[3], Sum = 0;
	if (A->dev) memcpy(s, A->dev + 3*((size_t)(i-A->i0)*A->nj + (j-A->j0)), sizeof s);
	else column_sums_host(W, A->n, i, j, s);
	Sum += x * s[0] * W->CH->dx[i] * C->dy[j];
	Sum += y * s[1] * C->dx[i] * W->CH->dy[j];
	Sum += z * s[2] * C->dx[i] * C->dy[j];
	return Sum;
}

float objectAbsorption(world W, object O)
{
	ALIS_PROF("objectAbsorption");
	coeffs *C = W->CE;
	float Sum = 0;
	int n = 0;
	for (int o=0; !same_object(O, C->O[o]); o++) if (needs_aux(C->M[o])) n++;
	const int i0 = MAX(W->iMIN, C->iMin[n]), i1 = MIN(W->iMAX, C->iMax[n]);
	const int j0 = MAX(W->jMIN, C->jMin[n]), j1 = MIN(W->jMAX, C->jMax[n]);
	const int k0 = MAX(W->kMIN, C->kMin[n]), k1 = MIN(W->kMAX, C->kMax[n]);
	abs_cols A = { W, n, i0, j0, j1-j0+1, 0 };
	if (i0 > i1 || j0 > j1 || k0 > k1) return 0;
	if (!alis_host_monitors()) {
		alis_bridge *B = alis_host_bridge(W);
		alis_bridge_commit(B);
		const int32_t box[6] = { i0-(W->iMIN-1), i1-(W->iMIN-1), j0-(W->jMIN-1), j1-(W->jMIN-1), k0-(W->kMIN-1), k1-(W->kMIN-1) };
		A.dev = xmalloc(sizeof(float)*3*(size_t)(i1-i0+1)*(j1-j0+1));
		if (alisb_absorption_columns(alis_bridge_engine(B), n, box, A.dev) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
	}
	else {
		alis_view_need(W, i0, i1, j0, j1, k0, k1);
		alis_view_need_ade(W);
	}

	for (int i=i0; i<=i1; i++) {
		for (int j=j0; j<=j1; j++) Sum += column_absorption(&A, i, j, 1, 1, 1);
		if (j0 == W->jMIN) Sum += column_absorption(&A, i, W->jMIN,-0.5,-1,-0.5);
		if (j1 == W->jMAX) Sum += column_absorption(&A, i, W->jMAX,-0.5, 0,-0.5);
	}
	if (i0 == W->iMIN) {                         /* (separate partial sums: see plane_flux) */
		float part = 0;
		for (int j=j0; j<=j1; j++) part += column_absorption(&A, W->iMIN, j,-1,-0.5,-0.5);
		Sum += part;
		if (j0 == W->jMIN) Sum += column_absorption(&A, W->iMIN, W->jMIN, 0.5, 0.5, 0.25);
		if (j1 == W->jMAX) Sum += column_absorption(&A, W->iMIN, W->jMAX, 0.5, 0.0, 0.25);
	}
	if (i1 == W->iMAX) {
		float part = 0;
		for (int j=j0; j<=j1; j++) part += column_absorption(&A, W->iMAX, j, 0,-0.5,-0.5);
		Sum += part;
		if (j0 == W->jMIN) Sum += column_absorption(&A, W->iMAX, W->jMIN, 0.0, 0.5, 0.25);
		if (j1 == W->jMAX) Sum += column_absorption(&A, W->iMAX, W->jMAX, 0.0, 0.0, 0.25);
	}

	free(A.dev);

	if (W->Dom.x.Min==W->Dom.x.Max) Sum /= W->CH->dx[W->iMax]+W->CH->dx[W->iMax-1];
	if (W->Dom.y.Min==W->Dom.y.Max) Sum /= W->CH->dy[W->jMax]+W->CH->dy[W->jMax-1];
	if (W->Dom.z.Min==W->Dom.z.Max) Sum /= W->CH->dz[W->kMax]+W->CH->dz[W->kMax-1];
	if (W->xMinSur==SYM) Sum *= 2;
	if (W->yMinSur==SYM) Sum *= 2;
	if (W->zMinSur==SYM) Sum *= 2;
	return Sum / W->dt;
}

/* shell helper used by createWorld and drivers (reference output.c:438-445) */
void exec(char *format, ...)
{
	char str[1024];
	va_list ap;
	va_start(ap, format);
	vsprintf(str, format, ap);
	va_end(ap);
	if (system(str)) { /* exit status ignored, as in the reference */ }
}
