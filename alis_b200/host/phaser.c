/*
 * phaser.c -- running DFT ("phasers") of the tangential fields on the faces of a box, per
 * frequency; Poynting flux from the phasors; spectrum writers.
 *
 * The accumulators live on the GPU (alisb_phaser_*, kernel phaser_face): updatePhaser computes the
 * two phase factors per frequency on the host (same cexpf as the reference) and launches ONE
 * kernel per face for all frequencies.  The host arrays P->zMax.Ex[i][j][k] ... are views that
 * are refreshed lazily: library routines call alisSyncPhaser; user code that indexes the arrays
 * directly (reference examples/Cross_Polarization_T.c:3-19) is served by page protection -- while
 * the device holds newer data the arrays are PROT_NONE, the first touch faults, the handler pulls
 * the accumulators from the device and the access is retried (the scheme of managed memory).
 * Definitions follow reference src/phaser.c (cited per function).
 */
#define _GNU_SOURCE
#include <sys/stat.h>
#include <sys/mman.h>
#include <signal.h>
#include <unistd.h>
#include <pthread.h>
#include <string.h>
#include "alis_c.h"
#include "host_internal.h"

static alisb_engine *engine_of(world W) { return alis_bridge_engine(alis_host_bridge(W)); }
static phaser base_of(phaser P) { return P - P->b200_index; }

/* ---- guarded host views -------------------------------------------------------------------------------
 * All host views of a phaser list live in ONE shared-memory block that is mapped twice:
 *   lo     the user mapping, the one P->zMax.Ex[i][j][k] points into.  PROT_NONE while the device holds
 *          newer data, readable/writable otherwise.
 *   alias  a second mapping of the same pages, always writable, never handed out.
 * A refresh (alisSyncPhaser, called by library routines or by the fault handler) fills the pages through
 * `alias` under guard_lock and unprotects `lo` only AFTER the data is complete and b200_stale is cleared:
 * a second thread that touches the arrays meanwhile (reference examples/Cross_Polarization_T.c:3-19 reads
 * them inside `#pragma omp parallel for`) still faults, blocks on the lock in its own handler, finds the
 * list fresh and retries -- no thread can ever see a half-filled or stale page.
 * The handler calls the CUDA runtime (a device-to-host copy), which is not async-signal-safe in general;
 * it is entered only synchronously, by a load/store of user code on these pages, never from inside the
 * runtime or libc.  ALIS_PHASER_GUARD=0 turns the protection off: the arrays are then plain memory and
 * valid only after an explicit alisSyncPhaser(P) (the conservative path; include/alis_api.h). */
typedef struct guard { char *lo, *alias; size_t bytes; phaser base; struct guard *next; } guard;
static guard *guards;
static struct sigaction chained;
static int handler_installed;
static pthread_mutex_t guard_lock = PTHREAD_MUTEX_INITIALIZER;

static int guard_enabled(void)
{
	static int mode = -1;
	if (mode < 0) { const char *s = getenv("ALIS_PHASER_GUARD"); mode = s ? atoi(s) != 0 : 1; }
	return mode;
}

static guard *guard_of(phaser base)
{
	for (guard *g = guards; g; g = g->next) if (g->base == base) return g;
	return 0;
}

static void guard_protect(phaser base, int prot)
{
	guard *g = guard_of(base);
	if (g) mprotect(g->lo, g->bytes, prot);
}

static void guard_fault(int sig, siginfo_t *si, void *ctx)
{
	const char *addr = (const char*)si->si_addr;
	for (guard *g = guards; g; g = g->next) if (g->lo <= addr && addr < g->lo + g->bytes) {
		alisSyncPhaser(g->base);        /* takes guard_lock; returns with the pages complete and accessible */
		return;
	}
	/* not ours: hand over to whoever was installed before (or the default action) */
	if ((chained.sa_flags & SA_SIGINFO) && chained.sa_sigaction) { chained.sa_sigaction(sig, si, ctx); return; }
	if (!(chained.sa_flags & SA_SIGINFO) && chained.sa_handler != SIG_DFL && chained.sa_handler != SIG_IGN) { chained.sa_handler(sig); return; }
	signal(sig, SIG_DFL);
}

/* the doubly mapped block for all host views of a list, registered for protection as a single range
 * (one mprotect per transition, however many frequencies the list has); 0 when the guard is off or
 * the kernel refuses the shared mapping (the views are then ordinary memory) */
static char *guard_arena(phaser base, size_t bytes)
{
	if (!guard_enabled() || !bytes) return 0;
	const size_t page = (size_t)sysconf(_SC_PAGESIZE);
	bytes = (bytes + page - 1)/page*page;
	const int fd = memfd_create("alis-phasor-views", MFD_CLOEXEC);
	if (fd < 0) return 0;
	char *lo = MAP_FAILED, *alias = MAP_FAILED;
	if (ftruncate(fd, (off_t)bytes) == 0) {
		lo = mmap(0, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
		alias = mmap(0, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	}
	close(fd);
	if (lo == MAP_FAILED || alias == MAP_FAILED) {
		if (lo != MAP_FAILED) munmap(lo, bytes);
		if (alias != MAP_FAILED) munmap(alias, bytes);
		return 0;
	}
	guard *g = xmalloc(sizeof *g);
	g->lo = lo; g->alias = alias; g->bytes = bytes; g->base = base;
	pthread_mutex_lock(&guard_lock);
	g->next = guards; guards = g;
	if (!handler_installed) {
		struct sigaction sa;
		memset(&sa, 0, sizeof sa);
		sa.sa_sigaction = guard_fault;
		sa.sa_flags = SA_SIGINFO | SA_NODEFER;
		sigemptyset(&sa.sa_mask);
		sigaction(SIGSEGV, &sa, &chained);
		handler_installed = 1;
	}
	pthread_mutex_unlock(&guard_lock);
	return lo;
}

/* makeComplexField over a given data block (or its own when data == 0) */
static float complex ***field_over(float complex *data, int iMin, int iMax, int jMin, int jMax, int kMin, int kMax)
{
	if (!data) return makeComplexField(iMin, iMax, jMin, jMax, kMin, kMax);
	const long ni = 1-iMin+iMax, nj = 1-jMin+jMax, nk = 1-kMin+kMax;
	float complex ***f;
	MALLOC(f, iMin, ni);
	MALLOC(f[iMin], jMin, ni*nj);
	f[iMin][jMin] = data - kMin;
	for (int i=iMin; i<=iMax; i++) {
		f[i] = f[iMin] + nj*(i-iMin);
		for (int j=jMin; j<=jMax; j++) f[i][j] = f[iMin][jMin] + nk*(nj*(i-iMin)+(j-jMin));
	}
	return f;
}
#define VIEW_BYTES(b) ((sizeof(float complex)*(size_t)((b)[1]-(b)[0]+1)*((b)[3]-(b)[2]+1)*((b)[5]-(b)[4]+1) + 63)/64*64)

static void guard_release(phaser base)
{
	pthread_mutex_lock(&guard_lock);
	for (guard **g = &guards; *g; ) {
		if ((*g)->base == base) { guard *dead = *g; *g = dead->next; munmap(dead->lo, dead->bytes); munmap(dead->alias, dead->bytes); free(dead); }
		else g = &(*g)->next;
	}
	pthread_mutex_unlock(&guard_lock);
}

/* face order used with the engine: 0 xMin, 1 xMax, 2 yMin, 3 yMax, 4 zMin, 5 zMax; the four
 * accumulators of a face are E(c+1), E(c+2), H(c+1), H(c+2) for face normal c */
static cfield *face_of(phaser P, int face)
{
	cfield *F[6] = { &P->xMin, &P->xMax, &P->yMin, &P->yMax, &P->zMin, &P->zMax };
	return F[face];
}

static float complex ****slot_of(cfield *F, int axis, int which)
{
	float complex ****E[3] = { &F->Ex, &F->Ey, &F->Ez }, ****H[3] = { &F->Hx, &F->Hy, &F->Hz };
	const int c = (axis + 1 + (which & 1)) % 3;
	return which < 2 ? E[c] : H[c];
}

static void face_box(phaser P, int face, int b[6])
{
	b[0] = P->iMin, b[1] = P->iMax, b[2] = P->jMin, b[3] = P->jMax, b[4] = P->kMin, b[5] = P->kMax;
	const int axis = face/2;
	b[2*axis] = b[2*axis+1] = (face & 1) ? b[2*axis+1] : b[2*axis];
}

/* reference phaser.c:21-66 */
phaser (createPhasers)(world W, float LMin, float LMax, float dL, float xMin, float xMax, float yMin, float yMax, float zMin, float zMax, ...)
{
	int nMin = floor(LMin/dL), nMax = ceil(LMax/dL);
	phaser P;
	/* the finite source whose duration defines W->N (searched in the TFSF auxiliary world too) */
	source S = W->SE && W->N==W->SE->N ? W->SE : W->SH && W->N==W->SH->N ? W->SH : 0;
	for (world w = W; !S && w->S && w->S->W; S = w->SE && W->N==w->SE->N ? w->SE : w->SH && W->N==w->SH->N ? w->SH : 0) w = W->S->W;
	const int count = 1-nMin+nMax;
	MALLOC(P, 0, count+1);                       /* zeroed terminator element at the end */

	int mask = 0;
	for (int n=0; n<count; n++) {
		phaser p = &P[n];
		p->W = W;
		p->N = nMin * nMax * dL / W->dt;
		p->f = (nMax-n) / (dL * nMin * nMax);
		p->A = 2.0 / (P->W->N ? P->W->N : P->N);
		p->iMin = MAX(W->iMin, W->xtoi(W, xMin)), p->iMax = MIN(W->iMax, W->xtoi(W, xMax));
		p->jMin = MAX(W->jMin, W->ytoj(W, yMin)), p->jMax = MIN(W->jMax, W->ytoj(W, yMax));
		p->kMin = MAX(W->kMin, W->ztok(W, zMin)), p->kMax = MIN(W->kMax, W->ztok(W, zMax));
		p->b200_index = n, p->b200_count = count;

		/* normalisation by the source spectrum at this frequency */
		if (S && W->N) {
			float complex c = 0;
			float complex iwdt = 2 * PI * I * p->f * W->dt;
			for (int q=1; q<W->N; q++) c += S->waveForm(W, S, W->dt*q, S->phase) * cexpf(iwdt*q);
			p->A = 1 / cabsf(c);
			if (W->sourceType == 2) p->A /= sqrtf(p->f * S->wavelength);
			if (W->sourceType == 3) p->A /= p->f * S->wavelength;
		}

		/* a monitor box wider than a periodic / symmetric world stands for xNUM x yNUM replicas */
		if ((xMax-xMin < INF && xMin < W->xMin && W->xMax < xMax) && (W->xMaxSur==SYM || W->xMaxSur==PBC || W->xMaxSur==BBC)) {
			p->xNUM = floor((xMax - xMin) / (W->xMax - W->xMin) + 0.5);
			p->expikx = (W->xMaxSur==BBC) ? (W->coskx+I*W->sinkx) : 1;
		}
		if ((yMax-yMin < INF && yMin < W->yMin && W->yMax < yMax) && (W->yMaxSur==SYM || W->yMaxSur==PBC || W->yMaxSur==BBC)) {
			p->yNUM = floor((yMax - yMin) / (W->yMax - W->yMin) + 0.5);
			p->expiky = (W->yMaxSur==BBC) ? (W->cosky+I*W->sinky) : 1;
		}

		/* faces exist only where the world face is absorbing */
		const int on[6] = { W->xMinSur==PML && p->iMin != p->iMax, W->xMaxSur==PML,
		                    W->yMinSur==PML && p->jMin != p->jMax, W->yMaxSur==PML,
		                    W->zMinSur==PML && p->kMin != p->kMax, W->zMaxSur==PML };
		mask = 0;
		for (int face=0; face<6; face++) if (on[face]) mask |= 1<<face;
	}

	/* host views of the accumulators, all in one guarded arena */
	if (mask) {
		size_t per = 0;
		for (int face=0; face<6; face++) if (mask & (1<<face)) { int b[6]; face_box(P, face, b); per += 4*VIEW_BYTES(b); }
		char *cur = guard_arena(P, per*count);
		P->b200_arena = cur;
		for (int n=0; n<count; n++) for (int face=0; face<6; face++) if (mask & (1<<face)) {
			int b[6];
			face_box(&P[n], face, b);
			for (int which=0; which<4; which++) {
				*slot_of(face_of(&P[n], face), face/2, which) = field_over((float complex*)cur, b[0], b[1], b[2], b[3], b[4], b[5]);
				if (cur) cur += VIEW_BYTES(b);
			}
		}
	}

	/* one device accumulator set for the whole frequency list */
	if (mask) {
		int32_t box[6] = { P->iMin-(W->iMIN-1), P->iMax-(W->iMIN-1), P->jMin-(W->jMIN-1), P->jMax-(W->jMIN-1), P->kMin-(W->kMIN-1), P->kMax-(W->kMIN-1) };
		alis_phasers *dev = 0;
		if (alis_phasers_create(alis_bridge_shards(alis_host_bridge(W)), box, mask, count, &dev) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
		P->b200_dev = dev;
	}
	return P;
}

phaser (createPhaser)(world W, float L, float xMin, float xMax, float yMin, float yMax, float zMin, float zMax, ...)
{
	return (createPhasers)(W, L, L, L, xMin, xMax, yMin, yMax, zMin, zMax);
}

/* the list is ordered by decreasing frequency; pick the entry closest to 1/wavelength
 * (reference phaser.c:75-80) */
phaser choosePhaser(phaser P, float wavelength)
{
	do P++;
	while (1/wavelength < P->f);
	return (1/wavelength - P->f < (P-1)->f - 1/wavelength) ? P : P-1;
}

/* reference phaser.c:94-110: one accumulation of every frequency from P to the end of its list */
void updatePhaser(world W, phaser P)
{
	ALIS_PROF("updatePhaser");
	phaser B = base_of(P);
	if (!B->b200_dev) return;
	const int count = B->b200_count;
	float complex *wE = xcalloc(count, sizeof *wE), *wH = xcalloc(count, sizeof *wH);
	for (phaser p = P; p->f; p++) {
		wE[p->b200_index] = cexpf(2 * PI * I * p->f * W->E->t);
		wH[p->b200_index] = cexpf(2 * PI * I * p->f * W->H->t);
	}
	if (alis_phasers_accumulate(B->b200_dev, (const float*)wE, (const float*)wH) != ALISB_OK) {
		printf("alis-b200: %s\n", alisb_last_error());
		exit(1);
	}
	if (!B->b200_stale) {                        /* the device is ahead of the host views now */
		pthread_mutex_lock(&guard_lock);
		B->b200_stale = 1;
		guard_protect(B, PROT_NONE);
		pthread_mutex_unlock(&guard_lock);
	}
	free(wE); free(wH);
}

/* device accumulator set and frequency index of one list element (far-field sums, farfld.c) */
alis_phasers *alis_phaser_device(phaser P, int *freq)
{
	if (freq) *freq = P->b200_index;
	return (alis_phasers*)base_of(P)->b200_dev;
}

/* refresh the host views of the whole frequency list */
void alisSyncPhaser(phaser P)
{
	phaser B = base_of(P);
	if (!B->b200_dev) return;
	pthread_mutex_lock(&guard_lock);
	if (B->b200_stale) {
		/* fill through the always-writable alias mapping; the user mapping opens only when all is there */
		const guard *g = guard_of(B);
		const ptrdiff_t shift = g ? g->alias - g->lo : 0;
		for (phaser p = B; p->f; p++) for (int face=0; face<6; face++) for (int which=0; which<4; which++) {
			float complex ***A = *slot_of(face_of(p, face), face/2, which);
			if (!A) continue;
			int b[6];
			face_box(p, face, b);
			/* the row-pointer tables are ordinary memory; only the data block is guarded */
			float *dst = (float*)((char*)(A[b[0]][b[2]] + b[4]) + shift);
			if (alis_phasers_read(B->b200_dev, p->b200_index, face, which, dst) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
		}
		__sync_synchronize();
		B->b200_stale = 0;
		guard_protect(B, PROT_READ | PROT_WRITE);
	}
	pthread_mutex_unlock(&guard_lock);
}

/* reference phaser.c:122-139 */
void deletePhaser(phaser P)
{
	phaser B = base_of(P);
	guard_release(B);
	if (B->b200_dev) alis_phasers_destroy(B->b200_dev);
	for (phaser p = B; p->f; p++) {
		for (int face=0; face<6; face++) for (int which=0; which<4; which++) {
			float complex ****slot = slot_of(face_of(p, face), face/2, which);
			if (*slot) {
				int b[6];
				face_box(p, face, b);
				float complex ***A = *slot;
				if (B->b200_arena) { free(&A[b[0]][b[2]]); free(&A[b[0]]); }     /* the data block belongs to the arena */
				else removeField(A, b[0], b[2], b[4]);
				*slot = 0;
			}
		}
		if (p->dE) free(&p->dE[p->kMIN]);
		if (p->dH) free(&p->dH[p->kMIN]);
		if (p->ex) free(&p->ex[p->kMIN]);
		if (p->ez) free(&p->ez[p->kMIN]);
	}
	free(B);                                     /* the arena's two mappings went with guard_release */
}

/* ---- flux of the phasors through the monitor box (reference phaser.c:142-245) ---------------------
 * Re(E x H*) integrated with trapezoidal weights; generic in the face normal a, b = a+1, c = a+2. */
static inline float complex cpick(float complex ***A, int a, int at, int b, int ib, int c, int ic)
{
	int q[3];
	q[a] = at; q[b] = ib; q[c] = ic;
	return A[q[0]][q[1]][q[2]];
}

static float face_flux(phaser P, cfield S, int a, int at)
{
	world W = P->W;
	const int b = (a+1)%3, c = (a+2)%3;
	float complex ***E[3] = { S.Ex, S.Ey, S.Ez }, ***H[3] = { S.Hx, S.Hy, S.Hz };
	const float *dE[3] = { W->CE->dx, W->CE->dy, W->CE->dz }, *dH[3] = { W->CH->dx, W->CH->dy, W->CH->dz };
	const int lo[3] = { P->iMin, P->jMin, P->kMin }, hi[3] = { P->iMax, P->jMax, P->kMax };
	float Sum = 0, Sum2 = 0;                     /* two partial sums, added at the end (as the reference's two reduction loops) */

	if (E[b] && H[c]) {
		for (int ib=lo[b]+1; ib<=hi[b]; ib++) {
			float sum = 0;
			sum += 0.5 * dE[c][lo[c]] * cpick(E[b],a,at,b,ib,c,lo[c]) * conjf(cpick(H[c],a,at,b,ib,c,lo[c]));
			sum += 0.5 * dE[c][hi[c]] * cpick(E[b],a,at,b,ib,c,hi[c]) * conjf(cpick(H[c],a,at,b,ib,c,hi[c]));
			for (int ic=lo[c]+1; ic<hi[c]; ic++) sum += dE[c][ic] * cpick(E[b],a,at,b,ib,c,ic) * conjf(cpick(H[c],a,at,b,ib,c,ic));
			Sum += dH[b][ib] * sum;
		}
	}
	if (E[c] && H[b]) {
		for (int ic=lo[c]+1; ic<=hi[c]; ic++) {
			float sum = 0;
			sum -= 0.5 * dE[b][lo[b]] * cpick(E[c],a,at,b,lo[b],c,ic) * conjf(cpick(H[b],a,at,b,lo[b],c,ic));
			sum -= 0.5 * dE[b][hi[b]] * cpick(E[c],a,at,b,hi[b],c,ic) * conjf(cpick(H[b],a,at,b,hi[b],c,ic));
			for (int ib=lo[b]+1; ib<hi[b]; ib++) sum -= dE[b][ib] * cpick(E[c],a,at,b,ib,c,ic) * conjf(cpick(H[b],a,at,b,ib,c,ic));
			Sum2 += dH[c][ic] * sum;
		}
	}
	Sum += Sum2;
	return Sum;
}

float phaserPoyntingOut(phaser P)
{
	world W = P->W;
	float Sum = 0;
	alisSyncPhaser(P);

	Sum += face_flux(P, P->xMax, 0, P->iMax) - face_flux(P, P->xMin, 0, P->iMin);
	Sum += face_flux(P, P->yMax, 1, P->jMax) - face_flux(P, P->yMin, 1, P->jMin);
	Sum += face_flux(P, P->zMax, 2, P->kMax) - face_flux(P, P->zMin, 2, P->kMin);

	if (W->Dom.x.Min==W->Dom.x.Max) Sum /= W->CH->dx[P->iMax]+W->CH->dx[P->iMax-1];
	if (W->Dom.y.Min==W->Dom.y.Max) Sum /= W->CH->dy[P->jMax]+W->CH->dy[P->jMax-1];
	if (W->Dom.z.Min==W->Dom.z.Max) Sum /= W->CH->dz[P->kMax]+W->CH->dz[P->kMax-1];

	if (W->xMinSur==SYM) Sum *= 2;
	if (W->yMinSur==SYM) Sum *= 2;
	if (W->zMinSur==SYM) Sum *= 2;

	return Sum * sq(P->A);
}

float phaserPoynting(phaser P)   { return fabsf(phaserPoyntingOut(P)); }
float phaserPoyntingIn(phaser P) { return -phaserPoyntingOut(P); }

/* ---- spectra of several phaser lists side by side (reference phaser.c:248-295) ------------------------ */
void (writePoyntingSpectrum)(world W, char *format, ...)
{
	int M = 0;
	char str[1024], filename[1024];
	va_list ap;
	phaser *L;

	va_start(ap, format);
	vsprintf(filename, format, ap);
	for (phaser p=va_arg(ap, phaser); p; M++) p=va_arg(ap, phaser);
	va_end(ap);
	va_start(ap, format);
	vsprintf(filename, format, ap);
	MALLOC(L, 0, M);
	for (int m=0; m<M; m++) L[m] = va_arg(ap, phaser);
	va_end(ap);
	for (int n=0; L[0][n].f; n++) {
		int s = sprintf(str, "%g", 1/L[0][n].f);
		for (int m=0; m<M; m++) s += sprintf(str+s, "\t%g", phaserPoynting(L[m]+n));
		sprintf(str+s, "\n");
		(writeTxt)(W, "%s", filename, "%s", str, "");
	}
	free(L);
}

void (writeFarFieldSpectrum)(world W, tfield pol(phaser,float,float), float theta, float phi, char *format, ...)
{
	int M = 0;
	char str[1024], filename[1024];
	va_list ap;
	phaser *L;

	va_start(ap, format);
	vsprintf(filename, format, ap);
	for (phaser p=va_arg(ap, phaser); p; M++) p=va_arg(ap, phaser);
	va_end(ap);
	va_start(ap, format);
	vsprintf(filename, format, ap);
	MALLOC(L, 0, M);
	for (int m=0; m<M; m++) L[m] = va_arg(ap, phaser);
	va_end(ap);
	for (int n=0; L[0][n].f; n++) {
		int s = sprintf(str, "%g", 1/L[0][n].f);
		for (int m=0; m<M; m++) s += sprintf(str+s, "\t%g", farFieldI(L[m]+n, pol, theta, phi));
		sprintf(str+s, "\n");
		(writeTxt)(W, "%s", filename, "%s", str, "");
	}
	free(L);
}
