/*
 * farfld.c -- near-to-far-field transform from the phasor faces.
 *
 * Reciprocity formulation of the reference (src/farfld.c): the far field towards (theta, phi) is
 * the overlap of the recorded surface phasors with a "test field" -- the plane wave arriving from
 * that direction, propagated through the layered background (eps(z) column under the monitor
 * box, including the dispersive poles at the phasor frequency) by a 1-D finite-difference
 * transfer recursion.  The x,y dependence of the test field is a plane-wave factor, so every face
 * integral is separable: a weighted sum with 1-D weight vectors.
 *
 * The face integrals run ON THE DEVICE: the phasor faces stay where
 * updatePhaser accumulated them, the host prepares the 1-D weight vectors and test fields of every
 * direction with the reference's expressions and libm (cited per function), and all directions of
 * a sweep (farFieldProfile, farFieldTheta/Phi, farFieldFlux) go to the GPU as one batched launch
 * per 4096 queries (alisb_phaser_surface_sums).
 */
#include <sys/stat.h>
#include <sys/time.h>
#include <string.h>
#include "alis_c.h"
#include "host_internal.h"

/* ---- projection maps: pixel (x,y) of an X x Y image -> (theta, phi) in degrees, theta = INF when
 * the pixel is outside the map (reference farfld.c:14-41) --------------------------------------------- */
void CylindericalMap(int x, int y, int X, int Y, float angle[2])
{
	angle[0] = 180.0 - 180.0 * y / Y;               /* polar angle: rows from the north pole down */
	angle[1] = 360.0 * x / X;                       /* azimuth: columns */
}

/* azimuthal equidistant chart of one hemisphere: radius from the image centre = polar distance from
 * that hemisphere's pole, 90 degrees at the rim; pixels outside the disc get theta = INF */
static void hemisphere_chart(int x, int y, int X, int Y, int southern, float angle[2])
{
	const float radius = 180 * sqrtf(sq(x/(X-1.0)-0.5) + sq(y/(Y-1.0)-0.5));
	angle[1] = 180 * atan2f(y/(Y-1.0)-0.5, x/(X-1.0)-0.5) / PI;
	if (southern) { angle[0] = 180 - radius; if (angle[0] < 90) angle[0] = INF; }
	else          { angle[0] = radius;       if (angle[0] > 90) angle[0] = INF; }
}

void NorthernAzimuthalMap(int x, int y, int X, int Y, float angle[2]) { hemisphere_chart(x, y, X, Y, 0, angle); }
void SouthernAzimuthalMap(int x, int y, int X, int Y, float angle[2]) { hemisphere_chart(x, y, X, Y, 1, angle); }

void AzimuthalMap(int x, int y, int X, int Y, float angle[2])      /* two hemispheres side by side */
{
	const int half = (X+1)/2;
	if (x < half) hemisphere_chart(x, y, half, Y, 0, angle);
	if (x > X/2) hemisphere_chart(x-X/2, y, half, Y, 1, angle);
}

/* direction (a,b) given relative to an axis tilted by theta, back in the global frame */
static float tilt_theta(float theta, float a, float b)
{
	return acosf(cosf(a*PI/180)*cosf(theta*PI/180)-(sinf(a*PI/180)*cosf(b*PI/180))*sinf(theta*PI/180))*180/PI;
}

static float tilt_phi(float theta, float a, float b)
{
	return atan2f(sinf(a*PI/180)*sinf(b*PI/180), (sinf(a*PI/180)*cosf(b*PI/180))*cosf(theta*PI/180)+cosf(a*PI/180)*sinf(theta*PI/180))*180/PI;
}

/* ---- layered background under the monitor box (reference farfld.c:52-117) ------------------------------ */
static int needs_aux(matter M)
{
	for (int a=0; a<3; a++) for (int b=0; b<6; b++) if (M.d[a][b]) return 2;
	if (M.P[0].w || M.P[0].r) return 1;
	return 0;
}

/* complex susceptibility of pole p at photon energy w (conductivity poles included) */
static float complex pole_chi(matter M, float w, int p)
{
	if (M.P[p].f || M.P[p].g)
		return 0.5
			*((M.P[p].f * M.P[p].w - M.P[p].g * M.P[p].r + I * M.P[p].g * M.P[p].w) / (M.P[p].w + w + I * M.P[p].r)
			+ (M.P[p].f * M.P[p].w - M.P[p].g * M.P[p].r - I * M.P[p].g * M.P[p].w) / (M.P[p].w - w - I * M.P[p].r));
	if (M.P[p].w) return - sq(M.P[p].w) / (sq(w) + I * w * M.P[p].r);
	return -I * M.P[p].r / w;
}

/* dE, dH: cell sizes along z; ex, ez: complex eps_x(z), eps_z(z) at the phasor frequency, read at
 * the (iMax, jMax) corner of the box; mirrored below z = 0 on a symmetric world */
static void build_layer_column(phaser P, world W)
{
	coeffs *C = W->CE;
	P->kMIN = (W->zMinSur==SYM) ? -P->kMax : P->kMin;
	const int len = 2-P->kMIN+P->kMax;
	MALLOC(P->dE, P->kMIN, len);
	MALLOC(P->dH, P->kMIN, len);
	MALLOC(P->ex, P->kMIN, len);
	MALLOC(P->ez, P->kMIN, len);
	const int ic = P->iMax, jc = P->jMax;

	for (int k=P->kMin; k<=P->kMax+1; k++) {
		P->dE[k] = 1 / C->dtdz[k];
		P->dH[k] = 1 / W->CH->dtdz[k];
		P->ex[k] = C->x ? 1/C->x[ic][jc][k] : 1;
		P->ez[k] = C->z ? 1/C->z[ic][jc][k] : 1;

		for (int n=0, o=0; n<C->N; n++, o++) {
			while (!needs_aux(C->M[o])) o++;
			if (!(C->iMin[n]<=ic && ic<=C->iMax[n] && C->jMin[n]<=jc && jc<=C->jMax[n])) continue;
			if (!(C->kMin[n]<=k && k<=C->kMax[n])) continue;
			matter M = C->M[o];
			int tri = M.e.y || M.e.z ? 1 : 0;
			P->ex[k] += -C->ax[n] * (M.e.x ? M.e.x : M.e.x) * C->Jx[n][ic][jc][k];
			P->ez[k] += -C->az[n] * (M.e.z ? M.e.z : M.e.x) * C->Jz[n][ic][jc][k];
			for (int p=0; p<MAXPOLES; p+=1+2*tri) {
				P->ex[k] += pole_chi(M, W->eV * P->f, p+0*tri) * C->Jx[n][ic][jc][k];
				P->ez[k] += pole_chi(M, W->eV * P->f, p+2*tri) * C->Jz[n][ic][jc][k];
			}
		}
	}
	if (W->zMinSur==SYM) {
		for (int k=P->kMIN; k<P->kMin; k++) {
			P->dE[k] = P->dE[-k];
			P->dH[k] = P->dH[1-k];
			P->ex[k] = P->ex[-k];
			P->ez[k] = P->ez[1-k];
		}
	}
}

/* ---- test fields (reference farfld.c:120-245) -------------------------------------------------------------
 * Unit plane wave incident from the far side, marched through the column with the discrete
 * (leap-frog) dispersion of the grid: s = sin(pi f dt) is the discrete frequency.  T.A converts
 * the overlap integral into the far-field amplitude (transmission normalisation). */
tfield Unpol(phaser P, float theta, float phi)
{
	(void)P; (void)theta; (void)phi;
	return (tfield) {0};
}

tfield Ppol(phaser P, float theta, float phi)
{
	tfield T = {0};
	float sinth = 0, sin2th = 0, sinphi = sinf(phi*PI/180), cosphi = cosf(phi*PI/180);
	const int lo = P->kMIN, hi = P->kMax, len = 2-P->kMIN+P->kMax;

	MALLOC(T.Ex, lo, len);
	MALLOC(T.Hy, lo, len);

	if (theta < 90 || 270 < theta) {             /* upward: start below, march up */
		sinth = sinf(theta*PI/180) * sqrtf(crealf(P->ex[hi])), sin2th = sq(sinth);
		T.Hy[lo] = -1;
		T.Ex[lo] = csqrtf((1-sin2th/P->ez[lo]) / P->ex[lo]);
		T.Ex[lo] *= cexpf(-I * casinf(sinf(PI*P->f*P->W->dt) * csqrtf(P->ex[lo]-sin2th) * P->dE[lo]));

		for (int k=lo; k<=hi; k++) {
			T.Hy[k+1] = T.Hy[k] + 2 * I * sinf(PI*P->f*P->W->dt) * P->dE[k] * T.Ex[k] * P->ex[k];
			T.Ex[k+1] = T.Ex[k] + 2 * I * sinf(PI*P->f*P->W->dt) * P->dH[k+1] * T.Hy[k+1] * (1-sin2th/P->ez[k+1]);
			if (cabsf(T.Hy[k+1])>1E10) for (int kk=lo; kk<=k+1; kk++) T.Ex[kk]/=1E10, T.Hy[kk]/=1E10;
		}
		T.A = sqrtf(sqrtf(P->ex[hi]));
		T.A /= cabsf(0.5*(T.Hy[hi]+T.Hy[hi+1]) / sqrtf(P->ex[hi]) - T.Ex[hi] / cosf(theta*PI/180));
	}

	if (90 < theta && theta < 270) {             /* downward: start above, march down */
		sinth = sinf(theta*PI/180) * sqrtf(crealf(P->ex[lo])), sin2th = sq(sinth);
		T.Hy[hi+1] = 1;
		T.Ex[hi+1] = csqrtf((1-sin2th/P->ez[hi+1]) / P->ex[hi+1]);
		T.Ex[hi+1] *= cexpf(I * casinf(sinf(PI*P->f*P->W->dt) * csqrtf(P->ex[hi+1]-sin2th) * P->dE[hi+1]));

		for (int k=hi; k>=lo; k--) {
			T.Ex[k] = T.Ex[k+1] - 2 * I * sinf(PI*P->f*P->W->dt) * P->dH[k+1] * T.Hy[k+1] * (1-sin2th/P->ez[k+1]);
			T.Hy[k] = T.Hy[k+1] - 2 * I * sinf(PI*P->f*P->W->dt) * P->dE[k] * T.Ex[k] * P->ex[k];
			if (cabsf(T.Hy[k+1])>1E10) for (int kk=hi; kk>=k; kk--) T.Ex[kk]/=1E10, T.Hy[kk]/=1E10;
		}
		T.A = sqrtf(sqrtf(P->ex[lo]));
		T.A /= cabsf(0.5*(T.Hy[lo]+T.Hy[lo+1]) / sqrtf(P->ex[lo]) - T.Ex[lo] / cosf(theta*PI/180));
	}

	if (theta) {                                 /* longitudinal E of the oblique wave */
		MALLOC(T.Ez, lo, len);
		for (int k=lo; k<=hi+1; k++) T.Ez[k] = T.Hy[k] * sinth/P->ez[k];
	}
	if (phi) {                                   /* rotate the transverse pair into (x,y) */
		MALLOC(T.Hx, lo, len);
		MALLOC(T.Ey, lo, len);
		for (int k=lo; k<=hi+1; k++) {
			T.Ey[k] = T.Ex[k] * sinphi;
			T.Ex[k] = T.Ex[k] * cosphi;
			T.Hx[k] =-T.Hy[k] * sinphi;
			T.Hy[k] = T.Hy[k] * cosphi;
		}
	}
	return T;
}

tfield Spol(phaser P, float theta, float phi)
{
	tfield T = {0};
	float sinth = 0, sin2th = 0, sinphi = sinf(phi*PI/180), cosphi = cosf(phi*PI/180);
	const int lo = P->kMIN, hi = P->kMax, len = 2-P->kMIN+P->kMax;

	MALLOC(T.Hx, lo, len);
	MALLOC(T.Ey, lo, len);

	if (theta < 90 || 270 < theta) {
		sinth = sinf(theta*PI/180) * sqrtf(crealf(P->ex[hi])), sin2th = sq(sinth);
		T.Hx[lo] = csqrtf(P->ex[lo] - sin2th);
		T.Ey[lo] = cexpf(-I * casinf(sinf(PI*P->f*P->W->dt) * csqrtf(P->ex[lo]-sin2th) * P->dE[lo]));

		for (int k=lo; k<=hi; k++) {
			T.Hx[k+1] = T.Hx[k] - 2 * I * sinf(PI*P->f*P->W->dt) * P->dE[k] * T.Ey[k] * (P->ex[k] - sin2th);
			T.Ey[k+1] = T.Ey[k] - 2 * I * sinf(PI*P->f*P->W->dt) * P->dH[k+1] * T.Hx[k+1];
			if (cabsf(T.Hx[k+1])>1E10) for (int kk=lo; kk<=k+1; kk++) T.Ey[kk]/=1E10, T.Hx[kk]/=1E10;
		}
		T.A = sqrtf(sqrtf(P->ex[hi]));
		T.A /= cabsf(T.Ey[hi] + 0.5*(T.Hx[hi]+T.Hx[hi+1]) / (sqrtf(P->ex[hi]) * cosf(theta*PI/180)));
	}

	if (90 < theta && theta < 270) {
		sinth = sinf(theta*PI/180) * sqrtf(crealf(P->ex[lo])), sin2th = sq(sinth);
		T.Ey[hi+1] = cexpf(I * casinf(sinf(PI*P->f*P->W->dt) * csqrtf(P->ex[hi+1]-sin2th) * P->dE[hi+1]));
		T.Hx[hi+1] = -csqrtf(P->ex[hi+1] - sin2th);

		for (int k=hi; k>=lo; k--) {
			T.Ey[k] = T.Ey[k+1] + 2 * I * sinf(PI*P->f*P->W->dt) * P->dH[k+1] * T.Hx[k+1];
			T.Hx[k] = T.Hx[k+1] + 2 * I * sinf(PI*P->f*P->W->dt) * P->dE[k] * T.Ey[k] * (P->ex[k] - sin2th);
			if (cabsf(T.Hx[k+1])>1E10) for (int kk=hi; kk>=k; kk--) T.Ey[kk]/=1E10, T.Hx[kk]/=1E10;
		}
		T.A = sqrtf(sqrtf(P->ex[lo]));
		T.A /= cabsf(T.Ey[lo] + 0.5*(T.Hx[lo]+T.Hx[lo+1]) / (sqrtf(P->ex[lo]) * cosf(theta*PI/180)));
	}

	if (theta) {
		MALLOC(T.Hz, lo, len);
		for (int k=lo; k<=hi+1; k++) T.Hz[k] = -T.Ey[k] * sinth;
	}
	if (phi) {
		MALLOC(T.Ex, lo, len);
		MALLOC(T.Hy, lo, len);
		for (int k=lo; k<=hi+1; k++) {
			T.Ex[k] =-T.Ey[k] * sinphi;
			T.Ey[k] = T.Ey[k] * cosphi;
			T.Hy[k] = T.Hx[k] * sinphi;
			T.Hx[k] = T.Hx[k] * cosphi;
		}
	}
	return T;
}

/* ---- separable face integrals (reference farfld.c:248-459) --------------------------------------------------
 * weight vectors:
 *   node-type along a transverse axis:  d_E[q] exp(i kq x_q)        mirror term  -conj(.) cos
 *   cell-type (staggered half a cell):  d_H[q] exp(i kq x_{q-1/2})  mirror term  +conj(.) cos
 *   along z they are the test-field samples times the cell size (node-type uses T[-k], cell-type
 *   T[1-k] for the mirror image).
 * surface sum: outer index runs lo+1..hi with cell-type weights, inner index is a trapezoid over
 * lo..hi with node-type weights.
 *
 * The double loops themselves run ON THE DEVICE (alisb_phaser_surface_sums, kernel surface_sums):
 * the phasor faces stay where updatePhaser accumulated them.  Every query (direction x
 * polarisation) is evaluated twice by the same code: a RECORD pass that builds the weight vectors
 * and lists the sums it needs (ff_job), and -- after one device call for all queries of a sweep --
 * a REPLAY pass in which surface_sum() hands back the device results in the same order. */
typedef struct { float complex *w; int lo, hi; } wvec;

typedef struct { int face, which; int lo_vec, hi_vec; } ff_job;       /* indices into ff_ctx.vec */
typedef struct { float complex *w; int len; const float complex *key; } ff_vec;
/* plane-wave weight vectors depend on (axis, staggering, range, k) only: the faces and the two
 * polarisations of one direction share them */
typedef struct { int axis, cell_type, lo, hi; float kq; wvec v; } ff_cached;
typedef struct {
	int replay;
	/* record */
	ff_job *job; int njob, capjob;
	ff_vec *vec; int nvec, capvec;
	ff_cached cache[8]; int ncache;
	/* test fields: built in the record pass, consumed (and released) by the replay pass */
	tfield *T; int nT, capT, iT;
	/* replay */
	const double *result; int next;
} ff_ctx;

static void ctx_free(ff_ctx *c)
{
	for (int q=0; q<c->nvec; q++) free(c->vec[q].w);
	for (int q=0; q<c->ncache; q++) free(&c->cache[q].v.w[c->cache[q].v.lo]);
	free(c->vec); free(c->job);
	tfield *T = c->T; const int nT = c->nT;          /* handed over to the replay pass */
	memset(c, 0, sizeof *c);
	c->T = T; c->nT = nT;
}

static wvec vec_alloc(ff_ctx *c, int lo, int hi)
{
	wvec v = { 0, lo, hi };
	if (!c->replay) MALLOC(v.w, lo, 1-lo+hi);
	return v;
}
/* a released vector's address may be handed out again by malloc: forget it as a sharing key */
static void vec_free(ff_ctx *c, wvec v)
{
	if (!v.w) return;
	for (int q=0; q<c->ncache; q++) if (c->cache[q].v.w == v.w) return;      /* owned by the cache until ctx_free */
	for (int q=0; q<c->nvec; q++) if (c->vec[q].key == v.w) c->vec[q].key = 0;
	free(&v.w[v.lo]);
}

static wvec plane_wave_weights(ff_ctx *c, world W, int axis, int lo, int hi, float kq, int cell_type)
{
	const float *d = cell_type ? (axis==0 ? W->CH->dx : W->CH->dy) : (axis==0 ? W->CE->dx : W->CE->dy);
	float (*coord)(world, float) = axis==0 ? W->itox : W->jtoy;
	const int sym = (axis==0 ? W->xMinSur : W->yMinSur) == SYM;
	const float cosk = axis==0 ? W->coskx : W->cosky;
	if (!c->replay) for (int q=0; q<c->ncache; q++) {
		const ff_cached *e = &c->cache[q];
		if (e->axis == axis && e->cell_type == cell_type && e->lo == lo && e->hi == hi && e->kq == kq) return e->v;
	}
	wvec v = vec_alloc(c, lo, hi);
	if (c->replay) return v;
	for (int q = cell_type ? lo+1 : lo; q<=hi; q++) {
		v.w[q] = d[q] * cexpf(I*kq*coord(W, cell_type ? q-0.5 : q));
		if (sym) { if (cell_type) v.w[q] += conjf(v.w[q]) * cosk; else v.w[q] -= conjf(v.w[q]) * cosk; }
	}
	if (c->ncache < 8) { ff_cached *e = &c->cache[c->ncache++]; e->axis = axis; e->cell_type = cell_type; e->lo = lo; e->hi = hi; e->kq = kq; e->v = v; }
	return v;
}

static wvec test_field_weights(ff_ctx *c, world W, int lo, int hi, const float complex *T, int cell_type)
{
	const float *d = cell_type ? W->CH->dz : W->CE->dz;
	const int sym = W->zMinSur == SYM;
	wvec v = vec_alloc(c, lo, hi);
	if (c->replay) return v;
	for (int k = cell_type ? lo+1 : lo; k<=hi; k++) {
		v.w[k] = d[k] * T[k];
		if (sym) { if (cell_type) v.w[k] += d[k] * T[1-k] * W->coskz; else v.w[k] -= d[k] * T[-k] * W->coskz; }
	}
	return v;
}

/* the device sums over the whole face box with these weights: the outer (cell-type) vector gets a
 * zero at lo, the inner (node-type) vector the trapezoid's halves at both ends */
static int ctx_vector(ff_ctx *c, wvec v, int outer)
{
	for (int q=0; q<c->nvec; q++) if (c->vec[q].key == v.w) return q;
	if (c->nvec == c->capvec) { c->capvec = c->capvec ? 2*c->capvec : 16; c->vec = xrealloc(c->vec, sizeof(ff_vec)*c->capvec); }
	ff_vec *f = &c->vec[c->nvec];
	f->len = v.hi - v.lo + 1; f->key = v.w;
	f->w = xmalloc(sizeof(float complex)*f->len);
	for (int q=v.lo; q<=v.hi; q++) {
		if (outer) f->w[q-v.lo] = q == v.lo ? 0 : v.w[q];
		else f->w[q-v.lo] = ((q==v.lo ? 0.5f : 0) + (q==v.hi ? 0.5f : 0) + (q!=v.lo && q!=v.hi ? 1.0f : 0)) * v.w[q];
	}
	return c->nvec++;
}

/* face `face` (normal a = face/2); field component `comp` of E (isH 0) or H (isH 1); outer axis o
 * (cell-type weights), inner axis n (node-type weights) */
static float complex surface_sum(ff_ctx *c, int face, int isH, int comp, int o, wvec Wo, int n, wvec Wn)
{
	if (c->replay) { const double *r = c->result + 2*(size_t)c->next++; return (float)r[0] + I*(float)r[1]; }
	const int a = face/2;
	if (c->njob == c->capjob) { c->capjob = c->capjob ? 2*c->capjob : 16; c->job = xrealloc(c->job, sizeof(ff_job)*c->capjob); }
	ff_job *j = &c->job[c->njob++];
	j->face = face; j->which = (comp - a + 3)%3 - 1 + (isH ? 2 : 0);
	const int vo = ctx_vector(c, Wo, 1), vn = ctx_vector(c, Wn, 0);
	if (o < n) j->lo_vec = vo, j->hi_vec = vn; else j->lo_vec = vn, j->hi_vec = vo;
	return 0;
}

static float complex face_x(ff_ctx *c, phaser P, int face, int i, tfield T, float kx, float ky)
{
	world W = P->W;
	float complex EyHz=0, EzHy=0, HyEz=0, HzEy=0, Sum=0;

	if (T.Hz || T.Ey) {
		wvec Y = plane_wave_weights(c, W, 1, P->jMin, P->jMax, ky, 1);
		if (T.Hz) { wvec Z = test_field_weights(c, W, P->kMin, P->kMax, T.Hz, 0); EyHz = surface_sum(c, face, 0, 1, 1, Y, 2, Z); vec_free(c, Z); }
		if (T.Ey) { wvec Z = test_field_weights(c, W, P->kMin, P->kMax, T.Ey, 0); HzEy = surface_sum(c, face, 1, 2, 1, Y, 2, Z); vec_free(c, Z); }
		vec_free(c, Y);
	}
	if (T.Hy || T.Ez) {
		wvec Y = plane_wave_weights(c, W, 1, P->jMin, P->jMax, ky, 0);
		if (T.Hy) { wvec Z = test_field_weights(c, W, P->kMin, P->kMax, T.Hy, 1); EzHy = surface_sum(c, face, 0, 2, 2, Z, 1, Y); vec_free(c, Z); }
		if (T.Ez) { wvec Z = test_field_weights(c, W, P->kMin, P->kMax, T.Ez, 1); HyEz = surface_sum(c, face, 1, 1, 2, Z, 1, Y); vec_free(c, Z); }
		vec_free(c, Y);
	}
	Sum += (EyHz-HzEy-EzHy+HyEz) * cexpf(I*kx*W->itox(W,i));
	if (W->xMinSur==SYM) Sum += (EyHz+HzEy-EzHy-HyEz) * cexpf(-I*kx*W->itox(W,i)) * W->coskx;
	return Sum;
}

static float complex face_y(ff_ctx *c, phaser P, int face, int j, tfield T, float kx, float ky)
{
	world W = P->W;
	float complex ExHz=0, EzHx=0, HxEz=0, HzEx=0, Sum=0;

	if (T.Hx || T.Ez) {
		wvec X = plane_wave_weights(c, W, 0, P->iMin, P->iMax, kx, 0);
		if (T.Hx) { wvec Z = test_field_weights(c, W, P->kMin, P->kMax, T.Hx, 1); EzHx = surface_sum(c, face, 0, 2, 2, Z, 0, X); vec_free(c, Z); }
		if (T.Ez) { wvec Z = test_field_weights(c, W, P->kMin, P->kMax, T.Ez, 1); HxEz = surface_sum(c, face, 1, 0, 2, Z, 0, X); vec_free(c, Z); }
		vec_free(c, X);
	}
	if (T.Hz || T.Ex) {
		wvec X = plane_wave_weights(c, W, 0, P->iMin, P->iMax, kx, 1);
		if (T.Hz) { wvec Z = test_field_weights(c, W, P->kMin, P->kMax, T.Hz, 0); ExHz = surface_sum(c, face, 0, 0, 0, X, 2, Z); vec_free(c, Z); }
		if (T.Ex) { wvec Z = test_field_weights(c, W, P->kMin, P->kMax, T.Ex, 0); HzEx = surface_sum(c, face, 1, 2, 0, X, 2, Z); vec_free(c, Z); }
		vec_free(c, X);
	}
	Sum += (EzHx-HxEz-ExHz+HzEx) * cexpf(I*ky*W->jtoy(W,j));
	if (W->yMinSur==SYM) Sum += (EzHx+HxEz-ExHz-HzEx) * cexpf(-I*ky*W->jtoy(W,j)) * W->cosky;
	return Sum;
}

static float complex face_z(ff_ctx *c, phaser P, int face, int k, tfield T, float kx, float ky)
{
	world W = P->W;
	float complex ExHy=0, EyHx=0, HxEy=0, HyEx=0, Sum=0;
	const int sym = W->zMinSur==SYM;

	if (T.Hy || T.Ex) {
		wvec X = plane_wave_weights(c, W, 0, P->iMin, P->iMax, kx, 1), Y = plane_wave_weights(c, W, 1, P->jMin, P->jMax, ky, 0);
		if (T.Hy) {
			ExHy = surface_sum(c, face, 0, 0, 0, X, 1, Y);
			Sum += ExHy * 0.5 * (T.Hy[k]+T.Hy[k+1]);
			if (sym) Sum += ExHy * 0.5 * (T.Hy[1-k]+T.Hy[-k]) * W->coskz;
		}
		if (T.Ex) {
			HyEx = surface_sum(c, face, 1, 1, 0, X, 1, Y);
			Sum -= HyEx * T.Ex[k];
			if (sym) Sum += HyEx * T.Ex[-k] * W->coskz;
		}
		vec_free(c, X); vec_free(c, Y);
	}
	if (T.Hx || T.Ey) {
		wvec X = plane_wave_weights(c, W, 0, P->iMin, P->iMax, kx, 0), Y = plane_wave_weights(c, W, 1, P->jMin, P->jMax, ky, 1);
		if (T.Hx) {
			EyHx = surface_sum(c, face, 0, 1, 1, Y, 0, X);
			Sum -= EyHx * 0.5 * (T.Hx[k]+T.Hx[k+1]);
			if (sym) Sum -= EyHx * 0.5 * (T.Hx[1-k]+T.Hx[-k]) * W->coskz;
		}
		if (T.Ey) {
			HxEy = surface_sum(c, face, 1, 0, 1, Y, 0, X);
			Sum += HxEy * T.Ey[k];
			if (sym) Sum -= HxEy * T.Ey[-k] * W->coskz;
		}
		vec_free(c, X); vec_free(c, Y);
	}
	return Sum;
}

/* ---- the transform (reference farfld.c:462-514) ------------------------------------------------------------- */
static void ensure_ready(phaser P)
{
	if (!P->dE) build_layer_column(P, P->W);
}

/* one query; called once in record mode and once in replay mode with identical control flow */
static float complex far_field_eval(ff_ctx *c, phaser P, tfield pol(phaser,float,float), float theta, float phi)
{
	float sinth = 0, kx, ky;
	float complex Sum = 0;
	world W = P->W;

	if (theta == INF || phi == INF) return 0;
	while (theta < 0) theta += 360;
	while (theta >= 360) theta -= 360;
	while (phi < 0) phi += 360;
	while (phi >= 360) phi -= 360;
	/* grazing directions: average the two neighbours */
	if ( 89 < theta && theta <  91) return 0.5 * (far_field_eval(c, P, pol,  89, phi) + far_field_eval(c, P, pol,  91, phi));
	if (269 < theta && theta < 271) return 0.5 * (far_field_eval(c, P, pol, 269, phi) + far_field_eval(c, P, pol, 271, phi));
	/* no far field inside an absorbing half space */
	if (theta < 90 || 270 < theta) {
		if (cimagf(P->ex[P->kMax])) return 0;
		else sinth = sinf(theta*PI/180) * sqrtf(crealf(P->ex[P->kMax]));
	}
	if (90 < theta && theta < 270) {
		if (cimagf(P->ex[P->kMIN])) return 0;
		else sinth = sinf(theta*PI/180) * sqrtf(crealf(P->ex[P->kMIN]));
	}

	tfield T;
	if (c->replay && c->iT < c->nT) T = c->T[c->iT++];             /* the record pass's test field */
	else {
		T = pol(P, theta, phi);
		if (!c->replay) {
			if (c->nT == c->capT) { c->capT = c->capT ? 2*c->capT : 4; c->T = xrealloc(c->T, sizeof(tfield)*c->capT); }
			c->T[c->nT++] = T;
		}
	}
	kx = - 2 * PI * P->f * sinth * cosf(phi*PI/180);
	ky = - 2 * PI * P->f * sinth * sinf(phi*PI/180);

	if (W->xMinSur==PML && P->xMin.Ey) Sum -= face_x(c, P, 0, P->iMin, T, kx, ky);
	if (W->xMaxSur==PML) Sum += face_x(c, P, 1, P->iMax, T, kx, ky);
	if (W->yMinSur==PML && P->yMin.Ex) Sum -= face_y(c, P, 2, P->jMin, T, kx, ky);
	if (W->yMaxSur==PML) Sum += face_y(c, P, 3, P->jMax, T, kx, ky);
	if (W->zMinSur==PML && P->zMin.Ex) Sum -= face_z(c, P, 4, P->kMin, T, kx, ky);
	if (W->zMaxSur==PML) Sum += face_z(c, P, 5, P->kMax, T, kx, ky);

	if (c->replay) {                                               /* in the record pass the context keeps T for the replay */
		float complex *parts[6] = { T.Ex, T.Ey, T.Ez, T.Hx, T.Hy, T.Hz };
		for (int q=0; q<6; q++) if (parts[q]) free(&parts[q][P->kMIN]);
	}

	/* array factor of xNUM x yNUM periodic replicas */
	if (P->xNUM) {
		float complex s = 0, q = 1, p = P->expikx * cexpf(I*kx*(W->xMax-W->xMin));
		for (int x=0; x<P->xNUM; x++) s += (q *= p);
		Sum *= s / P->xNUM;
	}
	if (P->yNUM) {
		float complex s = 0, q = 1, p = P->expiky * cexpf(I*ky*(W->yMax-W->yMin));
		for (int y=0; y<P->yNUM; y++) s += (q *= p);
		Sum *= s / P->yNUM;
	}
	return Sum * T.A * P->A * P->f;
}

/* n directions, each with one or two polarisations (P and S share the plane-wave weights) ->
 * out[q][k]; the surface sums of a chunk of directions are one device call */
#define FF_CHUNK 4096
typedef tfield (*polfn)(phaser, float, float);
typedef struct { polfn pol[2]; int npol; float theta, phi; } ff_query;

static double wall_seconds(void) { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6*tv.tv_usec; }

static void far_field_batch(phaser P, int n, const ff_query *Q, float complex (*out)[2])
{
	const int timing = getenv("ALIS_FF_TIMING") != 0;       /* prints where a sweep spends its time */
	double t_rec = 0, t_pack = 0, t_dev = 0, t_rep = 0, t0;
	size_t tot_jobs = 0, tot_w = 0;
	ensure_ready(P);
	int freq = 0;
	alis_phasers *dev = alis_phaser_device(P, &freq);
	if (!dev) { for (int q=0; q<n; q++) out[q][0] = out[q][1] = 0; return; }      /* no absorbing face: nothing was recorded */
	ff_ctx *ctx = xcalloc(FF_CHUNK, sizeof *ctx);
	size_t *job0 = xmalloc(sizeof(size_t)*(FF_CHUNK+1)), *w0 = xmalloc(sizeof(size_t)*(FF_CHUNK+1));
	for (int base=0; base<n; base+=FF_CHUNK) {
		const int m = MIN(FF_CHUNK, n-base);
		t0 = wall_seconds();
		#pragma omp parallel for schedule(dynamic, 16)
		for (int q=0; q<m; q++) for (int k=0; k<Q[base+q].npol; k++) far_field_eval(&ctx[q], P, Q[base+q].pol[k], Q[base+q].theta, Q[base+q].phi);
		t_rec += wall_seconds() - t0; t0 = wall_seconds();
		size_t njobs = 0, nw = 0;
		for (int q=0; q<m; q++) { job0[q] = njobs; w0[q] = nw; njobs += ctx[q].njob; for (int v=0; v<ctx[q].nvec; v++) nw += ctx[q].vec[v].len; }
		job0[m] = njobs; w0[m] = nw;
		if (nw > 0xFFFFFFFFu) { printf("alis-b200: far-field batch too large\n"); exit(1); }
		alisb_surface_job *jobs = xmalloc(sizeof *jobs*(njobs+1));
		float complex *weights = xmalloc(sizeof *weights*(nw+1));
		double *sums = xmalloc(sizeof(double)*2*(njobs+1));
		#pragma omp parallel for schedule(static)
		for (int q=0; q<m; q++) {
			size_t off[ctx[q].nvec+1], w = w0[q];
			for (int v=0; v<ctx[q].nvec; v++) { off[v] = w; memcpy(weights+w, ctx[q].vec[v].w, sizeof(float complex)*ctx[q].vec[v].len); w += ctx[q].vec[v].len; }
			for (int j=0; j<ctx[q].njob; j++) {
				alisb_surface_job *J = &jobs[job0[q]+j];
				J->face = ctx[q].job[j].face; J->which = ctx[q].job[j].which;
				J->w_lo = (uint32_t)off[ctx[q].job[j].lo_vec]; J->w_hi = (uint32_t)off[ctx[q].job[j].hi_vec];
			}
			ctx_free(&ctx[q]);
		}
		t_pack += wall_seconds() - t0; t0 = wall_seconds();
		if (alis_phasers_surface_sums(dev, freq, jobs, (int)njobs, (const float*)weights, nw, sums) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
		t_dev += wall_seconds() - t0; t0 = wall_seconds();
		tot_jobs += njobs; tot_w += nw;
		#pragma omp parallel for schedule(dynamic, 16)
		for (int q=0; q<m; q++) {
			ff_ctx r = { 0 };
			r.replay = 1; r.result = sums + 2*job0[q];
			r.T = ctx[q].T; r.nT = ctx[q].nT;
			out[base+q][0] = out[base+q][1] = 0;
			for (int k=0; k<Q[base+q].npol; k++) out[base+q][k] = far_field_eval(&r, P, Q[base+q].pol[k], Q[base+q].theta, Q[base+q].phi);
			free(ctx[q].T);
			ctx[q].T = 0; ctx[q].nT = 0;
		}
		t_rep += wall_seconds() - t0;
		free(jobs); free(weights); free(sums);
	}
	free(ctx); free(job0); free(w0);
	if (timing) fprintf(stderr, "[alis far field] %d directions, %zu surface sums, %zu weights: record %.3f s, pack %.3f s, engine %.3f s, replay %.3f s\n",
	                    n, tot_jobs, tot_w, t_rec, t_pack, t_dev, t_rep);
}

float complex farField(phaser P, tfield pol(phaser,float,float), float theta, float phi)
{
	float complex f[1][2];
	ff_query q = { { pol, 0 }, 1, theta, phi };
	far_field_batch(P, 1, &q, f);
	return f[0][0];
}

float farFieldI(phaser P, tfield pol(phaser,float,float), float theta, float phi)
{
	ALIS_PROF("farFieldI");
	if (pol == Unpol) return farFieldI(P, Ppol, theta, phi) + farFieldI(P, Spol, theta, phi);
	float complex f = farField(P, pol, theta, phi);
	return conjf(f) * f;
}

/* P- and S-polarised far field towards n directions */
static void far_field_pairs(phaser P, int n, const float *theta, const float *phi, float complex *p, float complex *s)
{
	ff_query *Q = xmalloc(sizeof(ff_query)*(n+1));
	float complex (*f)[2] = xmalloc(sizeof(float complex)*2*(n+1));
	for (int q=0; q<n; q++) { Q[q].pol[0] = Ppol; Q[q].pol[1] = Spol; Q[q].npol = 2; Q[q].theta = theta[q]; Q[q].phi = phi[q]; }
	far_field_batch(P, n, Q, f);
	for (int q=0; q<n; q++) { p[q] = f[q][0]; s[q] = f[q][1]; }
	free(Q); free(f);
}

/* ---- angular maps and cuts (reference farfld.c:525-662) -------------------------------------------------------- */
void farFieldProfile(world W, phaser P, projmap projMap, int X, int Y, ...)
{
	ALIS_PROF("farFieldProfile");
	char *format, fieldname[8] = "", filename[1024], fullname[1024];   /* png() looks at F[4] */
	va_list ap;
	va_start(ap, Y);
	method M = va_arg(ap, method);
	unsigned long int *cmap = (M == png) ? va_arg(ap, unsigned long int *) : 0;
	float Max = (M == png) ? va_arg(ap, double) : 1;
	float max = 0;
	format = va_arg(ap, char*);
	vsprintf(filename, format, ap);
	va_end(ap);
	strcat(strcpy(fullname, W->ID), filename);
	strcpy(fieldname, strstr(filename, "Log") ? "Log" : "");
	if (filename[0]=='/') mkdir(W->ID, 0755);
	if (fullname[strlen(fullname)-1]=='/') mkdir(fullname, 0755);

	slice S;
	MALLOC(S, 0, 1);
	*S = (struct slice) {0, 0, 0, X-1, 0, Y-1, 0, 0, 0, 0, X-1, X, 0, Y-1, Y};
	float ***map[7];                             /* X-pol, Y-pol, CL, CR, P, S, unpolarised */
	for (int q=0; q<7; q++) map[q] = makeField(S->iMIN, S->iMAX, S->jMIN, S->jMAX, S->kMIN, S->kMAX);

	printf("Performing Near-to-Far-Field Transformation.\n");
	/* every pixel's direction, both polarisations: one batched sweep on the device */
	float complex *pp = xcalloc((size_t)X*Y, sizeof *pp), *ss = xcalloc((size_t)X*Y, sizeof *ss);
	float *thetas = xcalloc((size_t)X*Y, sizeof *thetas), *phis = xcalloc((size_t)X*Y, sizeof *phis);
	for (int x=0; x<X; x++) for (int y=0; y<Y; y++) {
		float angle[2];
		projMap(x, y, X, Y, angle);
		thetas[(size_t)x*Y+y] = angle[0];
		phis[(size_t)x*Y+y] = angle[1];
	}
	far_field_pairs(P, X*Y, thetas, phis, pp, ss);
	free(thetas);
	for (int x=0; x<X; x++) for (int y=0; y<Y; y++) {
		const float complex p = pp[(size_t)x*Y+y], s = ss[(size_t)x*Y+y];
		const float a1 = phis[(size_t)x*Y+y];
		float v[7];
		v[0] = conjf(p*cosf(a1*PI/180)-s*sinf(a1*PI/180))*(p*cosf(a1*PI/180)-s*sinf(a1*PI/180));
		v[1] = conjf(s*cosf(a1*PI/180)+p*sinf(a1*PI/180))*(s*cosf(a1*PI/180)+p*sinf(a1*PI/180));
		v[2] = sq(crealf(p)+cimagf(s))+sq(cimagf(p)-crealf(s));
		v[3] = sq(crealf(p)-cimagf(s))+sq(cimagf(p)+crealf(s));
		v[4] = conjf(p)*p;
		v[5] = conjf(s)*s;
		if (fieldname[0]) {
			v[6] = log10f(conjf(p)*p+conjf(s)*s);
			for (int q=0; q<6; q++) v[q] = log10f(v[q]);
		}
		else v[6] = v[4]+v[5];
		for (int q=0; q<7; q++) map[q][0][x][y] = v[q];
		max = MAX(max, v[6]);
	}
	free(pp); free(ss); free(phis);

	if (Max <= 0) Max = max;
	static const char *suffix[7] = { "X-pol", "Y-pol", "CL-pol", "CR-pol", "P-pol", "S-pol", "Unpol" };
	for (int q=0; q<7; q++) {
		S->F = map[q];
		M(W, S, Max, fieldname, strcat(strcpy(filename, fullname), suffix[q]), cmap);
		if (q < 6) { float ***f = map[q]; removeField(f, 0, 0, 0); }
	}
	deleteSlice(S);                              /* frees the last map through S->F */
}

void farFieldTheta(world W, phaser P, float phi, char *format, ...)
{
	ALIS_PROF("farFieldTheta");
	char filename[1024];
	va_list ap;
	va_start(ap, format);
	vsprintf(filename, format, ap);
	va_end(ap);
	float th[360], ph[360];
	float complex p[360], s[360];
	int n = 0;
	for (float theta=0; theta<360 && n<360; theta+=1) th[n] = theta, ph[n] = phi, n++;
	far_field_pairs(P, n, th, ph, p, s);
	for (int q=0; q<n; q++) writeRow(W, "%s", filename, th[q], sq(cabsf(p[q]))+sq(cabsf(s[q])), sq(cabsf(p[q])), sq(cabsf(s[q])));
}

void farFieldPhi(world W, phaser P, float theta, char *format, ...)
{
	ALIS_PROF("farFieldPhi");
	char filename[1024];
	va_list ap;
	va_start(ap, format);
	vsprintf(filename, format, ap);
	va_end(ap);
	float th[360], ph[360];
	float complex p[360], s[360];
	int n = 0;
	for (float phi=0; phi<360 && n<360; phi+=1) th[n] = theta, ph[n] = phi, n++;
	far_field_pairs(P, n, th, ph, p, s);
	for (int q=0; q<n; q++) writeRow(W, "%s", filename, ph[q], sq(cabsf(p[q]))+sq(cabsf(s[q])), sq(cabsf(p[q])), sq(cabsf(s[q])));
}

/* power radiated into the cone aMin..aMax (degrees) around (theta, phi), integrated on rings of
 * width delta with 6 n samples on ring n (capped at 60 degrees) */
float (farFieldFlux)(phaser P, tfield pol(phaser,float,float), float theta, float phi, float aMin, float aMax, float delta, ...)
{
	ALIS_PROF("farFieldFlux");
	if (!delta) delta=1;
	if (!aMax) aMax=aMin, aMin=0;

	float Sum = 0;
	int nMin = floor(aMin/delta+0.5);
	int nMax = ceil(aMax/delta-0.5);

	if (!nMin) {
		Sum = 2 * PI * (cosf(aMin*PI/180) - cosf(MIN(aMax,0.5*delta)*PI/180)) * farFieldI(P,pol,theta,phi);
		aMin = 0.5 * delta;
		nMin++;
	}
	if (nMax) {
		/* all ring samples first (one batched sweep on the device), then the reference's sum in its order */
		int count = 0;
		for (int n=nMin; n<=nMax; n++) count += n<floor(60/delta+0.5) ? n*6 : floor(60/delta+0.5)*6;
		float *ta = xmalloc(sizeof(float)*(count+1)), *tb = xmalloc(sizeof(float)*(count+1));
		float complex *fp = xmalloc(sizeof(float complex)*(count+1)), *fs = xmalloc(sizeof(float complex)*(count+1));
		int q = 0;
		for (int n=nMin; n<=nMax; n++) {
			int N = n<floor(60/delta+0.5) ? n*6 : floor(60/delta+0.5)*6;
			for (int m=0; m<N; m++, q++) {
				ta[q] = tilt_theta(theta, delta*n, 360.0*m/N);
				tb[q] = tilt_phi(theta, delta*n, 360.0*m/N)+phi;
			}
		}
		far_field_pairs(P, count, ta, tb, fp, fs);
		q = 0;
		for (int n=nMin; n<=nMax; n++) {
			int N = n<floor(60/delta+0.5) ? n*6 : floor(60/delta+0.5)*6;
			float dth = cosf(MAX(aMin,delta*(n-0.5))*PI/180) - cosf(MIN(aMax,delta*(n+0.5))*PI/180);
			for (int m=0; m<N; m++, q++) {
				float a = ta[q];
				float b = tb[q];
				float c = tilt_phi(-a, theta, phi-b)+180-360.0*m/N;
				float complex p = fp[q];
				float complex s = fs[q];
				float pp = sq(cabsf((p * cosf(c*PI/180) - s * sinf(c*PI/180))));
				float ss = sq(cabsf((s * cosf(c*PI/180) + p * sinf(c*PI/180))));
				if (pol == Ppol) Sum += dth * (pp * (PI/N + sinf(PI/N)) + ss * (PI/N - sinf(PI/N)));
				if (pol == Spol) Sum += dth * (pp * (PI/N - sinf(PI/N)) + ss * (PI/N + sinf(PI/N)));
				if (pol == Unpol) Sum += 2 * PI/N * dth * (pp + ss);
			}
		}
		free(ta); free(tb); free(fp); free(fs);
	}
	return Sum;
}
