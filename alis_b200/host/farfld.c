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
 * The phasor faces are small (N^2 complex values per face), so they are pulled from the device
 * once (alisSyncPhaser) and the angular sweep runs on the host (device-side sweep: SURVEY.md 8
 * row f2).  farFieldProfile evaluates its pixels in parallel over the host cores.
 * Arithmetic follows the reference expression by expression (cited per function).
 */
#include <sys/stat.h>
#include <string.h>
#include "alis_c.h"
#include "host_internal.h"

/* ---- projection maps: pixel (x,y) of an X x Y image -> (theta, phi) in degrees, theta = INF when
 * the pixel is outside the map (reference farfld.c:14-41) --------------------------------------------- */
void CylindericalMap(int x, int y, int X, int Y, float angle[2])
{
	angle[0] = 180.0 - 180.0 * y / Y;
	angle[1] = 360.0 * x / X;
}

void NorthernAzimuthalMap(int x, int y, int X, int Y, float angle[2])
{
	angle[0] = 180 * sqrtf(sq(x/(X-1.0)-0.5) + sq(y/(Y-1.0)-0.5));
	angle[1] = 180 * atan2f(y/(Y-1.0)-0.5, x/(X-1.0)-0.5) / PI;
	if (angle[0] > 90) angle[0] = INF;
}

void SouthernAzimuthalMap(int x, int y, int X, int Y, float angle[2])
{
	angle[0] = 180 - 180 * sqrtf(sq(x/(X-1.0)-0.5) + sq(y/(Y-1.0)-0.5));
	angle[1] = 180 * atan2f(y/(Y-1.0)-0.5, x/(X-1.0)-0.5) / PI;
	if (angle[0] < 90) angle[0] = INF;
}

void AzimuthalMap(int x, int y, int X, int Y, float angle[2])      /* two hemispheres side by side */
{
	if (x < (X+1)/2) NorthernAzimuthalMap(x, y, (X+1)/2, Y, angle);
	if (x > X/2) SouthernAzimuthalMap(x-X/2, y, (X+1)/2, Y, angle);
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
 * lo..hi with node-type weights. */
typedef struct { float complex *w; int lo, hi; } wvec;

static wvec vec_alloc(int lo, int hi)
{
	wvec v = { 0, lo, hi };
	MALLOC(v.w, lo, 1-lo+hi);
	return v;
}
static void vec_free(wvec v) { free(&v.w[v.lo]); }

static wvec plane_wave_weights(world W, int axis, int lo, int hi, float kq, int cell_type)
{
	const float *d = cell_type ? (axis==0 ? W->CH->dx : W->CH->dy) : (axis==0 ? W->CE->dx : W->CE->dy);
	float (*coord)(world, float) = axis==0 ? W->itox : W->jtoy;
	const int sym = (axis==0 ? W->xMinSur : W->yMinSur) == SYM;
	const float cosk = axis==0 ? W->coskx : W->cosky;
	wvec v = vec_alloc(lo, hi);
	for (int q = cell_type ? lo+1 : lo; q<=hi; q++) {
		v.w[q] = d[q] * cexpf(I*kq*coord(W, cell_type ? q-0.5 : q));
		if (sym) { if (cell_type) v.w[q] += conjf(v.w[q]) * cosk; else v.w[q] -= conjf(v.w[q]) * cosk; }
	}
	return v;
}

static wvec test_field_weights(world W, int lo, int hi, const float complex *T, int cell_type)
{
	const float *d = cell_type ? W->CH->dz : W->CE->dz;
	const int sym = W->zMinSur == SYM;
	wvec v = vec_alloc(lo, hi);
	for (int k = cell_type ? lo+1 : lo; k<=hi; k++) {
		v.w[k] = d[k] * T[k];
		if (sym) { if (cell_type) v.w[k] += d[k] * T[1-k] * W->coskz; else v.w[k] -= d[k] * T[-k] * W->coskz; }
	}
	return v;
}

static inline float complex cget(float complex ***A, int a, int at, int b, int ib, int c, int ic)
{
	int q[3];
	q[a] = at; q[b] = ib; q[c] = ic;
	return A[q[0]][q[1]][q[2]];
}

/* face normal a at index `at`; outer axis o (cell-type weights), inner axis n (node-type weights) */
static float complex surface_sum(float complex ***A, int a, int at, int o, wvec Wo, int n, wvec Wn)
{
	float complex acc = 0;
	for (int io=Wo.lo+1; io<=Wo.hi; io++) {
		float complex sum = 0.5 * (cget(A,a,at,o,io,n,Wn.lo) * Wn.w[Wn.lo] + cget(A,a,at,o,io,n,Wn.hi) * Wn.w[Wn.hi]);
		for (int in=Wn.lo+1; in<Wn.hi; in++) sum += cget(A,a,at,o,io,n,in) * Wn.w[in];
		acc += Wo.w[io] * sum;
	}
	return acc;
}

static float complex face_x(phaser P, cfield S, int i, tfield T, float kx, float ky)
{
	world W = P->W;
	float complex EyHz=0, EzHy=0, HyEz=0, HzEy=0, Sum=0;

	if (T.Hz || T.Ey) {
		wvec Y = plane_wave_weights(W, 1, P->jMin, P->jMax, ky, 1);
		if (T.Hz) { wvec Z = test_field_weights(W, P->kMin, P->kMax, T.Hz, 0); EyHz = surface_sum(S.Ey, 0, i, 1, Y, 2, Z); vec_free(Z); }
		if (T.Ey) { wvec Z = test_field_weights(W, P->kMin, P->kMax, T.Ey, 0); HzEy = surface_sum(S.Hz, 0, i, 1, Y, 2, Z); vec_free(Z); }
		vec_free(Y);
	}
	if (T.Hy || T.Ez) {
		wvec Y = plane_wave_weights(W, 1, P->jMin, P->jMax, ky, 0);
		if (T.Hy) { wvec Z = test_field_weights(W, P->kMin, P->kMax, T.Hy, 1); EzHy = surface_sum(S.Ez, 0, i, 2, Z, 1, Y); vec_free(Z); }
		if (T.Ez) { wvec Z = test_field_weights(W, P->kMin, P->kMax, T.Ez, 1); HyEz = surface_sum(S.Hy, 0, i, 2, Z, 1, Y); vec_free(Z); }
		vec_free(Y);
	}
	Sum += (EyHz-HzEy-EzHy+HyEz) * cexpf(I*kx*W->itox(W,i));
	if (W->xMinSur==SYM) Sum += (EyHz+HzEy-EzHy-HyEz) * cexpf(-I*kx*W->itox(W,i)) * W->coskx;
	return Sum;
}

static float complex face_y(phaser P, cfield S, int j, tfield T, float kx, float ky)
{
	world W = P->W;
	float complex ExHz=0, EzHx=0, HxEz=0, HzEx=0, Sum=0;

	if (T.Hx || T.Ez) {
		wvec X = plane_wave_weights(W, 0, P->iMin, P->iMax, kx, 0);
		if (T.Hx) { wvec Z = test_field_weights(W, P->kMin, P->kMax, T.Hx, 1); EzHx = surface_sum(S.Ez, 1, j, 2, Z, 0, X); vec_free(Z); }
		if (T.Ez) { wvec Z = test_field_weights(W, P->kMin, P->kMax, T.Ez, 1); HxEz = surface_sum(S.Hx, 1, j, 2, Z, 0, X); vec_free(Z); }
		vec_free(X);
	}
	if (T.Hz || T.Ex) {
		wvec X = plane_wave_weights(W, 0, P->iMin, P->iMax, kx, 1);
		if (T.Hz) { wvec Z = test_field_weights(W, P->kMin, P->kMax, T.Hz, 0); ExHz = surface_sum(S.Ex, 1, j, 0, X, 2, Z); vec_free(Z); }
		if (T.Ex) { wvec Z = test_field_weights(W, P->kMin, P->kMax, T.Ex, 0); HzEx = surface_sum(S.Hz, 1, j, 0, X, 2, Z); vec_free(Z); }
		vec_free(X);
	}
	Sum += (EzHx-HxEz-ExHz+HzEx) * cexpf(I*ky*W->jtoy(W,j));
	if (W->yMinSur==SYM) Sum += (EzHx+HxEz-ExHz-HzEx) * cexpf(-I*ky*W->jtoy(W,j)) * W->cosky;
	return Sum;
}

static float complex face_z(phaser P, cfield S, int k, tfield T, float kx, float ky)
{
	world W = P->W;
	float complex ExHy=0, EyHx=0, HxEy=0, HyEx=0, Sum=0;
	const int sym = W->zMinSur==SYM;

	if (T.Hy || T.Ex) {
		wvec X = plane_wave_weights(W, 0, P->iMin, P->iMax, kx, 1), Y = plane_wave_weights(W, 1, P->jMin, P->jMax, ky, 0);
		if (T.Hy) {
			ExHy = surface_sum(S.Ex, 2, k, 0, X, 1, Y);
			Sum += ExHy * 0.5 * (T.Hy[k]+T.Hy[k+1]);
			if (sym) Sum += ExHy * 0.5 * (T.Hy[1-k]+T.Hy[-k]) * W->coskz;
		}
		if (T.Ex) {
			HyEx = surface_sum(S.Hy, 2, k, 0, X, 1, Y);
			Sum -= HyEx * T.Ex[k];
			if (sym) Sum += HyEx * T.Ex[-k] * W->coskz;
		}
		vec_free(X); vec_free(Y);
	}
	if (T.Hx || T.Ey) {
		wvec X = plane_wave_weights(W, 0, P->iMin, P->iMax, kx, 0), Y = plane_wave_weights(W, 1, P->jMin, P->jMax, ky, 1);
		if (T.Hx) {
			EyHx = surface_sum(S.Ey, 2, k, 1, Y, 0, X);
			Sum -= EyHx * 0.5 * (T.Hx[k]+T.Hx[k+1]);
			if (sym) Sum -= EyHx * 0.5 * (T.Hx[1-k]+T.Hx[-k]) * W->coskz;
		}
		if (T.Ey) {
			HxEy = surface_sum(S.Hx, 2, k, 1, Y, 0, X);
			Sum += HxEy * T.Ey[k];
			if (sym) Sum -= HxEy * T.Ey[-k] * W->coskz;
		}
		vec_free(X); vec_free(Y);
	}
	return Sum;
}

/* ---- the transform (reference farfld.c:462-514) ------------------------------------------------------------- */
static void ensure_ready(phaser P)
{
	alisSyncPhaser(P);
	if (!P->dE) build_layer_column(P, P->W);
}

float complex farField(phaser P, tfield pol(phaser,float,float), float theta, float phi)
{
	float sinth = 0, kx, ky;
	float complex Sum = 0;
	world W = P->W;

	ensure_ready(P);
	if (theta == INF || phi == INF) return 0;
	while (theta < 0) theta += 360;
	while (theta >= 360) theta -= 360;
	while (phi < 0) phi += 360;
	while (phi >= 360) phi -= 360;
	/* grazing directions: average the two neighbours */
	if ( 89 < theta && theta <  91) return 0.5 * (farField(P, pol,  89, phi) + farField(P, pol,  91, phi));
	if (269 < theta && theta < 271) return 0.5 * (farField(P, pol, 269, phi) + farField(P, pol, 271, phi));
	/* no far field inside an absorbing half space */
	if (theta < 90 || 270 < theta) {
		if (cimagf(P->ex[P->kMax])) return 0;
		else sinth = sinf(theta*PI/180) * sqrtf(crealf(P->ex[P->kMax]));
	}
	if (90 < theta && theta < 270) {
		if (cimagf(P->ex[P->kMIN])) return 0;
		else sinth = sinf(theta*PI/180) * sqrtf(crealf(P->ex[P->kMIN]));
	}

	tfield T = pol(P, theta, phi);
	kx = - 2 * PI * P->f * sinth * cosf(phi*PI/180);
	ky = - 2 * PI * P->f * sinth * sinf(phi*PI/180);

	if (W->xMinSur==PML) Sum -= face_x(P, P->xMin, P->iMin, T, kx, ky);
	if (W->xMaxSur==PML) Sum += face_x(P, P->xMax, P->iMax, T, kx, ky);
	if (W->yMinSur==PML) Sum -= face_y(P, P->yMin, P->jMin, T, kx, ky);
	if (W->yMaxSur==PML) Sum += face_y(P, P->yMax, P->jMax, T, kx, ky);
	if (W->zMinSur==PML) Sum -= face_z(P, P->zMin, P->kMin, T, kx, ky);
	if (W->zMaxSur==PML) Sum += face_z(P, P->zMax, P->kMax, T, kx, ky);

	float complex *parts[6] = { T.Ex, T.Ey, T.Ez, T.Hx, T.Hy, T.Hz };
	for (int q=0; q<6; q++) if (parts[q]) free(&parts[q][P->kMIN]);

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

float farFieldI(phaser P, tfield pol(phaser,float,float), float theta, float phi)
{
	if (pol == Unpol) return farFieldI(P, Ppol, theta, phi) + farFieldI(P, Spol, theta, phi);
	float complex f = farField(P, pol, theta, phi);
	return conjf(f) * f;
}

/* ---- angular maps and cuts (reference farfld.c:525-662) -------------------------------------------------------- */
void farFieldProfile(world W, phaser P, projmap projMap, int X, int Y, ...)
{
	char *format, fieldname[4], filename[1024], fullname[1024];
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
	ensure_ready(P);
	float complex *pp = calloc((size_t)X*Y, sizeof *pp), *ss = calloc((size_t)X*Y, sizeof *ss);
	float *phis = calloc((size_t)X*Y, sizeof *phis);
	#pragma omp parallel for collapse(2) schedule(dynamic, 16)
	for (int x=0; x<X; x++) for (int y=0; y<Y; y++) {
		float angle[2];
		projMap(x, y, X, Y, angle);
		pp[(size_t)x*Y+y] = farField(P, Ppol, angle[0], angle[1]);
		ss[(size_t)x*Y+y] = farField(P, Spol, angle[0], angle[1]);
		phis[(size_t)x*Y+y] = angle[1];
	}
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
	char filename[1024];
	va_list ap;
	va_start(ap, format);
	vsprintf(filename, format, ap);
	va_end(ap);
	for (float theta=0; theta<360; theta+=1) {
		float complex p = farField(P, Ppol, theta, phi);
		float complex s = farField(P, Spol, theta, phi);
		writeRow(W, "%s", filename, theta, sq(cabsf(p))+sq(cabsf(s)), sq(cabsf(p)), sq(cabsf(s)));
	}
}

void farFieldPhi(world W, phaser P, float theta, char *format, ...)
{
	char filename[1024];
	va_list ap;
	va_start(ap, format);
	vsprintf(filename, format, ap);
	va_end(ap);
	for (float phi=0; phi<360; phi+=1) {
		float complex p = farField(P, Ppol, theta, phi);
		float complex s = farField(P, Spol, theta, phi);
		writeRow(W, "%s", filename, phi, sq(cabsf(p))+sq(cabsf(s)), sq(cabsf(p)), sq(cabsf(s)));
	}
}

/* power radiated into the cone aMin..aMax (degrees) around (theta, phi), integrated on rings of
 * width delta with 6 n samples on ring n (capped at 60 degrees) */
float (farFieldFlux)(phaser P, tfield pol(phaser,float,float), float theta, float phi, float aMin, float aMax, float delta, ...)
{
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
		for (int n=nMin; n<=nMax; n++) {
			int N = n<floor(60/delta+0.5) ? n*6 : floor(60/delta+0.5)*6;
			float dth = cosf(MAX(aMin,delta*(n-0.5))*PI/180) - cosf(MIN(aMax,delta*(n+0.5))*PI/180);
			for (int m=0; m<N; m++) {
				float a = tilt_theta(theta, delta*n, 360.0*m/N);
				float b = tilt_phi(theta, delta*n, 360.0*m/N)+phi;
				float c = tilt_phi(-a, theta, phi-b)+180-360.0*m/N;
				float complex p = farField(P, Ppol, a, b);
				float complex s = farField(P, Spol, a, b);
				float pp = sq(cabsf((p * cosf(c*PI/180) - s * sinf(c*PI/180))));
				float ss = sq(cabsf((s * cosf(c*PI/180) + p * sinf(c*PI/180))));
				if (pol == Ppol) Sum += dth * (pp * (PI/N + sinf(PI/N)) + ss * (PI/N - sinf(PI/N)));
				if (pol == Spol) Sum += dth * (pp * (PI/N - sinf(PI/N)) + ss * (PI/N + sinf(PI/N)));
				if (pol == Unpol) Sum += 2 * PI/N * dth * (pp + ss);
			}
		}
	}
	return Sum;
}
