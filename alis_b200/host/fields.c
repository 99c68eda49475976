/*
 * fields.c -- point-wise derived quantities `float F(world W, int i, int j, int k)`:
 * co-located E/H components, energy densities, Poynting vector, Maxwell stress, current
 * density, scattered / source fields, effective permittivity and refractive index.
 *
 * Their ADDRESSES are part of the API (used as tags by pointDipole, sliceSnap, ...), so every
 * one is a real function.  They read the host views of the device-resident fields; each entry
 * point first makes sure the few cells it touches are present (alis_view_need -- a no-op when
 * the calling monitor has prefetched its region, which all library monitors do).
 *
 * Definitions follow reference src/fields.c (cited per group): the Yee staggering averages, the
 * mirror-symmetry parities and the summation order are the reference's, so monitor values agree
 * bit for bit when the fields do.
 */
#include <string.h>
#include "alis_c.h"
#include "host_internal.h"

/* ---- symmetric-boundary folding --------------------------------------------------------------------
 * A negative index on a SYM axis is folded to its mirror image and the value picks up the parity
 * of the component: E_c is even along its own axis c and odd along the other two, H_c the
 * opposite; the sign is carried by the world's cos factors (+-1 after the first source). */
typedef struct { int i, j, k; float p; } folded;

/* sx,sy,sz: +1 even / -1 odd; `shift` (0 or 1) adds 1 to the folded index on the component's own
 * axis (raw staggered quantities live half a cell back: mirror of -i is 1+i-... see rawJ, raweEff) */
static inline folded fold_cos(world W, int i, int j, int k, float p, int sx, int sy, int sz)
{
	folded r = { i, j, k, p };
	if (i<0 && W->xMinSur==SYM) r.i = -i, r.p *= sx > 0 ? W->coskx : -W->coskx;
	if (j<0 && W->yMinSur==SYM) r.j = -j, r.p *= sy > 0 ? W->cosky : -W->cosky;
	if (k<0 && W->zMinSur==SYM) r.k = -k, r.p *= sz > 0 ? W->coskz : -W->coskz;
	return r;
}

/* same, for products of two components: parity only, no cos factor */
static inline folded fold_sign(world W, int i, int j, int k, float p, int sx, int sy, int sz)
{
	folded r = { i, j, k, p };
	if (i<0 && W->xMinSur==SYM) r.i = -i, r.p *= sx;
	if (j<0 && W->yMinSur==SYM) r.j = -j, r.p *= sy;
	if (k<0 && W->zMinSur==SYM) r.k = -k, r.p *= sz;
	return r;
}

static inline folded fold_plain(world W, int i, int j, int k)
{
	return fold_sign(W, i, j, k, 1, 1, 1, 1);
}

#define NEED(W,r) alis_view_need(W, (r).i, (r).i+1, (r).j, (r).j+1, (r).k, (r).k+1)
/* only one array: (set 0 E / 1 H, component, field set m) */
#define NEED1(W,set,c,m,r) alis_view_need_arr(W, set, c, m, (r).i, (r).i+1, (r).j, (r).j+1, (r).k, (r).k+1)

/* ---- stencils of the Yee cell ---------------------------------------------------------------------- */
/* two samples of an E component along its own axis c */
static inline float e_pair(float ***A, int c, int i, int j, int k)
{
	return A[i][j][k] + A[i+(c==0)][j+(c==1)][k+(c==2)];
}
static inline float e_hi(float ***A, int c, int i, int j, int k) { return A[i+(c==0)][j+(c==1)][k+(c==2)]; }

/* four samples of an H component around the node, in the order (0,0), (+a1), (+a2), (+a1+a2)
 * with a1 = c+1, a2 = c+2 cyclic (the reference's summation order) */
static inline float h_quad(float ***A, int c, int i, int j, int k)
{
	const int a1 = (c+1)%3, a2 = (c+2)%3;
	const int d1[3] = { a1==0, a1==1, a1==2 }, d2[3] = { a2==0, a2==1, a2==2 };
	return A[i][j][k] + A[i+d1[0]][j+d1[1]][k+d1[2]] + A[i+d2[0]][j+d2[1]][k+d2[2]] + A[i+d1[0]+d2[0]][j+d1[1]+d2[1]][k+d1[2]+d2[2]];
}
static inline float h_quad_sq(float ***A, int c, int i, int j, int k)
{
	const int a1 = (c+1)%3, a2 = (c+2)%3;
	const int d1[3] = { a1==0, a1==1, a1==2 }, d2[3] = { a2==0, a2==1, a2==2 };
	return sq(A[i][j][k]) + sq(A[i+d1[0]][j+d1[1]][k+d1[2]]) + sq(A[i+d2[0]][j+d2[1]][k+d2[2]]) + sq(A[i+d1[0]+d2[0]][j+d1[1]+d2[1]][k+d1[2]+d2[2]]);
}
static inline float e_pair_sq(float ***A, int c, int i, int j, int k)
{
	return sq(A[i][j][k]) + sq(A[i+(c==0)][j+(c==1)][k+(c==2)]);
}

static inline float ***Ecomp(world W, int m, int c) { return c==0 ? W->E[m].x : c==1 ? W->E[m].y : W->E[m].z; }
static inline float ***Hcomp(world W, int m, int c) { return c==0 ? W->H[m].x : c==1 ? W->H[m].y : W->H[m].z; }
#define SGN(c,axis,own) ((c)==(axis) ? (own) : -(own))   /* E: own=+1, H: own=-1 */

/* ---- raw (staggered) components, reference fields.c:130-176 ----------------------------------------- */
static float raw_component(world W, int isH, int c, int i, int j, int k)
{
	const int own = isH ? -1 : 1;
	folded r = fold_cos(W, i, j, k, 1, SGN(c,0,own), SGN(c,1,own), SGN(c,2,own));
	NEED1(W, isH, c, 0, r);
	return r.p * (isH ? Hcomp(W,0,c) : Ecomp(W,0,c))[r.i][r.j][r.k];
}
float rawEx(world W, int i, int j, int k) { return raw_component(W, 0, 0, i, j, k); }
float rawEy(world W, int i, int j, int k) { return raw_component(W, 0, 1, i, j, k); }
float rawEz(world W, int i, int j, int k) { return raw_component(W, 0, 2, i, j, k); }
float rawHx(world W, int i, int j, int k) { return raw_component(W, 1, 0, i, j, k); }
float rawHy(world W, int i, int j, int k) { return raw_component(W, 1, 1, i, j, k); }
float rawHz(world W, int i, int j, int k) { return raw_component(W, 1, 2, i, j, k); }

/* ---- node-centred components, reference fields.c:179-291 ---------------------------------------------- */
static float e_component(world W, int m, int c, int i, int j, int k)
{
	if (m && !W->complexField) return 0;
	folded r = fold_cos(W, i, j, k, 0.5, SGN(c,0,1), SGN(c,1,1), SGN(c,2,1));
	NEED1(W, 0, c, m, r);
	return r.p * e_pair(Ecomp(W,m,c), c, r.i, r.j, r.k);
}
static float h_component(world W, int m, int c, int i, int j, int k)
{
	if (m && !W->complexField) return 0;
	folded r = fold_cos(W, i, j, k, 0.25, SGN(c,0,-1), SGN(c,1,-1), SGN(c,2,-1));
	NEED1(W, 1, c, m, r);
	return r.p * h_quad(Hcomp(W,m,c), c, r.i, r.j, r.k);
}
float Ex(world W, int i, int j, int k)  { return e_component(W, 0, 0, i, j, k); }
float Ey(world W, int i, int j, int k)  { return e_component(W, 0, 1, i, j, k); }
float Ez(world W, int i, int j, int k)  { return e_component(W, 0, 2, i, j, k); }
float iEx(world W, int i, int j, int k) { return e_component(W, 1, 0, i, j, k); }
float iEy(world W, int i, int j, int k) { return e_component(W, 1, 1, i, j, k); }
float iEz(world W, int i, int j, int k) { return e_component(W, 1, 2, i, j, k); }
float Hx(world W, int i, int j, int k)  { return h_component(W, 0, 0, i, j, k); }
float Hy(world W, int i, int j, int k)  { return h_component(W, 0, 1, i, j, k); }
float Hz(world W, int i, int j, int k)  { return h_component(W, 0, 2, i, j, k); }
float iHx(world W, int i, int j, int k) { return h_component(W, 1, 0, i, j, k); }
float iHy(world W, int i, int j, int k) { return h_component(W, 1, 1, i, j, k); }
float iHz(world W, int i, int j, int k) { return h_component(W, 1, 2, i, j, k); }

/* ---- effective permittivity / refractive index, reference fields.c:10-122 ------------------------------- */
static int needs_aux(matter M)
{
	for (int a=0; a<3; a++) for (int b=0; b<6; b++) if (M.d[a][b]) return 2;
	if (M.P[0].w || M.P[0].r) return 1;
	return 0;
}

/* susceptibility of pole p at photon energy w and its derivative (group-index term) */
static float pole_chi(matter M, float w, int p)
{
	if (M.P[p].f || M.P[p].g)
		return 0.5
			*((M.P[p].f * M.P[p].w - M.P[p].g * M.P[p].r + I * M.P[p].g * M.P[p].w) / (M.P[p].w + w + I * M.P[p].r)
			+ (M.P[p].f * M.P[p].w - M.P[p].g * M.P[p].r - I * M.P[p].g * M.P[p].w) / (M.P[p].w - w - I * M.P[p].r));
	if (M.P[p].w) return - sq(M.P[p].w) / (sq(w) + I * w * M.P[p].r);
	return 0;
}

static float pole_dchi(matter M, float w, int p)
{
	if (M.P[p].f || M.P[p].g)
		return 0.5
			*((M.P[p].f * M.P[p].w - M.P[p].g * M.P[p].r + I * M.P[p].g * M.P[p].w) / sq(M.P[p].w - w - I * M.P[p].r)
			- (M.P[p].f * M.P[p].w - M.P[p].g * M.P[p].r - I * M.P[p].g * M.P[p].w) / sq(M.P[p].w + w + I * M.P[p].r));
	if (M.P[p].w) return (2 * w + I * M.P[p].r) * sq(M.P[p].w) / sq(sq(w) + I * w * M.P[p].r);
	return 0;
}

/* energy-density permittivity d(w eps)/dw of component c at the world's source frequency; the
 * staggered component's mirror image of index -i on its own axis is 1-i */
static float raw_eps_eff(world W, int c, int i, int j, int k)
{
	coeffs *C = W->CE;
	float ***inv = c==0 ? C->x : c==1 ? C->y : C->z;
	if (!inv) return 1;
	if (i<0 && W->xMinSur==SYM) i = c==0 ? 1-i : -i;
	if (j<0 && W->yMinSur==SYM) j = c==1 ? 1-j : -j;
	if (k<0 && W->zMinSur==SYM) k = c==2 ? 1-k : -k;
	float f = 1 / inv[i][j][k];
	for (int n=0, o=0; n<C->N; n++, o++) {
		while (!needs_aux(C->M[o])) o++;
		if (C->iMin[n]<=i && i<=C->iMax[n] && C->jMin[n]<=j && j<=C->jMax[n] && C->kMin[n]<=k && k<=C->kMax[n]) {
			matter M = C->M[o];
			int tri = M.e.y || M.e.z ? 1 : 0;
			float a = c==0 ? C->ax[n] : c==1 ? C->ay[n] : C->az[n];
			float e = c==0 ? M.e.x : c==1 ? (M.e.y ? M.e.y : M.e.x) : (M.e.z ? M.e.z : M.e.x);
			float fill = (c==0 ? C->Jx[n] : c==1 ? C->Jy[n] : C->Jz[n])[i][j][k];
			f -= a * e * fill;
			for (int p=0; p<MAXPOLES; p+=1+2*tri)
				f += (pole_chi(M, W->f*W->eV, p+c*tri) + W->f*W->eV * pole_dchi(M, W->f*W->eV, p+c*tri)) * fill;
		}
	}
	return f;
}
float raweEffx(world W, int i, int j, int k) { return raw_eps_eff(W, 0, i, j, k); }
float raweEffy(world W, int i, int j, int k) { return raw_eps_eff(W, 1, i, j, k); }
float raweEffz(world W, int i, int j, int k) { return raw_eps_eff(W, 2, i, j, k); }

float eEff(world W, int i, int j, int k)
{
	folded r = fold_plain(W, i, j, k);
	i = r.i, j = r.j, k = r.k;
	return (raweEffx(W,i,j,k) + raweEffx(W,i+1,j,k) + raweEffy(W,i,j,k) + raweEffy(W,i,j+1,k) + raweEffz(W,i,j,k) + raweEffz(W,i,j,k+1)) / 6;
}

float RI(world W, int i, int j, int k)
{
	folded r = fold_plain(W, i, j, k);
	return sqrtf(fabsf(eEff(W, r.i, r.j, r.k)));
}

/* gradient magnitude of log(1+n) on a UNIFORM index raster (used to draw object outlines) */
float ContourRI(world W, int i, int j, int k)
{
	float x=0, y=0, z=0;
	i = W->xtoi(W, i*W->dx), j = W->ytoj(W, j*W->dy), k = W->ztok(W, k*W->dz);
	folded r = fold_plain(W, i, j, k);
	i = r.i, j = r.j, k = r.k;
	x = log10f(1+sqrtf(fabsf(raweEffx(W, i+1, j, k)))) - log10f(1+sqrtf(fabsf(raweEffx(W, i, j, k))));
	y = log10f(1+sqrtf(fabsf(raweEffy(W, i, j+1, k)))) - log10f(1+sqrtf(fabsf(raweEffy(W, i, j, k))));
	z = log10f(1+sqrtf(fabsf(raweEffz(W, i, j, k+1)))) - log10f(1+sqrtf(fabsf(raweEffz(W, i, j, k))));
	return sqrt(x*x+y*y+z*z);
}

float LogRI(world W, int i, int j, int k)   { return log10f(RI(W, i, j, k)); }
float Contour(world W, int i, int j, int k) { return ContourRI(W, i, j, k) ? 1 : 0; }

/* ---- current density, reference fields.c:296-376 ----------------------------------------------------------
 * J = (curl H term)(1 - eps^-1) + conduction + polarisation currents - impressed sources, / dt.
 * The box test is half-open on the two transverse axes of the component. */
static float raw_current(world W, int c, int i, int j, int k)
{
	coeffs *C = W->CE;
	float f = 0, p = 1;
	if (i<0 && W->xMinSur==SYM) i = c==0 ? 1-i : -i, p *= c==0 ? W->coskx : -W->coskx;
	if (j<0 && W->yMinSur==SYM) j = c==1 ? 1-j : -j, p *= c==1 ? W->cosky : -W->cosky;
	if (k<0 && W->zMinSur==SYM) k = c==2 ? 1-k : -k, p *= c==2 ? W->coskz : -W->coskz;
	alis_view_need(W, i, i+1, j, j+1, k, k+1);
	float ***inv = c==0 ? C->x : c==1 ? C->y : C->z;
	float ***Ec = Ecomp(W, 0, c);

	if (inv) {
		float ***hx = W->H->x, ***hy = W->H->y, ***hz = W->H->z;
		float curl = c==0 ? (C->dtdz[k] * (hy[i][j][k] - hy[i][j][k+1]) + C->dtdy[j] * (hz[i][j+1][k] - hz[i][j][k]))
		           : c==1 ? (C->dtdx[i] * (hz[i][j][k] - hz[i+1][j][k]) + C->dtdz[k] * (hx[i][j][k+1] - hx[i][j][k]))
		                  : (C->dtdy[j] * (hx[i][j][k] - hx[i][j+1][k]) + C->dtdx[i] * (hy[i+1][j][k] - hy[i][j][k]));
		f += curl * (1 - inv[i][j][k]);
		if (C->N) alis_view_need_ade_box(W, i, i, j, j, k, k);       /* the currents of this cell only */
		for (int n=0; n<C->N; n++) {
			const int ini = C->iMin[n]<=i && (c==0 ? i<=C->iMax[n] : i<C->iMax[n]);
			const int inj = C->jMin[n]<=j && (c==1 ? j<=C->jMax[n] : j<C->jMax[n]);
			const int ink = C->kMin[n]<=k && (c==2 ? k<=C->kMax[n] : k<C->kMax[n]);
			if (!(ini && inj && ink)) continue;
			float a = c==0 ? C->ax[n] : c==1 ? C->ay[n] : C->az[n];
			float ***fill = c==0 ? C->Jx[n] : c==1 ? C->Jy[n] : C->Jz[n];
			float ****J = c==0 ? W->E->Jx[n] : c==1 ? W->E->Jy[n] : W->E->Jz[n];
			if (a) f += 2 * fill[i][j][k] * Ec[i][j][k] / (1 + 1 / a);
			for (int q=0; q<C->NK[n]; q++) if (J[q]) f += J[q][i][j][k];
		}
	}
	for (source S = W->SE; S; S = S->S) {
		if (S->b200_set == 0 && S->b200_m == 0 && S->b200_comp == c && W->n < S->N &&
		    S->iMin <= i && i <= S->iMax && S->jMin <= j && j <= S->jMax && S->kMin <= k && k <= S->kMax)
			f -= (S->Amplitude ? S->Amplitude[i][j][k] : S->amplitude) * S->waveForm(W, S, W->H->t, S->Phase ? S->Phase[i][j][k] : S->phase);
	}
	return p * f / W->dt;
}
float rawJx(world W, int i, int j, int k) { return raw_current(W, 0, i, j, k); }
float rawJy(world W, int i, int j, int k) { return raw_current(W, 1, i, j, k); }
float rawJz(world W, int i, int j, int k) { return raw_current(W, 2, i, j, k); }
float Jx(world W, int i, int j, int k) { return 0.5 * (rawJx(W,i,j,k) + rawJx(W,i+1,j,k)); }
float Jy(world W, int i, int j, int k) { return 0.5 * (rawJy(W,i,j,k) + rawJy(W,i,j+1,k)); }
float Jz(world W, int i, int j, int k) { return 0.5 * (rawJz(W,i,j,k) + rawJz(W,i,j,k+1)); }

/* ---- E x H products (Poynting), reference fields.c:381-439, 483-485 -----------------------------------------
 * E_a H_b at the node: each of the two E_a samples times the sum of the two H_b samples that
 * flank it along the third axis. */
static float e_cross_h(world W, int a, int b, int i, int j, int k)
{
	const int t = 3-a-b;                                  /* the axis along which H_b is paired  */
	/* parity of the product: odd along the third axis only */
	folded r = fold_sign(W, i, j, k, 0.25, t==0 ? -1 : 1, t==1 ? -1 : 1, t==2 ? -1 : 1);
	NEED(W, r);
	i = r.i, j = r.j, k = r.k;
	float ***E = Ecomp(W,0,a), ***H = Hcomp(W,0,b);
	const int da[3] = { a==0, a==1, a==2 }, dt_[3] = { t==0, t==1, t==2 };
	return r.p * (E[i][j][k] * (H[i][j][k] + H[i+dt_[0]][j+dt_[1]][k+dt_[2]])
	            + E[i+da[0]][j+da[1]][k+da[2]] * (H[i+da[0]][j+da[1]][k+da[2]] + H[i+da[0]+dt_[0]][j+da[1]+dt_[1]][k+da[2]+dt_[2]]));
}
float ExHy(world W, int i, int j, int k) { return e_cross_h(W, 0, 1, i, j, k); }
float ExHz(world W, int i, int j, int k) { return e_cross_h(W, 0, 2, i, j, k); }
float EyHz(world W, int i, int j, int k) { return e_cross_h(W, 1, 2, i, j, k); }
float EyHx(world W, int i, int j, int k) { return e_cross_h(W, 1, 0, i, j, k); }
float EzHx(world W, int i, int j, int k) { return e_cross_h(W, 2, 0, i, j, k); }
float EzHy(world W, int i, int j, int k) { return e_cross_h(W, 2, 1, i, j, k); }
float Sx(world W, int i, int j, int k) { return EyHz(W,i,j,k) - EzHy(W,i,j,k); }
float Sy(world W, int i, int j, int k) { return EzHx(W,i,j,k) - ExHz(W,i,j,k); }
float Sz(world W, int i, int j, int k) { return ExHy(W,i,j,k) - EyHx(W,i,j,k); }

/* ---- energy-like scalars, reference fields.c:441-493 ---------------------------------------------------------- */
float JE(world W, int i, int j, int k)
{
	folded r = fold_plain(W, i, j, k);
	NEED(W, r);
	i = r.i, j = r.j, k = r.k;
	float f = 0;
	f += W->E->x[i][j][k] * rawJx(W,i,j,k) + W->E->x[i+1][j][k] * rawJx(W,i+1,j,k);
	f += W->E->y[i][j][k] * rawJy(W,i,j,k) + W->E->y[i][j+1][k] * rawJy(W,i,j+1,k);
	f += W->E->z[i][j][k] * rawJz(W,i,j,k) + W->E->z[i][j][k+1] * rawJz(W,i,j,k+1);
	return 0.5 * f;
}

float EE(world W, int i, int j, int k)
{
	folded r = fold_plain(W, i, j, k);
	NEED(W, r);
	float f = 0;
	for (int c=0; c<3; c++) f += e_pair_sq(Ecomp(W,0,c), c, r.i, r.j, r.k);
	return 0.5 * f;
}

float UE(world W, int i, int j, int k)
{
	folded r = fold_plain(W, i, j, k);
	NEED(W, r);
	i = r.i, j = r.j, k = r.k;
	float f = 0;
	f += sq(W->E->x[i][j][k]) * raweEffx(W, i, j, k) + sq(W->E->x[i+1][j][k]) * raweEffx(W, i+1, j, k);
	f += sq(W->E->y[i][j][k]) * raweEffy(W, i, j, k) + sq(W->E->y[i][j+1][k]) * raweEffy(W, i, j+1, k);
	f += sq(W->E->z[i][j][k]) * raweEffz(W, i, j, k) + sq(W->E->z[i][j][k+1]) * raweEffz(W, i, j, k+1);
	return 0.25 * f;
}

float HH(world W, int i, int j, int k)
{
	folded r = fold_plain(W, i, j, k);
	NEED(W, r);
	float f = 0;
	for (int c=0; c<3; c++) f += h_quad_sq(Hcomp(W,0,c), c, r.i, r.j, r.k);
	return 0.25 * f;
}

float UH(world W, int i, int j, int k)    { return 0.5*HH(W, i, j, k); }
float U(world W, int i, int j, int k)     { return UE(W, i, j, k)+UH(W, i, j, k); }
float LogJE(world W, int i, int j, int k) { return log10f(JE(W, i, j, k)); }
float LogEE(world W, int i, int j, int k) { return log10f(EE(W, i, j, k)); }
float LogUE(world W, int i, int j, int k) { return log10f(UE(W, i, j, k)); }
float LogHH(world W, int i, int j, int k) { return log10f(HH(W, i, j, k)); }
float LogUH(world W, int i, int j, int k) { return log10f(HH(W, i, j, k)); }   /* sic: as the reference */
float LogU(world W, int i, int j, int k)  { return log10f(U(W, i, j, k)); }

/* ---- Lorentz force density pieces, reference fields.c:498-562 -------------------------------------------------- */
static float j_cross_h(world W, int a, int b, int i, int j, int k)
{
	const int t = 3-a-b;
	folded r = fold_sign(W, i, j, k, 0.25, t==0 ? -1 : 1, t==1 ? -1 : 1, t==2 ? -1 : 1);
	NEED(W, r);
	i = r.i, j = r.j, k = r.k;
	float ***H = Hcomp(W,0,b);
	const int da[3] = { a==0, a==1, a==2 }, dt_[3] = { t==0, t==1, t==2 };
	return r.p * (raw_current(W, a, i, j, k) * (H[i][j][k] + H[i+dt_[0]][j+dt_[1]][k+dt_[2]])
	            + raw_current(W, a, i+da[0], j+da[1], k+da[2]) * (H[i+da[0]][j+da[1]][k+da[2]] + H[i+da[0]+dt_[0]][j+da[1]+dt_[1]][k+da[2]+dt_[2]]));
}
float JxHy(world W, int i, int j, int k) { return j_cross_h(W, 0, 1, i, j, k); }
float JxHz(world W, int i, int j, int k) { return j_cross_h(W, 0, 2, i, j, k); }
float JyHz(world W, int i, int j, int k) { return j_cross_h(W, 1, 2, i, j, k); }
float JyHx(world W, int i, int j, int k) { return j_cross_h(W, 1, 0, i, j, k); }
float JzHx(world W, int i, int j, int k) { return j_cross_h(W, 2, 0, i, j, k); }
float JzHy(world W, int i, int j, int k) { return j_cross_h(W, 2, 1, i, j, k); }

float DivE(world W, int i, int j, int k)
{
	folded r = fold_cos(W, i, j, k, 1, -1, -1, -1);
	NEED(W, r);
	i = r.i, j = r.j, k = r.k;
	return ((W->E->x[i+1][j][k] - W->E->x[i][j][k]) * W->CE->dtdx[i] +
	        (W->E->y[i][j+1][k] - W->E->y[i][j][k]) * W->CE->dtdy[j] +
	        (W->E->z[i][j][k+1] - W->E->z[i][j][k]) * W->CE->dtdz[k]) * r.p / W->dt;
}

/* ---- Maxwell stress tensor, reference fields.c:564-681 ------------------------------------------------------------ */
static float stress_diag(world W, int c, int i, int j, int k)
{
	folded r = fold_plain(W, i, j, k);
	NEED(W, r);
	float f = 0;
	for (int q=0; q<3; q++) { float h = h_quad_sq(Hcomp(W,0,q), q, r.i, r.j, r.k); if (q==c) f += h; else f -= h; }
	f *= 0.5;
	for (int q=0; q<3; q++) { float e = e_pair_sq(Ecomp(W,0,q), q, r.i, r.j, r.k); if (q==c) f += e; else f -= e; }
	return 0.25 * f;
}
float Txx(world W, int i, int j, int k) { return stress_diag(W, 0, i, j, k); }
float Tyy(world W, int i, int j, int k) { return stress_diag(W, 1, i, j, k); }
float Tzz(world W, int i, int j, int k) { return stress_diag(W, 2, i, j, k); }

/* off-diagonal T_ab: (first, second) is the factor order of the reference's products */
static float stress_off(world W, int first, int second, int i, int j, int k)
{
	const int t = 3-first-second;                         /* even along the axis not involved */
	folded r = fold_sign(W, i, j, k, 0.25, t==0 ? 1 : -1, t==1 ? 1 : -1, t==2 ? 1 : -1);
	NEED(W, r);
	float f = 0;
	f += e_pair(Ecomp(W,0,first), first, r.i, r.j, r.k) * e_pair(Ecomp(W,0,second), second, r.i, r.j, r.k);
	f += h_quad(Hcomp(W,0,first), first, r.i, r.j, r.k) * h_quad(Hcomp(W,0,second), second, r.i, r.j, r.k) * 0.25;
	return r.p * f;
}
float Txy(world W, int i, int j, int k) { return stress_off(W, 0, 1, i, j, k); }
float Tyx(world W, int i, int j, int k) { return stress_off(W, 0, 1, i, j, k); }
float Txz(world W, int i, int j, int k) { return stress_off(W, 2, 0, i, j, k); }
float Tzx(world W, int i, int j, int k) { return stress_off(W, 2, 0, i, j, k); }
float Tyz(world W, int i, int j, int k) { return stress_off(W, 1, 2, i, j, k); }
float Tzy(world W, int i, int j, int k) { return stress_off(W, 1, 2, i, j, k); }

float Fx(world W, int i, int j, int k) { return DivE(W, i, j, k) * Ex(W, i, j, k) + (JyHz(W, i, j, k) - JzHy(W, i, j, k)); }
float Fy(world W, int i, int j, int k) { return DivE(W, i, j, k) * Ey(W, i, j, k) + (JzHx(W, i, j, k) - JxHz(W, i, j, k)); }
float Fz(world W, int i, int j, int k) { return DivE(W, i, j, k) * Ez(W, i, j, k) + (JxHy(W, i, j, k) - JyHx(W, i, j, k)); }

/* ---- incident (source) and scattered E fields, reference fields.c:686-789 ------------------------------------------
 * The incident field lives in the TFSF auxiliary worlds.  The reference indexes their arrays with
 * main-world indices through row-pointer aliasing (source.c:223-248); here the index is clamped:
 * past a collapsed auxiliary axis it refers to row MAX+alt, alt = 1 for components staggered
 * along that axis. */
static float aux_E(world W, tfsf S, int c, int i, int j, int k)
{
	world A = S->W;
	if (A->iNUM < W->iNUM) { const int top = A->iMAX + (c==0); if (i > top) i = top; }
	if (A->jNUM < W->jNUM) { const int top = A->jMAX + (c==1); if (j > top) j = top; }
	return Ecomp(A,0,c)[i][j][k];
}

static float source_E(world W, int c, int i, int j, int k)
{
	folded r = fold_cos(W, i, j, k, 0.5, SGN(c,0,1), SGN(c,1,1), SGN(c,2,1));
	float f = 0;
	for (tfsf S = W->S; S; S = S->S) {
		alis_view_need_all(S->W);
		f += aux_E(W, S, c, r.i, r.j, r.k) + aux_E(W, S, c, r.i+(c==0), r.j+(c==1), r.k+(c==2));
	}
	return r.p * f;
}
float SoEx(world W, int i, int j, int k) { return source_E(W, 0, i, j, k); }
float SoEy(world W, int i, int j, int k) { return source_E(W, 1, i, j, k); }
float SoEz(world W, int i, int j, int k) { return source_E(W, 2, i, j, k); }
float SoEE(world W, int i, int j, int k) { return sq(SoEx(W,i,j,k))+sq(SoEy(W,i,j,k))+sq(SoEz(W,i,j,k)); }

/* total field minus the incident field inside the total-field region */
static float scattered_E(world W, int c, int i, int j, int k)
{
	folded r = fold_cos(W, i, j, k, 0.5, SGN(c,0,1), SGN(c,1,1), SGN(c,2,1));
	NEED1(W, 0, c, 0, r);                       /* E component c of the real set is all it reads of this world */
	i = r.i, j = r.j, k = r.k;
	float f = e_pair(Ecomp(W,0,c), c, i, j, k);
	for (tfsf S = W->S; S; S = S->S) {
		alis_view_need_all(S->W);
		const int lo[3] = { S->iMin, S->jMin, S->kMin }, hi[3] = { S->iMax, S->jMax, S->kMax };
		const int LO[3] = { S->iMIN, S->jMIN, S->kMIN }, HI[3] = { S->iMAX, S->jMAX, S->kMAX };
		const int q[3] = { i, j, k };
		int inside = 1;
		for (int a=0; a<3; a++) if (a != c && !(lo[a] <= q[a] && q[a] <= hi[a])) inside = 0;
		if (!inside) continue;
		if (LO[c] <= q[c] && q[c] <= HI[c]) f -= aux_E(W, S, c, i, j, k);
		if (LO[c] <= q[c]+1 && q[c]+1 <= HI[c]) f -= aux_E(W, S, c, i+(c==0), j+(c==1), k+(c==2));
	}
	return r.p * f;
}
float ScEx(world W, int i, int j, int k) { return scattered_E(W, 0, i, j, k); }
float ScEy(world W, int i, int j, int k) { return scattered_E(W, 1, i, j, k); }
float ScEz(world W, int i, int j, int k) { return scattered_E(W, 2, i, j, k); }
float ScEE(world W, int i, int j, int k) { return sq(ScEx(W,i,j,k))+sq(ScEy(W,i,j,k))+sq(ScEz(W,i,j,k)); }


/* ---- which host views a field function reads (bit (m*2+set)*3+comp); used by the region
 * prefetch of slices and world reductions so that e.g. a slice of Hz moves one array only ---------- */
unsigned alis_field_arrays(world W, field F)
{
	field *Ec[2][3] = { {Ex, Ey, Ez}, {iEx, iEy, iEz} }, *Hc[2][3] = { {Hx, Hy, Hz}, {iHx, iHy, iHz} };
	field *rawE[3] = { rawEx, rawEy, rawEz }, *rawH[3] = { rawHx, rawHy, rawHz };
	for (int m=0; m<2; m++) for (int c=0; c<3; c++) {
		if (F == Ec[m][c]) return 1u << ((m*2+0)*3+c);
		if (F == Hc[m][c]) return 1u << ((m*2+1)*3+c);
	}
	for (int c=0; c<3; c++) {
		if (F == rawE[c]) return 1u << c;
		if (F == rawH[c]) return 1u << (3+c);
	}
	if (F == EE) return 07;
	if (F == HH || F == UH) return 070;
	/* scattered / incident E (fields.c:686-789): one E component of this world (none for the incident field); the
	 * auxiliary worlds are fetched whole by the functions themselves */
	if (F == ScEx) return 01;
	if (F == ScEy) return 02;
	if (F == ScEz) return 04;
	if (F == ScEE) return 07;
	if (F == SoEx || F == SoEy || F == SoEz || F == SoEE) return 0;
	if (F == RI || F == LogRI || F == ContourRI || F == Contour || F == raweEffx || F == raweEffy || F == raweEffz || F == eEff) return 0;
	return W->complexField ? 07777 : 077;
}

/* ---- device-side evaluation ---------------------------------------------------------------
 * The field functions that read E and H only have a twin on the device (alisb_eval_box /
 * alisb_reduce_absmax / alisb_column_sums, csrc/monitor_eval.h): slices and world reductions of
 * those move the evaluated values (or one number) instead of the field arrays.  Everything else
 * (material-dependent quantities, currents, TFSF source/scattered fields, user callbacks) keeps
 * the host-view path above. */
static unsigned long device_calls;
unsigned long alisDeviceMonitorCalls(void) { return device_calls; }

/* ALIS_HOST_MONITORS=1 evaluates every monitor through the host views (the pre-f1 path): a
 * diagnostic switch, used by the tests to check the device monitors against the host ones */
int alis_host_monitors(void)
{
	static int mode = -1;
	if (mode < 0) {
		const char *s = getenv("ALIS_HOST_MONITORS");
		mode = s && atoi(s) != 0;
		/* several devices (ALIS_DEVICES, shard.c): a monitor's region may span z slabs; it is gathered through
		 * the host views, which read every slab's part (alis_bridge_read_box) */
		const char *d = getenv("ALIS_DEVICES");
		if (d && strchr(d, ',')) mode = 1;
	}
	return mode;
}

int alis_field_device_fn(world W, field F, alisb_field_fn *d, int *take_log)
{
	field *Ec[2][3] = { {Ex, Ey, Ez}, {iEx, iEy, iEz} }, *Hc[2][3] = { {Hx, Hy, Hz}, {iHx, iHy, iHz} };
	field *rawE[3] = { rawEx, rawEy, rawEz }, *rawH[3] = { rawHx, rawHy, rawHz };
	field *exh[6] = { ExHy, ExHz, EyHz, EyHx, EzHx, EzHy };
	static const int exh_ab[6][2] = { {0,1}, {0,2}, {1,2}, {1,0}, {2,0}, {2,1} };
	field *S[3] = { Sx, Sy, Sz }, *Td[3] = { Txx, Tyy, Tzz };
	field *To[6] = { Txy, Tyx, Txz, Tzx, Tyz, Tzy };
	static const int to_ab[6][2] = { {0,1}, {0,1}, {2,0}, {2,0}, {1,2}, {1,2} };
	memset(d, 0, sizeof *d);
	*take_log = 0;
	d->fn = -1;
	if (alis_host_monitors()) return 0;
	for (int m=0; m<2; m++) for (int c=0; c<3; c++) {
		if (F == Ec[m][c]) d->fn = ALISB_FN_NODE, d->set = 0, d->comp = c, d->m = m;
		if (F == Hc[m][c]) d->fn = ALISB_FN_NODE, d->set = 1, d->comp = c, d->m = m;
	}
	for (int c=0; c<3; c++) {
		if (F == rawE[c]) d->fn = ALISB_FN_RAW, d->set = 0, d->comp = c;
		if (F == rawH[c]) d->fn = ALISB_FN_RAW, d->set = 1, d->comp = c;
		if (F == S[c])    d->fn = ALISB_FN_S, d->comp = c;
		if (F == Td[c])   d->fn = ALISB_FN_TDIAG, d->comp = c;
	}
	for (int q=0; q<6; q++) {
		if (F == exh[q]) d->fn = ALISB_FN_EXH, d->comp = exh_ab[q][0], d->comp2 = exh_ab[q][1];
		if (F == To[q])  d->fn = ALISB_FN_TOFF, d->comp = to_ab[q][0], d->comp2 = to_ab[q][1];
	}
	if (F == EE) d->fn = ALISB_FN_EE;
	if (F == HH) d->fn = ALISB_FN_HH;
	if (F == UH) d->fn = ALISB_FN_UH;
	if (F == DivE) d->fn = ALISB_FN_DIVE;
	if (F == LogEE) d->fn = ALISB_FN_EE, *take_log = 1;
	if (F == LogHH || F == LogUH) d->fn = ALISB_FN_HH, *take_log = 1;      /* LogUH: sic, as the reference */
	if (d->fn < 0) return 0;
	d->sym[0] = W->xMinSur==SYM; d->sym[1] = W->yMinSur==SYM; d->sym[2] = W->zMinSur==SYM;
	d->cosk[0] = W->coskx; d->cosk[1] = W->cosky; d->cosk[2] = W->coskz;
	d->org[0] = 1-W->iMIN; d->org[1] = 1-W->jMIN; d->org[2] = 1-W->kMIN;
	return 1;
}

/* the index box, mirror images folded back, and the +1 reach of the stencils lie inside the arrays */
int alis_device_region_ok(world W, const alisb_field_fn *d, const int box[6])
{
	const int MIN_[3] = { W->iMIN, W->jMIN, W->kMIN }, MAX_[3] = { W->iMAX, W->jMAX, W->kMAX };
	const int reach = d->fn == ALISB_FN_RAW ? 0 : 1;
	for (int a=0; a<3; a++) {
		int lo = box[2*a], hi = box[2*a+1];
		if (lo > hi) return 0;
		if (d->sym[a] && lo < 0) { const int far = MAX(-lo, hi); lo = hi < 0 ? -hi : 0; hi = far; }
		if (lo < MIN_[a]-1 || hi + reach > MAX_[a]+1) return 0;
	}
	return 1;
}

/* F over the inclusive index box as a dense block (k fastest), evaluated on the device when F has a
 * device twin, through the host views otherwise.  Caller frees. */
float *alis_eval_region(world W, field F, int i0, int i1, int j0, int j1, int k0, int k1)
{
	const int ni = i1-i0+1, nj = j1-j0+1, nk = k1-k0+1;
	if (ni <= 0 || nj <= 0 || nk <= 0) return xcalloc(1, sizeof(float));
	float *v = xmalloc(sizeof(float)*(size_t)ni*nj*nk);
	alisb_field_fn d;
	int take_log;
	const int box[6] = { i0, i1, j0, j1, k0, k1 };
	if (alis_field_device_fn(W, F, &d, &take_log) && alis_device_region_ok(W, &d, box)) {
		const int32_t ib[6] = { i0, i1, j0, j1, k0, k1 };
		alis_bridge_commit(alis_host_bridge(W));                       /* DivE reads the dtd tables */
		if (alisb_eval_box(alis_bridge_engine(alis_host_bridge(W)), &d, ib, v) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
		if (take_log) for (size_t q=0; q<(size_t)ni*nj*nk; q++) v[q] = log10f(v[q]);
		device_calls++;
		return v;
	}
	int lo[3] = { i0, j0, k0 }, hi[3] = { i1, j1, k1 };
	const int sym[3] = { W->xMinSur==SYM, W->yMinSur==SYM, W->zMinSur==SYM };
	for (int a=0; a<3; a++) if (sym[a] && lo[a] < 0) { const int far = MAX(-lo[a], hi[a]); lo[a] = hi[a] < 0 ? -hi[a] : 0; hi[a] = far; }
	alis_view_need_for(W, F, lo[0], hi[0]+1, lo[1], hi[1]+1, lo[2], hi[2]+1);
	if (alis_field_arrays(W, F) == (W->complexField ? 07777u : 077u))      /* a function this file does not know better: it may read currents */
		alis_view_need_ade_box(W, lo[0], hi[0]+1, lo[1], hi[1]+1, lo[2], hi[2]+1);
	/* The reference evaluates a slice under `omp parallel for` (slices.c:262-266), so field functions -- its own and user
	 * callbacks -- are safe to call concurrently.  Here the region they read has just been fetched; the first cell is
	 * evaluated alone so that whatever a function fetches lazily (auxiliary worlds of the TFSF source, ADE currents, the
	 * engine itself) is in place before the team starts, and a fetch that still happens inside the team is serialised
	 * (view.c).  ALIS_PARALLEL_EVAL=0: one thread. */
	const size_t total = (size_t)ni*nj*nk;
	v[0] = F(W, i0, j0, k0);
	static int team = -1;
	if (team < 0) { const char *e = getenv("ALIS_PARALLEL_EVAL"); team = e ? atoi(e) != 0 : 1; }
	if (team && total >= 2048) {
		#pragma omp parallel for schedule(static)
		for (long t=1; t<(long)total; t++) {
			const int k = k0 + (int)(t % nk), j = j0 + (int)((t / nk) % nj), i = i0 + (int)(t / ((long)nk*nj));
			v[t] = F(W, i, j, k);
		}
	} else {
		size_t q = 0;
		for (int i=i0; i<=i1; i++) for (int j=j0; j<=j1; j++) for (int k=k0; k<=k1; k++, q++) if (q) v[q] = F(W, i, j, k);
	}
	return v;
}

/* max |F| over the inclusive index box; on the device when F has a twin there (Log variants are not
 * monotonic in |.|: host path).  Returns 0 and leaves *out alone when the host path must be used. */
int alis_device_absmax(world W, field F, const int box[6], float *out)
{
	alisb_field_fn d;
	int take_log;
	if (!alis_field_device_fn(W, F, &d, &take_log) || take_log || !alis_device_region_ok(W, &d, box)) return 0;
	const int32_t ib[6] = { box[0], box[1], box[2], box[3], box[4], box[5] };
	alis_bridge_commit(alis_host_bridge(W));
	if (alisb_reduce_absmax(alis_bridge_engine(alis_host_bridge(W)), &d, ib, out) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
	device_calls++;
	return 1;
}

/* z column sums of worldSum / sliceSum for every (i,j) of the box, from the device: cols[(i-i0)*nj +
 * (j-j0)].  Returns 0 when the host path must be used. */
int alis_device_column_sums(world W, field F, const int box[6], float *cols)
{
	alisb_field_fn d;
	int take_log;
	if (!alis_field_device_fn(W, F, &d, &take_log) || take_log || !alis_device_region_ok(W, &d, box)) return 0;
	if (box[4] < W->kMIN-1 || box[5] > W->kMAX+1) return 0;          /* dz is read at the unfolded index */
	const int32_t ib[6] = { box[0], box[1], box[2], box[3], box[4], box[5] };
	alis_bridge_commit(alis_host_bridge(W));
	if (alisb_column_sums(alis_bridge_engine(alis_host_bridge(W)), &d, ib, cols) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
	device_calls++;
	return 1;
}
