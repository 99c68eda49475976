/*
 * alis_oracle.c -- CPU ORACLE for the FDTD time-stepping path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference algorithm (hseelab/alis src/update.c,
 * src/phaser.c) behind the same C ABI as the CUDA engine (include/alis_b200.h), so the
 * two can be driven by identical descriptors and compared bit for bit.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this
 * library; nothing under alis_b200/ links it and there is no fallback to it.
 *
 * Parity status: PINNED -- validated bit-exactly against the unmodified reference built
 * from /root/reference (oracle/_ref, see oracle/build_ref.sh and tests/test_oracle_vs_ref.py)
 * and against the golden vectors under tests/golden/ that were generated from it.
 *
 * The monitor calls (field functions, slice / world reductions, absorption columns: reference
 * fields.c, slices.c, output.c) and the geometry calls (shape predicates and fill factors:
 * shapes.c, object.c) evaluate the restatements in alis_b200/csrc/monitor_eval.h and
 * geometry_eval.h -- the same source the CUDA kernels compile -- looped here in the reference's
 * order.  They are pinned the same way: through the host layer against the reference's monitor
 * values (tests/golden/monitor_values.json), its field hashes (shape_zoo and every configuration
 * with objects), and live against the reference on random worlds (tests/test_zzz_fuzz_reference.py).
 * The far-field surface sums (farfld.c:248-459) are restated below and additionally checked
 * against numpy (tests/test_farfield_sums.py).
 *
 * Layout here is the reference's: [a][b][c] with c (z) fastest, one ghost layer
 * (world.c:58-70,170-172).  Arithmetic is fp32 with the reference's expression order;
 * compile with -std=c99 -ffp-contract=off so no FMA is formed (the reference build,
 * gcc -O3 -std=c99 on x86-64, contains none).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
#include "alis_b200.h"

#define MAXOBJ 64
#define MAXSRC 256
#define MAXTFSF 8

typedef struct { int lo[3], hi[3]; } box3;

typedef struct {
	box3 bx; int n[3];
	float *v;
} block;                                  /* dense block over a box, z fastest */

typedef struct {
	box3 bx; int nj, nk; float a[3];
	float JJ[3][ALISB_MAXPOLES], JK[3][ALISB_MAXPOLES], JE[3][ALISB_MAXPOLES];
	float KJ[3][ALISB_MAXPOLES], KK[3][ALISB_MAXPOLES], KE[3][ALISB_MAXPOLES];
	block fill[3];
	block J[2][ALISB_MAXPOLES][3], K[2][ALISB_MAXPOLES][3];
} dispobj;

typedef struct {
	alisb_source d;
	float *Amp, *Pha;
} srcobj;

typedef struct {
	alisb_tfsf d;
	alisb_strip *sE, *sH;
} tfsfobj;

struct alisb_engine {
	alisb_grid g;
	int n[3];                             /* nx,ny,nz */
	size_t sb, sa;                        /* strides: c stride 1, b stride sb?, see IDX */
	float *F[2][2][3];                    /* [set][m][comp] */
	float *dtd[2][3], *P[2][3], *P1[2][3], *P2[2][3];
	block psi[2][2][3][2][2];             /* [set][m][axis][side][which] */
	float *eps[3];                        /* full (nx+2)(ny+2)(nz+2) arrays or NULL */
	int nobj; dispobj *obj[MAXOBJ];
	int nsrc[2]; srcobj *src[2][MAXSRC];
	int ntfsf; tfsfobj *tf[MAXTFSF];
	int ncopy[2]; alisb_copy *copy[2];
	uint64_t step; float t, tE, tH;
	float *halo_send[2][2][2], *halo_recv[2][2][2];   /* [set][m][comp x/y] */
	uint64_t launches;
	int f_lo, f_hi;                       /* plane filter of the current (partial) pass */
	float *spacing[2][3];
};

struct alisb_phaser {
	alisb_engine *e; box3 bx; int mask, nfreq;
	block acc[6][4];                      /* per face, E1,E2,H1,H2; v holds nfreq*2*cells floats */
};

static float *dup_table(const float *src, int n);
static char g_err[256] = "";
static int fail(int code, const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }

int alisb_abi_version(void) { return ALISB_ABI_VERSION; }
const char *alisb_last_error(void) { return g_err; }
const char *alisb_backend(void) { return "oracle-cpu"; }

#define IDX(e,a,b,c) (((size_t)(a)*((e)->n[1]+2) + (size_t)(b))*((e)->n[2]+2) + (size_t)(c))

static int block_alloc(block *B, const int lo[3], const int hi[3])
{
	size_t cnt = 1;
	for (int d=0; d<3; d++) { B->bx.lo[d]=lo[d]; B->bx.hi[d]=hi[d]; B->n[d]=hi[d]-lo[d]+1; if (B->n[d]<0) B->n[d]=0; cnt*=B->n[d]; }
	B->v = cnt ? calloc(cnt, sizeof(float)) : 0;
	return (cnt && !B->v) ? -1 : 0;
}
static inline size_t bidx(const block *B, int a, int b, int c)
{
	return (((size_t)(a-B->bx.lo[0]))*B->n[1] + (size_t)(b-B->bx.lo[1]))*B->n[2] + (size_t)(c-B->bx.lo[2]);
}
static void box_from(box3 *bx, const int32_t b[6]) { for (int d=0; d<3; d++) bx->lo[d]=b[2*d], bx->hi[d]=b[2*d+1]; }
static size_t box_count(const int32_t b[6]) { size_t c=1; for (int d=0; d<3; d++) c *= (b[2*d+1]>=b[2*d]) ? (size_t)(b[2*d+1]-b[2*d]+1) : 0; return c; }

int alisb_create(const alisb_grid *g, alisb_engine **out)
{
	if (!g || !out || g->nx<1 || g->ny<1 || g->nz<1) return fail(ALISB_EINVAL, "bad grid");
	alisb_engine *e = calloc(1, sizeof *e);
	if (!e) return fail(ALISB_ENOMEM, "oom");
	e->g = *g; e->n[0]=g->nx; e->n[1]=g->ny; e->n[2]=g->nz;
	size_t cnt = (size_t)(g->nx+2)*(g->ny+2)*(g->nz+2);
	for (int s=0; s<2; s++) for (int m=0; m<=g->complex_field; m++) for (int c=0; c<3; c++)
		if (!(e->F[s][m][c] = calloc(cnt, sizeof(float)))) return fail(ALISB_ENOMEM, "oom fields");
	for (int s=0; s<2; s++) for (int m=0; m<=g->complex_field; m++) for (int ax=0; ax<3; ax++) for (int sd=0; sd<2; sd++) {
		int lo[3] = {1,1,1}, hi[3] = {g->nx, g->ny, g->nz};
		lo[ax] = g->pml[s][ax][sd][0]; hi[ax] = g->pml[s][ax][sd][1];
		if (lo[ax] > hi[ax]) continue;
		for (int w=0; w<2; w++) if (block_alloc(&e->psi[s][m][ax][sd][w], lo, hi)) return fail(ALISB_ENOMEM, "oom psi");
	}
	size_t plane = (size_t)(g->nx+2)*(g->ny+2);
	for (int s=0; s<2; s++) for (int m=0; m<=g->complex_field; m++) for (int c=0; c<2; c++) {
		e->halo_send[s][m][c] = calloc(plane, sizeof(float));
		e->halo_recv[s][m][c] = calloc(plane, sizeof(float));
	}
	e->f_lo = 0; e->f_hi = g->nz + 1;
	*out = e;
	return ALISB_OK;
}

#define CLO(e,c) ((c) < (e)->f_lo ? (e)->f_lo : (c))
#define CHI(e,c) ((c) > (e)->f_hi ? (e)->f_hi : (c))

static void free_src(srcobj *s) { if (s) { free(s->Amp); free(s->Pha); free(s); } }

int alisb_clear_sources(alisb_engine *e)
{
	for (int s=0; s<2; s++) { for (int i=0; i<e->nsrc[s]; i++) free_src(e->src[s][i]); e->nsrc[s]=0; }
	for (int i=0; i<e->ntfsf; i++) { free(e->tf[i]->sE); free(e->tf[i]->sH); free(e->tf[i]); }
	e->ntfsf = 0;
	return ALISB_OK;
}

void alisb_destroy(alisb_engine *e)
{
	if (!e) return;
	alisb_clear_sources(e);
	for (int s=0; s<2; s++) for (int m=0; m<2; m++) for (int c=0; c<3; c++) free(e->F[s][m][c]);
	for (int s=0; s<2; s++) for (int ax=0; ax<3; ax++) { free(e->dtd[s][ax]); free(e->P[s][ax]); free(e->P1[s][ax]); free(e->P2[s][ax]); free(e->spacing[s][ax]); }
	for (int s=0; s<2; s++) for (int m=0; m<2; m++) for (int ax=0; ax<3; ax++) for (int sd=0; sd<2; sd++) for (int w=0; w<2; w++) free(e->psi[s][m][ax][sd][w].v);
	for (int c=0; c<3; c++) free(e->eps[c]);
	alisb_clear_dispersive(e);
	for (int s=0; s<2; s++) { free(e->copy[s]); for (int m=0; m<2; m++) for (int c=0; c<2; c++) { free(e->halo_send[s][m][c]); free(e->halo_recv[s][m][c]); } }
	free(e);
}

int alisb_set_stream(alisb_engine *e, void *s) { (void)e; (void)s; return ALISB_OK; }

int alisb_clear_dispersive(alisb_engine *e);
static float *dup_table(const float *src, int n)
{
	if (!src) return 0;
	float *p = malloc((size_t)(n+2)*sizeof(float));
	memcpy(p, src, (size_t)(n+2)*sizeof(float));
	return p;
}

int alisb_set_axis_tables(alisb_engine *e, int set, int axis, const float *dtd, const float *P, const float *P1, const float *P2)
{
	if (set<0 || set>1 || axis<0 || axis>2 || !dtd) return fail(ALISB_EINVAL, "bad tables");
	free(e->dtd[set][axis]); free(e->P[set][axis]); free(e->P1[set][axis]); free(e->P2[set][axis]);
	e->dtd[set][axis] = dup_table(dtd, e->n[axis]);
	e->P[set][axis] = dup_table(P, e->n[axis]);
	e->P1[set][axis] = dup_table(P1, e->n[axis]);
	e->P2[set][axis] = dup_table(P2, e->n[axis]);
	return ALISB_OK;
}

int alisb_set_inv_eps(alisb_engine *e, const float *cx, const float *cy, const float *cz)
{
	const float *src[3] = {cx, cy, cz};
	for (int c=0; c<3; c++) { free(e->eps[c]); e->eps[c]=0; }
	if (!cx && !cy && !cz) return ALISB_OK;
	if (!cx || !cy || !cz) return fail(ALISB_EINVAL, "inv_eps needs all three components");
	size_t cnt = (size_t)(e->n[0]+2)*(e->n[1]+2)*(e->n[2]+2);
	for (int c=0; c<3; c++) {
		int hi[3] = {e->n[0], e->n[1], e->n[2]}; hi[c]++;
		e->eps[c] = calloc(cnt, sizeof(float));
		size_t q = 0;
		for (int a=1; a<=hi[0]; a++) for (int b=1; b<=hi[1]; b++) for (int k=1; k<=hi[2]; k++) e->eps[c][IDX(e,a,b,k)] = src[c][q++];
	}
	return ALISB_OK;
}

int alisb_add_dispersive(alisb_engine *e, const alisb_dispersive *d)
{
	if (e->nobj >= MAXOBJ) return fail(ALISB_EINVAL, "too many dispersive objects");
	if (d->nk > ALISB_MAXPOLES || d->nj > d->nk) return fail(ALISB_EINVAL, "bad pole counts");
	dispobj *o = calloc(1, sizeof *o);
	box_from(&o->bx, d->box); o->nj=d->nj; o->nk=d->nk;
	memcpy(o->a, d->a, sizeof o->a);
	memcpy(o->JJ, d->JJ, sizeof o->JJ); memcpy(o->JK, d->JK, sizeof o->JK); memcpy(o->JE, d->JE, sizeof o->JE);
	memcpy(o->KJ, d->KJ, sizeof o->KJ); memcpy(o->KK, d->KK, sizeof o->KK); memcpy(o->KE, d->KE, sizeof o->KE);
	size_t cnt = box_count(d->box);
	for (int c=0; c<3; c++) {
		block_alloc(&o->fill[c], o->bx.lo, o->bx.hi);
		memcpy(o->fill[c].v, d->fill[c], cnt*sizeof(float));
		for (int m=0; m<=e->g.complex_field; m++) for (int p=0; p<o->nk; p++) {
			block_alloc(&o->J[m][p][c], o->bx.lo, o->bx.hi);
			if (p >= o->nj) block_alloc(&o->K[m][p][c], o->bx.lo, o->bx.hi);
		}
	}
	e->obj[e->nobj++] = o;
	return ALISB_OK;
}

static void free_obj(dispobj *o)
{
	for (int c=0; c<3; c++) { free(o->fill[c].v); for (int m=0; m<2; m++) for (int p=0; p<ALISB_MAXPOLES; p++) { free(o->J[m][p][c].v); free(o->K[m][p][c].v); } }
	free(o);
}
int alisb_clear_dispersive(alisb_engine *e)
{
	for (int n=0; n<e->nobj; n++) free_obj(e->obj[n]);
	e->nobj = 0;
	return ALISB_OK;
}

int alisb_add_source(alisb_engine *e, const alisb_source *s)
{
	if (s->set<0 || s->set>1 || s->comp<0 || s->comp>2 || s->m<0 || s->m>e->g.complex_field || !s->waveform) return fail(ALISB_EINVAL, "bad source");
	if (e->nsrc[s->set] >= MAXSRC) return fail(ALISB_EINVAL, "too many sources");
	srcobj *o = calloc(1, sizeof *o);
	o->d = *s;
	size_t cnt = box_count(s->box);
	if (s->Amplitude) { o->Amp = malloc(cnt*sizeof(float)); memcpy(o->Amp, s->Amplitude, cnt*sizeof(float)); }
	if (s->Phase) { o->Pha = malloc(cnt*sizeof(float)); memcpy(o->Pha, s->Phase, cnt*sizeof(float)); }
	e->src[s->set][e->nsrc[s->set]++] = o;
	return ALISB_OK;
}

int alisb_add_tfsf(alisb_engine *e, const alisb_tfsf *t)
{
	if (e->ntfsf >= MAXTFSF || !t->aux) return fail(ALISB_EINVAL, "bad tfsf");
	tfsfobj *o = calloc(1, sizeof *o);
	o->d = *t;
	o->sE = malloc(sizeof(alisb_strip)*(t->n_strips_E+1)); memcpy(o->sE, t->strips_E, sizeof(alisb_strip)*t->n_strips_E);
	o->sH = malloc(sizeof(alisb_strip)*(t->n_strips_H+1)); memcpy(o->sH, t->strips_H, sizeof(alisb_strip)*t->n_strips_H);
	e->tf[e->ntfsf++] = o;
	return ALISB_OK;
}

int alisb_set_copies(alisb_engine *e, int set, const alisb_copy *ops, int n)
{
	if (set<0 || set>1 || n<0) return fail(ALISB_EINVAL, "bad copies");
	free(e->copy[set]);
	e->copy[set] = malloc(sizeof(alisb_copy)*(n+1));
	if (n) memcpy(e->copy[set], ops, sizeof(alisb_copy)*n);
	e->ncopy[set] = n;
	return ALISB_OK;
}

/* ------------------------------------------------------------------------------------
 * The update itself.
 * ------------------------------------------------------------------------------------ */

/* reference update.c:48-84 (updatePx/Py/Pz) restated for a generic axis:
 * psi0 acts on component ax+1 using G component ax+2, psi1 on ax+2 using G comp ax+1. */
static void psi_slab(alisb_engine *e, int set, int m, int ax, int sd, int Alt)
{
	block *B0 = &e->psi[set][m][ax][sd][0], *B1 = &e->psi[set][m][ax][sd][1];
	if (!B0->v) return;
	const int c1 = (ax+1)%3, c2 = (ax+2)%3;
	float *F1 = e->F[set][m][c1], *F2 = e->F[set][m][c2];
	const float *G1 = e->F[1-set][m][c1], *G2 = e->F[1-set][m][c2];
	const float *P = e->P[set][ax], *Pa = e->P1[set][ax], *Pb = e->P2[set][ax];
	const float *C1 = set==ALISB_SET_E ? e->eps[c1] : 0, *C2 = set==ALISB_SET_E ? e->eps[c2] : 0;
	const int useC = (C1 && C2);
	long d[3] = {0,0,0}; d[ax] = Alt;
	const long off = (long)IDX(e,1,1,1) - (long)IDX(e,1-d[0],1-d[1],1-d[2]);   /* index shift of +Alt along ax */
	#pragma omp parallel for collapse(2)
	for (int a=B0->bx.lo[0]; a<=B0->bx.hi[0]; a++) for (int b=B0->bx.lo[1]; b<=B0->bx.hi[1]; b++) {
		for (int c=CLO(e,B0->bx.lo[2]); c<=CHI(e,B0->bx.hi[2]); c++) {
			const int q = ax==0 ? a : ax==1 ? b : c;
			const size_t i = IDX(e,a,b,c), p = bidx(B0,a,b,c);
			float v0 = B0->v[p] = P[q]*B0->v[p] + Pa[q]*(G2[i] - G2[i+off]);
			float v1 = B1->v[p] = P[q]*B1->v[p] + Pb[q]*(G1[i] - G1[i+off]);
			if (useC) { F1[i] += v0*C1[i]; F2[i] += v1*C2[i]; }
			else      { F1[i] += v0;       F2[i] += v1; }
		}
	}
}

/* reference update.c:87-125 (updateF) */
static void update_F(alisb_engine *e, int set, int m)
{
	const int Alt = set==ALISB_SET_E ? 1 : -1;
	const int nx=e->n[0], ny=e->n[1], nz=e->n[2];
	float *Fx=e->F[set][m][0], *Fy=e->F[set][m][1], *Fz=e->F[set][m][2];
	const float *Gx=e->F[1-set][m][0], *Gy=e->F[1-set][m][1], *Gz=e->F[1-set][m][2];
	const float *dtdx=e->dtd[set][0], *dtdy=e->dtd[set][1], *dtdz=e->dtd[set][2];
	const long sa = (long)(ny+2)*(nz+2)*Alt, sb = (long)(nz+2)*Alt, sc = Alt;

	if (set==ALISB_SET_E) {
		/* update.c:89-98: conductivity loss */
		for (int n=0; n<e->nobj; n++) {
			dispobj *o = e->obj[n];
			if (!(o->a[0] || o->a[1] || o->a[2])) continue;
			const int a0=o->bx.lo[0]<1?1:o->bx.lo[0], a1=o->bx.hi[0]>nx?nx:o->bx.hi[0];
			const int b0=o->bx.lo[1]<1?1:o->bx.lo[1], b1=o->bx.hi[1]>ny?ny:o->bx.hi[1];
			const int c0=CLO(e,o->bx.lo[2]<1?1:o->bx.lo[2]), c1=CHI(e,o->bx.hi[2]>nz?nz:o->bx.hi[2]);
			#pragma omp parallel for collapse(2)
			for (int a=a0; a<=a1; a++) for (int b=b0; b<=b1; b++) for (int c=c0; c<=c1; c++) {
				const size_t i=IDX(e,a,b,c), p=bidx(&o->fill[0],a,b,c);
				Fx[i] *= (1 - o->a[0]*o->fill[0].v[p]) / (1 + o->a[0]*o->fill[0].v[p]);
				Fy[i] *= (1 - o->a[1]*o->fill[1].v[p]) / (1 + o->a[1]*o->fill[1].v[p]);
				Fz[i] *= (1 - o->a[2]*o->fill[2].v[p]) / (1 + o->a[2]*o->fill[2].v[p]);
			}
		}
		/* update.c:99-108: subtract polarisation currents */
		for (int n=0; n<e->nobj; n++) {
			dispobj *o = e->obj[n];
			const int a0=o->bx.lo[0]<1?1:o->bx.lo[0], a1=o->bx.hi[0]>nx?nx:o->bx.hi[0];
			const int b0=o->bx.lo[1]<1?1:o->bx.lo[1], b1=o->bx.hi[1]>ny?ny:o->bx.hi[1];
			const int c0=CLO(e,o->bx.lo[2]<1?1:o->bx.lo[2]), c1=CHI(e,o->bx.hi[2]>nz?nz:o->bx.hi[2]);
			for (int p=0; p<o->nk; p++) {
				const float *Jx=o->J[m][p][0].v, *Jy=o->J[m][p][1].v, *Jz=o->J[m][p][2].v;
				#pragma omp parallel for collapse(2)
				for (int a=a0; a<=a1; a++) for (int b=b0; b<=b1; b++) for (int c=c0; c<=c1; c++) {
					const size_t i=IDX(e,a,b,c), q=bidx(&o->fill[0],a,b,c);
					Fx[i] -= Jx[q]; Fy[i] -= Jy[q]; Fz[i] -= Jz[q];
				}
			}
		}
	}

	/* update.c:109-118: curl */
	const float *Cx = set==ALISB_SET_E ? e->eps[0] : 0, *Cy = set==ALISB_SET_E ? e->eps[1] : 0, *Cz = set==ALISB_SET_E ? e->eps[2] : 0;
	if (Cx && Cy && Cz) {
		#pragma omp parallel for collapse(2)
		for (int a=1; a<=nx; a++) for (int b=1; b<=ny; b++) for (int c=CLO(e,1); c<=CHI(e,nz); c++) {
			const size_t i = IDX(e,a,b,c);
			Fx[i] += (dtdz[c]*(Gy[i]-Gy[i+sc]) + dtdy[b]*(Gz[i+sb]-Gz[i])) * Cx[i];
			Fy[i] += (dtdx[a]*(Gz[i]-Gz[i+sa]) + dtdz[c]*(Gx[i+sc]-Gx[i])) * Cy[i];
			Fz[i] += (dtdy[b]*(Gx[i]-Gx[i+sb]) + dtdx[a]*(Gy[i+sa]-Gy[i])) * Cz[i];
		}
	} else {
		#pragma omp parallel for collapse(2)
		for (int a=1; a<=nx; a++) for (int b=1; b<=ny; b++) for (int c=CLO(e,1); c<=CHI(e,nz); c++) {
			const size_t i = IDX(e,a,b,c);
			Fx[i] += (dtdz[c]*(Gy[i]-Gy[i+sc]) + dtdy[b]*(Gz[i+sb]-Gz[i]));
			Fy[i] += (dtdx[a]*(Gz[i]-Gz[i+sa]) + dtdz[c]*(Gx[i+sc]-Gx[i]));
			Fz[i] += (dtdy[b]*(Gx[i]-Gx[i+sb]) + dtdx[a]*(Gy[i+sa]-Gy[i]));
		}
	}
	/* update.c:119-124: slab order x-min, y-min, z-min, x-max, y-max, z-max */
	for (int sd=0; sd<2; sd++) for (int ax=0; ax<3; ax++) psi_slab(e, set, m, ax, sd, Alt);
}

/* reference update.c:128-158 (updateJ), acting on E from the H call */
static void update_J(alisb_engine *e, int m)
{
	const int nx=e->n[0], ny=e->n[1], nz=e->n[2];
	for (int n=0; n<e->nobj; n++) {
		dispobj *o = e->obj[n];
		const int a0=o->bx.lo[0]<1?1:o->bx.lo[0], a1=o->bx.hi[0]>nx?nx:o->bx.hi[0];
		const int b0=o->bx.lo[1]<1?1:o->bx.lo[1], b1=o->bx.hi[1]>ny?ny:o->bx.hi[1];
		const int c0=CLO(e,o->bx.lo[2]<1?1:o->bx.lo[2]), c1=CHI(e,o->bx.hi[2]>nz?nz:o->bx.hi[2]);
		for (int p=0; p<o->nk; p++) for (int k=0; k<3; k++) {
			const float *F = e->F[ALISB_SET_E][m][k], *fl = o->fill[k].v;
			float *J = o->J[m][p][k].v, *K = o->K[m][p][k].v;
			const float JJ=o->JJ[k][p], JK=o->JK[k][p], JE=o->JE[k][p], KJ=o->KJ[k][p], KK=o->KK[k][p], KE=o->KE[k][p];
			if (p < o->nj) {
				#pragma omp parallel for collapse(2)
				for (int a=a0; a<=a1; a++) for (int b=b0; b<=b1; b++) for (int c=c0; c<=c1; c++) {
					const size_t i=IDX(e,a,b,c), q=bidx(&o->fill[0],a,b,c);
					J[q] = JJ*J[q] + JE*fl[q]*F[i];
				}
			} else {
				#pragma omp parallel for collapse(2)
				for (int a=a0; a<=a1; a++) for (int b=b0; b<=b1; b++) for (int c=c0; c<=c1; c++) {
					const size_t i=IDX(e,a,b,c), q=bidx(&o->fill[0],a,b,c);
					J[q] = JJ*J[q] + JK*K[q] + JE*fl[q]*F[i];
					K[q] = KJ*J[q] + KK*K[q] + KE*fl[q]*F[i];
				}
			}
		}
	}
}

/* reference update.c:161-179 (updateS) */
static void apply_sources(alisb_engine *e, int set, float t)
{
	for (int n=0; n<e->nsrc[set]; n++) {
		srcobj *s = e->src[set][n];
		if (!(e->step < s->d.n_active)) continue;
		float *F = e->F[set][s->d.m][s->d.comp];
		const int32_t *bx = s->d.box;
		size_t q = 0;
		if (CLO(e,bx[4]) > CHI(e,bx[5])) continue;
		if (s->Pha) {
			for (int a=bx[0]; a<=bx[1]; a++) for (int b=bx[2]; b<=bx[3]; b++) for (int c=bx[4]; c<=bx[5]; c++, q++)
				if (c >= e->f_lo && c <= e->f_hi) F[IDX(e,a,b,c)] += (s->Amp ? s->Amp[q] : s->d.amplitude) * s->d.waveform(s->d.user, t, s->Pha[q]);
		} else {
			const float f = s->d.waveform(s->d.user, t, s->d.phase);
			for (int a=bx[0]; a<=bx[1]; a++) for (int b=bx[2]; b<=bx[3]; b++) for (int c=bx[4]; c<=bx[5]; c++, q++)
				if (c >= e->f_lo && c <= e->f_hi) F[IDX(e,a,b,c)] += (s->Amp ? s->Amp[q] : s->d.amplitude) * f;
		}
	}
}

/* reference update.c:194-272 (update, updateTFSFE/H) */
static void apply_tfsf(alisb_engine *e, int set, int step_aux)
{
	for (int n=0; n<e->ntfsf; n++) {
		tfsfobj *t = e->tf[n];
		alisb_engine *x = t->d.aux;
		if (step_aux) { if (set==ALISB_SET_E) alisb_step_E(x); else alisb_step_H(x); }
		const alisb_strip *S = set==ALISB_SET_E ? t->sE : t->sH;
		const int ns = set==ALISB_SET_E ? t->d.n_strips_E : t->d.n_strips_H;
		for (int m=0; m<=e->g.complex_field; m++) for (int q=0; q<ns; q++) {
			const alisb_strip *s = &S[q];
			float *F = e->F[set][m][s->comp];
			const float *G = x->F[1-set][m][s->aux_comp];
			const float *C = s->use_eps ? e->eps[s->comp] : 0;
			const int ca = t->d.clamp[1-set][s->aux_comp][0], cb = t->d.clamp[1-set][s->aux_comp][1];
			for (int a=s->box[0]; a<=s->box[1]; a++) for (int b=s->box[2]; b<=s->box[3]; b++) for (int c=CLO(e,s->box[4]); c<=CHI(e,s->box[5]); c++) {
				int xa = a + s->alt[0] + t->d.offset[0], xb = b + s->alt[1] + t->d.offset[1], xc = c + s->alt[2] + t->d.offset[2];
				if (xa > ca) xa = ca;
				if (xb > cb) xb = cb;
				const size_t i = IDX(e,a,b,c);
				if (C) F[i] += s->coef * G[IDX(x,xa,xb,xc)] * C[i];
				else   F[i] += s->coef * G[IDX(x,xa,xb,xc)];
			}
		}
	}
}

/* reference update.c:182-191 (copyBoundary) */
static void apply_copies(alisb_engine *e, int set)
{
	for (int n=0; n<e->ncopy[set]; n++) {
		const alisb_copy *k = &e->copy[set][n];
		float *F = e->F[k->set][0][k->comp], *G = e->g.complex_field ? e->F[k->set][1][k->comp] : 0;
		for (int a=k->box[0]; a<=k->box[1]; a++) for (int b=k->box[2]; b<=k->box[3]; b++) for (int c=CLO(e,k->box[4]); c<=CHI(e,k->box[5]); c++) {
			const size_t i = IDX(e,a,b,c), j = IDX(e,a+k->alt[0],b+k->alt[1],c+k->alt[2]);
			if (G) {
				F[i] = k->cosk*F[j] + k->sink*G[j];
				G[i] = k->cosk*G[j] - k->sink*F[j];
			} else F[i] = k->cosk*F[j];
		}
	}
}

int alisb_step_part(alisb_engine *e, int set, int part)
{
	if (set<0 || set>1 || part<0 || part>2) return fail(ALISB_EINVAL, "bad step part");
	const int send = set==ALISB_SET_E ? e->n[2] : 1;
	e->f_lo = 0; e->f_hi = e->n[2]+1;
	if (part == ALISB_PART_SEND_PLANE) e->f_lo = e->f_hi = send;
	if (part == ALISB_PART_REST) { if (set==ALISB_SET_E) e->f_hi = send-1; else e->f_lo = send+1; }
	for (int m=0; m<=e->g.complex_field; m++) { update_F(e, set, m); if (set==ALISB_SET_H) update_J(e, m); }
	apply_tfsf(e, set, part != ALISB_PART_REST);
	apply_sources(e, set, set==ALISB_SET_E ? e->tH : e->tE);
	apply_copies(e, set);
	e->f_lo = 0; e->f_hi = e->n[2]+1;
	if (part != ALISB_PART_SEND_PLANE) {
		if (set==ALISB_SET_E) e->t = e->tE = ++e->step * e->g.dt;         /* update.c:298 */
		else e->t = e->tH = e->tE + 0.5 * e->g.dt;                         /* update.c:340 (double arithmetic, then float) */
	}
	e->launches++;
	return ALISB_OK;
}

int alisb_step_E(alisb_engine *e) { return alisb_step_part(e, ALISB_SET_E, ALISB_PART_ALL); }
int alisb_step_H(alisb_engine *e) { return alisb_step_part(e, ALISB_SET_H, ALISB_PART_ALL); }

int alisb_run(alisb_engine *e, int steps) { for (int i=0; i<steps; i++) { alisb_step_E(e); alisb_step_H(e); } return ALISB_OK; }
int alisb_synchronize(alisb_engine *e) { (void)e; return ALISB_OK; }
/* the oracle always steps in two plain passes; fusion is an implementation choice of the CUDA engine */
int alisb_set_fusion(alisb_engine *e, int on) { (void)e; (void)on; return ALISB_OK; }
int alisb_flush(alisb_engine *e, int h) { (void)e; (void)h; return ALISB_OK; }
int alisb_begin_pass(alisb_engine *e, int set) { (void)e; (void)set; return ALISB_OK; }
uint64_t alisb_fused_launch_count(alisb_engine *e) { (void)e; return 0; }
int alisb_get_time(alisb_engine *e, uint64_t *n, float *t, float *tE, float *tH) { if (n) *n=e->step; if (t) *t=e->t; if (tE) *tE=e->tE; if (tH) *tH=e->tH; return ALISB_OK; }
int alisb_set_time(alisb_engine *e, uint64_t n, float t, float tE, float tH) { e->step=n; e->t=t; e->tE=tE; e->tH=tH; return ALISB_OK; }

int alisb_reset_fields(alisb_engine *e)
{
	size_t cnt = (size_t)(e->n[0]+2)*(e->n[1]+2)*(e->n[2]+2);
	for (int s=0; s<2; s++) for (int m=0; m<=e->g.complex_field; m++) {
		for (int c=0; c<3; c++) memset(e->F[s][m][c], 0, cnt*sizeof(float));
		for (int ax=0; ax<3; ax++) for (int sd=0; sd<2; sd++) for (int w=0; w<2; w++) {
			block *B = &e->psi[s][m][ax][sd][w];
			if (B->v) memset(B->v, 0, (size_t)B->n[0]*B->n[1]*B->n[2]*sizeof(float));
		}
	}
	for (int n=0; n<e->nobj; n++) for (int m=0; m<=e->g.complex_field; m++) for (int p=0; p<e->obj[n]->nk; p++) for (int c=0; c<3; c++) {
		block *B = &e->obj[n]->J[m][p][c];
		memset(B->v, 0, (size_t)B->n[0]*B->n[1]*B->n[2]*sizeof(float));
		B = &e->obj[n]->K[m][p][c];
		if (B->v) memset(B->v, 0, (size_t)B->n[0]*B->n[1]*B->n[2]*sizeof(float));
	}
	return ALISB_OK;
}

/* resolve (arr, ids) to either the full-grid array or a dense block */
static int resolve(alisb_engine *e, int arr, const int32_t ids[6], float **full, block **blk)
{
	*full = 0; *blk = 0;
	switch (arr) {
	case ALISB_ARR_FIELD:
		if (ids[0]<0||ids[0]>1||ids[1]<0||ids[1]>2||ids[2]<0||ids[2]>e->g.complex_field) return -1;
		*full = e->F[ids[0]][ids[2]][ids[1]]; return 0;
	case ALISB_ARR_INVEPS:
		if (ids[0]<0||ids[0]>2||!e->eps[ids[0]]) return -1;
		*full = e->eps[ids[0]]; return 0;
	case ALISB_ARR_PSI:
		if (ids[0]<0||ids[0]>1||ids[1]<0||ids[1]>e->g.complex_field||ids[2]<0||ids[2]>2||ids[3]<0||ids[3]>1||ids[4]<0||ids[4]>1) return -1;
		*blk = &e->psi[ids[0]][ids[1]][ids[2]][ids[3]][ids[4]]; return (*blk)->v ? 0 : -1;
	case ALISB_ARR_ADE_J: case ALISB_ARR_ADE_K:
		if (ids[0]<0||ids[0]>e->g.complex_field||ids[1]<0||ids[1]>=e->nobj||ids[2]<0||ids[2]>=e->obj[ids[1]]->nk||ids[3]<0||ids[3]>2) return -1;
		*blk = arr==ALISB_ARR_ADE_J ? &e->obj[ids[1]]->J[ids[0]][ids[2]][ids[3]] : &e->obj[ids[1]]->K[ids[0]][ids[2]][ids[3]];
		return (*blk)->v ? 0 : -1;
	case ALISB_ARR_ADE_FILL:
		if (ids[0]<0||ids[0]>=e->nobj||ids[1]<0||ids[1]>2) return -1;
		*blk = &e->obj[ids[0]]->fill[ids[1]]; return 0;
	}
	return -1;
}

static int rw_box(alisb_engine *e, int arr, const int32_t ids[6], const int32_t box[6], float *host, int write)
{
	float *full; block *B;
	if (resolve(e, arr, ids, &full, &B)) return fail(ALISB_EINVAL, "bad array id");
	size_t q = 0;
	for (int a=box[0]; a<=box[1]; a++) for (int b=box[2]; b<=box[3]; b++) for (int c=box[4]; c<=box[5]; c++, q++) {
		float *p;
		if (full) {
			if (a<0||a>e->n[0]+1||b<0||b>e->n[1]+1||c<0||c>e->n[2]+1) return fail(ALISB_EINVAL, "box outside array");
			p = &full[IDX(e,a,b,c)];
		} else {
			if (a<B->bx.lo[0]||a>B->bx.hi[0]||b<B->bx.lo[1]||b>B->bx.hi[1]||c<B->bx.lo[2]||c>B->bx.hi[2]) return fail(ALISB_EINVAL, "box outside block");
			p = &B->v[bidx(B,a,b,c)];
		}
		if (write) *p = host[q]; else host[q] = *p;
	}
	return ALISB_OK;
}
int alisb_read_box(alisb_engine *e, int arr, const int32_t ids[6], const int32_t box[6], float *host) { return rw_box(e, arr, ids, box, host, 0); }
int alisb_write_box(alisb_engine *e, int arr, const int32_t ids[6], const int32_t box[6], const float *host) { return rw_box(e, arr, ids, box, (float*)host, 1); }

/* ---- monitor reductions: reference output.c:233-332, same loop order and float arithmetic ---- */
int alisb_set_spacing(alisb_engine *e, int set, int axis, const float *d)
{
	if (set<0 || set>1 || axis<0 || axis>2 || !d) return fail(ALISB_EINVAL, "bad spacing table");
	free(e->spacing[set][axis]);
	e->spacing[set][axis] = dup_table(d, e->n[axis]);
	return ALISB_OK;
}

static inline float pickf(alisb_engine *e, const float *A, int a, int at, int ib, int ic)
{
	int q[3];
	q[a] = at; q[(a+1)%3] = ib; q[(a+2)%3] = ic;
	return A[IDX(e, q[0], q[1], q[2])];
}

int alisb_plane_flux(alisb_engine *e, const alisb_flux_plane *planes, int n, double *out)
{
	for (int q=0; q<n; q++) {
		const alisb_flux_plane *P = &planes[q];
		const int a = P->axis, b = (a+1)%3, c = (a+2)%3;
		const float *Eb = e->F[0][0][b], *Ec = e->F[0][0][c], *Hb = e->F[1][0][b], *Hc = e->F[1][0][c];
		const float *dEb = e->spacing[0][b], *dEc = e->spacing[0][c], *dHb = e->spacing[1][b], *dHc = e->spacing[1][c];
		float S1 = 0, S2 = 0;
		#define E_H(E, H, ib, ic) pickf(e,E,a,P->at,ib,ic) * (pickf(e,H,a,P->at,ib,ic) + pickf(e,H,a,P->at+1,ib,ic))
		for (int ib=P->b0+1; ib<=P->b1; ib++) {
			float sum = 0;
			sum += 0.5 * dEc[P->c0] * E_H(Eb, Hc, ib, P->c0);
			sum += 0.5 * dEc[P->c1] * E_H(Eb, Hc, ib, P->c1);
			for (int ic=P->c0+1; ic<P->c1; ic++) sum += dEc[ic] * E_H(Eb, Hc, ib, ic);
			S1 += dHb[ib] * sum;
		}
		for (int ic=P->c0+1; ic<=P->c1; ic++) {
			float sum = 0;
			sum -= 0.5 * dEb[P->b0] * E_H(Ec, Hb, P->b0, ic);
			sum -= 0.5 * dEb[P->b1] * E_H(Ec, Hb, P->b1, ic);
			for (int ib=P->b0+1; ib<P->b1; ib++) sum -= dEb[ib] * E_H(Ec, Hb, ib, ic);
			S2 += dHc[ic] * sum;
		}
		#undef E_H
		out[2*q] = S1; out[2*q+1] = S2;
	}
	return ALISB_OK;
}

/* ---- point-wise field functions, slices and world reductions: reference fields.c:130-681 through
 * the shared restatement alis_b200/csrc/monitor_eval.h, looped like slices.c:382-394 (max) and
 * output.c:73-80 (z column sums) ---------------------------------------------------------------- */
typedef struct { alisb_engine *e; const alisb_field_fn *f; const dispobj *o; } mon_ctx;
#define MON_FN static inline
#define MON_CTX mon_ctx
#define MON_DESC(X) ((X)->f)
#define MON_LD(X,set,m,c,a,b,cc) ((X)->e->F[set][m][c][IDX((X)->e, a, b, cc)])
#define MON_DTD(X,axis,q) ((X)->e->dtd[0][axis][q])
#define MON_DT(X) ((X)->e->g.dt)
#define MON_COMPLEX(X) ((X)->e->g.complex_field)
#define MON_INV(X,c,a,b,cc) ((X)->e->eps[c][IDX((X)->e, a, b, cc)])
#define MON_FILL(X,c,a,b,cc) ((X)->o->fill[c].v[bidx(&(X)->o->fill[c], a, b, cc)])
#define MON_J(X,p,c,a,b,cc) ((X)->o->J[0][p][c].v[bidx(&(X)->o->J[0][p][c], a, b, cc)])
#define MON_NPOLES(X) ((X)->o->nk)
#define MON_LOSS(X,c) ((X)->o->a[c])
#define MON_DZE(X,cc) ((X)->e->spacing[0][2][cc])
#define MON_DZH(X,cc) ((X)->e->spacing[1][2][cc])
#include "../alis_b200/csrc/monitor_eval.h"

static int eval_check(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], int columns)
{
	if (!f || !ibox) return fail(ALISB_EINVAL, "eval: null argument");
	if (f->fn < ALISB_FN_RAW || f->fn > ALISB_FN_DIVE || f->set<0 || f->set>1 || f->comp<0 || f->comp>2 || f->comp2<0 || f->comp2>2 || f->m<0 || f->m>1)
		return fail(ALISB_EINVAL, "eval: bad field function");
	if ((f->fn == ALISB_FN_EXH || f->fn == ALISB_FN_TOFF) && f->comp == f->comp2) return fail(ALISB_EINVAL, "eval: the two components must differ");
	if (f->m > e->g.complex_field && f->fn != ALISB_FN_NODE) return fail(ALISB_EINVAL, "eval: no such field set");
	const int reach = f->fn == ALISB_FN_RAW ? 0 : 1;
	for (int ax=0; ax<3; ax++) {
		int lo = ibox[2*ax], hi = ibox[2*ax+1];
		if (lo > hi) return fail(ALISB_EINVAL, "eval: empty box");
		if (columns && ax == 2 && (lo + f->org[2] < 0 || hi + f->org[2] > e->n[2] + 1)) return fail(ALISB_EINVAL, "eval: column range outside the spacing table");
		if (f->sym[ax] && lo < 0) { const int far = -lo > hi ? -lo : hi; lo = hi < 0 ? -hi : 0; hi = far; }
		if (lo + f->org[ax] < 0 || hi + f->org[ax] + reach > e->n[ax] + 1) return fail(ALISB_EINVAL, "eval: box outside the arrays");
	}
	if (columns && !e->spacing[0][2]) return fail(ALISB_ESTATE, "alisb_set_spacing has not been called");
	return ALISB_OK;
}

int alisb_eval_box(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], float *host)
{
	int rc = eval_check(e, f, ibox, 0);
	if (rc) return rc;
	const mon_ctx X = { e, f, 0 };
	size_t q = 0;
	for (int i=ibox[0]; i<=ibox[1]; i++) for (int j=ibox[2]; j<=ibox[3]; j++) for (int k=ibox[4]; k<=ibox[5]; k++) host[q++] = mon_eval(&X, i, j, k);
	return ALISB_OK;
}

/* reference slices.c:282-379: the per-point running sums of sliceTimeAvg / sliceFreqDom */
struct alisb_accum { alisb_engine *e; alisb_field_fn f; int32_t ibox[6]; size_t n; float *re, *im; };

int alisb_accum_create(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], int with_imag, alisb_accum **out)
{
	int rc = eval_check(e, f, ibox, 0);
	if (rc) return rc;
	alisb_accum *a = calloc(1, sizeof *a);
	a->e = e; a->f = *f; memcpy(a->ibox, ibox, sizeof a->ibox);
	a->n = (size_t)(ibox[1]-ibox[0]+1)*(ibox[3]-ibox[2]+1)*(ibox[5]-ibox[4]+1);
	a->re = calloc(a->n, sizeof(float));
	a->im = with_imag ? calloc(a->n, sizeof(float)) : 0;
	*out = a;
	return ALISB_OK;
}

int alisb_accum_add(alisb_accum *a, float wr, float wi)
{
	const mon_ctx X = { a->e, &a->f, 0 };
	size_t q = 0;
	for (int i=a->ibox[0]; i<=a->ibox[1]; i++) for (int j=a->ibox[2]; j<=a->ibox[3]; j++) for (int k=a->ibox[4]; k<=a->ibox[5]; k++, q++) {
		const float v = mon_eval(&X, i, j, k);
		a->re[q] += v * wr;                              /* slices.c:352 */
		if (a->im) a->im[q] += v * wi;                   /* slices.c:353 */
	}
	return ALISB_OK;
}

int alisb_accum_read(alisb_accum *a, float *re, float *im)
{
	memcpy(re, a->re, a->n*sizeof(float));
	if (im && a->im) memcpy(im, a->im, a->n*sizeof(float));
	return ALISB_OK;
}

int alisb_accum_clear(alisb_accum *a)
{
	memset(a->re, 0, a->n*sizeof(float));
	if (a->im) memset(a->im, 0, a->n*sizeof(float));
	return ALISB_OK;
}

void alisb_accum_destroy(alisb_accum *a) { if (a) { free(a->re); free(a->im); free(a); } }

int alisb_reduce_absmax(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], float *out)
{
	int rc = eval_check(e, f, ibox, 0);
	if (rc) return rc;
	const mon_ctx X = { e, f, 0 };
	float max = 0, v;
	for (int i=ibox[0]; i<=ibox[1]; i++) for (int j=ibox[2]; j<=ibox[3]; j++) for (int k=ibox[4]; k<=ibox[5]; k++)
		if (max < (v = fabsf(mon_eval(&X, i, j, k)))) max = v;
	*out = max;
	return ALISB_OK;
}

int alisb_column_sums(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], float *host)
{
	int rc = eval_check(e, f, ibox, 1);
	if (rc) return rc;
	const mon_ctx X = { e, f, 0 };
	const float *dz = e->spacing[0][2];
	size_t q = 0;
	for (int i=ibox[0]; i<=ibox[1]; i++) for (int j=ibox[2]; j<=ibox[3]; j++) {
		float s;
		MON_COLUMN_SUM(&X, dz, i, j, ibox[4], ibox[5], s);
		host[q++] = s;
	}
	return ALISB_OK;
}

int alisb_absorption_columns(alisb_engine *e, int obj, const int32_t box[6], float *host)
{
	if (!box || !host) return fail(ALISB_EINVAL, "absorption: null argument");
	if (obj < 0 || obj >= e->nobj) return fail(ALISB_EINVAL, "absorption: no such object");
	if (!e->eps[0] || !e->eps[1] || !e->eps[2]) return fail(ALISB_ESTATE, "absorption: no material arrays");
	if (!e->spacing[0][2] || !e->spacing[1][2]) return fail(ALISB_ESTATE, "alisb_set_spacing has not been called");
	const dispobj *o = e->obj[obj];
	for (int d=0; d<3; d++) {
		const int lo = o->bx.lo[d] > 1 ? o->bx.lo[d] : 1, hi = o->bx.hi[d] < e->n[d] ? o->bx.hi[d] : e->n[d];
		if (box[2*d] > box[2*d+1] || box[2*d] < lo || box[2*d+1] > hi) return fail(ALISB_EINVAL, "absorption: box outside the object");
	}
	static const alisb_field_fn none;
	const mon_ctx X = { e, &none, o };
	size_t q = 0;
	for (int a=box[0]; a<=box[1]; a++) for (int b=box[2]; b<=box[3]; b++) {
		float xs, ys, zs;
		MON_ABSORPTION_COLUMN(&X, a, b, box[4], box[5], 1, e->n[2], xs, ys, zs);
		host[q++] = xs; host[q++] = ys; host[q++] = zs;
	}
	return ALISB_OK;
}

/* ---- geometry rasterisation: reference shapes.c:17-221 + object.c:70-121,229-262 through the shared
 * restatement alis_b200/csrc/geometry_eval.h, looped like the reference ------------------------- */
#define GEO_FN static
#include "../alis_b200/csrc/geometry_eval.h"

static int geo_check(const alisb_geometry *g, const int32_t ibox[6])
{
	if (!g || !ibox || g->nobjects < 1 || !g->roots || g->nnodes < 1 || !g->nodes || g->subn < 1 || g->subn > 64) return fail(ALISB_EINVAL, "geometry: bad description");
	for (int a=0; a<3; a++) {
		if (!g->coord[a] || g->lo[a] > g->hi[a]) return fail(ALISB_EINVAL, "geometry: missing coordinate table");
		if (ibox[2*a] > ibox[2*a+1] || ibox[2*a] < g->lo[a] || ibox[2*a+1] > g->hi[a]) return fail(ALISB_EINVAL, "geometry: box outside the coordinate tables");
	}
	for (int o=0; o<g->nobjects; o++) if (g->roots[o] < 0 || g->roots[o] >= g->nnodes) return fail(ALISB_EINVAL, "geometry: bad root");
	for (int n=0; n<g->nnodes; n++) {
		const alisb_shape *s = &g->nodes[n];
		if (s->type < 0 || s->type >= ALISB_SHAPE_COUNT) return fail(ALISB_EINVAL, "geometry: unknown shape type");
		if (s->type > ALISB_SHAPE_HRODZY) {
			if (s->nchild < 1 || s->child <= n || s->child + s->nchild > g->nnodes) return fail(ALISB_EINVAL, "geometry: bad child range");
			const int need = s->type == ALISB_SHAPE_ROTATION ? 6 : (s->type >= ALISB_SHAPE_CLATTICEX ? 2*(int)s->f[0][0] : 0);
			if (need && (s->aux < 0 || s->aux + need > g->naux || !g->aux)) return fail(ALISB_EINVAL, "geometry: bad constant range");
		}
	}
	return ALISB_OK;
}

int alisb_geometry_bounds(int device, const alisb_geometry *g, const int32_t ibox[6], int32_t *bounds)
{
	(void)device;
	int rc = geo_check(g, ibox);
	if (rc) return rc;
	for (int o=0; o<g->nobjects; o++) for (int a=0; a<3; a++) { bounds[6*o+2*a] = INT32_MAX; bounds[6*o+2*a+1] = INT32_MIN; }
	for (int i=ibox[0]; i<=ibox[1]; i++) for (int j=ibox[2]; j<=ibox[3]; j++) for (int k=ibox[4]; k<=ibox[5]; k++) {
		const float x = geo_coord(g, 0, i, 0, 0), y = geo_coord(g, 1, j, 0, 0), z = geo_coord(g, 2, k, 0, 0);
		int o = 0;
		while (o < g->nobjects && !geo_inside(g->nodes, g->aux, g->roots[o], x, y, z)) o++;
		if (o == g->nobjects) continue;
		int32_t *b = bounds + 6*o;
		if (i < b[0]) b[0] = i;
		if (i > b[1]) b[1] = i;
		if (j < b[2]) b[2] = j;
		if (j > b[3]) b[3] = j;
		if (k < b[4]) b[4] = k;
		if (k > b[5]) b[5] = k;
	}
	return ALISB_OK;
}

int alisb_geometry_fill(int device, const alisb_geometry *g, int object, const int32_t ibox[6], float *fx, float *fy, float *fz)
{
	(void)device;
	int rc = geo_check(g, ibox);
	if (rc) return rc;
	if (object < 0 || object >= g->nobjects || !fx || !fy || !fz) return fail(ALISB_EINVAL, "geometry: bad object or null output");
	const int nj = ibox[3]-ibox[2]+1, nk = ibox[5]-ibox[4]+1;
	#pragma omp parallel for schedule(guided)
	for (int i=ibox[0]; i<=ibox[1]; i++) for (int j=ibox[2]; j<=ibox[3]; j++) for (int k=ibox[4]; k<=ibox[5]; k++) {
		const size_t q = ((size_t)(i-ibox[0])*nj + (j-ibox[2]))*nk + (k-ibox[4]);
		const int idx[3] = { i, j, k };
		const int hx[3] = { 1, 0, 0 }, hy[3] = { 0, 1, 0 }, hz[3] = { 0, 0, 1 };
		fx[q] = geo_fill(g, object, idx, hx);
		fy[q] = geo_fill(g, object, idx, hy);
		fz[q] = geo_fill(g, object, idx, hz);
	}
	return ALISB_OK;
}

/* ---- halo ------------------------------------------------------------------------------ */
/* ---- chi2 mixing: reference update.c:344-411 (coupleTWM), sweep by sweep ---------------------- */
static double chi2_at(alisb_engine *g, int m, int u, int v, int a, int b, int c)
{
	/* component u of g's E field seen from the Yee position of component v.  u == v: the value
	 * itself (float); else the 4-point average 0.25*(f0+f1+f2+f3) of update.c:347-352, where the
	 * sum is float arithmetic and the product with the double literal 0.25 is double.          */
	const float *G = g->F[ALISB_SET_E][m][u];
	int o[3] = {a, b, c};
	if (u == v) return (double)G[IDX(g, a, b, c)];
	int p1[3] = {o[0], o[1], o[2]}, p2[3] = {o[0], o[1], o[2]}, p3[3] = {o[0], o[1], o[2]};
	p1[v] -= 1; p2[u] += 1; p3[v] -= 1; p3[u] += 1;
	const float sum = ((G[IDX(g,a,b,c)] + G[IDX(g,p1[0],p1[1],p1[2])]) + G[IDX(g,p2[0],p2[1],p2[2])]) + G[IDX(g,p3[0],p3[1],p3[2])];
	return 0.25 * sum;
}

static void chi2_sweep(alisb_engine *e, const dispobj *o, int m, alisb_engine *g1, alisb_engine *g2, int v, int u1, int u2, float dq, float alt)
{
	if (!dq) return;                                                       /* twm(): if (C->M[n].d) */
	float *F = e->F[ALISB_SET_E][m][v];
	const float *Cv = e->eps[v];
	for (int a=o->bx.lo[0]; a<=o->bx.hi[0]-1; a++)
	for (int b=o->bx.lo[1]; b<=o->bx.hi[1]-1; b++)
	for (int c=o->bx.lo[2]; c<=o->bx.hi[2]-1; c++) {
		const size_t i = IDX(e, a, b, c);
		const float pre = 2 * alt * dq * Cv[i] * o->fill[v].v[bidx(&o->fill[v], a, b, c)];
		if (u1 == v && u2 == v) {
			const float G1 = g1->F[ALISB_SET_E][m][v][IDX(g1,a,b,c)], G2 = g2->F[ALISB_SET_E][m][v][IDX(g2,a,b,c)];
			F[i] += pre * (G1 * G2);
		} else {
			F[i] += pre * (chi2_at(g1, m, u1, v, a, b, c) * chi2_at(g2, m, u2, v, a, b, c));
		}
	}
}

int alisb_chi2_couple(alisb_engine *e, int obj, int m, alisb_engine *p1, alisb_engine *p2, const float *d, float alt)
{
	if (!e || !p1 || !p2 || !d || obj < 0 || obj >= e->nobj || m < 0 || m > e->g.complex_field) return fail(ALISB_EINVAL, "chi2: bad argument");
	if (!e->eps[0]) return fail(ALISB_EINVAL, "chi2: the signal world has no material arrays");
	for (int q=0; q<3; q++) if (p1->n[q] != e->n[q] || p2->n[q] != e->n[q]) return fail(ALISB_EINVAL, "chi2: grids differ");
	const dispobj *o = e->obj[obj];
	/* Voigt pairs in the order of update.c:364-390 */
	static const int U[9][2] = {{0,0},{1,1},{2,2},{1,2},{2,1},{2,0},{0,2},{0,1},{1,0}};
	static const int Q[9]    = {0, 1, 2, 3, 3, 4, 4, 5, 5};
	for (int v=0; v<3; v++) for (int s=0; s<9; s++)
		chi2_sweep(e, o, m, p1, p2, v, U[s][0], U[s][1], d[v*6 + Q[s]], alt);
	return ALISB_OK;
}

int alisb_halo_info(alisb_engine *e, int set, int m, int comp, void **send_ptr, void **recv_ptr, size_t *bytes)
{
	if (set<0||set>1||m<0||m>e->g.complex_field||comp<0||comp>1) return fail(ALISB_EINVAL, "bad halo id");
	if (send_ptr) *send_ptr = e->halo_send[set][m][comp];
	if (recv_ptr) *recv_ptr = e->halo_recv[set][m][comp];
	if (bytes) *bytes = (size_t)(e->n[0]+2)*(e->n[1]+2)*sizeof(float);
	return ALISB_OK;
}
static void halo_copy(alisb_engine *e, int set, int pack)
{
	const int c = pack ? (set==ALISB_SET_E ? e->n[2] : 1) : (set==ALISB_SET_E ? 0 : e->n[2]+1);
	for (int m=0; m<=e->g.complex_field; m++) for (int k=0; k<2; k++) {
		float *F = e->F[set][m][k], *buf = pack ? e->halo_send[set][m][k] : e->halo_recv[set][m][k];
		size_t q = 0;
		for (int b=0; b<=e->n[1]+1; b++) for (int a=0; a<=e->n[0]+1; a++, q++) {   /* x fastest, like the device layout */
			if (pack) buf[q] = F[IDX(e,a,b,c)]; else F[IDX(e,a,b,c)] = buf[q];
		}
	}
}
int alisb_halo_pack(alisb_engine *e, int set) { halo_copy(e, set, 1); return ALISB_OK; }
int alisb_halo_unpack(alisb_engine *e, int set) { halo_copy(e, set, 0); return ALISB_OK; }
int alisb_halo_exchange(alisb_engine *lower, alisb_engine *upper, int set)
{
	if (!lower || !upper || set < 0 || set > 1) return fail(ALISB_EINVAL, "halo exchange: bad argument");
	if (lower->n[0] != upper->n[0] || lower->n[1] != upper->n[1] || lower->g.complex_field != upper->g.complex_field) return fail(ALISB_EINVAL, "halo exchange: slabs differ in x, y or field sets");
	alisb_engine *src = set == ALISB_SET_E ? lower : upper, *dst = set == ALISB_SET_E ? upper : lower;
	const size_t bytes = (size_t)(src->n[0]+2)*(src->n[1]+2)*sizeof(float);
	halo_copy(src, set, 1);
	for (int m=0; m<=src->g.complex_field; m++) for (int k=0; k<2; k++) memcpy(dst->halo_recv[set][m][k], src->halo_send[set][m][k], bytes);
	halo_copy(dst, set, 0);
	return ALISB_OK;
}

/* ---- phasers: reference phaser.c:83-110 ------------------------------------------------ */
int alisb_phaser_create(alisb_engine *e, const int32_t box[6], int face_mask, int nfreq, alisb_phaser **out)
{
	if (nfreq < 1) return fail(ALISB_EINVAL, "nfreq");
	alisb_phaser *p = calloc(1, sizeof *p);
	p->e = e; box_from(&p->bx, box); p->mask = face_mask; p->nfreq = nfreq;
	for (int f=0; f<6; f++) if (face_mask & (1<<f)) {
		int lo[3], hi[3], ax=f/2, sd=f%2;
		for (int d=0; d<3; d++) lo[d]=p->bx.lo[d], hi[d]=p->bx.hi[d];
		lo[ax] = hi[ax] = sd ? p->bx.hi[ax] : p->bx.lo[ax];
		for (int w=0; w<4; w++) {
			block *B = &p->acc[f][w];
			for (int d=0; d<3; d++) { B->bx.lo[d]=lo[d]; B->bx.hi[d]=hi[d]; B->n[d]=hi[d]-lo[d]+1; }
			B->v = calloc((size_t)B->n[0]*B->n[1]*B->n[2]*2*nfreq, sizeof(float));
		}
	}
	*out = p;
	return ALISB_OK;
}
void alisb_phaser_destroy(alisb_phaser *p)
{
	if (!p) return;
	for (int f=0; f<6; f++) for (int w=0; w<4; w++) free(p->acc[f][w].v);
	free(p);
}
int alisb_phaser_accumulate(alisb_phaser *p, const float *expE, const float *expH)
{
	alisb_engine *e = p->e;
	for (int f=0; f<6; f++) if (p->mask & (1<<f)) {
		const int ax=f/2, c1=(ax+1)%3, c2=(ax+2)%3;
		const float *E1=e->F[0][0][c1], *E2=e->F[0][0][c2], *H1=e->F[1][0][c1], *H2=e->F[1][0][c2];
		block *B = p->acc[f];
		const size_t cells = (size_t)B->n[0]*B->n[1]*B->n[2];
		int d[3]={0,0,0}; d[ax]=1;
		for (int q=0; q<p->nfreq; q++) {
			const float eEr=expE[2*q], eEi=expE[2*q+1], eHr=expH[2*q], eHi=expH[2*q+1];
			float *A0=B[0].v+2*cells*q, *A1=B[1].v+2*cells*q, *A2=B[2].v+2*cells*q, *A3=B[3].v+2*cells*q;
			for (int a=B->bx.lo[0]; a<=B->bx.hi[0]; a++) for (int b=B->bx.lo[1]; b<=B->bx.hi[1]; b++) for (int c=B->bx.lo[2]; c<=B->bx.hi[2]; c++) {
				const size_t i=IDX(e,a,b,c), j=IDX(e,a+d[0],b+d[1],c+d[2]), s=2*bidx(B,a,b,c);
				/* P.E += expiwtE * E  : (float complex)*(float), componentwise, float add */
				A0[s] += eEr*E1[i]; A0[s+1] += eEi*E1[i];
				A1[s] += eEr*E2[i]; A1[s+1] += eEi*E2[i];
				/* P.H += expiwtH * (H+H') * 0.5 : the 0.5 is a double constant, so the product
				 * and the accumulation are done in double and rounded once (phaser.c:89-90) */
				const float h1 = H1[i]+H1[j], h2 = H2[i]+H2[j];
				A2[s]   = (float)((double)A2[s]   + (double)(eHr*h1)*0.5);
				A2[s+1] = (float)((double)A2[s+1] + (double)(eHi*h1)*0.5);
				A3[s]   = (float)((double)A3[s]   + (double)(eHr*h2)*0.5);
				A3[s+1] = (float)((double)A3[s+1] + (double)(eHi*h2)*0.5);
			}
		}
	}
	return ALISB_OK;
}
int alisb_phaser_read(alisb_phaser *p, int freq, int face, int which, float *host)
{
	if (face<0||face>5||which<0||which>3||freq<0||freq>=p->nfreq||!(p->mask&(1<<face))) return fail(ALISB_EINVAL, "bad phaser id");
	block *B = &p->acc[face][which];
	const size_t cells = (size_t)B->n[0]*B->n[1]*B->n[2];
	memcpy(host, B->v + 2*cells*freq, 2*cells*sizeof(float));
	return ALISB_OK;
}

/* far-field surface sums (reference farfld.c:248-459 inner loops, weights folded by the host):
 * fp32 products A*w_hi then *w_lo, double accumulation in the reference's loop nesting (u outer
 * loop = the face's lower in-face axis is the OUTER loop of the z-fastest block here) */
int alisb_phaser_surface_sums(alisb_phaser *p, int freq, const alisb_surface_job *jobs, int njobs, const float *weights, size_t nweights, double *out)
{
	if (!p || freq<0 || freq>=p->nfreq || njobs<0) return fail(ALISB_EINVAL, "bad surface-sum request");
	for (int q=0; q<njobs; q++) {
		const alisb_surface_job *J = &jobs[q];
		if (J->face<0 || J->face>5 || J->which<0 || J->which>3 || !(p->mask & (1<<J->face))) return fail(ALISB_EINVAL, "surface sum: face is not accumulated");
		const block *B = &p->acc[J->face][J->which];
		const int ax = J->face/2;
		if ((size_t)J->w_lo + B->n[ax==0 ? 1 : 0] > nweights || (size_t)J->w_hi + B->n[ax==2 ? 1 : 2] > nweights) return fail(ALISB_EINVAL, "surface sum: weight vector outside the buffer");
	}
	#pragma omp parallel for schedule(static)
	for (int q=0; q<njobs; q++) {
		const alisb_surface_job *J = &jobs[q];
		const block *B = &p->acc[J->face][J->which];
		const int ax = J->face/2, lo = ax==0 ? 1 : 0, hi = ax==2 ? 1 : 2;
		const int nlo = B->n[lo], nhi = B->n[hi];
		const size_t cells = (size_t)B->n[0]*B->n[1]*B->n[2];
		const float *A = B->v + 2*cells*freq, *Wlo = weights + 2*(size_t)J->w_lo, *Whi = weights + 2*(size_t)J->w_hi;
		double sr = 0, si = 0;
		for (int u=0; u<nlo; u++) for (int v=0; v<nhi; v++) {
			const size_t c = 2*((size_t)u*nhi + v);          /* z-fastest block: the higher axis is fastest */
			const float ar = A[c], ai = A[c+1], br = Whi[2*v], bi = Whi[2*v+1];
			const float tr = ar*br - ai*bi, ti = ar*bi + ai*br;
			const float cr = Wlo[2*u], ci = Wlo[2*u+1];
			sr += (double)(tr*cr - ti*ci); si += (double)(tr*ci + ti*cr);
		}
		out[2*q] = sr; out[2*q+1] = si;
	}
	return ALISB_OK;
}

uint64_t alisb_launch_count(alisb_engine *e) { return e->launches; }
int alisb_profile_dominant(alisb_engine *e, float *ms, int *n, double *b) { (void)e; if (ms) *ms=0; if (n) *n=0; if (b) *b=0; return ALISB_OK; }
int alisb_set_profiling(alisb_engine *e, int on) { (void)e; (void)on; return ALISB_OK; }
int alisb_last_run_kernel_ms(alisb_engine *e, float *ms, int *n) { (void)e; if (ms) *ms=0; if (n) *n=0; return ALISB_OK; }
