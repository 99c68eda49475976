/*
 * engine_bridge.c -- translates a host `world` (the ALiS public structs) into the POD
 * descriptors of the engine C ABI (include/alis_b200.h) and keeps the two in sync.
 *
 * It only touches struct members that are part of the public, source-compatible API
 * (reference alis_c.h:88-130,305-318), so the same file compiles
 *   (a) against include/alis_c.h   -> part of the product library (updateE/updateH), and
 *   (b) against the reference's own alis_c.h (oracle/harness, -DALIS_HARNESS_REF) -> lets the
 *       tests drive the CUDA engine and the CPU oracle from worlds that the UNMODIFIED
 *       reference set up, isolating the time-stepping path from our own setup code.
 *
 * What is restated here from the reference: the index ranges of the six PML slabs
 * (update.c:119-124), the ordered boundary-copy call lists (update.c:284-297,311-339) and
 * the 12+12 TFSF face strips (update.c:205-272) -- as data handed to the engine.
 */
#include <limits.h>
#include <string.h>
#include "alis_c.h"
#include "alis_b200.h"
#include "engine_bridge.h"
#include "shard.h"

typedef struct wave_ctx { world W; source S; } wave_ctx;

struct alis_bridge {
	world W;
	alis_shards *S;                 /* one engine, or one per device: z slabs of W (ALIS_DEVICES), or replicas of an auxiliary world */
	alisb_engine *E;                /* the first (with one device: the only) engine */
	int devices[16];
	struct alis_bridge *aux[8]; int naux;
	wave_ctx *ctx; int nctx, cap_ctx;
	int dirty;                      /* DIRTY_* bits: what must be re-uploaded before the next step */
	int owns_aux;
};

enum { DIRTY_TABLES = 1, DIRTY_MATERIALS = 2, DIRTY_SOURCES = 4, DIRTY_COPIES = 8, DIRTY_ALL = 15 };

#ifdef ALIS_B200_HOST
alis_bridge *alis_host_bridge(world W);        /* host layer: every world owns its bridge */
#endif

/* how the next bridge is created: a world of its own is cut into z slabs over the devices ALIS_DEVICES names;
 * the auxiliary world of a TFSF source gets one full copy per engine of its main world (push_tfsf) */
static int create_mode = ALIS_SHARDS_SPLIT, create_n = 0;
static const int *create_dev = 0;

static float wave_trampoline(void *user, float t, float phase)
{
	wave_ctx *c = (wave_ctx*)user;
	return c->S->waveForm(c->W, c->S, t, phase);
}

static void die(const char *what)
{
	printf("alis-b200 engine: %s: %s\n", what, alisb_last_error());
	exit(1);
}
#define CHECK(call) do { if ((call) != ALISB_OK) die(#call); } while (0)

static void set_box(int32_t b[6], world W, int i0, int i1, int j0, int j1, int k0, int k1)
{
	b[0] = i0-(W->iMIN-1); b[1] = i1-(W->iMIN-1);
	b[2] = j0-(W->jMIN-1); b[3] = j1-(W->jMIN-1);
	b[4] = k0-(W->kMIN-1); b[5] = k1-(W->kMIN-1);
}

/* 1-D table indexed MIN..MAX on the host -> n+2 floats in array coordinates */
static float *table(const float *t, int MIN, int NUM)
{
	float *p = calloc(NUM+2, sizeof(float));
	if (t) for (int q=1; q<=NUM; q++) p[q] = t[MIN-1+q];
	return p;
}

static void push_tables(alis_bridge *B, int set, coeffs *C)
{
	world W = B->W;
	const float *dtd[3] = {C->dtdx, C->dtdy, C->dtdz};
	const float *P[3] = {C->xP, C->yP, C->zP};
	const float *P1[3] = {C->xPy, C->yPz, C->zPx};
	const float *P2[3] = {C->xPz, C->yPx, C->zPy};
	const int MIN[3] = {W->iMIN, W->jMIN, W->kMIN}, NUM[3] = {W->iNUM, W->jNUM, W->kNUM};
	for (int ax=0; ax<3; ax++) {
		float *a = table(dtd[ax], MIN[ax], NUM[ax]);
		float *b = P[ax] ? table(P[ax], MIN[ax], NUM[ax]) : 0;
		float *c = P[ax] ? table(P1[ax], MIN[ax], NUM[ax]) : 0;
		float *d = P[ax] ? table(P2[ax], MIN[ax], NUM[ax]) : 0;
		CHECK(alis_shards_set_axis_tables(B->S, set, ax, a, b, c, d));
		free(a); free(b); free(c); free(d);
	}
	/* cell sizes, for the device-side flux integrals (reference output.c:233-332) */
	const float *ds[3] = {C->dx, C->dy, C->dz};
	for (int ax=0; ax<3; ax++) {
		float *a = table(ds[ax], MIN[ax], NUM[ax]);
		CHECK(alis_shards_set_spacing(B->S, set, ax, a));
		free(a);
	}
}

static void fill_grid(world W, alisb_grid *g)
{
	memset(g, 0, sizeof *g);
	g->nx = W->iNUM; g->ny = W->jNUM; g->nz = W->kNUM;
	g->complex_field = W->complexField;
	g->dt = W->dt;
	g->kshard_lo = 1; g->nz_global = W->kNUM;
	const int MIN[3] = {W->iMIN, W->jMIN, W->kMIN}, Min[3] = {W->iMin, W->jMin, W->kMin};
	const int MAX[3] = {W->iMAX, W->jMAX, W->kMAX}, Max[3] = {W->iMax, W->jMax, W->kMax};
	for (int set=0; set<2; set++) for (int ax=0; ax<3; ax++) {
		const int Alt = set==0 ? 1 : -1;
		/* update.c:119-121 (min side), :122-124 (max side) */
		g->pml[set][ax][0][0] = 1; g->pml[set][ax][0][1] = 0;
		g->pml[set][ax][1][0] = 1; g->pml[set][ax][1][1] = 0;
		if (MIN[ax] < Min[ax]) {
			g->pml[set][ax][0][0] = (MIN[ax]-(Alt-1)/2) - (MIN[ax]-1);
			g->pml[set][ax][0][1] = (Min[ax]-(Alt+3)/2) - (MIN[ax]-1);
		}
		if (MAX[ax] > Max[ax]) {
			g->pml[set][ax][1][0] = (Max[ax]+2) - (MIN[ax]-1);
			g->pml[set][ax][1][1] = MAX[ax] - (MIN[ax]-1);
		}
	}
}

static void add_copy(alisb_copy *L, int *n, world W, int set, int comp, float cosk, float sink,
                     int i0, int i1, int j0, int j1, int k0, int k1, int ia, int ja, int ka)
{
	alisb_copy *c = &L[(*n)++];
	c->set = set; c->comp = comp; c->cosk = cosk; c->sink = sink;
	set_box(c->box, W, i0, i1, j0, j1, k0, k1);
	c->alt[0] = ia; c->alt[1] = ja; c->alt[2] = ka;
}

/* the ordered copy lists of update.c:284-297 (E) and :311-339 (H) */
static void push_copies(alis_bridge *B)
{
	world W = B->W;
	alisb_copy L[24];
	int n = 0;
	const int X=0, Y=1, Z=2, E=ALISB_SET_E, H=ALISB_SET_H;
	const int iAlt = (W->xMinSur==PBC || W->xMinSur==BBC) ? W->iNUM-1 : 1;
	const int jAlt = (W->yMinSur==PBC || W->yMinSur==BBC) ? W->jNUM-1 : 1;
	const int kAlt = (W->zMinSur==PBC || W->zMinSur==BBC) ? W->kNUM-1 : 1;
	const int i0=W->iMIN, i1=W->iMAX, j0=W->jMIN, j1=W->jMAX, k0=W->kMIN, k1=W->kMAX;

	add_copy(L,&n,W, E,X, W->coskx,-W->sinkx, i1+1,i1+1, j0,j1, k0,k1, -iAlt,0,0);
	add_copy(L,&n,W, E,Z, W->coskz,-W->sinkz, i0,i1, j0,j1, k1+1,k1+1, 0,0,-kAlt);
	if (W->yMinSur==HBC) {
		add_copy(L,&n,W, E,Y, W->coskym,-W->sinkym, i0,0,    j1+1,j1+1, k0,k1, i1,1-W->jNUM,0);
		add_copy(L,&n,W, E,Y, W->coskyp,-W->sinkyp, 1,i1+1,  j1+1,j1+1, k0,k1, i0,1-W->jNUM,0);
	}
	else add_copy(L,&n,W, E,Y, W->cosky,-W->sinky, i0,i1, j1+1,j1+1, k0,k1, 0,-jAlt,0);
	CHECK(alis_shards_set_copies(B->S, E, L, n));

	n = 0;
	const float cxm = W->xMaxSur==PEC ? 1 : W->coskx, cym = W->yMaxSur==PEC ? 1 : W->cosky, czm = W->zMaxSur==PEC ? 1 : W->coskz;
	add_copy(L,&n,W, H,Y, cxm,-W->sinkx, i1+1,i1+1, j0,j1+1, k0,k1+1, -iAlt,0,0);
	add_copy(L,&n,W, H,Z, cxm,-W->sinkx, i1+1,i1+1, j0,j1+1, k0,k1+1, -iAlt,0,0);
	add_copy(L,&n,W, H,X, czm,-W->sinkz, i0,i1+1, j0,j1+1, k1+1,k1+1, 0,0,-kAlt);
	add_copy(L,&n,W, H,Y, czm,-W->sinkz, i0,i1+1, j0,j1+1, k1+1,k1+1, 0,0,-kAlt);
	add_copy(L,&n,W, H,Y, W->coskx, W->sinkx, i0,i0, j0,j1+1, k0,k1+1, iAlt,0,0);
	add_copy(L,&n,W, H,Z, W->coskx, W->sinkx, i0,i0, j0,j1+1, k0,k1+1, iAlt,0,0);
	add_copy(L,&n,W, H,X, W->coskz, W->sinkz, i0,i1+1, j0,j1+1, k0,k0, 0,0,kAlt);
	add_copy(L,&n,W, H,Y, W->coskz, W->sinkz, i0,i1+1, j0,j1+1, k0,k0, 0,0,kAlt);
	if (W->yMinSur==HBC) {
		const int up = W->jNUM-1, dn = 1-W->jNUM;
		add_copy(L,&n,W, H,X, W->coskyp, W->sinkyp, i0,0,   j0,j0,     k0,k1, i1,up,0);
		add_copy(L,&n,W, H,Z, W->coskyp, W->sinkyp, i0,0,   j0,j0,     k0,k1, i1,up,0);
		add_copy(L,&n,W, H,X, W->coskym,-W->sinkym, i0,0,   j1+1,j1+1, k0,k1, i1,dn,0);
		add_copy(L,&n,W, H,Z, W->coskym,-W->sinkym, i0,0,   j1+1,j1+1, k0,k1, i1,dn,0);
		add_copy(L,&n,W, H,X, W->coskym, W->sinkym, 1,i1+1, j0,j0,     k0,k1, i0,up,0);
		add_copy(L,&n,W, H,Z, W->coskym, W->sinkym, 1,i1+1, j0,j0,     k0,k1, i0,up,0);
		add_copy(L,&n,W, H,X, W->coskyp,-W->sinkyp, 1,i1+1, j1+1,j1+1, k0,k1, i0,dn,0);
		add_copy(L,&n,W, H,Z, W->coskyp,-W->sinkyp, 1,i1+1, j1+1,j1+1, k0,k1, i0,dn,0);
	}
	else {
		add_copy(L,&n,W, H,Z, cym,-W->sinky, i0,i1+1, j1+1,j1+1, k0,k1+1, 0,-jAlt,0);
		add_copy(L,&n,W, H,X, cym,-W->sinky, i0,i1+1, j1+1,j1+1, k0,k1+1, 0,-jAlt,0);
		add_copy(L,&n,W, H,Z, W->cosky, W->sinky, i0,i1+1, j0,j0, k0,k1+1, 0,jAlt,0);
		add_copy(L,&n,W, H,X, W->cosky, W->sinky, i0,i1+1, j0,j0, k0,k1+1, 0,jAlt,0);
	}
	CHECK(alis_shards_set_copies(B->S, H, L, n));
}

static void push_materials(alis_bridge *B)
{
	world W = B->W;
	coeffs *C = W->CE;
	CHECK(alis_shards_clear_dispersive(B->S));
	if (C->x && C->y && C->z)
		CHECK(alis_shards_set_inv_eps(B->S, &C->x[W->iMIN][W->jMIN][W->kMIN], &C->y[W->iMIN][W->jMIN][W->kMIN], &C->z[W->iMIN][W->jMIN][W->kMIN]));
	else
		CHECK(alis_shards_set_inv_eps(B->S, 0, 0, 0));
	for (int o=0; o<C->N; o++) {
		alisb_dispersive d;
		memset(&d, 0, sizeof d);
		set_box(d.box, W, C->iMin[o], C->iMax[o], C->jMin[o], C->jMax[o], C->kMin[o], C->kMax[o]);
		d.nj = C->NJ[o]; d.nk = C->NK[o];
		d.a[0] = C->ax[o]; d.a[1] = C->ay[o]; d.a[2] = C->az[o];
		float **JJ[3] = {C->JJx, C->JJy, C->JJz}, **JK[3] = {C->JKx, C->JKy, C->JKz}, **JE[3] = {C->JEx, C->JEy, C->JEz};
		float **KJ[3] = {C->KJx, C->KJy, C->KJz}, **KK[3] = {C->KKx, C->KKy, C->KKz}, **KE[3] = {C->KEx, C->KEy, C->KEz};
		for (int c=0; c<3; c++) for (int p=0; p<d.nk; p++) {
			d.JJ[c][p] = JJ[c][o][p]; d.JE[c][p] = JE[c][o][p];
			if (p >= d.nj) { d.JK[c][p] = JK[c][o][p]; d.KJ[c][p] = KJ[c][o][p]; d.KK[c][p] = KK[c][o][p]; d.KE[c][p] = KE[c][o][p]; }
		}
		d.fill[0] = &C->Jx[o][C->iMin[o]][C->jMin[o]][C->kMin[o]];
		d.fill[1] = &C->Jy[o][C->iMin[o]][C->jMin[o]][C->kMin[o]];
		d.fill[2] = &C->Jz[o][C->iMin[o]][C->jMin[o]][C->kMin[o]];
		CHECK(alis_shards_add_dispersive(B->S, &d));
	}
}

static void push_source_list(alis_bridge *B, source S, int set)
{
	world W = B->W;
	for (; S; S = S->S) {
		alisb_source d;
		memset(&d, 0, sizeof d);
		int comp = -1, m = 0;
		if (alis_source_target(W, S, set, &comp, &m)) { printf("alis-b200 engine: source with unknown target array\n"); exit(1); }
		d.set = set; d.comp = comp; d.m = m;
		set_box(d.box, W, S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);
		d.amplitude = S->amplitude; d.phase = S->phase;
		d.Amplitude = S->Amplitude ? &S->Amplitude[S->iMin][S->jMin][S->kMin] : 0;
		d.Phase = S->Phase ? &S->Phase[S->iMin][S->jMin][S->kMin] : 0;
		d.n_active = S->N;
		if (B->nctx == B->cap_ctx) { printf("alis-b200 engine: too many sources\n"); exit(1); }
		wave_ctx *c = &B->ctx[B->nctx++];
		c->W = W; c->S = S;
		d.waveform = wave_trampoline; d.user = c;
		CHECK(alis_shards_add_source(B->S, &d));
	}
}

static void add_strip(alisb_strip *L, int *n, world W, int comp, int aux_comp, float coef, int use_eps,
                      int i0, int i1, int j0, int j1, int k0, int k1, int ia, int ja, int ka)
{
	alisb_strip *s = &L[(*n)++];
	s->comp = comp; s->aux_comp = aux_comp; s->coef = coef; s->use_eps = use_eps;
	set_box(s->box, W, i0, i1, j0, j1, k0, k1);
	s->alt[0] = ia; s->alt[1] = ja; s->alt[2] = ka;
}

/* the face-strip lists of update.c:209-234 (E) and :244-269 (H) */
static void push_tfsf(alis_bridge *B, tfsf S)
{
	world W = B->W, A = S->W;
	const int X=0, Y=1, Z=2;
	alisb_strip LE[12], LH[12];
	int nE = 0, nH = 0;
	const int ue = (W->CE->x && W->CE->y && W->CE->z) ? 1 : 0;
	const float *ex = W->CE->dtdx, *ey = W->CE->dtdy, *ez = W->CE->dtdz;
	const float *hx = W->CH->dtdx, *hy = W->CH->dtdy, *hz = W->CH->dtdz;

	if (S->iMin > W->iMIN) {
		add_strip(LE,&nE,W, Z,Y, -ex[S->iMin], ue, S->iMin,S->iMin, S->jMin,S->jMax, S->kMIN,S->kMAX, 0,0,0);
		add_strip(LE,&nE,W, Y,Z, +ex[S->iMin], ue, S->iMin,S->iMin, S->jMIN,S->jMAX, S->kMin,S->kMax, 0,0,0);
		add_strip(LH,&nH,W, Z,Y, +hx[S->iMin], 0,  S->iMin,S->iMin, S->jMIN,S->jMAX, S->kMin,S->kMax, 0,0,0);
		add_strip(LH,&nH,W, Y,Z, -hx[S->iMin], 0,  S->iMin,S->iMin, S->jMin,S->jMax, S->kMIN,S->kMAX, 0,0,0);
	}
	if (S->jMin > W->jMIN) {
		add_strip(LE,&nE,W, X,Z, -ey[S->jMin], ue, S->iMIN,S->iMAX, S->jMin,S->jMin, S->kMin,S->kMax, 0,0,0);
		add_strip(LE,&nE,W, Z,X, +ey[S->jMin], ue, S->iMin,S->iMax, S->jMin,S->jMin, S->kMIN,S->kMAX, 0,0,0);
		add_strip(LH,&nH,W, X,Z, +hy[S->jMin], 0,  S->iMin,S->iMax, S->jMin,S->jMin, S->kMIN,S->kMAX, 0,0,0);
		add_strip(LH,&nH,W, Z,X, -hy[S->jMin], 0,  S->iMIN,S->iMAX, S->jMin,S->jMin, S->kMin,S->kMax, 0,0,0);
	}
	if (S->kMin > W->kMIN) {
		add_strip(LE,&nE,W, Y,X, -ez[S->kMin], ue, S->iMin,S->iMax, S->jMIN,S->jMAX, S->kMin,S->kMin, 0,0,0);
		add_strip(LE,&nE,W, X,Y, +ez[S->kMin], ue, S->iMIN,S->iMAX, S->jMin,S->jMax, S->kMin,S->kMin, 0,0,0);
		add_strip(LH,&nH,W, Y,X, +hz[S->kMin], 0,  S->iMIN,S->iMAX, S->jMin,S->jMax, S->kMin,S->kMin, 0,0,0);
		add_strip(LH,&nH,W, X,Y, -hz[S->kMin], 0,  S->iMin,S->iMax, S->jMIN,S->jMAX, S->kMin,S->kMin, 0,0,0);
	}
	if (S->iMax < W->iMAX) {
		add_strip(LE,&nE,W, Z,Y, +ex[S->iMax], ue, S->iMax,S->iMax, S->jMin,S->jMax, S->kMIN,S->kMAX, 1,0,0);
		add_strip(LE,&nE,W, Y,Z, -ex[S->iMax], ue, S->iMax,S->iMax, S->jMIN,S->jMAX, S->kMin,S->kMax, 1,0,0);
		add_strip(LH,&nH,W, Z,Y, -hx[S->iMax+1], 0, S->iMax+1,S->iMax+1, S->jMIN,S->jMAX, S->kMin,S->kMax, -1,0,0);
		add_strip(LH,&nH,W, Y,Z, +hx[S->iMax+1], 0, S->iMax+1,S->iMax+1, S->jMin,S->jMax, S->kMIN,S->kMAX, -1,0,0);
	}
	if (S->jMax < W->jMAX) {
		add_strip(LE,&nE,W, X,Z, +ey[S->jMax], ue, S->iMIN,S->iMAX, S->jMax,S->jMax, S->kMin,S->kMax, 0,1,0);
		add_strip(LE,&nE,W, Z,X, -ey[S->jMax], ue, S->iMin,S->iMax, S->jMax,S->jMax, S->kMIN,S->kMAX, 0,1,0);
		add_strip(LH,&nH,W, X,Z, -hy[S->jMax+1], 0, S->iMin,S->iMax, S->jMax+1,S->jMax+1, S->kMIN,S->kMAX, 0,-1,0);
		add_strip(LH,&nH,W, Z,X, +hy[S->jMax+1], 0, S->iMIN,S->iMAX, S->jMax+1,S->jMax+1, S->kMin,S->kMax, 0,-1,0);
	}
	if (S->kMax < W->kMAX) {
		add_strip(LE,&nE,W, Y,X, +ez[S->kMax], ue, S->iMin,S->iMax, S->jMIN,S->jMAX, S->kMax,S->kMax, 0,0,1);
		add_strip(LE,&nE,W, X,Y, -ez[S->kMax], ue, S->iMIN,S->iMAX, S->jMin,S->jMax, S->kMax,S->kMax, 0,0,1);
		add_strip(LH,&nH,W, Y,X, -hz[S->kMax+1], 0, S->iMIN,S->iMAX, S->jMin,S->jMax, S->kMax+1,S->kMax+1, 0,0,-1);
		add_strip(LH,&nH,W, X,Y, +hz[S->kMax+1], 0, S->iMin,S->iMax, S->jMIN,S->jMAX, S->kMax+1,S->kMax+1, 0,0,-1);
	}

	/* the auxiliary world: one engine per engine of this world (every z slab steps its own copy) */
	const int outer_mode = create_mode, outer_n = create_n;
	const int *outer_dev = create_dev;
	create_mode = ALIS_SHARDS_REPLICA; create_n = alis_shards_count(B->S); create_dev = B->devices;
#ifdef ALIS_B200_HOST
	alis_bridge *AB = alis_host_bridge(A);      /* owned by the auxiliary world itself */
	B->owns_aux = 0;
#else
	alis_bridge *AB = alis_bridge_create(A);
	B->owns_aux = 1;
#endif
	create_mode = outer_mode; create_n = outer_n; create_dev = outer_dev;
	if (alis_shards_count(AB->S) != alis_shards_count(B->S) || (alis_shards_count(B->S) > 1 && alis_shards_mode(AB->S) != ALIS_SHARDS_REPLICA)) {
		printf("alis-b200 engine: the auxiliary world of a TFSF source was stepped on its own before the main world was first used; with ALIS_DEVICES naming several devices that is not supported\n");
		exit(1);
	}
	alis_bridge_commit(AB);
	if (B->naux >= 8) { printf("alis-b200 engine: too many TFSF regions\n"); exit(1); }
	B->aux[B->naux++] = AB;

	alisb_tfsf t;
	memset(&t, 0, sizeof t);
	t.aux = AB->E;                              /* the shard layer substitutes the copy that lives with each slab */
	t.offset[0] = W->iMIN - A->iMIN; t.offset[1] = W->jMIN - A->jMIN; t.offset[2] = W->kMIN - A->kMIN;
	/* pointer aliasing of source.c:223-248,298-317: rows past the collapsed aux extent alias row
	 * MAX+Alt, Alt = 1 for the components staggered along that axis.                        */
	static const int altx[2][3] = {{1,0,0},{0,1,1}}, alty[2][3] = {{0,1,0},{1,0,1}};
	for (int s=0; s<2; s++) for (int c=0; c<3; c++) {
		t.clamp[s][c][0] = (A->iNUM < W->iNUM) ? A->iNUM + altx[s][c] : INT_MAX;
		t.clamp[s][c][1] = (A->jNUM < W->jNUM) ? A->jNUM + alty[s][c] : INT_MAX;
	}
	t.strips_E = LE; t.n_strips_E = nE;
	t.strips_H = LH; t.n_strips_H = nH;
	CHECK(alis_shards_add_tfsf(B->S, &t, AB->S));
}

alis_bridge *alis_bridge_create(world W)
{
	alis_bridge *B = calloc(1, sizeof *B);
	alisb_grid g;
	B->W = W;
	fill_grid(W, &g);
	int n = 0, mode = create_mode;
	if (create_n > 0) { n = create_n; memcpy(B->devices, create_dev, sizeof(int)*n); }       /* an auxiliary world follows its main world */
	else {
		n = alis_shards_env_devices(B->devices, 16);
		if (n == 1 && !getenv("ALIS_DEVICES")) B->devices[0] = alis_bridge_device();
		if (n > 1 && (W->zMinSur == PBC || W->zMinSur == BBC)) {
			printf("alis-b200: the world is periodic along z: not cut into z slabs, it runs on device %d alone\n", B->devices[0]);
			n = 1;
		}
	}
	g.device = B->devices[0];
	CHECK(alis_shards_create(&g, B->devices, n, mode, &B->S));
	B->E = alis_shards_engine(B->S, 0);
	B->cap_ctx = 512;
	B->ctx = calloc(B->cap_ctx, sizeof(wave_ctx));
	B->dirty = DIRTY_ALL;
	return B;
}

static void drop_aux(alis_bridge *B)
{
	if (B->owns_aux) for (int q=0; q<B->naux; q++) alis_bridge_destroy(B->aux[q]);
	B->naux = 0;
}

void alis_bridge_destroy(alis_bridge *B)
{
	if (!B) return;
	drop_aux(B);
	alis_shards_destroy(B->S);
	free(B->ctx);
	free(B);
}

alisb_engine *alis_bridge_engine(alis_bridge *B) { return B->E; }
int alis_bridge_engines(alis_bridge *B) { return alis_shards_count(B->S); }
alis_shards *alis_bridge_shards(alis_bridge *B) { return B->S; }
int alis_bridge_read_box(alis_bridge *B, int arr, const int32_t ids[6], const int32_t box[6], float *host) { return alis_shards_read_box(B->S, arr, ids, box, host); }
int alis_bridge_reset_fields(alis_bridge *B) { return alis_shards_reset_fields(B->S); }

/* (re)upload whatever changed since the last commit; the field state is never touched */
void alis_bridge_commit(alis_bridge *B)
{
	world W = B->W;
	if (!B->dirty) return;
	if (B->dirty & DIRTY_TABLES) {
		push_tables(B, ALISB_SET_E, W->CE);
		push_tables(B, ALISB_SET_H, W->CH);
	}
	if (B->dirty & DIRTY_MATERIALS) push_materials(B);
	if (B->dirty & (DIRTY_COPIES|DIRTY_SOURCES)) push_copies(B);
	if (B->dirty & DIRTY_SOURCES) {
		CHECK(alis_shards_clear_sources(B->S));
		drop_aux(B);
		B->nctx = 0;
		for (tfsf S = W->S; S; S = S->S) push_tfsf(B, S);
		push_source_list(B, W->SE, ALISB_SET_E);
		push_source_list(B, W->SH, ALISB_SET_H);
	}
	CHECK(alis_shards_set_time(B->S, W->n, W->t, W->E->t, W->H->t));
	B->dirty = 0;
}

void alis_bridge_mark(alis_bridge *B, int mask) { if (B) B->dirty |= mask; }

static void pull_time(alis_bridge *B)
{
	world W = B->W;
	uint64_t step; float t, tE, tH;
	CHECK(alis_shards_get_time(B->S, &step, &t, &tE, &tH));
	W->n = step; W->t = t; W->E->t = tE; W->H->t = tH;
	for (int q=0; q<B->naux; q++) pull_time(B->aux[q]);
}

void alis_bridge_step_E(alis_bridge *B)
{
	if (B->dirty) alis_bridge_commit(B);
	CHECK(alis_shards_step_E(B->S));
	pull_time(B);
}

void alis_bridge_step_H(alis_bridge *B)
{
	if (B->dirty) alis_bridge_commit(B);
	CHECK(alis_shards_step_H(B->S));
	pull_time(B);
}

/* nonLinearSHG / nonLinearTWM (update.c:394-411): object by object, field set by field set.
 * C->M is indexed with the ADE-object index n exactly as the reference does (update.c:354).   */
void alis_bridge_chi2(alis_bridge *S, alis_bridge *P1, alis_bridge *P2, float alt)
{
	if (S->dirty) alis_bridge_commit(S);
	if (P1->dirty) alis_bridge_commit(P1);
	if (P2->dirty) alis_bridge_commit(P2);
	if (alis_shards_count(S->S) > 1 || alis_shards_count(P1->S) > 1 || alis_shards_count(P2->S) > 1) { printf("alis-b200: nonLinearSHG/TWM are not available on a world cut into z slabs (ALIS_DEVICES)\n"); exit(1); }
	coeffs *C = S->W->CE;
	for (int n=0; n<C->N; n++) for (int m=0; m<=S->W->complexField; m++)
		CHECK(alisb_chi2_couple(S->E, n, m, P1->E, P2->E, &C->M[n].d[0][0], alt));
}

void alis_bridge_invalidate(alis_bridge *B) { alis_bridge_mark(B, DIRTY_ALL); }

/* dense host block (reference layout, incl. ghost layer) of one field component */
void alis_bridge_read_field(alis_bridge *B, int set, int comp, int m, float *host)
{
	world W = B->W;
	int32_t ids[6] = {set, comp, m, 0, 0, 0};
	int32_t box[6] = {0, W->iNUM+1, 0, W->jNUM+1, 0, W->kNUM+1};
	CHECK(alis_shards_read_box(B->S, ALISB_ARR_FIELD, ids, box, host));
}

void alis_bridge_write_field(alis_bridge *B, int set, int comp, int m, const float *host)
{
	world W = B->W;
	int32_t ids[6] = {set, comp, m, 0, 0, 0};
	int32_t box[6] = {0, W->iNUM+1, 0, W->jNUM+1, 0, W->kNUM+1};
	CHECK(alis_shards_write_box(B->S, ALISB_ARR_FIELD, ids, box, host));
}
