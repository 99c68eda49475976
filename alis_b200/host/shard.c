/*
 * shard.c -- one world on several engines: z-slab decomposition behind the public C API
 * (SURVEY.md 8e).  See shard.h.
 *
 * What the reference's time step needs across a cut between planes c and c+1 is one plane of two
 * components per half step (update.c:110-111: the E pass at plane c reads H.x, H.y at c+1; :115-116: the
 * H pass at c+1 reads E.x, E.y at c).  Everything else a pass does is local to a cell or confined to
 * an x-y plane, so the descriptors of the engine ABI are simply cut along z:
 *   1-D tables, spacings      the slab's window of the z tables
 *   CPML slabs                z slabs stay with the slabs that hold their planes
 *   material blocks, ADE      the planes of the slab, repacked (dense host blocks are z fastest)
 *   sources, TFSF strips      boxes clipped to the slab; the auxiliary world is replicated per slab
 *   boundary copies           x / y faces clipped to the slab, z faces on the end slabs only
 * Worlds that are periodic or Bloch-periodic along z are not cut (their z wrap would be a second
 * exchange; none of the reference's examples has one): they run on the first device with a notice.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shard.h"

#define MAXSH 16
#define MAXOBJ 256

struct alis_shards {
	int n, mode;
	alisb_grid g;                      /* the whole world                                        */
	alisb_engine *E[MAXSH];
	int c0[MAXSH], c1[MAXSH];          /* global planes owned by slab s (array coordinates)        */
	int nobj;                          /* dispersive objects of the world ...                       */
	int objmap[MAXSH][MAXOBJ];         /* ... and their index inside slab s, or -1                  */
	int32_t objbox[MAXOBJ][6];
};

#define TRYS(call) do { int rc_ = (call); if (rc_ != ALISB_OK) return rc_; } while (0)

int alis_shards_env_devices(int *devices, int cap)
{
	const char *s = getenv("ALIS_DEVICES");
	int n = 0;
	if (s && *s) {
		while (*s && n < cap) {
			char *end;
			long v = strtol(s, &end, 10);
			if (end == s) break;
			devices[n++] = (int)v;
			s = end;
			while (*s == ',' || *s == ' ') s++;
		}
	}
	if (n == 0) { const char *d = getenv("ALIS_DEVICE"); devices[n++] = d ? atoi(d) : 0; }
	return n;
}

int  alis_shards_count(const alis_shards *S) { return S->n; }
int  alis_shards_mode(const alis_shards *S) { return S->mode; }
alisb_engine *alis_shards_engine(alis_shards *S, int s) { return S->E[s]; }
void alis_shards_range(const alis_shards *S, int s, int *c0, int *c1) { *c0 = S->c0[s]; *c1 = S->c1[s]; }

static int split(const alis_shards *S) { return S->mode == ALIS_SHARDS_SPLIT && S->n > 1; }
/* planes slab s answers for when a box is read, written or clipped: its own, plus the ghost plane at either end of the world */
static int lo_of(const alis_shards *S, int s) { return s == 0 ? INT_MIN/2 : S->c0[s]; }
static int hi_of(const alis_shards *S, int s) { return s == S->n - 1 ? INT_MAX/2 : S->c1[s]; }
static int off_of(const alis_shards *S, int s) { return S->c0[s] - 1; }    /* local c = global c - off */

int alis_shards_create(const alisb_grid *g, const int *devices, int n, int mode, alis_shards **out)
{
	alis_shards *S = calloc(1, sizeof *S);
	if (!S) return ALISB_ENOMEM;
	S->g = *g; S->mode = mode;
	if (n < 1) n = 1;
	if (n > MAXSH) n = MAXSH;
	if (mode == ALIS_SHARDS_SPLIT) while (n > 1 && g->nz / n < 4) n--;      /* at least four planes per slab */
	S->n = n;
	for (int s=0; s<n; s++) {
		alisb_grid q = *g;
		q.device = devices[s];
		S->c0[s] = 1; S->c1[s] = g->nz;
		if (split(S)) {
			const int base = g->nz / n, rem = g->nz % n;
			S->c0[s] = 1 + s*base + (s < rem ? s : rem);
			S->c1[s] = S->c0[s] + base - 1 + (s < rem ? 1 : 0);
			const int off = off_of(S, s);
			q.nz = S->c1[s] - S->c0[s] + 1;
			q.kshard_lo = S->c0[s]; q.nz_global = g->nz;
			for (int set=0; set<2; set++) for (int sd=0; sd<2; sd++) {
				int f = g->pml[set][2][sd][0], l = g->pml[set][2][sd][1];
				if (f <= l) { if (f < S->c0[s]) f = S->c0[s]; if (l > S->c1[s]) l = S->c1[s]; }
				if (f > l) { q.pml[set][2][sd][0] = 1; q.pml[set][2][sd][1] = 0; }
				else       { q.pml[set][2][sd][0] = f - off; q.pml[set][2][sd][1] = l - off; }
			}
		}
		int rc = alisb_create(&q, &S->E[s]);
		if (rc != ALISB_OK) { for (int t=0; t<s; t++) alisb_destroy(S->E[t]); free(S); return rc; }
	}
	*out = S;
	return ALISB_OK;
}

void alis_shards_destroy(alis_shards *S)
{
	if (!S) return;
	for (int s=0; s<S->n; s++) alisb_synchronize(S->E[s]);      /* no plane may still be on its way into a slab that is being freed */
	for (int s=0; s<S->n; s++) alisb_destroy(S->E[s]);
	free(S);
}

int alis_shards_set_axis_tables(alis_shards *S, int set, int axis, const float *dtd, const float *P, const float *P1, const float *P2)
{
	for (int s=0; s<S->n; s++) {
		const int o = (split(S) && axis == 2) ? off_of(S, s) : 0;        /* the slab's window: entries off .. off + nz_local + 1 */
		TRYS(alisb_set_axis_tables(S->E[s], set, axis, dtd ? dtd + o : 0, P ? P + o : 0, P1 ? P1 + o : 0, P2 ? P2 + o : 0));
	}
	return ALISB_OK;
}

int alis_shards_set_spacing(alis_shards *S, int set, int axis, const float *ds)
{
	for (int s=0; s<S->n; s++) TRYS(alisb_set_spacing(S->E[s], set, axis, ds ? ds + ((split(S) && axis == 2) ? off_of(S, s) : 0) : 0));
	return ALISB_OK;
}

/* planes [k0, k0+nk) of a dense z-fastest block of na x nb x nc floats */
static float *cut_planes(const float *src, size_t na, size_t nb, size_t nc, size_t k0, size_t nk)
{
	float *dst = malloc(sizeof(float)*(na*nb*nk ? na*nb*nk : 1));
	if (!dst) return 0;
	for (size_t r=0; r<na*nb; r++) memcpy(dst + r*nk, src + r*nc + k0, sizeof(float)*nk);
	return dst;
}

int alis_shards_set_inv_eps(alis_shards *S, const float *cx, const float *cy, const float *cz)
{
	if (!split(S) || (!cx && !cy && !cz)) {
		for (int s=0; s<S->n; s++) TRYS(alisb_set_inv_eps(S->E[s], cx, cy, cz));
		return ALISB_OK;
	}
	if (!cx || !cy || !cz) return ALISB_EINVAL;
	const size_t nx = S->g.nx, ny = S->g.ny, nz = S->g.nz;
	for (int s=0; s<S->n; s++) {
		/* extents of the three blocks: alis_b200.h (cx one more along x, cy along y, cz along z) */
		const size_t k0 = off_of(S, s), nzl = S->c1[s] - S->c0[s] + 1;
		float *x = cut_planes(cx, nx + 1, ny, nz, k0, nzl), *y = cut_planes(cy, nx, ny + 1, nz, k0, nzl), *z = cut_planes(cz, nx, ny, nz + 1, k0, nzl + 1);
		int rc = (x && y && z) ? alisb_set_inv_eps(S->E[s], x, y, z) : ALISB_ENOMEM;
		free(x); free(y); free(z);
		TRYS(rc);
	}
	return ALISB_OK;
}

int alis_shards_clear_dispersive(alis_shards *S)
{
	S->nobj = 0;
	for (int s=0; s<S->n; s++) TRYS(alisb_clear_dispersive(S->E[s]));
	return ALISB_OK;
}

/* z range of `box` that slab s answers for, in LOCAL coordinates; 0 when empty */
static int clip_z(const alis_shards *S, int s, const int32_t box[6], int32_t out[6])
{
	memcpy(out, box, 6*sizeof(int32_t));
	if (!split(S)) return box[4] <= box[5];
	const int lo = lo_of(S, s), hi = hi_of(S, s), off = off_of(S, s);
	const int k0 = box[4] > lo ? box[4] : lo, k1 = box[5] < hi ? box[5] : hi;
	if (k0 > k1) return 0;
	out[4] = k0 - off; out[5] = k1 - off;
	return 1;
}

int alis_shards_add_dispersive(alis_shards *S, const alisb_dispersive *d)
{
	if (S->nobj >= MAXOBJ) return ALISB_EINVAL;
	const int g = S->nobj++;
	memcpy(S->objbox[g], d->box, sizeof S->objbox[g]);
	const size_t na = d->box[1] - d->box[0] + 1, nb = d->box[3] - d->box[2] + 1, nc = d->box[5] - d->box[4] + 1;
	for (int s=0; s<S->n; s++) {
		S->objmap[s][g] = -1;
		alisb_dispersive q = *d;
		if (!clip_z(S, s, d->box, q.box)) continue;
		int count = 0;
		for (int t=0; t<g; t++) if (S->objmap[s][t] >= 0) count++;
		S->objmap[s][g] = count;
		if (!split(S)) { TRYS(alisb_add_dispersive(S->E[s], &q)); continue; }
		const size_t k0 = q.box[4] + off_of(S, s) - d->box[4], nk = q.box[5] - q.box[4] + 1;
		float *f[3];
		for (int c=0; c<3; c++) f[c] = cut_planes(d->fill[c], na, nb, nc, k0, nk);
		for (int c=0; c<3; c++) q.fill[c] = f[c];
		int rc = (f[0] && f[1] && f[2]) ? alisb_add_dispersive(S->E[s], &q) : ALISB_ENOMEM;
		for (int c=0; c<3; c++) free(f[c]);
		TRYS(rc);
	}
	return ALISB_OK;
}

int alis_shards_clear_sources(alis_shards *S)
{
	for (int s=0; s<S->n; s++) TRYS(alisb_clear_sources(S->E[s]));
	return ALISB_OK;
}

int alis_shards_add_source(alis_shards *S, const alisb_source *d)
{
	const size_t na = d->box[1] - d->box[0] + 1, nb = d->box[3] - d->box[2] + 1, nc = d->box[5] - d->box[4] + 1;
	for (int s=0; s<S->n; s++) {
		alisb_source q = *d;
		if (!clip_z(S, s, d->box, q.box)) continue;
		if (!split(S)) { TRYS(alisb_add_source(S->E[s], &q)); continue; }
		const size_t k0 = q.box[4] + off_of(S, s) - d->box[4], nk = q.box[5] - q.box[4] + 1;
		float *A = d->Amplitude ? cut_planes(d->Amplitude, na, nb, nc, k0, nk) : 0, *P = d->Phase ? cut_planes(d->Phase, na, nb, nc, k0, nk) : 0;
		q.Amplitude = A; q.Phase = P;
		int rc = ((d->Amplitude && !A) || (d->Phase && !P)) ? ALISB_ENOMEM : alisb_add_source(S->E[s], &q);
		free(A); free(P);
		TRYS(rc);
	}
	return ALISB_OK;
}

int alis_shards_add_tfsf(alis_shards *S, const alisb_tfsf *t, alis_shards *aux)
{
	if (!aux || aux->n != S->n) return ALISB_EINVAL;
	if (split(S) && aux->mode != ALIS_SHARDS_REPLICA) return ALISB_EINVAL;     /* every slab steps its own copy of the auxiliary world */
	alisb_strip *LE = malloc(sizeof(alisb_strip)*(t->n_strips_E + 1)), *LH = malloc(sizeof(alisb_strip)*(t->n_strips_H + 1));
	if (!LE || !LH) { free(LE); free(LH); return ALISB_ENOMEM; }
	int rc = ALISB_OK;
	for (int s=0; s<S->n && rc == ALISB_OK; s++) {
		alisb_tfsf q = *t;
		q.aux = aux->E[s];
		int nE = 0, nH = 0;
		for (int k=0; k<t->n_strips_E; k++) { LE[nE] = t->strips_E[k]; if (clip_z(S, s, t->strips_E[k].box, LE[nE].box)) nE++; }
		for (int k=0; k<t->n_strips_H; k++) { LH[nH] = t->strips_H[k]; if (clip_z(S, s, t->strips_H[k].box, LH[nH].box)) nH++; }
		q.strips_E = LE; q.n_strips_E = nE; q.strips_H = LH; q.n_strips_H = nH;
		if (split(S)) q.offset[2] = t->offset[2] + off_of(S, s);               /* auxiliary coordinate = main LOCAL coordinate + offset */
		rc = alisb_add_tfsf(S->E[s], &q);
	}
	free(LE); free(LH);
	return rc;
}

int alis_shards_set_copies(alis_shards *S, int set, const alisb_copy *ops, int n)
{
	if (!split(S)) {
		for (int s=0; s<S->n; s++) TRYS(alisb_set_copies(S->E[s], set, ops, n));
		return ALISB_OK;
	}
	alisb_copy *L = malloc(sizeof(alisb_copy)*(n + 1));
	if (!L) return ALISB_ENOMEM;
	int rc = ALISB_OK;
	for (int s=0; s<S->n && rc == ALISB_OK; s++) {
		int m = 0;
		for (int k=0; k<n; k++) {
			L[m] = ops[k];
			if (ops[k].alt[2] != 0) {
				/* a z-face copy: target and source plane lie at the same end of the world (alis_shards_create refuses z-periodic worlds) */
				const int at_top = ops[k].box[4] > S->g.nz/2;
				if ((at_top && s != S->n - 1) || (!at_top && s != 0)) continue;
				L[m].box[4] = ops[k].box[4] - off_of(S, s); L[m].box[5] = ops[k].box[5] - off_of(S, s);
				m++;
			} else if (clip_z(S, s, ops[k].box, L[m].box)) m++;
		}
		rc = alisb_set_copies(S->E[s], set, L, m);
	}
	free(L);
	return rc;
}

int alis_shards_set_time(alis_shards *S, uint64_t n, float t, float tE, float tH)
{
	for (int s=0; s<S->n; s++) TRYS(alisb_set_time(S->E[s], n, t, tE, tH));
	return ALISB_OK;
}

int alis_shards_get_time(alis_shards *S, uint64_t *n, float *t, float *tE, float *tH) { return alisb_get_time(S->E[0], n, t, tE, tH); }

static int exchange(alis_shards *S, int set)
{
	if (split(S)) for (int s=0; s+1<S->n; s++) TRYS(alisb_halo_exchange(S->E[s], S->E[s+1], set));
	return ALISB_OK;
}

/* Every slab's pass is enqueued on its own device (the calls return at once: the slabs run side by side).  A slab with
 * a neighbour above computes the plane that neighbour needs first (alisb_step_part SEND_PLANE), the plane starts its
 * journey (alisb_halo_exchange: a peer copy on the owner's communication stream; the receiver waits for it only when the
 * pass that reads the ghost plane begins), then the rest of the pass follows -- E.x, E.y of a slab's top plane go up,
 * H.x, H.y of its bottom plane go down (update.c:110-111,115-116).  With fused E+H stepping the bulk of the E pass is
 * issued together with the H pass (alisb_flush(e, 1) in alis_shards_step_H): the buffers swap first (alisb_begin_pass),
 * so that the planes travel between the buffers the passes are about to write. */
int alis_shards_step_E(alis_shards *S)
{
	if (!split(S)) {
		for (int s=0; s<S->n; s++) TRYS(alisb_step_E(S->E[s]));
		return ALISB_OK;
	}
	for (int s=0; s<S->n; s++) TRYS(alisb_begin_pass(S->E[s], ALISB_SET_E));
	for (int s=0; s<S->n; s++) {
		if (s + 1 < S->n) {
			TRYS(alisb_step_part(S->E[s], ALISB_SET_E, ALISB_PART_SEND_PLANE));
			TRYS(alisb_halo_exchange(S->E[s], S->E[s+1], ALISB_SET_E));
			TRYS(alisb_step_part(S->E[s], ALISB_SET_E, ALISB_PART_REST));
		} else TRYS(alisb_step_E(S->E[s]));
	}
	return ALISB_OK;
}

int alis_shards_step_H(alis_shards *S)
{
	if (!split(S)) {
		for (int s=0; s<S->n; s++) TRYS(alisb_step_H(S->E[s]));
		return ALISB_OK;
	}
	for (int s=0; s<S->n; s++) { TRYS(alisb_flush(S->E[s], 1)); TRYS(alisb_begin_pass(S->E[s], ALISB_SET_H)); }
	/* from the top slab down: a slab's bottom plane is on its way before the slab below enqueues its own pass (the copy
	 * only has to wait for what that slab had enqueued before: its previous E pass, the last reader of the ghost plane) */
	for (int s=S->n-1; s>=0; s--) {
		if (s > 0) {
			TRYS(alisb_step_part(S->E[s], ALISB_SET_H, ALISB_PART_SEND_PLANE));
			TRYS(alisb_halo_exchange(S->E[s-1], S->E[s], ALISB_SET_H));
			TRYS(alisb_step_part(S->E[s], ALISB_SET_H, ALISB_PART_REST));
		} else TRYS(alisb_step_H(S->E[s]));
	}
	return ALISB_OK;
}

int alis_shards_reset_fields(alis_shards *S)
{
	for (int s=0; s<S->n; s++) TRYS(alisb_reset_fields(S->E[s]));
	return ALISB_OK;
}

static int rw_box(alis_shards *S, int arr, const int32_t ids[6], const int32_t box[6], float *host, int write)
{
	if (!split(S)) {
		if (!write) return alisb_read_box(S->E[0], arr, ids, box, host);
		for (int s=0; s<S->n; s++) TRYS(alisb_write_box(S->E[s], arr, ids, box, host));   /* replicas stay identical */
		return ALISB_OK;
	}
	const size_t rows = (size_t)(box[1] - box[0] + 1)*(box[3] - box[2] + 1), nc = box[5] - box[4] + 1;
	const int per_object = arr == ALISB_ARR_ADE_J || arr == ALISB_ARR_ADE_K || arr == ALISB_ARR_ADE_FILL;
	const int oslot = arr == ALISB_ARR_ADE_FILL ? 0 : 1;                       /* position of the object index in ids */
	for (int s=0; s<S->n; s++) {
		int32_t lb[6], lid[6];
		memcpy(lid, ids, sizeof lid);
		if (!clip_z(S, s, box, lb)) continue;
		if (per_object) {
			const int g = ids[oslot];
			if (g < 0 || g >= S->nobj) return ALISB_EINVAL;
			if (S->objmap[s][g] < 0) { if (!write) { /* the slab holds no part of the object: its planes read as zero */
					const size_t k0 = lb[4] + off_of(S, s) - box[4], nk = lb[5] - lb[4] + 1;
					for (size_t r=0; r<rows; r++) memset(host + r*nc + k0, 0, sizeof(float)*nk); }
				continue; }
			lid[oslot] = S->objmap[s][g];
			/* only the planes of the object that the slab holds */
			int32_t ob[6];
			if (!clip_z(S, s, S->objbox[g], ob)) continue;
			if (lb[4] < ob[4]) lb[4] = ob[4];
			if (lb[5] > ob[5]) lb[5] = ob[5];
			if (lb[4] > lb[5]) continue;
		}
		const size_t k0 = lb[4] + off_of(S, s) - box[4], nk = lb[5] - lb[4] + 1;
		if (nk == nc) { TRYS(write ? alisb_write_box(S->E[s], arr, lid, lb, host) : alisb_read_box(S->E[s], arr, lid, lb, host)); continue; }
		float *tmp = malloc(sizeof(float)*rows*nk);
		if (!tmp) return ALISB_ENOMEM;
		int rc = ALISB_OK;
		if (write) {
			for (size_t r=0; r<rows; r++) memcpy(tmp + r*nk, host + r*nc + k0, sizeof(float)*nk);
			rc = alisb_write_box(S->E[s], arr, lid, lb, tmp);
		} else {
			rc = alisb_read_box(S->E[s], arr, lid, lb, tmp);
			if (rc == ALISB_OK) for (size_t r=0; r<rows; r++) memcpy(host + r*nc + k0, tmp + r*nk, sizeof(float)*nk);
		}
		free(tmp);
		TRYS(rc);
	}
	if (write && arr == ALISB_ARR_FIELD) {
		/* a plane next to a cut may have changed: the neighbour's ghost copy follows */
		TRYS(exchange(S, ALISB_SET_E));
		TRYS(exchange(S, ALISB_SET_H));
	}
	return ALISB_OK;
}

int alis_shards_read_box(alis_shards *S, int arr, const int32_t ids[6], const int32_t box[6], float *host) { return rw_box(S, arr, ids, box, host, 0); }
int alis_shards_write_box(alis_shards *S, int arr, const int32_t ids[6], const int32_t box[6], const float *host) { return rw_box(S, arr, ids, box, (float*)host, 1); }

/* ---- running DFT on a cut world (reference phaser.c:83-110; far-field sums farfld.c:248-459) -------------------------
 * Every slab accumulates the part of the monitor box's faces that lies in its planes: the x and y faces are cut along z,
 * a z face belongs to the slab that owns its plane -- the H average over the planes k and k+1 (phaser.c:89-90) then reads
 * either an own plane or the ghost plane the H exchange fills.  Face arrays are stitched when they are read; a far-field
 * surface sum is the sum of the slabs' partial sums over their z ranges (weights along z start at the slab's offset). */
struct alis_phasers {
	alis_shards *S;
	int n;
	alisb_phaser *p[MAXSH];
	int32_t box[6];
	int mask;
	int k0[MAXSH], k1[MAXSH];          /* global planes of the monitor box held by slab s (k0 > k1: none) */
	int smask[MAXSH];
};

int alis_phasers_create(alis_shards *S, const int32_t box[6], int face_mask, int nfreq, alis_phasers **out)
{
	alis_phasers *P = calloc(1, sizeof *P);
	if (!P) return ALISB_ENOMEM;
	P->S = S; P->mask = face_mask; memcpy(P->box, box, sizeof P->box);
	P->n = split(S) ? S->n : 1;                   /* replicas are identical: the first one records */
	for (int s=0; s<P->n; s++) {
		int32_t b[6];
		memcpy(b, box, sizeof b);
		P->k0[s] = box[4]; P->k1[s] = box[5]; P->smask[s] = face_mask;
		if (split(S)) {
			P->k0[s] = box[4] > S->c0[s] ? box[4] : S->c0[s];
			P->k1[s] = box[5] < S->c1[s] ? box[5] : S->c1[s];
			if (P->k0[s] > P->k1[s]) { P->smask[s] = 0; continue; }
			b[4] = P->k0[s] - off_of(S, s); b[5] = P->k1[s] - off_of(S, s);
			if (P->k0[s] != box[4]) P->smask[s] &= ~(1 << 4);
			if (P->k1[s] != box[5]) P->smask[s] &= ~(1 << 5);
		}
		if (!P->smask[s]) continue;
		int rc = alisb_phaser_create(S->E[s], b, P->smask[s], nfreq, &P->p[s]);
		if (rc != ALISB_OK) { for (int t=0; t<s; t++) if (P->p[t]) alisb_phaser_destroy(P->p[t]); free(P); return rc; }
	}
	*out = P;
	return ALISB_OK;
}

void alis_phasers_destroy(alis_phasers *P)
{
	if (!P) return;
	for (int s=0; s<P->n; s++) if (P->p[s]) alisb_phaser_destroy(P->p[s]);
	free(P);
}

int alis_phasers_accumulate(alis_phasers *P, const float *expE, const float *expH)
{
	for (int s=0; s<P->n; s++) if (P->p[s]) TRYS(alisb_phaser_accumulate(P->p[s], expE, expH));
	return ALISB_OK;
}

int alis_phasers_read(alis_phasers *P, int freq, int face, int which, float *host_complex)
{
	if (P->n == 1) return P->p[0] ? alisb_phaser_read(P->p[0], freq, face, which, host_complex) : ALISB_EINVAL;
	const int ax = face/2;
	if (ax == 2) {                                /* the slab that owns the plane */
		for (int s=0; s<P->n; s++) if (P->p[s] && (P->smask[s] >> face & 1)) return alisb_phaser_read(P->p[s], freq, face, which, host_complex);
		return ALISB_EINVAL;
	}
	/* dense over the face, z fastest: rows of nc complex values, one per cell along the other in-face axis */
	const size_t rows = ax == 0 ? (size_t)(P->box[3] - P->box[2] + 1) : (size_t)(P->box[1] - P->box[0] + 1), nc = P->box[5] - P->box[4] + 1;
	for (int s=0; s<P->n; s++) {
		if (!P->p[s] || !(P->smask[s] >> face & 1)) continue;
		const size_t nk = P->k1[s] - P->k0[s] + 1, k0 = P->k0[s] - P->box[4];
		float *tmp = malloc(sizeof(float)*2*rows*nk);
		if (!tmp) return ALISB_ENOMEM;
		int rc = alisb_phaser_read(P->p[s], freq, face, which, tmp);
		if (rc == ALISB_OK) for (size_t r=0; r<rows; r++) memcpy(host_complex + 2*(r*nc + k0), tmp + 2*r*nk, sizeof(float)*2*nk);
		free(tmp);
		TRYS(rc);
	}
	return ALISB_OK;
}

int alis_phasers_surface_sums(alis_phasers *P, int freq, const alisb_surface_job *jobs, int njobs, const float *weights, size_t nweights, double *out)
{
	if (P->n == 1) return P->p[0] ? alisb_phaser_surface_sums(P->p[0], freq, jobs, njobs, weights, nweights, out) : ALISB_EINVAL;
	alisb_surface_job *mine = malloc(sizeof *mine*(njobs + 1));
	int *where = malloc(sizeof(int)*(njobs + 1));
	double *part = malloc(sizeof(double)*2*(njobs + 1));
	if (!mine || !where || !part) { free(mine); free(where); free(part); return ALISB_ENOMEM; }
	for (int j=0; j<2*njobs; j++) out[j] = 0;
	int rc = ALISB_OK;
	for (int s=0; s<P->n && rc == ALISB_OK; s++) {
		if (!P->p[s]) continue;
		int m = 0;
		for (int j=0; j<njobs; j++) {
			if (!(P->smask[s] >> jobs[j].face & 1)) continue;
			mine[m] = jobs[j];
			if (jobs[j].face < 4) mine[m].w_hi += (uint32_t)(P->k0[s] - P->box[4]);      /* v runs along z on the x and y faces */
			where[m++] = j;
		}
		if (!m) continue;
		rc = alisb_phaser_surface_sums(P->p[s], freq, mine, m, weights, nweights, part);
		if (rc == ALISB_OK) for (int q=0; q<m; q++) { out[2*where[q]] += part[2*q]; out[2*where[q]+1] += part[2*q+1]; }
	}
	free(mine); free(where); free(part);
	return rc;
}
