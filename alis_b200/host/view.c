/*
 * view.c -- host views of device-resident fields.
 *
 * The reference's monitors read W->E[m].x[i][j][k] directly (reference fields.c:130-291,
 * output.c:23-47, phaser.c:87-90).  Here those arrays are HOST VIEWS: allocated on first use with
 * the reference's shape ([iMIN-1..iMAX+1]..., world.c:170-172) and refreshed from the device only
 * over the region a monitor is about to read (alisb_read_box), so a point probe costs a few
 * bytes of D2H and a Poynting box six thin slabs -- never a full-field round trip unless the
 * caller asks for it (alisSyncFields).
 */
#include <string.h>
#include "alis_c.h"
#include "host_internal.h"

typedef struct { unsigned long synced; int box[6], have; } arr_state;

typedef struct {
	unsigned long version;        /* bumped whenever the device state advances              */
	arr_state arr[2][2][3];       /* [m][set][comp]: which box of each host view is current  */
	arr_state *ade; int nade;     /* per dispersive object: which box of its current arrays (all poles, components, field sets) is current */
} view_state;

static view_state *state(world W)
{
	if (!W->b200_view) W->b200_view = calloc(1, sizeof(view_state));
	return (view_state*)W->b200_view;
}

unsigned long alis_view_version(world W) { return state(W)->version; }

void alis_host_touch(world W)
{
	state(W)->version++;
	for (tfsf S = W->S; S; S = S->S) alis_host_touch(S->W);
}

static float ***view_array(world W, int set, int comp, int m)
{
	vfield *V = set == 0 ? &W->E[m] : &W->H[m];
	float ****slot = comp == 0 ? &V->x : comp == 1 ? &V->y : &V->z;
	if (!*slot) *slot = makeField(W->iMIN-1, W->iMAX+1, W->jMIN-1, W->jMAX+1, W->kMIN-1, W->kMAX+1);
	return *slot;
}

static void pull_box(world W, int set, int comp, int m, const int b[6])
{
	alis_bridge *E = alis_host_bridge(W);      /* all z slabs of the world (one engine unless ALIS_DEVICES names several) */
	const int nj = b[3]-b[2]+1, nk = b[5]-b[4]+1;
	int32_t box[6] = { b[0]-(W->iMIN-1), b[1]-(W->iMIN-1), b[2]-(W->jMIN-1), b[3]-(W->jMIN-1), b[4]-(W->kMIN-1), b[5]-(W->kMIN-1) };
	int32_t ids[6] = { set, comp, m, 0, 0, 0 };
	float ***A = view_array(W, set, comp, m);
	const int whole_rows = b[4] == W->kMIN-1 && b[5] == W->kMAX+1, whole_sheets = whole_rows && b[2] == W->jMIN-1 && b[3] == W->jMAX+1;
	if (whole_sheets) {                        /* contiguous in the host view: read straight into it */
		if (alis_bridge_read_box(E, ALISB_ARR_FIELD, ids, box, &A[b[0]][b[2]][b[4]]) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
		return;
	}
	float *tmp = malloc(sizeof(float)*(size_t)(b[1]-b[0]+1)*nj*nk);
	if (alis_bridge_read_box(E, ALISB_ARR_FIELD, ids, box, tmp) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
	size_t q = 0;
	for (int i=b[0]; i<=b[1]; i++) for (int j=b[2]; j<=b[3]; j++) { memcpy(&A[i][j][b[4]], tmp+q, sizeof(float)*nk); q += nk; }
	free(tmp);
}

/* make host view (set, comp, m) current on [i0-1..i1+1] x [j0-1..j1+1] x [k0-1..k1+1] (clipped) */
void alis_view_need_arr(world W, int set, int comp, int m, int i0, int i1, int j0, int j1, int k0, int k1)
{
	view_state *v = state(W);
	arr_state *a = &v->arr[m][set][comp];
	int b[6] = { MAX(i0-1, W->iMIN-1), MIN(i1+1, W->iMAX+1), MAX(j0-1, W->jMIN-1), MIN(j1+1, W->jMAX+1), MAX(k0-1, W->kMIN-1), MIN(k1+1, W->kMAX+1) };
	if (b[0] > b[1] || b[2] > b[3] || b[4] > b[5]) { view_array(W, set, comp, m); return; }
	if (a->have && a->synced == v->version &&
	    a->box[0] <= b[0] && b[1] <= a->box[1] && a->box[2] <= b[2] && b[3] <= a->box[3] && a->box[4] <= b[4] && b[5] <= a->box[5]) return;
	/* a fetch from inside a team of evaluating threads (alis_eval_region) is rare -- the region was fetched before the team
	 * started -- but must not run twice at once: one thread at a time, and whoever comes second looks again first */
	#pragma omp critical(alis_view_fetch)
	{
	const int covered = a->have && a->synced == v->version &&
	    a->box[0] <= b[0] && b[1] <= a->box[1] && a->box[2] <= b[2] && b[3] <= a->box[3] && a->box[4] <= b[4] && b[5] <= a->box[5];
	if (!covered) {
	/* grow the working set: monitors called every step (probes, slices) ask for the same regions
	 * again, so fetch the union with the previous request unless that is wasteful */
	if (a->have) {
		int u[6];
		for (int q=0; q<3; q++) u[2*q] = MIN(b[2*q], a->box[2*q]), u[2*q+1] = MAX(b[2*q+1], a->box[2*q+1]);
		const double vu = (double)(u[1]-u[0]+1)*(u[3]-u[2]+1)*(u[5]-u[4]+1);
		const double vb = (double)(b[1]-b[0]+1)*(b[3]-b[2]+1)*(b[5]-b[4]+1);
		const double vo = (double)(a->box[1]-a->box[0]+1)*(a->box[3]-a->box[2]+1)*(a->box[5]-a->box[4]+1);
		if (vu <= 4*(vb+vo) + 4096) memcpy(b, u, sizeof u);
	}
	pull_box(W, set, comp, m, b);
	memcpy(a->box, b, sizeof b);
	a->have = 1; a->synced = v->version;
	}
	}
}

/* all components of both field sets (monitors that mix E and H) */
void alis_view_need(world W, int i0, int i1, int j0, int j1, int k0, int k1)
{
	for (int m=0; m<=W->complexField; m++) for (int set=0; set<2; set++) for (int comp=0; comp<3; comp++)
		alis_view_need_arr(W, set, comp, m, i0, i1, j0, j1, k0, k1);
}

void alis_view_need_for(world W, field F, int i0, int i1, int j0, int j1, int k0, int k1)
{
	const unsigned mask = alis_field_arrays(W, F);
	for (int m=0; m<=W->complexField; m++) for (int set=0; set<2; set++) for (int comp=0; comp<3; comp++)
		if (mask & (1u << ((m*2+set)*3+comp))) alis_view_need_arr(W, set, comp, m, i0, i1, j0, j1, k0, k1);
}

void alis_view_need_all(world W)
{
	alis_view_need(W, W->iMIN-1, W->iMAX+1, W->jMIN-1, W->jMAX+1, W->kMIN-1, W->kMAX+1);
}

/* host views of the ADE polarisation currents W->E[m].Jx[n][p][i][j][k] (read by rawJ*, jx..jz) */
/* The polarisation currents of the dispersive objects over the index box (clipped to every object's own box): what a
 * function that reads currents at a few cells needs (a `get(W, JE, ...)` every half step used to bring ALL current arrays of
 * the world back, megabytes per call).  One box per object covers all its poles, components and field sets. */
void alis_view_need_ade_box(world W, int i0, int i1, int j0, int j1, int k0, int k1)
{
	view_state *v = state(W);
	coeffs *C = W->CE;
	if (!C->N) return;
	if (v->nade < C->N) {
		#pragma omp critical(alis_view_fetch)
		if (v->nade < C->N) {
			arr_state *grown = calloc(C->N, sizeof *grown);
			if (v->ade) memcpy(grown, v->ade, sizeof(*grown)*v->nade);
			v->ade = grown;                    /* (the old table, if any, is left to the end of the process: another thread may be reading it) */
			v->nade = C->N;
		}
	}
	for (int n=0; n<C->N; n++) {
		if (!C->NK[n]) continue;
		arr_state *a = &v->ade[n];
		int b[6] = { MAX(i0, C->iMin[n]), MIN(i1, C->iMax[n]), MAX(j0, C->jMin[n]), MIN(j1, C->jMax[n]), MAX(k0, C->kMin[n]), MIN(k1, C->kMax[n]) };
		if (b[0] > b[1] || b[2] > b[3] || b[4] > b[5]) continue;
		if (a->have && a->synced == v->version &&
		    a->box[0] <= b[0] && b[1] <= a->box[1] && a->box[2] <= b[2] && b[3] <= a->box[3] && a->box[4] <= b[4] && b[5] <= a->box[5]) continue;
		#pragma omp critical(alis_view_fetch)
		{
		const int covered = a->have && a->synced == v->version &&
		    a->box[0] <= b[0] && b[1] <= a->box[1] && a->box[2] <= b[2] && b[3] <= a->box[3] && a->box[4] <= b[4] && b[5] <= a->box[5];
		if (!covered) {
			if (a->have) {                     /* grow the working set, as alis_view_need_arr does */
				int u[6];
				for (int q=0; q<3; q++) u[2*q] = MIN(b[2*q], a->box[2*q]), u[2*q+1] = MAX(b[2*q+1], a->box[2*q+1]);
				const double vu = (double)(u[1]-u[0]+1)*(u[3]-u[2]+1)*(u[5]-u[4]+1);
				const double vb = (double)(b[1]-b[0]+1)*(b[3]-b[2]+1)*(b[5]-b[4]+1);
				const double vo = (double)(a->box[1]-a->box[0]+1)*(a->box[3]-a->box[2]+1)*(a->box[5]-a->box[4]+1);
				if (vu <= 4*(vb+vo) + 4096) memcpy(b, u, sizeof u);
			}
			alis_bridge *E = alis_host_bridge(W);
			const int nj = b[3]-b[2]+1, nk = b[5]-b[4]+1;
			const int whole = b[2] == C->jMin[n] && b[3] == C->jMax[n] && b[4] == C->kMin[n] && b[5] == C->kMax[n];   /* whole sheets: contiguous in the host array */
			float *tmp = whole ? 0 : malloc(sizeof(float)*(size_t)(b[1]-b[0]+1)*nj*nk);
			for (int m=0; m<=W->complexField; m++) {
				float *****J[3] = { W->E[m].Jx, W->E[m].Jy, W->E[m].Jz };
				for (int p=0; p<C->NK[n]; p++) for (int c=0; c<3; c++) {
					int32_t ids[6] = { m, n, p, c, 0, 0 };
					int32_t box[6] = { b[0]-(W->iMIN-1), b[1]-(W->iMIN-1), b[2]-(W->jMIN-1), b[3]-(W->jMIN-1), b[4]-(W->kMIN-1), b[5]-(W->kMIN-1) };
					float ***A = J[c][n][p];
					if (alis_bridge_read_box(E, ALISB_ARR_ADE_J, ids, box, whole ? &A[b[0]][b[2]][b[4]] : tmp) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
					if (!whole) {
						size_t q = 0;
						for (int i=b[0]; i<=b[1]; i++) for (int j=b[2]; j<=b[3]; j++) { memcpy(&A[i][j][b[4]], tmp+q, sizeof(float)*nk); q += nk; }
					}
				}
			}
			free(tmp);
			memcpy(a->box, b, sizeof b);
			a->have = 1; a->synced = v->version;
		}
		}
	}
}

/* all of them (alisSyncFields, the host path of objectAbsorption) */
void alis_view_need_ade(world W)
{
	alis_view_need_ade_box(W, W->iMIN-1, W->iMAX+1, W->jMIN-1, W->jMAX+1, W->kMIN-1, W->kMAX+1);
}

void alis_view_free(world W)
{
	for (int m=0; m<=W->complexField; m++) {
		vfield *V[2] = { &W->E[m], &W->H[m] };
		for (int s=0; s<2; s++) {
			if (V[s]->x) removeField(V[s]->x, W->iMIN-1, W->jMIN-1, W->kMIN-1);
			if (V[s]->y) removeField(V[s]->y, W->iMIN-1, W->jMIN-1, W->kMIN-1);
			if (V[s]->z) removeField(V[s]->z, W->iMIN-1, W->jMIN-1, W->kMIN-1);
		}
	}
	if (W->b200_view) free(((view_state*)W->b200_view)->ade);
	free(W->b200_view);
	W->b200_view = 0;
}

/* ---- public extensions --------------------------------------------------------------------------- */
void alisSyncFields(world W) { alis_view_need_all(W); alis_view_need_ade(W); }

void alisReadField(world W, int set, int comp, int m, float *dense)
{
	alis_bridge_read_field(alis_host_bridge(W), set, comp, m, dense);
}

void alisWriteField(world W, int set, int comp, int m, const float *dense)
{
	alis_bridge_write_field(alis_host_bridge(W), set, comp, m, dense);
	alis_host_touch(W);
}

const char *alisBackend(void) { return alisb_backend(); }
