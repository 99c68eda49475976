/*
 * engine.cu -- the B200 FDTD engine behind the C ABI of include/alis_b200.h.
 *
 * Host orchestration of one updateE / updateH call (reference update.c:275-341):
 *   E: [ADE loss, ADE current subtraction] -> fused curl+CPML volume kernel (per field set m)
 *      -> TFSF (aux engine stepped first, then face strips) -> soft sources -> boundary copies
 *   H: fused curl+CPML volume kernel -> ADE current update from the final E -> TFSF -> sources
 *      -> boundary copies
 * All launches go to one stream; nothing synchronises with the host inside a step.  Waveform
 * scalars are evaluated on the host with the caller's libm (bit parity with the reference)
 * and passed as kernel arguments.
 *
 * There is no CPU fallback: without a usable CUDA device alisb_create fails.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <limits.h>
#include <vector>
#include <algorithm>
#include <chrono>
#include <nvtx3/nvToolsExt.h>      /* header-only NVTX 3: ranges cost a few ns when no tool is attached */
#include "alis_b200.h"
#include "kernels.cuh"
#include "copy_resolve.h"
#include "strip_resolve.h"

/* the fused kernel needs more than the default 48 KB of dynamic shared memory (a per-device function attribute) */
template <int ROWS, bool EPS, bool PMZ> static cudaError_t fused_optin_one()
{
	return cudaFuncSetAttribute(yee_fused<ROWS,EPS,PMZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem<ROWS,EPS>));
}

/* NVTX range per pass / monitor call: visible in nsys / ncu --nvtx timelines */
struct NvtxRange { NvtxRange(const char *name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
	return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ALISB_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct SubBox { int a0, a1, b0, b1, c0, c1; int pml; int fusable; };   /* fusable: the slab-free, fix-up-free box of the fused E+H kernel */   /* cell ranges; pml: bit 0/1/2 = may touch x/y/z slabs */
struct ProfRec { size_t ev; int set, sub; double cells; };
struct alisb_engine;
/* a whole pass (all kernels of updateE or updateH, auxiliary TFSF worlds included) captured once as
 * a CUDA graph and replayed; the only per-step parameters are the source lists, whose kernel nodes
 * are updated in place before each launch */
struct SrcNode { alisb_engine *eng; int set; cudaGraphNode_t node; };
struct PassGraph { cudaGraphExec_t exec; cudaGraph_t graph; std::vector<SrcNode> nodes; uint64_t version; int nkernels; };
enum { GMODE_NONE = 0, GMODE_CAPTURE = 1 };
static thread_local int g_mode = GMODE_NONE;
static thread_local std::vector<SrcNode> *g_nodes = 0;

struct DispObj {
	alisb_dispersive d;
	Box full, upd;
	size_t cnt;
	float *fill[3];
	float *J[2][ALISB_MAXPOLES][3], *K[2][ALISB_MAXPOLES][3];
};

#define SRC_RING 4
struct SrcObj {
	alisb_source d;
	Box B; size_t cnt;
	float *dAmp;                       /* device dense box (x fastest) or NULL            */
	std::vector<float> hAmp, hPha;     /* host copies in device order (x fastest)         */
	float *hvals[SRC_RING], *dvals[SRC_RING]; cudaEvent_t ev[SRC_RING]; int slot;
};

struct TfsfObj {
	alisb_tfsf d;
	std::vector<alisb_strip> sE, sH;
};

/* a resolved copy list (kernels.cuh, copy_resolved): built once per distinct list, found again by comparing the list */
struct CopyPlan { CopyList key; ResolvedCopy P; int4 *d_ent; float2 *d_tmp; };
/* a strip list regrouped per cell (kernels.cuh, strip_resolved), found again the same way */
struct StripPlan { StripList key; ResolvedStrips P; RsCellDev *d_cells; RsTermDev *d_terms; };

struct alisb_engine {
	alisb_grid g;
	int nx, ny, nz, pitch;
	size_t plane, elems;               /* floats per z-plane / per array (incl. slack)    */
	float *F[2][2][3];                 /* current fields [set][m][comp]                    */
	float *Fb[2][2][3];                /* double-buffered mode: the other buffer (input of a running pass) */
	bool dbuf;                         /* passes run out of place (Fb -> F) and swap at pass begin */
	bool pass_open[2];                 /* the running pass of this set has already swapped  */
	int fuse_req;                      /* fused E+H kernel requested (ALISB_FUSE / alisb_set_fusion) */
	double tune_fuse_frac;             /* smallest fused share of the grid worth fusing     */
	long tune_fuse_min;                /* smallest grid (cells) worth fusing                */
	bool fuse_ok; uint64_t fuse_cfg;   /* feasibility for the current configuration          */
	int R[6];                          /* the fused box a0,a1,b0,b1,c0,c1                    */
	std::vector<std::pair<int,int> > fuse_z;   /* its clean z ranges                        */
	int pendingE;                      /* -1, or the part of an E pass prepared but not yet launched */
	bool fusing, h_fused;              /* launching the fused variant / H' of the fused box is already in the H output buffer */
	float *eps[3];
	float *dtd[2][3], *P[2][3], *P1[2][3], *P2[2][3];
	float *psi[2][2][3][2][2]; size_t psi_elems[2][3][2];
	float *psib[2][2][2][2];           /* double-buffered mode: the other buffer of the z-slab psi arrays [set][m][side][which] (they swap with the fields) */
	int tune_pmz;                      /* z CPML slabs inside the fused box (ALISB_FUSE_PMZ, default 0) */
	int xbase[2][2], xpitch[2][2];     /* [set][side]                                     */
	std::vector<DispObj*> obj;
	AdeTable *d_ade[2]; uint64_t ade_cfg;   /* device copies of the in-kernel object table per complex set; built for ade_cfg */
	int tune_ade;                       /* dispersive objects handled inside the volume kernels (ALISB_ADE_FUSE, default 1) */
	std::vector<SrcObj*> src[2];
	std::vector<TfsfObj*> tf;
	std::vector<alisb_copy> copies[2];
	cudaStream_t stream; bool own_stream;
	uint64_t step; float t, tE, tH;
	uint64_t launches, fused_launches;
	int resident_ctas;                 /* volume kernel CTAs resident on the whole GPU    */
	int sm_count, multi_resident[2];   /* SMs; resident CTAs of the two merged PML-shell launch classes */
	std::vector<SubBox> plan[2];       /* region-specialised volume launches per field set */
	int tune_wave;                     /* fused stepping in place (kernel yee_wave, no second set of field arrays); ALISB_WAVE */
	unsigned long long *d_wave_ticket, *d_wave_progress; int *d_wave_error; size_t wave_tiles_cap;
	unsigned long long wave_ticket_base, wave_epoch;
	int tune_frows;                    /* rows per CTA of the fused kernel (8 or 16)          */
	int tune_fchunk;                   /* planes per z-chunk of the fused kernel              */
	int tune_merge;                    /* PML shell boxes gathered into two merged launches  */
	long tune_small;                   /* grids below this many cells use the single-launch plan */
	int tune_debug, tune_conc;         /* tune_conc: 0 serial sub-launches, 1 PML shell launches concurrent after the interior, 2 all concurrent */
	cudaStream_t side[3]; cudaEvent_t ev_fork, ev_join[3];
	cudaStream_t hi; cudaEvent_t ev_hi;    /* the fused launches: a stream of the highest priority, so that the block scheduler hands SMs to the
	                                        * PML shell launches (which run beside them) only when the fused kernel has no CTA left to issue */
	int tune_prio;                         /* ALISB_FUSE_PRIO (1) */
	double trace_s[2][5]; long trace_n[2][5];   /* ALISB_DEBUG: host seconds per field set in {prepare, graph capture, node patching, graph launch, direct launches} */
	/* alisb_halo_exchange (both z neighbours in one process): the peer copy runs on the owner's `comm` stream;
	 * who has to wait for it is told lazily -- the receiver before the pass that reads the ghost plane (or any
	 * read of the device state), the owner before it rewrites the plane */
	cudaStream_t comm;
	cudaEvent_t ev_plane, ev_war;      /* owner: the plane is complete / receiver: the old ghost plane is no longer read */
	cudaEvent_t ev_out[2];             /* [set] owner: the copy of my plane has finished (recorded on comm)                  */
	cudaEvent_t ev_in[2];              /* [set] receiver: the OWNER's ev_out of the copy into my ghost plane (a handle, not owned: events
	                                    * can only be recorded on a stream of their own device, but waited for from any) */
	bool wait_in[2], wait_out[2];
	int tune_bx, tune_by, tune_stream, tune_waves, tune_minchunk;   /* launch tunables (env ALISB_*)  */
	int f_lo, f_hi;                    /* plane filter of the current (partial) pass: ops act on c in [f_lo,f_hi] */
	bool profiling; std::vector<cudaEvent_t> ev_pool; size_t ev_used; float last_ms; int last_launches;
	std::vector<ProfRec> prof;         /* one record per profiled volume sub-launch          */
	float *scratch; size_t scratch_elems;
	float *spacing[2][3];              /* cell sizes per field set and axis (array coords)  */
	float *srcv_h[SRC_RING], *srcv_d[SRC_RING]; cudaEvent_t srcv_ev[SRC_RING]; size_t srcv_cap; int srcv_slot;   /* per-cell source values of one pass */
	SrcList cur_src[2]; bool cur_src_fused[2]; int cur_src_slot[2];   /* what pass_prepare set up for the pass being issued */
	uint64_t cfg_version;              /* bumped by every call that changes what a pass launches */
	bool graph_ok;                     /* CUDA-graph replay of whole passes allowed (ALISB_GRAPH, no per-source fallback path) */
	PassGraph graph[2][3];             /* [set][part] */
	FluxPlane *d_planes; double *d_partial, *d_flux, *h_flux; int flux_cap;
	std::vector<CopyPlan*> cplan;      /* resolved copy lists (ALISB_RESOLVE, default 1) */
	std::vector<StripPlan*> splan;     /* regrouped strip lists (same switch) */
	int tune_resolve;
	int tune_onewave, vol_occ[2][2][2][2][8];   /* launch-bound grids: volume launches sized to one wave of resident CTAs (ALISB_ONEWAVE, default 0: measured slower); resident CTAs per SM per variant */
	int tune_aux_aside;                /* TFSF auxiliary worlds step on a side stream beside the volume launch (ALISB_AUX_ASIDE, default 1) */
	/* mailbox: small results are stored by the kernel into mapped pinned host memory and announced by a sequence number
	 * (ALISB_MAILBOX, default 1) instead of travelling by cudaMemcpy + stream synchronisation */
	int tune_mail; unsigned char *mail_h, *mail_d; unsigned mail_seq; unsigned *d_mail_counters; double *d_mail_partial;
};

struct alisb_phaser {
	alisb_engine *e;
	int box[6]; int mask, nfreq;
	Box face[6]; size_t cells[6];
	float2 *acc[6];
	SurfJob *d_jobs; float2 *d_w; double *d_sums; size_t jobs_cap, w_cap;   /* far-field surface sums (grown on demand) */
};

/* fused stepping defers the E pass (see alisb_step_part): every entry point that reads the device
 * state calls SETTLE, every one that changes fields or the configuration calls UNFUSE (the H' already
 * written for the fused box would be stale) */
static int launch_pending(alisb_engine *e, bool fused);
static bool ade_in_kernel(const alisb_engine *e);
/* halo planes in flight (alisb_halo_exchange): a pass of `set` reads the ghost plane of the OTHER field set and rewrites
 * the plane of its own set that was sent; anything else that looks at the device state waits for both */
static int halo_fence(alisb_engine *e, int set, int part = ALISB_PART_ALL)
{
	for (int q=0; q<2; q++) {
		if (e->wait_in[q] && (set < 0 || q == 1 - set)) { CU(cudaStreamWaitEvent(e->stream, e->ev_in[q], 0)); e->wait_in[q] = false; }
		/* the rest of a split pass does not touch the plane that has just been sent: it must not wait for the copy */
		if (e->wait_out[q] && (set < 0 || (q == set && part != ALISB_PART_REST))) { CU(cudaStreamWaitEvent(e->stream, e->ev_out[q], 0)); e->wait_out[q] = false; }
	}
	return ALISB_OK;
}
#define SETTLE(e) do { if (!(e)) return fail(ALISB_EINVAL, "null engine"); TRY(launch_pending((e), false)); TRY(halo_fence((e), -1)); } while (0)
#define UNFUSE(e) do { if (!(e)) return fail(ALISB_EINVAL, "null engine"); TRY(launch_pending((e), false)); TRY(halo_fence((e), -1)); (e)->h_fused = false; } while (0)

extern "C" int alisb_abi_version(void) { return ALISB_ABI_VERSION; }
extern "C" const char *alisb_last_error(void) { return g_err; }
extern "C" const char *alisb_backend(void) { return "cuda-sm_100a"; }

static inline int div_up(size_t a, size_t b) { return (int)((a + b - 1)/b); }
static inline Box mkbox(const int32_t b[6]) { Box B; B.a0=b[0]; B.b0=b[2]; B.c0=b[4]; B.na=b[1]-b[0]+1; B.nb=b[3]-b[2]+1; B.nc=b[5]-b[4]+1; if (B.na<0) B.na=0; if (B.nb<0) B.nb=0; if (B.nc<0) B.nc=0; return B; }
static inline size_t bcount(const Box &B) { return (size_t)B.na*B.nb*B.nc; }
/* clip a box to the plane filter [lo,hi] */
static inline Box clipc(Box B, int lo, int hi)
{
	const int c0 = std::max(B.c0, lo), c1 = std::min(B.c0 + B.nc - 1, hi);
	B.c0 = c0; B.nc = std::max(0, c1 - c0 + 1);
	return B;
}

static int dalloc(float **p, size_t n)
{
	*p = 0;
	if (!n) return ALISB_OK;
	cudaError_t e = cudaMalloc((void**)p, n*sizeof(float));
	if (e != cudaSuccess) return fail(ALISB_ENOMEM, "cudaMalloc(%zu floats): %s", n, cudaGetErrorString(e));
	e = cudaMemset(*p, 0, n*sizeof(float));
	if (e != cudaSuccess) return fail(ALISB_ECUDA, "cudaMemset: %s", cudaGetErrorString(e));
	return ALISB_OK;
}
#define TRY(call) do { int r_ = (call); if (r_ != ALISB_OK) return r_; } while (0)
#define DEV(e) CU(cudaSetDevice((e)->g.device))

/* host dense block, z fastest (reference layout)  ->  dense block, x fastest (device order);
 * the way back is done on the device (box_gather) */
static void to_xfast(const float *zf, float *xf, const Box &B)
{
	for (int a=0; a<B.na; a++) for (int b=0; b<B.nb; b++) for (int c=0; c<B.nc; c++)
		xf[((size_t)c*B.nb + b)*B.na + a] = zf[((size_t)a*B.nb + b)*B.nc + c];
}

extern "C" int alisb_create(const alisb_grid *g, alisb_engine **out)
{
	if (!g || !out || g->nx < 1 || g->ny < 1 || g->nz < 1) return fail(ALISB_EINVAL, "bad grid");
	/* ALISB_DEBUG=2: where the start-up time goes */
	const bool trace = getenv("ALISB_DEBUG") && atoi(getenv("ALISB_DEBUG")) >= 2;
	auto t_last = std::chrono::steady_clock::now();
	auto lap = [&](const char *what) {
		if (!trace) return;
		const auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[alisb] create: %-34s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
		t_last = now;
	};
	int ndev = 0;
	cudaError_t ce = cudaGetDeviceCount(&ndev);
	if (ce != cudaSuccess || ndev < 1) return fail(ALISB_ECUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(ce));
	if (g->device < 0 || g->device >= ndev) return fail(ALISB_EINVAL, "device %d out of range (%d devices)", g->device, ndev);
	lap("cudaGetDeviceCount (driver init)");
	CU(cudaSetDevice(g->device));
	CU(cudaFree(0));
	lap("cudaSetDevice + context");
	alisb_engine *e = new alisb_engine();
	e->g = *g; e->nx = g->nx; e->ny = g->ny; e->nz = g->nz;
	e->pitch = ((ALISB_XO + g->nx + 2) + 31)/32*32;
	e->plane = (size_t)(g->ny + 2)*e->pitch;
	e->elems = e->plane*(g->nz + 2) + 64;
	memset(e->F, 0, sizeof e->F); memset(e->Fb, 0, sizeof e->Fb); memset(e->eps, 0, sizeof e->eps);
	e->dbuf = false; e->pass_open[0] = e->pass_open[1] = false; e->fuse_ok = false; e->fuse_cfg = 0; e->pendingE = -1; e->fusing = false; e->h_fused = false; memset(e->psi, 0, sizeof e->psi); memset(e->psib, 0, sizeof e->psib);
	memset(e->dtd, 0, sizeof e->dtd); memset(e->P, 0, sizeof e->P); memset(e->P1, 0, sizeof e->P1); memset(e->P2, 0, sizeof e->P2);
	e->step = 0; e->t = e->tE = e->tH = 0; e->launches = 0; e->fused_launches = 0; e->profiling = false; e->ev_used = 0; e->last_ms = 0; e->last_launches = 0;
	e->scratch = 0; e->scratch_elems = 0;
	for (int r=0; r<SRC_RING; r++) { e->srcv_h[r] = 0; e->srcv_d[r] = 0; e->srcv_ev[r] = 0; }
	e->cfg_version = 1; e->cur_src_fused[0] = e->cur_src_fused[1] = false; e->cur_src_slot[0] = e->cur_src_slot[1] = -1;
	for (int a=0; a<2; a++) for (int b=0; b<3; b++) { e->graph[a][b].exec = 0; e->graph[a][b].graph = 0; e->graph[a][b].version = 0; e->graph[a][b].nkernels = 0; }
	{ const char *gs = getenv("ALISB_GRAPH"); e->graph_ok = gs ? atoi(gs) != 0 : true; }
	e->srcv_cap = 0; e->srcv_slot = 0;
	e->d_ade[0] = e->d_ade[1] = 0; e->ade_cfg = 0;
	e->d_wave_ticket = 0; e->d_wave_progress = 0; e->d_wave_error = 0; e->wave_tiles_cap = 0; e->wave_ticket_base = 0; e->wave_epoch = 0;
	{ const char *ws = getenv("ALISB_WAVE"); e->tune_wave = ws ? atoi(ws) : 0; }
	{ const char *as = getenv("ALISB_ADE_FUSE"); e->tune_ade = as ? atoi(as) : 1; }
	/* z CPML slabs inside the fused box: opt-in.  Measured on the 2048x2048x256 bench (profiles/r02_pmz_ab_v9.txt): the slab
	 * planes leave launches that run at 6.7 TB/s (yee_multi faces, 104 B per cell) for one that runs at 5.9 TB/s (80 B per
	 * cell) and whose psi loads are not staged through the ring: 11.37 ms per step against 10.59 without */
	{ const char *as = getenv("ALISB_FUSE_PMZ"); e->tune_pmz = as ? atoi(as) : 0; }
	memset(e->spacing, 0, sizeof e->spacing); e->d_planes = 0; e->d_partial = 0; e->d_flux = 0; e->h_flux = 0; e->flux_cap = 0;
	e->f_lo = 0; e->f_hi = g->nz + 1;
	CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)); e->own_stream = true;
	for (int s=0; s<2; s++) for (int m=0; m<=g->complex_field; m++) for (int c=0; c<3; c++) TRY(dalloc(&e->F[s][m][c], e->elems));
	for (int s=0; s<2; s++) {
		/* tables: x ones are stored shifted by XO (so float4 loads of a quad are aligned) */
		TRY(dalloc(&e->dtd[s][0], e->pitch + 8)); TRY(dalloc(&e->dtd[s][1], g->ny + 2)); TRY(dalloc(&e->dtd[s][2], g->nz + 2));
		TRY(dalloc(&e->P[s][0], e->pitch + 8)); TRY(dalloc(&e->P1[s][0], e->pitch + 8)); TRY(dalloc(&e->P2[s][0], e->pitch + 8));
		TRY(dalloc(&e->P[s][1], g->ny + 2)); TRY(dalloc(&e->P1[s][1], g->ny + 2)); TRY(dalloc(&e->P2[s][1], g->ny + 2));
		TRY(dalloc(&e->P[s][2], g->nz + 2)); TRY(dalloc(&e->P1[s][2], g->nz + 2)); TRY(dalloc(&e->P2[s][2], g->nz + 2));
		for (int sd=0; sd<2; sd++) {
			const int f = g->pml[s][0][sd][0], l = g->pml[s][0][sd][1];
			e->xbase[s][sd] = 1; e->xpitch[s][sd] = 4;
			if (f <= l) {
				e->xbase[s][sd] = sd == 0 ? 1 : 1 + 4*((f - 1)/4);
				e->xpitch[s][sd] = (l - e->xbase[s][sd] + 1 + 3)/4*4;
			}
			for (int ax=0; ax<3; ax++) {
				const int ff = g->pml[s][ax][sd][0], ll = g->pml[s][ax][sd][1];
				size_t n = 0;
				if (ff <= ll) {
					if (ax == 0) n = (size_t)(g->nz + 2)*(g->ny + 2)*e->xpitch[s][sd];
					if (ax == 1) n = (size_t)(g->nz + 2)*(ll - ff + 1)*e->pitch;
					if (ax == 2) n = (size_t)(ll - ff + 1)*(g->ny + 2)*e->pitch;
					n += 64;
				}
				e->psi_elems[s][ax][sd] = n;
				for (int m=0; m<=g->complex_field; m++) for (int w=0; w<2; w++) TRY(dalloc(&e->psi[s][m][ax][sd][w], n));
			}
		}
	}
	lap("stream + field/psi/table allocation");
	int nb = 0;
	CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, yee_volume<1,false,7,32,false>, 256, 0));
	lap("first kernel query (module load)");
	{
		const char *s;
		e->tune_bx = 1; e->tune_by = 8; e->tune_stream = 0; e->tune_waves = 100000; e->tune_minchunk = 8;   /* measured: profiles/sweep2_r01, sweep3_r01 */
		if ((s = getenv("ALISB_BLOCK")) && sscanf(s, "%dx%d", &e->tune_bx, &e->tune_by) == 2) { if (e->tune_bx*e->tune_by*32 > 256 || e->tune_bx < 1 || e->tune_by < 1) { e->tune_bx = 1; e->tune_by = 8; } }
		if ((s = getenv("ALISB_STREAM"))) e->tune_stream = atoi(s);
		e->tune_debug = (s = getenv("ALISB_DEBUG")) ? atoi(s) : 0;
		e->tune_conc = (s = getenv("ALISB_CONC")) ? atoi(s) : 1;
		e->tune_merge = (s = getenv("ALISB_MERGE")) ? atoi(s) : 1;
		e->fuse_req = (s = getenv("ALISB_FUSE")) ? atoi(s) : 1;      /* on by default; only grids of >= tune_fuse_min cells with a large clean box actually fuse */
		e->tune_fchunk = (s = getenv("ALISB_FCHUNK")) ? atoi(s) : 64;   /* measured (16-row tiles): 57..113 planes 11.27 ms/step, 32: 11.37 (profiles/fuse_r01.sh) */
		e->tune_frows = (s = getenv("ALISB_FROWS")) ? atoi(s) : 16;   /* measured: 16 rows 11.38 ms/step, 8 rows 11.72 */
		e->tune_fuse_frac = (s = getenv("ALISB_FUSE_FRAC")) ? atof(s) : 0.5;
		e->tune_fuse_min = (s = getenv("ALISB_FUSE_MIN")) ? atol(s) : 40000000L;   /* measured: equal at 320^3, +4 % at 384^3, +12 % at 512^3 (profiles/midfuse_r01.sh) */
		e->tune_small = (s = getenv("ALISB_SMALL")) ? atol(s) : 3000000L;   /* crossover measured at ~150^3 (profiles/small_grid_probe.py) */
		for (int q=0; q<3; q++) { CU(cudaStreamCreateWithFlags(&e->side[q], cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&e->ev_join[q], cudaEventDisableTiming)); }
		CU(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
		e->tune_prio = (s = getenv("ALISB_FUSE_PRIO")) ? atoi(s) : 1;
		e->tune_resolve = (s = getenv("ALISB_RESOLVE")) ? atoi(s) : 1;
		e->tune_mail = (s = getenv("ALISB_MAILBOX")) ? atoi(s) : 1;
		e->tune_aux_aside = (s = getenv("ALISB_AUX_ASIDE")) ? atoi(s) : 1;
		e->tune_onewave = (s = getenv("ALISB_ONEWAVE")) ? atoi(s) : 0;   /* measured slower: Nanoblock loop 0.94 s against 0.83, PhC 3.32 against 3.20 (profiles/r02_small_ab_v19.txt) */
		{ int least = 0, greatest = 0; CU(cudaDeviceGetStreamPriorityRange(&least, &greatest)); CU(cudaStreamCreateWithPriority(&e->hi, cudaStreamNonBlocking, greatest)); }
		CU(cudaEventCreateWithFlags(&e->ev_hi, cudaEventDisableTiming));
		if ((s = getenv("ALISB_WAVES"))) e->tune_waves = std::max(1, atoi(s));
		if ((s = getenv("ALISB_MINCHUNK"))) e->tune_minchunk = std::max(1, atoi(s));
	}
	cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, g->device));
	e->resident_ctas = std::max(1, nb)*prop.multiProcessorCount;
	e->sm_count = prop.multiProcessorCount;
	/* the > 48 KB dynamic shared memory opt-in is a per-device function attribute: set it for THIS engine's
	 * device (cheap; engines on several devices in one process each do it) */
	CU((fused_optin_one<8,false,false>())); CU((fused_optin_one<8,true,false>())); CU((fused_optin_one<16,false,false>())); CU((fused_optin_one<16,true,false>()));
	CU((fused_optin_one<8,false,true>()));  CU((fused_optin_one<8,true,true>()));  CU((fused_optin_one<16,false,true>()));  CU((fused_optin_one<16,true,true>()));
	CU((fused_optin_one<15,false,false>())); CU((fused_optin_one<15,true,false>())); CU((fused_optin_one<15,false,true>())); CU((fused_optin_one<15,true,true>()));
	CU(cudaFuncSetAttribute(yee_wave<16,false,3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WaveSmem<16,false,3>)));
	CU(cudaFuncSetAttribute(yee_wave<12,false,4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WaveSmem<12,false,4>)));
	CU(cudaFuncSetAttribute(yee_wave<8,false,4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WaveSmem<8,false,4>)));
	CU(cudaFuncSetAttribute(yee_wave<16,true,2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WaveSmem<16,true,2>)));
	CU(cudaFuncSetAttribute(yee_wave<8,true,3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WaveSmem<8,true,3>)));
	for (int cls=0; cls<2; cls++) {
		int per_sm = 0;
		if (cls == 0) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, yee_multi<1,false,0,false>, 256, 0));
		else          CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, yee_multi<1,false,1,false>, 256, 0));
		e->multi_resident[cls] = std::max(1, per_sm)*prop.multiProcessorCount;
	}
	lap("side streams, attributes, occupancy");
	*out = e;
	return ALISB_OK;
}

static void free_src(SrcObj *s)
{
	if (!s) return;
	cudaFree(s->dAmp);
	for (int r=0; r<SRC_RING; r++) { if (s->hvals[r]) cudaFreeHost(s->hvals[r]); cudaFree(s->dvals[r]); if (s->ev[r]) cudaEventDestroy(s->ev[r]); }
	delete s;
}

extern "C" int alisb_clear_sources(alisb_engine *e)
{
	UNFUSE(e);
	e->cfg_version++;
	DEV(e);
	CU(cudaStreamSynchronize(e->stream));
	for (int s=0; s<2; s++) { for (SrcObj *o : e->src[s]) free_src(o); e->src[s].clear(); }
	for (TfsfObj *t : e->tf) delete t;
	e->tf.clear();
	return ALISB_OK;
}

extern "C" int alisb_clear_dispersive(alisb_engine *e)
{
	UNFUSE(e);
	e->cfg_version++;
	DEV(e);
	CU(cudaStreamSynchronize(e->stream));
	for (DispObj *o : e->obj) {
		for (int c=0; c<3; c++) { cudaFree(o->fill[c]); for (int m=0; m<2; m++) for (int p=0; p<ALISB_MAXPOLES; p++) { cudaFree(o->J[m][p][c]); cudaFree(o->K[m][p][c]); } }
		delete o;
	}
	e->obj.clear();
	return ALISB_OK;
}

static void trace_print(const alisb_engine *e);
extern "C" void alisb_destroy(alisb_engine *e)
{
	if (!e) return;
	cudaSetDevice(e->g.device);
	cudaStreamSynchronize(e->stream);
	if (e->comm) cudaStreamSynchronize(e->comm);
	if (e->tune_debug) trace_print(e);
	e->pendingE = -1; e->h_fused = false;
	e->wait_in[0] = e->wait_in[1] = e->wait_out[0] = e->wait_out[1] = false;     /* a neighbour's events may be gone already */
	alisb_clear_sources(e);
	alisb_clear_dispersive(e);
	for (int s=0; s<2; s++) for (int m=0; m<2; m++) for (int c=0; c<3; c++) { cudaFree(e->F[s][m][c]); cudaFree(e->Fb[s][m][c]); }
	for (int c=0; c<3; c++) cudaFree(e->eps[c]);
	for (int s=0; s<2; s++) for (int ax=0; ax<3; ax++) { cudaFree(e->dtd[s][ax]); cudaFree(e->P[s][ax]); cudaFree(e->P1[s][ax]); cudaFree(e->P2[s][ax]); }
	for (int s=0; s<2; s++) for (int m=0; m<2; m++) for (int ax=0; ax<3; ax++) for (int sd=0; sd<2; sd++) for (int w=0; w<2; w++) cudaFree(e->psi[s][m][ax][sd][w]);
	for (int s=0; s<2; s++) for (int m=0; m<2; m++) for (int sd=0; sd<2; sd++) for (int w=0; w<2; w++) cudaFree(e->psib[s][m][sd][w]);
	for (int a=0; a<2; a++) for (int b=0; b<3; b++) { if (e->graph[a][b].exec) cudaGraphExecDestroy(e->graph[a][b].exec); if (e->graph[a][b].graph) cudaGraphDestroy(e->graph[a][b].graph); }
	for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
	for (int q=0; q<3; q++) { cudaStreamDestroy(e->side[q]); cudaEventDestroy(e->ev_join[q]); }
	cudaEventDestroy(e->ev_fork);
	if (e->hi) cudaStreamDestroy(e->hi);
	if (e->ev_hi) cudaEventDestroy(e->ev_hi);
	if (e->comm) cudaStreamDestroy(e->comm);
	if (e->ev_plane) cudaEventDestroy(e->ev_plane);
	if (e->ev_war) cudaEventDestroy(e->ev_war);
	for (int q=0; q<2; q++) if (e->ev_out[q]) cudaEventDestroy(e->ev_out[q]);
	cudaFree(e->scratch); cudaFree(e->d_ade[0]); cudaFree(e->d_ade[1]);
	cudaFree(e->d_wave_ticket); cudaFree(e->d_wave_progress); cudaFree(e->d_wave_error);
	for (int q=0; q<2; q++) for (int ax=0; ax<3; ax++) cudaFree(e->spacing[q][ax]);
	for (int r=0; r<SRC_RING; r++) { if (e->srcv_h[r]) cudaFreeHost(e->srcv_h[r]); cudaFree(e->srcv_d[r]); if (e->srcv_ev[r]) cudaEventDestroy(e->srcv_ev[r]); }
	cudaFree(e->d_planes); cudaFree(e->d_partial); cudaFree(e->d_flux); if (e->h_flux) cudaFreeHost(e->h_flux);
	for (CopyPlan *cp : e->cplan) { cudaFree(cp->d_ent); cudaFree(cp->d_tmp); delete cp; }
	for (StripPlan *sp : e->splan) { cudaFree(sp->d_cells); cudaFree(sp->d_terms); delete sp; }
	if (e->mail_h) cudaFreeHost(e->mail_h);
	cudaFree(e->d_mail_counters); cudaFree(e->d_mail_partial);
	if (e->own_stream) cudaStreamDestroy(e->stream);
	delete e;
}

extern "C" int alisb_set_stream(alisb_engine *e, void *s)
{
	UNFUSE(e);
	e->cfg_version++;
	DEV(e);
	CU(cudaStreamSynchronize(e->stream));
	if (e->own_stream) { cudaStreamDestroy(e->stream); e->own_stream = false; }
	if (s) e->stream = (cudaStream_t)s;
	else { CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)); e->own_stream = true; }
	return ALISB_OK;
}

static int upload_table(alisb_engine *e, float *dst, const float *src, int n, int shift)
{
	std::vector<float> h((size_t)n + 2 + shift, 0.f);
	if (src) for (int q=0; q<n+2; q++) h[q + shift] = src[q];
	CU(cudaMemcpyAsync(dst, h.data(), h.size()*sizeof(float), cudaMemcpyHostToDevice, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	return ALISB_OK;
}

extern "C" int alisb_set_axis_tables(alisb_engine *e, int set, int axis, const float *dtd, const float *P, const float *P1, const float *P2)
{
	UNFUSE(e);
	e->cfg_version++;
	if (set<0 || set>1 || axis<0 || axis>2 || !dtd) return fail(ALISB_EINVAL, "bad tables");
	DEV(e);
	const int n = axis==0 ? e->nx : axis==1 ? e->ny : e->nz, shift = axis==0 ? ALISB_XO : 0;
	TRY(upload_table(e, e->dtd[set][axis], dtd, n, shift));
	TRY(upload_table(e, e->P[set][axis], P, n, shift));
	TRY(upload_table(e, e->P1[set][axis], P1, n, shift));
	TRY(upload_table(e, e->P2[set][axis], P2, n, shift));
	return ALISB_OK;
}

extern "C" int alisb_set_inv_eps(alisb_engine *e, const float *cx, const float *cy, const float *cz)
{
	UNFUSE(e);
	e->cfg_version++;
	DEV(e);
	CU(cudaStreamSynchronize(e->stream));
	const float *src[3] = {cx, cy, cz};
	if (!cx && !cy && !cz) { for (int c=0; c<3; c++) { cudaFree(e->eps[c]); e->eps[c] = 0; } return ALISB_OK; }
	if (!cx || !cy || !cz) return fail(ALISB_EINVAL, "inv_eps needs all three components");
	std::vector<float> h(e->elems);
	for (int c=0; c<3; c++) {
		if (!e->eps[c]) TRY(dalloc(&e->eps[c], e->elems));
		int hi[3] = {e->nx, e->ny, e->nz}; hi[c]++;
		std::fill(h.begin(), h.end(), 0.f);
		const size_t sb = (size_t)hi[2], sa = (size_t)hi[1]*hi[2];
		#pragma omp parallel for
		for (int k=1; k<=hi[2]; k++) for (int b=1; b<=hi[1]; b++) for (int a=1; a<=hi[0]; a++)
			h[((size_t)k*(e->ny+2) + b)*e->pitch + ALISB_XO + a] = src[c][(size_t)(a-1)*sa + (size_t)(b-1)*sb + (k-1)];
		CU(cudaMemcpy(e->eps[c], h.data(), e->elems*sizeof(float), cudaMemcpyHostToDevice));
	}
	return ALISB_OK;
}

static int upload_box(float **dst, const float *zfast, const Box &B)
{
	const size_t n = bcount(B);
	TRY(dalloc(dst, n));
	if (!n) return ALISB_OK;
	std::vector<float> h(n);
	to_xfast(zfast, h.data(), B);
	CU(cudaMemcpy(*dst, h.data(), n*sizeof(float), cudaMemcpyHostToDevice));
	return ALISB_OK;
}

extern "C" int alisb_add_dispersive(alisb_engine *e, const alisb_dispersive *d)
{
	UNFUSE(e);
	e->cfg_version++;
	if (!d || d->nk > ALISB_MAXPOLES || d->nj > d->nk || d->nk < 0) return fail(ALISB_EINVAL, "bad pole counts");
	DEV(e);
	DispObj *o = new DispObj();
	memset(o->fill, 0, sizeof o->fill); memset(o->J, 0, sizeof o->J); memset(o->K, 0, sizeof o->K);
	o->d = *d;
	o->full = mkbox(d->box);
	int32_t u[6] = {std::max(d->box[0],1), std::min(d->box[1],e->nx), std::max(d->box[2],1), std::min(d->box[3],e->ny), std::max(d->box[4],1), std::min(d->box[5],e->nz)};
	o->upd = mkbox(u);
	o->cnt = bcount(o->full);
	int rc = ALISB_OK;
	for (int c=0; c<3 && rc == ALISB_OK; c++) {
		rc = upload_box(&o->fill[c], d->fill[c], o->full);
		for (int m=0; m<=e->g.complex_field && rc == ALISB_OK; m++) for (int p=0; p<d->nk && rc == ALISB_OK; p++) {
			rc = dalloc(&o->J[m][p][c], o->cnt);
			if (rc == ALISB_OK && p >= d->nj) rc = dalloc(&o->K[m][p][c], o->cnt);
		}
	}
	if (rc != ALISB_OK) {                                    /* release the partially built object */
		for (int c=0; c<3; c++) { cudaFree(o->fill[c]); for (int m=0; m<2; m++) for (int p=0; p<ALISB_MAXPOLES; p++) { cudaFree(o->J[m][p][c]); cudaFree(o->K[m][p][c]); } }
		delete o;
		return rc;
	}
	e->obj.push_back(o);
	return ALISB_OK;
}

extern "C" int alisb_add_source(alisb_engine *e, const alisb_source *s)
{
	UNFUSE(e);
	e->cfg_version++;
	if (!s || s->set<0 || s->set>1 || s->comp<0 || s->comp>2 || s->m<0 || s->m>e->g.complex_field || !s->waveform) return fail(ALISB_EINVAL, "bad source");
	DEV(e);
	SrcObj *o = new SrcObj();
	o->d = *s; o->B = mkbox(s->box); o->cnt = bcount(o->B); o->dAmp = 0; o->slot = 0;
	for (int r=0; r<SRC_RING; r++) { o->hvals[r] = 0; o->dvals[r] = 0; o->ev[r] = 0; }
	auto build = [&]() -> int {
		if (s->Amplitude) {
			o->hAmp.resize(o->cnt); to_xfast(s->Amplitude, o->hAmp.data(), o->B);
			TRY(dalloc(&o->dAmp, o->cnt));
			CU(cudaMemcpy(o->dAmp, o->hAmp.data(), o->cnt*sizeof(float), cudaMemcpyHostToDevice));
		}
		if (s->Phase) {
			o->hPha.resize(o->cnt); to_xfast(s->Phase, o->hPha.data(), o->B);
			for (int r=0; r<SRC_RING; r++) {
				CU(cudaMallocHost((void**)&o->hvals[r], o->cnt*sizeof(float)));
				TRY(dalloc(&o->dvals[r], o->cnt));
				CU(cudaEventCreateWithFlags(&o->ev[r], cudaEventDisableTiming));
			}
		}
		return ALISB_OK;
	};
	const int rc = build();
	if (rc != ALISB_OK) { free_src(o); return rc; }         /* release the partially built source */
	o->d.Amplitude = 0; o->d.Phase = 0;
	e->src[s->set].push_back(o);
	return ALISB_OK;
}

extern "C" int alisb_add_tfsf(alisb_engine *e, const alisb_tfsf *t)
{
	UNFUSE(e);
	e->cfg_version++;
	if (!t || !t->aux) return fail(ALISB_EINVAL, "bad tfsf");
	TfsfObj *o = new TfsfObj();
	o->d = *t;
	o->sE.assign(t->strips_E, t->strips_E + t->n_strips_E);
	o->sH.assign(t->strips_H, t->strips_H + t->n_strips_H);
	e->tf.push_back(o);
	return ALISB_OK;
}

extern "C" int alisb_set_copies(alisb_engine *e, int set, const alisb_copy *ops, int n)
{
	UNFUSE(e);
	e->cfg_version++;
	if (set<0 || set>1 || n<0) return fail(ALISB_EINVAL, "bad copies");
	e->copies[set].assign(ops, ops + n);
	return ALISB_OK;
}

/* ---- one pass ------------------------------------------------------------------------- */
/* occ != 0: no launch, *occ = resident CTAs per SM of the variant that would be launched */
#define VOL_GO(PMV, LANES) do { if (occ) cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, yee_volume<S,EPS,PMV,LANES,ADE>, 256, 0); \
                                else yee_volume<S,EPS,PMV,LANES,ADE><<<grid,block,0,st>>>(p, r); } while (0)
template <int S, bool EPS, bool ADE>
static void launch_volume_te(cudaStream_t st, const VolParams &p, const VolRegion &r, dim3 grid, dim3 block, int *occ)
{
	if (r.narrow) {                   /* 32-cell wide x strips: 8 lanes along x, 4 rows per warp */
		if (r.pm == 1) VOL_GO(1, 8);
		else           VOL_GO(5, 8);
		return;
	}
	switch (r.pm) {
	case 0: VOL_GO(0, 32); break;
	case 1: VOL_GO(1, 32); break;
	case 2: VOL_GO(2, 32); break;
	case 3: VOL_GO(3, 32); break;
	case 4: VOL_GO(4, 32); break;
	case 5: VOL_GO(5, 32); break;
	case 6: VOL_GO(6, 32); break;
	default: VOL_GO(7, 32); break;
	}
}
/* dispersive objects sit in material worlds: the ADE variants exist with the inverse-permittivity arrays only */
template <int S>
static void launch_volume_t(cudaStream_t st, const VolParams &p, const VolRegion &r, dim3 grid, dim3 block, bool eps, bool ade, int *occ = 0)
{
	if (ade)      launch_volume_te<S,(S > 0),true>(st, p, r, grid, block, occ);
	else if (eps) launch_volume_te<S,(S > 0),false>(st, p, r, grid, block, occ);
	else          launch_volume_te<S,false,false>(st, p, r, grid, block, occ);
}
template <int S, bool EPS, bool ADE>
static void launch_multi_te(cudaStream_t st, const MultiParams &mp, unsigned nctas, int cls)
{
	if (cls == 0) yee_multi<S,EPS,0,ADE><<<nctas, dim3(32, 8, 1), 0, st>>>(mp);
	else          yee_multi<S,EPS,1,ADE><<<nctas, dim3(32, 8, 1), 0, st>>>(mp);
}
static void launch_multi(cudaStream_t st, const MultiParams &mp, unsigned nctas, int cls, int set, bool eps, bool ade)
{
	if (set == ALISB_SET_E) { if (ade) launch_multi_te<1,true,true>(st, mp, nctas, cls); else if (eps) launch_multi_te<1,true,false>(st, mp, nctas, cls); else launch_multi_te<1,false,false>(st, mp, nctas, cls); }
	else                    { if (ade) launch_multi_te<-1,false,true>(st, mp, nctas, cls); else launch_multi_te<-1,false,false>(st, mp, nctas, cls); }
}

/* ---- fused E+H stepping (temporal blocking) ------------------------------------------------------
 * Where a step is nothing but the two curls -- no CPML slab, no material array, no ADE object, no
 * source / TFSF strip / boundary-copy target -- E' and H' of a cell can be formed in one pass over
 * memory (kernel yee_fused).  fuse_eval finds the largest such box R and its clean z ranges for the
 * current configuration.  R needs plain cells on its low sides (the kernel recomputes E' on the
 * halo a-1, b-1, c-1), so it starts one cell above the low slabs; its x range is a whole number of
 * 32-cell groups.  Everything outside R keeps the two-pass kernels.  Fields are double buffered in
 * this mode: every pass reads Fb and writes F, the pointers swap when a pass begins.             */
static void fuse_eval(alisb_engine *e)
{
	if (e->fuse_cfg == e->cfg_version) return;
	e->fuse_cfg = e->cfg_version;
	const bool was = e->fuse_ok;
	e->fuse_ok = false; e->fuse_z.clear();
	auto done = [&]() { if (was != e->fuse_ok) { e->plan[0].clear(); e->plan[1].clear(); } };
	const long cells = (long)e->nx*e->ny*e->nz;
	if (!e->fuse_req || cells < e->tune_fuse_min || cells < e->tune_small || e->tune_bx != 1 || e->tune_by != 8) return done();
	const int n[3] = {e->nx, e->ny, e->nz};
	int lo[3], hi[3];
	for (int ax=0; ax<3; ax++) {
		lo[ax] = 2; hi[ax] = n[ax];                 /* the low halo (cell 1 at least) must be an updated cell */
		for (int set=0; set<2; set++) {
			const int f0 = e->g.pml[set][ax][0][0], l0 = e->g.pml[set][ax][0][1], f1 = e->g.pml[set][ax][1][0], l1 = e->g.pml[set][ax][1][1];
			if (f0 <= l0) lo[ax] = std::max(lo[ax], l0 + (set == ALISB_SET_E ? 2 : 1));   /* E' is also formed on the halo cell below */
			if (f1 <= l1) hi[ax] = std::min(hi[ax], f1 - 1);
		}
	}
	if (e->tune_pmz && !e->tune_wave) { lo[2] = 2; hi[2] = n[2]; }   /* yee_fused<..,PMZ> carries the z slabs (the halo plane 1 is an updated cell) */
	hi[2] = std::min(hi[2], e->nz - 1);             /* plane nz is the E send plane of a sharded run (alisb_step_part) */
	/* x in whole 32-cell groups starting at a = 1 mod 32 */
	int a0 = (lo[0] - 1 + 31)/32*32 + 1;
	int a1 = a0 - 1 + (hi[0] - a0 + 1)/32*32;
	if (a1 < a0 || lo[1] > hi[1] || lo[2] > hi[2]) return done();
	e->R[0] = a0; e->R[1] = a1; e->R[2] = lo[1]; e->R[3] = hi[1]; e->R[4] = lo[2]; e->R[5] = hi[2];
	/* boundary copies on E change E after the curl: none may land in R or on its low halo (H' of R
	 * was formed from the plain E').  Copies on H are applied after the H pass, on top of H'.      */
	for (const alisb_copy &k : e->copies[ALISB_SET_E]) {
		bool hit = true;
		for (int ax=0; ax<3; ax++) if (k.box[2*ax+1] < e->R[2*ax] - 1 || k.box[2*ax] > e->R[2*ax+1]) hit = false;
		if (hit) return done();
	}
	/* planes whose H' depends on an E that a fix-up changes after the curl: sources on E, TFSF E strips */
	std::vector<char> dirty(e->nz + 3, 0);
	auto mark = [&](int c0, int c1) { for (int c=std::max(c0, 0); c<=std::min(c1 + 1, e->nz + 1); c++) dirty[c] = 1; };
	for (SrcObj *so : e->src[ALISB_SET_E]) mark(so->d.box[4], so->d.box[5]);
	for (TfsfObj *t : e->tf) for (const alisb_strip &st : t->sE) mark(st.box[4], st.box[5]);
	/* dispersive objects handled inside the two-pass kernels: their planes (and the plane above, whose H'
	 * the fused kernel would form from an E' recomputed without the currents) stay two-pass */
	if (ade_in_kernel(e)) for (DispObj *o : e->obj) mark(o->upd.c0, o->upd.c0 + o->upd.nc - 1);
	long fused_planes = 0;
	for (int c=lo[2]; c<=hi[2]; ) {
		if (dirty[c]) { c++; continue; }
		int d = c;
		while (d + 1 <= hi[2] && !dirty[d + 1]) d++;
		if (d - c + 1 >= 4) { e->fuse_z.push_back(std::make_pair(c, d)); fused_planes += d - c + 1; }
		c = d + 1;
	}
	const double frac = (double)(a1 - a0 + 1)*(hi[1] - lo[1] + 1)*fused_planes/(double)cells;
	if (frac < e->tune_fuse_frac) { e->fuse_z.clear(); return done(); }   /* not worth doubling the field memory */
	e->fuse_ok = true;
	if (e->tune_debug) fprintf(stderr, "[alisb] fused box a %d..%d b %d..%d c %d..%d, %zu clean z range(s), %.1f%% of the grid\n", a0, a1, lo[1], hi[1], lo[2], hi[2], e->fuse_z.size(), 100*frac);
	done();
}

/* switch to double-buffered passes (sticky): the second buffer starts as a copy so that ghost cells
 * and not-yet-updated cells agree */
static int enable_dbuf(alisb_engine *e)
{
	if (e->dbuf) return ALISB_OK;
	/* twice the field memory: only when it fits with room to spare */
	size_t free_b = 0, total_b = 0;
	CU(cudaMemGetInfo(&free_b, &total_b));
	const size_t need = (size_t)6*(e->g.complex_field + 1)*e->elems*sizeof(float);
	if (need + (total_b >> 4) > free_b) {
		if (e->tune_debug) fprintf(stderr, "[alisb] fused stepping off: %.1f GB more field memory needed, %.1f GB free\n", need*1e-9, free_b*1e-9);
		e->fuse_req = 0; e->fuse_ok = false; e->fuse_z.clear(); e->plan[0].clear(); e->plan[1].clear();
		return ALISB_OK;
	}
	for (int s=0; s<2; s++) for (int m=0; m<=e->g.complex_field; m++) for (int c=0; c<3; c++) {
		TRY(dalloc(&e->Fb[s][m][c], e->elems));
		CU(cudaMemcpyAsync(e->Fb[s][m][c], e->F[s][m][c], e->elems*sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
	}
	/* the z-slab psi arrays follow the same protocol (every kernel of a pass reads the old buffer and writes the
	 * current one): the fused kernel recomputes E' on halo cells whose psi the owning CTA may already have replaced */
	for (int s=0; s<2; s++) for (int m=0; m<=e->g.complex_field; m++) for (int sd=0; sd<2; sd++) for (int w=0; w<2; w++) {
		if (!e->psi[s][m][2][sd][w] || !e->psi_elems[s][2][sd]) continue;
		TRY(dalloc(&e->psib[s][m][sd][w], e->psi_elems[s][2][sd]));
		CU(cudaMemcpyAsync(e->psib[s][m][sd][w], e->psi[s][m][2][sd][w], e->psi_elems[s][2][sd]*sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
	}
	e->dbuf = true;
	e->cfg_version++; e->fuse_cfg = e->cfg_version;      /* drops captured graphs; feasibility is unchanged */
	return ALISB_OK;
}

/* a pass of `set` begins: in double-buffered mode the buffers swap once per pass */
static void open_pass(alisb_engine *e, int set)
{
	if (!e->dbuf || e->pass_open[set]) return;
	for (int m=0; m<=e->g.complex_field; m++) for (int c=0; c<3; c++) std::swap(e->F[set][m][c], e->Fb[set][m][c]);
	for (int m=0; m<=e->g.complex_field; m++) for (int sd=0; sd<2; sd++) for (int w=0; w<2; w++) if (e->psib[set][m][sd][w]) std::swap(e->psi[set][m][2][sd][w], e->psib[set][m][sd][w]);
	e->pass_open[set] = true;
}

/* Region-specialised launches.  Per axis the cells split into lo slab / slab-free middle / hi slab
 * (x in units of 128-cell tiles: tiles that contain slab cells vs. the rest).  Every region runs
 * the yee_volume instantiation that only carries the PML code (and registers) of the slabs it can
 * touch: the interior gets the lean PM=0 kernel, the big x-edge columns PM=1, the z-slab planes
 * away from the x/y edges PM=4, the thin edge/corner regions the full variants.  Cells are
 * independent within a pass, so the split cannot change results.                          */
static void plan_volume(alisb_engine *e, int set, std::vector<SubBox> &L)
{
	L.clear();
	int lo[3], hi[3];                 /* slab-free range per axis in array coords */
	const int n[3] = {e->nx, e->ny, e->nz};
	for (int ax=0; ax<3; ax++) {
		lo[ax] = 1; hi[ax] = n[ax];
		if (e->g.pml[set][ax][0][0] <= e->g.pml[set][ax][0][1]) lo[ax] = e->g.pml[set][ax][0][1] + 1;
		if (e->g.pml[set][ax][1][0] <= e->g.pml[set][ax][1][1]) hi[ax] = e->g.pml[set][ax][1][0] - 1;
		if (lo[ax] > hi[ax] + 1) lo[ax] = hi[ax] + 1;
	}
	/* x: the slab strips are widened to whole 32-cell (128-byte) groups so that every launch starts
	 * on a sector boundary and on a quad boundary (a = 1 mod 4) */
	int xl = 0, xh = e->nx + 1;       /* strips [1,xl] and [xh,nx] may hold x-slab cells */
	if (lo[0] > 1)     xl = std::min(e->nx, (lo[0] - 1 + 31)/32*32);
	if (hi[0] < e->nx) xh = std::max(xl + 1, hi[0]/32*32 + 1);
	struct Range { int a, b, pm; };
	std::vector<Range> X, Y, Z;
	const long cells = (long)e->nx*e->ny*e->nz;
	if (cells < e->tune_small) {
		/* small grids are launch-bound: one launch with every slab the grid has */
		int pm = 0;
		for (int ax=0; ax<3; ax++) if (lo[ax] > 1 || hi[ax] < n[ax]) pm |= 1 << ax;
		L.push_back({1, e->nx, 1, e->ny, 1, e->nz, pm, 0});
		return;
	}
	/* does [a,b] on axis ax hold slab cells of this field set? */
	auto slabby = [&](int ax, int a, int b) -> bool {
		for (int sd=0; sd<2; sd++) { const int f = e->g.pml[set][ax][sd][0], l = e->g.pml[set][ax][sd][1]; if (f <= l && a <= l && b >= f) return true; }
		return false; };
	if (e->fuse_ok) {
		/* the centre box is the fused box R (same for both field sets); what lies outside it is cut
		 * as usual and carries a PML bit wherever it holds slab cells */
		const int *R = e->R;
		if (R[0] > 1)       X.push_back({1, R[0] - 1, slabby(0, 1, R[0] - 1) ? 1 : 0});
		X.push_back({R[0], R[1], 0});
		if (R[1] < e->nx)   X.push_back({R[1] + 1, e->nx, slabby(0, R[1] + 1, e->nx) ? 1 : 0});
		if (R[2] > 1)       Y.push_back({1, R[2] - 1, slabby(1, 1, R[2] - 1) ? 2 : 0});
		Y.push_back({R[2], R[3], 0});
		if (R[3] < e->ny)   Y.push_back({R[3] + 1, e->ny, slabby(1, R[3] + 1, e->ny) ? 2 : 0});
		if (R[4] > 1)       Z.push_back({1, R[4] - 1, slabby(2, 1, R[4] - 1) ? 4 : 0});
		{
			/* the fused box may reach into the z slabs (yee_fused<..,PMZ>): its z range is cut where the slabs of
			 * either field set end / begin, so that what runs two-pass beside it (x and y strips, planes a fix-up
			 * touches) carries the z code only on slab planes */
			int zl = R[4] - 1, zh = R[5] + 1;
			for (int st=0; st<2; st++) {
				if (e->g.pml[st][2][0][0] <= e->g.pml[st][2][0][1]) zl = std::max(zl, std::min(R[5], e->g.pml[st][2][0][1]));
				if (e->g.pml[st][2][1][0] <= e->g.pml[st][2][1][1]) zh = std::min(zh, std::max(R[4], e->g.pml[st][2][1][0]));
			}
			if (zl >= zh) { zl = R[4] - 1; zh = R[5] + 1; Z.push_back({R[4], R[5], slabby(2, R[4], R[5]) ? 4 : 0}); }
			else {
				if (zl >= R[4]) Z.push_back({R[4], zl, slabby(2, R[4], zl) ? 4 : 0});
				if (zl + 1 <= zh - 1) Z.push_back({zl + 1, zh - 1, slabby(2, zl + 1, zh - 1) ? 4 : 0});
				if (zh <= R[5]) Z.push_back({zh, R[5], slabby(2, zh, R[5]) ? 4 : 0});
			}
		}
		if (R[5] < e->nz)   Z.push_back({R[5] + 1, e->nz, slabby(2, R[5] + 1, e->nz) ? 4 : 0});
	} else {
		if (xl >= 1)         X.push_back({1, xl, 1});
		if (xl + 1 <= xh - 1) X.push_back({xl + 1, xh - 1, 0});
		if (xh <= e->nx)     X.push_back({xh, e->nx, 1});
		if (lo[1] > 1)       Y.push_back({1, lo[1] - 1, 2});
		if (lo[1] <= hi[1])  Y.push_back({lo[1], hi[1], 0});
		if (hi[1] < e->ny)   Y.push_back({hi[1] + 1, e->ny, 2});
		if (lo[2] > 1)       Z.push_back({1, lo[2] - 1, 4});
		if (lo[2] <= hi[2])  Z.push_back({lo[2], hi[2], 0});
		if (hi[2] < e->nz)   Z.push_back({hi[2] + 1, e->nz, 4});
	}
	const int xany = slabby(0, 1, e->nx) ? 1 : 0;
	for (size_t zi=0; zi<Z.size(); zi++) for (size_t yi=0; yi<Y.size(); yi++) {
		const Range &z = Z[zi], &y = Y[yi];
		const bool ymid = e->fuse_ok ? (y.a == e->R[2] && y.b == e->R[3]) : y.pm == 0;
		if (!ymid && (y.b - y.a + 1) <= 64) {      /* thin y rows: one launch over the whole row with the x code included */
			L.push_back({1, e->nx, y.a, y.b, z.a, z.b, xany | y.pm | z.pm, 0});
			continue;
		}
		for (const Range &x : X) {
			const int fus = e->fuse_ok && x.a == e->R[0] && x.b == e->R[1] && y.a == e->R[2] && y.b == e->R[3] && z.a >= e->R[4] && z.b <= e->R[5];
			L.push_back({x.a, x.b, y.a, y.b, z.a, z.b, x.pm | y.pm | z.pm, fus});
		}
	}
	/* largest region first (it is also the dominant kernel reported by alisb_profile_dominant) */
	std::stable_sort(L.begin(), L.end(), [](const SubBox &p, const SubBox &q) {
		return (double)(p.a1-p.a0+1)*(p.b1-p.b0+1)*(p.c1-p.c0+1) > (double)(q.a1-q.a0+1)*(q.b1-q.b0+1)*(q.c1-q.c0+1); });
}

/* dispersive objects are handled inside the volume kernels (ADE variants of yee_volume / yee_multi) when
 * the world has its material arrays and at most ALISB_MAXADE objects; otherwise the three box kernels
 * ade_loss / ade_subtract / ade_currents run around the volume launches */
static bool ade_in_kernel(const alisb_engine *e)
{
	if (!(e->tune_ade && !e->obj.empty() && e->obj.size() <= ALISB_MAXADE && e->eps[0] && e->eps[1] && e->eps[2])) return false;
	/* grids large enough for fused E+H stepping keep the three box kernels: they act on the passes' input and
	 * output, so the fused kernel also covers the objects' planes (measured at 512^3 with a gold block on 1/8 of
	 * the interior: 46.1 Gcells/s against 40.5 with the objects' planes two-pass, profiles/r02_ade_bench_v5.txt);
	 * ALISB_ADE_FUSE=2 forces the in-kernel variant */
	const long cells = (long)e->nx*e->ny*e->nz;
	if (e->tune_ade < 2 && e->fuse_req && cells >= e->tune_fuse_min && cells >= e->tune_small) return false;
	return true;
}

static int ade_table(alisb_engine *e)
{
	if (e->ade_cfg == e->cfg_version && e->d_ade[0]) return ALISB_OK;
	for (int m=0; m<=e->g.complex_field; m++) {
		AdeTable T;
		memset(&T, 0, sizeof T);
		T.n = (int)e->obj.size();
		for (int n=0; n<T.n; n++) {
			const DispObj *o = e->obj[n];
			AdeObj &d = T.o[n];
			for (int c=0; c<3; c++) {
				d.fill[c] = o->fill[c]; d.a[c] = o->d.a[c];
				for (int q=0; q<ALISB_MAXPOLES; q++) { d.J[q][c] = o->J[m][q][c]; d.K[q][c] = o->K[m][q][c]; }
			}
			memcpy(d.JJ, o->d.JJ, sizeof d.JJ); memcpy(d.JK, o->d.JK, sizeof d.JK); memcpy(d.JE, o->d.JE, sizeof d.JE);
			memcpy(d.KJ, o->d.KJ, sizeof d.KJ); memcpy(d.KK, o->d.KK, sizeof d.KK); memcpy(d.KE, o->d.KE, sizeof d.KE);
			d.nj = o->d.nj; d.nk = o->d.nk; d.lossy = (o->d.a[0] || o->d.a[1] || o->d.a[2]) ? 1 : 0;
			d.full = o->full; d.upd = o->upd;
		}
		if (!e->d_ade[m]) CU(cudaMalloc((void**)&e->d_ade[m], sizeof(AdeTable)));
		CU(cudaStreamSynchronize(e->stream));            /* configuration changes are rare; no pass may still read the old table */
		CU(cudaMemcpy(e->d_ade[m], &T, sizeof T, cudaMemcpyHostToDevice));
	}
	e->ade_cfg = e->cfg_version;
	return ALISB_OK;
}

template <int ROWS, bool EPS> static void launch_fused_t(cudaStream_t st, const FusedParams &fp, dim3 grid, dim3 block, bool pmz)
{
	if (pmz) yee_fused<ROWS,EPS,true><<<grid, block, sizeof(FusedSmem<ROWS,EPS>), st>>>(fp);
	else     yee_fused<ROWS,EPS,false><<<grid, block, sizeof(FusedSmem<ROWS,EPS>), st>>>(fp);
}
static void launch_fused(cudaStream_t st, const FusedParams &fp, dim3 grid, dim3 block, int frows, bool eps, bool pmz)
{
	if (frows == 16) { if (eps) launch_fused_t<16,true>(st, fp, grid, block, pmz); else launch_fused_t<16,false>(st, fp, grid, block, pmz); }
	else if (frows == 15) { if (eps) launch_fused_t<15,true>(st, fp, grid, block, pmz); else launch_fused_t<15,false>(st, fp, grid, block, pmz); }
	else             { if (eps) launch_fused_t<8,true>(st, fp, grid, block, pmz);  else launch_fused_t<8,false>(st, fp, grid, block, pmz); }
}

static int launch_volume(alisb_engine *e, int set, int m)
{
	VolParams p;
	memset(&p, 0, sizeof p);
	const bool ade = ade_in_kernel(e);
	if (ade) p.ade = e->d_ade[m];
	for (int c=0; c<3; c++) {
		p.F[c] = e->F[set][m][c]; p.G[c] = e->F[1-set][m][c];
		p.Fin[c] = e->dbuf ? e->Fb[set][m][c] : e->F[set][m][c];
		p.C[c] = set == ALISB_SET_E ? e->eps[c] : 0;
		p.dtd[c] = e->dtd[set][c]; p.P[c] = e->P[set][c]; p.P1[c] = e->P1[set][c]; p.P2[c] = e->P2[set][c];
		for (int sd=0; sd<2; sd++) {
			p.slab[c][sd][0] = e->g.pml[set][c][sd][0]; p.slab[c][sd][1] = e->g.pml[set][c][sd][1];
			for (int w=0; w<2; w++) p.psi[c][sd][w] = e->psi[set][m][c][sd][w];
		}
	}
	for (int sd=0; sd<2; sd++) { p.xbase[sd] = e->xbase[set][sd]; p.xpitch[sd] = e->xpitch[set][sd]; }
	for (int sd=0; sd<2; sd++) for (int w=0; w<2; w++) p.psiz_in[sd][w] = (e->dbuf && e->psib[set][m][sd][w]) ? e->psib[set][m][sd][w] : e->psi[set][m][2][sd][w];
	p.nx = e->nx; p.ny = e->ny; p.nz = e->nz; p.pitch = e->pitch; p.plane = e->plane;
	const bool eps = set == ALISB_SET_E && e->eps[0] && e->eps[1] && e->eps[2];
	if (e->plan[set].empty()) plan_volume(e, set, e->plan[set]);

	/* Sub-launches touch disjoint cells, so they may run concurrently.  The slab-free interior is
	 * one launch of the lean kernel on the engine's stream; the PML shell boxes are gathered into
	 * two merged launches (faces / edges+corners, yee_multi) on side streams, fork/join with events.
	 * ALISB_MERGE=0: one launch per box, spread over three side streams.                        */
	int sub = -1, nlaunched = 0, rr = 0;
	bool forked = false, used[3] = {false, false, false}, hi_used = false;
	const bool conc = e->tune_conc > 0 && e->plan[set].size() > 2;
	const bool merge = e->tune_merge && e->plan[set].size() > 2 && e->tune_bx == 1 && e->tune_by == 8;
	MultiParams mp[2];
	unsigned mcta[2] = {0, 0}; double mcells[2] = {0, 0}, mplanes[2] = {0, 0}; bool made[2] = {false, false};
	if (merge) for (int q=0; q<2; q++) { mp[q].v = p; mp[q].nreg = 0; }
	if (conc && e->tune_conc == 2) { CU(cudaEventRecord(e->ev_fork, e->stream)); forked = true; }
	auto side_stream = [&](bool first_alone) -> cudaStream_t {
		if (!conc || (first_alone && nlaunched == 0 && e->tune_conc != 2)) return e->stream;
		if (!forked) { cudaEventRecord(e->ev_fork, e->stream); forked = true; }
		const int slot = rr++ % 4;                         /* slot 3 = the main stream itself */
		if (slot == 3) return e->stream;
		if (!used[slot]) { cudaStreamWaitEvent(e->side[slot], e->ev_fork, 0); used[slot] = true; }
		return e->side[slot];
	};
	auto prof_begin = [&](cudaStream_t st, int subidx, double cells) -> cudaEvent_t {
		if (!e->profiling) return (cudaEvent_t)0;
		if (e->ev_used + 2 > e->ev_pool.size()) for (int q=0; q<64; q++) { cudaEvent_t ev; cudaEventCreate(&ev); e->ev_pool.push_back(ev); }
		e->prof.push_back({e->ev_used, set, subidx, cells});
		cudaEventRecord(e->ev_pool[e->ev_used++], st);
		return e->ev_pool[e->ev_used++];
	};
	/* the fused box: in the E pass its clean z ranges run the fused E+H kernel (when the pass is
	 * issued as the fused variant), in the following H pass they are skipped; its other planes and
	 * every other box run the two-pass kernel */
	struct Piece { SubBox sb; int sub; bool fused; bool ade; };
	std::vector<Piece> work;
	for (const SubBox &sb0 : e->plan[set]) {
		sub++;
		const bool skipz = sb0.fusable && e->fuse_ok && ((set == ALISB_SET_E && e->fusing) || (set == ALISB_SET_H && e->h_fused));
		if (!skipz) { work.push_back({sb0, sub, false, ade}); continue; }
		int c = sb0.c0;
		for (const std::pair<int,int> &z0 : e->fuse_z) {
			const std::pair<int,int> z(std::max(z0.first, sb0.c0), std::min(z0.second, sb0.c1));   /* the part of the clean range inside this piece of the box */
			if (z.first > z.second) continue;
			if (z.first > c) { SubBox d = sb0; d.c0 = c; d.c1 = z.first - 1; work.push_back({d, sub, false, ade}); }
			if (set == ALISB_SET_E) { SubBox d = sb0; d.c0 = z.first; d.c1 = z.second; work.push_back({d, sub, true, false}); }
			c = z.second + 1;
		}
		if (c <= sb0.c1) { SubBox d = sb0; d.c0 = c; d.c1 = sb0.c1; work.push_back({d, sub, false, ade}); }
	}
	{
		/* fused pieces of the box that touch in z (it is planned in up to three z ranges: low slabs, slab-free,
		 * high slabs) become one launch: one march, no halo plane recomputed at the seams */
		std::vector<Piece> merged;
		std::vector<Piece> fz;
		for (const Piece &pc : work) { if (pc.fused) fz.push_back(pc); else merged.push_back(pc); }
		std::sort(fz.begin(), fz.end(), [](const Piece &p, const Piece &q) { return p.sb.c0 < q.sb.c0; });
		for (size_t q=0; q<fz.size(); q++) {
			Piece cur = fz[q];
			int big = cur.sb.c1 - cur.sb.c0;
			while (q + 1 < fz.size() && fz[q + 1].sb.c0 == cur.sb.c1 + 1) {
				q++;
				if (fz[q].sb.c1 - fz[q].sb.c0 > big) { big = fz[q].sb.c1 - fz[q].sb.c0; cur.sub = fz[q].sub; }   /* profiled under the largest part */
				cur.sb.c1 = fz[q].sb.c1; cur.sb.pml |= fz[q].sb.pml;
			}
			merged.insert(merged.begin(), cur);       /* the big launches first */
		}
		work.swap(merged);
	}
	if (ade) {
		/* The ADE variants carry the object code (and its registers: fewer resident CTAs).  Only the part of a
		 * box that can hold object cells runs them: a box is cut around the bounding box of the objects it
		 * meets (x cuts on 32-cell groups), everything else runs the plain variant.  Launch-bound grids keep
		 * their single launch. */
		const bool split = (long)e->nx*e->ny*e->nz >= e->tune_small;
		std::vector<Piece> cut;
		for (const Piece &pc : work) {
			if (pc.fused) { cut.push_back(pc); continue; }
			const SubBox &b = pc.sb;
			int u[6] = { INT_MAX, INT_MIN, INT_MAX, INT_MIN, INT_MAX, INT_MIN };
			for (const DispObj *o : e->obj) {
				const int lo[3] = { std::max(b.a0, o->upd.a0), std::max(b.b0, o->upd.b0), std::max(b.c0, o->upd.c0) };
				const int hi[3] = { std::min(b.a1, o->upd.a0 + o->upd.na - 1), std::min(b.b1, o->upd.b0 + o->upd.nb - 1), std::min(b.c1, o->upd.c0 + o->upd.nc - 1) };
				if (lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2]) continue;
				for (int q=0; q<3; q++) { u[2*q] = std::min(u[2*q], lo[q]); u[2*q+1] = std::max(u[2*q+1], hi[q]); }
			}
			Piece q = pc;
			if (u[0] > u[1]) { q.ade = false; cut.push_back(q); continue; }
			if (!split) { cut.push_back(q); continue; }
			u[0] = std::max(b.a0, (u[0] - 1)/32*32 + 1); u[1] = std::min(b.a1, (u[1] - 1)/32*32 + 32);
			auto piece = [&](int a0, int a1, int b0, int b1, int c0, int c1, bool with_ade) {
				if (a0 > a1 || b0 > b1 || c0 > c1) return;
				Piece r = pc; r.ade = with_ade;
				r.sb.a0 = a0; r.sb.a1 = a1; r.sb.b0 = b0; r.sb.b1 = b1; r.sb.c0 = c0; r.sb.c1 = c1;
				cut.push_back(r);
			};
			piece(u[0], u[1], u[2], u[3], u[4], u[5], true);
			piece(b.a0, b.a1, b.b0, b.b1, b.c0, u[4] - 1, false);
			piece(b.a0, b.a1, b.b0, b.b1, u[5] + 1, b.c1, false);
			piece(b.a0, b.a1, b.b0, u[2] - 1, u[4], u[5], false);
			piece(b.a0, b.a1, u[3] + 1, b.b1, u[4], u[5], false);
			piece(b.a0, u[0] - 1, u[2], u[3], u[4], u[5], false);
			piece(u[1] + 1, b.a1, u[2], u[3], u[4], u[5], false);
		}
		work.swap(cut);
	}
	std::vector<Piece> wave_later;          /* in-place fused pieces: launched after every two-pass piece of this E pass has finished */
	for (const Piece &pc : work) {
		sub = pc.sub;
		SubBox sb = pc.sb;
		sb.c0 = std::max(sb.c0, e->f_lo); sb.c1 = std::min(sb.c1, e->f_hi);
		if (sb.c0 > sb.c1) continue;
		if (pc.fused && e->tune_wave) { Piece q = pc; q.sb = sb; wave_later.push_back(q); continue; }
		if (pc.fused) {
			FusedParams fp;
			memset(&fp, 0, sizeof fp);
			for (int c=0; c<3; c++) {
				fp.Ei[c] = e->Fb[ALISB_SET_E][m][c]; fp.Eo[c] = e->F[ALISB_SET_E][m][c];
				fp.Hi[c] = e->F[ALISB_SET_H][m][c];  fp.Ho[c] = e->Fb[ALISB_SET_H][m][c];
				fp.dtdE[c] = e->dtd[ALISB_SET_E][c]; fp.dtdH[c] = e->dtd[ALISB_SET_H][c];
				fp.C[c] = eps ? e->eps[c] : 0;
			}
			fp.ny = e->ny; fp.pitch = e->pitch; fp.plane = e->plane;
			/* z CPML slabs inside the box: the E set's psi arrays were swapped when this pass opened (read the other
			 * buffer, write the current one); the H set's swap when the H pass opens (read the current one, write
			 * the other: like H' itself) */
			bool pmz = false;
			for (int sd=0; sd<2; sd++) {
				fp.zslabE[sd][0] = e->g.pml[ALISB_SET_E][2][sd][0]; fp.zslabE[sd][1] = e->g.pml[ALISB_SET_E][2][sd][1];
				fp.zslabH[sd][0] = e->g.pml[ALISB_SET_H][2][sd][0]; fp.zslabH[sd][1] = e->g.pml[ALISB_SET_H][2][sd][1];
				if (fp.zslabE[sd][0] <= fp.zslabE[sd][1] && fp.zslabE[sd][0] <= sb.c1 && fp.zslabE[sd][1] >= sb.c0 - 1) pmz = true;
				if (fp.zslabH[sd][0] <= fp.zslabH[sd][1] && fp.zslabH[sd][0] <= sb.c1 && fp.zslabH[sd][1] >= sb.c0) pmz = true;
				for (int w=0; w<2; w++) {
					fp.psiEi[sd][w] = e->psib[ALISB_SET_E][m][sd][w]; fp.psiEo[sd][w] = e->psi[ALISB_SET_E][m][2][sd][w];
					fp.psiHi[sd][w] = e->psi[ALISB_SET_H][m][2][sd][w]; fp.psiHo[sd][w] = e->psib[ALISB_SET_H][m][sd][w];
				}
			}
			fp.PzE[0] = e->P[ALISB_SET_E][2]; fp.PzE[1] = e->P1[ALISB_SET_E][2]; fp.PzE[2] = e->P2[ALISB_SET_E][2];
			fp.PzH[0] = e->P[ALISB_SET_H][2]; fp.PzH[1] = e->P1[ALISB_SET_H][2]; fp.PzH[2] = e->P2[ALISB_SET_H][2];
			fp.a_first = sb.a0; fp.a_last = sb.a1; fp.nseg = div_up(sb.a1 - sb.a0 + 1, 128);
			fp.b_first = sb.b0; fp.b_last = sb.b1; fp.c_first = sb.c0; fp.c_last = sb.c1;
			const int frows = e->tune_frows == 16 ? 16 : e->tune_frows == 15 ? 15 : 8;
			const int nzs = sb.c1 - sb.c0 + 1, gyf = div_up(sb.b1 - sb.b0 + 1, frows);
			fp.cchunk = std::min(nzs, std::max(1, e->tune_fchunk));
			dim3 grid(fp.nseg, gyf, div_up(nzs, fp.cchunk)), block(32, frows + 1, 1);
			cudaStream_t st;
			if (conc && e->tune_prio) {
				/* all fused pieces one after the other on the high-priority stream, forked from the engine's stream */
				if (!forked) { CU(cudaEventRecord(e->ev_fork, e->stream)); forked = true; }
				if (!hi_used) { CU(cudaStreamWaitEvent(e->hi, e->ev_fork, 0)); hi_used = true; }
				st = e->hi;
			} else st = side_stream(true);
			nlaunched++;
			/* profiling counts cell-passes: a fused launch does two passes per cell */
			cudaEvent_t e1 = prof_begin(st, sub, 2.0*(double)(sb.a1-sb.a0+1)*(sb.b1-sb.b0+1)*nzs);
			if (e->tune_debug && e->launches < 64) fprintf(stderr, "[alisb] fused E+H cells %d..%d rows %d..%d planes %d..%d grid %dx%dx%d chunk %d\n", sb.a0, sb.a1, sb.b0, sb.b1, sb.c0, sb.c1, grid.x, grid.y, grid.z, fp.cchunk);
			launch_fused(st, fp, grid, block, frows, eps, pmz);
			e->launches++; e->fused_launches++;
			if (e->profiling) CU(cudaEventRecord(e1, st));
			continue;
		}
		VolRegion r;
		memset(&r, 0, sizeof r);
		const bool narrow = (sb.pml == 1 || sb.pml == 5) && sb.a1 - sb.a0 + 1 <= 64;     /* x-PML strips */
		r.pm = sb.pml; r.narrow = narrow ? 1 : 0;
		r.a_first = sb.a0; r.a_last = sb.a1; r.nseg = div_up(sb.a1 - sb.a0 + 1, narrow ? 32 : 128);
		r.b_first = sb.b0; r.b_last = sb.b1; r.c_first = sb.c0; r.c_last = sb.c1;
		const bool merged = merge && sb.pml != 0 && mp[0].nreg < ALISB_MAXREG && mp[1].nreg < ALISB_MAXREG;
		const int bx = (narrow || merged) ? 1 : e->tune_bx, by = (narrow || merged) ? 8 : e->tune_by;
		const int nzs = sb.c1 - sb.c0 + 1, gx = div_up(r.nseg, bx), gy = div_up(sb.b1 - sb.b0 + 1, by*(narrow ? 4 : 1));
		const long tiles = (long)gx*gy;
		long want = ((long)e->tune_waves*e->resident_ctas + tiles - 1)/tiles;   /* many short z-chunks balance best  */
		long maxc = std::max(1, nzs/e->tune_minchunk);                          /* (the carry preload is an L2 hit)  */
		int nchunks = (int)std::max(1L, std::min(want, maxc));
		if (!merged && tiles*nchunks < 2L*e->resident_ctas) nchunks = (int)std::min((long)nzs, (2L*e->resident_ctas + tiles - 1)/tiles);   /* small grids: fill the GPU */
		if (!merged && e->tune_onewave && (long)e->nx*e->ny*e->nz < e->tune_small) {
			/* launch-bound grids: the launch is over when its slowest CTA is, so all CTAs should be resident at once -- as
			 * many z-chunks as the variant has resident slots (shared with the twin launch of the other complex set),
			 * not a few more (448 CTAs on 444 slots ran as two waves) */
			int &occ = e->vol_occ[set][eps ? 1 : 0][pc.ade ? 1 : 0][narrow ? 1 : 0][sb.pml & 7];
			if (!occ) {
				VolRegion probe = r;
				if (set == ALISB_SET_E) launch_volume_t<1>(0, p, probe, dim3(1), dim3(1), eps, pc.ade, &occ);
				else                    launch_volume_t<-1>(0, p, probe, dim3(1), dim3(1), false, pc.ade, &occ);
				if (occ < 1) occ = 1;
			}
			const long slots = (long)occ*e->sm_count/(e->g.complex_field && e->tune_conc > 0 ? 2 : 1);
			if (tiles*nchunks > slots && tiles <= slots) nchunks = (int)std::max(1L, slots/tiles);
		}
		r.cchunk = div_up(nzs, nchunks);
		nchunks = div_up(nzs, r.cchunk);
		r.gx = gx; r.gy = gy; r.gz = nchunks;
		const double cells = (double)(sb.a1-sb.a0+1)*(sb.b1-sb.b0+1)*(sb.c1-sb.c0+1);
		if (e->tune_debug && e->launches < 64) fprintf(stderr, "[alisb] set %d sub %d pm %d cells %d..%d rows %d..%d planes %d..%d grid %dx%dx%d chunk %d (resident %d)%s\n",
			set, sub, sb.pml, sb.a0, sb.a1, sb.b0, sb.b1, sb.c0, sb.c1, gx, gy, nchunks, r.cchunk, e->resident_ctas, merged ? " merged" : "");
		if (merged) {                                       /* chunking is decided per class below */
			const int cls = (sb.pml == 1 || sb.pml == 2 || sb.pml == 4) ? 0 : 1;
			mp[cls].r[mp[cls].nreg++] = r;
			mcells[cls] += cells; mplanes[cls] += (double)tiles*nzs;
			if (pc.ade) made[cls] = true;
			continue;
		}
		cudaStream_t st = side_stream(true);
		nlaunched++;
		cudaEvent_t e1 = prof_begin(st, sub, cells);
		dim3 grid(gx, gy, nchunks), block(32*bx, by, 1);
		if (set == ALISB_SET_E) launch_volume_t<1>(st, p, r, grid, block, eps, pc.ade);
		else                    launch_volume_t<-1>(st, p, r, grid, block, false, pc.ade);
		e->launches++;
		if (e->profiling) CU(cudaEventRecord(e1, st));
	}
	if (merge) for (int cls=0; cls<2; cls++) if (mp[cls].nreg) {
		/* z-chunk of a merged launch: tune_minchunk planes when that still gives >= 4 CTAs per resident
		 * slot over the whole class, fewer planes per chunk on mid-size grids (more, shorter marches) */
		const double slots = 4.0*e->multi_resident[cls];
		const int chunk = (int)std::max(1.0, std::min((double)e->tune_minchunk, floor(mplanes[cls]/slots)));
		for (int q=0; q<mp[cls].nreg; q++) {
			VolRegion &r = mp[cls].r[q];
			const int nzs = r.c_last - r.c_first + 1;
			r.cchunk = std::min(chunk, nzs); r.gz = div_up(nzs, r.cchunk);
			r.cta0 = mcta[cls];
			mcta[cls] += (unsigned)r.gx*r.gy*r.gz;
		}
		cudaStream_t st = side_stream(true);
		nlaunched++;
		cudaEvent_t e1 = prof_begin(st, -2 - cls, mcells[cls]);
		launch_multi(st, mp[cls], mcta[cls], cls, set, eps, made[cls]);
		e->launches++;
		if (e->profiling) CU(cudaEventRecord(e1, st));
	}
	for (int q=0; q<3; q++) if (used[q]) { CU(cudaEventRecord(e->ev_join[q], e->side[q])); CU(cudaStreamWaitEvent(e->stream, e->ev_join[q], 0)); }
	if (hi_used) { CU(cudaEventRecord(e->ev_hi, e->hi)); CU(cudaStreamWaitEvent(e->stream, e->ev_hi, 0)); }
	/* the in-place wavefront kernel: on the engine's stream, i.e. after all of the above (it reads the E' the
	 * two-pass E kernels formed on the box's low faces and the old H they left next to it) */
	for (const Piece &pc : wave_later) {
		const SubBox &sb = pc.sb;
		WaveParams wp;
		memset(&wp, 0, sizeof wp);
		for (int c=0; c<3; c++) {
			wp.E[c] = e->F[ALISB_SET_E][m][c]; wp.H[c] = e->F[ALISB_SET_H][m][c];
			wp.dtdE[c] = e->dtd[ALISB_SET_E][c]; wp.dtdH[c] = e->dtd[ALISB_SET_H][c];
			wp.C[c] = eps ? e->eps[c] : 0;
		}
		wp.ny = e->ny; wp.pitch = e->pitch; wp.plane = e->plane;
		wp.a_first = sb.a0; wp.a_last = sb.a1; wp.b_first = sb.b0; wp.b_last = sb.b1; wp.c_first = sb.c0; wp.c_last = sb.c1;
		const int nzs = sb.c1 - sb.c0 + 1;
		int frows = eps ? (e->tune_frows == 8 ? 8 : 16) : (e->tune_frows == 8 ? 8 : e->tune_frows == 12 ? 12 : 16);
		wp.cchunk = std::min(nzs, std::max(2, e->tune_fchunk));
		wp.ntx = div_up(sb.a1 - sb.a0 + 1, 128); wp.nty = div_up(sb.b1 - sb.b0 + 1, frows); wp.ntz = div_up(nzs, wp.cchunk);
		const size_t ntiles = (size_t)wp.ntx*wp.nty*wp.ntz;
		if (!e->d_wave_ticket) {
			CU(cudaMalloc((void**)&e->d_wave_ticket, sizeof(unsigned long long))); CU(cudaMemset(e->d_wave_ticket, 0, sizeof(unsigned long long)));
			CU(cudaMalloc((void**)&e->d_wave_error, sizeof(int))); CU(cudaMemset(e->d_wave_error, 0, sizeof(int)));
		}
		if (ntiles > e->wave_tiles_cap) {
			CU(cudaStreamSynchronize(e->stream));
			cudaFree(e->d_wave_progress); e->d_wave_progress = 0; e->wave_tiles_cap = 0;
			CU(cudaMalloc((void**)&e->d_wave_progress, ntiles*sizeof(unsigned long long)));
			CU(cudaMemset(e->d_wave_progress, 0, ntiles*sizeof(unsigned long long)));
			e->wave_tiles_cap = ntiles;
		}
		wp.ticket = e->d_wave_ticket; wp.ticket_base = e->wave_ticket_base; wp.progress = e->d_wave_progress;
		wp.epoch = (++e->wave_epoch) << 20; wp.error = e->d_wave_error;
		{ const char *ls = getenv("ALISB_WAVE_LAG"); wp.lag = ls ? std::max(0, atoi(ls)) : 4; }
		e->wave_ticket_base += ntiles;
		cudaEvent_t e1 = prof_begin(e->stream, pc.sub, 2.0*(double)(sb.a1-sb.a0+1)*(sb.b1-sb.b0+1)*nzs);
		if (e->tune_debug && e->launches < 64) fprintf(stderr, "[alisb] in-place fused E+H cells %d..%d rows %d..%d planes %d..%d tiles %dx%dx%d of 128x%d cells, %d planes per chunk\n", sb.a0, sb.a1, sb.b0, sb.b1, sb.c0, sb.c1, wp.ntx, wp.nty, wp.ntz, frows, wp.cchunk);
		const dim3 grid((unsigned)ntiles), block(32, frows + 2, 1);
		if (eps) { if (frows == 16) yee_wave<16,true,2><<<grid, block, sizeof(WaveSmem<16,true,2>), e->stream>>>(wp); else yee_wave<8,true,3><<<grid, block, sizeof(WaveSmem<8,true,3>), e->stream>>>(wp); }
		else if (frows == 16) yee_wave<16,false,3><<<grid, block, sizeof(WaveSmem<16,false,3>), e->stream>>>(wp);
		else if (frows == 12) yee_wave<12,false,4><<<grid, block, sizeof(WaveSmem<12,false,4>), e->stream>>>(wp);
		else                  yee_wave<8,false,4><<<grid, block, sizeof(WaveSmem<8,false,4>), e->stream>>>(wp);
		e->launches++; e->fused_launches++;
		if (e->profiling) CU(cudaEventRecord(e1, e->stream));
	}
	CU(cudaGetLastError());
	return ALISB_OK;
}

static void fill_ade(alisb_engine *e, DispObj *o, int m, AdeParams &p)
{
	memset(&p, 0, sizeof p);
	for (int c=0; c<3; c++) {
		p.F[c] = e->F[ALISB_SET_E][m][c]; p.fill[c] = o->fill[c]; p.a[c] = o->d.a[c];
		for (int s=0; s<ALISB_MAXPOLES; s++) { p.J[s][c] = o->J[m][s][c]; p.K[s][c] = o->K[m][s][c]; }
	}
	memcpy(p.JJ, o->d.JJ, sizeof p.JJ); memcpy(p.JK, o->d.JK, sizeof p.JK); memcpy(p.JE, o->d.JE, sizeof p.JE);
	memcpy(p.KJ, o->d.KJ, sizeof p.KJ); memcpy(p.KK, o->d.KK, sizeof p.KK); memcpy(p.KE, o->d.KE, sizeof p.KE);
	p.nj = o->d.nj; p.nk = o->d.nk; p.lossy = (o->d.a[0] || o->d.a[1] || o->d.a[2]) ? 1 : 0;
	p.full = o->full; p.upd = clipc(o->upd, e->f_lo, e->f_hi); p.ny = e->ny; p.pitch = e->pitch;
}

static int ade_pre(alisb_engine *e, int m)
{
	/* loss and J subtraction act on the OLD E (update.c:89-108), i.e. on the pass's input buffer */
	AdeParams p;
	for (DispObj *o : e->obj) {
		fill_ade(e, o, m, p);
		if (e->dbuf) for (int c=0; c<3; c++) p.F[c] = e->Fb[ALISB_SET_E][m][c];
		const size_t n = bcount(p.upd);
		if (!n || !p.lossy) continue;
		ade_loss<<<div_up(n, 256), 256, 0, e->stream>>>(p); e->launches++;
	}
	for (DispObj *o : e->obj) {
		fill_ade(e, o, m, p);
		if (e->dbuf) for (int c=0; c<3; c++) p.F[c] = e->Fb[ALISB_SET_E][m][c];
		const size_t n = bcount(p.upd);
		if (!n || !p.nk) continue;
		ade_subtract<<<div_up(n, 256), 256, 0, e->stream>>>(p); e->launches++;
	}
	CU(cudaGetLastError());
	return ALISB_OK;
}

static int ade_post(alisb_engine *e, int m)
{
	AdeParams p;
	for (DispObj *o : e->obj) {
		fill_ade(e, o, m, p);
		const size_t n = bcount(p.upd);
		if (!n || !p.nk) continue;
		ade_currents<<<div_up(n, 256), 256, 0, e->stream>>>(p); e->launches++;
	}
	CU(cudaGetLastError());
	return ALISB_OK;
}

/* Sources of a pass in three phases so that a pass can be replayed as a CUDA graph:
 *   prepare_sources  host: waveform samples with the caller's libm, per-cell values uploaded
 *                    (stream-ordered, before the pass), the SrcList kernel argument filled
 *   launch_sources   the source_list kernel (its node is the one updated on graph replay)
 *   finish_sources   ring-buffer bookkeeping                                                */
static int prepare_sources(alisb_engine *e, int set, float t)
{
	SrcList &L = e->cur_src[set];
	L.n = 0; L.ny = e->ny; L.pitch = e->pitch;
	e->cur_src_slot[set] = -1;
	e->cur_src_fused[set] = true;
	size_t nvals = 0;
	int nact = 0;
	for (SrcObj *s : e->src[set]) {
		if (s->cnt > 16384) e->cur_src_fused[set] = false;           /* big planes: per-source launches */
		if (!(e->step < s->d.n_active) || !s->cnt) continue;
		const Box B = clipc(s->B, e->f_lo, e->f_hi);
		if (!bcount(B)) continue;
		nact++;
		if (!s->hPha.empty()) nvals += s->cnt;
	}
	if (nact > ALISB_MAXSRCOPS) e->cur_src_fused[set] = false;
	if (!e->cur_src_fused[set] || nact == 0) return ALISB_OK;
	float *hv = 0, *dv = 0;
	if (nvals) {
		if (nvals > e->srcv_cap) {
			CU(cudaStreamSynchronize(e->stream));
			e->srcv_cap = 0;                                 /* until every slot has been re-allocated */
			for (int r=0; r<SRC_RING; r++) {
				if (e->srcv_h[r]) cudaFreeHost(e->srcv_h[r]);
				cudaFree(e->srcv_d[r]);
				e->srcv_h[r] = 0; e->srcv_d[r] = 0;
				CU(cudaMallocHost((void**)&e->srcv_h[r], nvals*sizeof(float)));
				TRY(dalloc(&e->srcv_d[r], nvals));
				if (!e->srcv_ev[r]) CU(cudaEventCreateWithFlags(&e->srcv_ev[r], cudaEventDisableTiming));
			}
			e->srcv_cap = nvals;
		}
		const int slot = e->srcv_slot; e->srcv_slot = (e->srcv_slot + 1) % SRC_RING;
		CU(cudaEventSynchronize(e->srcv_ev[slot]));
		hv = e->srcv_h[slot]; dv = e->srcv_d[slot];
		e->cur_src_slot[set] = slot;
	}
	size_t off = 0;
	for (SrcObj *s : e->src[set]) {
		if (!(e->step < s->d.n_active) || !s->cnt) continue;
		const Box B = clipc(s->B, e->f_lo, e->f_hi);
		if (!bcount(B)) continue;
		SrcOp &o = L.op[L.n++];
		o.F = e->F[set][s->d.m][s->d.comp]; o.B = B; o.full = s->B; o.vals = 0; o.amp = s->dAmp; o.amplitude = s->d.amplitude; o.f = 0.f;
		if (!s->hPha.empty()) {
			for (size_t q=0; q<s->cnt; q++)
				hv[off + q] = (s->hAmp.empty() ? s->d.amplitude : s->hAmp[q]) * s->d.waveform(s->d.user, t, s->hPha[q]);
			o.vals = dv + off; off += s->cnt;
		} else o.f = s->d.waveform(s->d.user, t, s->d.phase);
	}
	if (nvals) CU(cudaMemcpyAsync(dv, hv, nvals*sizeof(float), cudaMemcpyHostToDevice, e->stream));
	return ALISB_OK;
}

static int launch_sources(alisb_engine *e, int set, float t)
{
	if (e->src[set].empty()) return ALISB_OK;
	if (e->cur_src_fused[set]) {
		/* launched even when no source is active right now, so that the pass has the same kernels
		 * at every step (graph replay only rewrites this node's argument) */
		source_list<<<ALISB_LIST_CTAS, 1024, 0, e->stream>>>(e->cur_src[set]);
		e->launches++;
		if (g_mode == GMODE_CAPTURE) {
			cudaStreamCaptureStatus st; const cudaGraphNode_t *deps = 0; size_t nd = 0;
			CU(cudaStreamGetCaptureInfo(e->stream, &st, 0, 0, &deps, &nd));
			if (nd != 1 || !g_nodes) return fail(ALISB_ECUDA, "graph capture: cannot identify the source node");
			g_nodes->push_back({e, set, deps[0]});
		}
		CU(cudaGetLastError());
		return ALISB_OK;
	}
	for (SrcObj *s : e->src[set]) {
		if (!(e->step < s->d.n_active) || !s->cnt) continue;
		const Box B = clipc(s->B, e->f_lo, e->f_hi);
		const size_t nb = bcount(B);
		if (!nb) continue;
		float *F = e->F[set][s->d.m][s->d.comp];
		if (!s->hPha.empty()) {
			const int r = s->slot; s->slot = (s->slot + 1) % SRC_RING;
			CU(cudaEventSynchronize(s->ev[r]));
			float *h = s->hvals[r];
			for (size_t q=0; q<s->cnt; q++)
				h[q] = (s->hAmp.empty() ? s->d.amplitude : s->hAmp[q]) * s->d.waveform(s->d.user, t, s->hPha[q]);
			CU(cudaMemcpyAsync(s->dvals[r], h, s->cnt*sizeof(float), cudaMemcpyHostToDevice, e->stream));
			source_add<<<div_up(nb, 256), 256, 0, e->stream>>>(F, B, s->B, e->ny, e->pitch, s->dvals[r], 0, 0.f, 0.f);
			CU(cudaEventRecord(s->ev[r], e->stream));
		} else {
			const float f = s->d.waveform(s->d.user, t, s->d.phase);
			source_add<<<div_up(nb, 256), 256, 0, e->stream>>>(F, B, s->B, e->ny, e->pitch, 0, s->dAmp, s->d.amplitude, f);
		}
		e->launches++;
	}
	CU(cudaGetLastError());
	return ALISB_OK;
}

static int finish_sources(alisb_engine *e, int set)
{
	if (e->cur_src_slot[set] >= 0) CU(cudaEventRecord(e->srcv_ev[e->cur_src_slot[set]], e->stream));
	e->cur_src_slot[set] = -1;
	return ALISB_OK;
}

static int pass_launch(alisb_engine *e, int set, int part);

/* Orders the ops of a list kernel into chains (kernels.cuh, CopyList): stable partition by target array, then, inside
 * every chain, a barrier after op q-1 whenever op q conflicts with an op since the chain's last barrier.  Returns the
 * largest op (cells) so that the caller can choose between one CTA and one cluster per chain. */
template <typename Op, typename List, typename Conflict>
static size_t build_chains(List &L, Op *ops, Conflict conflict)
{
	Op tmp[sizeof(L.op)/sizeof(L.op[0])];
	float *key[ALISB_MAXCHAINS];
	int nkey = 0, chain_of[sizeof(L.op)/sizeof(L.op[0])];
	for (int q=0; q<L.n; q++) {
		int c = 0;
		while (c < nkey && key[c] != ops[q].F) c++;
		if (c == nkey) { if (nkey < ALISB_MAXCHAINS) key[nkey++] = ops[q].F; else c = nkey - 1; }   /* more arrays than chains: the last chain takes the rest, in order */
		chain_of[q] = c;
	}
	int n = 0;
	size_t biggest = 0;
	for (int c=0; c<nkey; c++) {
		L.first[c] = n;
		int level_start = n;
		for (int q=0; q<L.n; q++) if (chain_of[q] == c) {
			tmp[n] = ops[q]; tmp[n].sync_after = 0;
			bool hit = false;
			for (int r=level_start; r<n && !hit; r++) hit = conflict(tmp[r], tmp[n]);
			if (hit) { tmp[n-1].sync_after = 1; level_start = n; }
			biggest = std::max(biggest, (size_t)tmp[n].B.na*tmp[n].B.nb*tmp[n].B.nc);
			n++;
		}
	}
	L.first[nkey] = n; L.nchain = nkey;
	for (int c=nkey+1; c<=ALISB_MAXCHAINS; c++) L.first[c] = n;
	memcpy(ops, tmp, sizeof(Op)*n);
	return biggest;
}
static bool boxes_meet(const Box &A, const Box &B)
{
	return A.a0 < B.a0 + B.na && B.a0 < A.a0 + A.na && A.b0 < B.b0 + B.nb && B.b0 < A.b0 + A.nb && A.c0 < B.c0 + B.nc && B.c0 < A.c0 + A.nc;
}

/* the strips of one TFSF region for this pass (clipped to the planes of the current part) as a list for the list kernels;
 * false: the faces are large, one launch per strip */
static bool small_strip_list(alisb_engine *e, const TfsfObj *t, int set, StripList &SL)
{
	const alisb_engine *x = t->d.aux;
	const std::vector<alisb_strip> &L = set == ALISB_SET_E ? t->sE : t->sH;
	memset(&SL, 0, sizeof SL);             /* the list is also the key of its regrouped plan */
	SL.n = 0; SL.ny = e->ny; SL.pitch = e->pitch; SL.any = x->ny; SL.apitch = x->pitch;
	for (int m=0; m<=e->g.complex_field; m++) for (const alisb_strip &s : L) {
		Box B = clipc(mkbox(s.box), e->f_lo, e->f_hi);
		const size_t n = bcount(B);
		if (!n) continue;
		if (n > 16384 || SL.n == ALISB_MAXSTRIPOPS) return false;
		StripOp &o = SL.op[SL.n++];
		o.F = e->F[set][m][s.comp]; o.G = x->F[1-set][m][s.aux_comp]; o.C = s.use_eps ? e->eps[s.comp] : 0; o.coef = s.coef; o.B = B;
		o.alt0 = s.alt[0] + t->d.offset[0]; o.alt1 = s.alt[1] + t->d.offset[1]; o.alt2 = s.alt[2] + t->d.offset[2];
		o.clampa = t->d.clamp[1-set][s.aux_comp][0]; o.clampb = t->d.clamp[1-set][s.aux_comp][1];
	}
	return true;
}

static StripPlan *build_strip_plan(alisb_engine *e, const alisb_engine *x, const StripList &SL)
{
	if ((e->nz + 2)*e->plane >= ((size_t)1 << 31) || (x->nz + 2)*x->plane >= ((size_t)1 << 31)) return 0;
	RsOp ops[ALISB_MAXSTRIPOPS];
	for (int q=0; q<SL.n; q++) {
		const StripOp &o = SL.op[q];
		ops[q].F = o.F; ops[q].G = o.G; ops[q].C = o.C; ops[q].coef = o.coef;
		ops[q].a0 = o.B.a0; ops[q].b0 = o.B.b0; ops[q].c0 = o.B.c0; ops[q].na = o.B.na; ops[q].nb = o.B.nb; ops[q].nc = o.B.nc;
		ops[q].alt0 = o.alt0; ops[q].alt1 = o.alt1; ops[q].alt2 = o.alt2; ops[q].clampa = o.clampa; ops[q].clampb = o.clampb;
	}
	RsPlan R;
	if (!rs_resolve(ops, SL.n, SL.ny, SL.pitch, SL.any, SL.apitch, ALISB_XO, ALISB_RS_ARRAYS, 15, R)) return 0;
	std::vector<RsCellDev> cells(R.cells.size());
	std::vector<RsTermDev> terms;
	for (size_t q=0; q<R.cells.size(); q++) {
		const RsCell &c = R.cells[q];
		const RsTerm &t0 = R.terms[c.first];
		RsCellDev d;
		d.tgt = c.tgt; d.src = t0.src; d.coef = t0.coef; d.pad = 0;
		d.meta = c.farr | (t0.garr << 4) | ((t0.carr + 1) << 8) | (c.count << 12);
		d.more = (int)terms.size();
		for (int k=1; k<c.count; k++) {
			const RsTerm &t = R.terms[c.first + k];
			RsTermDev td; td.src = t.src; td.coef = t.coef; td.meta = t.garr | ((t.carr + 1) << 4); td.pad = 0;
			terms.push_back(td);
		}
		cells[q] = d;
	}
	StripPlan *sp = new StripPlan();
	memset(&sp->P, 0, sizeof sp->P);
	sp->key = SL; sp->d_cells = 0; sp->d_terms = 0;
	for (size_t q=0; q<R.F.size(); q++) sp->P.F[q] = (float*)R.F[q];
	for (size_t q=0; q<R.G.size(); q++) sp->P.G[q] = (const float*)R.G[q];
	for (size_t q=0; q<R.C.size(); q++) sp->P.C[q] = (const float*)R.C[q];
	sp->P.n = (int)cells.size();
	if (sp->P.n) {
		bool ok = cudaMalloc((void**)&sp->d_cells, sizeof(RsCellDev)*cells.size()) == cudaSuccess
		       && cudaMemcpy(sp->d_cells, cells.data(), sizeof(RsCellDev)*cells.size(), cudaMemcpyHostToDevice) == cudaSuccess;
		if (ok && !terms.empty()) ok = cudaMalloc((void**)&sp->d_terms, sizeof(RsTermDev)*terms.size()) == cudaSuccess
		                            && cudaMemcpy(sp->d_terms, terms.data(), sizeof(RsTermDev)*terms.size(), cudaMemcpyHostToDevice) == cudaSuccess;
		if (!ok) { cudaGetLastError(); cudaFree(sp->d_cells); cudaFree(sp->d_terms); delete sp; return 0; }
	}
	sp->P.cells = sp->d_cells; sp->P.terms = sp->d_terms;
	return sp;
}

static StripPlan *strip_plan(alisb_engine *e, const alisb_engine *x, const StripList &SL, bool may_build)
{
	if (!e->tune_resolve || e->dbuf || x->dbuf || SL.n < 2) return 0;
	for (StripPlan *sp : e->splan) if (!memcmp(&sp->key, &SL, sizeof SL)) return sp->P.n >= 0 ? sp : 0;
	if (!may_build) return 0;
	if (e->splan.size() >= 16) {
		cudaStreamSynchronize(e->stream);
		for (StripPlan *sp : e->splan) { cudaFree(sp->d_cells); cudaFree(sp->d_terms); delete sp; }
		e->splan.clear(); e->cfg_version++;
	}
	StripPlan *sp = build_strip_plan(e, x, SL);
	if (!sp) { sp = new StripPlan(); sp->key = SL; sp->d_cells = 0; sp->d_terms = 0; memset(&sp->P, 0, sizeof sp->P); sp->P.n = -1; }
	e->splan.push_back(sp);
	return sp->P.n >= 0 ? sp : 0;
}

static int apply_tfsf(alisb_engine *e, int set, bool step_aux)
{
	for (TfsfObj *t : e->tf) {
		alisb_engine *x = t->d.aux;
		if (step_aux) TRY(pass_launch(x, set, ALISB_PART_ALL));
		const std::vector<alisb_strip> &L = set == ALISB_SET_E ? t->sE : t->sH;
		/* small faces: all strips of this TFSF region in one launch */
		{
			StripList SL;
			const bool ok = small_strip_list(e, t, set, SL);
			if (ok) {
				if (StripPlan *sp = strip_plan(e, x, SL, g_mode != GMODE_CAPTURE)) {
					if (sp->P.n) { strip_resolved<<<div_up(sp->P.n, 256), 256, 0, e->stream>>>(sp->P); e->launches++; }
					continue;
				}
				if (SL.n) {
					/* strips that add to the same array keep their order where their boxes meet (float addition is not associative);
					 * the chains of the three components (and of the complex sets) run side by side */
					const size_t biggest = build_chains(SL, SL.op, [](const StripOp &p, const StripOp &q) { return p.F == q.F && boxes_meet(p.B, q.B); });
					if (biggest <= 1024) strip_list1<<<SL.nchain, 1024, 0, e->stream>>>(SL);        /* one cell per thread: a CTA per chain is enough */
					else                 strip_list<<<SL.nchain*ALISB_LIST_CTAS, 1024, 0, e->stream>>>(SL);
					e->launches++;
				}
				continue;
			}
		}
		for (int m=0; m<=e->g.complex_field; m++) for (const alisb_strip &s : L) {
			Box B = clipc(mkbox(s.box), e->f_lo, e->f_hi);
			const size_t n = bcount(B);
			if (!n) continue;
			const float *C = s.use_eps ? e->eps[s.comp] : 0;
			tfsf_strip<<<div_up(n, 128), 128, 0, e->stream>>>(e->F[set][m][s.comp], x->F[1-set][m][s.aux_comp], C, s.coef, B, e->ny, e->pitch,
				x->ny, x->pitch, s.alt[0] + t->d.offset[0], s.alt[1] + t->d.offset[1], s.alt[2] + t->d.offset[2],
				t->d.clamp[1-set][s.aux_comp][0], t->d.clamp[1-set][s.aux_comp][1]);
			e->launches++;
		}
	}
	CU(cudaGetLastError());
	return ALISB_OK;
}

/* the copies of a pass (clipped to the planes of the current part) as a list for the list kernels; false: the planes
 * are large, one launch per copy */
static bool small_copy_list(alisb_engine *e, int set, CopyList &L)
{
	memset(&L, 0, sizeof L);             /* the list is also the key of its resolved plan: no indeterminate padding */
	L.n = 0; L.ny = e->ny; L.pitch = e->pitch;
	if (e->copies[set].size() > ALISB_MAXCOPIES) return false;
	size_t total = 0;
	for (const alisb_copy &k : e->copies[set]) {
		Box B = clipc(mkbox(k.box), e->f_lo, e->f_hi);
		const size_t n = bcount(B);
		if (!n) continue;
		if (n > 32768) return false;
		total += n;
		CopyOp &o = L.op[L.n++];
		o.F = e->F[k.set][0][k.comp]; o.G = e->g.complex_field ? e->F[k.set][1][k.comp] : 0;
		o.cosk = k.cosk; o.sink = k.sink; o.B = B; o.alt0 = k.alt[0]; o.alt1 = k.alt[1]; o.alt2 = k.alt[2];
		o.sync_after = 0;
	}
	return total <= 400000;
}

/* The list resolved by copy_resolve.h into the tables of copy_resolved.  Returns 0 when the list does not fit them (the
 * list kernels take it then). */
static CopyPlan *build_copy_plan(alisb_engine *e, const CopyList &L)
{
	if ((e->nz + 2)*e->plane >= ((size_t)1 << 31)) return 0;
	RcOp ops[ALISB_MAXCOPIES];
	for (int q=0; q<L.n; q++) {
		const CopyOp &o = L.op[q];
		ops[q].F = o.F; ops[q].G = o.G; ops[q].a0 = o.B.a0; ops[q].b0 = o.B.b0; ops[q].c0 = o.B.c0; ops[q].na = o.B.na; ops[q].nb = o.B.nb; ops[q].nc = o.B.nc;
		ops[q].alt0 = o.alt0; ops[q].alt1 = o.alt1; ops[q].alt2 = o.alt2;
	}
	RcPlan R;
	if (!rc_resolve(ops, L.n, L.ny, L.pitch, ALISB_XO, ALISB_RC_ARRAYS, ALISB_RC_CHAINS, ALISB_RC_DEPTH, R)) return 0;
	std::vector<int4> ent(R.ent.size());
	for (size_t q=0; q<R.ent.size(); q++) ent[q] = make_int4(R.ent[q].tgt, R.ent[q].src, (R.ent[q].arr << 8) | R.ent[q].chain, 0);
	CopyPlan *cp = new CopyPlan();
	memset(&cp->P, 0, sizeof cp->P);
	cp->key = L; cp->d_ent = 0; cp->d_tmp = 0;
	for (size_t arr=0; arr<R.F.size(); arr++) { cp->P.F[arr] = (float*)R.F[arr]; cp->P.G[arr] = (float*)R.G[arr]; }
	for (size_t ch=0; ch<R.chains.size(); ch++) {
		cp->P.depth[ch] = (int)R.chains[ch].size();
		for (size_t d=0; d<R.chains[ch].size(); d++) { cp->P.cosk[ch][d] = L.op[R.chains[ch][d]].cosk; cp->P.sink[ch][d] = L.op[R.chains[ch][d]].sink; }
	}
	cp->P.n = (int)ent.size(); cp->P.nbound = (int)R.nbound;
	if (cp->P.n) {
		if (cudaMalloc((void**)&cp->d_ent, sizeof(int4)*ent.size()) != cudaSuccess || cudaMalloc((void**)&cp->d_tmp, sizeof(float2)*ent.size()) != cudaSuccess
		    || cudaMemcpy(cp->d_ent, ent.data(), sizeof(int4)*ent.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
			cudaGetLastError(); cudaFree(cp->d_ent); cudaFree(cp->d_tmp); delete cp; return 0;
		}
	}
	cp->P.ent = cp->d_ent; cp->P.tmp = cp->d_tmp;
	return cp;
}

/* the resolved plan of a list: looked up by the list itself; built on first sight unless the stream is being captured
 * (pass_prepare has then built it already) */
static CopyPlan *copy_plan(alisb_engine *e, const CopyList &L, bool may_build)
{
	if (!e->tune_resolve || e->dbuf || L.n < 2) return 0;          /* double-buffered passes alternate their arrays: the list kernels take those */
	for (CopyPlan *cp : e->cplan) if (!memcmp(&cp->key, &L, sizeof L)) return cp->P.n >= 0 ? cp : 0;
	if (!may_build) return 0;
	if (e->cplan.size() >= 16) {                                   /* the configuration keeps changing: start over (captured passes hold plan pointers) */
		cudaStreamSynchronize(e->stream);
		for (CopyPlan *cp : e->cplan) { cudaFree(cp->d_ent); cudaFree(cp->d_tmp); delete cp; }
		e->cplan.clear(); e->cfg_version++;
	}
	CopyPlan *cp = build_copy_plan(e, L);
	if (!cp) { cp = new CopyPlan(); cp->key = L; cp->d_ent = 0; cp->d_tmp = 0; memset(&cp->P, 0, sizeof cp->P); cp->P.n = -1; }   /* remembered as "does not resolve" */
	e->cplan.push_back(cp);
	return cp->P.n >= 0 ? cp : 0;
}

static int apply_copies(alisb_engine *e, int set)
{
	/* small planes: the whole ordered list in one launch */
	CopyList L;
	const bool small = small_copy_list(e, set, L);
	if (small) {
		if (CopyPlan *cp = copy_plan(e, L, g_mode != GMODE_CAPTURE)) {
			if (cp->P.n) {
				/* one cluster for the bound entries (if any) and one per 16 K free entries (two per thread) */
				const int nfree = cp->P.n - cp->P.nbound;
				const int clusters = (cp->P.nbound ? 1 : 0) + std::min(16, div_up(nfree, 2*1024*ALISB_LIST_CTAS));
				copy_resolved<<<ALISB_LIST_CTAS*clusters, 1024, 0, e->stream>>>(cp->P);
				e->launches++;
			}
			CU(cudaGetLastError());
			return ALISB_OK;
		}
		/* a barrier is needed between op q and the ops of its chain since the last barrier only if they conflict:
		 * (target x target, target x source or source x target boxes intersect) */
		auto shifted = [](const CopyOp &o) { Box S = o.B; S.a0 += o.alt0; S.b0 += o.alt1; S.c0 += o.alt2; return S; };
		const size_t biggest = build_chains(L, L.op, [&](const CopyOp &p, const CopyOp &q) {
			return p.F == q.F && (boxes_meet(p.B, q.B) || boxes_meet(p.B, shifted(q)) || boxes_meet(shifted(p), q.B)); });
		if (L.n) {
			if (biggest <= 1024) copy_list1<<<L.nchain, 1024, 0, e->stream>>>(L);        /* one cell per thread: a CTA per chain is enough */
			else                 copy_list<<<L.nchain*ALISB_LIST_CTAS, 1024, 0, e->stream>>>(L);
			e->launches++;
		}
		CU(cudaGetLastError());
		return ALISB_OK;
	}
	/* large planes: one launch per copy.  A copy reads and writes one component (and its imaginary twin), so the
	 * ordered list falls into up to three independent chains, one per component: they run side by side (the
	 * launches are latency-bound: a 2048 x 256 x-plane is 0.5 M scattered 4-byte accesses) */
	float *chain_of[3] = {0, 0, 0};
	int nchain = 0;
	const bool conc = e->tune_conc > 0;
	bool forked = false, used[3] = {false, false, false};
	for (const alisb_copy &k : e->copies[set]) {
		Box B = clipc(mkbox(k.box), e->f_lo, e->f_hi);
		const size_t n = bcount(B);
		if (!n) continue;
		float *F = e->F[k.set][0][k.comp], *G = e->g.complex_field ? e->F[k.set][1][k.comp] : 0;
		int q = 0;
		while (q < nchain && chain_of[q] != F) q++;
		if (q == nchain && nchain < 3) chain_of[nchain++] = F;
		cudaStream_t st = e->stream;
		if (conc && q > 0 && q < 3) {                     /* chain 0 stays on the engine's stream */
			if (!forked) { CU(cudaEventRecord(e->ev_fork, e->stream)); forked = true; }
			if (!used[q]) { CU(cudaStreamWaitEvent(e->side[q], e->ev_fork, 0)); used[q] = true; }
			st = e->side[q];
		}
		copy_plane<<<div_up(n, 256), 256, 0, st>>>(F, G, k.cosk, k.sink, B, e->ny, e->pitch, k.alt[0], k.alt[1], k.alt[2]);
		e->launches++;
	}
	for (int q=1; q<3; q++) if (used[q]) { CU(cudaEventRecord(e->ev_join[q], e->side[q])); CU(cudaStreamWaitEvent(e->stream, e->ev_join[q], 0)); }
	CU(cudaGetLastError());
	return ALISB_OK;
}

static void set_filter(alisb_engine *e, int set, int part)
{
	const int send = set == ALISB_SET_E ? e->nz : 1;          /* the plane the z-neighbour needs */
	e->f_lo = 0; e->f_hi = e->nz + 1;
	if (part == ALISB_PART_SEND_PLANE) e->f_lo = e->f_hi = send;
	if (part == ALISB_PART_REST) { if (set == ALISB_SET_E) e->f_hi = send - 1; else e->f_lo = send + 1; }
}

/* host side of a pass (recursively for the TFSF auxiliary worlds that the pass steps first) */
static int pass_prepare(alisb_engine *e, int set, int part)
{
	CU(cudaSetDevice(e->g.device));
	if (part != ALISB_PART_REST) for (TfsfObj *t : e->tf) {
		alisb_engine *x = t->d.aux;
		if (x->stream != e->stream && !x->own_stream) return fail(ALISB_ESTATE, "aux engine on a foreign stream");
		if (x->stream != e->stream) { CU(cudaStreamSynchronize(x->stream)); if (x->own_stream) { cudaStreamDestroy(x->stream); x->own_stream = false; } x->stream = e->stream; x->cfg_version++; }
		TRY(pass_prepare(x, set, ALISB_PART_ALL));
	}
	open_pass(e, set);
	set_filter(e, set, part);
	if (ade_in_kernel(e)) TRY(ade_table(e));
	TRY(prepare_sources(e, set, set == ALISB_SET_E ? e->tH : e->tE));
	if (e->tune_resolve && !e->dbuf) {              /* the pass may be captured: plans are built here */
		CopyList L;
		if (small_copy_list(e, set, L)) copy_plan(e, L, true);
		for (TfsfObj *t : e->tf) { StripList SL; if (small_strip_list(e, t, set, SL)) strip_plan(e, t->d.aux, SL, true); }
	}
	return ALISB_OK;
}

/* all kernels of a pass, in the reference's order (update.c:275-341) */
static int pass_launch(alisb_engine *e, int set, int part)
{
	set_filter(e, set, part);
	/* launch-bound grids with complex fields: the real and the imaginary set are independent within a pass and their
	 * (single) volume launches are latency bound -- they run side by side, the imaginary set on a side stream */
	const bool pair = e->g.complex_field && e->tune_conc > 0 && (long)e->nx*e->ny*e->nz < e->tune_small;
	cudaStream_t const main_stream = e->stream;
	/* launch-bound grids with TFSF regions: the auxiliary worlds' passes (volume, source, copies: a chain of latency-bound
	 * launches) run on a side stream beside this world's whole pass.  Nothing here waits for them: the strips of an E pass
	 * read the auxiliary world's H, which its E pass (update.c:207 before update.c:209-234) does not touch, and vice versa
	 * (update.c:242-269); the pass as a whole ends when both have finished */
	const bool aux_aside = part != ALISB_PART_REST && !e->tf.empty() && e->tune_conc > 0 && e->tune_aux_aside && (long)e->nx*e->ny*e->nz < e->tune_small;
	if (aux_aside) {
		/* on the stream of the highest priority: this world's volume launch fills every SM (two CTAs take the whole register
		 * file), so the few CTAs of the auxiliary chain must be first in line whenever a slot frees up */
		CU(cudaEventRecord(e->ev_fork, main_stream)); CU(cudaStreamWaitEvent(e->hi, e->ev_fork, 0));
		int rca = ALISB_OK;
		for (TfsfObj *t : e->tf) {
			alisb_engine *x = t->d.aux;
			cudaStream_t const xs = x->stream;
			x->stream = e->hi;
			if (rca == ALISB_OK) rca = pass_launch(x, set, ALISB_PART_ALL);
			x->stream = xs;
		}
		CU(cudaEventRecord(e->ev_hi, e->hi));
		TRY(rca);
	}
	if (pair) { CU(cudaEventRecord(e->ev_fork, main_stream)); CU(cudaStreamWaitEvent(e->side[0], e->ev_fork, 0)); }
	int rc = ALISB_OK;
	for (int m=0; m<=e->g.complex_field && rc == ALISB_OK; m++) {
		if (pair && m == 1) e->stream = e->side[0];
		if (set == ALISB_SET_E) {
			if (!e->obj.empty() && !ade_in_kernel(e)) rc = ade_pre(e, m);
			if (rc == ALISB_OK) rc = launch_volume(e, ALISB_SET_E, m);
		} else {
			rc = launch_volume(e, ALISB_SET_H, m);
			if (rc == ALISB_OK && !e->obj.empty() && !ade_in_kernel(e)) rc = ade_post(e, m);
		}
		e->stream = main_stream;
	}
	if (pair) { CU(cudaEventRecord(e->ev_join[0], e->side[0])); CU(cudaStreamWaitEvent(main_stream, e->ev_join[0], 0)); }
	TRY(rc);
	TRY(apply_tfsf(e, set, part != ALISB_PART_REST && !aux_aside));
	set_filter(e, set, part);
	TRY(launch_sources(e, set, set == ALISB_SET_E ? e->tH : e->tE));
	TRY(apply_copies(e, set));
	if (aux_aside) CU(cudaStreamWaitEvent(main_stream, e->ev_hi, 0));
	e->f_lo = 0; e->f_hi = e->nz + 1;
	return ALISB_OK;
}

/* bookkeeping after the pass has been issued (times: update.c:298,340) */
static void pass_time(alisb_engine *e, int set, int part)
{
	if (part == ALISB_PART_SEND_PLANE) return;
	if (set == ALISB_SET_E) e->t = e->tE = ++e->step * e->g.dt;
	else e->t = e->tH = e->tE + 0.5 * e->g.dt;
	e->pass_open[set] = false;
}

static int pass_finish(alisb_engine *e, int set, int part, bool time_done = false)
{
	if (part != ALISB_PART_REST) for (TfsfObj *t : e->tf) TRY(pass_finish(t->d.aux, set, ALISB_PART_ALL));
	TRY(finish_sources(e, set));
	if (!time_done) pass_time(e, set, part);
	return ALISB_OK;
}

static uint64_t config_version(alisb_engine *e)
{
	uint64_t v = e->cfg_version*1000003ull;
	for (TfsfObj *t : e->tf) v = v*31 + config_version(t->d.aux);
	return v;
}

static bool graphable(alisb_engine *e, int set)
{
	if (!e->graph_ok || e->profiling || e->dbuf || e->fuse_ok) return false;
	if (!e->src[set].empty() && !e->cur_src_fused[set]) return false;
	for (TfsfObj *t : e->tf) if (!graphable(t->d.aux, set)) return false;
	return true;
}

/* issue a deferred E pass: as the fused variant (the caller is about to run the H pass and nothing
 * touched the fields in between) or as the plain two-pass variant */
static int launch_pending(alisb_engine *e, bool fused)
{
	if (e->pendingE < 0) return ALISB_OK;
	const int part = e->pendingE;
	e->pendingE = -1;
	TRY(halo_fence(e, ALISB_SET_E, part));                 /* the fused kernel also forms H', but of no plane that travels and from no ghost plane */
	e->fusing = fused && e->fuse_ok;
	const int r = pass_launch(e, ALISB_SET_E, part);
	const bool did = e->fusing;
	e->fusing = false;
	if (r != ALISB_OK) return r;
	e->h_fused = did;
	return pass_finish(e, ALISB_SET_E, part, true);
}

extern "C" uint64_t alisb_fused_launch_count(alisb_engine *e) { return e->fused_launches; }

extern "C" int alisb_set_fusion(alisb_engine *e, int on)
{
	UNFUSE(e);
	e->fuse_req = on ? 1 : 0; e->cfg_version++;
	return ALISB_OK;
}

extern "C" int alisb_flush(alisb_engine *e, int h_pass_follows)
{
	DEV(e);
	return launch_pending(e, h_pass_follows != 0);
}

extern "C" int alisb_begin_pass(alisb_engine *e, int set)
{
	if (set < 0 || set > 1) return fail(ALISB_EINVAL, "bad field set");
	DEV(e);
	if (set == ALISB_SET_E) UNFUSE(e);
	fuse_eval(e);
	if (e->fuse_ok && !e->tune_wave) TRY(enable_dbuf(e));
	open_pass(e, set);
	return ALISB_OK;
}

static void trace_print(const alisb_engine *e)
{
	static const char *const what[5] = { "prepare (host)", "graph capture", "source-node patching", "graph launch", "direct launches" };
	for (int st=0; st<2; st++) for (int q=0; q<5; q++) if (e->trace_n[st][q])
		fprintf(stderr, "[alisb] %dx%dx%d pass %c %-22s %8ld calls %10.4f s (%.1f us each)\n", e->nx, e->ny, e->nz, st ? 'H' : 'E', what[q], e->trace_n[st][q], e->trace_s[st][q], 1e6*e->trace_s[st][q]/e->trace_n[st][q]);
}

extern "C" int alisb_step_part(alisb_engine *e, int set, int part)
{
	if (set < 0 || set > 1 || part < 0 || part > 2) return fail(ALISB_EINVAL, "bad step part");
	static const char *const names[2][3] = { { "alisb:updateE", "alisb:updateE:send-plane", "alisb:updateE:rest" }, { "alisb:updateH", "alisb:updateH:send-plane", "alisb:updateH:rest" } };
	NvtxRange range(names[set][part]);
	if (e->fuse_req || e->dbuf) {
		CU(cudaSetDevice(e->g.device));
		if (set == ALISB_SET_E) {
			if (e->pendingE >= 0 || (e->h_fused && part != ALISB_PART_REST)) UNFUSE(e);
			fuse_eval(e);
			if (e->fuse_ok && !e->tune_wave) TRY(enable_dbuf(e));
			bool list_sources = true;
			for (SrcObj *so : e->src[ALISB_SET_E]) if (so->cnt > 16384) list_sources = false;
			if (e->fuse_ok && part != ALISB_PART_SEND_PLANE && list_sources && (int)e->src[ALISB_SET_E].size() <= ALISB_MAXSRCOPS) {
				/* host side and bookkeeping now, kernels when the H pass (or anything else) arrives */
				TRY(pass_prepare(e, set, part));
				pass_time(e, set, part);
				e->pendingE = part;
				return ALISB_OK;
			}
		} else {
			TRY(launch_pending(e, true));
		}
	}
	const bool trace = e->tune_debug != 0;
	auto tick = [&]() { return trace ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point(); };
	auto tock = [&](int what, std::chrono::steady_clock::time_point t0) { if (trace) { e->trace_s[set][what] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); e->trace_n[set][what]++; } };
	auto t0 = tick();
	TRY(pass_prepare(e, set, part));
	tock(0, t0);
	TRY(halo_fence(e, set, part));
	if (graphable(e, set)) {
		PassGraph &G = e->graph[set][part];
		const uint64_t ver = config_version(e);
		if (!G.exec || G.version != ver) {
			/* capture the pass once */
			t0 = tick();
			if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = 0; }
			if (G.graph) { cudaGraphDestroy(G.graph); G.graph = 0; }
			G.nodes.clear();
			const uint64_t l0 = e->launches;
			CU(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
			g_mode = GMODE_CAPTURE; g_nodes = &G.nodes;
			const int r = pass_launch(e, set, part);
			g_mode = GMODE_NONE; g_nodes = 0;
			cudaError_t ce = cudaStreamEndCapture(e->stream, &G.graph);
			if (r != ALISB_OK || ce != cudaSuccess || !G.graph) {
				if (G.graph) { cudaGraphDestroy(G.graph); G.graph = 0; }
				cudaGetLastError();
				e->graph_ok = false;                      /* fall back to plain launches for good */
				TRY(pass_launch(e, set, part));
				return pass_finish(e, set, part);
			}
			CU(cudaGraphInstantiate(&G.exec, G.graph, 0));
			G.version = ver;
			G.nkernels = (int)(e->launches - l0);
			e->launches = l0;
			tock(1, t0);
		} else {
			t0 = tick();
			for (const SrcNode &sn : G.nodes) {
				cudaKernelNodeParams kp;
				memset(&kp, 0, sizeof kp);
				void *args[1] = { (void*)&sn.eng->cur_src[sn.set] };
				kp.func = (void*)source_list; kp.gridDim = dim3(ALISB_LIST_CTAS); kp.blockDim = dim3(1024); kp.kernelParams = args;
				CU(cudaGraphExecKernelNodeSetParams(G.exec, sn.node, &kp));
			}
			tock(2, t0);
		}
		t0 = tick();
		CU(cudaGraphLaunch(G.exec, e->stream));
		tock(3, t0);
		if (trace && e->tune_debug >= 2 && set == ALISB_SET_H && e->trace_n[set][3] % 2000 == 0) trace_print(e);
		e->launches += G.nkernels;
		return pass_finish(e, set, part);
	}
	t0 = tick();
	TRY(pass_launch(e, set, part));
	tock(4, t0);
	if (set == ALISB_SET_H && part != ALISB_PART_SEND_PLANE) e->h_fused = false;      /* the H pass is complete */
	return pass_finish(e, set, part);
}

extern "C" int alisb_step_E(alisb_engine *e) { return alisb_step_part(e, ALISB_SET_E, ALISB_PART_ALL); }
extern "C" int alisb_step_H(alisb_engine *e) { return alisb_step_part(e, ALISB_SET_H, ALISB_PART_ALL); }

extern "C" int alisb_run(alisb_engine *e, int steps)
{
	e->ev_used = 0; e->prof.clear();
	for (int s=0; s<steps; s++) { TRY(alisb_step_E(e)); TRY(alisb_step_H(e)); }
	return ALISB_OK;
}

extern "C" int alisb_synchronize(alisb_engine *e)
{
	DEV(e); SETTLE(e);
	CU(cudaStreamSynchronize(e->stream));
	if (e->d_wave_error) {
		int bad = 0;
		CU(cudaMemcpy(&bad, e->d_wave_error, sizeof bad, cudaMemcpyDeviceToHost));
		if (bad) return fail(ALISB_ECUDA, "yee_wave: a tile dependency never arrived (fields are invalid)");
	}
	return ALISB_OK;
}
extern "C" int alisb_get_time(alisb_engine *e, uint64_t *n, float *t, float *tE, float *tH) { if (n) *n = e->step; if (t) *t = e->t; if (tE) *tE = e->tE; if (tH) *tH = e->tH; return ALISB_OK; }
extern "C" int alisb_set_time(alisb_engine *e, uint64_t n, float t, float tE, float tH) { SETTLE(e); e->step = n; e->t = t; e->tE = tE; e->tH = tH; return ALISB_OK; }

extern "C" int alisb_reset_fields(alisb_engine *e)
{
	DEV(e);
	UNFUSE(e);
	for (int s=0; s<2; s++) for (int m=0; m<=e->g.complex_field; m++) {
		for (int c=0; c<3; c++) CU(cudaMemsetAsync(e->F[s][m][c], 0, e->elems*sizeof(float), e->stream));
		if (e->dbuf) for (int c=0; c<3; c++) CU(cudaMemsetAsync(e->Fb[s][m][c], 0, e->elems*sizeof(float), e->stream));
		for (int ax=0; ax<3; ax++) for (int sd=0; sd<2; sd++) for (int w=0; w<2; w++)
			if (e->psi[s][m][ax][sd][w]) CU(cudaMemsetAsync(e->psi[s][m][ax][sd][w], 0, e->psi_elems[s][ax][sd]*sizeof(float), e->stream));
		for (int sd=0; sd<2; sd++) for (int w=0; w<2; w++)
			if (e->psib[s][m][sd][w]) CU(cudaMemsetAsync(e->psib[s][m][sd][w], 0, e->psi_elems[s][2][sd]*sizeof(float), e->stream));
	}
	for (DispObj *o : e->obj) for (int m=0; m<=e->g.complex_field; m++) for (int p=0; p<o->d.nk; p++) for (int c=0; c<3; c++) {
		CU(cudaMemsetAsync(o->J[m][p][c], 0, o->cnt*sizeof(float), e->stream));
		if (o->K[m][p][c]) CU(cudaMemsetAsync(o->K[m][p][c], 0, o->cnt*sizeof(float), e->stream));
	}
	return ALISB_OK;
}

/* ---- data access ------------------------------------------------------------------------- */
static int need_scratch(alisb_engine *e, size_t n)
{
	if (n <= e->scratch_elems) return ALISB_OK;
	cudaFree(e->scratch); e->scratch = 0; e->scratch_elems = 0;
	TRY(dalloc(&e->scratch, n));
	e->scratch_elems = n;
	return ALISB_OK;
}

/* resolves to a padded full-grid style array (ny2 rows per plane, row pitch) or a dense box */
struct ArrRef { float *p; int kind; int ny2; int pitch; int shift; Box dense; };
static int resolve(alisb_engine *e, int arr, const int32_t ids[6], ArrRef &r)
{
	memset(&r, 0, sizeof r);
	switch (arr) {
	case ALISB_ARR_FIELD:
		if (ids[0]<0||ids[0]>1||ids[1]<0||ids[1]>2||ids[2]<0||ids[2]>e->g.complex_field) break;
		r.p = e->F[ids[0]][ids[2]][ids[1]]; r.kind = 0; return ALISB_OK;
	case ALISB_ARR_INVEPS:
		if (ids[0]<0||ids[0]>2||!e->eps[ids[0]]) break;
		r.p = e->eps[ids[0]]; r.kind = 0; return ALISB_OK;
	case ALISB_ARR_PSI: {
		if (ids[0]<0||ids[0]>1||ids[1]<0||ids[1]>e->g.complex_field||ids[2]<0||ids[2]>2||ids[3]<0||ids[3]>1||ids[4]<0||ids[4]>1) break;
		const int s=ids[0], m=ids[1], ax=ids[2], sd=ids[3];
		if (!e->psi[s][m][ax][sd][ids[4]]) break;
		r.p = e->psi[s][m][ax][sd][ids[4]]; r.kind = 1 + ax;
		r.dense.a0 = e->xbase[s][sd]; r.dense.na = e->xpitch[s][sd];
		r.dense.b0 = e->g.pml[s][1][sd][0]; r.dense.nb = e->g.pml[s][1][sd][1] - e->g.pml[s][1][sd][0] + 1;
		r.dense.c0 = e->g.pml[s][2][sd][0]; r.dense.nc = e->g.pml[s][2][sd][1] - e->g.pml[s][2][sd][0] + 1;
		return ALISB_OK; }
	case ALISB_ARR_ADE_J: case ALISB_ARR_ADE_K: {
		if (ids[0]<0||ids[0]>e->g.complex_field||ids[1]<0||ids[1]>=(int)e->obj.size()||ids[2]<0||ids[2]>=e->obj[ids[1]]->d.nk||ids[3]<0||ids[3]>2) break;
		DispObj *o = e->obj[ids[1]];
		r.p = arr==ALISB_ARR_ADE_J ? o->J[ids[0]][ids[2]][ids[3]] : o->K[ids[0]][ids[2]][ids[3]];
		if (!r.p) break;
		r.kind = 4; r.dense = o->full; return ALISB_OK; }
	case ALISB_ARR_ADE_FILL:
		if (ids[0]<0||ids[0]>=(int)e->obj.size()||ids[1]<0||ids[1]>2) break;
		r.p = e->obj[ids[0]]->fill[ids[1]]; r.kind = 4; r.dense = e->obj[ids[0]]->full; return ALISB_OK;
	}
	return fail(ALISB_EINVAL, "bad array id");
}

/* element offset of (a,b,c) in the resolved array, or -1 if outside */
static long long arr_offset(alisb_engine *e, const ArrRef &r, int a, int b, int c)
{
	switch (r.kind) {
	case 0: if (a<0||a>e->nx+1||b<0||b>e->ny+1||c<0||c>e->nz+1) return -1;
	        return ((long long)c*(e->ny+2) + b)*e->pitch + ALISB_XO + a;
	case 1: if (a<r.dense.a0||a>=r.dense.a0+r.dense.na||b<0||b>e->ny+1||c<0||c>e->nz+1) return -1;
	        return ((long long)c*(e->ny+2) + b)*r.dense.na + (a - r.dense.a0);
	case 2: if (a<0||a>e->nx+1||b<r.dense.b0||b>=r.dense.b0+r.dense.nb||c<0||c>e->nz+1) return -1;
	        return ((long long)c*r.dense.nb + (b - r.dense.b0))*e->pitch + ALISB_XO + a;
	case 3: if (a<0||a>e->nx+1||b<0||b>e->ny+1||c<r.dense.c0||c>=r.dense.c0+r.dense.nc) return -1;
	        return ((long long)(c - r.dense.c0)*(e->ny+2) + b)*e->pitch + ALISB_XO + a;
	case 4: if (a<r.dense.a0||a>=r.dense.a0+r.dense.na||b<r.dense.b0||b>=r.dense.b0+r.dense.nb||c<r.dense.c0||c>=r.dense.c0+r.dense.nc) return -1;
	        return ((long long)(c - r.dense.c0)*r.dense.nb + (b - r.dense.b0))*r.dense.na + (a - r.dense.a0);
	}
	return -1;
}

/* ---- mailbox (see box_gather_mail / flux_mail in kernels.cuh) ------------------------------------------ */
#define MAIL_BYTES 4096                 /* payload; the sequence word follows it */
static int mail_open(alisb_engine *e)
{
	if (e->mail_h) return ALISB_OK;
	unsigned char *h = 0, *d = 0;
	CU(cudaHostAlloc((void**)&h, MAIL_BYTES + 64, cudaHostAllocMapped));
	memset(h, 0, MAIL_BYTES + 64);
	if (cudaHostGetDevicePointer((void**)&d, h, 0) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(h); return fail(ALISB_ECUDA, "mailbox: no device pointer for mapped host memory"); }
	unsigned *cnt = 0; double *part = 0;
	if (cudaMalloc((void**)&cnt, sizeof(unsigned)*(1 + ALISB_FLUX_MAIL_PLANES)) != cudaSuccess || cudaMalloc((void**)&part, sizeof(double)*2*(128 + 1)*ALISB_FLUX_MAIL_PLANES) != cudaSuccess
	    || cudaMemset(cnt, 0, sizeof(unsigned)*(1 + ALISB_FLUX_MAIL_PLANES)) != cudaSuccess) {
		cudaGetLastError(); cudaFree(cnt); cudaFree(part); cudaFreeHost(h); return fail(ALISB_ECUDA, "mailbox: allocation failed");
	}
	e->mail_h = h; e->mail_d = d; e->d_mail_counters = cnt; e->d_mail_partial = part; e->mail_seq = 0;
	return ALISB_OK;
}
/* waits for the kernel that was launched with sequence number `seq`; polls the stream now and then so that a launch
 * that died does not hang the host */
static int mail_wait(alisb_engine *e, unsigned seq, bool pair = false)
{
	/* pair: a single value and its sequence number in one 8-byte word (floats 1022, 1023 of the payload: value, number) */
	volatile unsigned *flag = pair ? (volatile unsigned*)(e->mail_h + 1023*sizeof(float)) : (volatile unsigned*)(e->mail_h + MAIL_BYTES);
	for (unsigned long spins = 1; *flag != seq; spins++) {
		if ((spins & 0xffff) == 0) {
			const cudaError_t st = cudaStreamQuery(e->stream);
			if (st == cudaSuccess) { if (*flag == seq) break; return fail(ALISB_ECUDA, "mailbox: the stream drained without the result arriving"); }
			if (st != cudaErrorNotReady) return fail(ALISB_ECUDA, "mailbox: %s", cudaGetErrorString(st));
		}
#if defined(__x86_64__) || defined(__i386__)
		__builtin_ia32_pause();
#endif
	}
	__sync_synchronize();
	return ALISB_OK;
}

static int rw_box_one(alisb_engine *e, int arr, const int32_t ids[6], const int32_t box[6], float *host, bool write)
{
	DEV(e);
	ArrRef r;
	TRY(resolve(e, arr, ids, r));
	Box B = mkbox(box);
	const size_t n = bcount(B);
	if (!n) return ALISB_OK;
	if (arr_offset(e, r, box[0], box[2], box[4]) < 0 || arr_offset(e, r, box[1], box[3], box[5]) < 0) return fail(ALISB_EINVAL, "box outside array");
	/* strides of the resolved array: element (a,b,c) at c*sC + b*sB + a + off */
	long long sC, sB, off;
	switch (r.kind) {
	case 0:  sB = e->pitch; sC = (long long)(e->ny+2)*e->pitch; off = ALISB_XO; break;
	case 1:  sB = r.dense.na; sC = (long long)(e->ny+2)*r.dense.na; off = -(long long)r.dense.a0; break;
	case 2:  sB = e->pitch; sC = (long long)r.dense.nb*e->pitch; off = ALISB_XO - (long long)r.dense.b0*sB; break;
	case 3:  sB = e->pitch; sC = (long long)(e->ny+2)*e->pitch; off = ALISB_XO - (long long)r.dense.c0*sC; break;
	default: sB = r.dense.na; sC = (long long)r.dense.nb*r.dense.na; off = -((long long)r.dense.c0*sC + (long long)r.dense.b0*sB + r.dense.a0); break;
	}
	if (!write && e->tune_mail && n*sizeof(float) <= MAIL_BYTES - 2*sizeof(float)) {
		TRY(mail_open(e));
		const unsigned seq = ++e->mail_seq;
		box_gather_mail<<<1, n == 1 ? 32 : 256, 0, e->stream>>>(r.p, (float*)e->mail_d, B, sC, sB, off, (volatile unsigned*)(e->mail_d + MAIL_BYTES), seq);
		CU(cudaGetLastError());
		e->launches++;
		TRY(mail_wait(e, seq, n == 1));
		if (n == 1) {
			/* value and number were stored as one 8-byte word; read them the same way */
			const unsigned long long w = *(volatile unsigned long long*)(e->mail_h + 1022*sizeof(float));
			if ((unsigned)(w >> 32) != seq) return fail(ALISB_ECUDA, "mailbox: torn probe word");
			const unsigned bits = (unsigned)w;
			memcpy(host, &bits, sizeof(float));
		} else memcpy(host, e->mail_h, n*sizeof(float));
		return ALISB_OK;
	}
	TRY(need_scratch(e, n));
	if (write) {
		CU(cudaMemcpyAsync(e->scratch, host, n*sizeof(float), cudaMemcpyHostToDevice, e->stream));
		box_scatter<<<div_up(n, 256), 256, 0, e->stream>>>(r.p, e->scratch, B, sC, sB, off);
		CU(cudaStreamSynchronize(e->stream));
	} else {
		box_gather<<<div_up(n, 256), 256, 0, e->stream>>>(r.p, e->scratch, B, sC, sB, off);
		CU(cudaMemcpyAsync(host, e->scratch, n*sizeof(float), cudaMemcpyDeviceToHost, e->stream));
		CU(cudaStreamSynchronize(e->stream));
	}
	e->launches++;
	return ALISB_OK;
}
static int rw_box(alisb_engine *e, int arr, const int32_t ids[6], const int32_t box[6], float *host, bool write)
{
	DEV(e);
	/* in-place fused stepping cannot take back the H' it has already formed: after alisb_flush(e, 1) the H pass must
	 * follow before anything is written (the public API never opens this window: updateH issues both at once) */
	if (write && e->h_fused && e->tune_wave > 0) return fail(ALISB_ESTATE, "in-place fused stepping: no writes between alisb_flush(e, 1) and the H pass");
	if (write) UNFUSE(e); else SETTLE(e);
	TRY(rw_box_one(e, arr, ids, box, host, write));
	if (write && arr == ALISB_ARR_FIELD && e->dbuf) {
		/* both buffers, so that ghost cells and cells no pass rewrites stay in agreement */
		float **cur = &e->F[ids[0]][ids[2]][ids[1]], **oth = &e->Fb[ids[0]][ids[2]][ids[1]];
		std::swap(*cur, *oth);
		const int r = rw_box_one(e, arr, ids, box, host, true);
		std::swap(*cur, *oth);
		return r;
	}
	if (write && arr == ALISB_ARR_PSI && ids[2] == 2 && e->dbuf && e->psib[ids[0]][ids[1]][ids[3]][ids[4]]) {
		float **cur = &e->psi[ids[0]][ids[1]][2][ids[3]][ids[4]], **oth = &e->psib[ids[0]][ids[1]][ids[3]][ids[4]];
		std::swap(*cur, *oth);
		const int r = rw_box_one(e, arr, ids, box, host, true);
		std::swap(*cur, *oth);
		return r;
	}
	return ALISB_OK;
}

extern "C" int alisb_read_box(alisb_engine *e, int arr, const int32_t ids[6], const int32_t box[6], float *host) { return rw_box(e, arr, ids, box, host, false); }
extern "C" int alisb_write_box(alisb_engine *e, int arr, const int32_t ids[6], const int32_t box[6], const float *host) { return rw_box(e, arr, ids, box, (float*)host, true); }

/* ---- second-order mixing between worlds (nonLinearSHG / nonLinearTWM, update.c:344-411) -------- */
extern "C" int alisb_chi2_couple(alisb_engine *e, int obj, int m, alisb_engine *p1, alisb_engine *p2, const float *d, float alt)
{
	if (!e || !p1 || !p2 || !d) return fail(ALISB_EINVAL, "chi2: null argument");
	if (obj < 0 || obj >= (int)e->obj.size()) return fail(ALISB_EINVAL, "chi2: no such object");
	if (m < 0 || m > e->g.complex_field || m > p1->g.complex_field || m > p2->g.complex_field) return fail(ALISB_EINVAL, "chi2: no such field set");
	for (alisb_engine *q : {p1, p2})
		if (q->nx != e->nx || q->ny != e->ny || q->nz != e->nz || q->pitch != e->pitch || q->g.device != e->g.device)
			return fail(ALISB_EINVAL, "chi2: the coupled worlds must share one grid and one device");
	if (!e->eps[0]) return fail(ALISB_EINVAL, "chi2: the signal world has no material arrays");
	DEV(e);
	UNFUSE(e); SETTLE(p1); SETTLE(p2);
	const DispObj *o = e->obj[obj];
	Chi2Params p;
	memset(&p, 0, sizeof p);
	for (int c=0; c<3; c++) {
		p.F[c] = e->F[ALISB_SET_E][m][c]; p.G1[c] = p1->F[ALISB_SET_E][m][c]; p.G2[c] = p2->F[ALISB_SET_E][m][c];
		p.eps[c] = e->eps[c]; p.fill[c] = o->fill[c];
		for (int q=0; q<6; q++) p.d[c][q] = d[c*6 + q];
	}
	p.alt = alt; p.full = o->full;
	p.upd = o->full; p.upd.na--; p.upd.nb--; p.upd.nc--;                  /* iMin..iMax-1 etc. (update.c:359-361) */
	p.ny = e->ny; p.pitch = e->pitch; p.plane = (long long)e->plane;
	if (p.upd.na <= 0 || p.upd.nb <= 0 || p.upd.nc <= 0) return ALISB_OK;
	/* order after whatever the pump worlds have queued, and make them wait for us afterwards */
	for (alisb_engine *q : {p1, p2}) if (q->stream != e->stream) {
		CU(cudaEventRecord(q->ev_fork, q->stream)); CU(cudaStreamWaitEvent(e->stream, q->ev_fork, 0));
	}
	const size_t n = bcount(p.upd);
	const int T = 256; const unsigned nb = (unsigned)((n + T - 1)/T);
	chi2_couple<0><<<nb, T, 0, e->stream>>>(p);
	chi2_couple<1><<<nb, T, 0, e->stream>>>(p);
	chi2_couple<2><<<nb, T, 0, e->stream>>>(p);
	e->launches += 3;
	CU(cudaGetLastError());
	for (alisb_engine *q : {p1, p2}) if (q->stream != e->stream) {
		CU(cudaEventRecord(e->ev_fork, e->stream)); CU(cudaStreamWaitEvent(q->stream, e->ev_fork, 0));
	}
	return ALISB_OK;
}

/* ---- device-side monitor reductions ------------------------------------------------------------- */
extern "C" int alisb_set_spacing(alisb_engine *e, int set, int axis, const float *d)
{
	if (set<0 || set>1 || axis<0 || axis>2 || !d) return fail(ALISB_EINVAL, "bad spacing table");
	DEV(e);
	const int n = axis==0 ? e->nx : axis==1 ? e->ny : e->nz;
	if (!e->spacing[set][axis]) TRY(dalloc(&e->spacing[set][axis], n + 2));
	CU(cudaMemcpyAsync(e->spacing[set][axis], d, (size_t)(n + 2)*sizeof(float), cudaMemcpyHostToDevice, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	return ALISB_OK;
}

#define FLUX_BLOCKS 64
extern "C" int alisb_plane_flux(alisb_engine *e, const alisb_flux_plane *planes, int n, double *out)
{
	NvtxRange range("alisb:plane_flux");
	if (n < 1 || !planes || !out) return fail(ALISB_EINVAL, "bad flux request");
	DEV(e);
	SETTLE(e);
	for (int s=0; s<2; s++) for (int ax=0; ax<3; ax++) if (!e->spacing[s][ax]) return fail(ALISB_ESTATE, "alisb_set_spacing has not been called");
	if (n > e->flux_cap) {
		/* the old buffers go first and the capacity only grows once every new one exists: a failed
		 * allocation leaves an engine with capacity 0 and null pointers, never dangling ones */
		cudaFree(e->d_planes); cudaFree(e->d_partial); cudaFree(e->d_flux); if (e->h_flux) cudaFreeHost(e->h_flux);
		e->d_planes = 0; e->d_partial = 0; e->d_flux = 0; e->h_flux = 0; e->flux_cap = 0;
		const int cap = std::max(16, n);
		CU(cudaMalloc((void**)&e->d_planes, sizeof(FluxPlane)*cap));
		CU(cudaMalloc((void**)&e->d_partial, sizeof(double)*2*FLUX_BLOCKS*cap));
		CU(cudaMalloc((void**)&e->d_flux, sizeof(double)*2*cap));
		CU(cudaMallocHost((void**)&e->h_flux, sizeof(double)*2*cap + sizeof(FluxPlane)*cap));
		e->flux_cap = cap;
	}
	FluxPlane *hp = (FluxPlane*)(e->h_flux + 2*e->flux_cap);
	const int N[3] = {e->nx, e->ny, e->nz};
	for (int q=0; q<n; q++) {
		const alisb_flux_plane &f = planes[q];
		if (f.axis<0 || f.axis>2 || f.at<0 || f.at>N[f.axis] || f.b0<0 || f.b1>N[(f.axis+1)%3]+1 || f.c0<0 || f.c1>N[(f.axis+2)%3]+1) return fail(ALISB_EINVAL, "flux plane outside the grid");
		hp[q].axis = f.axis; hp[q].at = f.at; hp[q].b0 = f.b0; hp[q].b1 = f.b1; hp[q].c0 = f.c0; hp[q].c1 = f.c1;
	}
	FluxParams p;
	for (int c=0; c<3; c++) { p.E[c] = e->F[0][0][c]; p.H[c] = e->F[1][0][c]; p.dE[c] = e->spacing[0][c]; p.dH[c] = e->spacing[1][c]; }
	p.ny = e->ny; p.pitch = e->pitch;
	if (e->tune_mail && n <= ALISB_FLUX_MAIL_PLANES && sizeof(double)*2*n <= MAIL_BYTES) {
		/* one launch, planes as kernel arguments, sums into the mailbox (same partial sums added in the same order) */
		TRY(mail_open(e));
		FluxPlaneList PL;
		memset(&PL, 0, sizeof PL);
		memcpy(PL.P, hp, sizeof(FluxPlane)*n);
		const unsigned seq = ++e->mail_seq;
		flux_mail<<<dim3(FLUX_BLOCKS, n), 256, 0, e->stream>>>(p, PL, e->d_mail_partial, e->d_mail_counters, (double*)e->mail_d, (volatile unsigned*)(e->mail_d + MAIL_BYTES), seq);
		CU(cudaGetLastError());
		e->launches++;
		TRY(mail_wait(e, seq));
		memcpy(out, e->mail_h, sizeof(double)*2*n);
		return ALISB_OK;
	}
	CU(cudaMemcpyAsync(e->d_planes, hp, sizeof(FluxPlane)*n, cudaMemcpyHostToDevice, e->stream));
	flux_partial<<<dim3(FLUX_BLOCKS, n), 256, 0, e->stream>>>(p, e->d_planes, e->d_partial);
	flux_finish<<<n, 32, 0, e->stream>>>(e->d_partial, FLUX_BLOCKS, e->d_flux);
	CU(cudaMemcpyAsync(e->h_flux, e->d_flux, sizeof(double)*2*n, cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	memcpy(out, e->h_flux, sizeof(double)*2*n);
	e->launches += 2;
	return ALISB_OK;
}

/* ---- device-side field functions (kernels eval_box, reduce_absmax, column_sums) ------------------ */
static int eval_setup(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], bool columns, EvalParams &p)
{
	if (!f || !ibox) return fail(ALISB_EINVAL, "eval: null argument");
	if (f->fn < ALISB_FN_RAW || f->fn > ALISB_FN_DIVE || f->set < 0 || f->set > 1 || f->comp < 0 || f->comp > 2 || f->comp2 < 0 || f->comp2 > 2 || f->m < 0 || f->m > 1)
		return fail(ALISB_EINVAL, "eval: bad field function");
	if ((f->fn == ALISB_FN_EXH || f->fn == ALISB_FN_TOFF) && f->comp == f->comp2) return fail(ALISB_EINVAL, "eval: the two components must differ");
	if (f->m > e->g.complex_field && f->fn != ALISB_FN_NODE) return fail(ALISB_EINVAL, "eval: no such field set");
	const int n[3] = { e->nx, e->ny, e->nz };
	const int reach = f->fn == ALISB_FN_RAW ? 0 : 1;           /* the stencils read index +1 */
	for (int ax=0; ax<3; ax++) {
		int lo = ibox[2*ax], hi = ibox[2*ax+1];
		if (lo > hi) return fail(ALISB_EINVAL, "eval: empty box");
		if (columns && ax == 2) { if (lo + f->org[2] < 0 || hi + f->org[2] > n[2] + 1) return fail(ALISB_EINVAL, "eval: column range outside the spacing table"); }
		if (f->sym[ax] && lo < 0) { const int far = std::max(-lo, hi); lo = hi < 0 ? -hi : 0; hi = far; }
		if (lo + f->org[ax] < 0 || hi + f->org[ax] + reach > n[ax] + 1) return fail(ALISB_EINVAL, "eval: box outside the arrays");
	}
	memset(&p, 0, sizeof p);
	for (int s=0; s<2; s++) for (int m=0; m<=e->g.complex_field; m++) for (int c=0; c<3; c++) p.F[s][m][c] = e->F[s][m][c];
	for (int ax=0; ax<3; ax++) p.dtd[ax] = e->dtd[0][ax];
	p.dz = e->spacing[0][2];
	p.f = *f; p.ny = e->ny; p.pitch = e->pitch; p.complex_field = e->g.complex_field; p.dt = e->g.dt;
	p.i0 = ibox[0]; p.j0 = ibox[2]; p.k0 = ibox[4];
	p.ni = ibox[1]-ibox[0]+1; p.nj = ibox[3]-ibox[2]+1; p.nk = ibox[5]-ibox[4]+1;
	if (columns && !p.dz) return fail(ALISB_ESTATE, "alisb_set_spacing has not been called");
	return ALISB_OK;
}

extern "C" int alisb_eval_box(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], float *host)
{
	NvtxRange range("alisb:eval_box");
	DEV(e);
	EvalParams p;
	TRY(eval_setup(e, f, ibox, false, p));
	if (!host) return fail(ALISB_EINVAL, "eval: null output");
	SETTLE(e);
	const size_t n = (size_t)p.ni*p.nj*p.nk;
	TRY(need_scratch(e, n));
	eval_box<<<div_up(n, 256), 256, 0, e->stream>>>(p, e->scratch);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(host, e->scratch, n*sizeof(float), cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	e->launches++;
	return ALISB_OK;
}

struct alisb_accum { alisb_engine *e; alisb_field_fn f; int32_t ibox[6]; size_t n; float *re, *im; };

extern "C" int alisb_accum_create(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], int with_imag, alisb_accum **out)
{
	if (!e || !out) return fail(ALISB_EINVAL, "accum: null argument");
	DEV(e);
	EvalParams p;
	TRY(eval_setup(e, f, ibox, false, p));
	alisb_accum *a = new alisb_accum();
	a->e = e; a->f = *f; memcpy(a->ibox, ibox, sizeof a->ibox); a->n = (size_t)p.ni*p.nj*p.nk; a->re = a->im = 0;
	int rc = dalloc(&a->re, a->n);
	if (rc == ALISB_OK && with_imag) rc = dalloc(&a->im, a->n);
	if (rc != ALISB_OK) { cudaFree(a->re); cudaFree(a->im); delete a; return rc; }
	*out = a;
	return ALISB_OK;
}

extern "C" int alisb_accum_add(alisb_accum *a, float wr, float wi)
{
	if (!a) return fail(ALISB_EINVAL, "accum: null argument");
	NvtxRange range("alisb:accum_add");
	alisb_engine *e = a->e;
	DEV(e);
	EvalParams p;
	TRY(eval_setup(e, &a->f, a->ibox, false, p));          /* the field pointers may have changed (double-buffered passes) */
	SETTLE(e);
	for (int s=0; s<2; s++) for (int m=0; m<=e->g.complex_field; m++) for (int c=0; c<3; c++) p.F[s][m][c] = e->F[s][m][c];
	eval_accumulate<<<div_up(a->n, 256), 256, 0, e->stream>>>(p, a->re, a->im, wr, wi);
	CU(cudaGetLastError());
	e->launches++;
	return ALISB_OK;
}

extern "C" int alisb_accum_read(alisb_accum *a, float *re, float *im)
{
	if (!a || !re || (im && !a->im)) return fail(ALISB_EINVAL, "accum: bad read");
	alisb_engine *e = a->e;
	DEV(e);
	CU(cudaMemcpyAsync(re, a->re, a->n*sizeof(float), cudaMemcpyDeviceToHost, e->stream));
	if (im) CU(cudaMemcpyAsync(im, a->im, a->n*sizeof(float), cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	return ALISB_OK;
}

extern "C" int alisb_accum_clear(alisb_accum *a)
{
	if (!a) return fail(ALISB_EINVAL, "accum: null argument");
	alisb_engine *e = a->e;
	DEV(e);
	CU(cudaMemsetAsync(a->re, 0, a->n*sizeof(float), e->stream));
	if (a->im) CU(cudaMemsetAsync(a->im, 0, a->n*sizeof(float), e->stream));
	return ALISB_OK;
}

extern "C" void alisb_accum_destroy(alisb_accum *a)
{
	if (!a) return;
	cudaSetDevice(a->e->g.device);
	cudaStreamSynchronize(a->e->stream);
	cudaFree(a->re); cudaFree(a->im);
	delete a;
}

extern "C" int alisb_reduce_absmax(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], float *out)
{
	NvtxRange range("alisb:reduce_absmax");
	DEV(e);
	EvalParams p;
	TRY(eval_setup(e, f, ibox, false, p));
	if (!out) return fail(ALISB_EINVAL, "eval: null output");
	SETTLE(e);
	const size_t n = (size_t)p.ni*p.nj*p.nk;
	TRY(need_scratch(e, 1));
	CU(cudaMemsetAsync(e->scratch, 0, sizeof(float), e->stream));
	const int blocks = (int)std::min<size_t>(div_up(n, 256), (size_t)e->resident_ctas*4);
	reduce_absmax<<<blocks, 256, 0, e->stream>>>(p, (unsigned*)e->scratch);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out, e->scratch, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	e->launches++;
	return ALISB_OK;
}

extern "C" int alisb_column_sums(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], float *host)
{
	NvtxRange range("alisb:column_sums");
	DEV(e);
	EvalParams p;
	TRY(eval_setup(e, f, ibox, true, p));
	if (!host) return fail(ALISB_EINVAL, "eval: null output");
	SETTLE(e);
	const size_t n = (size_t)p.ni*p.nj;
	TRY(need_scratch(e, n));
	column_sums<<<div_up(n, 128), 128, 0, e->stream>>>(p, e->scratch);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(host, e->scratch, n*sizeof(float), cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	e->launches++;
	return ALISB_OK;
}

extern "C" int alisb_absorption_columns(alisb_engine *e, int obj, const int32_t box[6], float *host)
{
	NvtxRange range("alisb:absorption_columns");
	if (!box || !host) return fail(ALISB_EINVAL, "absorption: null argument");
	if (obj < 0 || obj >= (int)e->obj.size()) return fail(ALISB_EINVAL, "absorption: no such object");
	if (!e->eps[0] || !e->eps[1] || !e->eps[2]) return fail(ALISB_ESTATE, "absorption: no material arrays");
	if (!e->spacing[0][2] || !e->spacing[1][2]) return fail(ALISB_ESTATE, "alisb_set_spacing has not been called");
	DEV(e);
	const DispObj *o = e->obj[obj];
	const Box &B = o->full;
	if (box[0] > box[1] || box[2] > box[3] || box[4] > box[5] ||
	    box[0] < std::max(B.a0, 1) || box[1] > std::min(B.a0 + B.na - 1, e->nx) || box[2] < std::max(B.b0, 1) || box[3] > std::min(B.b0 + B.nb - 1, e->ny) ||
	    box[4] < std::max(B.c0, 1) || box[5] > std::min(B.c0 + B.nc - 1, e->nz)) return fail(ALISB_EINVAL, "absorption: box outside the object");
	SETTLE(e);
	EvalParams p;
	memset(&p, 0, sizeof p);
	for (int c=0; c<3; c++) {
		p.F[0][0][c] = e->F[0][0][c]; p.inv[c] = e->eps[c]; p.fill[c] = o->fill[c]; p.loss[c] = o->d.a[c];
		for (int q=0; q<o->d.nk; q++) p.J[q][c] = o->J[0][q][c];
	}
	p.npoles = o->d.nk; p.full = B; p.dz = e->spacing[0][2]; p.dzH = e->spacing[1][2];
	p.ny = e->ny; p.pitch = e->pitch;
	p.i0 = box[0]; p.j0 = box[2]; p.k0 = box[4];
	p.ni = box[1]-box[0]+1; p.nj = box[3]-box[2]+1; p.nk = box[5]-box[4]+1;
	const size_t n = (size_t)p.ni*p.nj;
	TRY(need_scratch(e, 3*n));
	absorption_columns<<<div_up(n, 128), 128, 0, e->stream>>>(p, 1, e->nz, e->scratch);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(host, e->scratch, 3*n*sizeof(float), cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	e->launches++;
	return ALISB_OK;
}

/* ---- geometry rasterisation (kernels geometry_bounds, geometry_fill) ------------------------------- */
struct GeoDev {
	alisb_geometry g; void *p[6]; size_t old_stack; bool stack_set;
	GeoDev() { memset(&g, 0, sizeof g); memset(p, 0, sizeof p); old_stack = 0; stack_set = false; }
	~GeoDev() { for (void *q : p) cudaFree(q); if (stack_set) cudaDeviceSetLimit(cudaLimitStackSize, old_stack); }
};

static int geo_check(const alisb_geometry *g, const int32_t ibox[6])
{
	if (!g || !ibox || g->nobjects < 1 || !g->roots || g->nnodes < 1 || !g->nodes || g->subn < 1 || g->subn > 64) return fail(ALISB_EINVAL, "geometry: bad description");
	for (int a=0; a<3; a++) {
		if (!g->coord[a] || g->lo[a] > g->hi[a]) return fail(ALISB_EINVAL, "geometry: missing coordinate table");
		if (ibox[2*a] > ibox[2*a+1] || ibox[2*a] < g->lo[a] || ibox[2*a+1] > g->hi[a]) return fail(ALISB_EINVAL, "geometry: box outside the coordinate tables");
	}
	for (int o=0; o<g->nobjects; o++) if (g->roots[o] < 0 || g->roots[o] >= g->nnodes) return fail(ALISB_EINVAL, "geometry: bad root");
	for (int n=0; n<g->nnodes; n++) {
		const alisb_shape &s = g->nodes[n];
		if (s.type < 0 || s.type >= ALISB_SHAPE_COUNT) return fail(ALISB_EINVAL, "geometry: unknown shape type");
		if (s.type > ALISB_SHAPE_HRODZY) {
			/* children come after their parent: the tree is acyclic and recursion ends */
			if (s.nchild < 1 || s.child <= n || s.child + s.nchild > g->nnodes) return fail(ALISB_EINVAL, "geometry: bad child range");
			const int need = s.type == ALISB_SHAPE_ROTATION ? 6 : (s.type >= ALISB_SHAPE_CLATTICEX ? 2*(int)s.f[0][0] : 0);
			if (need && (s.aux < 0 || s.aux + need > g->naux || !g->aux)) return fail(ALISB_EINVAL, "geometry: bad constant range");
		}
	}
	return ALISB_OK;
}

/* device copy of the description; raises the per-thread stack for the recursive interpreter */
static int geo_upload(int device, const alisb_geometry *g, GeoDev &d)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(ALISB_ECUDA, "geometry: no CUDA device %d (this library has no CPU fallback)", device);
	CU(cudaSetDevice(device));
	d.g = *g;
	const size_t slots = (size_t)g->subn + 3;
	const void *src[6] = { g->roots, g->nodes, g->aux, g->coord[0], g->coord[1], g->coord[2] };
	const size_t bytes[6] = { sizeof(int32_t)*g->nobjects, sizeof(alisb_shape)*g->nnodes, sizeof(float)*std::max(g->naux, 1),
		sizeof(float)*2*slots*(g->hi[0]-g->lo[0]+1), sizeof(float)*2*slots*(g->hi[1]-g->lo[1]+1), sizeof(float)*2*slots*(g->hi[2]-g->lo[2]+1) };
	for (int q=0; q<6; q++) {
		CU(cudaMalloc(&d.p[q], bytes[q]));
		if (src[q] && !(q == 2 && g->naux == 0)) CU(cudaMemcpy(d.p[q], src[q], bytes[q], cudaMemcpyHostToDevice));
	}
	d.g.roots = (const int32_t*)d.p[0]; d.g.nodes = (const alisb_shape*)d.p[1]; d.g.aux = (const float*)d.p[2];
	for (int a=0; a<3; a++) d.g.coord[a] = (const float*)d.p[3+a];
	CU(cudaDeviceGetLimit(&d.old_stack, cudaLimitStackSize));
	if (d.old_stack < 4096) { CU(cudaDeviceSetLimit(cudaLimitStackSize, 4096)); d.stack_set = true; }   /* geo_inside: 144 B per level */
	return ALISB_OK;
}

extern "C" int alisb_geometry_bounds(int device, const alisb_geometry *g, const int32_t ibox[6], int32_t *bounds)
{
	TRY(geo_check(g, ibox));
	if (!bounds) return fail(ALISB_EINVAL, "geometry: null output");
	GeoDev d;
	TRY(geo_upload(device, g, d));
	const int ni = ibox[1]-ibox[0]+1, nj = ibox[3]-ibox[2]+1, nk = ibox[5]-ibox[4]+1;
	std::vector<int32_t> h((size_t)6*g->nobjects);
	for (int o=0; o<g->nobjects; o++) for (int a=0; a<3; a++) { h[6*o+2*a] = INT32_MAX; h[6*o+2*a+1] = INT32_MIN; }
	int *db = 0;
	CU(cudaMalloc((void**)&db, h.size()*sizeof(int32_t)));
	cudaError_t ce = cudaMemcpy(db, h.data(), h.size()*sizeof(int32_t), cudaMemcpyHostToDevice);
	if (ce == cudaSuccess) {
		const size_t n = (size_t)ni*nj*nk;
		geometry_bounds<<<div_up(n, 128), 128>>>(d.g, ibox[0], ibox[2], ibox[4], ni, nj, nk, db);
		ce = cudaGetLastError();
		if (ce == cudaSuccess) ce = cudaMemcpy(h.data(), db, h.size()*sizeof(int32_t), cudaMemcpyDeviceToHost);
	}
	cudaFree(db);
	if (ce != cudaSuccess) return fail(ALISB_ECUDA, "geometry_bounds: %s", cudaGetErrorString(ce));
	memcpy(bounds, h.data(), h.size()*sizeof(int32_t));
	return ALISB_OK;
}

extern "C" int alisb_geometry_fill(int device, const alisb_geometry *g, int object, const int32_t ibox[6], float *fx, float *fy, float *fz)
{
	TRY(geo_check(g, ibox));
	if (object < 0 || object >= g->nobjects || !fx || !fy || !fz) return fail(ALISB_EINVAL, "geometry: bad object or null output");
	for (int a=0; a<3; a++) if (ibox[2*a] - 1 < g->lo[a] - 1 || ibox[2*a] < g->lo[a]) return fail(ALISB_EINVAL, "geometry: box outside the coordinate tables");
	GeoDev d;
	TRY(geo_upload(device, g, d));
	const int ni = ibox[1]-ibox[0]+1, nj = ibox[3]-ibox[2]+1, nk = ibox[5]-ibox[4]+1;
	const size_t n = (size_t)ni*nj*nk;
	float *df = 0;
	CU(cudaMalloc((void**)&df, 3*n*sizeof(float)));
	geometry_fill<<<div_up(3*n, 128), 128>>>(d.g, object, ibox[0], ibox[2], ibox[4], ni, nj, nk, df, df + n, df + 2*n);
	cudaError_t ce = cudaGetLastError();
	float *dst[3] = { fx, fy, fz };
	for (int c=0; c<3 && ce == cudaSuccess; c++) ce = cudaMemcpy(dst[c], df + c*n, n*sizeof(float), cudaMemcpyDeviceToHost);
	cudaFree(df);
	if (ce != cudaSuccess) return fail(ALISB_ECUDA, "geometry_fill: %s", cudaGetErrorString(ce));
	return ALISB_OK;
}

/* ---- halo: planes are contiguous in this layout, hand out pointers into the arrays -------- */
extern "C" int alisb_halo_info(alisb_engine *e, int set, int m, int comp, void **send_ptr, void **recv_ptr, size_t *bytes)
{
	if (set<0||set>1||m<0||m>e->g.complex_field||comp<0||comp>1) return fail(ALISB_EINVAL, "bad halo id");
	float *F = e->F[set][m][comp];
	if (send_ptr) *send_ptr = F + (set == ALISB_SET_E ? (size_t)e->nz*e->plane : e->plane);
	if (recv_ptr) *recv_ptr = F + (set == ALISB_SET_E ? 0 : (size_t)(e->nz + 1)*e->plane);
	if (bytes) *bytes = e->plane*sizeof(float);
	return ALISB_OK;
}
extern "C" int alisb_halo_pack(alisb_engine *e, int set) { (void)e; (void)set; return ALISB_OK; }
extern "C" int alisb_halo_unpack(alisb_engine *e, int set) { (void)e; (void)set; return ALISB_OK; }

/* both slabs in one process (host/shard.c): the plane crosses the cut as a peer copy on the owner's stream */
extern "C" int alisb_halo_exchange(alisb_engine *lower, alisb_engine *upper, int set)
{
	if (!lower || !upper || set < 0 || set > 1) return fail(ALISB_EINVAL, "halo exchange: bad argument");
	if (lower->nx != upper->nx || lower->ny != upper->ny || lower->plane != upper->plane || lower->g.complex_field != upper->g.complex_field)
		return fail(ALISB_EINVAL, "halo exchange: slabs differ in x, y or field sets");
	alisb_engine *src = set == ALISB_SET_E ? lower : upper, *dst = set == ALISB_SET_E ? upper : lower;
	if (src->g.device != dst->g.device) {
		/* direct peer access (NVLink / NVSwitch) once per ordered pair; without it the copy is staged by the driver */
		static unsigned long long tried[64];
		const int a = src->g.device & 63, b = dst->g.device & 63;
		if (!(tried[a] >> b & 1ull)) {
			tried[a] |= 1ull << b; tried[b] |= 1ull << a;
			int ok = 0;
			if (cudaDeviceCanAccessPeer(&ok, a, b) == cudaSuccess && ok) { cudaSetDevice(a); cudaDeviceEnablePeerAccess(b, 0); }
			ok = 0;
			if (cudaDeviceCanAccessPeer(&ok, b, a) == cudaSuccess && ok) { cudaSetDevice(b); cudaDeviceEnablePeerAccess(a, 0); }
			cudaGetLastError();                             /* "already enabled" is fine */
		}
	}
	/* the plane must exist: a deferred E pass is issued now, as the plain variant (a caller that wants the fused
	 * variant sends the plane first: alisb_step_part(SEND_PLANE), this call, then the rest of the pass) */
	{ alisb_engine *e = src; DEV(e); TRY(launch_pending(e, false)); }
	{ alisb_engine *e = dst; DEV(e); TRY(launch_pending(e, false)); }
	for (alisb_engine *e : {src, dst}) {
		DEV(e);
		if (!e->comm) {
			CU(cudaStreamCreateWithFlags(&e->comm, cudaStreamNonBlocking));
			CU(cudaEventCreateWithFlags(&e->ev_plane, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&e->ev_war, cudaEventDisableTiming));
			for (int q=0; q<2; q++) CU(cudaEventCreateWithFlags(&e->ev_out[q], cudaEventDisableTiming));
		}
	}
	{
		alisb_engine *e = dst; DEV(e);
		/* an earlier copy into the same ghost plane that nobody has waited for yet is still ordered before this one */
		if (e->wait_in[set]) { CU(cudaStreamWaitEvent(e->stream, e->ev_in[set], 0)); e->wait_in[set] = false; }
		CU(cudaEventRecord(e->ev_war, e->stream));           /* what the receiver has enqueued so far may still read the old ghost plane */
	}
	{
		alisb_engine *e = src; DEV(e);
		CU(cudaEventRecord(e->ev_plane, e->stream));
		CU(cudaStreamWaitEvent(e->comm, e->ev_plane, 0));
		CU(cudaStreamWaitEvent(e->comm, dst->ev_war, 0));
		for (int m=0; m<=e->g.complex_field; m++) for (int comp=0; comp<2; comp++) {
			void *sp = 0, *rp = 0; size_t nb = 0;
			TRY(alisb_halo_info(src, set, m, comp, &sp, 0, &nb));
			TRY(alisb_halo_info(dst, set, m, comp, 0, &rp, 0));
			CU(cudaMemcpyPeerAsync(rp, dst->g.device, sp, src->g.device, nb, e->comm));
		}
		CU(cudaEventRecord(e->ev_out[set], e->comm));
		e->wait_out[set] = true;
		dst->ev_in[set] = e->ev_out[set]; dst->wait_in[set] = true;
	}
	return ALISB_OK;
}

/* ---- phasers -------------------------------------------------------------------------------- */
extern "C" int alisb_phaser_create(alisb_engine *e, const int32_t box[6], int face_mask, int nfreq, alisb_phaser **out)
{
	if (nfreq < 1 || !out) return fail(ALISB_EINVAL, "nfreq");
	DEV(e);
	alisb_phaser *p = new alisb_phaser();
	p->e = e; memcpy(p->box, box, sizeof p->box); p->mask = face_mask; p->nfreq = nfreq;
	p->d_jobs = 0; p->d_w = 0; p->d_sums = 0; p->jobs_cap = 0; p->w_cap = 0;
	for (int f=0; f<6; f++) {
		p->acc[f] = 0; p->cells[f] = 0;
		if (!(face_mask & (1<<f))) continue;
		int32_t fb[6]; memcpy(fb, box, sizeof fb);
		const int ax = f/2, sd = f%2;
		fb[2*ax] = fb[2*ax+1] = sd ? box[2*ax+1] : box[2*ax];
		p->face[f] = mkbox(fb); p->cells[f] = bcount(p->face[f]);
		float *a = 0; TRY(dalloc(&a, p->cells[f]*8*(size_t)nfreq)); p->acc[f] = (float2*)a;
	}
	*out = p;
	return ALISB_OK;
}

extern "C" void alisb_phaser_destroy(alisb_phaser *p)
{
	if (!p) return;
	cudaSetDevice(p->e->g.device);
	cudaStreamSynchronize(p->e->stream);
	for (int f=0; f<6; f++) cudaFree(p->acc[f]);
	cudaFree(p->d_jobs); cudaFree(p->d_w); cudaFree(p->d_sums);
	delete p;
}

/* one updatePhaser call (reference phaser.c:94-110): every face and every frequency in one launch
 * (kernel phaser_accumulate), phase factors as kernel arguments.  Lists longer than
 * ALISB_PHASER_MAXF frequencies take one launch per ALISB_PHASER_MAXF. */
extern "C" int alisb_phaser_accumulate(alisb_phaser *p, const float *expE, const float *expH)
{
	alisb_engine *e = p->e;
	NvtxRange range("alisb:updatePhaser");
	DEV(e);
	SETTLE(e);
	PhaserParams pp;
	memset(&pp, 0, sizeof pp);
	pp.ny = e->ny; pp.pitch = e->pitch; pp.nfreq = p->nfreq;
	int nfaces = 0;
	for (int f=0; f<6; f++) if (p->acc[f] && p->cells[f]) {
		const int ax = f/2, c1 = (ax+1)%3, c2 = (ax+2)%3, k = nfaces++;
		pp.acc[k] = p->acc[f]; pp.B[k] = p->face[f];
		pp.E1[k] = e->F[0][0][c1]; pp.E2[k] = e->F[0][0][c2]; pp.H1[k] = e->F[1][0][c1]; pp.H2[k] = e->F[1][0][c2];
		pp.hstep[k] = ax == 0 ? 1 : ax == 1 ? (unsigned long long)e->pitch : (unsigned long long)e->plane;
	}
	pp.nfaces = nfaces;
	if (!nfaces) return ALISB_OK;
	for (int f0=0; f0<p->nfreq; f0+=ALISB_PHASER_MAXF) {
		const int nf = std::min(ALISB_PHASER_MAXF, p->nfreq - f0);
		pp.f0 = f0; pp.nf = nf;
		memcpy(pp.wE, expE + 2*(size_t)f0, 2*(size_t)nf*sizeof(float));
		memcpy(pp.wH, expH + 2*(size_t)f0, 2*(size_t)nf*sizeof(float));
		unsigned long long tot = 0;
		for (int k=0; k<nfaces; k++) { pp.start[k] = tot; tot += (unsigned long long)bcount(pp.B[k])*nf; }
		pp.start[nfaces] = tot;
		phaser_accumulate<<<div_up(tot, 256), 256, 0, e->stream>>>(pp);
		e->launches++;
	}
	CU(cudaGetLastError());
	return ALISB_OK;
}

extern "C" int alisb_phaser_read(alisb_phaser *p, int freq, int face, int which, float *host)
{
	if (face<0||face>5||which<0||which>3||freq<0||freq>=p->nfreq||!p->acc[face]) return fail(ALISB_EINVAL, "bad phaser id");
	alisb_engine *e = p->e;
	DEV(e);
	const size_t n = p->cells[face];
	std::vector<float2> h(n);
	CU(cudaStreamSynchronize(e->stream));
	CU(cudaMemcpy(h.data(), p->acc[face] + ((size_t)freq*4 + which)*n, n*sizeof(float2), cudaMemcpyDeviceToHost));
	const Box &B = p->face[face];
	float2 *o = (float2*)host;
	for (int a=0; a<B.na; a++) for (int b=0; b<B.nb; b++) for (int c=0; c<B.nc; c++)
		o[((size_t)a*B.nb + b)*B.nc + c] = h[((size_t)c*B.nb + b)*B.na + a];
	return ALISB_OK;
}

/* near-to-far-field surface sums: every job of an angular sweep in one launch (kernel surface_sums) */
extern "C" int alisb_phaser_surface_sums(alisb_phaser *p, int freq, const alisb_surface_job *jobs, int njobs, const float *weights, size_t nweights, double *out)
{
	NvtxRange range("alisb:phaser_surface_sums");
	if (!p || freq < 0 || freq >= p->nfreq || njobs < 0 || (njobs && (!jobs || !weights || !out))) return fail(ALISB_EINVAL, "bad surface-sum request");
	if (!njobs) return ALISB_OK;
	alisb_engine *e = p->e;
	DEV(e);
	SurfParams sp;
	memset(&sp, 0, sizeof sp);
	sp.freq = freq;
	for (int f=0; f<6; f++) {
		sp.acc[f] = p->acc[f];
		const Box &B = p->face[f];
		const int ax = f/2;
		if (p->acc[f]) { sp.nlo[f] = ax == 0 ? B.nb : B.na; sp.nhi[f] = ax == 2 ? B.nb : B.nc; }
	}
	std::vector<SurfJob> hj((size_t)njobs);
	for (int q=0; q<njobs; q++) {
		const alisb_surface_job &j = jobs[q];
		if (j.face < 0 || j.face > 5 || !p->acc[j.face] || j.which < 0 || j.which > 3) return fail(ALISB_EINVAL, "surface sum: face %d is not accumulated", j.face);
		if ((size_t)j.w_lo + sp.nlo[j.face] > nweights || (size_t)j.w_hi + sp.nhi[j.face] > nweights) return fail(ALISB_EINVAL, "surface sum: weight vector outside the buffer");
		hj[q].face = j.face; hj[q].which = j.which; hj[q].w_lo = j.w_lo; hj[q].w_hi = j.w_hi;
	}
	if ((size_t)njobs > p->jobs_cap) {
		cudaFree(p->d_jobs); cudaFree(p->d_sums); p->d_jobs = 0; p->d_sums = 0; p->jobs_cap = 0;
		const size_t cap = std::max((size_t)njobs, (size_t)1024);
		CU(cudaMalloc((void**)&p->d_jobs, cap*sizeof(SurfJob)));
		CU(cudaMalloc((void**)&p->d_sums, cap*2*sizeof(double)));
		p->jobs_cap = cap;
	}
	if (nweights > p->w_cap) {
		cudaFree(p->d_w); p->d_w = 0; p->w_cap = 0;
		const size_t cap = std::max(nweights, (size_t)4096);
		CU(cudaMalloc((void**)&p->d_w, cap*sizeof(float2)));
		p->w_cap = cap;
	}
	CU(cudaMemcpyAsync(p->d_jobs, hj.data(), (size_t)njobs*sizeof(SurfJob), cudaMemcpyHostToDevice, e->stream));
	CU(cudaMemcpyAsync(p->d_w, weights, nweights*sizeof(float2), cudaMemcpyHostToDevice, e->stream));
	surface_sums<<<(unsigned)njobs, 128, 0, e->stream>>>(sp, p->d_jobs, p->d_w, p->d_sums);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out, p->d_sums, (size_t)njobs*2*sizeof(double), cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	e->launches++;
	return ALISB_OK;
}

/* ---- instrumentation -------------------------------------------------------------------------- */
extern "C" uint64_t alisb_launch_count(alisb_engine *e) { return e->launches; }
extern "C" int alisb_set_profiling(alisb_engine *e, int on) { e->cfg_version++; e->profiling = on != 0; e->ev_used = 0; e->prof.clear(); return ALISB_OK; }
extern "C" int alisb_last_run_kernel_ms(alisb_engine *e, float *ms_total, int *launches)
{
	DEV(e);
	CU(cudaStreamSynchronize(e->stream));
	float tot = 0; int n = 0;
	for (const ProfRec &r : e->prof) { float ms = 0; CU(cudaEventElapsedTime(&ms, e->ev_pool[r.ev], e->ev_pool[r.ev+1])); tot += ms; n++; }
	if (ms_total) *ms_total = tot;
	if (launches) *launches = n;
	return ALISB_OK;
}

/* the dominant kernel = the volume sub-launch covering the most cells (the interior, no-PML
 * instantiation whenever the grid has an interior); bytes per launch = 36 B per cell
 * (read 3 F + 3 G, write 3 F: SURVEY 8d) for that launch's cells.                          */
extern "C" int alisb_profile_dominant(alisb_engine *e, float *ms_total, int *launches, double *bytes_per_launch)
{
	DEV(e);
	CU(cudaStreamSynchronize(e->stream));
	int best[2] = {-1, -1}; double cells[2] = {0, 0};
	for (int s=0; s<2; s++) for (size_t q=0; q<e->plan[s].size(); q++) {
		const SubBox &b = e->plan[s][q];
		const double c = (double)(b.a1 - b.a0 + 1)*(b.b1 - b.b0 + 1)*(b.c1 - b.c0 + 1);
		if (c > cells[s]) { cells[s] = c; best[s] = (int)q; }
	}
	/* launches of the largest region; when a pass is split for overlap (alisb_step_part) only the
	 * launches that carry at least half of the region count, and the bytes are those of the cells
	 * they actually processed */
	float tot = 0; int n = 0; double done = 0;
	for (const ProfRec &r : e->prof) if (r.sub == best[r.set] && r.cells >= 0.5*cells[r.set]) {
		float ms = 0; CU(cudaEventElapsedTime(&ms, e->ev_pool[r.ev], e->ev_pool[r.ev+1])); tot += ms; n++; done += r.cells;
	}
	if (ms_total) *ms_total = tot;
	if (launches) *launches = n;
	if (bytes_per_launch) *bytes_per_launch = n ? 36.0*done/n : 0;
	return ALISB_OK;
}
