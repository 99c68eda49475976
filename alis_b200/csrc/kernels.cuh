/*
 * kernels.cuh -- hand-written sm_100a kernels of the FDTD time-stepping path.
 *
 * Device layout (DESIGN.md "Data layout in HBM"): every scalar array is
 * [c][b][pitch] with x (a) fastest and z (c) slowest; element (a,b,c) lives at
 *     ((size_t)c*(ny+2) + b)*pitch + XO + a ,   XO = 31,
 * so the first updated cell a=1 of every row is 128-byte aligned and one warp reading
 * float4 per lane covers exactly four 128-byte lines.  z-planes are contiguous, which
 * makes the multi-GPU halo planes, k-faces and z-phasor faces plain 2-D blocks.
 *
 * Arithmetic: fp32, no FMA contraction (the translation unit is compiled with
 * -fmad=false), IEEE division, and the reference's expression order, so that results are
 * bit-identical to reference update.c (see the per-kernel citations).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ALISB_XO 31               /* array coordinate a sits at row offset XO + a */

struct Box { int a0, b0, c0, na, nb, nc; };

/* Dispersive objects handled INSIDE the volume kernels (template flag ADE): the E pass applies the
 * conductivity loss and subtracts the polarisation currents from the E it has just loaded, before
 * the curl (reference update.c:89-108); the H pass advances J/K from the final E it reads anyway
 * (updateJ, update.c:128-158).  The table lives in device memory (one per complex field set). */
#define ALISB_MAXADE 8
struct AdeObj {
	const float *fill[3];         /* dense box arrays, x fastest */
	float *J[ALISB_MAXPOLES][3], *K[ALISB_MAXPOLES][3];
	float a[3];
	float JJ[3][ALISB_MAXPOLES], JK[3][ALISB_MAXPOLES], JE[3][ALISB_MAXPOLES];
	float KJ[3][ALISB_MAXPOLES], KK[3][ALISB_MAXPOLES], KE[3][ALISB_MAXPOLES];
	int nj, nk, lossy;
	Box full;                     /* allocation box of fill/J/K                         */
	Box upd;                      /* full clipped to the updated cells 1..n              */
};
struct AdeTable { int n, pad; AdeObj o[ALISB_MAXADE]; };

struct VolParams {
	const AdeTable *ade;          /* ADE variants only */
	float       *F[3];            /* field set being updated (E or H), components x,y,z */
	const float *Fin[3];          /* where the old values are read: == F (in place) or the other buffer (double-buffered mode) */
	const float *G[3];            /* the other field set                                */
	const float *C[3];            /* inverse permittivity (E pass, EPS variant)         */
	const float *dtd[3];          /* dtdx (stored shifted by XO), dtdy, dtdz            */
	const float *P[3], *P1[3], *P2[3];   /* CPML tables per axis (x ones shifted by XO) */
	float       *psi[3][2][2];    /* [axis][side][which]                                */
	const float *psiz_in[2][2];   /* z slabs: where the old values are read ([side][which]): == psi[2] (in place) or the other buffer (double-buffered mode) */
	int slab[3][2][2];            /* [axis][side]{first,last} in array coordinates      */
	int xbase[2], xpitch[2];      /* x-slab psi arrays: slot = a - xbase, row pitch      */
	int nx, ny, nz, pitch;
	size_t plane;                 /* (ny+2)*pitch                                        */
};

/* one box of the region plan and the CTA grid that covers it */
struct VolRegion {
	int a_first, a_last;          /* x range updated (a_first = 1 mod 4); a warp covers 4*XL cells from a_first */
	int nseg;                     /* number of warp segments covering [a_first, a_last]  */
	int b_first, b_last;          /* rows updated                                        */
	int c_first, c_last;          /* planes updated                                      */
	int cchunk;                   /* planes per z-chunk                                  */
	int pm, narrow;               /* PML bitmask of the box; 1 = XL=8 mapping (x strips) */
	int gx, gy, gz;               /* CTA grid of the box                                 */
	unsigned cta0;                /* first linear CTA of the box inside a merged launch  */
};
#define ALISB_MAXREG 28
struct MultiParams { VolParams v; int nreg; VolRegion r[ALISB_MAXREG]; };

static __device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4*>(p); }
static __device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
/* streaming (evict-first) variants for data that is touched once per pass */
template <bool STREAM> static __device__ __forceinline__ float4 ld4s(const float *p) { return STREAM ? __ldcs(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p); }
template <bool STREAM> static __device__ __forceinline__ void st4s(float *p, float4 v) { if (STREAM) __stcs(reinterpret_cast<float4*>(p), v); else *reinterpret_cast<float4*>(p) = v; }

/*
 * Fused Yee curl + CPML volume kernel: reference updateF (update.c:87-125) without the
 * ADE pre-pass, i.e. the curl of update.c:109-118 followed by the six psi slabs of
 * update.c:48-84,119-124 applied in the reference's order, in ONE pass over memory.
 *
 *   S = +1 : E pass (forward differences of H),  S = -1 : H pass (backward differences of E)
 *
 * Thread mapping: one warp = one x-row segment of 128 cells (float4 per lane), blockDim.y
 * rows per CTA, and the CTA marches along z (2.5-D blocking) in direction -S so that the
 * plane c+S needed by the z-differences is the one carried in registers from the previous
 * iteration.  The y-neighbour row is fetched with a second coalesced load that hits L1
 * (the warp above/below in the same CTA loads the same line in the same iteration), the
 * x-neighbour comes from a warp shuffle.
 */
/* XL = lanes of a warp along x: 32 (a warp is one 128-cell row segment) or 8 (a warp is a 32-cell
 * wide, 4-row patch: used for the narrow x-PML strips so that all lanes have work). */
template <int S, bool EPS, int PM, int XL, bool ADE>  /* PM: bit 0/1/2 = this box may touch x / y / z PML slabs */
static __device__ __forceinline__ void yee_body(const VolParams &p, const VolRegion &r, const int bix, const int biy, const int biz)
{
	/* the object table is read by every thread in every plane: the CTA keeps its own copy in shared memory (one cooperative
	 * fetch instead of chains of dependent global loads, and loads the compiler may keep across the field stores) */
	const AdeTable *adeT = 0;
	if (ADE) {
		__shared__ __align__(16) unsigned long long ade_words[sizeof(AdeTable)/8];
		const int tid = threadIdx.y*blockDim.x + threadIdx.x, nth = blockDim.x*blockDim.y;
		const int nobj = min(max(p.ade->n, 0), ALISB_MAXADE);
		const int words = (int)((offsetof(AdeTable, o) + (size_t)nobj*sizeof(AdeObj))/8);
		const unsigned long long *src = reinterpret_cast<const unsigned long long*>(p.ade);
		for (int q = tid; q < words; q += nth) ade_words[q] = src[q];
		__syncthreads();
		adeT = reinterpret_cast<const AdeTable*>(ade_words);
	}
	constexpr int RW = 32/XL;                               /* rows per warp */
	const int lane = threadIdx.x & 31;
	const int lx = lane % XL;                               /* lane position along x */
	const int seg = bix*(blockDim.x >> 5) + (threadIdx.x >> 5);            /* (4*XL)-cell segment within the box */
	const int a0 = r.a_first + 4*(seg*XL + lx);
	const int bw = r.b_first + (biy*blockDim.y + threadIdx.y)*RW;          /* first row of this warp */
	const int b  = bw + lane/XL;
	if (bw > r.b_last || seg >= r.nseg) return;             /* whole warp leaves */
	const bool rowok = b <= r.b_last;
	const bool act = rowok && a0 <= r.a_last;               /* owns at least one cell of this box */
	const bool ldg = rowok && a0 <= r.a_last + 1;           /* G quad holds data the lane before needs (a_last+1) */
	const bool full = rowok && a0 + 3 <= r.a_last;

	const int cf = r.c_first + biz*r.cchunk;
	const int cl = min(cf + r.cchunk - 1, r.c_last);
	if (cf > cl) return;
	const int nsteps = cl - cf + 1;
	int c = (S > 0) ? cl : cf;

	const size_t row = (size_t)b*p.pitch + ALISB_XO + a0;    /* offset inside a plane */
	const ptrdiff_t yn = (ptrdiff_t)S*p.pitch;               /* to row b+S            */

	float *F0 = p.F[0], *F1 = p.F[1], *F2 = p.F[2];
	const float *I0 = p.Fin[0], *I1 = p.Fin[1], *I2 = p.Fin[2];
	const float *__restrict__ G0 = p.G[0], *__restrict__ G1 = p.G[1], *__restrict__ G2 = p.G[2];

	const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
	const float4 dx4 = act ? ld4(p.dtd[0] + ALISB_XO + a0) : zero4;
	const float  dy  = rowok ? __ldg(p.dtd[1] + b) : 0.f;

	/* loop-invariant slab membership */
	int ys = -1, xs = -1;
	float yP = 0.f, yP1 = 0.f, yP2 = 0.f;
	float4 xP = zero4, xP1 = zero4, xP2 = zero4;
	unsigned xmask = 0;
	constexpr bool STREAM = false;              /* evict-first hints measured slower (profiles/sweep_r01) */
	if (PM & 2) {
		if (b >= p.slab[1][0][0] && b <= p.slab[1][0][1]) ys = 0;
		else if (b >= p.slab[1][1][0] && b <= p.slab[1][1][1]) ys = 1;
		if (!rowok) ys = -1;
		if (ys >= 0) { yP = __ldg(p.P[1] + b); yP1 = __ldg(p.P1[1] + b); yP2 = __ldg(p.P2[1] + b); }
	}
	if (PM & 1) {
		if (act) {
			if (a0 + 3 >= p.slab[0][0][0] && a0 <= p.slab[0][0][1]) xs = 0;
			else if (a0 + 3 >= p.slab[0][1][0] && a0 <= p.slab[0][1][1]) xs = 1;
			if (xs >= 0) {
				#pragma unroll
				for (int e = 0; e < 4; e++) if (a0 + e >= p.slab[0][xs][0] && a0 + e <= p.slab[0][xs][1]) xmask |= 1u << e;
				xP = ld4(p.P[0] + ALISB_XO + a0); xP1 = ld4(p.P1[0] + ALISB_XO + a0); xP2 = ld4(p.P2[0] + ALISB_XO + a0);
			}
		}
	}

	/* plane c+S: carried */
	float4 gxp = zero4, gyp = zero4;
	if (ldg) {
		const size_t o = (size_t)(c + S)*p.plane + row;
		gxp = ld4(G0 + o); gyp = ld4(G1 + o);
	}

	for (int it = 0; it < nsteps; ++it, c -= S) {
		const size_t o = (size_t)c*p.plane + row;
		float4 fx = zero4, fy = zero4, fz = zero4, gx = zero4, gy = zero4, gz = zero4, gxn = zero4, gzn = zero4;
		float4 cx = zero4, cy = zero4, cz = zero4;
		if (ldg) { gx = ld4(G0 + o); gy = ld4s<STREAM>(G1 + o); gz = ld4(G2 + o); }
		if (act) {
			gxn = ld4(G0 + o + yn); gzn = ld4(G2 + o + yn);
			fx = ld4s<STREAM>(I0 + o); fy = ld4s<STREAM>(I1 + o); fz = ld4s<STREAM>(I2 + o);
			if (EPS) { cx = ld4(p.C[0] + o); cy = ld4(p.C[1] + o); cz = ld4(p.C[2] + o); }
		}
		/* x-neighbour element (a0+4 for S=+1, a0-1 for S=-1) from the adjacent lane */
		float gye, gze;
		if (S > 0) {
			gye = __shfl_down_sync(0xffffffffu, gy.x, 1);
			gze = __shfl_down_sync(0xffffffffu, gz.x, 1);
			if (lx == XL-1 && act) { gye = G1[o + 4]; gze = G2[o + 4]; }
		} else {
			gye = __shfl_up_sync(0xffffffffu, gy.w, 1);
			gze = __shfl_up_sync(0xffffffffu, gz.w, 1);
			if (lx == 0 && act) { gye = *(G1 + o - 1); gze = *(G2 + o - 1); }
		}
		const float dz = __ldg(p.dtd[2] + c);

		/* 5-wide windows along x: self(e) and neighbour(e) = element a+S */
		float gy5[5], gz5[5];
		if (S > 0) { gy5[0]=gy.x; gy5[1]=gy.y; gy5[2]=gy.z; gy5[3]=gy.w; gy5[4]=gye;  gz5[0]=gz.x; gz5[1]=gz.y; gz5[2]=gz.z; gz5[3]=gz.w; gz5[4]=gze; }
		else       { gy5[0]=gye;  gy5[1]=gy.x; gy5[2]=gy.y; gy5[3]=gy.z; gy5[4]=gy.w; gz5[0]=gze;  gz5[1]=gz.x; gz5[2]=gz.y; gz5[3]=gz.z; gz5[4]=gz.w; }
		constexpr int SELF = (S > 0) ? 0 : 1, NBR = (S > 0) ? 1 : 0;

		float Fx[4] = {fx.x, fx.y, fx.z, fx.w}, Fy[4] = {fy.x, fy.y, fy.z, fy.w}, Fz[4] = {fz.x, fz.y, fz.z, fz.w};
		const float Gx[4] = {gx.x, gx.y, gx.z, gx.w}, Gy[4] = {gy.x, gy.y, gy.z, gy.w}, Gz[4] = {gz.x, gz.y, gz.z, gz.w};
		const float Gxn[4] = {gxn.x, gxn.y, gxn.z, gxn.w}, Gzn[4] = {gzn.x, gzn.y, gzn.z, gzn.w};
		const float Gxp[4] = {gxp.x, gxp.y, gxp.z, gxp.w}, Gyp[4] = {gyp.x, gyp.y, gyp.z, gyp.w};
		const float Dx[4] = {dx4.x, dx4.y, dx4.z, dx4.w};
		const float Cx[4] = {cx.x, cx.y, cx.z, cx.w}, Cy[4] = {cy.x, cy.y, cy.z, cy.w}, Cz[4] = {cz.x, cz.y, cz.z, cz.w};

		if (ADE && S > 0 && act) {
			/* update.c:89-108 on the old E in registers: every lossy object's factor, then every object's
			 * currents, objects and poles in list order (cells are independent, so this is the reference's
			 * sequence of whole-box sweeps seen from one cell) */
			const AdeTable &T = *adeT;
			for (int n = 0; n < T.n; n++) {
				const AdeObj &ob = T.o[n];
				if (!ob.lossy || c < ob.upd.c0 || c >= ob.upd.c0 + ob.upd.nc || b < ob.upd.b0 || b >= ob.upd.b0 + ob.upd.nb) continue;
				const size_t jr = ((size_t)(c - ob.full.c0)*ob.full.nb + (b - ob.full.b0))*ob.full.na;
				#pragma unroll
				for (int e = 0; e < 4; e++) {
					const int a = a0 + e;
					if (a < ob.upd.a0 || a >= ob.upd.a0 + ob.upd.na || a > r.a_last) continue;
					const size_t j = jr + (a - ob.full.a0);
					const float f0 = ob.fill[0][j], f1 = ob.fill[1][j], f2 = ob.fill[2][j];
					Fx[e] = Fx[e] * ((1 - ob.a[0]*f0) / (1 + ob.a[0]*f0));       /* update.c:94-96 */
					Fy[e] = Fy[e] * ((1 - ob.a[1]*f1) / (1 + ob.a[1]*f1));
					Fz[e] = Fz[e] * ((1 - ob.a[2]*f2) / (1 + ob.a[2]*f2));
				}
			}
			for (int n = 0; n < T.n; n++) {
				const AdeObj &ob = T.o[n];
				if (!ob.nk || c < ob.upd.c0 || c >= ob.upd.c0 + ob.upd.nc || b < ob.upd.b0 || b >= ob.upd.b0 + ob.upd.nb) continue;
				const size_t jr = ((size_t)(c - ob.full.c0)*ob.full.nb + (b - ob.full.b0))*ob.full.na;
				#pragma unroll
				for (int e = 0; e < 4; e++) {
					const int a = a0 + e;
					if (a < ob.upd.a0 || a >= ob.upd.a0 + ob.upd.na || a > r.a_last) continue;
					const size_t j = jr + (a - ob.full.a0);
					for (int s = 0; s < ob.nk; s++) { Fx[e] -= ob.J[s][0][j]; Fy[e] -= ob.J[s][1][j]; Fz[e] -= ob.J[s][2][j]; }   /* update.c:104-106 */
				}
			}
		}

		/* update.c:110-112 / :115-117 */
		#pragma unroll
		for (int e = 0; e < 4; e++) {
			const float ux = dz*(Gy[e] - Gyp[e]) + dy*(Gzn[e] - Gz[e]);
			const float uy = Dx[e]*(gz5[e+SELF] - gz5[e+NBR]) + dz*(Gxp[e] - Gx[e]);
			const float uz = dy*(Gx[e] - Gxn[e]) + Dx[e]*(gy5[e+NBR] - gy5[e+SELF]);
			if (EPS) { Fx[e] += ux*Cx[e]; Fy[e] += uy*Cy[e]; Fz[e] += uz*Cz[e]; }
			else     { Fx[e] += ux;       Fy[e] += uy;       Fz[e] += uz; }
		}

		if (PM) {
			int zs = -1;
			if (PM & 4) {
				if (c >= p.slab[2][0][0] && c <= p.slab[2][0][1]) zs = 0;
				else if (c >= p.slab[2][1][0] && c <= p.slab[2][1][1]) zs = 1;
			}
			if (act && (xs >= 0 || ys >= 0 || zs >= 0)) {
				/* psi contributions; slab order x-min,y-min,z-min,x-max,y-max,z-max (update.c:119-124) */
				float xq0[4] = {0,0,0,0}, xq1[4] = {0,0,0,0};   /* x slab: psi0 -> Fy, psi1 -> Fz */
				float yq0[4] = {0,0,0,0}, yq1[4] = {0,0,0,0};   /* y slab: psi0 -> Fz, psi1 -> Fx */
				float zq0[4] = {0,0,0,0}, zq1[4] = {0,0,0,0};   /* z slab: psi0 -> Fx, psi1 -> Fy */
				if ((PM & 1) && xs >= 0) {
					const size_t po = ((size_t)c*(p.ny + 2) + b)*p.xpitch[xs] + (a0 - p.xbase[xs]);
					float *q0p = p.psi[0][xs][0] + po, *q1p = p.psi[0][xs][1] + po;
					float4 q0 = ld4(q0p), q1 = ld4(q1p);
					float v0[4] = {q0.x, q0.y, q0.z, q0.w}, v1[4] = {q1.x, q1.y, q1.z, q1.w};
					const float Pq[4] = {xP.x, xP.y, xP.z, xP.w}, Pa[4] = {xP1.x, xP1.y, xP1.z, xP1.w}, Pb[4] = {xP2.x, xP2.y, xP2.z, xP2.w};
					#pragma unroll
					for (int e = 0; e < 4; e++) if (xmask & (1u << e)) {
						v0[e] = Pq[e]*v0[e] + Pa[e]*(gz5[e+SELF] - gz5[e+NBR]);   /* update.c:51 */
						v1[e] = Pq[e]*v1[e] + Pb[e]*(gy5[e+SELF] - gy5[e+NBR]);   /* update.c:52 */
						xq0[e] = v0[e]; xq1[e] = v1[e];
					}
					st4(q0p, make_float4(v0[0], v0[1], v0[2], v0[3]));
					st4(q1p, make_float4(v1[0], v1[1], v1[2], v1[3]));
				}
				if ((PM & 2) && ys >= 0) {
					const int nys = p.slab[1][ys][1] - p.slab[1][ys][0] + 1;
					const size_t po = ((size_t)c*nys + (b - p.slab[1][ys][0]))*p.pitch + ALISB_XO + a0;
					float *q0p = p.psi[1][ys][0] + po, *q1p = p.psi[1][ys][1] + po;
					float4 q0 = ld4(q0p), q1 = ld4(q1p);
					float v0[4] = {q0.x, q0.y, q0.z, q0.w}, v1[4] = {q1.x, q1.y, q1.z, q1.w};
					#pragma unroll
					for (int e = 0; e < 4; e++) {
						v0[e] = yP*v0[e] + yP1*(Gx[e] - Gxn[e]);                 /* update.c:64 */
						v1[e] = yP*v1[e] + yP2*(Gz[e] - Gzn[e]);                 /* update.c:65 */
						yq0[e] = v0[e]; yq1[e] = v1[e];
					}
					st4(q0p, make_float4(v0[0], v0[1], v0[2], v0[3]));
					st4(q1p, make_float4(v1[0], v1[1], v1[2], v1[3]));
				}
				if ((PM & 4) && zs >= 0) {
					const float zP = __ldg(p.P[2] + c), zP1 = __ldg(p.P1[2] + c), zP2 = __ldg(p.P2[2] + c);
					const size_t po = ((size_t)(c - p.slab[2][zs][0])*(p.ny + 2) + b)*p.pitch + ALISB_XO + a0;
					float *q0p = p.psi[2][zs][0] + po, *q1p = p.psi[2][zs][1] + po;
					float4 q0 = ld4(p.psiz_in[zs][0] + po), q1 = ld4(p.psiz_in[zs][1] + po);
					float v0[4] = {q0.x, q0.y, q0.z, q0.w}, v1[4] = {q1.x, q1.y, q1.z, q1.w};
					#pragma unroll
					for (int e = 0; e < 4; e++) {
						v0[e] = zP*v0[e] + zP1*(Gy[e] - Gyp[e]);                 /* update.c:77 */
						v1[e] = zP*v1[e] + zP2*(Gx[e] - Gxp[e]);                 /* update.c:78 */
						zq0[e] = v0[e]; zq1[e] = v1[e];
					}
					st4(q0p, make_float4(v0[0], v0[1], v0[2], v0[3]));
					st4(q1p, make_float4(v1[0], v1[1], v1[2], v1[3]));
				}
				#pragma unroll
				for (int e = 0; e < 4; e++) {
					const bool xin = (xmask >> e) & 1u;
					#pragma unroll
					for (int side = 0; side < 2; side++) {
						if ((PM & 1) && xs == side && xin) {
							if (EPS) { Fy[e] += xq0[e]*Cy[e]; Fz[e] += xq1[e]*Cz[e]; } else { Fy[e] += xq0[e]; Fz[e] += xq1[e]; }
						}
						if ((PM & 2) && ys == side) {
							if (EPS) { Fz[e] += yq0[e]*Cz[e]; Fx[e] += yq1[e]*Cx[e]; } else { Fz[e] += yq0[e]; Fx[e] += yq1[e]; }
						}
						if ((PM & 4) && zs == side) {
							if (EPS) { Fx[e] += zq0[e]*Cx[e]; Fy[e] += zq1[e]*Cy[e]; } else { Fx[e] += zq0[e]; Fy[e] += zq1[e]; }
						}
					}
				}
			}
		}

		if (full) {
			st4s<STREAM>(F0 + o, make_float4(Fx[0], Fx[1], Fx[2], Fx[3]));
			st4s<STREAM>(F1 + o, make_float4(Fy[0], Fy[1], Fy[2], Fy[3]));
			st4s<STREAM>(F2 + o, make_float4(Fz[0], Fz[1], Fz[2], Fz[3]));
		} else if (act) {
			#pragma unroll
			for (int e = 0; e < 4; e++) if (a0 + e <= r.a_last) { F0[o + e] = Fx[e]; F1[o + e] = Fy[e]; F2[o + e] = Fz[e]; }
		}
		if (ADE && S < 0 && act) {
			/* updateJ (update.c:128-158) from the final E of this thread's own cells, which the H pass has
			 * in registers as G: Drude poles, then critical-point / Lorentz pairs (K uses the new J) */
			const AdeTable &T = *adeT;
			for (int n = 0; n < T.n; n++) {
				const AdeObj &ob = T.o[n];
				if (!ob.nk || c < ob.upd.c0 || c >= ob.upd.c0 + ob.upd.nc || b < ob.upd.b0 || b >= ob.upd.b0 + ob.upd.nb) continue;
				const size_t jr = ((size_t)(c - ob.full.c0)*ob.full.nb + (b - ob.full.b0))*ob.full.na;
				bool in[4];
				#pragma unroll
				for (int e = 0; e < 4; e++) { const int a = a0 + e; in[e] = !(a < ob.upd.a0 || a >= ob.upd.a0 + ob.upd.na || a > r.a_last); }
				if (!(in[0] || in[1] || in[2] || in[3])) continue;
				const size_t j0 = jr + (a0 - ob.full.a0);        /* cell e of the quad sits at j0 + e (only dereferenced where in[e]) */
				/* per component: every load of the quad's currents first (they are independent; issued one behind the other
				 * they were a chain of dependent global round trips: the stores between them may alias), then the arithmetic
				 * of update.c:135-142 pole by pole, then the stores */
				#pragma unroll
				for (int k = 0; k < 3; k++) {
					float Jv[ALISB_MAXPOLES][4], Kv[ALISB_MAXPOLES][4], fv[4];
					#pragma unroll
					for (int e = 0; e < 4; e++) fv[e] = in[e] ? ob.fill[k][j0 + e] : 0.f;
					#pragma unroll
					for (int s = 0; s < ALISB_MAXPOLES; s++) {
						#pragma unroll
						for (int e = 0; e < 4; e++) {
							Jv[s][e] = (s < ob.nk && in[e]) ? ob.J[s][k][j0 + e] : 0.f;
							Kv[s][e] = (s < ob.nk && s >= ob.nj && in[e]) ? ob.K[s][k][j0 + e] : 0.f;
						}
					}
					#pragma unroll
					for (int e = 0; e < 4; e++) {
						if (!in[e]) continue;
						const float Fv = k == 0 ? Gx[e] : k == 1 ? Gy[e] : Gz[e], f = fv[e];
						#pragma unroll
						for (int s = 0; s < ALISB_MAXPOLES; s++) {
							if (s >= ob.nk) continue;
							if (s < ob.nj) {
								ob.J[s][k][j0 + e] = ob.JJ[k][s]*Jv[s][e] + ob.JE[k][s]*f*Fv;                              /* update.c:135 */
							} else {
								const float Kold = Kv[s][e];
								const float Jn = ob.JJ[k][s]*Jv[s][e] + ob.JK[k][s]*Kold + ob.JE[k][s]*f*Fv;            /* update.c:140 */
								ob.J[s][k][j0 + e] = Jn;
								ob.K[s][k][j0 + e] = ob.KJ[k][s]*Jn + ob.KK[k][s]*Kold + ob.KE[k][s]*f*Fv;               /* update.c:142 */
							}
						}
					}
				}
			}
		}
		gxp = gx; gyp = gy;
	}
}

/*
 * Fused E+H step on a slab-free, fix-up-free box (temporal blocking of update.c:275-341 where
 * nothing but the two curls happens): E' = E + curl H,  H' = H + curl E'  in ONE pass over memory,
 * 48 B per cell-step instead of 72.  Out of place (reads Ei/Hi, writes Eo/Ho), because a CTA
 * recomputes E' on a one-cell halo (row b-1, column a-1, plane c-1 of its tile) from old values
 * that the owning CTA may already have replaced.  Same expressions and order as yee_body<+1> and
 * yee_body<-1>, so the results are bit-identical to the two-pass path.
 *
 * CTA = 32 x (ROWS+1) threads: warp wy=0 is the halo row b0-1 (computes E', stores nothing), warps
 * 1..ROWS own rows b0..b0+ROWS-1; a warp is a 128-cell row segment (float4 per lane).  The CTA marches DOWN in
 * z: at plane c it forms E'(c) from H(c), H(c+1) [carried], then H'(c+1) from the carried E'(c+1)
 * and the fresh E'(c).  E'(row b-1) travels through shared memory, E'(a-1) through a shuffle; lane 0
 * recomputes the single cell a0-1 with scalar loads.
 */
struct FusedParams {
	const float *Ei[3], *Hi[3];
	float *Eo[3], *Ho[3];
	const float *dtdE[3], *dtdH[3];   /* x tables shifted by XO                             */
	const float *C[3];                /* inverse permittivity (EPS variant), full-grid arrays */
	int ny, pitch; size_t plane;
	int a_first, a_last, nseg;        /* a_first = 1 mod 4, (a_last - a_first + 1) % 4 == 0  */
	int b_first, b_last, c_first, c_last, cchunk;
	/* PMZ variant: planes of the box may lie in z CPML slabs (update.c:74-84 inside the fused sweep) */
	const float *PzE[3], *PzH[3];     /* z tables P, P1, P2 of the E and of the H set         */
	const float *psiEi[2][2], *psiHi[2][2];   /* [side][which] z-slab psi arrays: old values ...  */
	float *psiEo[2][2], *psiHo[2][2];         /* ... and where the new ones go (the other buffer) */
	int zslabE[2][2], zslabH[2][2];   /* [side]{first,last} plane of the slab (first > last: none) */
};

/* cp.async staging: every global load of the kernel is issued two planes ahead into shared memory
 * (no registers are held while the data is in flight), three stages deep */
static __device__ __forceinline__ void cp_async16(void *smem, const void *g)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(g) : "memory");
}
static __device__ __forceinline__ void cp_async4(void *smem, const void *g)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s), "l"(g) : "memory");
}
static __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> static __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

/* vacuum: 3 stages (loads two planes ahead); with material arrays the ring also carries the three
 * inverse-permittivity rows and is 2 stages deep (one plane ahead) so that it still fits */
template <int ROWS, bool EPS> struct FusedStage {      /* ROWS owned rows + the halo row below them */
	float4 e[3][ROWS+1][32];       /* E x,y,z: rows b0-1 .. b0+ROWS-1                     */
	float4 hy[ROWS+1][32];
	float4 hx[ROWS+2][32], hz[ROWS+2][32];   /* rows b0-1 .. b0+ROWS (row b+1 feeds the y differences) */
	float4 c[EPS ? 3 : 1][EPS ? ROWS+1 : 1][EPS ? 32 : 1];   /* inverse permittivity x,y,z of the same rows */
	float  edge[ROWS+1][12];       /* per warp: lane 0's cell a0-1 {Ey,Ez,Hz,Hy,Hx,Hx(b+1)}, lane 31's cell a0+128 {Hy,Hz}, lane 0's {Cy,Cz} */
};
template <int ROWS, bool EPS> struct FusedSmem {
	static constexpr int NST = EPS ? 2 : 3;
	FusedStage<ROWS,EPS> st[NST];
	float4 sx[2][ROWS+1][32], sz[2][ROWS+1][32];   /* E'x, E'z rows of a plane, for the row above */
};

template <int ROWS, bool EPS, bool PMZ>
__global__ void __launch_bounds__(32*(ROWS+1), ROWS == 8 ? 2 : 1)
yee_fused(const __grid_constant__ FusedParams p)
{
	extern __shared__ __align__(16) unsigned char fused_smem_raw[];
	FusedSmem<ROWS,EPS> &sm = *reinterpret_cast<FusedSmem<ROWS,EPS>*>(fused_smem_raw);
	constexpr int NST = FusedSmem<ROWS,EPS>::NST, AHEAD = NST - 1;     /* planes in flight */
	const int lane = threadIdx.x, wy = threadIdx.y;
	const int a0 = p.a_first + 4*((int)blockIdx.x*32 + lane);
	const int b  = p.b_first - 1 + (int)blockIdx.y*ROWS + wy;
	const bool rowok = b <= p.b_last;
	const bool act = rowok && a0 <= p.a_last;              /* forms E' for its quad                  */
	const bool ldg = rowok && a0 <= p.a_last + 1;          /* its H quad feeds the lane before       */
	const bool own = act && wy > 0;                        /* stores E', forms and stores H'         */
	const bool xh  = act && lane == 0;                     /* recomputes the cell a0-1               */
	const bool xt  = act && lane == 31;                    /* needs H of the cell a0+128             */
	const bool top = wy == ROWS || b == p.b_last;             /* nobody stages row b+1 for this warp    */

	const int cf = p.c_first + (int)blockIdx.z*p.cchunk;
	const int cl = min(cf + p.cchunk - 1, p.c_last);
	if (cf > cl) return;                                    /* whole CTA */

	const size_t row = (size_t)b*p.pitch + ALISB_XO + a0;
	const ptrdiff_t yp = p.pitch;
	const float *E0 = p.Ei[0], *E1 = p.Ei[1], *E2 = p.Ei[2];
	const float *H0 = p.Hi[0], *H1 = p.Hi[1], *H2 = p.Hi[2];

	/* all loads of plane c, asynchronously, into stage s */
	auto stage_plane = [&](int c, int s) {
		FusedStage<ROWS,EPS> &S = sm.st[s];
		const size_t o = (size_t)c*p.plane + row;
		if (ldg) { cp_async16(&S.hx[wy][lane], H0 + o); cp_async16(&S.hy[wy][lane], H1 + o); cp_async16(&S.hz[wy][lane], H2 + o); }
		if (act) {
			cp_async16(&S.e[0][wy][lane], E0 + o); cp_async16(&S.e[1][wy][lane], E1 + o); cp_async16(&S.e[2][wy][lane], E2 + o);
			if (top) { cp_async16(&S.hx[wy + 1][lane], H0 + o + yp); cp_async16(&S.hz[wy + 1][lane], H2 + o + yp); }
			if (EPS) { cp_async16(&S.c[0][wy][lane], p.C[0] + o); cp_async16(&S.c[1][wy][lane], p.C[1] + o); cp_async16(&S.c[2][wy][lane], p.C[2] + o); }
		}
		if (xh) {
			cp_async4(&S.edge[wy][0], E1 + o - 1); cp_async4(&S.edge[wy][1], E2 + o - 1);
			cp_async4(&S.edge[wy][2], H2 + o - 1); cp_async4(&S.edge[wy][3], H1 + o - 1);
			cp_async4(&S.edge[wy][4], H0 + o - 1); cp_async4(&S.edge[wy][5], H0 + o - 1 + yp);
		}
		if (xt) { cp_async4(&S.edge[wy][6], H1 + o + 4); cp_async4(&S.edge[wy][7], H2 + o + 4); }
		if (EPS && xh) { cp_async4(&S.edge[wy][8], p.C[1] + o - 1); cp_async4(&S.edge[wy][9], p.C[2] + o - 1); }
	};

	const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
	const float4 dxe4 = act ? ld4(p.dtdE[0] + ALISB_XO + a0) : zero4;
	const float4 dxh4 = act ? ld4(p.dtdH[0] + ALISB_XO + a0) : zero4;
	const float dye = rowok ? __ldg(p.dtdE[1] + b) : 0.f, dyh = rowok ? __ldg(p.dtdH[1] + b) : 0.f;
	const float dxeh = xh ? __ldg(p.dtdE[0] + ALISB_XO + a0 - 1) : 0.f;    /* E table at the cell a0-1 */
	const float DxE[4] = {dxe4.x, dxe4.y, dxe4.z, dxe4.w}, DxH[4] = {dxh4.x, dxh4.y, dxh4.z, dxh4.w};

	/* pipeline prologue: the first AHEAD planes */
	#pragma unroll
	for (int q = 0; q < AHEAD; q++) {
		if (cl - q >= cf - 1) stage_plane(cl - q, q);
		cp_async_commit();
	}

	/* carried state of plane c+1 */
	float4 hxp = zero4, hyp = zero4, hzp = zero4;           /* H(c+1)                                  */
	float4 exp_ = zero4, eyp = zero4, ezp = zero4;          /* E'(c+1)                                 */
	float eym1 = 0.f, ezm1 = 0.f;                           /* E'y, E'z of the cell a0-1 at plane c+1 (lane 0) */
	float hxp1 = 0.f;                                       /* Hx of the cell a0-1 at plane c+1 (lane 0)       */
	if (ldg) {
		const size_t o = (size_t)(cl + 1)*p.plane + row;
		hxp = ld4(H0 + o); hyp = ld4(H1 + o);
		if (xh) hxp1 = H0[o - 1];
	}
	int c = cl;
	/* the z coefficients are fetched one iteration ahead: a dependent global load at its point of use was the
	 * largest single stall of this kernel (ncu source page, profiles/r02_ncu_fused_v2_notes.md) */
	float dze_next = __ldg(p.dtdE[2] + cl), dzh_next = __ldg(p.dtdH[2] + cl + 1);
	for (int it = 0; c >= cf - 1; ++it, --c) {
		const size_t o = (size_t)c*p.plane + row;
		const float dze = dze_next, dzh = dzh_next;
		if (c > cf - 1) { dze_next = __ldg(p.dtdE[2] + c - 1); dzh_next = __ldg(p.dtdH[2] + c); }
		/* z CPML slabs (PMZ): the psi values of plane c (E set) and of plane c+1 (H set) are requested before the
		 * barrier and used after the curls; only slab planes pay for them */
		int zsE = -1, zsH = -1;
		float4 qe0 = zero4, qe1 = zero4, qh0 = zero4, qh1 = zero4;
		float qeh = 0.f;
		size_t pe = 0, ph = 0;
		if (PMZ) {
			if (c >= p.zslabE[0][0] && c <= p.zslabE[0][1]) zsE = 0; else if (c >= p.zslabE[1][0] && c <= p.zslabE[1][1]) zsE = 1;
			if (it > 0) { if (c + 1 >= p.zslabH[0][0] && c + 1 <= p.zslabH[0][1]) zsH = 0; else if (c + 1 >= p.zslabH[1][0] && c + 1 <= p.zslabH[1][1]) zsH = 1; }
			if (zsE >= 0 && act) {
				pe = (size_t)(c - p.zslabE[zsE][0])*p.plane + row;
				qe0 = ld4(p.psiEi[zsE][0] + pe); qe1 = ld4(p.psiEi[zsE][1] + pe);
				if (xh) qeh = p.psiEi[zsE][1][pe - 1];
			}
			if (zsH >= 0 && own) {
				ph = (size_t)(c + 1 - p.zslabH[zsH][0])*p.plane + row;
				qh0 = ld4(p.psiHi[zsH][0] + ph); qh1 = ld4(p.psiHi[zsH][1] + ph);
			}
		}
		cp_async_wait<AHEAD - 1>();                         /* this thread's copies of plane c have landed */
		__syncthreads();                                    /* ... and everybody else's; iteration it-1 is over everywhere */
		{                                                   /* plane c-AHEAD goes into the stage iteration it-1 used */
			if (c - AHEAD >= cf - 1) stage_plane(c - AHEAD, (it + AHEAD) % NST);
			cp_async_commit();
		}
		const FusedStage<ROWS,EPS> &S = sm.st[it % NST];
		float4 cx4 = zero4, cy4 = zero4, cz4 = zero4;
		float cyh = 0.f, czh = 0.f;
		if (EPS && act) { cx4 = S.c[0][wy][lane]; cy4 = S.c[1][wy][lane]; cz4 = S.c[2][wy][lane]; if (xh) { cyh = S.edge[wy][8]; czh = S.edge[wy][9]; } }
		float4 ex = zero4, ey = zero4, ez = zero4, hx = zero4, hy = zero4, hz = zero4, hxn = zero4, hzn = zero4;
		if (ldg) { hx = S.hx[wy][lane]; hy = S.hy[wy][lane]; hz = S.hz[wy][lane]; }
		if (act) {
			hxn = S.hx[wy + 1][lane]; hzn = S.hz[wy + 1][lane];
			ex = S.e[0][wy][lane]; ey = S.e[1][wy][lane]; ez = S.e[2][wy][lane];
		}
		float hye = __shfl_down_sync(0xffffffffu, hy.x, 1), hze = __shfl_down_sync(0xffffffffu, hz.x, 1);
		if (xt) { hye = S.edge[wy][6]; hze = S.edge[wy][7]; }
		/* E'x, E'z of row b-1 at plane c+1, written by the warp below during iteration it-1 */
		float4 exm = zero4, ezm = zero4;
		if (wy > 0 && it > 0) { exm = sm.sx[(it - 1) & 1][wy - 1][lane]; ezm = sm.sz[(it - 1) & 1][wy - 1][lane]; }

		/* ---- E'(c), update.c:110-112 as in yee_body<+1> ---- */
		float Ex[4] = {ex.x, ex.y, ex.z, ex.w}, Ey[4] = {ey.x, ey.y, ey.z, ey.w}, Ez[4] = {ez.x, ez.y, ez.z, ez.w};
		{
			const float Hx[4] = {hx.x, hx.y, hx.z, hx.w}, Hy[4] = {hy.x, hy.y, hy.z, hy.w}, Hz[4] = {hz.x, hz.y, hz.z, hz.w};
			const float Hxn[4] = {hxn.x, hxn.y, hxn.z, hxn.w}, Hzn[4] = {hzn.x, hzn.y, hzn.z, hzn.w};
			const float Hxp[4] = {hxp.x, hxp.y, hxp.z, hxp.w}, Hyp[4] = {hyp.x, hyp.y, hyp.z, hyp.w};
			const float hy5[5] = {hy.x, hy.y, hy.z, hy.w, hye}, hz5[5] = {hz.x, hz.y, hz.z, hz.w, hze};
			#pragma unroll
			for (int e = 0; e < 4; e++) {
				const float ux = dze*(Hy[e] - Hyp[e]) + dye*(Hzn[e] - Hz[e]);
				const float uy = DxE[e]*(hz5[e] - hz5[e+1]) + dze*(Hxp[e] - Hx[e]);
				const float uz = dye*(Hx[e] - Hxn[e]) + DxE[e]*(hy5[e+1] - hy5[e]);
				if (EPS) {
					const float Cx[4] = {cx4.x, cx4.y, cx4.z, cx4.w}, Cy[4] = {cy4.x, cy4.y, cy4.z, cy4.w}, Cz[4] = {cz4.x, cz4.y, cz4.z, cz4.w};
					Ex[e] += ux*Cx[e]; Ey[e] += uy*Cy[e]; Ez[e] += uz*Cz[e];
				} else { Ex[e] += ux; Ey[e] += uy; Ez[e] += uz; }
			}
			if (PMZ && zsE >= 0 && act) {
				/* update.c:77-78 (psi) and :79-80 (their contribution) as in yee_body<+1>, PM & 4; also on the halo
				 * row and the halo plane, whose new psi is formed but not stored */
				const float zP = __ldg(p.PzE[0] + c), zP1 = __ldg(p.PzE[1] + c), zP2 = __ldg(p.PzE[2] + c);
				float v0[4] = {qe0.x, qe0.y, qe0.z, qe0.w}, v1[4] = {qe1.x, qe1.y, qe1.z, qe1.w};
				#pragma unroll
				for (int e = 0; e < 4; e++) {
					v0[e] = zP*v0[e] + zP1*(Hy[e] - Hyp[e]);
					v1[e] = zP*v1[e] + zP2*(Hx[e] - Hxp[e]);
					if (EPS) {
						const float Cx[4] = {cx4.x, cx4.y, cx4.z, cx4.w}, Cy[4] = {cy4.x, cy4.y, cy4.z, cy4.w};
						Ex[e] += v0[e]*Cx[e]; Ey[e] += v1[e]*Cy[e];
					} else { Ex[e] += v0[e]; Ey[e] += v1[e]; }
				}
				if (own && c >= cf) {
					st4(p.psiEo[zsE][0] + pe, make_float4(v0[0], v0[1], v0[2], v0[3]));
					st4(p.psiEo[zsE][1] + pe, make_float4(v1[0], v1[1], v1[2], v1[3]));
				}
			}
		}
		/* the cell a0-1 (lane 0): E'y and E'z only */
		float eyh = 0.f, ezh = 0.f, hxh = 0.f;
		if (xh) {
			const float eyo = S.edge[wy][0], ezo = S.edge[wy][1];
			const float hzh = S.edge[wy][2], hyh = S.edge[wy][3], hxnh = S.edge[wy][5];
			hxh = S.edge[wy][4];
			if (EPS) {
				eyh = eyo + (dxeh*(hzh - hz.x) + dze*(hxp1 - hxh))*cyh;
				ezh = ezo + (dye*(hxh - hxnh) + dxeh*(hy.x - hyh))*czh;
			} else {
				eyh = eyo + (dxeh*(hzh - hz.x) + dze*(hxp1 - hxh));
				ezh = ezo + (dye*(hxh - hxnh) + dxeh*(hy.x - hyh));
			}
			if (PMZ && zsE >= 0) {
				const float zP = __ldg(p.PzE[0] + c), zP2 = __ldg(p.PzE[2] + c);
				const float v1h = zP*qeh + zP2*(hxh - hxp1);
				if (EPS) eyh += v1h*cyh; else eyh += v1h;
			}
		}
		const float4 nex = make_float4(Ex[0], Ex[1], Ex[2], Ex[3]), ney = make_float4(Ey[0], Ey[1], Ey[2], Ey[3]), nez = make_float4(Ez[0], Ez[1], Ez[2], Ez[3]);
		if (own && c >= cf) { st4(p.Eo[0] + o, nex); st4(p.Eo[1] + o, ney); st4(p.Eo[2] + o, nez); }
		sm.sx[it & 1][wy][lane] = nex; sm.sz[it & 1][wy][lane] = nez;     /* read by the row above in iteration it+1 */

		/* ---- H'(c+1), update.c:115-117 as in yee_body<-1> ---- */
		if (it > 0) {
			float eya = __shfl_up_sync(0xffffffffu, eyp.w, 1), eza = __shfl_up_sync(0xffffffffu, ezp.w, 1);
			if (lane == 0) { eya = eym1; eza = ezm1; }
			if (own) {
				const size_t op = o + p.plane;
				float Hx[4] = {hxp.x, hxp.y, hxp.z, hxp.w}, Hy[4] = {hyp.x, hyp.y, hyp.z, hyp.w}, Hz[4] = {hzp.x, hzp.y, hzp.z, hzp.w};
				const float Gx[4] = {exp_.x, exp_.y, exp_.z, exp_.w}, Gy[4] = {eyp.x, eyp.y, eyp.z, eyp.w}, Gz[4] = {ezp.x, ezp.y, ezp.z, ezp.w};
				const float Gxn[4] = {exm.x, exm.y, exm.z, exm.w}, Gzn[4] = {ezm.x, ezm.y, ezm.z, ezm.w};
				const float gy5[5] = {eya, eyp.x, eyp.y, eyp.z, eyp.w}, gz5[5] = {eza, ezp.x, ezp.y, ezp.z, ezp.w};
				#pragma unroll
				for (int e = 0; e < 4; e++) {
					const float ux = dzh*(Gy[e] - Ey[e]) + dyh*(Gzn[e] - Gz[e]);
					const float uy = DxH[e]*(gz5[e+1] - gz5[e]) + dzh*(Ex[e] - Gx[e]);
					const float uz = dyh*(Gx[e] - Gxn[e]) + DxH[e]*(gy5[e] - gy5[e+1]);
					Hx[e] += ux; Hy[e] += uy; Hz[e] += uz;
				}
				if (PMZ && zsH >= 0) {
					/* update.c:77-80 as in yee_body<-1>: plane c+1 of the H set, differences against E'(c) */
					const float zP = __ldg(p.PzH[0] + c + 1), zP1 = __ldg(p.PzH[1] + c + 1), zP2 = __ldg(p.PzH[2] + c + 1);
					float v0[4] = {qh0.x, qh0.y, qh0.z, qh0.w}, v1[4] = {qh1.x, qh1.y, qh1.z, qh1.w};
					#pragma unroll
					for (int e = 0; e < 4; e++) {
						v0[e] = zP*v0[e] + zP1*(Gy[e] - Ey[e]);
						v1[e] = zP*v1[e] + zP2*(Gx[e] - Ex[e]);
						Hx[e] += v0[e]; Hy[e] += v1[e];
					}
					st4(p.psiHo[zsH][0] + ph, make_float4(v0[0], v0[1], v0[2], v0[3]));
					st4(p.psiHo[zsH][1] + ph, make_float4(v1[0], v1[1], v1[2], v1[3]));
				}
				st4(p.Ho[0] + op, make_float4(Hx[0], Hx[1], Hx[2], Hx[3]));
				st4(p.Ho[1] + op, make_float4(Hy[0], Hy[1], Hy[2], Hy[3]));
				st4(p.Ho[2] + op, make_float4(Hz[0], Hz[1], Hz[2], Hz[3]));
			}
		}
		/* plane c becomes plane c+1 */
		hxp = hx; hyp = hy; hzp = hz;
		exp_ = nex; eyp = ney; ezp = nez;
		eym1 = eyh; ezm1 = ezh; hxp1 = hxh;
	}
}

/*
 * yee_wave: the fused E+H step IN PLACE (no second set of field arrays), as a dependency wavefront.
 *
 * E'(X) reads the OLD H of X and of its +x, +y, +z neighbours; H'(X) reads the NEW E of X and of its
 * -x, -y, -z neighbours (update.c:110-117).  So a cell depends only on cells that precede it in every
 * axis: tiles can update their cells in place as long as a tile forms H' of a plane after its -x and -y
 * neighbour tiles have stored E' of that plane (true dependency; it also guarantees that the neighbour
 * has read the old H it needs from this tile, so nothing is overwritten too early).  Nothing is
 * recomputed and nothing is double buffered: every field value is read once and written once per step.
 *
 *   - tile = 128 x ROWS cells of an x-y patch, marching DOWN through one z-chunk (planes cl..cf);
 *   - tiles take their number from a global ticket (z-chunk, then y, then x): every tile a tile waits
 *     for has a smaller ticket, i.e. is already running or finished -- no deadlock, whatever the
 *     hardware's block scheduling order;
 *   - per tile one 64-bit progress word `epoch<<20 | planes stored`, published (fence, then store) by a
 *     helper warp; a second helper warp polls the -x / -y tiles' words one plane ahead of need and
 *     fetches their boundary E' (row b0-1: E'x, E'z; column a0-1: E'y, E'z) from L2 into shared memory;
 *   - the z-chunk below must have stored its top plane before this tile starts (its E' is this tile's
 *     plane cf-1, and it has then read this tile's old H(cf)); chunks of one column run concurrently;
 *   - cells outside the fused box are updated by the two-pass kernels: their E pass runs BEFORE this
 *     kernel (it needs the old H next to the box and provides E' on the box's low faces), their H
 *     pass after it.
 * The polls are bounded: a dependency that never arrives sets *error instead of hanging the GPU.
 * Arithmetic, expression order and operands are those of yee_body<+1> / yee_body<-1>: bit-identical.
 */
struct WaveParams {
	float *E[3], *H[3];
	const float *dtdE[3], *dtdH[3];   /* x tables shifted by XO                             */
	const float *C[3];                /* inverse permittivity (EPS variant)                  */
	int ny, pitch; size_t plane;
	int a_first, a_last;              /* a_first = 1 mod 4, (a_last - a_first + 1) % 4 == 0  */
	int b_first, b_last, c_first, c_last, cchunk;
	int ntx, nty, ntz;
	unsigned long long *ticket;       /* global counter; this launch owns [ticket_base, ticket_base + ntx*nty*ntz) */
	unsigned long long ticket_base;
	unsigned long long *progress;     /* one word per tile                                   */
	unsigned long long epoch;         /* launch number << 20                                 */
	int *error;
	int lag;                          /* planes a tile stays behind its -x / -y neighbours (slack against their jitter) */
};

#define ALISB_WAVE_NH 8              /* boundary-E' slots: the fetch warp runs up to six planes ahead of the workers */
template <int ROWS, bool EPS, int NST> struct WaveSmem {
	struct Stage {
		float4 e[3][ROWS][32];
		float4 hy[ROWS][32];
		float4 hx[ROWS+1][32], hz[ROWS+1][32];   /* rows b0 .. b0+ROWS (row b+1 feeds the y differences) */
		float4 c[EPS ? 3 : 1][EPS ? ROWS : 1][EPS ? 32 : 1];
		float  edge[ROWS][2];                    /* Hy, Hz of the cell a0+128 (read by the last lane of a row) */
	} st[NST];
	float4 sx[2][ROWS][32], sz[2][ROWS][32];     /* E'x, E'z rows of a plane, for the row above   */
	float4 hrow[ALISB_WAVE_NH][2][32];           /* [slot][E'x, E'z] of row b0-1 (from the -y tile) */
	float  hcol[ALISB_WAVE_NH][2][ROWS];         /* [slot][E'y, E'z] of column a0-1 (from the -x tile) */
	unsigned long long ticket;
	int iter_done;                               /* workers -> helpers: iterations < iter_done are complete */
	int halo_ready;                              /* fetch warp -> workers: boundary E' of plane indices < halo_ready are in shared memory */
};

/* progress words: release store by the publishing tile's helper, acquire load by the polling one (the stores
 * of the publishing tile's workers reach the helper through the shared-memory counter below, released and
 * acquired at CTA scope: causality order is cumulative).  Returns the value seen (>= need), or 0 after giving up. */
static __device__ __forceinline__ unsigned long long wave_wait(const unsigned long long *word, unsigned long long need, int *error)
{
	unsigned spins = 0;
	for (;;) {
		unsigned long long v;
		asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(word) : "memory");
		if (v >= need) return v;
		__nanosleep(20);
		if ((++spins & 1023u) == 0 && *(volatile int*)error) return 0;       /* somebody already gave up: do not wait again */
		if (spins > (1u << 22)) { atomicExch(error, 1); return 0; }          /* seconds: give up rather than hang */
	}
}
static __device__ __forceinline__ void wave_publish(unsigned long long *word, unsigned long long v)
{
	asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(word), "l"(v) : "memory");
}
static __device__ __forceinline__ void smem_wait_ge(const int *word, int need)
{
	const unsigned a = (unsigned)__cvta_generic_to_shared(word);
	int v;
	do { asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); } while (v < need);
}
static __device__ __forceinline__ void smem_publish(int *word, int value)
{
	const unsigned a = (unsigned)__cvta_generic_to_shared(word);
	asm volatile("st.release.cta.shared.s32 [%0], %1;" :: "r"(a), "r"(value) : "memory");
}
static __device__ __forceinline__ void worker_barrier(int nthreads)
{
	asm volatile("bar.sync 1, %0;" :: "r"(nthreads) : "memory");
}

/* The two helper warps never take part in the workers' per-plane barrier: they talk to them through the
 * two shared-memory counters above, so a poll of a neighbour tile's progress word or a memory fence
 * is never on the workers' critical path. */
template <int ROWS, bool EPS, int NST>
__global__ void __launch_bounds__(32*(ROWS+2), 1)
yee_wave(const __grid_constant__ WaveParams p)
{
	constexpr int AHEAD = NST - 1, NH = ALISB_WAVE_NH;
	extern __shared__ __align__(16) unsigned char wave_smem_raw[];
	WaveSmem<ROWS,EPS,NST> &sm = *reinterpret_cast<WaveSmem<ROWS,EPS,NST>*>(wave_smem_raw);
	typedef typename WaveSmem<ROWS,EPS,NST>::Stage Stage;
	const int lane = threadIdx.x, wy = threadIdx.y;
	const bool worker = wy < ROWS, fetcher = wy == ROWS;

	if (lane == 0 && wy == 0) { sm.ticket = atomicAdd(p.ticket, 1ull) - p.ticket_base; sm.iter_done = 0; sm.halo_ready = 0; }
	__syncthreads();
	const unsigned tile = (unsigned)sm.ticket;
	const int tx = (int)(tile % (unsigned)p.ntx), ty = (int)((tile / (unsigned)p.ntx) % (unsigned)p.nty), tz = (int)(tile / ((unsigned)p.ntx*(unsigned)p.nty));
	const unsigned long long *prog_x = tx > 0 ? p.progress + tile - 1 : 0;
	const unsigned long long *prog_y = ty > 0 ? p.progress + tile - p.ntx : 0;
	const unsigned long long *prog_z = tz > 0 ? p.progress + tile - (unsigned)p.ntx*(unsigned)p.nty : 0;

	const int a0 = p.a_first + 4*(tx*32 + lane);
	const int b0 = p.b_first + ty*ROWS;
	const int cf = p.c_first + tz*p.cchunk;
	const int cl = min(cf + p.cchunk - 1, p.c_last);
	const int nplanes = cl - cf + 1;
	float *E0 = p.E[0], *E1 = p.E[1], *E2 = p.E[2];
	float *H0 = p.H[0], *H1 = p.H[1], *H2 = p.H[2];

	/* the z-chunk below has stored its top plane (= this tile's plane cf-1) and read the old H(cf) */
	if (fetcher && lane == 0 && prog_z) (void)wave_wait(prog_z, p.epoch + 1, p.error);
	__syncthreads();

	if (!worker) {
		if (fetcher) {
			/* boundary E' of plane index k (plane cl-k) from the -y tile (row b0-1) and the -x tile (column a0-1),
			 * used by the workers in iteration k+1; the neighbours must have stored k+1 planes */
			/* Every round fetches as many planes as are allowed right now -- the workers have freed the slots
			 * (at most NH-1 planes ahead of the iteration they are in) and both neighbours have stored them
			 * (a poll returns how far a neighbour has come: usually several planes, so most rounds need no
			 * poll) -- then waits once for the copies and publishes them together. */
			unsigned long long known_x = 0, known_y = 0;
			int k = 0;
			while (k < nplanes) {
				int kend = k;
				if (lane == 0) {
					const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm.iter_done);
					int itw;
					for (;;) { asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(itw) : "r"(sa) : "memory"); if (itw >= k - NH + 2) break; __nanosleep(100); }
					/* stay `lag` planes behind the neighbours (they may be at most nplanes ahead): a neighbour's
					 * hiccup then eats slack instead of stalling this tile and everything downstream of it */
					const unsigned long long need = p.epoch + (unsigned long long)min(k + 1 + p.lag, nplanes);
					if (prog_x && known_x < need) known_x = wave_wait(prog_x, need, p.error);
					if (prog_y && known_y < need) known_y = wave_wait(prog_y, need, p.error);
					kend = min(nplanes - 1, itw + NH - 2);
					if (prog_x) kend = min(kend, known_x >= need ? (int)(known_x - p.epoch) - 1 : k);
					if (prog_y) kend = min(kend, known_y >= need ? (int)(known_y - p.epoch) - 1 : k);
					kend = max(kend, k);                             /* plane k itself is covered by `need` (lag >= 0) */
				}
				kend = __shfl_sync(0xffffffffu, kend, 0);
				for (int j = k; j <= kend; j++) {
					const int c = cl - j, slot = j & (NH - 1);
					const size_t o = (size_t)c*p.plane + (size_t)(b0 - 1)*p.pitch + ALISB_XO + a0;
					if (a0 <= p.a_last) { cp_async16(&sm.hrow[slot][0][lane], E0 + o); cp_async16(&sm.hrow[slot][1][lane], E2 + o); }
					if (lane < ROWS && b0 + lane <= p.b_last) {
						const size_t q = (size_t)c*p.plane + (size_t)(b0 + lane)*p.pitch + ALISB_XO + (p.a_first + 128*tx) - 1;
						cp_async4(&sm.hcol[slot][0][lane], E1 + q); cp_async4(&sm.hcol[slot][1][lane], E2 + q);
					}
				}
				cp_async_commit();
				cp_async_wait<0>();                                  /* one L2 round trip per round */
				__syncwarp();
				if (lane == 0) smem_publish(&sm.halo_ready, kend + 1);
				k = kend + 1;
			}
		} else if (lane == 0) {
			/* every worker stored E' of plane indices < it before the barrier that opened iteration it.  A release
			 * costs a memory barrier (it waits for the SM's stores in flight): publish whatever has been reached
			 * since the last one instead of one word per plane -- the successors then see progress in batches but
			 * never faster planes than barriers */
			int done = 0;
			const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm.iter_done);
			while (done < nplanes) {
				int itw;
				for (;;) { asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(itw) : "r"(sa) : "memory"); if (itw > done) break; __nanosleep(100); }
				wave_publish(p.progress + tile, p.epoch + (unsigned long long)itw);
				done = itw;
			}
		}
		return;
	}

	const int b  = b0 + wy;
	const bool rowok = b <= p.b_last;
	const bool act = rowok && a0 <= p.a_last;              /* owns a quad                              */
	const bool ldg = rowok && a0 <= p.a_last + 1;          /* its H quad feeds the lane before         */
	const bool xt  = act && lane == 31;                    /* needs H of the cell a0+128               */
	const bool top = wy == ROWS - 1 || b == p.b_last;      /* nobody stages row b+1 for this warp      */
	const size_t row = (size_t)b*p.pitch + ALISB_XO + a0;
	const ptrdiff_t yp = p.pitch;

	/* all loads of plane c into stage s; plane cf-1 only supplies the final E'x, E'y of the chunk below */
	auto stage_plane = [&](int c, int s) {
		Stage &S = sm.st[s];
		const size_t o = (size_t)c*p.plane + row;
		if (c < cf) { if (act) { cp_async16(&S.e[0][wy][lane], E0 + o); cp_async16(&S.e[1][wy][lane], E1 + o); } return; }
		if (ldg) { cp_async16(&S.hx[wy][lane], H0 + o); cp_async16(&S.hy[wy][lane], H1 + o); cp_async16(&S.hz[wy][lane], H2 + o); }
		if (act) {
			cp_async16(&S.e[0][wy][lane], E0 + o); cp_async16(&S.e[1][wy][lane], E1 + o); cp_async16(&S.e[2][wy][lane], E2 + o);
			if (top) { cp_async16(&S.hx[wy + 1][lane], H0 + o + yp); cp_async16(&S.hz[wy + 1][lane], H2 + o + yp); }
			if (EPS) { cp_async16(&S.c[0][wy][lane], p.C[0] + o); cp_async16(&S.c[1][wy][lane], p.C[1] + o); cp_async16(&S.c[2][wy][lane], p.C[2] + o); }
		}
		if (xt) { cp_async4(&S.edge[wy][0], H1 + o + 4); cp_async4(&S.edge[wy][1], H2 + o + 4); }
	};

	const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
	const float4 dxe4 = act ? ld4(p.dtdE[0] + ALISB_XO + a0) : zero4;
	const float4 dxh4 = act ? ld4(p.dtdH[0] + ALISB_XO + a0) : zero4;
	const float dye = rowok ? __ldg(p.dtdE[1] + b) : 0.f, dyh = rowok ? __ldg(p.dtdH[1] + b) : 0.f;
	const float DxE[4] = {dxe4.x, dxe4.y, dxe4.z, dxe4.w}, DxH[4] = {dxh4.x, dxh4.y, dxh4.z, dxh4.w};

	/* pipeline prologue: the first AHEAD planes */
	#pragma unroll
	for (int q = 0; q < AHEAD; q++) {
		if (cl - q >= cf - 1) stage_plane(cl - q, q);
		cp_async_commit();
	}

	/* carried state of plane c+1 */
	float4 hxp = zero4, hyp = zero4, hzp = zero4;           /* old H(c+1)                              */
	float4 exp_ = zero4, eyp = zero4, ezp = zero4;          /* E'(c+1)                                 */
	if (ldg) {
		const size_t o = (size_t)(cl + 1)*p.plane + row;
		hxp = ld4(H0 + o); hyp = ld4(H1 + o);
	}
	float dze_next = __ldg(p.dtdE[2] + cl), dzh_next = __ldg(p.dtdH[2] + cl + 1);   /* z coefficients one iteration ahead */
	int halo_known = 0;                                     /* boundary planes known to be in shared memory (uniform in the warp) */
	int c = cl;
	for (int it = 0; it <= nplanes; ++it, --c) {
		const size_t o = (size_t)c*p.plane + row;
		const float dze = dze_next, dzh = dzh_next;
		if (it < nplanes) { dze_next = __ldg(p.dtdE[2] + c - 1); dzh_next = __ldg(p.dtdH[2] + c); }
		cp_async_wait<AHEAD - 1>();                         /* this thread's copies of plane c have landed */
		worker_barrier(32*ROWS);                            /* ... and every worker's; iteration it-1 is over in all of them */
		if (wy == 0 && lane == 0) smem_publish(&sm.iter_done, it);
		{                                                   /* plane c-AHEAD goes into the stage iteration it-1 used */
			if (c - AHEAD >= cf - 1) stage_plane(c - AHEAD, (it + AHEAD) % NST);
			cp_async_commit();
		}
		const Stage &S = sm.st[it % NST];
		float4 ex = zero4, ey = zero4, ez = zero4, hx = zero4, hy = zero4, hz = zero4;
		if (act) { ex = S.e[0][wy][lane]; ey = S.e[1][wy][lane]; }
		float Ex[4] = {ex.x, ex.y, ex.z, ex.w}, Ey[4] = {ey.x, ey.y, ey.z, ey.w}, Ez[4] = {0.f, 0.f, 0.f, 0.f};
		float4 nex = ex, ney = ey, nez = zero4;
		/* E'x, E'z of row b-1 at plane c+1 from the warp below (its iteration it-1) */
		float4 exm = zero4, ezm = zero4;
		if (it > 0 && wy > 0) { exm = sm.sx[(it - 1) & 1][wy - 1][lane]; ezm = sm.sz[(it - 1) & 1][wy - 1][lane]; }
		if (it < nplanes) {
			/* ---- E'(c), update.c:110-112 as in yee_body<+1> ---- */
			float4 hxn = zero4, hzn = zero4, cx4 = zero4, cy4 = zero4, cz4 = zero4;
			if (ldg) { hx = S.hx[wy][lane]; hy = S.hy[wy][lane]; hz = S.hz[wy][lane]; }
			if (act) {
				hxn = S.hx[wy + 1][lane]; hzn = S.hz[wy + 1][lane];
				ez = S.e[2][wy][lane];
				if (EPS) { cx4 = S.c[0][wy][lane]; cy4 = S.c[1][wy][lane]; cz4 = S.c[2][wy][lane]; }
			}
			Ez[0] = ez.x; Ez[1] = ez.y; Ez[2] = ez.z; Ez[3] = ez.w;
			float hye = __shfl_down_sync(0xffffffffu, hy.x, 1), hze = __shfl_down_sync(0xffffffffu, hz.x, 1);
			if (xt) { hye = S.edge[wy][0]; hze = S.edge[wy][1]; }
			const float Hx[4] = {hx.x, hx.y, hx.z, hx.w}, Hy[4] = {hy.x, hy.y, hy.z, hy.w}, Hz[4] = {hz.x, hz.y, hz.z, hz.w};
			const float Hxn[4] = {hxn.x, hxn.y, hxn.z, hxn.w}, Hzn[4] = {hzn.x, hzn.y, hzn.z, hzn.w};
			const float Hxp[4] = {hxp.x, hxp.y, hxp.z, hxp.w}, Hyp[4] = {hyp.x, hyp.y, hyp.z, hyp.w};
			const float hy5[5] = {hy.x, hy.y, hy.z, hy.w, hye}, hz5[5] = {hz.x, hz.y, hz.z, hz.w, hze};
			#pragma unroll
			for (int e = 0; e < 4; e++) {
				const float ux = dze*(Hy[e] - Hyp[e]) + dye*(Hzn[e] - Hz[e]);
				const float uy = DxE[e]*(hz5[e] - hz5[e+1]) + dze*(Hxp[e] - Hx[e]);
				const float uz = dye*(Hx[e] - Hxn[e]) + DxE[e]*(hy5[e+1] - hy5[e]);
				if (EPS) {
					const float Cx[4] = {cx4.x, cx4.y, cx4.z, cx4.w}, Cy[4] = {cy4.x, cy4.y, cy4.z, cy4.w}, Cz[4] = {cz4.x, cz4.y, cz4.z, cz4.w};
					Ex[e] += ux*Cx[e]; Ey[e] += uy*Cy[e]; Ez[e] += uz*Cz[e];
				} else { Ex[e] += ux; Ey[e] += uy; Ez[e] += uz; }
			}
			nex = make_float4(Ex[0], Ex[1], Ex[2], Ex[3]); ney = make_float4(Ey[0], Ey[1], Ey[2], Ey[3]); nez = make_float4(Ez[0], Ez[1], Ez[2], Ez[3]);
			if (act) { st4(E0 + o, nex); st4(E1 + o, ney); st4(E2 + o, nez); }
			sm.sx[it & 1][wy][lane] = nex; sm.sz[it & 1][wy][lane] = nez;     /* read by the row above in iteration it+1 */
		}

		/* ---- H'(c+1), update.c:115-117 as in yee_body<-1> ---- */
		if (it > 0) {
			/* boundary E' of plane c+1 (plane index it-1) from the neighbour tiles */
			if (halo_known < it) {
				if (lane == 0) {
					const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm.halo_ready);
					do { asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(halo_known) : "r"(sa) : "memory"); } while (halo_known < it);
				}
				halo_known = __shfl_sync(0xffffffffu, halo_known, 0);
			}
			const int slot = (it - 1) & (NH - 1);
			if (wy == 0) { exm = sm.hrow[slot][0][lane]; ezm = sm.hrow[slot][1][lane]; }
			float eya = __shfl_up_sync(0xffffffffu, eyp.w, 1), eza = __shfl_up_sync(0xffffffffu, ezp.w, 1);
			if (lane == 0) { eya = sm.hcol[slot][0][wy]; eza = sm.hcol[slot][1][wy]; }
			if (act) {
				const size_t op = o + p.plane;
				float Hx[4] = {hxp.x, hxp.y, hxp.z, hxp.w}, Hy[4] = {hyp.x, hyp.y, hyp.z, hyp.w}, Hz[4] = {hzp.x, hzp.y, hzp.z, hzp.w};
				const float Gx[4] = {exp_.x, exp_.y, exp_.z, exp_.w}, Gy[4] = {eyp.x, eyp.y, eyp.z, eyp.w}, Gz[4] = {ezp.x, ezp.y, ezp.z, ezp.w};
				const float Gxn[4] = {exm.x, exm.y, exm.z, exm.w}, Gzn[4] = {ezm.x, ezm.y, ezm.z, ezm.w};
				const float gy5[5] = {eya, eyp.x, eyp.y, eyp.z, eyp.w}, gz5[5] = {eza, ezp.x, ezp.y, ezp.z, ezp.w};
				#pragma unroll
				for (int e = 0; e < 4; e++) {
					const float ux = dzh*(Gy[e] - Ey[e]) + dyh*(Gzn[e] - Gz[e]);
					const float uy = DxH[e]*(gz5[e+1] - gz5[e]) + dzh*(Ex[e] - Gx[e]);
					const float uz = dyh*(Gx[e] - Gxn[e]) + DxH[e]*(gy5[e] - gy5[e+1]);
					Hx[e] += ux; Hy[e] += uy; Hz[e] += uz;
				}
				st4(H0 + op, make_float4(Hx[0], Hx[1], Hx[2], Hx[3]));
				st4(H1 + op, make_float4(Hy[0], Hy[1], Hy[2], Hy[3]));
				st4(H2 + op, make_float4(Hz[0], Hz[1], Hz[2], Hz[3]));
			}
		}
		/* plane c becomes plane c+1 */
		hxp = hx; hyp = hy; hzp = hz;
		exp_ = nex; eyp = ney; ezp = nez;
	}
	/* the last iteration is complete too: releases the helpers' final waits (nothing reads iter_done beyond nplanes) */
}

/* one box per launch: the slab-free interior (PM=0) and, with ALISB_MERGE=0, every other box */
template <int S, bool EPS, int PM, int XL, bool ADE>
__global__ void __launch_bounds__(256, (PM == 1 || PM == 2 || PM == 4) ? 3 : ADE ? 2 : 1)   /* single-axis slab variants: keep 3 CTAs/SM; interior with objects: 2 */
yee_volume(const __grid_constant__ VolParams p, const __grid_constant__ VolRegion r)
{
	yee_body<S,EPS,PM,XL,ADE>(p, r, blockIdx.x, blockIdx.y, blockIdx.z);
}

/* many boxes per launch: the PML shell of a grid is up to 26 boxes (6 faces, 12 edges, 8 corners),
 * most of them far too small to fill the GPU on their own.  A CTA looks its box up in the table and
 * runs the body specialised for that box's PML bitmask.  Two classes so that the big single-axis
 * faces keep 3 CTAs/SM: CLS 0 = faces (PM 1,2,4), CLS 1 = edges and corners (PM 3,5,6,7).   */
template <int S, bool EPS, int CLS, bool ADE>
__global__ void __launch_bounds__(256, CLS == 0 ? 3 : 2)
yee_multi(const __grid_constant__ MultiParams mp)
{
	unsigned id = blockIdx.x;
	int q = 0;
	while (q + 1 < mp.nreg && id >= mp.r[q + 1].cta0) q++;
	const VolRegion &r = mp.r[q];
	id -= r.cta0;
	const int bix = (int)(id % (unsigned)r.gx), biy = (int)((id / (unsigned)r.gx) % (unsigned)r.gy), biz = (int)(id / ((unsigned)r.gx*(unsigned)r.gy));
	if (CLS == 0) {
		if (r.pm == 1) { if (r.narrow) yee_body<S,EPS,1,8,ADE>(mp.v, r, bix, biy, biz); else yee_body<S,EPS,1,32,ADE>(mp.v, r, bix, biy, biz); }
		else if (r.pm == 2) yee_body<S,EPS,2,32,ADE>(mp.v, r, bix, biy, biz);
		else                yee_body<S,EPS,4,32,ADE>(mp.v, r, bix, biy, biz);
	} else {
		if (r.pm == 5) { if (r.narrow) yee_body<S,EPS,5,8,ADE>(mp.v, r, bix, biy, biz); else yee_body<S,EPS,5,32,ADE>(mp.v, r, bix, biy, biz); }
		else if (r.pm == 3) yee_body<S,EPS,3,32,ADE>(mp.v, r, bix, biy, biz);
		else if (r.pm == 6) yee_body<S,EPS,6,32,ADE>(mp.v, r, bix, biy, biz);
		else                yee_body<S,EPS,7,32,ADE>(mp.v, r, bix, biy, biz);
	}
}

/* ------------------------------------------------------------------------------------------
 * Box kernels (regions, planes, strips).  A box is [lo,hi] inclusive in array coordinates,
 * threads run x fastest.
 * ------------------------------------------------------------------------------------------ */
static __device__ __forceinline__ bool box_cell(const Box &B, size_t t, int &a, int &b, int &c, size_t &q)
{
	const size_t n = (size_t)B.na*B.nb*B.nc;
	if (t >= n) return false;
	const int ia = (int)(t % B.na); const size_t r = t / B.na;
	const int ib = (int)(r % B.nb); const int ic = (int)(r / B.nb);
	a = B.a0 + ia; b = B.b0 + ib; c = B.c0 + ic; q = t;
	return true;
}
#define GIDX(a,b,c) (((size_t)(c)*(ny+2) + (size_t)(b))*pitch + ALISB_XO + (a))

/* reference update.c:89-108: conductivity loss then subtraction of all polarisation currents */
struct AdeParams {
	float *F[3];                  /* E components */
	const float *fill[3];         /* dense box arrays, x fastest */
	float *J[ALISB_MAXPOLES][3], *K[ALISB_MAXPOLES][3];
	float a[3];
	float JJ[3][ALISB_MAXPOLES], JK[3][ALISB_MAXPOLES], JE[3][ALISB_MAXPOLES];
	float KJ[3][ALISB_MAXPOLES], KK[3][ALISB_MAXPOLES], KE[3][ALISB_MAXPOLES];
	int nj, nk, lossy;
	Box full;                     /* allocation box of fill/J/K                         */
	Box upd;                      /* full clipped to the updated cells 1..n              */
	int ny, pitch;
};

__global__ void ade_loss(const __grid_constant__ AdeParams p)
{
	const int ny = p.ny, pitch = p.pitch;
	int a, b, c; size_t q;
	if (!box_cell(p.upd, (size_t)blockIdx.x*blockDim.x + threadIdx.x, a, b, c, q)) return;
	const size_t i = GIDX(a, b, c);
	const size_t j = ((size_t)(c - p.full.c0)*p.full.nb + (b - p.full.b0))*p.full.na + (a - p.full.a0);
	#pragma unroll
	for (int k = 0; k < 3; k++) {
		const float f = p.fill[k][j];
		p.F[k][i] = p.F[k][i] * ((1 - p.a[k]*f) / (1 + p.a[k]*f));       /* update.c:94-96 */
	}
}

__global__ void ade_subtract(const __grid_constant__ AdeParams p)
{
	const int ny = p.ny, pitch = p.pitch;
	int a, b, c; size_t q;
	if (!box_cell(p.upd, (size_t)blockIdx.x*blockDim.x + threadIdx.x, a, b, c, q)) return;
	const size_t i = GIDX(a, b, c);
	const size_t j = ((size_t)(c - p.full.c0)*p.full.nb + (b - p.full.b0))*p.full.na + (a - p.full.a0);
	#pragma unroll
	for (int k = 0; k < 3; k++) {
		float f = p.F[k][i];
		for (int s = 0; s < p.nk; s++) f -= p.J[s][k][j];                  /* update.c:104-106 */
		p.F[k][i] = f;
	}
}

/* reference update.c:128-158 (updateJ): Drude and critical-point/Lorentz ADE recursions */
__global__ void ade_currents(const __grid_constant__ AdeParams p)
{
	const int ny = p.ny, pitch = p.pitch;
	int a, b, c; size_t q;
	if (!box_cell(p.upd, (size_t)blockIdx.x*blockDim.x + threadIdx.x, a, b, c, q)) return;
	const size_t i = GIDX(a, b, c);
	const size_t j = ((size_t)(c - p.full.c0)*p.full.nb + (b - p.full.b0))*p.full.na + (a - p.full.a0);
	#pragma unroll
	for (int k = 0; k < 3; k++) {
		const float F = p.F[k][i], f = p.fill[k][j];
		for (int s = 0; s < p.nk; s++) {
			if (s < p.nj) {
				p.J[s][k][j] = p.JJ[k][s]*p.J[s][k][j] + p.JE[k][s]*f*F;                       /* update.c:135 */
			} else {
				const float Kold = p.K[s][k][j];
				const float Jn = p.JJ[k][s]*p.J[s][k][j] + p.JK[k][s]*Kold + p.JE[k][s]*f*F;   /* update.c:140 */
				p.J[s][k][j] = Jn;
				p.K[s][k][j] = p.KJ[k][s]*Jn + p.KK[k][s]*Kold + p.KE[k][s]*f*F;               /* update.c:142 */
			}
		}
	}
}

/* reference update.c:344-391 (coupleTWM): second-order mixing  F.v += 2*Alt*d[v][q]*C.v*fill.v*(G1..*G2..)
 * inside an object's box.  The reference makes 9 full sweeps per component, each a read-modify-
 * write of F; one thread replays them for its own cell in the same order.  Field values that do
 * not sit on F.v's Yee position are 4-point averages formed as 0.25*(float sum), a double (the
 * literal 0.25 is a double in C), so every product that involves an average is done in double and
 * the result is rounded to float once per sweep, exactly like `F += double` in C.  G1/G2 may alias
 * F (same world on both sides): own-cell values are therefore re-read after every sweep (volatile
 * accesses, no __restrict__).  One launch per component v, in x, y, z order.                   */
struct Chi2Params {
	float *F[3];
	const float *G1[3], *G2[3];
	const float *eps[3];          /* C->x/y/z on the full grid                           */
	const float *fill[3];         /* dense box arrays of the object, x fastest           */
	float d[3][6];
	float alt;
	Box full, upd;                /* fill allocation box; cells to update (iMin..iMax-1) */
	int ny, pitch; long long plane;
};

/* average of component U onto the position of component V (update.c:347-352) */
template<int V, int U>
__device__ __forceinline__ double chi2_avg(const float *g, size_t i, const long long *e)
{
	return 0.25*(double)(((g[i] + g[i - e[V]]) + g[i + e[U]]) + g[i - e[V] + e[U]]);
}

template<int V, int U1, int U2>
__device__ __forceinline__ void chi2_term(const Chi2Params &p, size_t i, size_t j, float dq, const long long *e)
{
	if (!dq) return;                                                          /* twm(): if (C->M[n].d) */
	volatile float *F = p.F[V];
	const float pre = 2*p.alt*dq*p.eps[V][i]*p.fill[V][j];                    /* float, left to right */
	if (U1 == V && U2 == V) {
		const float g1 = ((const volatile float*)p.G1[V])[i], g2 = ((const volatile float*)p.G2[V])[i];
		F[i] = F[i] + pre*(g1*g2);
	} else {
		const double g1 = U1 == V ? (double)((const volatile float*)p.G1[V])[i] : chi2_avg<V,U1>(p.G1[U1], i, e);
		const double g2 = U2 == V ? (double)((const volatile float*)p.G2[V])[i] : chi2_avg<V,U2>(p.G2[U2], i, e);
		F[i] = (float)((double)F[i] + (double)pre*(g1*g2));
	}
}

template<int V>
__global__ void chi2_couple(const __grid_constant__ Chi2Params p)
{
	const int ny = p.ny, pitch = p.pitch;
	int a, b, c; size_t q;
	if (!box_cell(p.upd, (size_t)blockIdx.x*blockDim.x + threadIdx.x, a, b, c, q)) return;
	const size_t i = GIDX(a, b, c);
	const size_t j = ((size_t)(c - p.full.c0)*p.full.nb + (b - p.full.b0))*p.full.na + (a - p.full.a0);
	const long long e[3] = {1, (long long)pitch, p.plane};
	constexpr int A = V, B1 = (V+1)%3, B2 = (V+2)%3;       /* x: (x,y,z)  y: (y,z,x)  z: (z,x,y) */
	(void)A; (void)B1; (void)B2;
	/* Voigt order of update.c:364-390: xx, yy, zz, then yz+zy, zx+xz, xy+yx */
	chi2_term<V,0,0>(p, i, j, p.d[V][0], e);
	chi2_term<V,1,1>(p, i, j, p.d[V][1], e);
	chi2_term<V,2,2>(p, i, j, p.d[V][2], e);
	chi2_term<V,1,2>(p, i, j, p.d[V][3], e);
	chi2_term<V,2,1>(p, i, j, p.d[V][3], e);
	chi2_term<V,2,0>(p, i, j, p.d[V][4], e);
	chi2_term<V,0,2>(p, i, j, p.d[V][4], e);
	chi2_term<V,0,1>(p, i, j, p.d[V][5], e);
	chi2_term<V,1,0>(p, i, j, p.d[V][5], e);
}

/* reference update.c:161-179 (updateS).  vals != 0: per-cell products Amplitude*wave(t,Phase)
 * evaluated on the host; else amp[q]*f or amplitude*f with the host-evaluated scalar f.   */
__global__ void source_add(float *F, Box B, Box full, int ny, int pitch, const float *vals, const float *amp, float amplitude, float f)
{
	int a, b, c; size_t q;
	if (!box_cell(B, (size_t)blockIdx.x*blockDim.x + threadIdx.x, a, b, c, q)) return;
	const size_t i = GIDX(a, b, c);
	q = ((size_t)(c - full.c0)*full.nb + (b - full.b0))*full.na + (a - full.a0);     /* index into the source's own arrays */
	float v;
	if (vals) v = vals[q];
	else v = (amp ? amp[q] : amplitude) * f;
	F[i] += v;
}

/* reference update.c:194-202 (update): one TFSF face strip, aux array read through the
 * aliasing clamp of source.c:223-248 */
__global__ void tfsf_strip(float *F, const float *G, const float *C, float coef, Box B, int ny, int pitch,
                           int any, int apitch, int alt0, int alt1, int alt2, int clampa, int clampb)
{
	int a, b, c; size_t q;
	if (!box_cell(B, (size_t)blockIdx.x*blockDim.x + threadIdx.x, a, b, c, q)) return;
	const size_t i = GIDX(a, b, c);
	const int xa = min(a + alt0, clampa), xb = min(b + alt1, clampb), xc = c + alt2;
	const float g = G[((size_t)xc*(any + 2) + xb)*apitch + ALISB_XO + xa];
	if (C) F[i] += coef*g*C[i];
	else   F[i] += coef*g;
}

/* reference update.c:182-191 (copyBoundary) */
__global__ void copy_plane(float *F, float *G, float cosk, float sink, Box B, int ny, int pitch, int alt0, int alt1, int alt2)
{
	int a, b, c; size_t q;
	if (!box_cell(B, (size_t)blockIdx.x*blockDim.x + threadIdx.x, a, b, c, q)) return;
	const size_t i = GIDX(a, b, c), j = GIDX(a + alt0, b + alt1, c + alt2);
	if (G) {
		const float f = F[j], g = G[j];
		F[i] = cosk*f + sink*g;
		G[i] = cosk*g - sink*f;
	} else F[i] = cosk*F[j];
}

/* The whole ordered copy list of a pass in ONE launch, for grids whose boundary planes are small:
 * a single CTA walks the list and separates the (order-dependent) copies with __syncthreads().
 * Cuts a Bloch/hexagonal H pass from 16 tiny launches to one. */
#define ALISB_MAXCOPIES 24
#define ALISB_LIST_CTAS 8          /* one thread-block cluster walks a list; cluster.sync() separates dependent ops */
struct CopyOp { float *F, *G; float cosk, sink; Box B; int alt0, alt1, alt2; int sync_after; };
/* Ops are grouped into CHAINS by the array they act on (a copy reads and writes one component, a strip or a source adds
 * to one): chains are independent and run side by side -- chain q is walked by CTAs q*CL .. q*CL+CL-1 (CL = 1: one CTA,
 * __syncthreads between dependent ops; CL = 8: one thread-block cluster, barrier.cluster) -- and inside a chain the ops keep
 * their list order with a barrier wherever an op touches what an earlier one of the same level wrote or read. */
#define ALISB_MAXCHAINS 6
struct CopyList { CopyOp op[ALISB_MAXCOPIES]; int n, ny, pitch; int nchain, first[ALISB_MAXCHAINS + 1]; };

static __device__ __forceinline__ void list_sync()
{
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CL> static __device__ __forceinline__ void chain_sync()
{
	if (CL == 1) __syncthreads(); else list_sync();
}

template <int CL> static __device__ __forceinline__ void copy_chain(const CopyList &L)
{
	const int ny = L.ny, pitch = L.pitch;
	const int chain = blockIdx.x / CL;
	const size_t tid = (size_t)(blockIdx.x % CL)*blockDim.x + threadIdx.x, nth = (size_t)CL*blockDim.x;
	for (int k = L.first[chain]; k < L.first[chain + 1]; k++) {
		const CopyOp &o = L.op[k];
		const size_t cells = (size_t)o.B.na*o.B.nb*o.B.nc;
		/* four cells per thread and round: all loads first (a copy's source cells are never targets of the same copy),
		 * so a thread with several cells pays one memory round trip, not one per cell */
		for (size_t t0 = tid; t0 < cells; t0 += 4*nth) {
			size_t i[4]; float f[4], g[4]; bool ok[4];
			#pragma unroll
			for (int u = 0; u < 4; u++) {
				const size_t t = t0 + (size_t)u*nth;
				ok[u] = t < cells; i[u] = 0; f[u] = 0.f; g[u] = 0.f;
				if (!ok[u]) continue;
				int a, b, c; size_t q;
				box_cell(o.B, t, a, b, c, q);
				i[u] = GIDX(a, b, c);
				const size_t j = GIDX(a + o.alt0, b + o.alt1, c + o.alt2);
				f[u] = o.F[j];
				if (o.G) g[u] = o.G[j];
			}
			#pragma unroll
			for (int u = 0; u < 4; u++) {
				if (!ok[u]) continue;
				if (o.G) { o.F[i[u]] = o.cosk*f[u] + o.sink*g[u]; o.G[i[u]] = o.cosk*g[u] - o.sink*f[u]; }
				else o.F[i[u]] = o.cosk*f[u];
			}
		}
		if (o.sync_after) chain_sync<CL>();     /* the next op of the chain reads or overwrites what this level wrote */
	}
}
__global__ void __cluster_dims__(ALISB_LIST_CTAS, 1, 1) __launch_bounds__(1024) copy_list(const __grid_constant__ CopyList L) { copy_chain<ALISB_LIST_CTAS>(L); }
__global__ void __launch_bounds__(1024) copy_list1(const __grid_constant__ CopyList L) { copy_chain<1>(L); }

/* The ordered copy list RESOLVED on the host (engine.cu, build_copy_plan): walking the list symbolically gives, for every
 * cell some copy writes, the cell whose value from BEFORE the list it ends up with and the sequence of Bloch rotations
 * (cosk, sink) that value passes through on its way (a ghost corner of a periodic world goes through up to three copies).
 * The kernel applies those rotations one after the other in registers -- the same multiplications and additions in the
 * same order as the chain of copies (update.c:182-191), so the result is bit-identical -- and the whole list needs ONE
 * barrier (every load before any store: a source cell may itself be rewritten) instead of one per dependent level. */
#define ALISB_RC_ARRAYS 8
#define ALISB_RC_CHAINS 48
#define ALISB_RC_DEPTH 4
#define ALISB_RC_PER_THREAD 8
struct ResolvedCopy {
	float *F[ALISB_RC_ARRAYS], *G[ALISB_RC_ARRAYS];
	float cosk[ALISB_RC_CHAINS][ALISB_RC_DEPTH], sink[ALISB_RC_CHAINS][ALISB_RC_DEPTH];
	int depth[ALISB_RC_CHAINS];
	const int4 *ent;        /* x: target index, y: source index, z: array << 8 | chain; the bound entries first */
	float2 *tmp;            /* one slot per bound entry, used when a thread owns more than ALISB_RC_PER_THREAD of them */
	int n, nbound;          /* entries [0,nbound) read or write a cell that the list both reads and writes: every load before any
	                         * store (one cluster, one barrier); entries [nbound,n) touch nothing another entry touches */
};
static __device__ __forceinline__ float2 rc_value(const ResolvedCopy &P, const int4 en)
{
	const int arr = en.z >> 8, ch = en.z & 255;
	float f = P.F[arr][en.y], g = P.G[arr] ? P.G[arr][en.y] : 0.f;
	const int depth = P.depth[ch];
	for (int d = 0; d < depth; d++) {
		const float c = P.cosk[ch][d], s = P.sink[ch][d];
		if (P.G[arr]) { const float nf = c*f + s*g, ng = c*g - s*f; f = nf; g = ng; }
		else f = c*f;
	}
	return make_float2(f, g);
}
static __device__ __forceinline__ void rc_store(const ResolvedCopy &P, const int4 en, const float2 v)
{
	const int arr = en.z >> 8;
	P.F[arr][en.x] = v.x; if (P.G[arr]) P.G[arr][en.x] = v.y;
}
/* grid: [one cluster for the bound entries, if any] + clusters that share the free entries (two per thread and round, both
 * loads in flight together) */
__global__ void __cluster_dims__(ALISB_LIST_CTAS, 1, 1) __launch_bounds__(1024) copy_resolved(const __grid_constant__ ResolvedCopy P)
{
	const int lead = P.nbound ? ALISB_LIST_CTAS : 0;
	if ((int)blockIdx.x >= lead) {
		const int tid = (blockIdx.x - lead)*blockDim.x + threadIdx.x, nth = (gridDim.x - lead)*blockDim.x;
		for (int t0 = P.nbound + tid; t0 < P.n; t0 += 2*nth) {
			const int t1 = t0 + nth;
			const int4 e0 = P.ent[t0]; int4 e1 = e0;
			if (t1 < P.n) e1 = P.ent[t1];
			const float2 v0 = rc_value(P, e0); float2 v1 = v0;
			if (t1 < P.n) v1 = rc_value(P, e1);
			rc_store(P, e0, v0);
			if (t1 < P.n) rc_store(P, e1, v1);
		}
		return;
	}
	const int tid = blockIdx.x*blockDim.x + threadIdx.x, nth = ALISB_LIST_CTAS*blockDim.x;
	if (P.nbound <= nth*ALISB_RC_PER_THREAD) {
		int4 en[ALISB_RC_PER_THREAD]; float2 v[ALISB_RC_PER_THREAD];
		#pragma unroll
		for (int u = 0; u < ALISB_RC_PER_THREAD; u++) { const int t = tid + u*nth; if (t < P.nbound) en[u] = P.ent[t]; }
		#pragma unroll
		for (int u = 0; u < ALISB_RC_PER_THREAD; u++) { const int t = tid + u*nth; if (t < P.nbound) v[u] = rc_value(P, en[u]); }
		list_sync();
		#pragma unroll
		for (int u = 0; u < ALISB_RC_PER_THREAD; u++) { const int t = tid + u*nth; if (t < P.nbound) rc_store(P, en[u], v[u]); }
		return;
	}
	for (int t = tid; t < P.nbound; t += nth) P.tmp[t] = rc_value(P, P.ent[t]);
	list_sync();
	for (int t = tid; t < P.nbound; t += nth) rc_store(P, P.ent[t], P.tmp[t]);
}

/* All soft sources of a pass (reference updateS list, update.c:161-179) and all TFSF strips of a
 * pass (update.c:205-272) in one single-CTA launch each, in list order (overlapping boxes add in
 * the reference's order). */
#define ALISB_MAXSRCOPS 16
struct SrcOp { float *F; Box B, full; const float *vals, *amp; float amplitude, f; };
struct SrcList { SrcOp op[ALISB_MAXSRCOPS]; int n, ny, pitch; };

__global__ void __cluster_dims__(ALISB_LIST_CTAS, 1, 1) __launch_bounds__(1024) source_list(const __grid_constant__ SrcList L)
{
	const int ny = L.ny, pitch = L.pitch;
	const size_t tid = (size_t)blockIdx.x*blockDim.x + threadIdx.x, nth = (size_t)gridDim.x*blockDim.x;
	for (int k = 0; k < L.n; k++) {
		const SrcOp &o = L.op[k];
		const size_t cells = (size_t)o.B.na*o.B.nb*o.B.nc;
		for (size_t t = tid; t < cells; t += nth) {
			int a, b, c; size_t q;
			box_cell(o.B, t, a, b, c, q);
			q = ((size_t)(c - o.full.c0)*o.full.nb + (b - o.full.b0))*o.full.na + (a - o.full.a0);
			const float v = o.vals ? o.vals[q] : (o.amp ? o.amp[q] : o.amplitude) * o.f;
			L.op[k].F[GIDX(a, b, c)] += v;
		}
		if (k + 1 < L.n) list_sync();
	}
}

#define ALISB_MAXSTRIPOPS 48
struct StripOp { float *F; const float *G, *C; float coef; Box B; int alt0, alt1, alt2, clampa, clampb; int sync_after; };
struct StripList { StripOp op[ALISB_MAXSTRIPOPS]; int n, ny, pitch, any, apitch; int nchain, first[ALISB_MAXCHAINS + 1]; };

template <int CL> static __device__ __forceinline__ void strip_chain(const StripList &L)
{
	const int ny = L.ny, pitch = L.pitch;
	const int chain = blockIdx.x / CL;
	const size_t tid = (size_t)(blockIdx.x % CL)*blockDim.x + threadIdx.x, nth = (size_t)CL*blockDim.x;
	for (int k = L.first[chain]; k < L.first[chain + 1]; k++) {
		const StripOp &o = L.op[k];
		const size_t cells = (size_t)o.B.na*o.B.nb*o.B.nc;
		for (size_t t = tid; t < cells; t += nth) {
			int a, b, c; size_t q;
			box_cell(o.B, t, a, b, c, q);
			const size_t i = GIDX(a, b, c);
			const int xa = min(a + o.alt0, o.clampa), xb = min(b + o.alt1, o.clampb), xc = c + o.alt2;
			const float g = o.G[((size_t)xc*(L.any + 2) + xb)*L.apitch + ALISB_XO + xa];
			if (o.C) o.F[i] += o.coef*g*o.C[i];
			else     o.F[i] += o.coef*g;
		}
		if (o.sync_after) chain_sync<CL>();     /* the next strip of the chain adds to cells this level added to (update.c:205-272: in list order) */
	}
}
__global__ void __cluster_dims__(ALISB_LIST_CTAS, 1, 1) __launch_bounds__(1024) strip_list(const __grid_constant__ StripList L) { strip_chain<ALISB_LIST_CTAS>(L); }
__global__ void __launch_bounds__(1024) strip_list1(const __grid_constant__ StripList L) { strip_chain<1>(L); }

/* The strip list regrouped per cell (engine.cu, build_strip_plan; strip_resolve.h): one thread per cell adds the cell's
 * terms in list order -- the same products and the same sequence of additions as the strips one after the other
 * (update.c:193-202), no ordering between cells, no barrier.  A cell record carries its first term (most cells have one). */
#define ALISB_RS_ARRAYS 8
struct RsCellDev { int tgt, meta, src; float coef; int more, pad; };   /* meta: farr | garr << 4 | (carr+1) << 8 | count << 12; more: index of the 2nd term */
struct RsTermDev { int src; float coef; int meta, pad; };              /* meta: garr | (carr+1) << 4 */
struct ResolvedStrips { float *F[ALISB_RS_ARRAYS]; const float *G[ALISB_RS_ARRAYS], *C[ALISB_RS_ARRAYS]; const RsCellDev *cells; const RsTermDev *terms; int n; };
__global__ void __launch_bounds__(256) strip_resolved(const __grid_constant__ ResolvedStrips P)
{
	const int t = blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= P.n) return;
	const RsCellDev c = P.cells[t];
	const int farr = c.meta & 15, garr = (c.meta >> 4) & 15, carr = ((c.meta >> 8) & 15) - 1, count = c.meta >> 12;
	float f = P.F[farr][c.tgt];
	{
		const float g = P.G[garr][c.src];
		if (carr >= 0) f += c.coef*g*P.C[carr][c.tgt];
		else           f += c.coef*g;
	}
	for (int k = 1; k < count; k++) {
		const RsTermDev tm = P.terms[c.more + k - 1];
		const int ga = tm.meta & 15, ca = ((tm.meta >> 4) & 15) - 1;
		const float g = P.G[ga][tm.src];
		if (ca >= 0) f += tm.coef*g*P.C[ca][c.tgt];
		else         f += tm.coef*g;
	}
	P.F[farr][c.tgt] = f;
}

/* gather / scatter a box between any engine array (element (a,b,c) at c*sC + b*sB + a + off) and a
 * dense buffer in the HOST order (z fastest, reference makeField layout): the transposition is
 * done here so that the host side is a plain memcpy */
__global__ void box_gather(const float *A, float *buf, Box B, long long sC, long long sB, long long off)
{
	int a, b, c; size_t q;
	if (!box_cell(B, (size_t)blockIdx.x*blockDim.x + threadIdx.x, a, b, c, q)) return;
	buf[((size_t)(a - B.a0)*B.nb + (b - B.b0))*B.nc + (c - B.c0)] = A[c*sC + b*sB + a + off];
}
__global__ void box_scatter(float *A, const float *buf, Box B, long long sC, long long sB, long long off)
{
	int a, b, c; size_t q;
	if (!box_cell(B, (size_t)blockIdx.x*blockDim.x + threadIdx.x, a, b, c, q)) return;
	A[c*sC + b*sB + a + off] = buf[((size_t)(a - B.a0)*B.nb + (b - B.b0))*B.nc + (c - B.c0)];
}

/* Small results go straight into mapped, pinned host memory ("mailbox"): the kernel stores the values, fences at system
 * scope and then publishes a sequence number; the host spins on that word instead of paying for a device-to-host copy
 * and a stream synchronisation (the launch-bound examples read one probe value or a few flux sums per step). */
__global__ void __launch_bounds__(256) box_gather_mail(const float *A, float *host_buf, Box B, long long sC, long long sB, long long off, volatile unsigned *flag, unsigned seq)
{
	const size_t n = (size_t)B.na*B.nb*B.nc;
	if (n == 1) {
		/* a single value (the probe of a `get`) travels together with its sequence number in ONE aligned 8-byte store,
		 * which cannot be torn: no fence, no second store (the host reads the pair at host_buf + 1022, see mail_wait) */
		if (threadIdx.x == 0) {
			const float v = A[B.c0*sC + B.b0*sB + B.a0 + off];
			*reinterpret_cast<volatile unsigned long long*>(host_buf + 1022) = ((unsigned long long)seq << 32) | __float_as_uint(v);
		}
		return;
	}
	for (size_t t = threadIdx.x; t < n; t += blockDim.x) {
		int a, b, c; size_t q;
		box_cell(B, t, a, b, c, q);
		host_buf[((size_t)(a - B.a0)*B.nb + (b - B.b0))*B.nc + (c - B.c0)] = A[c*sC + b*sB + a + off];
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) *flag = seq;           /* the end of the kernel flushes it */
}

/* reference phaser.c:83-110 (updatePhaser): the running DFT of one call -- every face of the monitor
 * box, every frequency of the list -- in ONE launch, one thread per (face cell, frequency).  The
 * phase factors exp(i w t_E), exp(i w t_H) are evaluated on the host with the caller's cexpf and
 * travel as kernel arguments (no staging copy, nothing to wait for).
 * acc layout per face: [freq][which 0..3][cell] float2, cells x fastest over the face box; the four
 * accumulators are E(c+1), E(c+2), H(c+1), H(c+2) for face normal c.  H is the average of the face
 * plane and the next one; the H terms reproduce the reference's promotion to double (0.5 is a
 * double constant).  Each accumulator is owned by exactly one thread: results do not depend on the
 * launch shape and are bit-identical to the reference's loops. */
#define ALISB_PHASER_MAXF 96      /* frequencies per launch (longer lists take several launches) */
struct PhaserParams {
	float2 *acc[6];
	const float *E1[6], *E2[6], *H1[6], *H2[6];
	Box B[6];
	unsigned long long hstep[6];      /* to the next plane along the face normal             */
	unsigned long long start[7];      /* first thread of each face: prefix sums of cells*nf   */
	int nfaces, nfreq, f0, nf;        /* this launch covers frequencies f0 .. f0+nf-1 of nfreq */
	int ny, pitch;
	float2 wE[ALISB_PHASER_MAXF], wH[ALISB_PHASER_MAXF];
};

__global__ void __launch_bounds__(256) phaser_accumulate(const __grid_constant__ PhaserParams p)
{
	const int ny = p.ny, pitch = p.pitch;
	unsigned long long t = (unsigned long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= p.start[p.nfaces]) return;
	int k = 0;
	while (k + 1 < p.nfaces && t >= p.start[k + 1]) k++;
	t -= p.start[k];
	const Box &B = p.B[k];
	const size_t cells = (size_t)B.na*B.nb*B.nc;
	const int fl = (int)(t / cells);                 /* frequency within this launch */
	int a, b, c; size_t q;
	if (!box_cell(B, (size_t)(t - (unsigned long long)fl*cells), a, b, c, q)) return;
	const size_t i = GIDX(a, b, c);
	const float e1 = p.E1[k][i], e2 = p.E2[k][i];
	const float h1 = p.H1[k][i] + p.H1[k][i + p.hstep[k]], h2 = p.H2[k][i] + p.H2[k][i + p.hstep[k]];
	const float2 wE = p.wE[fl], wH = p.wH[fl];
	float2 *A = p.acc[k] + (size_t)(p.f0 + fl)*4*cells + q;
	float2 v = A[0];       v.x += wE.x*e1; v.y += wE.y*e1; A[0] = v;                 /* phaser.c:87 */
	v = A[cells];          v.x += wE.x*e2; v.y += wE.y*e2; A[cells] = v;             /* phaser.c:88 */
	v = A[2*cells];                                                                /* phaser.c:89 */
	v.x = (float)((double)v.x + (double)(wH.x*h1)*0.5); v.y = (float)((double)v.y + (double)(wH.y*h1)*0.5); A[2*cells] = v;
	v = A[3*cells];                                                                /* phaser.c:90 */
	v.x = (float)((double)v.x + (double)(wH.x*h2)*0.5); v.y = (float)((double)v.y + (double)(wH.y*h2)*0.5); A[3*cells] = v;
}

/* ------------------------------------------------------------------------------------------
 * Poynting flux through rectangles (reference output.c:233-332).  blockIdx.y = plane.  Every
 * thread accumulates its cells in double, the block reduces with shuffles in a fixed order and
 * writes one partial per block; flux_finish adds the partials in index order, so the result is
 * deterministic.
 * ------------------------------------------------------------------------------------------ */
struct FluxPlane { int axis, at, b0, b1, c0, c1; };
struct FluxParams {
	const float *E[3], *H[3];
	const float *dE[3], *dH[3];       /* spacing tables in array coordinates */
	int ny, pitch;
};

static __device__ __forceinline__ size_t flux_idx(int axis, int at, int ib, int ic, int ny, int pitch)
{
	int q[3];
	q[axis] = at; q[(axis+1)%3] = ib; q[(axis+2)%3] = ic;
	return GIDX(q[0], q[1], q[2]);
}

__global__ void flux_partial(const __grid_constant__ FluxParams p, const FluxPlane *planes, double *partial)
{
	const FluxPlane P = planes[blockIdx.y];
	const int a = P.axis, b = (a+1)%3, c = (a+2)%3;
	const int ny = p.ny, pitch = p.pitch;
	const int nb = P.b1 - P.b0 + 1, nc = P.c1 - P.c0 + 1;
	const size_t step = a == 0 ? 1 : a == 1 ? (size_t)pitch : (size_t)(ny+2)*pitch;      /* to plane at+1 */
	double s1 = 0, s2 = 0;
	for (size_t t = (size_t)blockIdx.x*blockDim.x + threadIdx.x; t < (size_t)nb*nc; t += (size_t)gridDim.x*blockDim.x) {
		/* run fastest along whichever tangential axis is x (contiguous) if any */
		int ib, ic;
		if (b == 0) { ib = P.b0 + (int)(t % nb); ic = P.c0 + (int)(t / nb); }
		else        { ic = P.c0 + (int)(t % nc); ib = P.b0 + (int)(t / nc); }
		const size_t i = flux_idx(a, P.at, ib, ic, ny, pitch);
		if (ib > P.b0) {                       /* + E_b (H_c + H_c'), trapezoid along c */
			const float wc = (ic == P.c0 || ic == P.c1) ? 0.5f : 1.0f;
			s1 += (double)p.dH[b][ib] * ((double)wc * (double)p.dE[c][ic] * (double)p.E[b][i] * (double)(p.H[c][i] + p.H[c][i + step]));
		}
		if (ic > P.c0) {                       /* - E_c (H_b + H_b'), trapezoid along b */
			const float wb = (ib == P.b0 || ib == P.b1) ? 0.5f : 1.0f;
			s2 -= (double)p.dH[c][ic] * ((double)wb * (double)p.dE[b][ib] * (double)p.E[c][i] * (double)(p.H[b][i] + p.H[b][i + step]));
		}
	}
	__shared__ double sh[2][32];
	for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); s2 += __shfl_down_sync(0xffffffffu, s2, o); }
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) { sh[0][w] = s1; sh[1][w] = s2; }
	__syncthreads();
	if (w == 0) {
		s1 = l < (blockDim.x >> 5) ? sh[0][l] : 0; s2 = l < (blockDim.x >> 5) ? sh[1][l] : 0;
		for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); s2 += __shfl_down_sync(0xffffffffu, s2, o); }
		if (l == 0) { partial[2*((size_t)blockIdx.y*gridDim.x + blockIdx.x)] = s1; partial[2*((size_t)blockIdx.y*gridDim.x + blockIdx.x) + 1] = s2; }
	}
}

__global__ void flux_finish(const double *partial, int nblocks, double *out)
{
	/* the partials are fetched by all threads at once (one round trip instead of a chain of dependent loads), then added
	 * by one thread in index order, so the result does not depend on the launch shape */
	__shared__ double sh[2][256];
	const int q = blockIdx.x;
	double s1 = 0, s2 = 0;
	for (int k0 = 0; k0 < nblocks; k0 += 128) {
		const int n = min(128, nblocks - k0);
		for (int k = threadIdx.x; k < n; k += blockDim.x) { sh[0][k] = partial[2*((size_t)q*nblocks + k0 + k)]; sh[1][k] = partial[2*((size_t)q*nblocks + k0 + k) + 1]; }
		__syncthreads();
		if (threadIdx.x == 0) for (int k = 0; k < n; k++) { s1 += sh[0][k]; s2 += sh[1][k]; }
		__syncthreads();
	}
	if (threadIdx.x == 0) { out[2*q] = s1; out[2*q + 1] = s2; }
}

/* flux_partial + flux_finish in one launch, planes as kernel arguments, sums into the mailbox: the block that finishes a
 * plane last adds the plane's partials in index order (so the result equals flux_finish's, whatever the order in which the
 * blocks ran), and the last plane to finish publishes the sequence number */
#define ALISB_FLUX_MAIL_PLANES 32
struct FluxPlaneList { FluxPlane P[ALISB_FLUX_MAIL_PLANES]; };
__global__ void __launch_bounds__(256) flux_mail(const __grid_constant__ FluxParams p, const __grid_constant__ FluxPlaneList planes, double *partial, unsigned *counters,
                                                   double *host_out, volatile unsigned *flag, unsigned seq)
{
	const FluxPlane &P = planes.P[blockIdx.y];
	const int a = P.axis, b = (a+1)%3, c = (a+2)%3;
	const int ny = p.ny, pitch = p.pitch;
	const int nb = P.b1 - P.b0 + 1, nc = P.c1 - P.c0 + 1;
	const size_t step = a == 0 ? 1 : a == 1 ? (size_t)pitch : (size_t)(ny+2)*pitch;
	double s1 = 0, s2 = 0;
	for (size_t t = (size_t)blockIdx.x*blockDim.x + threadIdx.x; t < (size_t)nb*nc; t += (size_t)gridDim.x*blockDim.x) {
		int ib, ic;
		if (b == 0) { ib = P.b0 + (int)(t % nb); ic = P.c0 + (int)(t / nb); }
		else        { ic = P.c0 + (int)(t % nc); ib = P.b0 + (int)(t / nc); }
		const size_t i = flux_idx(a, P.at, ib, ic, ny, pitch);
		if (ib > P.b0) {
			const float wc = (ic == P.c0 || ic == P.c1) ? 0.5f : 1.0f;
			s1 += (double)p.dH[b][ib] * ((double)wc * (double)p.dE[c][ic] * (double)p.E[b][i] * (double)(p.H[c][i] + p.H[c][i + step]));
		}
		if (ic > P.c0) {
			const float wb = (ib == P.b0 || ib == P.b1) ? 0.5f : 1.0f;
			s2 -= (double)p.dH[c][ic] * ((double)wb * (double)p.dE[b][ib] * (double)p.E[c][i] * (double)(p.H[b][i] + p.H[b][i + step]));
		}
	}
	__shared__ double sh[2][128];
	__shared__ bool last;
	for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); s2 += __shfl_down_sync(0xffffffffu, s2, o); }
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) { sh[0][w] = s1; sh[1][w] = s2; }
	__syncthreads();
	if (w == 0) {
		s1 = l < (blockDim.x >> 5) ? sh[0][l] : 0; s2 = l < (blockDim.x >> 5) ? sh[1][l] : 0;
		for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); s2 += __shfl_down_sync(0xffffffffu, s2, o); }
		if (l == 0) {
			partial[2*((size_t)blockIdx.y*gridDim.x + blockIdx.x)] = s1; partial[2*((size_t)blockIdx.y*gridDim.x + blockIdx.x) + 1] = s2;
			__threadfence();
			last = atomicAdd(&counters[1 + blockIdx.y], 1u) == gridDim.x - 1;
		}
	}
	__syncthreads();
	if (!last) return;
	__threadfence();
	for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {          /* gridDim.x <= 128 */
		sh[0][k] = __ldcg(&partial[2*((size_t)blockIdx.y*gridDim.x + k)]); sh[1][k] = __ldcg(&partial[2*((size_t)blockIdx.y*gridDim.x + k) + 1]);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		s1 = 0; s2 = 0;
		for (int k = 0; k < (int)gridDim.x; k++) { s1 += sh[0][k]; s2 += sh[1][k]; }
		/* the plane's sums stay on the device; the plane that finishes last carries all of them to the host in one go: one
		 * system-scope fence per request instead of one per plane */
		double *dev_out = partial + 2*(size_t)gridDim.x*gridDim.y;
		dev_out[2*blockIdx.y] = s1; dev_out[2*blockIdx.y + 1] = s2;
		counters[1 + blockIdx.y] = 0;
		__threadfence();
		if (atomicAdd(&counters[0], 1u) == gridDim.y - 1) {
			counters[0] = 0;
			__threadfence();
			for (int q = 0; q < 2*(int)gridDim.y; q++) host_out[q] = __ldcg(&dev_out[q]);
			__threadfence_system();
			*flag = seq;                          /* the end of the kernel flushes it */
		}
	}
}

/* ------------------------------------------------------------------------------------------
 * Near-to-far-field surface sums (reference farfld.c:248-459, the double loops of farFieldX/Y/Z):
 *     out[job] = sum_v sum_u  Wlo[u] * A[u][v] * Whi[v]
 * over one phasor face, A = accumulator `which` at frequency `freq`, u the lower-numbered in-face
 * axis (fastest in the accumulator layout), v the higher one.  The trapezoid end weights and the
 * staggered ranges of the reference loops are folded into the weight vectors by the host.
 * One CTA per job; all (theta, phi, polarisation) queries of an angular sweep are jobs of one
 * launch.  Products are fp32 (A*Whi first, then *Wlo, no FMA), the sum is double and reduced in a
 * fixed order, so the result is deterministic.
 * ------------------------------------------------------------------------------------------ */
struct SurfJob { int face, which; unsigned w_lo, w_hi; };
struct SurfParams {
	const float2 *acc[6];             /* per face: [freq][which][cell]                       */
	int nlo[6], nhi[6];               /* face extents along the lower / higher in-face axis  */
	int freq;
};

static __device__ __forceinline__ float2 cmulf(float2 a, float2 b)
{
	return make_float2(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x);
}

__global__ void __launch_bounds__(128) surface_sums(const __grid_constant__ SurfParams p, const SurfJob *jobs, const float2 *W, double *out)
{
	const SurfJob J = jobs[blockIdx.x];
	const int nlo = p.nlo[J.face], nhi = p.nhi[J.face];
	const size_t cells = (size_t)nlo*nhi;
	const float2 *A = p.acc[J.face] + ((size_t)p.freq*4 + J.which)*cells;
	const float2 *Wlo = W + J.w_lo, *Whi = W + J.w_hi;
	double sr = 0, si = 0;
	for (size_t q = threadIdx.x; q < cells; q += blockDim.x) {
		const int v = (int)(q / nlo), u = (int)(q - (size_t)v*nlo);
		const float2 t = cmulf(cmulf(A[q], Whi[v]), Wlo[u]);
		sr += (double)t.x; si += (double)t.y;
	}
	__shared__ double sh[2][4];
	for (int o = 16; o > 0; o >>= 1) { sr += __shfl_down_sync(0xffffffffu, sr, o); si += __shfl_down_sync(0xffffffffu, si, o); }
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) { sh[0][w] = sr; sh[1][w] = si; }
	__syncthreads();
	if (threadIdx.x == 0) {
		out[2*(size_t)blockIdx.x]     = ((sh[0][0] + sh[0][1]) + sh[0][2]) + sh[0][3];
		out[2*(size_t)blockIdx.x + 1] = ((sh[1][0] + sh[1][1]) + sh[1][2]) + sh[1][3];
	}
}

/* ------------------------------------------------------------------------------------------
 * Device-side monitors: the reference's point-wise field functions
 * (fields.c:130-681, the ones that read E/H only) evaluated where the fields live, so that a
 * slice moves one float per pixel and a world reduction one number (or one float per z column)
 * instead of up to six field arrays.  The arithmetic is in monitor_eval.h (shared with the CPU
 * oracle); threads run x fastest for coalesced loads.
 * ------------------------------------------------------------------------------------------ */
struct EvalParams {
	const float *F[2][2][3];          /* [set][m][comp]                                     */
	const float *dtd[3];              /* CE->dtdx (shifted by XO), dtdy, dtdz                */
	const float *dz;                  /* CE->dz, array coordinates (column sums)             */
	alisb_field_fn f;
	int ny, pitch, complex_field;
	float dt;
	int i0, j0, k0, ni, nj, nk;       /* the box, world INDEX coordinates                    */
	/* objectAbsorption columns only: */
	const float *inv[3], *fill[3], *J[ALISB_MAXPOLES][3];
	const float *dzH;
	float loss[3]; int npoles;
	Box full;                         /* the object's dense ADE box                          */
};
#define MON_FN static __device__ __forceinline__
#define MON_CTX EvalParams
#define MON_DESC(X) (&(X)->f)
#define MON_LD(X,set,m,c,a,b,cc) ((X)->F[set][m][c][((size_t)(cc)*((X)->ny+2) + (size_t)(b))*(X)->pitch + ALISB_XO + (a)])
#define MON_DTD(X,axis,q) ((axis) == 0 ? (X)->dtd[0][ALISB_XO + (q)] : (X)->dtd[axis][q])
#define MON_DT(X) ((X)->dt)
#define MON_COMPLEX(X) ((X)->complex_field)
#define MON_GIDX(X,a,b,cc) (((size_t)(cc)*((X)->ny+2) + (size_t)(b))*(X)->pitch + ALISB_XO + (a))
#define MON_OIDX(X,a,b,cc) (((size_t)((cc) - (X)->full.c0)*(X)->full.nb + ((b) - (X)->full.b0))*(X)->full.na + ((a) - (X)->full.a0))
#define MON_INV(X,c,a,b,cc) ((X)->inv[c][MON_GIDX(X,a,b,cc)])
#define MON_FILL(X,c,a,b,cc) ((X)->fill[c][MON_OIDX(X,a,b,cc)])
#define MON_J(X,p,c,a,b,cc) ((X)->J[p][c][MON_OIDX(X,a,b,cc)])
#define MON_NPOLES(X) ((X)->npoles)
#define MON_LOSS(X,c) ((X)->loss[c])
#define MON_DZE(X,cc) ((X)->dz[cc])
#define MON_DZH(X,cc) ((X)->dzH[cc])
#include "monitor_eval.h"

/* out: dense block in the HOST order (z fastest), like box_gather */
__global__ void eval_box(const __grid_constant__ EvalParams p, float *out)
{
	const size_t t = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= (size_t)p.ni*p.nj*p.nk) return;
	const int ia = (int)(t % p.ni); const size_t r = t / p.ni;
	const int ib = (int)(r % p.nj), ic = (int)(r / p.nj);
	out[((size_t)ia*p.nj + ib)*p.nk + ic] = mon_eval(&p, p.i0 + ia, p.j0 + ib, p.k0 + ic);
}

/* running sums over a box (sliceTimeAvg / sliceFreqDom accumulations, slices.c:282-379): re += F*wr, im += F*wi,
 * accumulators dense in the HOST order like eval_box's output; one thread owns one point */
__global__ void eval_accumulate(const __grid_constant__ EvalParams p, float *re, float *im, float wr, float wi)
{
	const size_t t = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= (size_t)p.ni*p.nj*p.nk) return;
	const int ia = (int)(t % p.ni); const size_t r = t / p.ni;
	const int ib = (int)(r % p.nj), ic = (int)(r / p.nj);
	const float v = mon_eval(&p, p.i0 + ia, p.j0 + ib, p.k0 + ic);
	const size_t q = ((size_t)ia*p.nj + ib)*p.nk + ic;
	re[q] += v * wr;
	if (im) im[q] += v * wi;
}

/* max |F| over the box (worldMax output.c:56-70, sliceMax slices.c:382-394): order independent;
 * NaNs are skipped exactly like the reference's `if (max < f)`.  *out must be zeroed. */
__global__ void reduce_absmax(const __grid_constant__ EvalParams p, unsigned *out)
{
	const size_t n = (size_t)p.ni*p.nj*p.nk;
	float m = 0;
	for (size_t t = (size_t)blockIdx.x*blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x*blockDim.x) {
		const int ia = (int)(t % p.ni); const size_t r = t / p.ni;
		const int ib = (int)(r % p.nj), ic = (int)(r / p.nj);
		const float v = fabsf(mon_eval(&p, p.i0 + ia, p.j0 + ib, p.k0 + ic));
		if (m < v) m = v;
	}
	for (int o = 16; o > 0; o >>= 1) { const float u = __shfl_down_sync(0xffffffffu, m, o); if (m < u) m = u; }
	if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, __float_as_uint(m));     /* non-negative floats order like their bit patterns */
}

/* one thread per (i,j): the z column sum of worldSum / sliceSum in the reference's order */
__global__ void column_sums(const __grid_constant__ EvalParams p, float *out)
{
	const size_t t = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= (size_t)p.ni*p.nj) return;
	const int ia = (int)(t % p.ni), ib = (int)(t / p.ni);
	float s;
	MON_COLUMN_SUM(&p, p.dz, p.i0 + ia, p.j0 + ib, p.k0, p.k0 + p.nk - 1, s);
	out[(size_t)ia*p.nj + ib] = s;
}

/* objectAbsorption (output.c:365-435): one thread per (a,b) column of the object's box, (xs, ys, zs)
 * per column; the host applies the reference's edge weights and sums in its order.  Box: array
 * coordinates (i0,j0 = first column; k0..k0+nk-1 = planes); cmin/cmax: the world's first/last plane. */
__global__ void absorption_columns(const __grid_constant__ EvalParams p, int cmin, int cmax, float *out)
{
	const size_t t = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= (size_t)p.ni*p.nj) return;
	const int ia = (int)(t % p.ni), ib = (int)(t / p.ni);
	const int a = p.i0 + ia, b = p.j0 + ib;
	float xs, ys, zs;
	MON_ABSORPTION_COLUMN(&p, a, b, p.k0, p.k0 + p.nk - 1, cmin, cmax, xs, ys, zs);
	float *o = out + 3*((size_t)ia*p.nj + ib);
	o[0] = xs; o[1] = ys; o[2] = zs;
}

/* ------------------------------------------------------------------------------------------
 * Geometry rasterisation: the shape tree interpreter of geometry_eval.h (shared
 * with the CPU oracle) over the cells of a box.  Recursive device functions: the engine raises the
 * per-thread stack limit around these launches.
 * ------------------------------------------------------------------------------------------ */
#define GEO_FN static __device__
#include "geometry_eval.h"

/* object.c:229-262: first object owning the node (i,j,k) -> atomic min/max of its index box */
__global__ void geometry_bounds(const alisb_geometry g, int i0, int j0, int k0, int ni, int nj, int nk, int *bounds)
{
	const size_t t = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= (size_t)ni*nj*nk) return;
	const int ia = (int)(t % ni); const size_t r = t / ni;
	const int i = i0 + ia, j = j0 + (int)(r % nj), k = k0 + (int)(r / nj);
	const float x = geo_coord(&g, 0, i, 0, 0), y = geo_coord(&g, 1, j, 0, 0), z = geo_coord(&g, 2, k, 0, 0);
	int o = 0;
	while (o < g.nobjects && !geo_inside(g.nodes, g.aux, g.roots[o], x, y, z)) o++;
	if (o == g.nobjects) return;
	int *b = bounds + 6*o;
	atomicMin(&b[0], i); atomicMax(&b[1], i); atomicMin(&b[2], j); atomicMax(&b[3], j); atomicMin(&b[4], k); atomicMax(&b[5], k);
}

/* one thread per (cell, component); out[comp] dense over the box in the HOST order (z fastest) */
__global__ void geometry_fill(const alisb_geometry g, int object, int i0, int j0, int k0, int ni, int nj, int nk, float *fx, float *fy, float *fz)
{
	const size_t n = (size_t)ni*nj*nk;
	const size_t t = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
	if (t >= 3*n) return;
	const int comp = (int)(t / n); const size_t q = t - (size_t)comp*n;
	const int ia = (int)(q % ni); const size_t r = q / ni;
	const int ib = (int)(r % nj), ic = (int)(r / nj);
	const int idx[3] = { i0 + ia, j0 + ib, k0 + ic };
	const int half[3] = { comp == 0, comp == 1, comp == 2 };
	float *out = comp == 0 ? fx : comp == 1 ? fy : fz;
	out[((size_t)ia*nj + ib)*nk + ic] = geo_fill(&g, object, idx, half);
}
