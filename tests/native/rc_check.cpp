/*
 * rc_check -- CPU check of alis_b200/csrc/copy_resolve.h: random ordered copy lists (overlapping planes, Bloch rotations on
 * complex pairs, plain factors on real arrays) are applied (a) one after the other as the reference does (copyBoundary,
 * update.c:182-191) and (b) through the resolved plan the way kernel copy_resolved does (every value formed from the state
 * BEFORE the list, then stored).  Both must agree bit for bit.  Also walks the periodic / Bloch / hexagonal lists of
 * update.c:284-297 and 311-339 on a small grid.  Built with -ffp-contract=off (the engine is built with -fmad=false).
 * usage: rc_check [seeds]     prints "ok <cases> <resolved> <refused>" and exits 0, or the first difference and exits 1
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "../../alis_b200/csrc/copy_resolve.h"
#include "../../alis_b200/csrc/strip_resolve.h"

struct Op { int arr; float cosk, sink; RcOp r; };
static const int XO = 31;
static int NX, NY, NZ, PITCH; static size_t ELEMS;
static size_t idx(int a, int b, int c) { return ((size_t)c*(NY+2) + b)*PITCH + XO + a; }

static uint64_t rng_state;
static uint32_t rnd() { rng_state = rng_state*6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }
static int rint_(int lo, int hi) { return lo + (int)(rnd() % (uint32_t)(hi - lo + 1)); }
static float rflt() { return (float)((int)(rnd() % 20001) - 10000) / 8192.0f; }

static void apply_in_order(const std::vector<Op> &ops, std::vector<std::vector<float> > &F, std::vector<std::vector<float> > &G, const std::vector<bool> &cplx)
{
	for (const Op &o : ops) {
		const RcOp &r = o.r;
		for (int c=r.c0; c<r.c0+r.nc; c++) for (int b=r.b0; b<r.b0+r.nb; b++) for (int a=r.a0; a<r.a0+r.na; a++) {
			const size_t i = idx(a, b, c), j = idx(a + r.alt0, b + r.alt1, c + r.alt2);
			if (cplx[o.arr]) {
				const float f = F[o.arr][j], g = G[o.arr][j];
				F[o.arr][i] = o.cosk*f + o.sink*g;
				G[o.arr][i] = o.cosk*g - o.sink*f;
			} else F[o.arr][i] = o.cosk*F[o.arr][j];
		}
	}
}

/* the way kernel copy_resolved works: the bound entries form their values from the state before the list and store them after a
 * barrier; the free entries run whenever (here: before or after the bound ones, in reverse order, each one load-then-store) */
static void apply_resolved(const std::vector<Op> &ops, const RcPlan &P, std::vector<std::vector<float> > &F, std::vector<std::vector<float> > &G, const std::vector<bool> &cplx, const std::vector<int> &arr_of_plan, bool free_first)
{
	auto value = [&](const RcEntry &en, float &f, float &g) {
		const int arr = arr_of_plan[en.arr];
		f = F[arr][en.src]; g = cplx[arr] ? G[arr][en.src] : 0.f;
		for (int q : P.chains[en.chain]) {
			const float c = ops[q].cosk, s = ops[q].sink;
			if (cplx[arr]) { const float nf = c*f + s*g, ng = c*g - s*f; f = nf; g = ng; }
			else f = c*f;
		}
	};
	auto store = [&](const RcEntry &en, float f, float g) { const int arr = arr_of_plan[en.arr]; F[arr][en.tgt] = f; if (cplx[arr]) G[arr][en.tgt] = g; };
	auto run_free = [&]() { for (size_t t=P.ent.size(); t-- > P.nbound; ) { float f, g; value(P.ent[t], f, g); store(P.ent[t], f, g); } };
	if (free_first) run_free();
	std::vector<float> vf(P.nbound), vg(P.nbound);
	for (size_t t=0; t<P.nbound; t++) value(P.ent[t], vf[t], vg[t]);
	for (size_t t=0; t<P.nbound; t++) store(P.ent[t], vf[t], vg[t]);
	if (!free_first) run_free();
}

static size_t g_bound, g_entries;
static int run_case(const std::vector<Op> &ops, int narr, const std::vector<bool> &cplx, int &resolved, int &refused, const char *what)
{
	std::vector<std::vector<float> > F(narr, std::vector<float>(ELEMS)), G(narr, std::vector<float>(ELEMS));
	for (int q=0; q<narr; q++) for (size_t i=0; i<ELEMS; i++) { F[q][i] = rflt(); G[q][i] = rflt(); }
	std::vector<std::vector<float> > F2 = F, G2 = G, F3 = F, G3 = G;
	std::vector<RcOp> r(ops.size());
	for (size_t q=0; q<ops.size(); q++) { r[q] = ops[q].r; r[q].F = &F[ops[q].arr]; r[q].G = cplx[ops[q].arr] ? (const void*)&G[ops[q].arr] : 0; }
	RcPlan P;
	if (!rc_resolve(r.data(), (int)r.size(), NY, PITCH, XO, 8, 48, 4, P)) { refused++; return 0; }
	resolved++;
	std::vector<int> arr_of_plan(P.F.size());
	for (size_t k=0; k<P.F.size(); k++) for (int q=0; q<narr; q++) if (P.F[k] == (const void*)&F[q]) arr_of_plan[k] = q;
	apply_in_order(ops, F, G, cplx);
	apply_resolved(ops, P, F2, G2, cplx, arr_of_plan, true);
	apply_resolved(ops, P, F3, G3, cplx, arr_of_plan, false);
	g_bound += P.nbound; g_entries += P.ent.size();
	for (int q=0; q<narr; q++) {
		if (memcmp(F[q].data(), F2[q].data(), ELEMS*sizeof(float)) || memcmp(G[q].data(), G2[q].data(), ELEMS*sizeof(float))
		    || memcmp(F[q].data(), F3[q].data(), ELEMS*sizeof(float)) || memcmp(G[q].data(), G3[q].data(), ELEMS*sizeof(float))) {
			printf("%s: array %d differs between the ordered copies and the resolved plan (%zu ops, %zu entries, %zu chains)\n", what, q, ops.size(), P.ent.size(), P.chains.size());
			return 1;
		}
	}
	return 0;
}

static Op mk(int arr, float c, float s, int i0, int i1, int j0, int j1, int k0, int k1, int ia, int ja, int ka)
{
	Op o; memset(&o, 0, sizeof o);
	o.arr = arr; o.cosk = c; o.sink = s;
	o.r.a0 = i0; o.r.b0 = j0; o.r.c0 = k0; o.r.na = i1-i0+1; o.r.nb = j1-j0+1; o.r.nc = k1-k0+1; o.r.alt0 = ia; o.r.alt1 = ja; o.r.alt2 = ka;
	return o;
}

int main(int argc, char **argv)
{
	const int seeds = argc > 1 ? atoi(argv[1]) : 200;
	int cases = 0, resolved = 0, refused = 0;
	NX = 9; NY = 7; NZ = 6; PITCH = 64; ELEMS = (size_t)(NZ+2)*(NY+2)*PITCH;
	/* (1) the H-pass list of a world periodic in x, y and z (update.c:311-339, iAlt = iNUM-1 etc.) with Bloch rotations, and the
	 *     hexagonal variant (HBC along y: update.c:322-331); array coordinates: iMIN = 1, iMAX = NX, ghost 0 and NX+1 */
	for (int hex=0; hex<2; hex++) for (int cp=0; cp<2; cp++) {
		rng_state = 99 + hex*2 + cp;
		const int iMIN = 1, iMAX = NX, jMIN = 1, jMAX = NY, kMIN = 1, kMAX = NZ, iAlt = NX-1, jAlt = NY-1, kAlt = NZ-1, jNUM = NY;
		const float cx = 0.8f, sx = 0.6f, cy = 0.28f, sy = 0.96f, cz = 0.6f, sz = -0.8f, cyp = 0.96f, syp = 0.28f, cym = -0.6f, sym = 0.8f;
		std::vector<Op> L;
		enum { X = 0, Y = 1, Z = 2 };
		L.push_back(mk(Y, cx, -sx, iMAX+1, iMAX+1, jMIN, jMAX+1, kMIN, kMAX+1, -iAlt, 0, 0));
		L.push_back(mk(Z, cx, -sx, iMAX+1, iMAX+1, jMIN, jMAX+1, kMIN, kMAX+1, -iAlt, 0, 0));
		L.push_back(mk(X, cz, -sz, iMIN, iMAX+1, jMIN, jMAX+1, kMAX+1, kMAX+1, 0, 0, -kAlt));
		L.push_back(mk(Y, cz, -sz, iMIN, iMAX+1, jMIN, jMAX+1, kMAX+1, kMAX+1, 0, 0, -kAlt));
		L.push_back(mk(Y, cx, sx, iMIN, iMIN, jMIN, jMAX+1, kMIN, kMAX+1, iAlt, 0, 0));
		L.push_back(mk(Z, cx, sx, iMIN, iMIN, jMIN, jMAX+1, kMIN, kMAX+1, iAlt, 0, 0));
		L.push_back(mk(X, cz, sz, iMIN, iMAX+1, jMIN, jMAX+1, kMIN, kMIN, 0, 0, kAlt));
		L.push_back(mk(Y, cz, sz, iMIN, iMAX+1, jMIN, jMAX+1, kMIN, kMIN, 0, 0, kAlt));
		if (hex) {
			const int half = NX/2;            /* the hexagonal lists shift by half a period along x: (iMIN..0) in the reference's centred coordinates */
			L.push_back(mk(X, cyp, syp, iMIN, half, jMIN, jMIN, kMIN, kMAX, NX-half, jNUM-1, 0));
			L.push_back(mk(Z, cyp, syp, iMIN, half, jMIN, jMIN, kMIN, kMAX, NX-half, jNUM-1, 0));
			L.push_back(mk(X, cym, -sym, iMIN, half, jMAX+1, jMAX+1, kMIN, kMAX, NX-half, 1-jNUM, 0));
			L.push_back(mk(Z, cym, -sym, iMIN, half, jMAX+1, jMAX+1, kMIN, kMAX, NX-half, 1-jNUM, 0));
			L.push_back(mk(X, cym, sym, half+1, iMAX+1, jMIN, jMIN, kMIN, kMAX, -half, jNUM-1, 0));
			L.push_back(mk(Z, cym, sym, half+1, iMAX+1, jMIN, jMIN, kMIN, kMAX, -half, jNUM-1, 0));
			L.push_back(mk(X, cyp, -syp, half+1, iMAX+1, jMAX+1, jMAX+1, kMIN, kMAX, -half, 1-jNUM, 0));
			L.push_back(mk(Z, cyp, -syp, half+1, iMAX+1, jMAX+1, jMAX+1, kMIN, kMAX, -half, 1-jNUM, 0));
		} else {
			L.push_back(mk(Z, cy, -sy, iMIN, iMAX+1, jMAX+1, jMAX+1, kMIN, kMAX+1, 0, -jAlt, 0));
			L.push_back(mk(X, cy, -sy, iMIN, iMAX+1, jMAX+1, jMAX+1, kMIN, kMAX+1, 0, -jAlt, 0));
			L.push_back(mk(Z, cy, sy, iMIN, iMAX+1, jMIN, jMIN, kMIN, kMAX+1, 0, jAlt, 0));
			L.push_back(mk(X, cy, sy, iMIN, iMAX+1, jMIN, jMIN, kMIN, kMAX+1, 0, jAlt, 0));
		}
		std::vector<bool> cplx(3, cp != 0);
		cases++;
		const int before = resolved;
		if (run_case(L, 3, cplx, resolved, refused, hex ? "hexagonal H list" : "periodic H list")) return 1;
		if (resolved == before) { printf("%s (complex %d) was refused: it must resolve\n", hex ? "hexagonal H list" : "periodic H list", cp); return 1; }
	}
	/* (2) random lists: planes and small boxes anywhere in the padded grid, shifts along any axes */
	for (int seed=0; seed<seeds; seed++) {
		rng_state = 1000 + seed;
		NX = rint_(3, 12); NY = rint_(3, 9); NZ = rint_(3, 8); PITCH = 64; ELEMS = (size_t)(NZ+2)*(NY+2)*PITCH;
		const int narr = rint_(1, 4);
		std::vector<bool> cplx(narr);
		for (int q=0; q<narr; q++) cplx[q] = rnd() & 1;
		const int nops = rint_(2, 24);
		std::vector<Op> L;
		for (int q=0; q<nops; q++) {
			int lo[3], hi[3], alt[3];
			const int N[3] = {NX, NY, NZ};
			const int thin = rint_(0, 2);
			for (int ax=0; ax<3; ax++) {
				if (ax == thin) { lo[ax] = hi[ax] = rint_(0, N[ax]+1); }
				else { lo[ax] = rint_(0, N[ax]); hi[ax] = rint_(lo[ax], N[ax]+1); }
				alt[ax] = 0;
			}
			/* shift along the thin axis (never onto itself), sometimes along a second axis too */
			do { alt[thin] = rint_(-lo[thin], N[thin]+1-hi[thin]); } while (alt[thin] == 0);
			if (rnd() % 4 == 0) { const int ax = (thin + 1) % 3; alt[ax] = rint_(-lo[ax], N[ax]+1-hi[ax]); }
			L.push_back(mk(rint_(0, narr-1), rflt(), rflt(), lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], alt[0], alt[1], alt[2]));
		}
		cases++;
		if (run_case(L, narr, cplx, resolved, refused, "random list")) { printf("seed %d\n", seed); return 1; }
	}
	/* (3) strip lists (strip_resolve.h): overlapping face strips added one after the other (update(), update.c:193-202) against
	 *     the per-cell regrouping that kernel strip_resolved applies */
	size_t strip_cells = 0, strip_multi = 0;
	for (int seed=0; seed<seeds; seed++) {
		rng_state = 5000 + seed;
		NX = rint_(4, 12); NY = rint_(4, 9); NZ = rint_(4, 8); PITCH = 64; ELEMS = (size_t)(NZ+2)*(NY+2)*PITCH;
		const int ANY = rint_(1, NY), APITCH = 64, ANZ = NZ;                 /* auxiliary world: fewer rows, same planes */
		const size_t AELEMS = (size_t)(ANZ+2)*(ANY+2)*APITCH;
		const int nf = rint_(1, 3), ng = rint_(1, 3), nc = rint_(1, 2);
		std::vector<std::vector<float> > F(nf, std::vector<float>(ELEMS)), G(ng, std::vector<float>(AELEMS)), C(nc, std::vector<float>(ELEMS));
		for (auto &v : F) for (float &x : v) x = rflt();
		for (auto &v : G) for (float &x : v) x = rflt();
		for (auto &v : C) for (float &x : v) x = rflt();
		std::vector<std::vector<float> > F2 = F;
		const int nops = rint_(2, 30);
		std::vector<RsOp> ops(nops);
		for (int q=0; q<nops; q++) {
			RsOp &o = ops[q];
			const int N[3] = {NX, NY, NZ};
			int lo[3], hi[3];
			const int thin = rint_(0, 2);
			for (int ax=0; ax<3; ax++) {
				if (ax == thin) lo[ax] = hi[ax] = rint_(1, N[ax]);
				else { lo[ax] = rint_(1, N[ax]); hi[ax] = rint_(lo[ax], N[ax]); }
			}
			o.F = &F[rint_(0, nf-1)]; o.G = &G[rint_(0, ng-1)]; o.C = (rnd() & 1) ? (const void*)&C[rint_(0, nc-1)] : 0; o.coef = rflt();
			o.a0 = lo[0]; o.b0 = lo[1]; o.c0 = lo[2]; o.na = hi[0]-lo[0]+1; o.nb = hi[1]-lo[1]+1; o.nc = hi[2]-lo[2]+1;
			o.alt0 = rint_(0, 1); o.alt1 = rint_(-1, 1); o.alt2 = rint_(-1, 1);
			o.clampa = rint_(1, 31); o.clampb = ANY;                           /* rows beyond the auxiliary world clamp to its last row */
			if (o.b0 + o.alt1 < 0) o.alt1 = 0;
		}
		/* in order, as kernel strip_list / the reference */
		for (const RsOp &o : ops) {
			std::vector<float> &f = *(std::vector<float>*)o.F; const std::vector<float> &g = *(const std::vector<float>*)o.G;
			for (int c=o.c0; c<o.c0+o.nc; c++) for (int b=o.b0; b<o.b0+o.nb; b++) for (int a=o.a0; a<o.a0+o.na; a++) {
				const size_t i = idx(a, b, c);
				const int xa = std::min(a + o.alt0, o.clampa), xb = std::min(b + o.alt1, o.clampb), xc = c + o.alt2;
				const float gv = g[((size_t)xc*(ANY+2) + xb)*APITCH + XO + xa];
				if (o.C) f[i] += o.coef*gv*(*(const std::vector<float>*)o.C)[i];
				else     f[i] += o.coef*gv;
			}
		}
		RsPlan P;
		cases++;
		if (!rs_resolve(ops.data(), nops, NY, PITCH, ANY, APITCH, XO, 8, 15, P)) { refused++; continue; }
		resolved++;
		for (size_t t=P.cells.size(); t-- > 0; ) {                          /* any order: here from the last cell to the first */
			const RsCell &cl = P.cells[t];
			int fa = -1;
			for (int q=0; q<nf; q++) if (P.F[cl.farr] == (const void*)&F[q]) fa = q;
			float f = F2[fa][cl.tgt];
			for (int k=0; k<cl.count; k++) {
				const RsTerm &tm = P.terms[cl.first + k];
				const float gv = (*(const std::vector<float>*)P.G[tm.garr])[tm.src];
				if (tm.carr >= 0) f += tm.coef*gv*(*(const std::vector<float>*)P.C[tm.carr])[cl.tgt];
				else              f += tm.coef*gv;
			}
			F2[fa][cl.tgt] = f;
			strip_cells++; if (cl.count > 1) strip_multi++;
		}
		for (int q=0; q<nf; q++) if (memcmp(F[q].data(), F2[q].data(), ELEMS*sizeof(float))) { printf("strip list seed %d: array %d differs between the ordered strips and the regrouped plan\n", seed, q); return 1; }
	}
	if (!strip_multi) { printf("no cell with more than one strip term was exercised\n"); return 1; }
	if (resolved < cases/2) { printf("only %d of %d lists resolved\n", resolved, cases); return 1; }
	printf("ok %d %d %d bound %zu of %zu strips %zu of %zu\n", cases, resolved, refused, g_bound, g_entries, strip_multi, strip_cells);
	return 0;
}
