/*
 * strip_resolve.h -- the ordered TFSF strip list of a pass, regrouped per cell (host side, no CUDA in here so that it
 * can be checked on the CPU: tests/native/rc_check.cpp).
 *
 * The reference adds the incident field on the faces of the total-field box strip by strip (update(), update.c:193-202,
 * called in the order of update.c:209-234 and 244-269): F[i][j][k] += dtds * G[i+iAlt][..][..] (* C[i][j][k]).  Strips
 * of one component overlap on the box edges, where a cell receives two or three terms, and float addition is not
 * associative: the order matters.  The terms only READ the auxiliary world and the inverse-permittivity arrays, which no
 * strip writes, so regrouping by target cell -- the cell's terms kept in list order -- gives the same sums with no
 * ordering between cells at all: one thread per cell, no barrier (kernels.cuh, strip_resolved).
 */
#ifndef ALISB_STRIP_RESOLVE_H
#define ALISB_STRIP_RESOLVE_H
#include <vector>
#include <map>
#include <algorithm>

struct RsOp { const void *F, *G, *C; float coef; int a0, b0, c0, na, nb, nc, alt0, alt1, alt2, clampa, clampb; };
struct RsTerm { int src; float coef; int garr, carr; };        /* carr < 0: no factor array */
struct RsCell { int tgt, farr, first, count; };
struct RsPlan {
	std::vector<const void*> F, G, C;
	std::vector<RsCell> cells;        /* sorted by (farr, tgt) */
	std::vector<RsTerm> terms;        /* the terms of a cell are contiguous and in list order */
};

static inline int rs_array(std::vector<const void*> &tab, const void *p, size_t cap)
{
	for (size_t q=0; q<tab.size(); q++) if (tab[q] == p) return (int)q;
	if (tab.size() == cap) return -1;
	tab.push_back(p);
	return (int)tab.size() - 1;
}

/* target cell (a,b,c) of the world: ((c*(ny+2) + b)*pitch + xo + a); source cell of the auxiliary world (rows of apitch floats,
 * any+2 rows per plane): a + alt0 and b + alt1 clamped from above (1-D and 2-D auxiliary worlds), as kernel strip_list does */
static inline bool rs_resolve(const RsOp *ops, int n, int ny, int pitch, int any, int apitch, int xo, size_t max_arrays, size_t max_terms_per_cell, RsPlan &out)
{
	out.F.clear(); out.G.clear(); out.C.clear(); out.cells.clear(); out.terms.clear();
	std::vector<std::map<int, std::vector<RsTerm> > > acc;
	for (int q=0; q<n; q++) {
		const RsOp &o = ops[q];
		const int farr = rs_array(out.F, o.F, max_arrays), garr = rs_array(out.G, o.G, max_arrays), carr = o.C ? rs_array(out.C, o.C, max_arrays) : -1;
		if (farr < 0 || garr < 0 || (o.C && carr < 0)) return false;
		for (size_t k=0; k<out.F.size(); k++) if (out.F[k] == o.G || out.F[k] == o.C) return false;      /* a strip must not read what strips write */
		for (size_t k=0; k<out.G.size(); k++) if (out.G[k] == o.F) return false;
		for (size_t k=0; k<out.C.size(); k++) if (out.C[k] == o.F) return false;
		if ((int)acc.size() <= farr) acc.resize(farr + 1);
		for (int c=o.c0; c<o.c0+o.nc; c++) for (int b=o.b0; b<o.b0+o.nb; b++) for (int a=o.a0; a<o.a0+o.na; a++) {
			const long long i = ((long long)c*(ny+2) + b)*pitch + xo + a;
			const int xa = std::min(a + o.alt0, o.clampa), xb = std::min(b + o.alt1, o.clampb), xc = c + o.alt2;
			const long long j = ((long long)xc*(any+2) + xb)*apitch + xo + xa;
			if (i < 0 || j < 0 || i > 0x7fffffffLL || j > 0x7fffffffLL) return false;
			std::vector<RsTerm> &t = acc[farr][(int)i];
			if (t.size() == max_terms_per_cell) return false;
			RsTerm tm = { (int)j, o.coef, garr, carr };
			t.push_back(tm);
		}
	}
	for (size_t farr=0; farr<acc.size(); farr++)
		for (std::map<int, std::vector<RsTerm> >::const_iterator kv = acc[farr].begin(); kv != acc[farr].end(); ++kv) {
			RsCell cell = { kv->first, (int)farr, (int)out.terms.size(), (int)kv->second.size() };
			out.cells.push_back(cell);
			out.terms.insert(out.terms.end(), kv->second.begin(), kv->second.end());
		}
	return true;
}
#endif
