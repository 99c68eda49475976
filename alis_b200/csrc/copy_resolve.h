/*
 * copy_resolve.h -- symbolic walk of an ordered boundary-copy list (host side, no CUDA in here so that the
 * walk can be checked on the CPU: tests/native/rc_check.cpp).
 *
 * The reference applies its boundary copies one after the other (copyBoundary, update.c:182-191, called in the
 * order of update.c:284-297 and 311-339); a later copy may read what an earlier one wrote (the ghost edges and
 * corners of a periodic world pass through two or three copies).  Walking the list once with symbols instead of
 * values yields, for every cell that any copy writes, (a) the cell whose value from BEFORE the list it finally
 * carries and (b) the copies that value passed through, in order.  The device then needs one barrier for the
 * whole list (kernels.cuh, copy_resolved) and applies exactly the same float operations in the same order.
 */
#ifndef ALISB_COPY_RESOLVE_H
#define ALISB_COPY_RESOLVE_H
#include <vector>
#include <map>
#include <unordered_map>
#include <algorithm>
#include <utility>

struct RcOp { const void *F, *G; int a0, b0, c0, na, nb, nc, alt0, alt1, alt2; };   /* target box, source = target + alt; G: imaginary twin or 0 */
struct RcEntry { int tgt, src, arr, chain; };
struct RcPlan {
	std::vector<const void*> F, G;                   /* the arrays (pairs) the list acts on */
	std::vector<std::vector<int> > chains;           /* chains[0] is empty; chains[k]: indices of the ops a value passed through, in order */
	std::vector<RcEntry> ent;                        /* the bound entries first, then the free ones; each group sorted by (arr, tgt) */
	size_t nbound;                                   /* entries that read or write a cell which is both read and written by the list:
	                                                  * among those every load must precede every store; the others touch nothing
	                                                  * that any other entry touches and may run whenever */
};

static inline bool rc_boxes_meet(const RcOp &p, int pa, int pb, int pc, const RcOp &q, int qa, int qb, int qc)
{
	return p.a0 + pa < q.a0 + qa + q.na && q.a0 + qa < p.a0 + pa + p.na
	    && p.b0 + pb < q.b0 + qb + q.nb && q.b0 + qb < p.b0 + pb + p.nb
	    && p.c0 + pc < q.c0 + qc + q.nc && q.c0 + qc < p.c0 + pc + p.nc;
}

/* index of array cell (a,b,c): ((c*(ny+2) + b)*pitch + xo + a), as GIDX in kernels.cuh.
 * false: the list does not fit (too many arrays / chains, too deep, a copy that reads its own targets, arrays that
 * appear both as a real and as an imaginary part) -- the caller keeps the level-by-level kernels. */
static inline bool rc_resolve(const RcOp *ops, int n, int ny, int pitch, int xo, size_t max_arrays, size_t max_chains, size_t max_depth, RcPlan &out)
{
	struct Sym { int src, chain; };
	out.F.clear(); out.G.clear(); out.ent.clear();
	out.chains.assign(1, std::vector<int>());
	std::vector<std::unordered_map<int, Sym> > state;
	std::map<std::pair<int,int>, int> extend;                         /* (chain, op) -> chain */
	for (int q=0; q<n; q++) {
		const RcOp &o = ops[q];
		if (rc_boxes_meet(o, 0, 0, 0, o, o.alt0, o.alt1, o.alt2)) return false;
		size_t arr = 0;
		while (arr < out.F.size() && out.F[arr] != o.F) arr++;
		if (arr == out.F.size()) { if (arr == max_arrays) return false; out.F.push_back(o.F); out.G.push_back(o.G); state.emplace_back(); }
		if (out.G[arr] != o.G) return false;
		for (size_t other=0; other<out.F.size(); other++) {
			if (other == arr) continue;
			if (out.F[other] == o.G && o.G) return false;
			if (out.G[other] && (out.G[other] == o.F || out.G[other] == o.G)) return false;
		}
		std::unordered_map<int, Sym> &st = state[arr];
		for (int c=o.c0; c<o.c0+o.nc; c++) for (int b=o.b0; b<o.b0+o.nb; b++) for (int a=o.a0; a<o.a0+o.na; a++) {
			const long long i = ((long long)c*(ny+2) + b)*pitch + xo + a;
			const long long j = ((long long)(c+o.alt2)*(ny+2) + (b+o.alt1))*pitch + xo + (a+o.alt0);
			if (i < 0 || j < 0 || i > 0x7fffffffLL || j > 0x7fffffffLL) return false;
			Sym v = { (int)j, 0 };
			std::unordered_map<int, Sym>::const_iterator it = st.find((int)j);
			if (it != st.end()) v = it->second;
			const std::pair<int,int> key(v.chain, q);
			std::map<std::pair<int,int>, int>::const_iterator ex = extend.find(key);
			int next;
			if (ex == extend.end()) {
				if (out.chains[v.chain].size() == max_depth || out.chains.size() == max_chains) return false;
				std::vector<int> longer = out.chains[v.chain];
				longer.push_back(q);
				out.chains.push_back(longer);
				next = (int)out.chains.size() - 1;
				extend[key] = next;
			} else next = ex->second;
			v.chain = next;
			st[(int)i] = v;
		}
	}
	std::vector<RcEntry> bound, loose;
	for (size_t arr=0; arr<state.size(); arr++) {
		const std::unordered_map<int, Sym> &st = state[arr];
		std::unordered_map<int, char> read;                               /* cells whose old value some entry needs */
		for (std::unordered_map<int, Sym>::const_iterator kv = st.begin(); kv != st.end(); ++kv) read[kv->second.src] = 1;
		const size_t b0 = bound.size(), l0 = loose.size();
		for (std::unordered_map<int, Sym>::const_iterator kv = st.begin(); kv != st.end(); ++kv) {
			RcEntry en = { kv->first, kv->second.src, (int)arr, kv->second.chain };
			const bool hot = st.count(en.src) || read.count(en.tgt);        /* its source is rewritten, or its target is somebody's source */
			(hot ? bound : loose).push_back(en);
		}
		std::sort(bound.begin() + b0, bound.end(), [](const RcEntry &p, const RcEntry &q) { return p.tgt < q.tgt; });
		std::sort(loose.begin() + l0, loose.end(), [](const RcEntry &p, const RcEntry &q) { return p.tgt < q.tgt; });
	}
	out.nbound = bound.size();
	out.ent = bound;
	out.ent.insert(out.ent.end(), loose.begin(), loose.end());
	return true;
}
#endif
