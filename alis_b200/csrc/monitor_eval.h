/*
 * monitor_eval.h -- the reference's point-wise `field` functions (src/fields.c) that depend on the
 * E/H arrays only, written once for the device (kernels.cuh: eval_box, reduce_absmax, column_sums)
 * and for the CPU restatement of the same ABI calls in oracle/alis_oracle.c.
 *
 * The includer defines
 *   MON_FN                         function qualifiers (static inline / __device__ __forceinline__)
 *   MON_CTX                        context type, passed as `const MON_CTX *X`
 *   MON_DESC(X)                    const alisb_field_fn* of the request
 *   MON_LD(X,set,m,c,a,b,cc)       field value of array (set, m, comp c) at ARRAY coordinates
 *   MON_DTD(X,axis,q)              CE->dtdx/dtdy/dtdz at array coordinate q
 *   MON_DT(X), MON_COMPLEX(X)      W->dt, W->complexField
 *
 * Every expression keeps the reference's operand order and its float/double promotions (a bare
 * 0.5 is a double constant in C), so values are bit-identical to fields.c when the fields are
 * (no FMA contraction: -fmad=false / -ffp-contract=off).  Cited per function.
 */
#ifndef ALIS_B200_MONITOR_EVAL_H
#define ALIS_B200_MONITOR_EVAL_H

typedef struct { int i, j, k; float p; } mon_fold_t;

/* mirror folding of a negative index on a SYM axis; the value picks up the component's parity
 * through the world's cos factors (fields.c:130-291: `if (i<0 && W->xMinSur==SYM) ...`) */
MON_FN mon_fold_t mon_fold_cos(const alisb_field_fn *f, int i, int j, int k, float p, int sx, int sy, int sz)
{
	mon_fold_t r; r.i = i; r.j = j; r.k = k; r.p = p;
	if (i<0 && f->sym[0]) { r.i = -i; r.p *= sx > 0 ? f->cosk[0] : -f->cosk[0]; }
	if (j<0 && f->sym[1]) { r.j = -j; r.p *= sy > 0 ? f->cosk[1] : -f->cosk[1]; }
	if (k<0 && f->sym[2]) { r.k = -k; r.p *= sz > 0 ? f->cosk[2] : -f->cosk[2]; }
	return r;
}
/* products of two components: parity only (fields.c:381-493) */
MON_FN mon_fold_t mon_fold_sign(const alisb_field_fn *f, int i, int j, int k, float p, int sx, int sy, int sz)
{
	mon_fold_t r; r.i = i; r.j = j; r.k = k; r.p = p;
	if (i<0 && f->sym[0]) { r.i = -i; r.p *= sx; }
	if (j<0 && f->sym[1]) { r.j = -j; r.p *= sy; }
	if (k<0 && f->sym[2]) { r.k = -k; r.p *= sz; }
	return r;
}
#define MON_SGN(c,axis,own) ((c)==(axis) ? (own) : -(own))      /* E: own=+1, H: own=-1 */
#define MON_AT(X,set,m,c,i,j,k) MON_LD(X, set, m, c, (i)+MON_DESC(X)->org[0], (j)+MON_DESC(X)->org[1], (k)+MON_DESC(X)->org[2])
#define MON_SQ(x) ((x)*(x))

/* the two samples of an E component along its own axis, the four samples of an H component around
 * the node in the order (0,0), (+a1), (+a2), (+a1+a2), a1 = c+1, a2 = c+2 cyclic */
MON_FN float mon_e_pair(const MON_CTX *X, int m, int c, int i, int j, int k)
{
	return MON_AT(X,0,m,c,i,j,k) + MON_AT(X,0,m,c,i+(c==0),j+(c==1),k+(c==2));
}
MON_FN float mon_e_pair_sq(const MON_CTX *X, int c, int i, int j, int k)
{
	return MON_SQ(MON_AT(X,0,0,c,i,j,k)) + MON_SQ(MON_AT(X,0,0,c,i+(c==0),j+(c==1),k+(c==2)));
}
MON_FN float mon_h_quad(const MON_CTX *X, int m, int c, int i, int j, int k)
{
	const int a1 = (c+1)%3, a2 = (c+2)%3;
	const int x1 = a1==0, y1 = a1==1, z1 = a1==2, x2 = a2==0, y2 = a2==1, z2 = a2==2;
	return MON_AT(X,1,m,c,i,j,k) + MON_AT(X,1,m,c,i+x1,j+y1,k+z1) + MON_AT(X,1,m,c,i+x2,j+y2,k+z2) + MON_AT(X,1,m,c,i+x1+x2,j+y1+y2,k+z1+z2);
}
MON_FN float mon_h_quad_sq(const MON_CTX *X, int c, int i, int j, int k)
{
	const int a1 = (c+1)%3, a2 = (c+2)%3;
	const int x1 = a1==0, y1 = a1==1, z1 = a1==2, x2 = a2==0, y2 = a2==1, z2 = a2==2;
	return MON_SQ(MON_AT(X,1,0,c,i,j,k)) + MON_SQ(MON_AT(X,1,0,c,i+x1,j+y1,k+z1)) + MON_SQ(MON_AT(X,1,0,c,i+x2,j+y2,k+z2)) + MON_SQ(MON_AT(X,1,0,c,i+x1+x2,j+y1+y2,k+z1+z2));
}

/* E_a H_b at the node (fields.c:381-439): each of the two E_a samples times the sum of the two H_b
 * samples that flank it along the third axis; odd along that axis */
MON_FN float mon_e_cross_h(const MON_CTX *X, int a, int b, int i, int j, int k)
{
	const int t = 3-a-b;
	const mon_fold_t r = mon_fold_sign(MON_DESC(X), i, j, k, 0.25, t==0 ? -1 : 1, t==1 ? -1 : 1, t==2 ? -1 : 1);
	const int ax = a==0, ay = a==1, az = a==2, tx = t==0, ty = t==1, tz = t==2;
	i = r.i; j = r.j; k = r.k;
	return r.p * (MON_AT(X,0,0,a,i,j,k) * (MON_AT(X,1,0,b,i,j,k) + MON_AT(X,1,0,b,i+tx,j+ty,k+tz))
	            + MON_AT(X,0,0,a,i+ax,j+ay,k+az) * (MON_AT(X,1,0,b,i+ax,j+ay,k+az) + MON_AT(X,1,0,b,i+ax+tx,j+ay+ty,k+az+tz)));
}

/* the request's field function at world INDEX (i,j,k) */
MON_FN float mon_eval(const MON_CTX *X, int i, int j, int k)
{
	const alisb_field_fn *f = MON_DESC(X);
	const int c = f->comp;
	switch (f->fn) {
	case ALISB_FN_RAW: {                                   /* rawEx..rawHz, fields.c:130-176 */
		const int own = f->set ? -1 : 1;
		const mon_fold_t r = mon_fold_cos(f, i, j, k, 1, MON_SGN(c,0,own), MON_SGN(c,1,own), MON_SGN(c,2,own));
		return r.p * MON_AT(X, f->set, 0, c, r.i, r.j, r.k);
	}
	case ALISB_FN_NODE: {                                  /* Ex..Hz, iEx..iHz, fields.c:179-291 */
		if (f->m && !MON_COMPLEX(X)) return 0;
		if (!f->set) {
			const mon_fold_t r = mon_fold_cos(f, i, j, k, 0.5, MON_SGN(c,0,1), MON_SGN(c,1,1), MON_SGN(c,2,1));
			return r.p * mon_e_pair(X, f->m, c, r.i, r.j, r.k);
		} else {
			const mon_fold_t r = mon_fold_cos(f, i, j, k, 0.25, MON_SGN(c,0,-1), MON_SGN(c,1,-1), MON_SGN(c,2,-1));
			return r.p * mon_h_quad(X, f->m, c, r.i, r.j, r.k);
		}
	}
	case ALISB_FN_EE: {                                    /* fields.c:453-461 */
		const mon_fold_t r = mon_fold_sign(f, i, j, k, 1, 1, 1, 1);
		float s = 0;
		for (int q=0; q<3; q++) s += mon_e_pair_sq(X, q, r.i, r.j, r.k);
		return 0.5 * s;
	}
	case ALISB_FN_HH: case ALISB_FN_UH: {                  /* fields.c:476-487 */
		const mon_fold_t r = mon_fold_sign(f, i, j, k, 1, 1, 1, 1);
		float s = 0;
		for (int q=0; q<3; q++) s += mon_h_quad_sq(X, q, r.i, r.j, r.k);
		const float hh = 0.25 * s;
		return f->fn == ALISB_FN_HH ? hh : (float)(0.5*hh);
	}
	case ALISB_FN_EXH:                                     /* ExHy ... EzHy */
		return mon_e_cross_h(X, c, f->comp2, i, j, k);
	case ALISB_FN_S: {                                     /* Sx, Sy, Sz, fields.c:483-485 */
		const int a = (c+1)%3, b = (c+2)%3;
		return mon_e_cross_h(X, a, b, i, j, k) - mon_e_cross_h(X, b, a, i, j, k);
	}
	case ALISB_FN_TDIAG: {                                 /* Txx, Tyy, Tzz, fields.c:564-611 */
		const mon_fold_t r = mon_fold_sign(f, i, j, k, 1, 1, 1, 1);
		float s = 0;
		for (int q=0; q<3; q++) { const float h = mon_h_quad_sq(X, q, r.i, r.j, r.k); if (q==c) s += h; else s -= h; }
		s *= 0.5;
		for (int q=0; q<3; q++) { const float e = mon_e_pair_sq(X, q, r.i, r.j, r.k); if (q==c) s += e; else s -= e; }
		return 0.25 * s;
	}
	case ALISB_FN_TOFF: {                                  /* Txy ... Tzy, fields.c:613-681; comp, comp2 = factor order */
		const int first = c, second = f->comp2, t = 3-first-second;
		const mon_fold_t r = mon_fold_sign(f, i, j, k, 0.25, t==0 ? 1 : -1, t==1 ? 1 : -1, t==2 ? 1 : -1);
		float s = 0;
		s += mon_e_pair(X, 0, first, r.i, r.j, r.k) * mon_e_pair(X, 0, second, r.i, r.j, r.k);
		s += mon_h_quad(X, 0, first, r.i, r.j, r.k) * mon_h_quad(X, 0, second, r.i, r.j, r.k) * 0.25;
		return r.p * s;
	}
	case ALISB_FN_DIVE: {                                  /* fields.c:552-562 */
		const mon_fold_t r = mon_fold_cos(f, i, j, k, 1, -1, -1, -1);
		i = r.i; j = r.j; k = r.k;
		return ((MON_AT(X,0,0,0,i+1,j,k) - MON_AT(X,0,0,0,i,j,k)) * MON_DTD(X, 0, i + f->org[0]) +
		        (MON_AT(X,0,0,1,i,j+1,k) - MON_AT(X,0,0,1,i,j,k)) * MON_DTD(X, 1, j + f->org[1]) +
		        (MON_AT(X,0,0,2,i,j,k+1) - MON_AT(X,0,0,2,i,j,k)) * MON_DTD(X, 2, k + f->org[2])) * r.p / MON_DT(X);
	}
	}
	return 0;
}

/* one z column of worldSum / sliceSum (output.c:73-80 worldSumZ, slices.c:396-403 sliceSumZ):
 * the interior points first, then the two half-weighted ends; dz = CE->dz in array coordinates */
#define MON_COLUMN_SUM(X, dz, i, j, k0, k1, result) do { \
	float Sum_ = 0; \
	for (int k_ = (k0)+1; k_ < (k1); k_++) Sum_ += (dz)[k_ + MON_DESC(X)->org[2]] * mon_eval(X, i, j, k_); \
	Sum_ += 0.5 * (dz)[(k0) + MON_DESC(X)->org[2]] * mon_eval(X, i, j, k0); \
	Sum_ += 0.5 * (dz)[(k1) + MON_DESC(X)->org[2]] * mon_eval(X, i, j, k1); \
	(result) = Sum_; \
} while (0)

/* one z column of objectAbsorption (output.c:23-47 rawJ, :375-404 partialAbsorption): the three
 * sums xs, ys, zs of E.J over the planes c0..c1 of ADE object n at array coordinates (a,b), with
 * the half-weight corrections at the world's first / last plane.  The includer also defines
 *   MON_INV(X,c,a,b,cc)  inverse permittivity     MON_FILL(X,c,a,b,cc)  fill factor of the object
 *   MON_J(X,p,c,a,b,cc)  polarisation current p   MON_NPOLES(X), MON_LOSS(X,c)  NK and ax/ay/az
 *   MON_DZE(X,cc), MON_DZH(X,cc)  CE->dz, CH->dz at array coordinate cc                          */
#ifdef MON_INV
MON_FN float mon_object_current(const MON_CTX *X, int c, int a, int b, int cc)
{
	const float loss = MON_LOSS(X, c);
	float f = 0;
	if (loss) f += 2 * MON_FILL(X,c,a,b,cc) * MON_LD(X,0,0,c,a,b,cc) / (1 + 1 / loss);
	for (int p=0; p<MON_NPOLES(X); p++) f += MON_J(X,p,c,a,b,cc);
	return f / MON_INV(X,c,a,b,cc);
}
#define MON_ABSORPTION_COLUMN(X, a, b, c0, c1, cMIN, cMAX, xs, ys, zs) do { \
	float xs_ = 0, ys_ = 0, zs_ = 0; \
	for (int c_ = (c0); c_ <= (c1); c_++) { \
		if (MON_INV(X,0,a,b,c_)) xs_ += MON_DZE(X,c_) * MON_LD(X,0,0,0,a,b,c_) * mon_object_current(X, 0, a, b, c_); \
		if (MON_INV(X,1,a,b,c_)) ys_ += MON_DZE(X,c_) * MON_LD(X,0,0,1,a,b,c_) * mon_object_current(X, 1, a, b, c_); \
		if (MON_INV(X,2,a,b,c_)) zs_ += MON_DZH(X,c_) * MON_LD(X,0,0,2,a,b,c_) * mon_object_current(X, 2, a, b, c_); \
	} \
	if ((c0) == (cMIN)) { \
		const int c_ = (cMIN); \
		if (MON_INV(X,0,a,b,c_)) xs_ -= MON_DZE(X,c_) * MON_LD(X,0,0,0,a,b,c_) * mon_object_current(X, 0, a, b, c_) * 0.5; \
		if (MON_INV(X,1,a,b,c_)) ys_ -= MON_DZE(X,c_) * MON_LD(X,0,0,1,a,b,c_) * mon_object_current(X, 1, a, b, c_) * 0.5; \
		if (MON_INV(X,2,a,b,c_)) zs_ -= MON_DZH(X,c_) * MON_LD(X,0,0,2,a,b,c_) * mon_object_current(X, 2, a, b, c_); \
	} \
	if ((c1) == (cMAX)) { \
		const int c_ = (cMAX); \
		if (MON_INV(X,0,a,b,c_)) xs_ -= MON_DZE(X,c_) * MON_LD(X,0,0,0,a,b,c_) * mon_object_current(X, 0, a, b, c_) * 0.5; \
		if (MON_INV(X,1,a,b,c_)) ys_ -= MON_DZE(X,c_) * MON_LD(X,0,0,1,a,b,c_) * mon_object_current(X, 1, a, b, c_) * 0.5; \
	} \
	(xs) = xs_; (ys) = ys_; (zs) = zs_; \
} while (0)
#endif

#endif
