/*
 * geometry_eval.h -- the reference's built-in shape predicates (src/shapes.c:17-221) and sub-pixel
 * fill factor (src/object.c:70-121) over a POD shape tree, written once for the device
 * (kernels.cuh: geometry_bounds, geometry_fill) and for the CPU restatement in oracle/alis_oracle.c.
 *
 * The host flattens an `object` (function pointer + float f[3][3] + child pointer) into alisb_shape
 * nodes; everything that needs libm (the cos/sin of Rotation and CLattice angles, the grid's
 * index -> coordinate map incl. graded grids) is evaluated on the host with the caller's libm and
 * passed as constants / tables, so the predicates use only +,-,*,/, comparisons, floor and sqrt
 * (IEEE on both sides): inside/outside decisions, and therefore fill factors, are bit-identical.
 * Float/double promotions are the reference's (a bare 0.5 is a double).
 *
 * The includer defines GEO_FN (function qualifiers).
 */
#ifndef ALIS_B200_GEOMETRY_EVAL_H
#define ALIS_B200_GEOMETRY_EVAL_H

typedef struct { float x0, x1, y0, y1, z0, z1; } geo_extent;

/* shapes.c:9-13 (MINMAX): {{xMin,xMax},{yMin,yMax},{zMin,zMax}}, or centre + half-width when the z
 * pair is degenerate -- with the reference's operator binding `(c-r) ? r : rx` for y and z */
GEO_FN geo_extent geo_extent_of(const alisb_shape *O)
{
	geo_extent e;
	if (O->f[2][1]-O->f[2][0]) {
		e.x0 = O->f[0][0]; e.x1 = O->f[0][1];
		e.y0 = O->f[1][0]; e.y1 = O->f[1][1];
		e.z0 = O->f[2][0]; e.z1 = O->f[2][1];
	}
	else {
		e.x0 = O->f[0][0]-O->f[1][0];
		e.y0 = O->f[0][1]-O->f[1][1] ? O->f[1][1] : O->f[1][0];
		e.z0 = O->f[0][2]-O->f[1][2] ? O->f[1][2] : O->f[1][0];
		e.x1 = O->f[0][0]+O->f[1][0];
		e.y1 = O->f[0][1]+O->f[1][1] ? O->f[1][1] : O->f[1][0];
		e.z1 = O->f[0][2]+O->f[1][2] ? O->f[1][2] : O->f[1][0];
	}
	return e;
}

GEO_FN float geo_unit(float v, float lo, float hi) { return (2*v-lo-hi)/(hi-lo); }
GEO_FN int geo_within(float v, float lo, float hi) { return lo <= v && v <= hi; }
GEO_FN int geo_strictly(float v, float lo, float hi) { return lo < v && v < hi; }
#define GEO_SQ(x) ((x)*(x))

/* shapes.c:167-189 (Lattice): fold one coordinate into the unit cell */
GEO_FN float geo_fold(float v, float period, float count, float lo, float hi)
{
	if (!period) return v;
	if (count) v += 0.5 * (count-1) * period;
	float cell = floor((v + 0.5 * (period - lo - hi)) / period);
	float lim = cell;
	if (count > 0) { lim = cell < count-1 ? cell : count-1; lim = 0 > lim ? 0 : lim; }      /* MAX(0, MIN(cell, count-1)) */
	v -= period * lim;
	return v;
}

GEO_FN int geo_inside(const alisb_shape *nodes, const float *aux, int n, float x, float y, float z);

GEO_FN int geo_lattice(const alisb_shape *nodes, const float *aux, const alisb_shape *O, float x, float y, float z)
{
	const alisb_shape *child = &nodes[O->child];
	x = geo_fold(x, O->f[0][0], O->f[0][1], child->f[0][0], child->f[0][1]);
	y = geo_fold(y, O->f[1][0], O->f[1][1], child->f[1][0], child->f[1][1]);
	z = geo_fold(z, O->f[2][0], O->f[2][1], child->f[2][0], child->f[2][1]);
	return geo_inside(nodes, aux, O->child, x, y, z) ? 1 : 0;
}

/* the predicate of node n at (x,y,z) */
GEO_FN int geo_inside(const alisb_shape *nodes, const float *aux, int n, float x, float y, float z)
{
	const alisb_shape *O = &nodes[n];
	if (O->type <= ALISB_SHAPE_HRODZY) {                    /* primitives, shapes.c:17-121 */
		const geo_extent e = geo_extent_of(O);
		const float ux = geo_unit(x, e.x0, e.x1), uy = geo_unit(y, e.y0, e.y1), uz = geo_unit(z, e.z0, e.z1);
		switch (O->type) {
		case ALISB_SHAPE_BOX:     return geo_within(x, e.x0, e.x1) && geo_within(y, e.y0, e.y1) && geo_within(z, e.z0, e.z1);
		case ALISB_SHAPE_BALL:    return GEO_SQ(ux) + GEO_SQ(uy) + GEO_SQ(uz) < 1;
		case ALISB_SHAPE_DIAMOND: return fabsf(ux) + fabsf(uy) + fabsf(uz) <= 1;
		case ALISB_SHAPE_BICONEX: return fabsf(ux) + sqrt((double)(GEO_SQ(uy) + GEO_SQ(uz))) < 1;
		case ALISB_SHAPE_BICONEY: return fabsf(uy) + sqrt((double)(GEO_SQ(uz) + GEO_SQ(ux))) < 1;
		case ALISB_SHAPE_BICONEZ: return fabsf(uz) + sqrt((double)(GEO_SQ(ux) + GEO_SQ(uy))) < 1;
		case ALISB_SHAPE_RODX:    return geo_within(x, e.x0, e.x1) && GEO_SQ(uy) + GEO_SQ(uz) < 1;
		case ALISB_SHAPE_RODY:    return geo_within(y, e.y0, e.y1) && GEO_SQ(ux) + GEO_SQ(uz) < 1;
		case ALISB_SHAPE_RODZ:    return geo_within(z, e.z0, e.z1) && GEO_SQ(ux) + GEO_SQ(uy) < 1;
		case ALISB_SHAPE_DRODX:   return geo_within(x, e.x0, e.x1) && fabsf(uy) + fabsf(uz) <= 1;
		case ALISB_SHAPE_DRODY:   return geo_within(y, e.y0, e.y1) && fabsf(uz) + fabsf(ux) <= 1;
		case ALISB_SHAPE_DRODZ:   return geo_within(z, e.z0, e.z1) && fabsf(ux) + fabsf(uy) <= 1;
		case ALISB_SHAPE_HRODXY:  return geo_within(x, e.x0, e.x1) && geo_strictly(z, e.z0, e.z1) && fabsf(uz) + fabsf(uy)*2 < 2;
		case ALISB_SHAPE_HRODYZ:  return geo_within(y, e.y0, e.y1) && geo_strictly(x, e.x0, e.x1) && fabsf(ux) + fabsf(uz)*2 < 2;
		case ALISB_SHAPE_HRODZX:  return geo_within(z, e.z0, e.z1) && geo_strictly(y, e.y0, e.y1) && fabsf(uy) + fabsf(ux)*2 < 2;
		case ALISB_SHAPE_HRODXZ:  return geo_within(x, e.x0, e.x1) && geo_strictly(y, e.y0, e.y1) && fabsf(uy) + fabsf(uz)*2 < 2;
		case ALISB_SHAPE_HRODYX:  return geo_within(y, e.y0, e.y1) && geo_strictly(z, e.z0, e.z1) && fabsf(uz) + fabsf(ux)*2 < 2;
		case ALISB_SHAPE_HRODZY:  return geo_within(z, e.z0, e.z1) && geo_strictly(x, e.x0, e.x1) && fabsf(ux) + fabsf(uy)*2 < 2;
		}
		return 0;
	}
	switch (O->type) {                                      /* compound objects, shapes.c:126-221 */
	case ALISB_SHAPE_TRANSLATION:
		return geo_inside(nodes, aux, O->child, x-O->f[0][0], y-O->f[0][1], z-O->f[0][2]);
	case ALISB_SHAPE_ROTATION: {                            /* aux: cx, sx, cy, sy, cz, sz from the host's libm */
		const float *t = aux + O->aux;
		const float cx = t[0], sx = t[1], cy = t[2], sy = t[3], cz = t[4], sz = t[5];
		x -= O->f[1][0]; y -= O->f[1][1]; z -= O->f[1][2];
		return geo_inside(nodes, aux, O->child,
			cy*(cz*x+sz*y)-sy*z,
			cx*(cz*y-sz*x)+sx*(cy*z+sy*(cz*x+sz*y)),
			cx*(cy*z+sy*(cz*x+sz*y))-sx*(cz*y-sz*x));
	}
	case ALISB_SHAPE_COMBINATION:
		for (int q=0; q<O->nchild; q++) if (geo_inside(nodes, aux, O->child+q, x, y, z)) return 1;
		return 0;
	case ALISB_SHAPE_INTERSECTION:
		for (int q=0; q<O->nchild; q++) if (!geo_inside(nodes, aux, O->child+q, x, y, z)) return 0;
		return 1;
	case ALISB_SHAPE_DIFFERENCE:
		if (!geo_inside(nodes, aux, O->child, x, y, z)) return 0;
		for (int q=1; q<O->nchild; q++) if (geo_inside(nodes, aux, O->child+q, x, y, z)) return 0;
		return 1;
	case ALISB_SHAPE_LATTICE:
		return geo_lattice(nodes, aux, O, x, y, z);
	case ALISB_SHAPE_DLATTICE: {                            /* a second lattice with one copy fewer per axis */
		if (geo_lattice(nodes, aux, O, x, y, z)) return 1;
		alisb_shape D = *O;
		if (D.f[0][0]) D.f[0][1] -= D.f[0][1] ? 1 : 2;
		if (D.f[1][0]) D.f[1][1] -= D.f[1][1] ? 1 : 2;
		if (D.f[2][0]) D.f[2][1] -= D.f[2][1] ? 1 : 2;
		return geo_lattice(nodes, aux, &D, x, y, z) ? 1 : 0;
	}
	case ALISB_SHAPE_CLATTICEX: case ALISB_SHAPE_CLATTICEY: case ALISB_SHAPE_CLATTICEZ: {
		/* aux: (cosf(theta_q), sinf(theta_q)) for q = 0..count-1 from the host's libm; count = f[0][0] */
		const float u0 = O->f[1][0], v0 = O->f[1][1];
		const int count = (int)O->f[0][0];
		for (int q=0; q<count; q++) {
			const float c = aux[O->aux + 2*q], s = aux[O->aux + 2*q + 1];
			int in;
			if (O->type == ALISB_SHAPE_CLATTICEX)      in = geo_inside(nodes, aux, O->child, x, (y-u0)*c + (z-v0)*s + u0, (z-v0)*c - (y-u0)*s + v0);
			else if (O->type == ALISB_SHAPE_CLATTICEY) in = geo_inside(nodes, aux, O->child, (x-u0)*c + (z-v0)*s + u0, y, (z-v0)*c - (x-u0)*s + v0);
			else                                       in = geo_inside(nodes, aux, O->child, (x-u0)*c + (y-v0)*s + u0, (y-v0)*c - (x-u0)*s + v0, z);
			if (in) return 1;
		}
		return 0;
	}
	}
	return 0;
}

/* coordinate of (index idx, half-cell flag, sample slot) along one axis: tables prepared by the host
 * with the world's itox/jtoy/ktoz (uniform, graded or user grid function) */
GEO_FN float geo_coord(const alisb_geometry *g, int axis, int idx, int half, int slot)
{
	return g->coord[axis][((size_t)(idx - g->lo[axis])*2 + half)*(g->subn + 3) + slot];
}

/* object.c:70-75 (inThisObject): inside object o and inside none of the objects listed before it */
GEO_FN float geo_owned(const alisb_geometry *g, int o, float x, float y, float z)
{
	if (!geo_inside(g->nodes, g->aux, g->roots[o], x, y, z)) return 0;
	while (o--) if (geo_inside(g->nodes, g->aux, g->roots[o], x, y, z)) return 0;
	return 1;
}

/* object.c:78-121 (fillFactor) for subN >= 1 at the component position (idx[a] - half[a]/2):
 * slots: 0 centre, 1/2 the outermost sub-samples -/+, 3.. the subN regular sub-samples */
GEO_FN float geo_fill(const alisb_geometry *g, int o, const int idx[3], const int half[3])
{
	const float centre = geo_owned(g, o, geo_coord(g,0,idx[0],half[0],0), geo_coord(g,1,idx[1],half[1],0), geo_coord(g,2,idx[2],half[2],0));
	if (g->subn < 2) return centre;
	int uniform = 1;
	for (int c=0; c<8; c++) {
		const float s = geo_owned(g, o, geo_coord(g,0,idx[0],half[0],c&4 ? 2 : 1), geo_coord(g,1,idx[1],half[1],c&2 ? 2 : 1), geo_coord(g,2,idx[2],half[2],c&1 ? 2 : 1));
		if (s != centre) { uniform = 0; break; }
	}
	if (uniform) return centre;
	float inside = 0;
	for (int a=0; a<g->subn; a++) for (int b=0; b<g->subn; b++) for (int c=0; c<g->subn; c++)
		inside += geo_owned(g, o, geo_coord(g,0,idx[0],half[0],3+a), geo_coord(g,1,idx[1],half[1],3+b), geo_coord(g,2,idx[2],half[2],3+c));
	return inside / (g->subn * g->subn * g->subn);
}

#endif
