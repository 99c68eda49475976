/*
 * shapes.c -- inside/outside predicates `int S(object O, float x, float y, float z)`.
 *
 * Users compose these through object.p pointers and may add their own, so they are ordinary
 * host functions; the rasteriser (object.c) calls them with sub-pixel sampling.  Semantics,
 * including the strict/non-strict inequalities, follow reference src/shapes.c:17-235 so that the
 * fill factors -- and therefore the inverse-permittivity arrays uploaded to the GPU -- are
 * bit-identical.  All arithmetic is in float exactly as written there.
 *
 * Extent convention (reference shapes.c:9-13): O.f = {{xMin,xMax},{yMin,yMax},{zMin,zMax}} when the
 * z pair is non-degenerate, otherwise the "centre + half-width" form {{cx,cy,cz},{rx,ry,rz}}.
 * In the second form the reference's conditional expressions bind as `(c-r) ? r : rx`; that
 * quirk is reproduced (it only matters for the y and z extents of centre-form objects).
 */
#include "alis_c.h"

typedef struct { float x0, x1, y0, y1, z0, z1; } extent;

static extent extent_of(object O)
{
	extent e;
	if (O.f[2][1]-O.f[2][0]) {
		e.x0 = O.f[0][0], e.x1 = O.f[0][1];
		e.y0 = O.f[1][0], e.y1 = O.f[1][1];
		e.z0 = O.f[2][0], e.z1 = O.f[2][1];
	}
	else {
		e.x0 = O.f[0][0]-O.f[1][0];
		e.y0 = O.f[0][1]-O.f[1][1] ? O.f[1][1] : O.f[1][0];
		e.z0 = O.f[0][2]-O.f[1][2] ? O.f[1][2] : O.f[1][0];
		e.x1 = O.f[0][0]+O.f[1][0];
		e.y1 = O.f[0][1]+O.f[1][1] ? O.f[1][1] : O.f[1][0];
		e.z1 = O.f[0][2]+O.f[1][2] ? O.f[1][2] : O.f[1][0];
	}
	return e;
}

/* coordinate mapped to [-1,1] across [lo,hi] */
static inline float unit(float v, float lo, float hi) { return (2*v-lo-hi)/(hi-lo); }
static inline int within(float v, float lo, float hi) { return lo <= v && v <= hi; }
static inline int strictly(float v, float lo, float hi) { return lo < v && v < hi; }

/* ---- primitives -------------------------------------------------------------------------------- */

int Box(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(x, e.x0, e.x1) && within(y, e.y0, e.y1) && within(z, e.z0, e.z1);
}

int Ball(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return sq(unit(x, e.x0, e.x1)) + sq(unit(y, e.y0, e.y1)) + sq(unit(z, e.z0, e.z1)) < 1;
}

int Diamond(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return fabsf(unit(x, e.x0, e.x1)) + fabsf(unit(y, e.y0, e.y1)) + fabsf(unit(z, e.z0, e.z1)) <= 1;
}

/* double cones: |axis coordinate| + radial distance < 1 (the square root is taken in double) */
int BiconeX(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return fabsf(unit(x, e.x0, e.x1)) + sqrt(sq(unit(y, e.y0, e.y1)) + sq(unit(z, e.z0, e.z1))) < 1;
}

int BiconeY(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return fabsf(unit(y, e.y0, e.y1)) + sqrt(sq(unit(z, e.z0, e.z1)) + sq(unit(x, e.x0, e.x1))) < 1;
}

int BiconeZ(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return fabsf(unit(z, e.z0, e.z1)) + sqrt(sq(unit(x, e.x0, e.x1)) + sq(unit(y, e.y0, e.y1))) < 1;
}

/* elliptic cylinders along one axis */
int RodX(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(x, e.x0, e.x1) && sq(unit(y, e.y0, e.y1)) + sq(unit(z, e.z0, e.z1)) < 1;
}

int RodY(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(y, e.y0, e.y1) && sq(unit(x, e.x0, e.x1)) + sq(unit(z, e.z0, e.z1)) < 1;
}

int RodZ(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(z, e.z0, e.z1) && sq(unit(x, e.x0, e.x1)) + sq(unit(y, e.y0, e.y1)) < 1;
}

/* diamond (rhombic) prisms */
int DRodX(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(x, e.x0, e.x1) && fabsf(unit(y, e.y0, e.y1)) + fabsf(unit(z, e.z0, e.z1)) <= 1;
}

int DRodY(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(y, e.y0, e.y1) && fabsf(unit(z, e.z0, e.z1)) + fabsf(unit(x, e.x0, e.x1)) <= 1;
}

int DRodZ(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(z, e.z0, e.z1) && fabsf(unit(x, e.x0, e.x1)) + fabsf(unit(y, e.y0, e.y1)) <= 1;
}

/* hexagonal prisms HRod<axis><pointing direction>: |flat| + 2|pointed| < 2 inside the flat pair */
int HRodXY(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(x, e.x0, e.x1) && strictly(z, e.z0, e.z1) && fabsf(unit(z, e.z0, e.z1)) + fabsf(unit(y, e.y0, e.y1))*2 < 2;
}

int HRodYZ(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(y, e.y0, e.y1) && strictly(x, e.x0, e.x1) && fabsf(unit(x, e.x0, e.x1)) + fabsf(unit(z, e.z0, e.z1))*2 < 2;
}

int HRodZX(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(z, e.z0, e.z1) && strictly(y, e.y0, e.y1) && fabsf(unit(y, e.y0, e.y1)) + fabsf(unit(x, e.x0, e.x1))*2 < 2;
}

int HRodXZ(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(x, e.x0, e.x1) && strictly(y, e.y0, e.y1) && fabsf(unit(y, e.y0, e.y1)) + fabsf(unit(z, e.z0, e.z1))*2 < 2;
}

int HRodYX(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(y, e.y0, e.y1) && strictly(z, e.z0, e.z1) && fabsf(unit(z, e.z0, e.z1)) + fabsf(unit(x, e.x0, e.x1))*2 < 2;
}

int HRodZY(object O, float x, float y, float z)
{
	extent e = extent_of(O);
	return within(z, e.z0, e.z1) && strictly(x, e.x0, e.x1) && fabsf(unit(x, e.x0, e.x1)) + fabsf(unit(y, e.y0, e.y1))*2 < 2;
}

/* ---- transformations and set operations on child objects (O.p) -------------------------------- */

int Translation(object O, float x, float y, float z)
{
	object *child = O.p;
	return child->S(*child, x-O.f[0][0], y-O.f[0][1], z-O.f[0][2]);
}

/* rotation by Euler angles f[0] (degrees, about x, y, z) around the pivot f[1] */
int Rotation(object O, float x, float y, float z)
{
	object *child = O.p;
	x -= O.f[1][0], y -= O.f[1][1], z -= O.f[1][2];
	float cx = cosf(O.f[0][0]*PI/180.), sx = sinf(O.f[0][0]*PI/180.);
	float cy = cosf(O.f[0][1]*PI/180.), sy = sinf(O.f[0][1]*PI/180.);
	float cz = cosf(O.f[0][2]*PI/180.), sz = sinf(O.f[0][2]*PI/180.);
	return child->S(*child,
		cy*(cz*x+sz*y)-sy*z,
		cx*(cz*y-sz*x)+sx*(cy*z+sy*(cz*x+sz*y)),
		cx*(cy*z+sy*(cz*x+sz*y))-sx*(cz*y-sz*x));
}

int Combination(object O, float x, float y, float z)          /* union of f[0][0] children */
{
	object *child = O.p;
	const int count = O.f[0][0];
	for (int q=0; q<count; q++) if (child[q].S(child[q], x, y, z)) return 1;
	return 0;
}

int Intersection(object O, float x, float y, float z)
{
	object *child = O.p;
	const int count = O.f[0][0];
	for (int q=0; q<count; q++) if (!child[q].S(child[q], x, y, z)) return 0;
	return 1;
}

int Difference(object O, float x, float y, float z)           /* first child minus the others */
{
	object *child = O.p;
	const int count = O.f[0][0];
	if (!child[0].S(child[0], x, y, z)) return 0;
	for (int q=1; q<count; q++) if (child[q].S(child[q], x, y, z)) return 0;
	return 1;
}

/* fold one coordinate into the unit cell of a lattice with period `period` and `count` copies
 * (count 0 = infinite); lo/hi are the child's extent along that axis */
static float fold(float v, float period, float count, float lo, float hi)
{
	if (!period) return v;
	if (count) v += 0.5 * (count-1) * period;
	float cell = floor((v + 0.5 * (period - lo - hi)) / period);
	v -= period * (count > 0 ? MAX(0, MIN(cell, count-1)) : cell);
	return v;
}

int Lattice(object O, float x, float y, float z)
{
	object *child = O.p;
	x = fold(x, O.f[0][0], O.f[0][1], child->f[0][0], child->f[0][1]);
	y = fold(y, O.f[1][0], O.f[1][1], child->f[1][0], child->f[1][1]);
	z = fold(z, O.f[2][0], O.f[2][1], child->f[2][0], child->f[2][1]);
	return child->S(*child, x, y, z) ? 1 : 0;
}

/* two interleaved lattices (e.g. the triangular hole lattice of a photonic-crystal slab): the
 * second one has one copy fewer per axis (two fewer when infinite), which shifts it by half a period */
int DLattice(object O, float x, float y, float z)
{
	if (Lattice(O, x, y, z)) return 1;
	if (O.f[0][0]) O.f[0][1] -= O.f[0][1] ? 1 : 2;
	if (O.f[1][0]) O.f[1][1] -= O.f[1][1] ? 1 : 2;
	if (O.f[2][0]) O.f[2][1] -= O.f[2][1] ? 1 : 2;
	return Lattice(O, x, y, z) ? 1 : 0;
}

/* f[0][0] copies rotated about an axis through (f[1][0], f[1][1]) */
int CLatticeX(object O, float x, float y, float z)
{
	object *child = O.p;
	const float y0 = O.f[1][0], z0 = O.f[1][1];
	const int count = O.f[0][0];
	for (int q=0; q<count; q++) {
		float theta = 2 * PI * q / count;
		if (child->S(*child, x, (y-y0)*cosf(theta) + (z-z0)*sinf(theta) + y0, (z-z0)*cosf(theta) - (y-y0)*sinf(theta) + z0)) return 1;
	}
	return 0;
}

int CLatticeY(object O, float x, float y, float z)
{
	object *child = O.p;
	const float x0 = O.f[1][0], z0 = O.f[1][1];
	const int count = O.f[0][0];
	for (int q=0; q<count; q++) {
		float theta = 2 * PI * q / count;
		if (child->S(*child, (x-x0)*cosf(theta) + (z-z0)*sinf(theta) + x0, y, (z-z0)*cosf(theta) - (x-x0)*sinf(theta) + z0)) return 1;
	}
	return 0;
}

int CLatticeZ(object O, float x, float y, float z)
{
	object *child = O.p;
	const float x0 = O.f[1][0], y0 = O.f[1][1];
	const int count = O.f[0][0];
	for (int q=0; q<count; q++) {
		float theta = 2 * PI * q / count;
		if (child->S(*child, (x-x0)*cosf(theta) + (y-y0)*sinf(theta) + x0, (y-y0)*cosf(theta) - (x-x0)*sinf(theta) + y0, z)) return 1;
	}
	return 0;
}

/* 2-D height map (float** in O.p, f[0][2] x f[1][2] samples over the x,y extent): inside where
 * the sample lies in [f[2][0], f[2][1]] */
int ImportContour(object O, float x, float y, float z)
{
	(void)z;
	float **map = O.p;
	if (strictly(x, O.f[0][0], O.f[0][1]) && strictly(y, O.f[1][0], O.f[1][1])) {
		int i = O.f[0][2] * (x - O.f[0][0]) / (O.f[0][1] - O.f[0][0]);
		int j = O.f[1][2] * (y - O.f[1][0]) / (O.f[1][1] - O.f[1][0]);
		if (O.f[2][0] <= map[i][j] && map[i][j] <= O.f[2][1]) return 1;
	}
	return 0;
}
