/*
 * object.c -- rasterises objects into the per-component inverse-permittivity arrays and the
 * per-dispersive-object ADE data (bounding box, fill factors, Drude / Lorentz / critical-point
 * recursion coefficients) that the GPU engine consumes (engine_bridge.c: push_materials).
 *
 * Shapes are user-extensible host callbacks, so rasterisation stays on the host and the results
 * are uploaded once.  Numerical behaviour follows reference src/object.c (cited per function):
 * the same sub-pixel sampling pattern and the same float/complex-float expressions, so that the
 * uploaded arrays are bit-identical to the reference's.
 */
#include <string.h>
#include "alis_c.h"
#include "host_internal.h"

typedef struct { int iMin, iMax, jMin, jMax, kMin, kMax; } bounds;

/* ---- ADE coefficient formulas (reference object.c:23-67) ------------------------------------------
 * pole p of matter M: w (resonance, eV), r (damping), f, g (Lorentz / critical-point weights).
 * Drude:            J' = a J + b E         a = (1 - pi r dt/eV)/(1 + pi r dt/eV),  b = (2 pi w dt/eV)^2/(1 + pi r dt/eV)
 * critical point:   complex recursion with Gamma = r + i w and sigma = i (f w - g Gamma)                */
static float complex pole_gamma(matter M, int p) { return M.P[p].r + I * M.P[p].w; }
static float complex pole_sigma(matter M, int p) { return I * (M.P[p].f * M.P[p].w - M.P[p].g * pole_gamma(M, p)); }

static float drude_a(world W, matter M, int p)
{
	return (1 - PI * M.P[p].r * W->dt / W->eV) / (1 + PI * M.P[p].r * W->dt / W->eV);
}

static float drude_b(world W, matter M, int p)
{
	return sq(2 * PI * M.P[p].w * W->dt / W->eV) / (1 + PI * M.P[p].r * W->dt / W->eV);
}

static float complex cp_a(world W, matter M, int p)
{
	return (1 - PI * pole_gamma(M, p) * W->dt / W->eV) / (1 + PI * pole_gamma(M, p) * W->dt / W->eV);
}

static float complex cp_b(world W, matter M, int p)
{
	return - pole_gamma(M, p) * pole_sigma(M, p) * sq(2 * PI * W->dt / W->eV) / (1 + PI * pole_gamma(M, p) * W->dt / W->eV);
}

/* static conductivity of component q (0,1,2): pure conductivity poles {0,s} plus the g w part of
 * critical-point poles; anisotropic matter stores its poles in x,y,z triples */
static float conductivity(world W, matter M, int q)
{
	const int tri = M.e.y || M.e.z ? 1 : 0;
	float s = 0;
	for (int p=0; p<MAXPOLES; p+=1+2*tri) {
		if (!M.P[p+q*tri].w && M.P[p+q*tri].r) s += M.P[p+q*tri].r;
		if (M.P[p+q*tri].g) s += M.P[p+q*tri].g * M.P[p+q*tri].w;
	}
	return s / W->eV;
}

/* effective eps_inf of component q including the conductivity half-step term */
static float eps_eff(world W, matter M, int q)
{
	return ((q==2 && M.e.z) ? M.e.z : (q==1 && M.e.y) ? M.e.y : M.e.x) + PI * W->dt * conductivity(W, M, q);
}

/* 2: chi(2) matter, 1: dispersive/lossy matter (needs auxiliary fields), 0: plain dielectric */
static int needs_aux(matter M)
{
	for (int a=0; a<3; a++) for (int b=0; b<6; b++) if (M.d[a][b]) return 2;
	if (M.P[0].w || M.P[0].r) return 1;
	return 0;
}

/* ---- sub-pixel rasterisation (reference object.c:70-121) ---------------------------------------------- */

/* 1 if (i,j,k) (fractional indices) is inside object o and in none of the objects listed before it */
static float owned_by(world W, object *O, int o, float i, float j, float k)
{
	if (!(O[o].S(O[o], W->itox(W, i), W->jtoy(W, j), W->ktoz(W, k)))) return 0;
	while (o--) if (O[o].S(O[o], W->itox(W, i), W->jtoy(W, j), W->ktoz(W, k))) return 0;
	return 1;
}

static float fill_factor(world W, object *O, int o, float i, float j, float k)
{
	float centre = owned_by(W, O, o, i, j, k);

	if (W->subN < 2) {
		if (centre || W->subN >= 0) return centre;
		/* subN < 0: a component counts as inside if an adjacent pair of nodes is inside */
		if (floorf(i) < ceilf(i)) {
			if (owned_by(W, O, o, i-0.5, j-0.5, k) && owned_by(W, O, o, i+0.5, j-0.5, k)) return 1;
			if (owned_by(W, O, o, i-0.5, j+0.5, k) && owned_by(W, O, o, i+0.5, j+0.5, k)) return 1;
			if (owned_by(W, O, o, i-0.5, j, k-0.5) && owned_by(W, O, o, i+0.5, j, k-0.5)) return 1;
			if (owned_by(W, O, o, i-0.5, j, k+0.5) && owned_by(W, O, o, i+0.5, j, k+0.5)) return 1;
		}
		else if (floorf(j) < ceilf(j)) {
			if (owned_by(W, O, o, i-0.5, j-0.5, k) && owned_by(W, O, o, i-0.5, j+0.5, k)) return 1;
			if (owned_by(W, O, o, i+0.5, j-0.5, k) && owned_by(W, O, o, i+0.5, j+0.5, k)) return 1;
			if (owned_by(W, O, o, i, j-0.5, k-0.5) && owned_by(W, O, o, i, j+0.5, k-0.5)) return 1;
			if (owned_by(W, O, o, i, j-0.5, k+0.5) && owned_by(W, O, o, i, j+0.5, k+0.5)) return 1;
		}
		else if (floorf(k) < ceilf(k)) {
			if (owned_by(W, O, o, i-0.5, j, k-0.5) && owned_by(W, O, o, i-0.5, j, k+0.5)) return 1;
			if (owned_by(W, O, o, i+0.5, j, k-0.5) && owned_by(W, O, o, i+0.5, j, k+0.5)) return 1;
			if (owned_by(W, O, o, i, j-0.5, k-0.5) && owned_by(W, O, o, i, j-0.5, k+0.5)) return 1;
			if (owned_by(W, O, o, i, j+0.5, k-0.5) && owned_by(W, O, o, i, j+0.5, k+0.5)) return 1;
		}
		return 0;
	}

	/* quick exit: centre and the eight outermost sub-samples agree -> cell is uniform */
	float edge = (W->subN-1.0) / (W->subN*2.0);
	int uniform = 1;
	for (int c=0; c<8 && uniform; c++) {
		float s = owned_by(W, O, o, i + (c&4 ? edge : -edge), j + (c&2 ? edge : -edge), k + (c&1 ? edge : -edge));
		if (s != centre) uniform = 0;
	}
	if (uniform) return centre;

	/* subN^3 regular sub-samples */
	float inside = 0;
	for (int di=1-W->subN; di<W->subN; di+=2) for (int dj=1-W->subN; dj<W->subN; dj+=2) for (int dk=1-W->subN; dk<W->subN; dk+=2)
		inside += owned_by(W, O, o, i+0.5*di/W->subN, j+0.5*dj/W->subN, k+0.5*dk/W->subN);
	return inside / (W->subN * W->subN * W->subN);
}

/* ---- allocation (reference object.c:124-226) -------------------------------------------------------------- */

static void alloc_material_tables(world W, coeffs *C, float ex, float ey, float ez)
{
	for (int m=0; m<=W->complexField; m++) {
		MALLOC(W->E[m].Jx, 0, C->N); MALLOC(W->E[m].Jy, 0, C->N); MALLOC(W->E[m].Jz, 0, C->N);
		MALLOC(W->E[m].Kx, 0, C->N); MALLOC(W->E[m].Ky, 0, C->N); MALLOC(W->E[m].Kz, 0, C->N);
	}
	MALLOC(C->NJ, 0, C->N); MALLOC(C->NK, 0, C->N);
	MALLOC(C->iMin, 0, C->N); MALLOC(C->iMax, 0, C->N);
	MALLOC(C->jMin, 0, C->N); MALLOC(C->jMax, 0, C->N);
	MALLOC(C->kMin, 0, C->N); MALLOC(C->kMax, 0, C->N);
	MALLOC(C->Jx, 0, C->N); MALLOC(C->Jy, 0, C->N); MALLOC(C->Jz, 0, C->N);
	MALLOC(C->ax, 0, C->N); MALLOC(C->ay, 0, C->N); MALLOC(C->az, 0, C->N);
	float ***tabs[] = { &C->JJx, &C->JJy, &C->JJz, &C->JKx, &C->JKy, &C->JKz, &C->JEx, &C->JEy, &C->JEz,
	                    &C->KJx, &C->KJy, &C->KJz, &C->KKx, &C->KKy, &C->KKz, &C->KEx, &C->KEy, &C->KEz };
	for (unsigned q=0; q<sizeof tabs/sizeof *tabs; q++) MALLOC(*tabs[q], 0, C->N);

	/* inverse permittivity: each component is one cell longer along its own axis */
	C->x = makeField(W->iMIN, W->iMAX+1, W->jMIN, W->jMAX, W->kMIN, W->kMAX);
	C->y = makeField(W->iMIN, W->iMAX, W->jMIN, W->jMAX+1, W->kMIN, W->kMAX);
	C->z = makeField(W->iMIN, W->iMAX, W->jMIN, W->jMAX, W->kMIN, W->kMAX+1);
	#pragma omp parallel for
	for (int i=W->iMIN; i<=W->iMAX+1; i++) for (int j=W->jMIN; j<=W->jMAX; j++) for (int k=W->kMIN; k<=W->kMAX; k++) C->x[i][j][k] = 1/ex;
	#pragma omp parallel for
	for (int i=W->iMIN; i<=W->iMAX; i++) for (int j=W->jMIN; j<=W->jMAX+1; j++) for (int k=W->kMIN; k<=W->kMAX; k++) C->y[i][j][k] = 1/ey;
	#pragma omp parallel for
	for (int i=W->iMIN; i<=W->iMAX; i++) for (int j=W->jMIN; j<=W->jMAX; j++) for (int k=W->kMIN; k<=W->kMAX+1; k++) C->z[i][j][k] = 1/ez;
}

static void alloc_dispersive(world W, coeffs *C, int n, bounds B)
{
	C->iMin[n] = B.iMin, C->iMax[n] = B.iMax + (B.iMax == W->iMAX ? 1 : 0);
	C->jMin[n] = B.jMin, C->jMax[n] = B.jMax + (B.jMax == W->jMAX ? 1 : 0);
	C->kMin[n] = B.kMin, C->kMax[n] = B.kMax + (B.kMax == W->kMAX ? 1 : 0);

	float **all[] = { &C->JJx[n], &C->JJy[n], &C->JJz[n], &C->JKx[n], &C->JKy[n], &C->JKz[n], &C->JEx[n], &C->JEy[n], &C->JEz[n] };
	float **cp[]  = { &C->KJx[n], &C->KJy[n], &C->KJz[n], &C->KKx[n], &C->KKy[n], &C->KKz[n], &C->KEx[n], &C->KEy[n], &C->KEz[n] };
	for (int q=0; q<9; q++) MALLOC(*all[q], 0, C->NK[n]);
	for (int q=0; q<9; q++) MALLOC(*cp[q], C->NJ[n], C->NK[n] - C->NJ[n]);

	C->Jx[n] = makeField(C->iMin[n], C->iMax[n], C->jMin[n], C->jMax[n], C->kMin[n], C->kMax[n]);
	C->Jy[n] = makeField(C->iMin[n], C->iMax[n], C->jMin[n], C->jMax[n], C->kMin[n], C->kMax[n]);
	C->Jz[n] = makeField(C->iMin[n], C->iMax[n], C->jMin[n], C->jMax[n], C->kMin[n], C->kMax[n]);

	/* host views of the polarisation currents (device resident; refreshed by alis_view_need_ade) */
	for (int m=0; m<=W->complexField; m++) {
		float ******J[3] = { &W->E[m].Jx, &W->E[m].Jy, &W->E[m].Jz };
		for (int c=0; c<3; c++) {
			MALLOC((*J[c])[n], 0, C->NK[n]);
			for (int p=0; p<C->NK[n]; p++) (*J[c])[n][p] = makeField(C->iMin[n], C->iMax[n], C->jMin[n], C->jMax[n], C->kMin[n], C->kMax[n]);
		}
	}
}

/* bounding boxes of every object (the cells whose NODE lies in the object and in no earlier one),
 * widened by one cell below and two above, clipped to the grid (reference object.c:229-279) */
static bounds nothing_yet(world W)       /* inverted box: any cell shrinks it to itself */
{
	return (bounds){ W->iMAX, W->iMIN, W->jMAX, W->jMIN, W->kMAX, W->kMIN };
}

static void cover(bounds *b, bounds c)   /* smallest box holding both */
{
	b->iMin = MIN(b->iMin, c.iMin), b->iMax = MAX(b->iMax, c.iMax);
	b->jMin = MIN(b->jMin, c.jMin), b->jMax = MAX(b->jMax, c.jMax);
	b->kMin = MIN(b->kMin, c.kMin), b->kMax = MAX(b->kMax, c.kMax);
}

static int owner_of_point(const object *O, float x, float y, float z)     /* first object holding the point; the list ends with S == 0 */
{
	int o = 0;
	while (O[o].S && !O[o].S(O[o], x, y, z)) o++;
	return o;
}

static bounds *object_bounds(world W, object *O, int N)
{
	bounds *B;
	MALLOC(B, 0, N);
	for (int o=0; o<N; o++) B[o] = nothing_yet(W);

	/* node coordinates once per axis instead of once per cell */
	float *x, *y, *z;
	MALLOC(x, W->iMIN, W->iMAX - W->iMIN + 1);
	MALLOC(y, W->jMIN, W->jMAX - W->jMIN + 1);
	MALLOC(z, W->kMIN, W->kMAX - W->kMIN + 1);
	for (int i=W->iMIN; i<=W->iMAX; i++) x[i] = W->itox(W, i);
	for (int j=W->jMIN; j<=W->jMAX; j++) y[j] = W->jtoy(W, j);
	for (int k=W->kMIN; k<=W->kMAX; k++) z[k] = W->ktoz(W, k);

	#pragma omp parallel
	{
		bounds *mine;                       /* per-thread boxes, merged at the end (min/max: order-free) */
		MALLOC(mine, 0, N);
		for (int o=0; o<N; o++) mine[o] = nothing_yet(W);
		#pragma omp for
		for (int i=W->iMIN; i<=W->iMAX; i++) for (int j=W->jMIN; j<=W->jMAX; j++) for (int k=W->kMIN; k<=W->kMAX; k++) {
			const int o = owner_of_point(O, x[i], y[j], z[k]);
			if (o < N) cover(&mine[o], (bounds){ i, i, j, j, k, k });
		}
		#pragma omp critical
		for (int o=0; o<N; o++) cover(&B[o], mine[o]);
		free(mine);
	}
	free(x + W->iMIN), free(y + W->jMIN), free(z + W->kMIN);

	const bounds grid = { W->iMIN, W->iMAX, W->jMIN, W->jMAX, W->kMIN, W->kMAX };
	for (int o=0; o<N; o++) {
		if (B[o].iMin > B[o].iMax || B[o].jMin > B[o].jMax || B[o].kMin > B[o].kMax) {
			printf("\rVolume of object %d/%d is zero!\n", o+1, N);
			exit(0);
		}
		/* one cell below, two above (the fill factors look at the neighbours), inside the grid */
		B[o] = (bounds){ MAX(B[o].iMin-1, grid.iMin), MIN(B[o].iMax+2, grid.iMax),
		                 MAX(B[o].jMin-1, grid.jMin), MIN(B[o].jMax+2, grid.jMax),
		                 MAX(B[o].kMin-1, grid.kMin), MIN(B[o].kMax+2, grid.kMax) };
	}
	return B;
}

/* ---- rasterisation on the device ------------------------------------------------------------
 * Default (ALIS_DEVICE_GEOMETRY=0 keeps everything on the host): the object list is flattened into alisb_shape nodes and the two
 * O(cells x sub-samples) loops -- object_bounds and fill_factor -- run on the GPU
 * (alisb_geometry_bounds / alisb_geometry_fill).  Everything that needs libm stays here: the cos/sin
 * of Rotation / CLattice angles and the coordinates of every sample position (uniform, graded or
 * user grid functions alike) are tabulated with the same expressions the host path evaluates, so the
 * fill factors are bit-identical.  Objects with a user-written shape, ImportContour height maps
 * and subN < 0 keep the host path. */
typedef struct {
	alisb_geometry g;
	alisb_shape *nodes; int nnodes, capnodes;
	float *aux; int naux, capaux;
	int32_t *roots;
	float *coord[3];
} geo_ctx;

static int geo_type(shape *S)
{
	shape *known[ALISB_SHAPE_COUNT] = { Box, Ball, Diamond, BiconeX, BiconeY, BiconeZ, RodX, RodY, RodZ, DRodX, DRodY, DRodZ,
		HRodXY, HRodYZ, HRodZX, HRodXZ, HRodYX, HRodZY, Translation, Rotation, Combination, Intersection, Difference,
		Lattice, DLattice, CLatticeX, CLatticeY, CLatticeZ };
	for (int t=0; t<ALISB_SHAPE_COUNT; t++) if (S == known[t]) return t;
	return -1;
}

static int geo_new_nodes(geo_ctx *G, int count)
{
	while (G->nnodes + count > G->capnodes) { G->capnodes = G->capnodes ? 2*G->capnodes : 32; G->nodes = xrealloc(G->nodes, sizeof(alisb_shape)*G->capnodes); }
	memset(G->nodes + G->nnodes, 0, sizeof(alisb_shape)*count);
	G->nnodes += count;
	return G->nnodes - count;
}

static void geo_push_aux(geo_ctx *G, float v)
{
	if (G->naux == G->capaux) { G->capaux = G->capaux ? 2*G->capaux : 32; G->aux = xrealloc(G->aux, sizeof(float)*G->capaux); }
	G->aux[G->naux++] = v;
}

/* object O -> node `slot` (children appended behind); 0 when some shape has no device twin */
static int geo_flatten(geo_ctx *G, object O, int slot)
{
	const int type = geo_type(O.S);
	if (type < 0) return 0;
	int count = 0;
	if (type > ALISB_SHAPE_HRODZY) {
		if (!O.p) return 0;
		count = (type == ALISB_SHAPE_COMBINATION || type == ALISB_SHAPE_INTERSECTION || type == ALISB_SHAPE_DIFFERENCE) ? (int)O.f[0][0] : 1;
		if (count < 1) return 0;
	}
	int aux = 0;
	if (type == ALISB_SHAPE_ROTATION) {                   /* shapes.c:133-140 */
		aux = G->naux;
		for (int q=0; q<3; q++) { geo_push_aux(G, cosf(O.f[0][q]*PI/180.)); geo_push_aux(G, sinf(O.f[0][q]*PI/180.)); }
	}
	if (type >= ALISB_SHAPE_CLATTICEX) {                  /* shapes.c:191-221 */
		aux = G->naux;
		const int copies = O.f[0][0];
		for (int q=0; q<copies; q++) { float theta = 2 * PI * q / copies; geo_push_aux(G, cosf(theta)); geo_push_aux(G, sinf(theta)); }
	}
	const int first = count ? geo_new_nodes(G, count) : 0;
	alisb_shape *n = &G->nodes[slot];
	n->type = type; n->child = first; n->nchild = count; n->aux = aux;
	for (int a=0; a<3; a++) for (int b=0; b<3; b++) n->f[a][b] = O.f[a][b];
	object *child = O.p;
	for (int q=0; q<count; q++) if (!geo_flatten(G, child[q], first+q)) return 0;
	return 1;
}

static void geo_free(geo_ctx *G)
{
	if (!G) return;
	free(G->nodes); free(G->aux); free(G->roots);
	for (int a=0; a<3; a++) free(G->coord[a]);
	free(G);
}

static unsigned long geo_objects;
unsigned long alisDeviceGeometryObjects(void) { return geo_objects; }   /* objects rasterised by the engine so far (diagnostics) */

int alis_device_geometry(void)
{
	static int mode = -1;
	if (mode < 0) { const char *s = getenv("ALIS_DEVICE_GEOMETRY"); mode = s ? atoi(s) != 0 : 1; }     /* on unless ALIS_DEVICE_GEOMETRY=0 */
	return mode;
}

/* grids below this many cells are rasterised on the host: the device path costs a CUDA start-up plus a stack
 * resize before the first step (measured on the B200: putObjects of the 1e5..6e5-cell named examples 1.0-3.6 s
 * on the device against 0.1-0.3 s on 16 host threads, profiles/r02_named_profiles_v4.txt; at 22.7 M cells the
 * device wins 1.5 s to 3.7 s).  ALIS_DEVICE_GEOMETRY_MIN overrides. */
static long geo_min_cells(void)
{
	static long n = -1;
	if (n < 0) { const char *s = getenv("ALIS_DEVICE_GEOMETRY_MIN"); n = s ? atol(s) : 4000000L; }
	return n;
}

static geo_ctx *geo_build(world W, object *O, int N)
{
	if (!alis_device_geometry() || W->subN < 1 || N < 1) return 0;
	if ((long)W->iNUM*W->jNUM*W->kNUM < geo_min_cells()) return 0;
	geo_ctx *G = xcalloc(1, sizeof *G);
	G->roots = xmalloc(sizeof(int32_t)*N);
	for (int o=0; o<N; o++) {
		G->roots[o] = geo_new_nodes(G, 1);
		if (!geo_flatten(G, O[o], G->roots[o])) { geo_free(G); return 0; }
	}
	/* sample coordinates, [(idx-lo)*2 + half][slot] (object.c:72, 104-118 evaluate W->itox at exactly these arguments) */
	const int lo[3] = { W->iMIN, W->jMIN, W->kMIN }, hi[3] = { W->iMAX+1, W->jMAX+1, W->kMAX+1 };
	float (*map[3])(world, float) = { W->itox, W->jtoy, W->ktoz };
	const int slots = W->subN + 3;
	const float edge = (W->subN-1.0) / (W->subN*2.0);
	for (int a=0; a<3; a++) {
		G->coord[a] = xmalloc(sizeof(float)*2*slots*(hi[a]-lo[a]+1));
		for (int idx=lo[a]; idx<=hi[a]; idx++) for (int half=0; half<2; half++) {
			float *t = G->coord[a] + ((size_t)(idx-lo[a])*2 + half)*slots;
			const float base = half ? idx-0.5 : idx;
			t[0] = map[a](W, base);
			t[1] = map[a](W, base + -edge);
			t[2] = map[a](W, base + edge);
			for (int q=0; q<W->subN; q++) { const int d = 1-W->subN+2*q; t[3+q] = map[a](W, base+0.5*d/W->subN); }
		}
		G->g.lo[a] = lo[a]; G->g.hi[a] = hi[a]; G->g.coord[a] = G->coord[a];
	}
	G->g.nobjects = N; G->g.roots = G->roots; G->g.nnodes = G->nnodes; G->g.nodes = G->nodes;
	G->g.naux = G->naux; G->g.aux = G->aux; G->g.subn = W->subN;
	return G;
}

/* object_bounds with the node scan on the device */
static bounds *object_bounds_device(world W, geo_ctx *G, int N)
{
	bounds *B;
	MALLOC(B, 0, N);
	int32_t *raw = xmalloc(sizeof(int32_t)*6*N);
	const int32_t box[6] = { W->iMIN, W->iMAX, W->jMIN, W->jMAX, W->kMIN, W->kMAX };
	if (alisb_geometry_bounds(alis_bridge_device(), &G->g, box, raw) != ALISB_OK) {
		printf("alis-b200: device rasteriser unavailable (%s), rasterising on the host\n", alisb_last_error());
		free(raw); free(B);
		return 0;
	}
	for (int o=0; o<N; o++) {
		B[o].iMin = raw[6*o], B[o].iMax = raw[6*o+1], B[o].jMin = raw[6*o+2], B[o].jMax = raw[6*o+3], B[o].kMin = raw[6*o+4], B[o].kMax = raw[6*o+5];
		if (B[o].iMin > B[o].iMax || B[o].jMin > B[o].jMax || B[o].kMin > B[o].kMax) {
			printf("\rVolume of object %d/%d is zero!\n", o+1, N);
			exit(0);
		}
		B[o].iMin = MAX(B[o].iMin-1, W->iMIN), B[o].iMax = MIN(B[o].iMax+2, W->iMAX);
		B[o].jMin = MAX(B[o].jMin-1, W->jMIN), B[o].jMax = MIN(B[o].jMax+2, W->jMAX);
		B[o].kMin = MAX(B[o].kMin-1, W->kMIN), B[o].kMax = MIN(B[o].kMax+2, W->kMAX);
	}
	free(raw);
	return B;
}

/* the three fill-factor arrays of object o over its box widened by the boundary layers, from the
 * device: f[c][((i-i0)*nj + (j-j0))*nk + (k-k0)] */
typedef struct { float *f[3]; int i0, j0, k0, nj, nk; } geo_fill_t;

static geo_fill_t geo_fill_object(world W, geo_ctx *G, int o, bounds b)
{
	geo_fill_t F;
	const int32_t box[6] = { b.iMin, b.iMax + (b.iMax == W->iMAX), b.jMin, b.jMax + (b.jMax == W->jMAX), b.kMin, b.kMax + (b.kMax == W->kMAX) };
	const size_t n = (size_t)(box[1]-box[0]+1)*(box[3]-box[2]+1)*(box[5]-box[4]+1);
	for (int c=0; c<3; c++) F.f[c] = xmalloc(sizeof(float)*n);
	F.i0 = box[0]; F.j0 = box[2]; F.k0 = box[4]; F.nj = box[3]-box[2]+1; F.nk = box[5]-box[4]+1;
	if (alisb_geometry_fill(alis_bridge_device(), &G->g, o, box, F.f[0], F.f[1], F.f[2]) != ALISB_OK) {
		printf("alis-b200: device rasteriser unavailable (%s), rasterising on the host\n", alisb_last_error());
		for (int c=0; c<3; c++) { free(F.f[c]); F.f[c] = 0; }
		return F;                                   /* f[0] == 0: the caller switches to the host rasteriser */
	}
	geo_objects++;
	return F;
}
#define GEO_AT(F,c,i,j,k) ((F).f[c][((size_t)((i)-(F).i0)*(F).nj + ((j)-(F).j0))*(F).nk + ((k)-(F).k0)])

/* fill factors of the three staggered components + the pole coefficients of dispersive object n
 * (= list entry o); tri = 1 for anisotropic matter (reference object.c:290-336) */
static void set_dispersive(world W, coeffs *C, matter *M, object *O, int n, int o, int tri, const geo_fill_t *dev)
{
	if (dev) {
		#pragma omp parallel for
		for (int i=C->iMin[n]; i<=C->iMax[n]; i++) for (int j=C->jMin[n]; j<=C->jMax[n]; j++) for (int k=C->kMin[n]; k<=C->kMax[n]; k++) {
			C->Jx[n][i][j][k] = GEO_AT(*dev, 0, i, j, k);
			C->Jy[n][i][j][k] = GEO_AT(*dev, 1, i, j, k);
			C->Jz[n][i][j][k] = GEO_AT(*dev, 2, i, j, k);
		}
	}
	else {
	#pragma omp parallel for schedule(guided)
	for (int i=C->iMin[n]; i<=C->iMax[n]; i++) for (int j=C->jMin[n]; j<=C->jMax[n]; j++) for (int k=C->kMin[n]; k<=C->kMax[n]; k++) {
		C->Jx[n][i][j][k] = fill_factor(W, O, o, i-0.5, j, k);
		C->Jy[n][i][j][k] = fill_factor(W, O, o, i, j-0.5, k);
		C->Jz[n][i][j][k] = fill_factor(W, O, o, i, j, k-0.5);
	}
	}

	C->ax[n] = PI * W->dt * conductivity(W, M[o], 0) / (M[o].e.x ? M[o].e.x : M[o].e.x);
	C->ay[n] = PI * W->dt * conductivity(W, M[o], 1) / (M[o].e.y ? M[o].e.y : M[o].e.x);
	C->az[n] = PI * W->dt * conductivity(W, M[o], 2) / (M[o].e.z ? M[o].e.z : M[o].e.x);

	float *JJ[3] = { C->JJx[n], C->JJy[n], C->JJz[n] }, *JK[3] = { C->JKx[n], C->JKy[n], C->JKz[n] }, *JE[3] = { C->JEx[n], C->JEy[n], C->JEz[n] };
	float *KJ[3] = { C->KJx[n], C->KJy[n], C->KJz[n] }, *KK[3] = { C->KKx[n], C->KKy[n], C->KKz[n] }, *KE[3] = { C->KEx[n], C->KEy[n], C->KEz[n] };
	int drude = 0, cp = C->NJ[n];                 /* Drude poles first, then the two-variable ones */
	for (int p=0; p<MAXPOLES; p+=1+2*tri) {
		if (!M[o].P[p].w) continue;
		if (!M[o].P[p].f && !M[o].P[p].g) {
			for (int c=0; c<3; c++) {
				JJ[c][drude] = drude_a(W, M[o], p+c*tri);
				JE[c][drude] = drude_b(W, M[o], p+c*tri) / eps_eff(W, M[o], c);
			}
			drude++;
		}
		else {
			for (int c=0; c<3; c++) {
				JJ[c][cp] = creal(cp_a(W, M[o], p+c*tri));
				JK[c][cp] =-cimag(cp_a(W, M[o], p+c*tri));
				JE[c][cp] = creal(cp_b(W, M[o], p+c*tri)) / eps_eff(W, M[o], c);
				KJ[c][cp] =-JK[c][cp] / JJ[c][cp];
				KK[c][cp] = JK[c][cp] * JK[c][cp] / JJ[c][cp] + JJ[c][cp];
				KE[c][cp] = JE[c][cp] * JK[c][cp] / JJ[c][cp] + cimag(cp_b(W, M[o], p+c*tri)) / eps_eff(W, M[o], c);
			}
			cp++;
		}
	}
}

/* eps^-1 <- 1/(1/eps^-1 + d_eps * fill): arithmetic averaging of the permittivity inside a cell */
#define BLEND(C, de, fill) (C) = 1 / (1 / (C) + (de) * (fill))

/* reference object.c:339-405 */
void putObjectArray(world W, matter *M, object *O)
{
	removeObjects(W);
	int N;
	for (N=0; O[N].S; N++) if (needs_aux(M[N])) W->CE->N++;
	printf("\rPreparing to put %d objects.\n", N);
	/* sub-pixel sampling defaults: 1 if anything is dispersive, else 12 per axis */
	W->subN = W->Res.spec.subN = W->Res.spec.subN ? W->Res.spec.subN : W->CE->N ? 1 : 12;
	geo_ctx *G = geo_build(W, O, N);                /* non-NULL: rasterise on the device (default; ALIS_DEVICE_GEOMETRY=0 or a user-written shape: host) */
	bounds *B = G ? object_bounds_device(W, G, N) : 0;
	if (!B) { geo_free(G); G = 0; B = object_bounds(W, O, N); }
	coeffs *C = W->CE;
	alloc_material_tables(W, C, eps_eff(W, M[N], 0), eps_eff(W, M[N], 1), eps_eff(W, M[N], 2));
	MALLOC(C->O, 0, N+1);
	MALLOC(C->M, 0, N+1);
	C->M[N] = M[N];                                 /* background matter */

	/* last object first: earlier list entries take precedence (owned_by) */
	for (int o=N-1, n=C->N; o>=0; o--) {
		C->O[o] = O[o];
		C->M[o] = M[o];
		float cx = eps_eff(W, M[o], 0) - eps_eff(W, M[N], 0);
		float cy = eps_eff(W, M[o], 1) - eps_eff(W, M[N], 1);
		float cz = eps_eff(W, M[o], 2) - eps_eff(W, M[N], 2);
		const bounds b = B[o];
		geo_fill_t dev;
		if (G) { dev = geo_fill_object(W, G, o, b); if (!dev.f[0]) { geo_free(G); G = 0; } }

		printf("\rPutting objects: %d left.", o+1);
		fflush(stdout);

		if (needs_aux(M[o])) {
			n--;
			int tri = M[o].e.y || M[o].e.z ? 1 : 0;
			for (int p=0; p<MAXPOLES; p++) {
				if (M[o].P[p].w && !M[o].P[p].f && !M[o].P[p].g) C->NJ[n]++;
				if (M[o].P[p].w) C->NK[n]++;
			}
			if (tri) C->NJ[n] /= 3, C->NK[n] /= 3;
			alloc_dispersive(W, C, n, b);
			set_dispersive(W, C, M, O, n, o, tri, G ? &dev : 0);

			#pragma omp parallel for
			for (int i=b.iMin; i<=b.iMax; i++) for (int j=b.jMin; j<=b.jMax; j++) for (int k=b.kMin; k<=b.kMax; k++) {
				BLEND(C->x[i][j][k], cx, C->Jx[n][i][j][k]);
				BLEND(C->y[i][j][k], cy, C->Jy[n][i][j][k]);
				BLEND(C->z[i][j][k], cz, C->Jz[n][i][j][k]);
			}
			if (b.iMax == W->iMAX) for (int j=b.jMin; j<=b.jMax; j++) for (int k=b.kMin; k<=b.kMax; k++)
				BLEND(C->x[b.iMax+1][j][k], cx, C->Jx[n][b.iMax+1][j][k]);
			if (b.jMax == W->jMAX) for (int i=b.iMin; i<=b.iMax; i++) for (int k=b.kMin; k<=b.kMax; k++)
				BLEND(C->y[i][b.jMax+1][k], cy, C->Jy[n][i][b.jMax+1][k]);
			if (b.kMax == W->kMAX) for (int i=b.iMin; i<=b.iMax; i++) for (int j=b.jMin; j<=b.jMax; j++)
				BLEND(C->z[i][j][b.kMax+1], cz, C->Jz[n][i][j][b.kMax+1]);
		}
		else if (G) {
			#pragma omp parallel for
			for (int i=b.iMin; i<=b.iMax; i++) for (int j=b.jMin; j<=b.jMax; j++) for (int k=b.kMin; k<=b.kMax; k++) {
				BLEND(C->x[i][j][k], cx, GEO_AT(dev, 0, i, j, k));
				BLEND(C->y[i][j][k], cy, GEO_AT(dev, 1, i, j, k));
				BLEND(C->z[i][j][k], cz, GEO_AT(dev, 2, i, j, k));
			}
			if (b.iMax == W->iMAX) for (int j=b.jMin; j<=b.jMax; j++) for (int k=b.kMin; k<=b.kMax; k++)
				BLEND(C->x[b.iMax+1][j][k], cx, GEO_AT(dev, 0, b.iMax+1, j, k));
			if (b.jMax == W->jMAX) for (int i=b.iMin; i<=b.iMax; i++) for (int k=b.kMin; k<=b.kMax; k++)
				BLEND(C->y[i][b.jMax+1][k], cy, GEO_AT(dev, 1, i, b.jMax+1, k));
			if (b.kMax == W->kMAX) for (int i=b.iMin; i<=b.iMax; i++) for (int j=b.jMin; j<=b.jMax; j++)
				BLEND(C->z[i][j][b.kMax+1], cz, GEO_AT(dev, 2, i, j, b.kMax+1));
		}
		else {
			#pragma omp parallel for schedule(guided)
			for (int i=b.iMin; i<=b.iMax; i++) for (int j=b.jMin; j<=b.jMax; j++) for (int k=b.kMin; k<=b.kMax; k++) {
				BLEND(C->x[i][j][k], cx, fill_factor(W, O, o, i-0.5, j, k));
				BLEND(C->y[i][j][k], cy, fill_factor(W, O, o, i, j-0.5, k));
				BLEND(C->z[i][j][k], cz, fill_factor(W, O, o, i, j, k-0.5));
			}
			if (b.iMax == W->iMAX) for (int j=b.jMin; j<=b.jMax; j++) for (int k=b.kMin; k<=b.kMax; k++)
				BLEND(C->x[b.iMax+1][j][k], cx, fill_factor(W, O, o, b.iMax+0.5, j, k));
			if (b.jMax == W->jMAX) for (int i=b.iMin; i<=b.iMax; i++) for (int k=b.kMin; k<=b.kMax; k++)
				BLEND(C->y[i][b.jMax+1][k], cy, fill_factor(W, O, o, i, b.jMax+0.5, k));
			if (b.kMax == W->kMAX) for (int i=b.iMin; i<=b.iMax; i++) for (int j=b.jMin; j<=b.jMax; j++)
				BLEND(C->z[i][j][b.kMax+1], cz, fill_factor(W, O, o, i, j, b.kMax+0.5));
		}
		if (G) for (int c=0; c<3; c++) free(dev.f[c]);
	}

	free(B);
	geo_free(G);
	printf("\rPutting objects: %d done.\n", N);
	fflush(stdout);
	alis_host_dirty(W, ALIS_DIRTY_MATERIALS | ALIS_DIRTY_SOURCES);
}

/* putObjects(W, M1, O1, M2, O2, ..., Mbackground): the macro appends a zeroed object as
 * terminator; matter and object structs travel by value through the varargs
 * (reference object.c:408-429) */
void (putObjects)(world W, ...)
{
	ALIS_PROF("putObjects");
	int N = 0;
	va_list ap;
	matter *M;
	object *O;

	va_start(ap, W);
	while (va_arg(ap, matter), va_arg(ap, object).S) N++;
	va_end(ap);
	MALLOC(M, 0, N+1);
	MALLOC(O, 0, N+1);
	va_start(ap, W);
	for (int o=0; o<=N; o++) {
		M[o] = va_arg(ap, matter);
		O[o] = va_arg(ap, object);
	}
	va_end(ap);
	putObjectArray(W, M, O);
	free(M);
	free(O);
}

/* reference object.c:432-497 */
void removeObjects(world W)
{
	coeffs *C = W->CE;
	if (!(C->O && C->M)) return;
	for (int m=0; m<=W->complexField; m++) {
		float *****J[3] = { W->E[m].Jx, W->E[m].Jy, W->E[m].Jz };
		for (int n=0; n<C->N; n++) for (int c=0; c<3; c++) {
			for (int p=0; p<C->NK[n]; p++) removeField(J[c][n][p], C->iMin[n], C->jMin[n], C->kMin[n]);
			free(J[c][n]);
		}
		free(W->E[m].Jx); free(W->E[m].Jy); free(W->E[m].Jz);
		free(W->E[m].Kx); free(W->E[m].Ky); free(W->E[m].Kz);
		W->E[m].Jx = W->E[m].Jy = W->E[m].Jz = W->E[m].Kx = W->E[m].Ky = W->E[m].Kz = 0;
	}
	for (int n=0; n<C->N; n++) {
		removeField(C->Jx[n], C->iMin[n], C->jMin[n], C->kMin[n]);
		removeField(C->Jy[n], C->iMin[n], C->jMin[n], C->kMin[n]);
		removeField(C->Jz[n], C->iMin[n], C->jMin[n], C->kMin[n]);
		free(C->JJx[n]); free(C->JKx[n]); free(C->JEx[n]);
		free(C->JJy[n]); free(C->JKy[n]); free(C->JEy[n]);
		free(C->JJz[n]); free(C->JKz[n]); free(C->JEz[n]);
		free(C->KJx[n]+C->NJ[n]); free(C->KKx[n]+C->NJ[n]); free(C->KEx[n]+C->NJ[n]);
		free(C->KJy[n]+C->NJ[n]); free(C->KKy[n]+C->NJ[n]); free(C->KEy[n]+C->NJ[n]);
		free(C->KJz[n]+C->NJ[n]); free(C->KKz[n]+C->NJ[n]); free(C->KEz[n]+C->NJ[n]);
	}
	removeField(C->x, W->iMIN, W->jMIN, W->kMIN);
	removeField(C->y, W->iMIN, W->jMIN, W->kMIN);
	removeField(C->z, W->iMIN, W->jMIN, W->kMIN);
	free(C->NJ); free(C->NK);
	free(C->iMin); free(C->iMax); free(C->jMin); free(C->jMax); free(C->kMin); free(C->kMax);
	free(C->Jx); free(C->Jy); free(C->Jz);
	free(C->ax); free(C->ay); free(C->az);
	free(C->JJx); free(C->JKx); free(C->JEx); free(C->KJx); free(C->KKx); free(C->KEx);
	free(C->JJy); free(C->JKy); free(C->JEy); free(C->KJy); free(C->KKy); free(C->KEy);
	free(C->JJz); free(C->JKz); free(C->JEz); free(C->KJz); free(C->KKz); free(C->KEz);
	free(C->O);
	free(C->M);
	C->O = 0; C->M = 0; C->N = 0;
	alis_host_dirty(W, ALIS_DIRTY_MATERIALS);
}

/* loss scaling of all poles ("low temperature" factor T), reference object.c:500-507 */
matter LT(matter M, float T)
{
	for (int p=0; p<MAXPOLES; p++) {
		M.P[p].r *= T;
		M.P[p].g *= T;
	}
	return M;
}

/* constant-(n,k) matter interpolated from a text table of "wavelength n k" rows at wavelength l
 * (reference object.c:510-549) */
matter nkFile(char *name, float l, float f)
{
	FILE *file = fopen(name, "r");
	if (!file) {
		printf("The (n,k) file not found.\n");
		exit(0);
	}
	int rows = 0, cap = 0;
	float *L = 0, *Nr = 0, *Kr = 0;
	char line[1024];
	while (fgets(line, 256, file)) {
		if (cap <= rows) {
			cap += 256;
			L = realloc(L, cap*sizeof(float)); Nr = realloc(Nr, cap*sizeof(float)); Kr = realloc(Kr, cap*sizeof(float));
		}
		char *s = line;
		L[rows] = atof(s);
		s = strchr(s,'\t') ? strchr(s,'\t')+1 : strchr(s,' ') ? strchr(s,' ')+1 : strchr(s,',')+1;
		Nr[rows] = atof(s);
		s = strchr(s,'\t') ? strchr(s,'\t')+1 : strchr(s,' ') ? strchr(s,' ')+1 : strchr(s,',')+1;
		Kr[rows] = atof(s);
		rows++;
	}
	fclose(file);
	int q = 1;
	while (q<rows-1 && l>L[q]) q++;
	float n = (Nr[q-1]*(L[q]-l) + Nr[q]*(l-L[q-1])) / (L[q]-L[q-1]);
	float k = (Kr[q-1]*(L[q]-l) + Kr[q]*(l-L[q-1])) / (L[q]-L[q-1]);
	free(L); free(Nr); free(Kr);
	return (matter) {{(n)*(n)-(k)*(k)},{0,2*n*k*((float)f)}};
}
