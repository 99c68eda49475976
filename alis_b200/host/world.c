/*
 * world.c -- grid construction, boundary resolution, 1-D spacing and CPML tables, world life
 * cycle.  Host C; produces what the engine consumes through engine_bridge.c.
 *
 * Behaviour follows reference src/world.c (cited per function).  The arithmetic keeps the
 * reference's float/double mix expression by expression, because the tables built here feed the
 * bit-exact parity tests (tests/test_host_layer.py).
 */
#include <string.h>
#include "alis_c.h"
#include "host_internal.h"
#include "h5lite.h"

/* ---- coordinate <-> index maps ------------------------------------------------------------ */

/* nearest index to coordinate x for a monotone index->coordinate map: 16-step bisection over
 * +-2^16, then a +-1 refinement (reference world.c:12-25) */
static int nearest_index(world W, float (*coord)(world, float), float x)
{
	int i = 0;
	for (int bit=15; bit>=0; bit--) {
		float y = coord(W, i);
		if (x > y) i += 1u<<bit;
		if (x < y) i -= 1u<<bit;
	}
	float d = fabsf(x - coord(W, i));
	if (fabsf(x - coord(W, i+1)) < d) i++;
	if (fabsf(x - coord(W, i-1)) < d) i--;
	return i;
}

/* graded grid: uniform spacing ds between min and max, then geometric growth with ratio r over n
 * cells up to m*ds, then uniform m*ds (reference world.c:27-42).  r solves
 * r^n - (d*m - m/2 - 1/2) ln r - 1 = 0 by Newton iteration, d = 10. */
static float graded_coord(int m, float min, float max, float ds, float i)
{
	float d = 10;
	float iLo = min/ds-0.5, iHi = max/ds+0.5;
	float r = d/(d-1), prev;
	int n = floor((d*m-0.5*m-0.5)*logf(m)/(m-1)+0.5);

	do prev = r, r = prev - (powf(prev,n)-(d*m-0.5*m-0.5)*logf(prev)-1)/(n*powf(prev,n-1)-(d*m-0.5*m-0.5)/prev);
	while (fabsf(r-prev) > 1E-6);

	if (i >= iHi+n) return ds * (iHi + (d*m-0.5*m-0.5) + (i-iHi-n) * m);
	if (i <= iLo-n) return ds * (iLo - (d*m-0.5*m-0.5) + (i-iLo+n) * m);
	if (i > iHi) return ds * (iHi + (powf(r,i-iHi)-1) / logf(r));
	if (i < iLo) return ds * (iLo - (powf(r,iLo-i)-1) / logf(r));
	return ds * i;
}

static int x_to_i(world W, float x) { return nearest_index(W, W->itox, x); }
static int y_to_j(world W, float y) { return nearest_index(W, W->jtoy, y); }
static int z_to_k(world W, float z) { return nearest_index(W, W->ktoz, z); }

static float i_to_x_uniform(world W, float i) { return W->dx*i; }
static float j_to_y_uniform(world W, float j) { return W->dy*j; }
static float k_to_z_uniform(world W, float k) { return W->dz*k; }

static float i_to_x_graded(world W, float i) { return graded_coord(W->Res.nug.x.m, W->Res.nug.x.Min, W->Res.nug.x.Max, W->dx, i); }
static float j_to_y_graded(world W, float j) { return graded_coord(W->Res.nug.y.m, W->Res.nug.y.Min, W->Res.nug.y.Max, W->dy, j); }
static float k_to_z_graded(world W, float k) { return graded_coord(W->Res.nug.z.m, W->Res.nug.z.Min, W->Res.nug.z.Max, W->dz, k); }

/* ---- host array helpers (reference world.c:58-85) --------------------------------------------- */

float ***makeField(int iMin, int iMax, int jMin, int jMax, int kMin, int kMax)
{
	const long ni = 1-iMin+iMax, nj = 1-jMin+jMax, nk = 1-kMin+kMax;
	float ***f;
	MALLOC(f, iMin, ni);
	MALLOC(f[iMin], jMin, ni*nj);
	MALLOC(f[iMin][jMin], kMin, ni*nj*nk);
	for (int i=iMin; i<=iMax; i++) {
		f[i] = f[iMin] + nj*(i-iMin);
		for (int j=jMin; j<=jMax; j++) f[i][j] = f[iMin][jMin] + nk*(nj*(i-iMin)+(j-jMin));
	}
	return f;
}

float complex ***makeComplexField(int iMin, int iMax, int jMin, int jMax, int kMin, int kMax)
{
	const long ni = 1-iMin+iMax, nj = 1-jMin+jMax, nk = 1-kMin+kMax;
	float complex ***f;
	MALLOC(f, iMin, ni);
	MALLOC(f[iMin], jMin, ni*nj);
	MALLOC(f[iMin][jMin], kMin, ni*nj*nk);
	for (int i=iMin; i<=iMax; i++) {
		f[i] = f[iMin] + nj*(i-iMin);
		for (int j=jMin; j<=jMax; j++) f[i][j] = f[iMin][jMin] + nk*(nj*(i-iMin)+(j-jMin));
	}
	return f;
}

/* ---- CPML / ADE-PML coefficient tables (reference world.c:98-132) ------------------------------ */
/* Per PML cell n = 1..Num (counted from the interior):  graded sigma, kappa, alpha;
 * dtds is divided by kappa IN PLACE; the tables are C0 = Ca, C1 = -dtds*Cb, C2 = +dtds*Cb with
 * Ca = (1 - (sigma*dtds + alpha*dt)/2) / (1 + ...),  Cb = sigma*dtds / (1 + ...).
 * Defaults: sigmaMax 1, grading power 3.5; Num < 24 -> kappa 1, alpha 0 (plain ADE-PML),
 * Num >= 24 -> kappaMax 10, alphaMax 0.01 (CPML). */
static void pml_tables(sur S, int offset, float **C0, float **C1, float **C2, float *dtds, float dt,
                       int minSur, int maxSur, int Min, int Max, int MINi, int NUM)
{
	if (!S.pml.SigmaMax) S.pml.SigmaMax = 1;
	if (!S.pml.KappaMax) S.pml.KappaMax = S.pml.Num < 24 ? 1 : 10;
	if (!S.pml.AlphaMax) S.pml.AlphaMax = S.pml.Num < 24 ? 0 : 0.01;
	if (!S.pml.SigmaPower) S.pml.SigmaPower = 3.5;
	if (!S.pml.AlphaPower) S.pml.AlphaPower = 1;
	if (minSur != PML && maxSur != PML) return;

	MALLOC(*C0, MINi, NUM);
	MALLOC(*C1, MINi, NUM);
	MALLOC(*C2, MINi, NUM);
	for (int n=1; n<=S.pml.Num; n++) {
		float Sigma = 0.8*S.pml.SigmaMax*(S.pml.SigmaPower+1)*powf((n-0.5*offset)/S.pml.Num, S.pml.SigmaPower);
		float Kappa = 1+(S.pml.KappaMax-1)*powf((n-0.5*offset)/S.pml.Num, S.pml.SigmaPower);
		float Alpha = S.pml.AlphaMax*powf((S.pml.Num-(n-0.5*offset))/S.pml.Num, S.pml.AlphaPower);
		int cell[2] = { Min-n-1+offset, Max+n+1 };
		int on[2] = { minSur==PML, maxSur==PML };
		for (int side=0; side<2; side++) if (on[side]) {
			const int q = cell[side];
			dtds[q] /= Kappa;
			float Ca = (1-0.5*(Sigma*dtds[q]+Alpha*dt))/(1+0.5*(Sigma*dtds[q]+Alpha*dt));
			float Cb = Sigma*dtds[q]/(1+0.5*(Sigma*dtds[q]+Alpha*dt));
			(*C0)[q] = Ca;
			(*C1)[q] = -dtds[q]*Cb;
			(*C2)[q] = +dtds[q]*Cb;
		}
	}
}

/* spacings and dt/ds tables of one field set; offset 0 for E (node-centred widths), 1 for H
 * (cell widths) (reference world.c:135-161) */
static coeffs *make_set_coeffs(world W, sur S, int offset)
{
	coeffs *C;
	MALLOC(C, 0, 1);
	S.pml.Num = S.pml.Num ? S.pml.Num : 12;

	/* per axis: spacing table (mirrored below 0 on a symmetric axis), dt/ds table, then the PML tables,
	 * which divide dt/ds by kappa in place */
	struct {
		float **d, **dtd, **P0, **P1, **P2;
		int mirror, min_sur, max_sur, Min, Max, MIN, MAX, NUM;
		float (*to_coord)(world, float);
	} ax[3] = {
		{ &C->dx, &C->dtdx, &C->xP, &C->xPy, &C->xPz, W->xMinSur==SYM, W->xMinSur, W->xMaxSur, W->iMin, W->iMax, W->iMIN, W->iMAX, W->iNUM, W->itox },
		{ &C->dy, &C->dtdy, &C->yP, &C->yPz, &C->yPx, W->yMinSur==SYM, W->yMinSur, W->yMaxSur, W->jMin, W->jMax, W->jMIN, W->jMAX, W->jNUM, W->jtoy },
		{ &C->dz, &C->dtdz, &C->zP, &C->zPx, &C->zPy, W->zMinSur==SYM, W->zMinSur, W->zMaxSur, W->kMin, W->kMax, W->kMIN, W->kMAX, W->kNUM, W->ktoz },
	};
	for (int a=0; a<3; a++) {
		float *d, *dtd;
		MALLOC(d, ax[a].mirror ? -ax[a].MAX : ax[a].MIN, ax[a].mirror ? 2*ax[a].NUM-1 : ax[a].NUM);
		MALLOC(dtd, ax[a].MIN, ax[a].NUM);
		for (int q=ax[a].MIN; q<=ax[a].MAX; q++) {
			d[q] = ax[a].to_coord(W, q-0.5*(offset-1)) - ax[a].to_coord(W, q-0.5*(offset+1));
			dtd[q] = W->dt/d[q];
		}
		if (ax[a].mirror) for (int q=1; q<=ax[a].MAX; q++) d[-q] = d[q];
		*ax[a].d = d, *ax[a].dtd = dtd;
		pml_tables(S, offset, ax[a].P0, ax[a].P1, ax[a].P2, dtd, W->dt, ax[a].min_sur, ax[a].max_sur, ax[a].Min, ax[a].Max, ax[a].MIN, ax[a].NUM);
	}
	C->N = 0;
	return C;
}

/* Field sets: the arrays themselves live on the device (engine); the host struct only carries
 * the time stamp and lazily created host views (see view.c).  (reference world.c:164-189
 * allocates them with calloc here.) */
static vfield *make_field_set(world W)
{
	vfield *F;
	MALLOC(F, 0, 1+W->complexField);
	return F;
}

/* where one axis' index ranges and extents live in the world, and what the caller asked for */
typedef struct {
	int *Min, *Max, *Num, *MIN, *MAX, *NUM;          /* interior range; range with the PML shells      */
	float *lo, *hi, *size, *LO, *HI;                 /* interior extent; extent with the shells        */
	int min_sur, max_sur;
	float want_lo, want_hi;
	int (*to_index)(world, float);
	float (*to_coord)(world, float);
} axis_slots;

/* index ranges and coordinate extents of one axis.  An axis with no extent given (0, 0) collapses to
 * a 3-cell strip (2 with a mirror); a mirrored axis starts at index 0 and its extent is +-max */
static void lay_out_axis(world W, axis_slots A, int shell)
{
	const int given = A.want_lo || A.want_hi, mirror = A.min_sur==SYM;
	*A.Min = mirror ? 0 : given ? A.to_index(W, A.want_lo) : -1;
	*A.Max = given ? A.to_index(W, A.want_hi) : 1;
	*A.Num = 1 - *A.Min + *A.Max;
	*A.MIN = *A.Min - (A.min_sur==PML ? shell : 0);
	*A.MAX = *A.Max + (A.max_sur==PML ? shell : 0);
	*A.NUM = 1 - *A.MIN + *A.MAX;

	*A.hi = A.to_coord(W, *A.Max);
	*A.lo = mirror ? -*A.hi : A.to_coord(W, *A.Min);
	*A.size = *A.hi - *A.lo;
	*A.HI = A.to_coord(W, *A.MAX);
	*A.LO = mirror ? -*A.HI : A.to_coord(W, *A.MIN);
}

/* ---- index ranges, boundary types, Bloch phases (reference world.c:192-258) ---------------------- */
static world new_world(res R, dom D, sur S)
{
	world W;
	MALLOC(W, 0, 1);

	if (R.nug.x.Min>0) R.nug.x.Max = R.nug.x.Min, R.nug.x.Min = -R.nug.x.Max;
	if (R.nug.y.Min>0) R.nug.y.Max = R.nug.y.Min, R.nug.y.Min = -R.nug.y.Max;
	if (R.nug.z.Min>0) R.nug.z.Max = R.nug.z.Min, R.nug.z.Min = -R.nug.z.Max;
	/* a periodic axis given a Bloch period becomes a Bloch axis */
	if (S.x.MinSur==PBC && S.x.MaxSur) S.x.MinSur=BBC;
	if (S.y.MinSur==PBC && S.y.MaxSur) S.y.MinSur=BBC;
	if (S.z.MinSur==PBC && S.z.MaxSur) S.z.MinSur=BBC;
	if (!S.pml.Num) S.pml.Num = 12;

	*W = (struct world) { D, R, S, 0, R.dt, R.d.x, R.d.y?R.d.y:R.d.x, R.d.z?R.d.z:R.d.x, R.spec.eV?R.spec.eV:1, INF };
	/* a single positive number, or a symmetric axis, means +-Max */
	if (!D.x.Max && D.x.Min>0 || S.x.MinSur==SYM) D.x.Min = -(D.x.Max = D.x.Max ? D.x.Max : D.x.Min);
	if (!D.y.Max && D.y.Min>0 || S.y.MinSur==SYM) D.y.Min = -(D.y.Max = D.y.Max ? D.y.Max : D.y.Min);
	if (!D.z.Max && D.z.Min>0 || S.z.MinSur==SYM) D.z.Min = -(D.z.Max = D.z.Max ? D.z.Max : D.z.Min);

	W->itox = R.gridFunction.itox ? R.gridFunction.itox : R.nug.x.m > 1 ? i_to_x_graded : i_to_x_uniform;
	W->jtoy = R.gridFunction.jtoy ? R.gridFunction.jtoy : R.nug.y.m > 1 ? j_to_y_graded : j_to_y_uniform;
	W->ktoz = R.gridFunction.ktoz ? R.gridFunction.ktoz : R.nug.z.m > 1 ? k_to_z_graded : k_to_z_uniform;
	W->xtoi = x_to_i, W->ytoj = y_to_j, W->ztok = z_to_k;

	/* unspecified surfaces inherit: y,z from x; max side from min side */
	W->xMinSur = S.x.MinSur;
	W->yMinSur = S.y.MinSur ? S.y.MinSur : S.x.MinSur;
	W->zMinSur = S.z.MinSur ? S.z.MinSur : S.x.MinSur;
	W->xMaxSur = S.x.MinSur!=BBC && S.x.MaxSur ? S.x.MaxSur : S.x.MinSur;
	W->yMaxSur = S.y.MinSur!=BBC && S.y.MaxSur ? S.y.MaxSur : S.y.MinSur ? S.y.MinSur : S.x.MinSur;
	W->zMaxSur = S.z.MinSur!=BBC && S.z.MaxSur ? S.z.MaxSur : S.z.MinSur ? S.z.MinSur : S.x.MinSur;

	const int shell = S.pml.Num+1;            /* cells per PML face */
	const axis_slots ax[3] = {
		{ &W->iMin, &W->iMax, &W->iNum, &W->iMIN, &W->iMAX, &W->iNUM, &W->xMin, &W->xMax, &W->xSize, &W->xMIN, &W->xMAX, W->xMinSur, W->xMaxSur, D.x.Min, D.x.Max, W->xtoi, W->itox },
		{ &W->jMin, &W->jMax, &W->jNum, &W->jMIN, &W->jMAX, &W->jNUM, &W->yMin, &W->yMax, &W->ySize, &W->yMIN, &W->yMAX, W->yMinSur, W->yMaxSur, D.y.Min, D.y.Max, W->ytoj, W->jtoy },
		{ &W->kMin, &W->kMax, &W->kNum, &W->kMIN, &W->kMAX, &W->kNUM, &W->zMin, &W->zMax, &W->zSize, &W->zMIN, &W->zMAX, W->zMinSur, W->zMaxSur, D.z.Min, D.z.Max, W->ztok, W->ktoz },
	};
	for (int a=0; a<3; a++) lay_out_axis(W, ax[a], shell);

	/* phase factors of the boundary copies: Bloch axes rotate by 2 pi L / period; symmetric
	 * axes start at 0 and get their parity (+-1) from the first source; the rest copy with 1 */
	const int bx = S.x.MinSur==BBC && S.x.MaxSur, by = S.y.MinSur==BBC && S.y.MaxSur, bz = S.z.MinSur==BBC && S.z.MaxSur;
	W->coskx = bx ? cosf(2*PI*(W->xMax-W->xMin)/S.x.MaxSur) : S.x.MinSur==SYM ? 0 : 1;
	W->cosky = by ? cosf(2*PI*(W->yMax-W->yMin)/S.y.MaxSur) : S.y.MinSur==SYM ? 0 : 1;
	W->coskz = bz ? cosf(2*PI*(W->zMax-W->zMin)/S.z.MaxSur) : S.z.MinSur==SYM ? 0 : 1;
	W->sinkx = bx ? sinf(2*PI*(W->xMax-W->xMin)/S.x.MaxSur) : 0;
	W->sinky = by ? sinf(2*PI*(W->yMax-W->yMin)/S.y.MaxSur) : 0;
	W->sinkz = bz ? sinf(2*PI*(W->zMax-W->zMin)/S.z.MaxSur) : 0;

	/* hexagonal Bloch boundary in y: two half-period shifted phases */
	const int hy = S.y.MinSur==HBC && S.y.MaxSur;
	W->coskyp = hy ? cosf(PI*(2*(W->yMax-W->yMin)/S.y.MaxSur+(W->xMax-W->xMin)/S.x.MaxSur)) : 1;
	W->coskym = hy ? cosf(PI*(2*(W->yMax-W->yMin)/S.y.MaxSur-(W->xMax-W->xMin)/S.x.MaxSur)) : 1;
	W->sinkyp = hy ? sinf(PI*(2*(W->yMax-W->yMin)/S.y.MaxSur+(W->xMax-W->xMin)/S.x.MaxSur)) : 0;
	W->sinkym = hy ? sinf(PI*(2*(W->yMax-W->yMin)/S.y.MaxSur-(W->xMax-W->xMin)/S.x.MaxSur)) : 0;

	W->xyArea = W->xSize * W->ySize, W->xzArea = W->xSize * W->zSize, W->yzArea = W->ySize * W->zSize;
	if (S.x.MinSur==BBC || S.y.MinSur==BBC || S.z.MinSur==BBC) W->complexField = 1;
	return W;
}

/* reference world.c:261-280 */
world createWorld(dom D, res R, sur S, char *format, ...)
{
	ALIS_PROF("createWorld");
	va_list ap;
	world W = new_world(R, D, S);

	W->E = make_field_set(W);
	W->H = make_field_set(W);
	W->CE = make_set_coeffs(W, S, 0);
	W->CH = make_set_coeffs(W, S, 1);

	va_start(ap, format);
	vsprintf(W->ID, format, ap);
	va_end(ap);
	if (W->ID[0]) {
		/* like the reference: start from a clean output directory named after the driver */
		printf("Deleting previous files.");
		exec("rm -fr %s %s.c?* %s[#--0-~]* %s.[#-bd-~]* 2> /dev/null", W->ID, W->ID, W->ID, W->ID);
	}
	alis_host_attach(W);
	return W;
}

/* a world sharing the coefficients (grid, materials) of SW with its own, zeroed, field state
 * (reference world.c:318-331) */
world cloneWorld(world SW)
{
	world W;
	MALLOC(W, 0, 1);
	*W = *SW;
	W->E = make_field_set(W);
	W->H = make_field_set(W);
	W->b200_bridge = 0, W->b200_view = 0;
	W->b200_shared_coeffs = 1;
	W->SE = W->SH = 0, W->S = 0;          /* sources stay with the source world */
	alis_host_attach(W);
	return W;
}

/* reference world.c:371-380: zero fields, psi, J/K; drop sources; H->t = 0.5 (sic) */
void rebootWorld(world W)
{
	removeSources(W);
	alis_host_reset_fields(W);
	W->t = W->E->t = 0;
	W->H->t = 0.5;
	W->n = 0;
	alis_host_invalidate(W);
	printf("\n");
}

static void free_set_coeffs(world W, coeffs *C)
{
	if (C->xP) free(&C->xP[W->iMIN]), free(&C->xPy[W->iMIN]), free(&C->xPz[W->iMIN]);
	if (C->yP) free(&C->yP[W->jMIN]), free(&C->yPz[W->jMIN]), free(&C->yPx[W->jMIN]);
	if (C->zP) free(&C->zP[W->kMIN]), free(&C->zPx[W->kMIN]), free(&C->zPy[W->kMIN]);
	free(&C->dtdx[W->iMIN]), free(&C->dx[W->xMinSur!=SYM ? W->iMIN : -W->iMAX]);
	free(&C->dtdy[W->jMIN]), free(&C->dy[W->yMinSur!=SYM ? W->jMIN : -W->jMAX]);
	free(&C->dtdz[W->kMIN]), free(&C->dz[W->zMinSur!=SYM ? W->kMIN : -W->kMAX]);
	free(C);
}

/* reference world.c:445-466 */
void deleteWorld(world W)
{
	removeSources(W);
	alis_host_detach(W);
	if (!W->b200_shared_coeffs) {
		removeObjects(W);
		free_set_coeffs(W, W->CE);
		free_set_coeffs(W, W->CH);
	}
	free(W->E);
	free(W->H);
	free(W);
}

void deleteCloneWorld(world W)
{
	deleteWorld(W);
}

/* HDF5 is not available in this image: the I/O helpers (reference world.c:469-537) report it */
/* reference world.c:469-517: the first dataset of <name>.h5 as float*, float** or float*** (row
 * pointers into one block), dimensions into dim[].  libhdf5 is replaced by h5lite.c. */
void *h5read(int *dim, char *format, ...)
{
	va_list ap;
	char name[1024], err[200];
	unsigned long long dims[3];
	int N = 0;
	va_start(ap, format);
	vsprintf(name, format, ap);
	va_end(ap);
	float *data = h5lite_read(strcat(name, ".h5"), &N, dims, err, sizeof err);
	if (!data) { printf("alis-b200: h5read(\"%s\"): %s\n", name, err); exit(0); }
	if (dim) for (int n=0; n<N; n++) dim[n] = dims[n];
	if (N == 1) return data;
	if (N == 2) {
		float **f = xcalloc(dims[0], sizeof(float*));
		for (unsigned long long i=0; i<dims[0]; i++) f[i] = data + dims[1]*i;
		return f;
	}
	float ***f = xcalloc(dims[0], sizeof(float**));
	f[0] = xcalloc(dims[0]*dims[1], sizeof(float*));
	for (unsigned long long i=1; i<dims[0]; i++) f[i] = f[0] + dims[1]*i;
	for (unsigned long long i=0; i<dims[0]; i++) for (unsigned long long j=0; j<dims[1]; j++) f[i][j] = data + dims[2]*(dims[1]*i + j);
	return f;
}

/* reference world.c:520-537: dataset "data" of the non-zero dimensions in <name>.h5 */
void h5write(int *dim, float *p, char *format, ...)
{
	va_list ap;
	char filename[1024];
	int d = 0;
	unsigned long long dims[3] = {0};
	va_start(ap, format);
	vsprintf(filename, format, ap);
	va_end(ap);
	for (int n=0; n<3; n++) if (dim[n]) dims[d++] = dim[n];
	if (h5lite_write(strcat(filename, ".h5"), d, dims, p)) { printf("alis-b200: h5write: cannot write %s\n", filename); exit(0); }
}
