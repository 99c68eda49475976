/*
 * alis_b200.h -- C ABI of the B200 FDTD time-stepping engine.
 *
 * This is the boundary between host C code (the ALiS-compatible public API in
 * include/alis.h, implemented in alis_b200/host/) and the hand-written sm_100a
 * CUDA kernels in alis_b200/csrc/.  Plain pointers, sizes and POD structs only.
 *
 * What each entry point replaces in the reference (hseelab/alis, src/):
 *   alisb_create / alisb_set_axis_tables   fields+psi allocation and 1-D tables
 *                                          world.c:164-189 (makeVectorField),
 *                                          world.c:98-161 (makeCpmlCoeff, makeVFieldCoeff)
 *   alisb_set_inv_eps                      CE->x/y/z  object.c:168-174,374-398
 *   alisb_add_dispersive                   ADE boxes  object.c:178-226,290-336
 *   alisb_add_source                       source lists W->SE/W->SH  source.c:52-75, update.c:161-179
 *   alisb_add_tfsf                         TFSF list W->S  source.c:269-334, update.c:194-272
 *   alisb_set_copies                       copyBoundary call lists  update.c:182-191,284-297,311-339
 *   alisb_step_E / alisb_step_H            updateE / updateH  update.c:275-341
 *   alisb_phaser_*                         createPhasers/updatePhaser  phaser.c:21-110
 *   alisb_phaser_surface_sums              farFieldX/Y/Z double loops  farfld.c:248-459
 *   alisb_geometry_bounds / _fill          getObjectsBounds, fillFactor + the built-in shapes
 *                                          object.c:70-121,229-279, shapes.c:17-221
 *   alisb_read_box / alisb_write_box       the host float*** views  world.c:58-70
 *   alisb_eval_box / _reduce_absmax / _column_sums   field functions + slices/world reductions
 *                                          fields.c:130-681, slices.c:242-427, output.c:56-101
 *
 * Coordinates.  The engine uses ARRAY coordinates: a = i-(iMIN-1), b = j-(jMIN-1),
 * c = k-(kMIN-1), so a runs 0..nx+1 where 1..nx are the updated cells (iMIN..iMAX) and
 * 0 / nx+1 are the reference's ghost layer (world.c:170-172).  All boxes are inclusive
 * [lo,hi] in array coordinates.  Host-side dense blocks are packed like the reference's
 * makeField blocks: c (z) fastest, then b, then a  (world.c:58-70).
 * On the device the layout is x fastest / z slowest.
 *
 * Errors: every function returns 0 on success, a negative ALISB_E* code otherwise;
 * alisb_last_error() returns a static description.  There is NO CPU fallback: if no
 * CUDA device is usable alisb_create fails.
 */
#ifndef ALIS_B200_H
#define ALIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALISB_ABI_VERSION 1

enum { ALISB_OK = 0, ALISB_EINVAL = -1, ALISB_ENOMEM = -2, ALISB_ECUDA = -3, ALISB_ESTATE = -4 };

enum { ALISB_SET_E = 0, ALISB_SET_H = 1 };
enum { ALISB_X = 0, ALISB_Y = 1, ALISB_Z = 2 };
enum { ALISB_LO = 0, ALISB_HI = 1 };

/* array identifiers for alisb_read_box / alisb_write_box */
enum {
	ALISB_ARR_FIELD   = 0,  /* set, comp, m                                       */
	ALISB_ARR_INVEPS  = 1,  /* comp                      (E only)                 */
	ALISB_ARR_PSI     = 2,  /* set, m, axis, side, which (0: comp axis+1, 1: axis+2) */
	ALISB_ARR_ADE_J   = 3,  /* m, object, pole, comp                              */
	ALISB_ARR_ADE_K   = 4,  /* m, object, pole, comp                              */
	ALISB_ARR_ADE_FILL= 5   /* object, comp                                       */
};

typedef struct alisb_engine alisb_engine;
typedef struct alisb_phaser alisb_phaser;

/* Grid description (reference struct world, alis_c.h:109-130, index part). */
typedef struct {
	int32_t nx, ny, nz;          /* iNUM, jNUM, kNUM: updated cells per axis                  */
	int32_t complex_field;       /* W->complexField: 1 => a second (imaginary) field set      */
	float   dt;                  /* W->dt                                                     */
	/* CPML/ADE-PML slabs, update.c:119-124.  pml[set][axis][side] = {first,last} cell in
	 * array coordinates (inclusive), or first>last for "no slab".                        */
	int32_t pml[2][3][2][2];
	/* z-slab sharding: this engine instance owns global planes
	 * c = kshard_lo .. kshard_lo+nz-1 of a global grid of nz_global planes.  For a
	 * single-GPU world kshard_lo = 1 and nz_global = nz.  pml[][Z] and all boxes are
	 * LOCAL to the shard.                                                               */
	int32_t kshard_lo, nz_global;
	int32_t device;              /* CUDA device ordinal                                       */
	int32_t reserved[5];
} alisb_grid;

/* Dispersive object: per-object ADE data, reference coeffs (alis_c.h:96-107),
 * object.c:178-226 (allocation), :290-336 (coefficients), update.c:89-108,128-158. */
#define ALISB_MAXPOLES 6
typedef struct {
	int32_t box[6];              /* a0,a1,b0,b1,c0,c1 inclusive (bbox clipped to updated cells +1) */
	int32_t nj, nk;              /* NJ (Drude poles), NK (all poles)                           */
	float   a[3];                /* ax, ay, az loss factors                                    */
	float   JJ[3][ALISB_MAXPOLES], JK[3][ALISB_MAXPOLES], JE[3][ALISB_MAXPOLES];
	float   KJ[3][ALISB_MAXPOLES], KK[3][ALISB_MAXPOLES], KE[3][ALISB_MAXPOLES];
	const float *fill[3];        /* CE->Jx/Jy/Jz[n]: dense host blocks over box, z fastest      */
} alisb_dispersive;

/* Waveform callback evaluated ON THE HOST with the caller's libm (so that the value
 * is bit-identical to the reference's S->waveForm(W,S,t,phase), update.c:166-171). */
typedef float (*alisb_waveform_fn)(void *user, float t, float phase);

/* Soft source (reference struct source, alis_c.h:305-312; update.c:161-179). */
typedef struct {
	int32_t set, comp, m;        /* target array: E or H, component, field set (0 real, 1 imag) */
	int32_t box[6];
	float   amplitude, phase;    /* scalars used when the arrays below are NULL                 */
	const float *Amplitude;      /* dense host block over box or NULL                           */
	const float *Phase;          /* dense host block over box or NULL                           */
	uint64_t n_active;           /* S->N: source adds while engine step counter n < n_active    */
	alisb_waveform_fn waveform;
	void   *user;
} alisb_source;

/* One TFSF face strip: F[a,b,c] += coef * G_aux[clamp(a+alt)...] (* inv_eps) -- update.c:194-202
 * with the call lists of update.c:205-272.  The aux array is indexed through the
 * reference's pointer aliasing (source.c:223-248): aux index = min(index, clamp). */
typedef struct {
	int32_t comp;                /* target component of the pass's field set                    */
	int32_t aux_comp;            /* component of the aux world's OTHER field set                */
	float   coef;                /* +-dtd?[face]                                                */
	int32_t box[6];              /* in MAIN array coordinates                                   */
	int32_t alt[3];              /* iAlt,jAlt,kAlt added before the aux lookup                  */
	int32_t use_eps;             /* multiply by inv_eps[comp] (E pass with CE->x allocated)     */
} alisb_strip;

typedef struct {
	alisb_engine *aux;           /* auxiliary world engine; stepped first (update.c:207,242)    */
	int32_t offset[3];           /* aux array coord = main array coord + offset                 */
	int32_t clamp[2][3][2];      /* [aux set][aux comp][0:x,1:y] max aux array coord (aliasing) */
	const alisb_strip *strips_E; int32_t n_strips_E;   /* applied to every field set m         */
	const alisb_strip *strips_H; int32_t n_strips_H;
} alisb_tfsf;

/* One boundary copy (reference copyBoundary, update.c:182-191).  Executed in list order.
 * real:    F[p] = cosk*F[p+alt]
 * complex: F0[p] = cosk*F0[p+alt] + sink*F1[p+alt];  F1[p] = cosk*F1[p+alt] - sink*F0[p+alt] */
typedef struct {
	int32_t set, comp;
	float   cosk, sink;
	int32_t box[6];
	int32_t alt[3];
} alisb_copy;

/* ---- life cycle ------------------------------------------------------------------ */
int  alisb_abi_version(void);
const char *alisb_last_error(void);
const char *alisb_backend(void);                 /* "cuda-sm_100a" or "oracle-cpu"            */
int  alisb_create(const alisb_grid *grid, alisb_engine **out);
void alisb_destroy(alisb_engine *e);
int  alisb_set_stream(alisb_engine *e, void *cuda_stream);   /* NULL = engine-owned stream    */

/* ---- coefficients ---------------------------------------------------------------- */
/* dtd, P, P1, P2: host arrays of n+2 floats in array coordinates (n = nx/ny/nz of the axis).
 * P/P1/P2 = C->xP,xPy,xPz | yP,yPz,yPx | zP,zPx,zPy (world.c:155-157); NULL if no PML.   */
int  alisb_set_axis_tables(alisb_engine *e, int set, int axis, const float *dtd,
                           const float *P, const float *P1, const float *P2);
/* dense host blocks exactly like the reference's allocations (object.c:168-170):
 *   cx covers a in [1,nx+1], b in [1,ny],   c in [1,nz]
 *   cy covers a in [1,nx],   b in [1,ny+1], c in [1,nz]
 *   cz covers a in [1,nx],   b in [1,ny],   c in [1,nz+1]      (z fastest)
 * NULL for all three = vacuum (CE->x == NULL, update.c:109,114).                          */
int  alisb_set_inv_eps(alisb_engine *e, const float *cx, const float *cy, const float *cz);
int  alisb_add_dispersive(alisb_engine *e, const alisb_dispersive *obj);
int  alisb_clear_dispersive(alisb_engine *e);    /* removeObjects: ADE boxes and their J/K     */
int  alisb_add_source(alisb_engine *e, const alisb_source *src);
int  alisb_add_tfsf(alisb_engine *e, const alisb_tfsf *t);
int  alisb_set_copies(alisb_engine *e, int set, const alisb_copy *ops, int n);
int  alisb_clear_sources(alisb_engine *e);       /* removeSources: sources + tfsf links        */

/* ---- time stepping --------------------------------------------------------------- */
int  alisb_step_E(alisb_engine *e);              /* updateE, update.c:275-299                 */
int  alisb_step_H(alisb_engine *e);              /* updateH, update.c:302-341                 */
int  alisb_run(alisb_engine *e, int steps);      /* steps x (E then H), no host sync between  */
/* A pass split in two for communication/computation overlap (multi-GPU):
 *   ALISB_PART_SEND_PLANE  everything of the pass that lands in the plane the neighbour needs
 *                          (E: c = nz, H: c = 1): volume update, ADE, TFSF, sources, x/y copies
 *   ALISB_PART_REST        the rest of the pass (and the time bookkeeping)
 * SEND_PLANE followed by REST is identical to alisb_step_E / alisb_step_H (= ALISB_PART_ALL) on
 * the side of a shard that faces a neighbour, i.e. where the plane beyond the send plane is a halo
 * and not the target/source of a z boundary copy (rank r splits its E pass if it has an upper
 * neighbour and its H pass if it has a lower one; alis_b200/sharding.py).                      */
enum { ALISB_PART_ALL = 0, ALISB_PART_SEND_PLANE = 1, ALISB_PART_REST = 2 };
int  alisb_step_part(alisb_engine *e, int set, int part);
/* Fused E+H stepping (temporal blocking of update.c:275-341).  Where a step is nothing but the two
 * curls -- outside CPML slabs, material arrays, ADE boxes, sources, TFSF strips -- both half steps of
 * a cell are formed in one pass over memory (48 B per cell-step instead of 72), bit-identical to the
 * two-pass path.  With fusion on (alisb_set_fusion, or ALISB_FUSE=1 in the environment) an E pass is
 * prepared when alisb_step_E / alisb_step_part(E) is called but issued only when the H pass, or any
 * other call that looks at the device state, arrives; fields are double buffered (2x field memory).
 *   alisb_flush(e, 1)       issue the deferred E pass now as the fused variant; the caller promises
 *                           that the next thing to touch this engine's fields is the H pass
 *   alisb_flush(e, 0)       issue it as the plain two-pass variant
 *   alisb_begin_pass(e,set) swap the buffers for the coming pass now, so that alisb_halo_info returns
 *                           the planes that pass will write (alis_b200/sharding.py posts its receives
 *                           before the pass is issued)                                            */
int  alisb_set_fusion(alisb_engine *e, int on);
int  alisb_flush(alisb_engine *e, int h_pass_follows);
int  alisb_begin_pass(alisb_engine *e, int set);
uint64_t alisb_fused_launch_count(alisb_engine *e);   /* launches of the fused kernel so far (tests, bench) */
int  alisb_synchronize(alisb_engine *e);
int  alisb_get_time(alisb_engine *e, uint64_t *n, float *t, float *tE, float *tH);
int  alisb_set_time(alisb_engine *e, uint64_t n, float t, float tE, float tH);
int  alisb_reset_fields(alisb_engine *e);        /* rebootWorld field part, world.c:334-368   */

/* ---- second-order (chi2) mixing between worlds -------------------------------------- */
/* coupleTWM (update.c:356-391) for ADE object `obj` (index in add order) and field set m of the
 * signal engine e:  E(e).v += 2*alt*d[v][q]*inv_eps.v*fill.v*(E(p1)..*E(p2)..), nine ordered sweeps
 * per component with the reference's double-precision averages.  d is the 3x6 contracted tensor,
 * row-major (matter.d, alis_c.h:155).  nonLinearSHG(WS,WP,Alt) = couple(WS, WP, WP); nonLinearTWM
 * (WS,WP1,WP2,Alt) = couple(WS, WP1, WP2) (update.c:394-411).  All three engines must have the same
 * grid and device; p1/p2 may be e itself.                                                   */
int  alisb_chi2_couple(alisb_engine *e, int obj, int m, alisb_engine *p1, alisb_engine *p2, const float *d, float alt);

/* ---- data access (parity dumps, slices, get) -------------------------------------- */
/* ids: meaning depends on arr (see enum).  box inclusive, host block dense z-fastest.      */
int  alisb_read_box(alisb_engine *e, int arr, const int32_t ids[6], const int32_t box[6], float *host);
int  alisb_write_box(alisb_engine *e, int arr, const int32_t ids[6], const int32_t box[6], const float *host);

/* ---- device-side monitor reductions ------------------------------------------ */
/* cell sizes C->dx/dy/dz of field set `set` (reference world.c:148-150) in array coordinates,
 * n+2 floats; needed by the flux integrals                                                   */
int  alisb_set_spacing(alisb_engine *e, int set, int axis, const float *d);
/* Time-domain Poynting flux through axis-aligned rectangles (reference output.c:233-332,
 * poyntingX/Y/Z): plane normal `axis` at array coordinate `at`, tangential index ranges
 * [b0,b1] x [c0,c1] with b = axis+1, c = axis+2 (cyclic).  out[2*q] = + sum of E_b (H_c + H_c')
 * with trapezoidal weights, out[2*q+1] = - sum of E_c (H_b + H_b'), exactly the two partial sums of
 * the reference's two loops, accumulated in double on the device.  One synchronisation for all
 * n planes (a Poynting box is one call).                                                      */
typedef struct { int32_t axis, at, b0, b1, c0, c1; } alisb_flux_plane;
int  alisb_plane_flux(alisb_engine *e, const alisb_flux_plane *planes, int n, double *out);

/* Point-wise field functions evaluated on the device (reference fields.c:130-681, those that read
 * E and H only), for slices and world reductions (slices.c:242-427, output.c:56-101).            */
enum {
	ALISB_FN_RAW   = 0,   /* rawEx..rawHz: set, comp                       fields.c:130-176      */
	ALISB_FN_NODE  = 1,   /* Ex..Hz, iEx..iHz at the node: set, comp, m     fields.c:179-291      */
	ALISB_FN_EE    = 2,   /* |E|^2                                         fields.c:453-461      */
	ALISB_FN_HH    = 3,   /* |H|^2                                         fields.c:476-484      */
	ALISB_FN_UH    = 4,   /* HH/2                                          fields.c:486          */
	ALISB_FN_EXH   = 5,   /* E_comp * H_comp2 (ExHy ...)                   fields.c:381-439      */
	ALISB_FN_S     = 6,   /* Poynting vector component comp                fields.c:483-485      */
	ALISB_FN_TDIAG = 7,   /* Maxwell stress Txx, Tyy, Tzz: comp            fields.c:564-611      */
	ALISB_FN_TOFF  = 8,   /* off-diagonal stress, factors comp, comp2      fields.c:613-681      */
	ALISB_FN_DIVE  = 9    /* div E                                         fields.c:552-562      */
};
typedef struct {
	int32_t fn, set, comp, comp2, m;
	int32_t sym[3];       /* the axis has a mirror plane at index 0 (W->xMinSur==SYM ...): negative
	                       * indices fold to their mirror image with the component's parity       */
	float   cosk[3];      /* W->coskx, W->cosky, W->coskz                                          */
	int32_t org[3];       /* array coordinate of world index 0: a = i + org[0] (= i-(iMIN-1)) ...  */
} alisb_field_fn;
/* ibox: inclusive box in world INDEX coordinates (may be negative on mirror axes).  After folding,
 * every point and (except for ALISB_FN_RAW) its +1 neighbours must lie inside the arrays.
 * alisb_eval_box: values into a dense host block, z fastest.  alisb_reduce_absmax: max |F| (NaNs
 * skipped, like the reference's comparisons).  alisb_column_sums: for every (i,j) the z column sum
 * of worldSum/sliceSum -- interior points weighted dz[k], then the two ends weighted dz/2, in the
 * reference's order and promotions (output.c:73-80) -- into host[(i-i0)*nj + (j-j0)]; needs
 * alisb_set_spacing(e, ALISB_SET_E, ALISB_Z, ...) and a k range inside the arrays.              */
int  alisb_eval_box(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], float *host);
/* Running sums of a field function over a box, kept where the fields live (reference slices.c:282-379:
 * the per-pixel accumulations of sliceTimeAvg `Q->F += F` and sliceFreqDom `Q->F += F*real, Q->G += F*imag`,
 * called every step).  alisb_accum_add queues  re += F*wr  [, im += F*wi]  for every point of the box and
 * returns without waiting for the device; alisb_accum_read brings the sums back (dense blocks, z fastest).
 * Float accumulation per point in call order: bit-identical to the reference's loops. */
typedef struct alisb_accum alisb_accum;
int  alisb_accum_create(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], int with_imag, alisb_accum **out);
int  alisb_accum_add(alisb_accum *a, float wr, float wi);
int  alisb_accum_read(alisb_accum *a, float *re, float *im);
int  alisb_accum_clear(alisb_accum *a);
void alisb_accum_destroy(alisb_accum *a);
int  alisb_reduce_absmax(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], float *out);
int  alisb_column_sums(alisb_engine *e, const alisb_field_fn *f, const int32_t ibox[6], float *host);
/* objectAbsorption (reference output.c:23-47,365-435): for ADE object `obj` (index in add order) and
 * every column (a,b) of box[0..3] (ARRAY coordinates, inside the object's box) the three z sums
 * xs, ys, zs of E.J over the planes box[4]..box[5], including the half-weight corrections at the
 * world's first / last plane (array coordinate 1 / nz): host[((a-a0)*nb + (b-b0))*3 + {0,1,2}].
 * Needs alisb_set_inv_eps and the z spacing tables of both field sets.  The host applies the
 * reference's edge weights in x,y and adds the columns in its order.                          */
int  alisb_absorption_columns(alisb_engine *e, int obj, const int32_t box[6], float *host);

/* ---- geometry rasterisation on the device (reference shapes.c:17-221,
 * object.c:70-121,229-279) ---------------------------------------------------------------------
 * The host flattens the object list into alisb_shape nodes (children of a node are contiguous);
 * whatever needs libm is evaluated by the host and passed in: aux holds cos/sin constants
 * (ROTATION: cx,sx,cy,sy,cz,sz at node.aux; CLATTICE*: count (cos,sin) pairs at node.aux), coord[axis]
 * the coordinates itox/jtoy/ktoz of every sample position, [(idx-lo)*2 + half][slot], subn+3 slots:
 * 0 = the position itself, 1/2 = minus/plus the outermost sub-sample offset (object.c:104),
 * 3+q = the q-th regular sub-sample (object.c:116-118); half = 1 means idx - 1/2.              */
enum {
	ALISB_SHAPE_BOX = 0, ALISB_SHAPE_BALL, ALISB_SHAPE_DIAMOND, ALISB_SHAPE_BICONEX, ALISB_SHAPE_BICONEY, ALISB_SHAPE_BICONEZ,
	ALISB_SHAPE_RODX, ALISB_SHAPE_RODY, ALISB_SHAPE_RODZ, ALISB_SHAPE_DRODX, ALISB_SHAPE_DRODY, ALISB_SHAPE_DRODZ,
	ALISB_SHAPE_HRODXY, ALISB_SHAPE_HRODYZ, ALISB_SHAPE_HRODZX, ALISB_SHAPE_HRODXZ, ALISB_SHAPE_HRODYX, ALISB_SHAPE_HRODZY,
	ALISB_SHAPE_TRANSLATION, ALISB_SHAPE_ROTATION, ALISB_SHAPE_COMBINATION, ALISB_SHAPE_INTERSECTION, ALISB_SHAPE_DIFFERENCE,
	ALISB_SHAPE_LATTICE, ALISB_SHAPE_DLATTICE, ALISB_SHAPE_CLATTICEX, ALISB_SHAPE_CLATTICEY, ALISB_SHAPE_CLATTICEZ,
	ALISB_SHAPE_COUNT
};
typedef struct { int32_t type, child, nchild, aux; float f[3][3]; } alisb_shape;
typedef struct {
	int32_t nobjects; const int32_t *roots;      /* root node of every object, list order (earlier wins) */
	int32_t nnodes;   const alisb_shape *nodes;
	int32_t naux;     const float *aux;
	int32_t subn;                                /* W->subN, >= 1                                         */
	int32_t lo[3], hi[3];                        /* index range the coordinate tables cover (inclusive)   */
	const float *coord[3];
} alisb_geometry;
/* getObjectsBounds (object.c:229-262): for every object the index bounding box of the nodes of
 * ibox (inclusive, world indices) that lie in it and in no earlier object: bounds[o*6 + ...] =
 * iMin,iMax,jMin,jMax,kMin,kMax; an object that owns no node gets iMin > iMax.                  */
int  alisb_geometry_bounds(int device, const alisb_geometry *g, const int32_t ibox[6], int32_t *bounds);
/* fillFactor (object.c:78-121) of object `object` for the three staggered components at every cell
 * of ibox: fx at (i-1/2,j,k), fy at (i,j-1/2,k), fz at (i,j,k-1/2); dense host blocks, z fastest. */
int  alisb_geometry_fill(int device, const alisb_geometry *g, int object, const int32_t ibox[6], float *fx, float *fy, float *fz);

/* ---- z-slab halo (multi-GPU) -------------------------------------------- */
/* Buffers (device memory for the CUDA engine) holding one full z-plane incl. ghosts of
 * component comp (x or y) of field set `set`, copy m:
 *   send : the owned plane a neighbour needs (E: top plane c=nz -> upper neighbour's c=0;
 *          H: bottom plane c=1 -> lower neighbour's c=nz+1)     [update.c:110-111]
 *   recv : the ghost plane filled by the neighbour (E: c=0, H: c=nz+1)
 * The CUDA engine hands out pointers straight into the field arrays (planes are
 * contiguous in its x-fastest/z-slowest layout), so pack/unpack are no-ops there.          */
int  alisb_halo_info(alisb_engine *e, int set, int m, int comp, void **send_ptr, void **recv_ptr, size_t *bytes);
int  alisb_halo_pack(alisb_engine *e, int set);      /* owned planes -> send buffers (on stream) */
int  alisb_halo_unpack(alisb_engine *e, int set);    /* recv buffers -> ghost planes (on stream) */
/* Both engines in ONE process (alis_b200/host/shard.c: a world cut into z slabs behind the public C API): the
 * plane of field set `set` that crosses the cut between `lower` (the slab below) and `upper` moves from the
 * engine that owns it into the other's ghost plane -- set E: E.x, E.y of lower's plane nz -> upper's plane 0;
 * set H: H.x, H.y of upper's plane 1 -> lower's plane nz+1; every complex field set.  CUDA engine: a peer copy
 * (cudaMemcpyPeerAsync, NVLink between the GPUs of one box) ordered after the owner's stream, and the receiver's
 * stream waits for it; nothing blocks the host.  The engines must agree in nx, ny and complex_field.      */
int  alisb_halo_exchange(alisb_engine *lower, alisb_engine *upper, int set);

/* ---- running DFT (phasers) ---------------------------------------------------------- */
/* box = interior-clipped monitor box; face_mask bit (axis*2+side) set => that face is
 * accumulated (phaser.c:57-62).  nfreq accumulator sets.                                  */
int  alisb_phaser_create(alisb_engine *e, const int32_t box[6], int face_mask, int nfreq, alisb_phaser **out);
void alisb_phaser_destroy(alisb_phaser *p);
/* expE/expH: nfreq (re,im) pairs = cexpf(2*pi*i*f*tE), cexpf(2*pi*i*f*tH) from the host.  */
int  alisb_phaser_accumulate(alisb_phaser *p, const float *expE, const float *expH);
/* which: 0,1 = E1,E2 ; 2,3 = H1,H2 of that face (phaser.c:57-62 ordering); host gets
 * (re,im) pairs, dense over the face, z fastest.                                           */
int  alisb_phaser_read(alisb_phaser *p, int freq, int face, int which, float *host_complex);

/* Near-to-far-field transform, the surface sums (reference farfld.c:248-459, the
 * double loops of farFieldX/Y/Z).  One job = one separable sum over one accumulator of one face,
 *     out[job] = sum_v sum_u  w_lo[u] * A[u][v] * w_hi[v],
 * u running over the face box along its LOWER-numbered in-face axis and v along the higher one
 * (x face: u=y, v=z; y face: u=x, v=z; z face: u=x, v=y).  weights holds nweights complex floats
 * (re,im); a job names its two vectors by their offsets (in complex elements).  The host folds the
 * reference's trapezoid end weights, staggered ranges, cell sizes, plane-wave factors and test-field
 * samples into the vectors.  All directions and polarisations of an angular sweep are the jobs of
 * ONE call (one launch, one synchronisation); the phasor faces never leave the device.
 * out: (re,im) doubles per job.  Products in fp32 (A*w_hi, then *w_lo), sums in double.          */
typedef struct { int32_t face, which; uint32_t w_lo, w_hi; } alisb_surface_job;
int  alisb_phaser_surface_sums(alisb_phaser *p, int freq, const alisb_surface_job *jobs, int njobs,
                               const float *weights, size_t nweights, double *out);

/* ---- instrumentation ---------------------------------------------------------------- */
uint64_t alisb_launch_count(alisb_engine *e);    /* kernels launched by this engine so far   */
/* With profiling on, every volume-kernel launch is bracketed by CUDA events on the engine's
 * stream; alisb_last_run_kernel_ms returns their summed duration (ms) and count for the
 * last alisb_run (it synchronises the stream).                                             */
int  alisb_set_profiling(alisb_engine *e, int on);
int  alisb_last_run_kernel_ms(alisb_engine *e, float *ms_total, int *launches);
/* the same for the dominant kernel only (the volume launch covering the most cells), with
 * its algorithmic bytes per launch (36 B per cell per half-step pass)       */
int  alisb_profile_dominant(alisb_engine *e, float *ms_total, int *launches, double *bytes_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* ALIS_B200_H */
