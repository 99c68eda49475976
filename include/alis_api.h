/*
 * alis_api.h -- public C API of alis-b200, source-compatible with the API that the
 * hseelab/alis drivers (the .c files under the reference examples directory) are written against (reference src/alis_c.h).
 *
 * A simulation is a C program: build a `world`, put `object`s in it, add sources, create
 * `slice` / `phaser` monitors and run your own loop of updateE(W); updateH(W);  Here the
 * fields live in B200 HBM and updateE/updateH launch hand-written sm_100a kernels through the
 * engine C ABI (include/alis_b200.h); everything else in this header is host C.
 *
 * Compatibility notes
 *  - Aggregate types that drivers initialise positionally (dom, res, sur, matter, object) keep
 *    the reference's member order.  Handle types (world, source, tfsf, slice, phaser) keep every
 *    member NAME drivers read (W->T, W->N, W->dt, W->xMIN, P->N, P->f, ...); extra members used by
 *    this implementation are appended at the end and prefixed b200_.
 *  - Function identity is part of the API (a `field` or `waveform` pointer is compared against the
 *    library's own functions), so all of them are real functions, never macros.
 *  - The trailing-default macros (putObjects, pointDipole, sliceSnap, ...) behave as in the
 *    reference: they append sentinel / zero arguments to a variadic function.
 *  - vfield.x/y/z are HOST VIEWS of device arrays: they are valid after alisSyncFields(W) (the
 *    library's own monitors evaluate on the device or sync the region they need on demand).
 *  - The phasor arrays P->zMax.Ex[i][j][k] ... may be indexed directly, as in the reference: they
 *    are page-protected while the device holds newer data and refreshed on first touch
 *    (ALIS_PHASER_GUARD=0 turns that off; then call alisSyncPhaser(P) first).
 *  - h5read / h5write / the h5 slice writer / guidedWaveX,Y,Z work without libhdf5 (one contiguous
 *    float dataset per file, classic HDF5 layout).
 *  - Where the reference's behaviour is undefined this library is defined: get() of a point outside
 *    the domain on a non-mirror axis is 0; a clone made by cloneWorld starts with an empty source
 *    list (the reference shares the list, whose nodes keep writing into the original's arrays);
 *    deleteCloneWorld and deletePhasers free everything exactly once.
 */
#ifndef ALIS_API_H
#define ALIS_API_H

#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <math.h>
#include <values.h>
#include <complex.h>
#include <ieee754.h>

#ifdef __cplusplus
#error "alis_api.h uses C99 complex numbers and compound literals: compile drivers as C"
#endif

#define ALIS_B200_HOST 1

/* ---- surface codes (sur.?.MinSur / MaxSur) and small helpers ---------------------------- */
enum { ALIS_SUR_SYM = 1, ALIS_SUR_PEC = 2, ALIS_SUR_PML = 3, ALIS_SUR_PBC = 4, ALIS_SUR_HBC = 5, ALIS_SUR_BBC = 6, ALIS_SUR_NIL = 7 };
#define SYM 1
#define PEC 2
#define PML 3
#define PBC 4
#define HBC 5
#define BBC 6
#define NIL 7
#define MAXPOLES 6

#define INF FLT_MAX
#define PI  3.14159265358979323846f

#define sq(x)    ((x)*(x))
#define MIN(x,y) ((x)<(y)?(x):(y))
#define MAX(x,y) ((x)>(y)?(x):(y))
#define n(n)          ((matter){(n)*(n)})
#define nk(n,k,f)     ((matter){{(n)*(n)-(k)*(k)},{0,2*n*k*((float)f)}})
#define objects       (object[])
#define png(cmap,max) png,cmap,(double)max

/* offset-base allocation used all over driver code and the library (index m .. m+n-1) */
#define MALLOC(p,m,n) if ((p)=calloc((n),sizeof(*(p)))) (p)-=(m); else printf("Insufficient memory!\n"), exit(1)
/* an integer smuggled through a double vararg as a NaN payload: list terminator of writeRow etc. */
#define INT2NAN(i) ((union ieee754_double){.ieee.exponent=-1,.ieee.mantissa1=i}.d)
#define NAN2INT(n) ((union ieee754_double){.ieee.exponent=-1}.ieee.exponent==(union ieee754_double){.d=n}.ieee.exponent&&!(union ieee754_double){.d=n}.ieee.mantissa0?(union ieee754_double){.d=n}.ieee.mantissa1:0)

/* ---- handles ------------------------------------------------------------------------------ */
typedef struct world  *world;
typedef struct source *source;
typedef struct tfsf   *tfsf;
typedef struct slice  *slice;
typedef struct phaser *phaser;
typedef struct matter matter;
typedef struct object object;

typedef float field(world, int, int, int);               /* point-wise derived quantity       */
typedef int   shape(object, float, float, float);        /* inside/outside predicate          */
typedef float waveform(world, source, float, float);     /* (W, S, t, phase) -> sample        */

/* ---- aggregates initialised positionally by drivers (member order is API) ----------------- */
typedef struct {                                         /* resolution                        */
	float dt;
	struct { float x, y, z; } d;
	struct { float eV; int subN; } spec;
	struct { struct { int m; float Min, Max; } x, y, z; } nug;
	struct {
		float (*itox)(world, float);
		float (*jtoy)(world, float);
		float (*ktoz)(world, float);
	} gridFunction;
} res;

typedef struct { struct { float Min, Max; } x, y, z; } dom;   /* domain                       */

typedef struct {                                         /* boundary surfaces + PML parameters */
	struct { float MinSur, MaxSur; } x, y, z;
	struct { int Num; float SigmaMax, KappaMax, AlphaMax, SigmaPower, AlphaPower; } pml;
} sur;

struct matter {
	struct { float x, y, z; } e;                         /* eps_inf (y,z only if anisotropic) */
	struct { float w, r, f, g; } P[MAXPOLES];            /* poles: {w,r} Drude, {0,s} conductivity, {w,r,f} Lorentz, {w,r,f,g} critical point */
	float d[3][6];                                       /* chi(2) tensor (not supported here) */
};

struct object {
	shape *S;
	float f[3][3];
	void *p;
};

/* ---- field storage -------------------------------------------------------------------------
 * Member names (and their order) are the reference's: drivers and the engine bridge address
 * them by name. */
typedef struct {
	float ***x;                                          /* host views of the three components,         */
	float ***y;                                          /*   [iMIN-1..iMAX+1][jMIN-1..jMAX+1][kMIN-1..kMAX+1] */
	float ***z;
	float ***yMinPx, ***zMinPx;                          /* CPML psi arrays (device only, never allocated */
	float ***xMinPy, ***zMinPy;                          /*   on the host): <slab axis><side>P<component> */
	float ***xMinPz, ***yMinPz;
	float ***yMaxPx, ***zMaxPx;
	float ***xMaxPy, ***zMaxPy;
	float ***xMaxPz, ***yMaxPz;
	float *****Jx, *****Jy, *****Jz;                     /* ADE polarisation currents, host views [n][p]  */
	float *****Kx, *****Ky, *****Kz;                     /* second ADE variable of critical-point poles   */
	float t;                                             /* time of this field set                        */
} vfield;

typedef struct {
	float *dx, *dy, *dz;                                 /* cell sizes seen by this field set             */
	float *dtdx, *dtdy, *dtdz;                           /* dt / (kappa * cell size)                      */
	float *xP, *yP, *zP;                                 /* CPML tables: psi decay ...                    */
	float *xPy, *xPz, *yPz, *yPx, *zPx, *zPy;            /*   ... and psi drive per (axis, component)     */
	float ***x, ***y, ***z;                              /* inverse permittivity (E only)                 */
	float ****Jx, ****Jy, ****Jz;                        /* fill factors of dispersive objects [n]        */
	float *ax, *ay, *az;                                 /* conductivity loss factor per object           */
	float **JJx, **JJy, **JJz;                           /* pole coefficients [n][p]: J <- JJ J + JK K + JE E */
	float **JKx, **JKy, **JKz;
	float **JEx, **JEy, **JEz;
	float **KJx, **KJy, **KJz;                           /*                           K <- KJ J + KK K + KE E */
	float **KKx, **KKy, **KKz;
	float **KEx, **KEy, **KEz;
	int N;                                               /* number of dispersive objects                  */
	int *NJ, *NK;                                        /* Drude poles / all poles per object            */
	int *iMin, *iMax, *jMin, *jMax, *kMin, *kMax;        /* bounding box per object                       */
	object *O;                                           /* the object list as given to putObjects ...    */
	matter *M;                                           /* ... and its matter list (last = background)   */
} coeffs;

struct world {
	dom Dom;                                             /* the arguments of createWorld, normalised      */
	res Res;
	sur Sur;
	float t, dt;                                         /* time after the last H update; time step       */
	float dx, dy, dz;                                    /* nominal cell sizes                            */
	float eV, f;                                         /* energy unit of the material tables; source frequency */
	float xMin, xMax, yMin, yMax, zMin, zMax;            /* interior domain ...                           */
	float xMIN, xMAX, yMIN, yMAX, zMIN, zMAX;            /* ... and domain including the absorbing layers */
	float coskx, sinkx, cosky, sinky, coskz, sinkz;      /* Bloch phase per period / mirror parity (+-1)  */
	float coskyp, coskym, sinkyp, sinkym;                /* hexagonal Bloch boundary: the two skew phases */
	float xSize, ySize, zSize;
	float xyArea, xzArea, yzArea;
	int iMin, iMax, iNum, jMin, jMax, jNum, kMin, kMax, kNum;   /* interior index ranges                  */
	int iMIN, iMAX, iNUM, jMIN, jMAX, jNUM, kMIN, kMAX, kNUM;   /* index ranges incl. absorbing layers    */
	int xMinSur, xMaxSur, yMinSur, yMaxSur, zMinSur, zMaxSur;   /* resolved surface codes                 */
	int subN;                                            /* sub-pixel samples per axis                    */
	int complexField;                                    /* 1: a second (imaginary) field set exists      */
	int sourceType;
	unsigned long int n;                                 /* completed steps                               */
	unsigned long int N;                                 /* length of the finite sources, in steps        */
	unsigned long int T;                                 /* source period, in steps                       */
	int (*xtoi)(world, float), (*ytoj)(world, float), (*ztok)(world, float);     /* coordinate -> index   */
	float (*itox)(world, float), (*jtoy)(world, float), (*ktoz)(world, float);   /* (fractional) index -> coordinate */
	char ID[128];                                        /* output prefix                                 */
	vfield *E, *H;                                       /* field sets [1 + complexField]                 */
	coeffs *CE, *CH;
	source SE, SH;                                       /* soft source lists                             */
	tfsf S;                                              /* TFSF list                                     */
	/* --- alis-b200 --- */
	void *b200_bridge;                                   /* alis_bridge*: engine + descriptors */
	void *b200_view;                                     /* host-view bookkeeping              */
	int b200_shared_coeffs;                              /* cloneWorld: CE/CH owned by the source world */
};

struct source {
	waveform *waveForm;
	float wavelength, supplement, resolution, amplitude, phase, frequency, deviation, peaktime;
	float ***Amplitude, ***Phase, ***F;
	int iMin, iMax, jMin, jMax, kMin, kMax, nMin, nMax;
	unsigned long int N;
	source S;
	/* --- alis-b200 --- */
	int b200_set, b200_comp, b200_m;                     /* target array: (E|H, x|y|z, re|im)  */
};

struct tfsf {
	int iMin, iMax, jMin, jMax, kMin, kMax;
	int iMIN, iMAX, jMIN, jMAX, kMIN, kMAX;
	world W;
	tfsf S;
};

struct slice {
	int iMin, iMax, jMin, jMax, kMin, kMax;
	int iMIN, iMAX, iNUM, jMIN, jMAX, jNUM, kMIN, kMAX, kNUM;
	float ***C, ***F;
	float ***(*interpolate)(world, slice, float***);
};

typedef struct { float A; float complex *Ex, *Ey, *Ez, *Hx, *Hy, *Hz; } tfield;
typedef struct { float complex ***Ex, ***Ey, ***Ez, ***Hx, ***Hy, ***Hz; } cfield;

struct phaser {
	world W;
	int iMin, iMax, jMin, jMax, kMin, kMax, kMIN, xNUM, yNUM;
	unsigned long int N;
	float f, A, *dE, *dH;
	float complex expikx, expiky, *ex, *ez;
	cfield xMin, xMax, yMin, yMax, zMin, zMax;           /* host views of the device accumulators */
	/* --- alis-b200 --- */
	void *b200_dev;                                      /* shared device accumulator set (element 0 owns it) */
	int b200_index, b200_count, b200_stale;
	void *b200_arena;                                    /* element 0: page-aligned block holding every host view of the list */
};

/* ---- world ----------------------------------------------------------------------------------- */
world createWorld(dom, res, sur, char *fmt, ...);
world cloneWorld(world);
void  rebootWorld(world);
void  deleteWorld(world);
void  deleteCloneWorld(world);
float ***makeField(int iMin, int iMax, int jMin, int jMax, int kMin, int kMax);
float complex ***makeComplexField(int iMin, int iMax, int jMin, int jMax, int kMin, int kMax);
#define removeField(F,i,j,k) free(&F[i][j][k]),free(&F[i][j]),free(&F[i]),F=0
void *h5read(int *dim, char *fmt, ...);
void  h5write(int *dim, float *data, char *fmt, ...);

/* ---- objects and shapes ------------------------------------------------------------------------ */
void putObjectArray(world, matter*, object*);
void putObjects(world, ...);
#define putObjects(...) ((putObjects)(__VA_ARGS__,(object){0}))
void removeObjects(world);
matter LT(matter, float);
matter nkFile(char*, float, float);
shape Box, Ball, Diamond;
shape RodX, RodY, RodZ, DRodX, DRodY, DRodZ, BiconeX, BiconeY, BiconeZ;
shape HRodXY, HRodXZ, HRodYZ, HRodYX, HRodZX, HRodZY;
shape Translation, Rotation;
shape Combination, Intersection, Difference;
shape Lattice, DLattice, CLatticeX, CLatticeY, CLatticeZ;
shape ImportContour;

/* ---- time stepping (the hot path: CUDA) -------------------------------------------------------- */
void updateE(world);
void updateH(world);
void copyBoundary(float***, float***, float, float, int, int, int, int, int, int, int, int, int);
#define copyBoundary(W,F,x,...) ((copyBoundary)(W->F[0].x,W->complexField?W->F[1].x:0,__VA_ARGS__))
void nonLinearSHG(world, world, float);
void nonLinearTWM(world, world, world, float);
int  timer(long int, long int, char*, ...);
#define timer(...) ((timer)(__VA_ARGS__,""))

/* ---- fields, slices, output ---------------------------------------------------------------------- */
typedef void (*method)(world, slice, float, char*, char*, unsigned long int*);
typedef unsigned long int colormap[];

void exec(char*, ...);
void writeTxt(world, char*, ...);
void writeRow(world, char*, ...);
void writeSpectrum(world, int, float, float, char*, ...);
#define writeTxt(...)      ((writeTxt)(__VA_ARGS__,""))
#define writeRow(...)      ((writeRow)(__VA_ARGS__,INT2NAN(1)))
#define writeSpectrum(...) ((writeSpectrum)(__VA_ARGS__,INT2NAN(1)))

slice createSlice(world, float xMin, float xMax, float yMin, float yMax, float zMin, float zMax);
#define createSliceX(W,yCut,zCut) createSlice(W, W->xMin, W->xMax, yCut, yCut, zCut, zCut)
#define createSliceY(W,xCut,zCut) createSlice(W, xCut, xCut, W->yMin, W->yMax, zCut, zCut)
#define createSliceZ(W,xCut,yCut) createSlice(W, xCut, xCut, yCut, yCut, W->zMin, W->zMax)
#define createSliceXY(W,zCut)     createSlice(W, W->xMin, W->xMax, W->yMin, W->yMax, zCut, zCut)
#define createSliceXZ(W,yCut)     createSlice(W, W->xMin, W->xMax, yCut, yCut, W->zMin, W->zMax)
#define createSliceYZ(W,xCut)     createSlice(W, xCut, xCut, W->yMin, W->yMax, W->zMin, W->zMax)
void deleteSlice(slice);
void (h5)(world, slice, float, char*, char*, colormap);
void (txt)(world, slice, float, char*, char*, colormap);
void (png)(world, slice, float, char*, char*, colormap);
void sliceSnap(world, field, slice, char*, char*, ...);
#define sliceSnap(W,F,S,...)    ((sliceSnap)(W, F, S, #F, #S, __VA_ARGS__))
void sliceTimeAvg(world, field, slice, char*, char*, ...);
#define sliceTimeAvg(W,F,S,...) ((sliceTimeAvg)(W, F, S, #F, #S, __VA_ARGS__))
void sliceFreqDom(world, field, slice, char*, char*, int, float, ...);
#define sliceFreqDom(W,F,S,...) ((sliceFreqDom)(W, F, S, #F, #S, __VA_ARGS__))
float sliceMax(world, field, slice);
float sliceSum(world, field, slice);
float worldMax(world, field);
float worldSum(world, field);
float get(world, field, float x, float y, float z);
float objectAbsorption(world, object);
float poyntingX(world W, float x, float yMin, float yMax, float zMin, float zMax, ...);
float poyntingY(world W, float y, float xMin, float xMax, float zMin, float zMax, ...);
float poyntingZ(world W, float z, float xMin, float xMax, float yMin, float yMax, ...);
float poyntingOut(world, float, float, float, float, float, float);
#define poyntingX(...)  ((poyntingX)(__VA_ARGS__,-INF,INF,-INF,INF))
#define poyntingY(...)  ((poyntingY)(__VA_ARGS__,-INF,INF,-INF,INF))
#define poyntingZ(...)  ((poyntingZ)(__VA_ARGS__,-INF,INF,-INF,INF))
#define poyntingIn(...) (-poyntingOut(__VA_ARGS__))

field RI, LogRI, ContourRI, Contour;
field raweEffx, raweEffy, raweEffz, eEff;
field rawEx, rawEy, rawEz, rawHx, rawHy, rawHz, rawJx, rawJy, rawJz;
field Ex, Ey, Ez, Hx, Hy, Hz, iEx, iEy, iEz, iHx, iHy, iHz, Jx, Jy, Jz;
field ExHy, ExHz, EyHz, EyHx, EzHx, EzHy, Sx, Sy, Sz;
field JE, EE, UE, HH, UH, U, LogJE, LogEE, LogUE, LogHH, LogUH, LogU;
field ScEx, ScEy, ScEz, ScEE, SoEx, SoEy, SoEz, SoEE;
field DivE, JxHy, JxHz, JyHz, JyHx, JzHx, JzHy, Fx, Fy, Fz;
field Txx, Txy, Txz, Tyx, Tyy, Tyz, Tzx, Tzy, Tzz;

/* ---- phasers (running DFT) and far field ------------------------------------------------------------ */
typedef void projmap(int, int, int, int, float[2]);

phaser createPhaser(world, float wavelength, float xMin, float xMax, float yMin, float yMax, float zMin, float zMax, ...);
#define createPhaser(...)  ((createPhaser)(__VA_ARGS__,-INF,INF,-INF,INF,-INF,INF))
phaser createPhasers(world, float LMin, float LMax, float dL, float xMin, float xMax, float yMin, float yMax, float zMin, float zMax, ...);
#define createPhasers(...) ((createPhasers)(__VA_ARGS__,-INF,INF,-INF,INF,-INF,INF))
phaser choosePhaser(phaser, float wavelength);
void updatePhaser(world, phaser);
#define updatePhasers(...) updatePhaser(__VA_ARGS__)
void deletePhaser(phaser);
#define deletePhasers(...) deletePhaser(__VA_ARGS__)
tfield Ppol(phaser, float, float);
tfield Spol(phaser, float, float);
tfield Unpol(phaser, float, float);
float phaserPoynting(phaser);
float phaserPoyntingIn(phaser);
float phaserPoyntingOut(phaser);
float complex farField(phaser, tfield(phaser,float,float), float theta, float phi);
float farFieldI(phaser, tfield(phaser,float,float), float theta, float phi);
float farFieldFlux(phaser, tfield(phaser,float,float), float, float, float, float, float, ...);
#define farFieldFlux(...) ((farFieldFlux)(__VA_ARGS__,0,0))
void farFieldTheta(world, phaser, float, char*, ...);
void farFieldPhi(world, phaser, float, char*, ...);
void farFieldProfile(world, phaser, projmap, int, int, ...);
void writePoyntingSpectrum(world, char*, ...);
#define writePoyntingSpectrum(...) ((writePoyntingSpectrum)(__VA_ARGS__,0))
void writeFarFieldSpectrum(world, tfield(phaser,float,float), float, float, char*, ...);
#define writeFarFieldSpectrum(...) ((writeFarFieldSpectrum)(__VA_ARGS__,0))
projmap CylindericalMap, AzimuthalMap, NorthernAzimuthalMap, SouthernAzimuthalMap;

/* ---- sources -------------------------------------------------------------------------------------------- */
waveform Sine, Pulse, Band;
tfsf createTFSF(world, dom, sur);
void pointDipole(world, field, float x, float y, float z, waveform, float wavelength, float supplement, float amplitude, float phase, ...);
void guidedWaveX(world W, char *file, field F, float x, waveform waveForm, float wavelength, float supplement, float amplitude, float phase, ...);
void guidedWaveY(world W, char *file, field F, float y, waveform waveForm, float wavelength, float supplement, float amplitude, float phase, ...);
void guidedWaveZ(world W, char *file, field F, float z, waveform waveForm, float wavelength, float supplement, float amplitude, float phase, ...);
void planewave(world, dom, tfield(phaser,float,float), float theta, float phi, waveform, float wavelength, float supplement, float amplitude, float phase, ...);
void focusedBeam(world, dom, field, float x, float y, float z, float theta, waveform, float wavelength, float supplement, float spotsize, float amplitude, float phase, ...);
void removeSources(world W);
#define pointDipole(...) ((pointDipole)(__VA_ARGS__,0,0,0,0,0))
#define guidedWaveX(...) ((guidedWaveX)(__VA_ARGS__,0,0,0,0,0))
#define guidedWaveY(...) ((guidedWaveY)(__VA_ARGS__,0,0,0,0,0))
#define guidedWaveZ(...) ((guidedWaveZ)(__VA_ARGS__,0,0,0,0,0))
#define planewave(...)   ((planewave)(__VA_ARGS__,0,0,0,0,0))
#define focusedBeam(...) ((focusedBeam)(__VA_ARGS__,0,0,0,0,0,0))

/* ---- alis-b200 extensions (not in the reference) ------------------------------------------------------------ */
void alisSyncFields(world W);                 /* refresh every host view W->E[m].x/y/z, W->H[m].x/y/z (full D2H)      */
void alisSyncPhaser(phaser P);                /* refresh the host views P->xMin.Ey[...] ... of a phaser array now        */
void alisReadField(world W, int set, int comp, int m, float *dense);   /* raw block incl. ghosts, reference layout  */
void alisWriteField(world W, int set, int comp, int m, const float *dense);
unsigned long alisDeviceGeometryObjects(void); /* objects rasterised on the device so far (ALIS_DEVICE_GEOMETRY=0: host only) */
unsigned long alisDeviceMonitorCalls(void);   /* slices / world reductions evaluated on the device so far (diagnostics) */
const char *alisBackend(void);                /* "cuda-sm_100a"                                                     */

#endif /* ALIS_API_H */
