/*
 * source.c -- waveforms, point dipoles, plane-wave / focused-beam injection through TFSF
 * auxiliary worlds.  Host C: sources are descriptors (box, amplitude/phase scalars or arrays,
 * waveform callback) that engine_bridge.c hands to the GPU engine; per step the engine evaluates
 * the waveform on the host with this libm and adds the sample in a kernel (reference
 * update.c:161-179).  Behaviour follows reference src/source.c (cited per function).
 */
#include <string.h>
#include "alis_c.h"
#include "host_internal.h"

/* ---- waveforms (reference source.c:11-49) -------------------------------------------------------
 * The first call with S->N == 0 initialises frequency / peaktime / deviation / duration N. */

/* CW with a raised-cosine ramp over `supplement` periods; never switches off (N = -1) */
float Sine(world W, source S, float t, float phase)
{
	if (!S->N) {
		S->frequency = 2 * PI / S->wavelength;
		S->peaktime = W->dt * floorf(S->wavelength * S->supplement / W->dt);
		S->deviation = S->peaktime / PI;
		S->N = -1;
	}
	if (t < S->peaktime)
		return sinf(S->frequency * (t-S->peaktime) - phase) * (0.5 * (1 - cosf(t / S->deviation)));
	return sinf(S->frequency * (t-S->peaktime) - phase);
}

/* Gaussian-modulated carrier at `wavelength`, spectral support up to about `supplement` */
float Pulse(world W, source S, float t, float phase)
{
	if (!S->N) {
		S->frequency = 2 * PI / S->wavelength;
		S->deviation = 2 * sq(S->wavelength) / (PI * S->supplement);
		S->peaktime = W->dt * floorf(PI * S->deviation / W->dt);
		S->N = 2 * floorf(PI * S->deviation / W->dt);
	}
	return sinf(S->frequency * (t-S->peaktime) - phase) * expf(-sq((t-S->peaktime) / S->deviation));
}

/* flat band between `wavelength` and `supplement`: difference of two sinc pulses */
float Band(world W, source S, float t, float phase)
{
	(void)phase;
	if (!S->N) {
		S->resolution = S->supplement > S->wavelength ? ceilf(1.25 * S->supplement / S->wavelength + 0.5) : 2;
		S->frequency = 2 * PI / S->wavelength;
		S->deviation = 2 * PI / S->wavelength / (S->resolution - 0.5);
		S->peaktime = floorf(S->wavelength * (S->resolution - 0.5) / W->dt) * W->dt;
		S->N = 2 * S->wavelength * (S->resolution - 0.5) / W->dt;
	}
	if (t == S->peaktime) return 0;
	return 2 * (cosf((t-S->peaktime) * S->deviation * S->resolution) - cosf((t-S->peaktime) * S->deviation)) / ((t-S->peaktime) * S->deviation);
}

/* ---- target bookkeeping ---------------------------------------------------------------------------- */

/* decode a field tag: which set (0 E, 1 H), component, and real/imaginary field set */
static int decode_target(field F, int *set, int *comp, int *m)
{
	field *tags[2][2][3] = { { {Ex, Ey, Ez}, {iEx, iEy, iEz} }, { {Hx, Hy, Hz}, {iHx, iHy, iHz} } };
	for (int s=0; s<2; s++) for (int im=0; im<2; im++) for (int c=0; c<3; c++)
		if (F == tags[s][im][c]) { *set = s; *comp = c; *m = im; return 1; }
	return 0;
}

/* staggering: a component is "along" axis a if it sits half a cell back on that axis
 * (Ex,Hy,Hz along x; Ey,Hz,Hx along y; Ez,Hx,Hy along z) */
static int along(field F, int axis)
{
	int set, comp, m;
	if (!decode_target(F, &set, &comp, &m)) return 0;
	return set == 0 ? comp == axis : comp != axis;
}

/* reference source.c:52-75 */
static source new_source(world W, field F, waveform waveForm, float wavelength, float supplement, float resolution)
{
	source S;
	MALLOC(S, 0, 1);
	*S = (struct source) { waveForm, wavelength, supplement, resolution };
	waveForm(W, S, 0, 0);                          /* initialises S->N, peaktime, ... */

	int set = -1, comp = 0, m = 0;
	decode_target(F, &set, &comp, &m);
	S->b200_set = set, S->b200_comp = comp, S->b200_m = m;
	if (set >= 0) {                                /* append to the E or H source list */
		source *tail = set == 0 ? &W->SE : &W->SH;
		while (*tail) tail = &(*tail)->S;
		*tail = S;
	}

	if (W->N < S->N && S->N != -1) W->N = S->N;
	if (wavelength && W->T != wavelength/W->dt) W->f = 1/wavelength, W->T = wavelength/W->dt;
	if (!W->sourceType) W->sourceType = 1;
	alis_host_dirty(W, ALIS_DIRTY_SOURCES | ALIS_DIRTY_COPIES);
	return S;
}

/* parity of a symmetric boundary for a source of component F sitting on (on_plane) or off the
 * mirror plane (reference source.c:91-96) */
static void set_mirror_parity(world W, field F, int i, int j, int k)
{
	if (W->xMinSur==SYM && !i) W->coskx = along(F, 0) ? 1 : -1;
	if (W->yMinSur==SYM && !j) W->cosky = along(F, 1) ? 1 : -1;
	if (W->zMinSur==SYM && !k) W->coskz = along(F, 2) ? 1 : -1;
	if (W->xMinSur==SYM && !W->coskx) W->coskx = along(F, 0) ? -1 : 1;
	if (W->yMinSur==SYM && !W->cosky) W->cosky = along(F, 1) ? -1 : 1;
	if (W->zMinSur==SYM && !W->coskz) W->coskz = along(F, 2) ? -1 : 1;
	alis_host_dirty(W, ALIS_DIRTY_COPIES);
}

static void refresh_period(world W, float wavelength)
{
	if (wavelength && W->T != wavelength/W->dt) W->f = 1/wavelength, W->T = wavelength/W->dt;
}

/* ---- point dipole (reference source.c:78-124) ---------------------------------------------------- */
void (pointDipole)(world W, field F, float x, float y, float z, waveform waveForm, float wavelength, float supplement, float amplitude, float phase, ...)
{
	ALIS_PROF("pointDipole");
	int i = W->xtoi(W, x), j = W->ytoj(W, y), k = W->ztok(W, z);
	set_mirror_parity(W, F, i, j, k);

	if (waveForm) {
		const int flatx = W->Dom.x.Min==W->Dom.x.Max, flaty = W->Dom.y.Min==W->Dom.y.Max, flatz = W->Dom.z.Min==W->Dom.z.Max;
		if (!W->sourceType) W->sourceType = (flatx || flaty || flatz) ? 2 : 3;
		source S = new_source(W, F, waveForm, wavelength, supplement, amplitude);

		/* two cells along the component's own axis so the dipole is centred on the node */
		S->iMin = S->iMax = i, S->jMin = S->jMax = j, S->kMin = S->kMax = k;
		if (along(F, 0)) S->iMax++;
		if (along(F, 1)) S->jMax++;
		if (along(F, 2)) S->kMax++;

		/* normalisation to unit radiated power: 2-D line source or 3-D dipole */
		if (!amplitude) amplitude = 1;
		if      (flatx) amplitude *= W->dx * sqrtf(2.5) * sqrtf(wavelength) / sq(W->dt);
		else if (flaty) amplitude *= W->dy * sqrtf(2.5) * sqrtf(wavelength) / sq(W->dt);
		else if (flatz) amplitude *= W->dz * sqrtf(2.5) * sqrtf(wavelength) / sq(W->dt);
		else amplitude *= sqrtf(1.5 / PI) * wavelength / sq(W->dt);

		S->phase = phase * PI / 180;
		S->Amplitude = makeField(S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);

		/* amplitude * dt/dx * dt/dy * dt/dz [* eps^-1]; the spacing table of the dual grid is used
		 * along the component's own axis for E, along the other two for H */
		int set, comp, m;
		decode_target(F, &set, &comp, &m);
		coeffs *CE = W->CE, *CH = W->CH;
		float *tx = (set==0) == (comp==0) ? CH->dtdx : CE->dtdx;
		float *ty = (set==0) == (comp==1) ? CH->dtdy : CE->dtdy;
		float *tz = (set==0) == (comp==2) ? CH->dtdz : CE->dtdz;
		float ***eps = set==0 ? (comp==0 ? CE->x : comp==1 ? CE->y : CE->z) : (comp==0 ? CH->x : comp==1 ? CH->y : CH->z);
		float a = set==0 ? amplitude*0.5 : amplitude*0.25;
		for (int ii=S->iMin; ii<=S->iMax; ii++) for (int jj=S->jMin; jj<=S->jMax; jj++) for (int kk=S->kMin; kk<=S->kMax; kk++)
			S->Amplitude[ii][jj][kk] = a * tx[ii] * ty[jj] * tz[kk] * (eps ? eps[ii][jj][kk] : 1);
	}
	refresh_period(W, wavelength);
}

/* ---- guided-wave sources (reference source.c:127-220) ----------------------------------------------
 * A plane source over the whole cross-section at x, y or z whose per-cell amplitude is a mode
 * profile read from <file>.h5 (normally written earlier by sliceFreqDom(..., h5, ...) of a mode
 * calculation: Waveguide_Dispersion.c -> Waveguide_Propagation.c).  The profile's last row/column
 * is aligned with the world's upper edge, so a profile computed on a mirror-symmetric world (which
 * stores the mirror image too) fits a world with or without the mirror. */
static void guided_wave(world W, int axis, char *file, field F, float pos, waveform waveForm, float wavelength, float supplement, float amplitude, float phase)
{
	const int b = axis == 0 ? 1 : 0, c = axis == 2 ? 1 : 2;          /* the two transverse axes, in index order */
	const int minSur[3] = { W->xMinSur, W->yMinSur, W->zMinSur };
	const int Min[3] = { W->iMin, W->jMin, W->kMin }, Max[3] = { W->iMax, W->jMax, W->kMax };
	float *cosk[3] = { &W->coskx, &W->cosky, &W->coskz };
	const int at = axis == 0 ? W->xtoi(W, pos) : axis == 1 ? W->ytoj(W, pos) : W->ztok(W, pos);

	if (minSur[b]==SYM) *cosk[b] = along(F, b) ? 1 : -1;
	if (minSur[c]==SYM) *cosk[c] = along(F, c) ? 1 : -1;
	if (minSur[axis]==SYM && !at) *cosk[axis] = along(F, axis) ? 1 : -1;
	if (minSur[axis]==SYM && !*cosk[axis]) *cosk[axis] = along(F, axis) ? -1 : 1;
	alis_host_dirty(W, ALIS_DIRTY_COPIES);

	if (waveForm) {
		int d[2];
		float **f = h5read(d, "%s", file);
		if (!amplitude) amplitude = 1;
		source S = new_source(W, F, waveForm, wavelength, supplement, amplitude);

		if (d[0] != 1+Max[b]+(minSur[b]==SYM?Max[b]:-Min[b])) printf("The dimension of guided wave does not match!\n"), exit(0);
		if (d[1] != 1+Max[c]+(minSur[c]==SYM?Max[c]:-Min[c])) printf("The dimension of guided wave does not match!\n"), exit(0);
		S->phase = phase * PI / 180;
		S->iMin = W->iMin, S->iMax = W->iMax;
		S->jMin = W->jMin, S->jMax = W->jMax;
		S->kMin = W->kMin, S->kMax = W->kMax;
		if (axis == 0) S->iMin = S->iMax = at;
		if (axis == 1) S->jMin = S->jMax = at;
		if (axis == 2) S->kMin = S->kMax = at;
		S->Amplitude = makeField(S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);
		for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++) {
			const int q[3] = { i, j, k };
			S->Amplitude[i][j][k] = amplitude * f[q[b]-Max[b]+d[0]-1][q[c]-Max[c]+d[1]-1];
		}
		free(f[0]); free(f);
	}
	refresh_period(W, wavelength);
}

void (guidedWaveX)(world W, char *file, field F, float x, waveform waveForm, float wavelength, float supplement, float amplitude, float phase, ...) { guided_wave(W, 0, file, F, x, waveForm, wavelength, supplement, amplitude, phase); }
void (guidedWaveY)(world W, char *file, field F, float y, waveform waveForm, float wavelength, float supplement, float amplitude, float phase, ...) { guided_wave(W, 1, file, F, y, waveForm, wavelength, supplement, amplitude, phase); }
void (guidedWaveZ)(world W, char *file, field F, float z, waveform waveForm, float wavelength, float supplement, float amplitude, float phase, ...) { guided_wave(W, 2, file, F, z, waveForm, wavelength, supplement, amplitude, phase); }

/* ---- total-field / scattered-field region (reference source.c:251-334) ---------------------------- */

/* give the auxiliary world the layer stack found OUTSIDE the TFSF box: skip the leading objects
 * that never occur there (reference source.c:251-266) */
static void copy_background_stack(world W, tfsf S, matter *M, object *O)
{
	int N = 0;
	while (O[N].S) N++;
	for (int i=W->iMIN; i<=W->iMAX; i++) for (int j=W->jMIN; j<=W->jMAX; j++) for (int k=W->kMIN; k<=W->kMAX; k++) {
		if (i < S->iMin || S->iMax < i || j < S->jMin || S->jMax < j || k < S->kMin || S->kMax < k) {
			int first = 0;
			while (O[first].S && !(O[first].S(O[first], W->itox(W, i), W->jtoy(W, j), W->ktoz(W, k)))) first++;
			if (first < N) N = first;
		}
	}
	if (O[N].S || M[N].e.x != 1 || M[N].e.y && M[N].e.y != 1 || M[N].e.z && M[N].e.z != 1 || M[N].P->w || M[N].P->g)
		putObjectArray(S->W, M+N, O+N);
}

/* The auxiliary world carries the incident field.  Axes marked NIL in SUR are collapsed to a
 * two-cell periodic (or Bloch / symmetric) strip -- the engine reads it with clamped indices
 * where the reference aliases row pointers (source.c:223-248, 298-317). */
/* one axis of the main world's grid, so that the per-axis rules below are written once */
typedef struct {
	int min_sur;
	float outer_lo;                               /* xMIN: coordinate of the outermost cell      */
	int iMin, iMax, iMIN, iMAX;                   /* interior and outer index range              */
	int (*to_index)(world, float);
	float (*to_coord)(world, float);
} grid_axis;

static grid_axis grid_axis_of(world W, int a)
{
	const grid_axis x = { W->xMinSur, W->xMIN, W->iMin, W->iMax, W->iMIN, W->iMAX, W->xtoi, W->itox };
	const grid_axis y = { W->yMinSur, W->yMIN, W->jMin, W->jMax, W->jMIN, W->jMAX, W->ytoj, W->jtoy };
	const grid_axis z = { W->zMinSur, W->zMIN, W->kMin, W->kMax, W->kMIN, W->kMAX, W->ztok, W->ktoz };
	return a == 0 ? x : a == 1 ? y : z;
}

/* transverse axis the caller left open (NIL): the auxiliary world is one cell wide there and closes
 * on itself the way the main world does (mirror, Bloch or plain periodic) */
static void collapse_transverse(world W, int a, float *lo, float *hi, float *min_sur)
{
	const grid_axis A = grid_axis_of(W, a);
	const int mirror = A.min_sur == SYM;
	*lo = mirror ? 0 : A.outer_lo;
	*hi = mirror ? 0 : A.to_coord(W, 1 + A.to_index(W, A.outer_lo));
	*min_sur = mirror ? SYM : (A.min_sur == BBC) ? BBC : PBC;
}

/* requested TFSF extent along one axis -> index range lo..hi of the box (snapped to the grid, pushed
 * to the outer boundary where it reaches into the PML) and tangential range LO..HI of the strips */
static void snap_box_axis(world W, int a, int mirrored, float want_lo, float want_hi, int *lo, int *hi, int *LO, int *HI)
{
	const grid_axis A = grid_axis_of(W, a);
	if ((!want_hi && want_lo > 0) || mirrored) {          /* a single positive number = symmetric extent */
		want_hi = want_hi ? want_hi : want_lo;
		want_lo = -want_hi;
	}
	const int whole = want_lo == want_hi;                 /* degenerate request: the whole interior */
	*lo = A.min_sur == SYM ? 0 : whole ? A.iMin : A.iMin > A.to_index(W, want_lo) ? A.iMIN : A.to_index(W, want_lo);
	*hi = whole ? A.iMax : A.iMax < A.to_index(W, want_hi) ? A.iMAX : A.to_index(W, want_hi);
	*LO = (*lo == A.iMIN) ? *lo : *lo + 1;
	*HI = (*hi == A.iMAX) ? *hi + 1 : *hi;
}

tfsf (createTFSF)(world W, dom DOM, sur SUR)
{
	ALIS_PROF("createTFSF");
	tfsf S;
	MALLOC(S, 0, 1);
	tfsf *tail = &W->S;
	while (*tail) tail = &(*tail)->S;
	*tail = S;

	/* the auxiliary world that carries the incident field: the main world's domain, collapsed along
	 * the open transverse axes and padded by one cell beyond both z faces */
	dom aux_dom = W->Dom;
	sur aux_sur = SUR;
	if (SUR.x.MinSur==NIL) collapse_transverse(W, 0, &aux_dom.x.Min, &aux_dom.x.Max, &aux_sur.x.MinSur);
	if (SUR.y.MinSur==NIL) collapse_transverse(W, 1, &aux_dom.y.Min, &aux_dom.y.Max, &aux_sur.y.MinSur);
	if (SUR.z.MinSur==NIL) {
		aux_dom.z.Min = W->ktoz(W, W->ztok(W, W->zMin)-1);
		aux_dom.z.Max = W->ktoz(W, W->ztok(W, W->zMax)+1);
		aux_sur.z.MinSur = aux_sur.z.MaxSur = PML;
	}
	S->W = createWorld(aux_dom, W->Res, aux_sur, "");
	strcpy(S->W->ID, W->ID);

	snap_box_axis(W, 0, SUR.x.MinSur==SYM, DOM.x.Min, DOM.x.Max, &S->iMin, &S->iMax, &S->iMIN, &S->iMAX);
	snap_box_axis(W, 1, SUR.y.MinSur==SYM, DOM.y.Min, DOM.y.Max, &S->jMin, &S->jMax, &S->jMIN, &S->jMAX);
	snap_box_axis(W, 2, SUR.z.MinSur==SYM, DOM.z.Min, DOM.z.Max, &S->kMin, &S->kMax, &S->kMIN, &S->kMAX);

	if (W->CE->M && W->CE->O) copy_background_stack(W, S, W->CE->M, W->CE->O);
	alis_host_dirty(W, ALIS_DIRTY_SOURCES | ALIS_DIRTY_COPIES);
	return S;
}

/* ---- plane sources (reference source.c:337-411) --------------------------------------------------- */

/* a source covering the whole plane z = k, absorbing layers included */
static void whole_plane(source S, world W, int k)
{
	S->iMin = W->iMIN; S->iMax = W->iMAX;
	S->jMin = W->jMIN; S->jMax = W->jMAX;
	S->kMin = S->kMax = k;
}

/* one plane source z = k of component F; on a complex (Bloch) world every cell gets its own phase
 * phase + kx x + ky y so the plane radiates obliquely */
static void plane_source(world W, field F, int k, float kx, float ky, waveform waveForm, float wavelength, float supplement, float resolution, float amplitude, float phase)
{
	source S = new_source(W, F, waveForm, wavelength, supplement, resolution);
	S->amplitude = amplitude;
	S->phase = phase;
	whole_plane(S, W, k);

	if (W->complexField) {
		float di = (F==Ex||F==Hy||F==iEx||F==iHy) ? 0.5 : 0;
		float dj = (F==Ey||F==Hx||F==iEy||F==iHx) ? 0.5 : 0;
		S->Phase = makeField(S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);
		for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++)
			S->Phase[i][j][k] = phase + kx * W->itox(W, i-di) + ky * W->jtoy(W, j-dj);
	}
}

typedef struct { field *re, *im; int isH; float amp; int on; } plane_spec;

/* Sets the Bloch / mirror phases of W for a plane wave from (theta, phi) and, when a waveform is
 * given, adds the E and H plane sources (and their imaginary twins on complex worlds). */
static void plane_sources(world W, tfield pol(phaser,float,float), float theta, float phi, waveform waveForm, float wavelength, float supplement, float resolution, float amplitude, float phase)
{
	const int down = 90<theta && theta<270, up = theta<90 || 270<theta;
	int kE = down ? W->kMin : up ? W->kMax : 0;
	int kH = down ? W->kMin : up ? W->kMax+1 : 0;
	float sinth = sinf(theta*PI/180) * sqrtf(W->CE->z?1/W->CE->z[W->iMin][W->jMin][kH]:1);
	float costh = cosf(theta*PI/180) * sqrtf(W->CE->z?1/W->CE->z[W->iMin][W->jMin][kH]:1);
	float sinph = sinf(phi*PI/180);
	float cosph = cosf(phi*PI/180);
	float kx = -2 * PI * sinth * cosph / wavelength;
	float ky = -2 * PI * sinth * sinph / wavelength;

	if (W->xMinSur == BBC) W->coskx = cosf(kx*(W->xMax-W->xMin)), W->sinkx = sinf(kx*(W->xMax-W->xMin));
	if (W->yMinSur == BBC) W->cosky = cosf(ky*(W->yMax-W->yMin)), W->sinky = sinf(ky*(W->yMax-W->yMin));
	if ((W->xMinSur == SYM) && (phi==90 || phi==270)) W->coskx = pol==Ppol ? -1 : 1;
	if ((W->xMinSur == SYM) && (phi==0  || phi==180) && (theta==0 || theta==180)) W->coskx = pol==Ppol ? 1 : -1;
	if ((W->yMinSur == SYM) && (phi==0  || phi==180)) W->cosky = pol==Ppol ? -1 : 1;
	if ((W->yMinSur == SYM) && (phi==90 || phi==270) && (theta==0 || theta==180)) W->cosky = pol==Ppol ? 1 : -1;
	alis_host_dirty(W, ALIS_DIRTY_COPIES);

	if (waveForm) {
		float n = (W->CE->z ? 1/sqrtf(W->CE->z[W->iMin][W->jMin][kH]) : 1);
		float kzdz = W->CE->dz[kE] * PI * fabsf(costh) / wavelength;
		float ampE = W->CE->dtdz[kE] * (amplitude ? amplitude : 1) * 2 / ((1+n) * sqrtf(n) * n);
		float ampH = W->CH->dtdz[kH] * (amplitude ? amplitude : 1) * 2 / ((1+n) * sqrtf(n));
		const int needx = phi != 90 && phi != 270;       /* wave vector has an x-z plane part  */
		const int needy = phi !=  0 && phi != 180;       /* ... a y-z plane part               */

		/* (real tag, imaginary tag, is H, amplitude, emitted when) in the reference's order */
		plane_spec list[4];
		if (pol == Ppol) {
			list[0] = (plane_spec){ Ex, iEx, 0, ampE*cosph*costh, needx };
			list[1] = (plane_spec){ Ey, iEy, 0, ampE*sinph*costh, needy };
			list[2] = (plane_spec){ Hx, iHx, 1, ampH*sinph,       needy };
			list[3] = (plane_spec){ Hy, iHy, 1,-ampH*cosph,       needx };
		}
		else if (pol == Spol) {
			list[0] = (plane_spec){ Ex, iEx, 0,-ampE*sinph,       needy };
			list[1] = (plane_spec){ Ey, iEy, 0, ampE*cosph,       needx };
			list[2] = (plane_spec){ Hx, iHx, 1, ampH*cosph*costh, needx };
			list[3] = (plane_spec){ Hy, iHy, 1, ampH*sinph*costh, needy };
		}
		else return;
		/* real parts lead by 90 degrees; H planes sit half a cell further along z */
		for (int q=0; q<4; q++) if (list[q].on)
			plane_source(W, list[q].re, list[q].isH ? kH : kE, kx, ky, waveForm, wavelength, supplement, resolution, list[q].amp,
			             list[q].isH ? (phase+90)*PI/180+kzdz : (phase+90)*PI/180);
		if (W->complexField) for (int q=0; q<4; q++) if (list[q].on)
			plane_source(W, list[q].im, list[q].isH ? kH : kE, kx, ky, waveForm, wavelength, supplement, resolution, list[q].amp,
			             list[q].isH ? phase*PI/180+kzdz : phase*PI/180);
	}
	refresh_period(W, wavelength);
}

/* reference source.c:414-432 */
void (planewave)(world W, dom Dom, tfield pol(phaser,float,float), float theta, float phi, waveform waveForm, float wavelength, float supplement, float amplitude, float phase, ...)
{
	ALIS_PROF("planewave");
	while (theta < 0) theta += 360;
	while (theta >= 360) theta -= 360;
	while (phi < 0) phi += 360;
	while (phi >= 360) phi -= 360;

	/* oblique incidence needs Bloch-periodic auxiliary axes; the rest are collapsed */
	sur Sur = {{NIL}, {NIL}, {NIL}};
	if ((theta !=0 && theta != 180) && (phi != 90 && phi != 270)) Sur.x.MinSur = BBC;
	if ((theta !=0 && theta != 180) && (phi !=  0 && phi != 180)) Sur.y.MinSur = BBC;

	float resolution = amplitude;
	if (!amplitude) amplitude = 1;

	tfsf S = createTFSF(W, Dom, Sur);
	plane_sources(W, pol, theta, phi, 0, wavelength, supplement, resolution, 0, 0);
	plane_sources(S->W, pol, theta, phi, waveForm, wavelength, supplement, resolution, amplitude, phase);
	W->N = S->W->N;
}

/* ---- Gaussian beam (reference source.c:435-496) ----------------------------------------------------
 * A plane source over a whole z plane whose per-cell amplitude and phase are those of a paraxial
 * Gaussian beam focused at height z: waist w0 from the FWHM spot size, local width w, wavefront
 * radius R and Gouy phase at the distance of the source plane from the focus.  The expressions are
 * the reference's (operand order included): the profile is bit-identical. */
typedef struct { float k, w0, w, R, gouy; } beam_optics;

static beam_optics beam_optics_at(float wavelength, float index, float spotsize, float distance)
{
	beam_optics b;
	b.k = 2 * PI * index / wavelength;
	b.w0 = wavelength * (spotsize ? spotsize : 1) / (index * sqrtf(-2.0 * logf(0.5)));
	const float rayleigh = PI * index * sq(b.w0) / wavelength;
	b.w = b.w0 * sqrtf(1 + sq(distance/rayleigh));
	b.R = distance * (1 + sq(rayleigh/distance));
	b.gouy = atanf(distance/rayleigh);
	return b;
}

/* the source plane of a wave travelling along theta: the lower face for downward, the upper one for
 * upward propagation */
static int launch_plane(world W, float theta)
{
	if (90<theta && theta<270) return W->kMin;
	if (theta<90 || 270<theta) return W->kMax;
	return 0;
}

static void beam(world W, field F, float x, float y, float z, float theta, waveform waveForm, float wavelength, float supplement, float spotsize, float amplitude, float phase)
{
	while (theta < 0) theta += 360;
	while (theta >= 360) theta -= 360;

	/* mirror parities of a beam polarised along F */
	if (W->xMinSur==SYM) W->coskx = (F==Ex || F==iEx) ? 1 : -1;
	if (W->yMinSur==SYM) W->cosky = (F==Ey || F==iEy) ? 1 : -1;
	alis_host_dirty(W, ALIS_DIRTY_COPIES);

	if (waveForm) {
		const int k = launch_plane(W, theta);
		float index = 1;                            /* refractive index where the beam axis crosses the source plane */
		if (W->CE->z) index = 1/sqrtf(W->CE->z[W->xtoi(W,MIN(x,W->xMax))][W->ytoj(W,MIN(y,W->yMax))][k]);
		source S = new_source(W, F, waveForm, wavelength, supplement, amplitude);
		whole_plane(S, W, k);
		if (!amplitude) amplitude = 1;
		amplitude /= sqrtf(index*index*index);

		if (!spotsize) {                            /* plane wave */
			S->amplitude = amplitude;
			S->phase = phase * PI / 180;
		}
		else {
			const beam_optics B = beam_optics_at(wavelength, index, spotsize, fabsf(z - W->ktoz(W,k)));
			const int focus_x = x < INF, focus_y = y < INF;      /* an infinite coordinate = no focusing along that axis */
			S->Amplitude = makeField(S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);
			S->Phase = makeField(S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);
			for (int i=S->iMin; i<=S->iMax; i++) {
				const float ux = focus_x ? sq(x - W->itox(W,F==Ex?i-0.5:i)) : 0;
				for (int j=S->jMin; j<=S->jMax; j++) {
					float r2 = 0;
					if (focus_x) r2 += ux;
					if (focus_y) r2 += sq(y - W->jtoy(W,F==Ey?j-0.5:j));
					S->Amplitude[i][j][k] = amplitude * B.w0/B.w * expf(-r2/sq(B.w));
					S->Phase[i][j][k] = phase * PI / 180 - B.k * (0.5 * r2 / B.R) + B.gouy;
				}
			}
		}
	}
	refresh_period(W, wavelength);
}

void (focusedBeam)(world W, dom Dom, field F, float x, float y, float z, float theta, waveform waveForm, float wavelength, float supplement, float spotsize, float amplitude, float phase, ...)
{
	const int whole = Dom.x.Min==W->Dom.x.Min && Dom.x.Max==W->Dom.x.Max &&
	                  Dom.y.Min==W->Dom.y.Min && Dom.y.Max==W->Dom.y.Max &&
	                  Dom.z.Min==W->Dom.z.Min && Dom.z.Max==W->Dom.z.Max;
	if (whole) {
		beam(W, F, x, y, z, theta, waveForm, wavelength, supplement, spotsize, amplitude, phase);
		return;
	}
	/* TFSF injection: the auxiliary world keeps the real boundaries along the focused axes */
	sur Sur = {{NIL}, {NIL}, {NIL}};
	if (x < INF) Sur.x.MinSur = W->Sur.x.MinSur, Sur.x.MaxSur = W->Sur.x.MaxSur;
	if (y < INF) Sur.y.MinSur = W->Sur.y.MinSur, Sur.y.MaxSur = W->Sur.y.MaxSur;
	tfsf S = createTFSF(W, Dom, Sur);
	beam(W, F, x, y, z, theta, 0, wavelength, supplement, 0, 0, 0);
	beam(S->W, F, x, y, z, theta, waveForm, wavelength, supplement, spotsize, amplitude, phase);
	W->N = S->W->N;
}

/* ---- removal (reference source.c:499-524) ------------------------------------------------------------ */
static void free_source_list(source S)
{
	while (S) {
		source next = S->S;
		if (S->Amplitude) removeField(S->Amplitude, S->iMin, S->jMin, S->kMin);
		if (S->Phase) removeField(S->Phase, S->iMin, S->jMin, S->kMin);
		free(S);
		S = next;
	}
}

void removeSources(world W)
{
	free_source_list(W->SE), W->SE = 0;
	free_source_list(W->SH), W->SH = 0;
	for (tfsf S = W->S; S; ) {
		tfsf next = S->S;
		deleteWorld(S->W);
		free(S);
		S = next;
	}
	W->S = 0;
	W->f = 0;
	W->T = 0;
	W->N = 0;
	alis_host_dirty(W, ALIS_DIRTY_SOURCES);
}
