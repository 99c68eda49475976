/*
 * configs.h -- test configurations written purely against the ALiS public API
 * (createWorld / putObjects / pointDipole / planewave ...).  The same source is compiled
 *   - against the unmodified reference (oracle/_ref)            -> golden behaviour
 *   - against the reference setup + our engine through the bridge -> engine parity
 *   - against our own host layer + engine (the drop-in)          -> end-to-end parity
 * Configurations follow reference examples/*.c (cited per function); IDs are "" so that
 * createWorld does not shell out to `rm -fr` (reference world.c:275-278).
 */
#ifndef HARNESS_CONFIGS_H
#define HARNESS_CONFIGS_H
#include <string.h>

/* pump1/pump2: optional further worlds on the same grid, stepped in lock step and mixed into the
 * main (signal) world after every E update: pump2 == 0 -> nonLinearSHG(W, P1, alt), else
 * nonLinearTWM(W, P1, P2, alt); alias != 0 -> nonLinearTWM(W, W, P1, alt) (signal is its own pump) */
typedef struct { const char *name; world (*make)(void); int default_steps; int order;
                 world (*pump1)(void); world (*pump2)(void); float alt; int alias; } harness_config;

/* examples/Dipole_PML.c:6-13 */
static world cfg_dipole_pml(void)
{
	dom Dom = {{1000}, {1000}, {1000}};
	res Res = {10, {20}, {1240}};
	sur Sur = {{SYM, PML}, {SYM, PML}, {SYM, PML}, {24}};
	world W = createWorld(Dom, Res, Sur, "");
	pointDipole(W, Ez, 0, 0, 0, Pulse, 400, 800, 10000000);
	return W;
}

/* SURVEY 8d synthetic config at a small size: all-PML vacuum cube, Num=12 */
static world cfg_vacuum_cpml(void)
{
	dom Dom = {{-250, 240}, {-190, 180}, {-150, 140}};
	res Res = {5, {10}, {1240}};
	sur Sur = {{PML, PML}, {PML, PML}, {PML, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	pointDipole(W, Ez, 0, 0, 0, Pulse, 400, 800, 1);
	return W;
}

/* PEC cavity, no PML, off-centre Hy dipole */
static world cfg_pec_box(void)
{
	dom Dom = {{-200, 200}, {-150, 150}, {-100, 100}};
	res Res = {5, {10}, {1240}};
	sur Sur = {{PEC, PEC}, {PEC, PEC}, {PEC, PEC}};
	world W = createWorld(Dom, Res, Sur, "");
	pointDipole(W, Hy, 30, -20, 10, Sine, 300, 3, 1);
	return W;
}

/* periodic (PBC) in x,y + PML in z with a dielectric slab and a Band dipole */
static world cfg_pbc_slab(void)
{
	dom Dom = {{-150, 150}, {-100, 100}, {-300, 300}};
	res Res = {5, {10}, {1240, 4}};
	sur Sur = {{PBC}, {PBC}, {PML, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	object Slab = {Box, {{-INF, INF}, {-INF, INF}, {-55, 45}}};
	object Ball1 = {Ball, {{-60, 60}, {-60, 60}, {100, 220}}};
	putObjects(W, n(2.0), Ball1, n(3.4), Slab, Air);
	pointDipole(W, Ex, 10, 20, 150, Band, 500, 900, 1);
	return W;
}

/* examples/Nanoblock_Scattering.c:6-15 : CPML + Au (Drude + 2 critical points) + substrate
 * + oblique TFSF plane wave (complex Bloch 2-D auxiliary world) */
static world cfg_nanoblock(void)
{
	dom DTF = {{300}, {300}, {300}};
	dom DSF = {{400}, {400}, {400}};
	res Res = {5, {10}, {1240}};
	sur Sur = {{PML}, {SYM, PML}, {PML}};
	world W = createWorld(DSF, Res, Sur, "");
	object Nanoblock = {Box, {{-200, 200}, {-200, 200}, {-200, 200}}};
	object Substrate = {Box, {{-INF, INF}, {-INF, INF}, {-INF,-200}}};
	putObjects(W, Au, Nanoblock, n(1.5), Substrate, Air);
	planewave(W, DTF, Ppol, 55, 0, Sine, 500, 20);
	return W;
}

/* normal-incidence S-polarised TFSF (1-D auxiliary world) on a lossy + Lorentz particle */
static world cfg_tfsf_normal(void)
{
	dom DTF = {{200}, {200}, {200}};
	dom DSF = {{300}, {300}, {300}};
	res Res = {5, {10}, {1240}};
	sur Sur = {{SYM, PML}, {SYM, PML}, {PML}};
	world W = createWorld(DSF, Res, Sur, "");
	matter Lossy = {{2.25}, {{0, 0.35}, {2.1, 0.3, 1.7}}};
	object Particle = {Ball, {{-120, 120}, {-120, 120}, {-120, 120}}};
	putObjects(W, Lossy, Particle, Air);
	planewave(W, DTF, Spol, 0, 0, Pulse, 600, 900);
	return W;
}

/* examples/PhC_Band_Diagram.c:6-24 : Bloch (BBC) in x, hexagonal Bloch (HBC) in y, complex fields */
static world cfg_phc(void)
{
	float kx, ky, k=1.2;
	if (k<=0 || k>=3) kx = 0, ky = 0;
	else if (k < 1) kx =   0, ky =   k;
	else if (k < 2) kx = k-1, ky =   1;
	else if (k < 3) kx = 3-k, ky = 3-k;
	dom Dom = {{250}, {0, 250*sqrtf(3)}, {500}};
	res Res = {5, {10, 5*sqrtf(3), 10}, {1240}};
	sur Sur = {{BBC, 500*3/kx}, {HBC, 500*sqrtf(3)/ky}, {SYM, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	object Slab = {Difference, {2}, objects {
						{Box, {{-INF, INF}, {-INF, INF}, {-100, 100}}},
						{DLattice, {{500}, {500*sqrtf(3)}}, objects {
							{RodZ, {{-200, 200}, {-200, 200}, {-INF, INF}}},
						}},
					}};
	putObjects(W, n(3.5), Slab, Air);
	pointDipole(W, iHz, 200, 100*sqrtf(3), 0, Band, 750, 2000);
	return W;
}

/* examples/Metasurface_Farfield.c:6-15 */
static world cfg_metasurface(void)
{
	dom DTF = {{500}, {500}, {300}};
	dom DSF = {{500}, {500}, {500}};
	res Res = {10, {20}, {1240}};
	sur Sur = {{BBC}, {SYM}, {PML}};
	world W = createWorld(DSF, Res, Sur, "");
	object Nanoblock = {Lattice, {{500}, {500}}, objects {{Box, {{0, 200}, {-100, 100}, {-100, 100}}}}};
	object Substrate = {Box, {{-INF, INF}, {-INF, INF}, {-INF,-100}}};
	putObjects(W, n(4), Nanoblock, n(1.5), Substrate, Air);
	planewave(W, DTF, Ppol, 55, 0, Sine, 530, 5);
	return W;
}

/* examples/Dipole_Band.c:6-11 : non-uniform grid, SYM + PML, Band dipole */
static world cfg_dipole_band(void)
{
	dom Dom = {{500}, {500}, {500}};
	res Res = {10, {20}, {1240}, {{2, 100}, {2, 100}, {2, 100}}};
	sur Sur = {{SYM, PML}, {SYM, PML}, {SYM, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	pointDipole(W, Ex, 0, 0, 0, Band, 400, 800);
	return W;
}

/* anisotropic Drude/critical-point matter (poles consumed in x,y,z triples, object.c:364-370),
 * two overlapping dispersive objects, conductivity, PML on one side only */
static world cfg_aniso_ade(void)
{
	dom Dom = {{-200, 200}, {-200, 200}, {-200, 200}};
	res Res = {5, {10}, {1240}};
	sur Sur = {{PML, PEC}, {PBC}, {PEC, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	matter Aniso = {{2.0, 3.0, 4.0}, {{8.0, 0.05}, {7.0, 0.06}, {6.0, 0.07}, {2.5, 0.3, 0.4, 0.2}, {2.7, 0.25, 0.5, 0.1}, {2.9, 0.2, 0.6, 0.3}}};
	object A1 = {Box, {{-90, 60}, {-50, 110}, {-100, 20}}};
	object A2 = {Ball, {{-20, 140}, {-80, 80}, {-40, 120}}};
	putObjects(W, Ag, A2, Aniso, A1, n(1.2));
	pointDipole(W, Ey, -120, 10, 130, Pulse, 450, 700, 1);
	pointDipole(W, Hz, 100, -100, -100, Sine, 500, 4, 1, 30);
	return W;
}

/* chi(2) mixing (update.c:354-411): a nonlinear crystal with a fairly full contracted d tensor and
 * a Lorentz pole, next to a metal (second ADE object, d = 0), all-PML box.  which: 0 signal world,
 * 1 pump at 800 nm (Ey dipole), 2 pump at 1000 nm (Ex dipole, other corner). */
static world chi2_world(int which, int bloch)
{
	dom Dom = {{-200, 200}, {-150, 150}, {-180, 180}};
	res Res = {10, {20}, {1240}};
	sur SurP = {{PML, PML}, {PML, PML}, {PML, PML}};
	sur SurB = {{BBC, 1700}, {PBC}, {PML, PML}};
	world W = createWorld(Dom, Res, bloch ? SurB : SurP, "");
	matter NL = {{2.25, 2.4, 2.6}, {{5.5, 0.4, 1.1}, {5.7, 0.3, 0.9}, {5.9, 0.2, 0.8}},
	             {{0.011, 0.023, -0.017, 0.031, 0.0, 0.027}, {0.013, 0.0, 0.029, -0.019, 0.037, 0.021}, {-0.015, 0.033, 0.041, 0.0, 0.025, 0.0}}};
	object Crystal = {Box, {{-90, 110}, {-80, 60}, {-100, 70}}};
	object Tip = {Ball, {{60, 160}, {-30, 70}, {-20, 80}}};
	putObjects(W, NL, Crystal, Au, Tip, Air);
	if (which == 0) pointDipole(W, Ez, -120, 90, 110, Pulse, 450, 900, 0.2);
	if (which == 1) pointDipole(W, Ey, 0, -20, -140, Pulse, 800, 1600, 1);
	if (which == 2) pointDipole(W, Ex, -40, 100, 20, Sine, 1000, 2, 1);
	return W;
}
static world cfg_chi2_signal(void) { return chi2_world(0, 0); }
static world cfg_chi2_pump1(void)  { return chi2_world(1, 0); }
static world cfg_chi2_pump2(void)  { return chi2_world(2, 0); }
static world cfg_chi2b_signal(void) { return chi2_world(0, 1); }
static world cfg_chi2b_pump1(void)  { return chi2_world(1, 1); }

/* a Bloch-periodic (complex-field) world wide enough for the fused E+H box of the CUDA engine:
 * BBC in x (no slab), PBC in y, PML in z, a tilted dielectric bar and a lossy Drude ball */
static world cfg_bloch_wide(void)
{
	dom Dom = {{-1100, 1100}, {-300, 300}, {-400, 400}};
	res Res = {10, {20}, {1240}};
	sur Sur = {{BBC, 5100}, {PBC}, {PML, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	object Bar = {Box, {{-700, 550}, {-120, 190}, {-90, 110}}};
	object Ball_ = {Ball, {{200, 560}, {-180, 180}, {40, 400}}};
	putObjects(W, n(2.1), Bar, Ag, Ball_, Air);
	pointDipole(W, Ey, -750, 40, -240, Pulse, 500, 900, 1);
	pointDipole(W, iHz, 780, -80, 260, Sine, 600, 3, 0.5);
	return W;
}

/* examples/Waveguide_Propagation.c:6-16: a guided mode launched from a profile file.  The profiles
 * are synthetic (analytic), written with h5write and read back by guidedWaveY/X/Z (reference
 * source.c:127-220): covers the HDF5 round trip, the alignment of a profile that includes the
 * mirror image with a SYM world, and the per-cell amplitude plane source. */
static void guided_profile(const char *name, int n0, int n1, int o0, int o1, float scale, int odd)
{
	int d[3] = { n0, 0, n1 };
	float *p = calloc((size_t)n0*n1, sizeof(float));
	for (int a=0; a<n0; a++) for (int b=0; b<n1; b++) {
		const float u = 10.0f*(a-o0), v = 10.0f*(b-o1);
		p[(size_t)a*n1+b] = scale * (odd ? u/100 : 1) * expf(-(u*u+v*v)/sq(120.0f));
	}
	h5write(d, p, "%s", name);
	free(p);
}

static world cfg_guided_y(void)
{
	dom Dom = {{300}, {400}, {-200, 300}};
	res Res = {5, {10}, {1240}};
	sur Sur = {{SYM, PML}, {PML, PML}, {PML, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	object Waveguide = {Box, {{ -50,  50}, {-INF, INF}, { -50,  50}}};
	object Substrate = {Box, {{-INF, INF}, {-INF, INF}, {-INF, -50}}};
	putObjects(W, n(4), Waveguide, n(1.5), Substrate, Air);
	guided_profile("harness_guided_a", 1+2*W->iMax, 1+W->kMax-W->kMin, W->iMax, -W->kMin, 1.0f, 0);
	guided_profile("harness_guided_b", 1+2*W->iMax, 1+W->kMax-W->kMin, W->iMax, -W->kMin, 0.4f, 1);
	guidedWaveY(W, "harness_guided_a", Ex, -200, Sine, 625, 10, 1, 0);
	guidedWaveY(W, "harness_guided_b", Ez, -200, Sine, 625, 10, 0.5, 90);
	return W;
}

static world cfg_guided_xz(void)
{
	dom Dom = {{-250, 350}, {300}, {350}};
	res Res = {5, {10}, {1240}};
	sur Sur = {{PML, PML}, {PML, PML}, {SYM, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	object Waveguide = {Box, {{-INF, INF}, { -60,  60}, { -40,  40}}};
	putObjects(W, n(3.2), Waveguide, Air);
	guided_profile("harness_guided_c", 1+W->jMax-W->jMin, 1+2*W->kMax, -W->jMin, W->kMax, 1.0f, 0);
	guided_profile("harness_guided_d", 1+W->iMax-W->iMin, 1+W->jMax-W->jMin, -W->iMin, -W->jMin, 0.2f, 0);
	guidedWaveX(W, "harness_guided_c", Ey, -150, Pulse, 600, 900, 1, 0);
	guidedWaveZ(W, "harness_guided_d", Hy, 200, Sine, 625, 10, 1, 30);
	return W;
}

/* every built-in shape type of reference shapes.c:17-221 once (28 objects on a 4 x 3 x 3 raster of
 * 200 nm cells, sub-pixel sampling 3^3), so that each predicate -- and each branch of the device
 * shape interpreter -- is pinned against the unmodified reference through the permittivity arrays */
#define ZOO_BOX(cx,cy,cz,rx,ry,rz) {{(cx)-(rx), (cx)+(rx)}, {(cy)-(ry), (cy)+(ry)}, {(cz)-(rz), (cz)+(rz)}}
static world cfg_shape_zoo(void)
{
	dom Dom = {{-400, 400}, {-300, 300}, {-300, 300}};
	res Res = {5, {10}, {1240, 3}};
	sur Sur = {{PML, PML}, {PML, PML}, {PML, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	static object O[29];
	static matter M[29];
	shape *prim[18] = { Box, Ball, Diamond, BiconeX, BiconeY, BiconeZ, RodX, RodY, RodZ, DRodX, DRodY, DRodZ,
	                    HRodXY, HRodYZ, HRodZX, HRodXZ, HRodYX, HRodZY };
	int q = 0;
	#define ZOO_AT(q) const float cx = -300 + 200*((q)%4), cy = -200 + 200*(((q)/4)%3), cz = -200 + 200*((q)/12)
	for (; q<18; q++) {
		ZOO_AT(q);
		O[q] = (object){prim[q], ZOO_BOX(cx, cy, cz, 71, 63, 77)};
	}
	/* compound objects: the children live in static arrays (a compound literal dies with its block) */
	static object kids[10][3];
	#define ZOO_KID(slot, idx, S_, ...) kids[slot][idx] = (object){S_, __VA_ARGS__}
	{ ZOO_AT(q); ZOO_KID(0, 0, Ball, ZOO_BOX(cx-13.5, cy+7.25, cz-21, 60, 70, 50));
	  O[q++] = (object){Translation, {{13.5, -7.25, 21}}, kids[0]}; }
	{ ZOO_AT(q); ZOO_KID(1, 0, Box, ZOO_BOX(0, 0, 0, 70, 30, 45));
	  O[q++] = (object){Rotation, {{20, 35, 50}, {cx, cy, cz}}, kids[1]}; }
	{ ZOO_AT(q); ZOO_KID(2, 0, RodZ, ZOO_BOX(cx-30, cy, cz, 30, 40, 80)); ZOO_KID(2, 1, Box, ZOO_BOX(cx+20, cy+10, cz, 50, 20, 30)); ZOO_KID(2, 2, Ball, ZOO_BOX(cx, cy-40, cz+30, 35, 35, 35));
	  O[q++] = (object){Combination, {{3}}, kids[2]}; }
	{ ZOO_AT(q); ZOO_KID(3, 0, Ball, ZOO_BOX(cx, cy, cz, 80, 80, 80)); ZOO_KID(3, 1, DRodZ, ZOO_BOX(cx+10, cy, cz, 60, 90, 50));
	  O[q++] = (object){Intersection, {{2}}, kids[3]}; }
	{ ZOO_AT(q); ZOO_KID(4, 0, Box, ZOO_BOX(cx, cy, cz, 80, 80, 40)); ZOO_KID(4, 1, RodZ, ZOO_BOX(cx-30, cy-30, cz, 25, 25, 100)); ZOO_KID(4, 2, Ball, ZOO_BOX(cx+35, cy+35, cz+40, 30, 30, 30));
	  O[q++] = (object){Difference, {{3}}, kids[4]}; }
	{ ZOO_AT(q); ZOO_KID(5, 0, Box, ZOO_BOX(cx, cy, cz, 18, 60, 22));
	  O[q++] = (object){Lattice, {{60, 3}, {0}, {70, 2}}, kids[5]}; }
	{ ZOO_AT(q); ZOO_KID(6, 0, RodZ, ZOO_BOX(cx, cy, cz, 22, 22, 1000));
	  ZOO_KID(7, 0, Box, ZOO_BOX(cx, cy, cz, 85, 85, 60)); ZOO_KID(7, 1, DLattice, {{80}, {80*1.7320508f}}, kids[6]);
	  O[q++] = (object){Intersection, {{2}}, kids[7]}; }
	{ ZOO_AT(q); ZOO_KID(8, 0, Box, ZOO_BOX(cx, cy+45, cz, 60, 18, 14));
	  O[q++] = (object){CLatticeX, {{3}, {cy, cz}}, kids[8]}; }
	{ ZOO_AT(q); ZOO_KID(8, 1, Diamond, ZOO_BOX(cx+45, cy, cz, 30, 60, 25));
	  O[q++] = (object){CLatticeY, {{4}, {cx, cz}}, &kids[8][1]}; }
	{ ZOO_AT(q); ZOO_KID(8, 2, Ball, ZOO_BOX(cx+50, cy, cz, 28, 20, 60));
	  O[q++] = (object){CLatticeZ, {{5}, {cx, cy}}, &kids[8][2]}; }
	for (int o=0; o<q; o++) M[o] = n(1.3f + 0.1f*(o%9));
	M[q] = Air;
	O[q] = (object){0};
	putObjectArray(W, M, O);
	pointDipole(W, Ex, 5, -5, 15, Pulse, 500, 700);
	return W;
}

/* examples/Focused_Beam.c:6-11 (whole-domain beam: a plane source with per-cell amplitude and phase), and
 * a beam injected through a TFSF box in a dielectric half space, focused in x and y, going down */
static world cfg_focused_beam(void)
{
	dom Dom = {{1200}, {1200}, {1200}};
	res Res = {20, {40}, {1240}};
	sur Sur = {{SYM, PML}, {SYM, PML}, {PML}};
	world W = createWorld(Dom, Res, Sur, "");
	focusedBeam(W, Dom, Ex, 0, INF, 0, 0, Sine, 500, 3, 2);
	return W;
}

static world cfg_focused_tfsf(void)
{
	dom Dom = {{-500, 500}, {-400, 400}, {-500, 500}};
	dom DTF = {{-300, 300}, {-250, 250}, {-300, 300}};
	res Res = {10, {20}, {1240}};
	sur Sur = {{PML, PML}, {PML, PML}, {PML, PML}};
	world W = createWorld(Dom, Res, Sur, "");
	object Substrate = {Box, {{-INF, INF}, {-INF, INF}, {-INF, -100}}};
	putObjects(W, n(1.5), Substrate, Air);
	focusedBeam(W, DTF, Ey, 40, -60, 50, 180, Pulse, 600, 900, 1.5, 1, 30);
	return W;
}

/* randomised differential test: a world drawn from the seed in the environment variable FUZZ_SEED --
 * cell sizes (isotropic or not), time step, domain, flat (2-D) axes, boundary types per axis (PML with
 * several depths incl. true CPML, mirror, PEC, periodic, Bloch -> complex fields), graded grids, sub-pixel
 * sampling, up to three objects of dielectric / dispersive / anisotropic matter, one or two point
 * dipoles of random component, waveform and phase.  Only the public API is used; every draw stays
 * inside what the reference handles without undefined behaviour. */
static unsigned fuzz_state;
static unsigned fuzz_next(void) { fuzz_state = fuzz_state*1664525u + 1013904223u; return fuzz_state >> 8; }
static int fuzz_int(int lo, int hi) { return lo + (int)(fuzz_next() % (unsigned)(hi-lo+1)); }
static float fuzz_pick(const float *v, int n) { return v[fuzz_next() % n]; }

static world cfg_fuzz(void)
{
	const char *seed = getenv("FUZZ_SEED");
	fuzz_state = seed ? (unsigned)atoi(seed)*2654435761u + 12345u : 1u;
	for (int q=0; q<4; q++) fuzz_next();
	static const float cell[3] = {10, 12.5f, 20};
	const float dx = fuzz_pick(cell, 3);
	const int aniso = fuzz_int(0, 3) == 0;
	const float dy = aniso ? fuzz_pick(cell, 3) : dx, dz = aniso ? fuzz_pick(cell, 3) : dx;
	const float dmin = MIN(dx, MIN(dy, dz));
	res Res = {(fuzz_int(0, 1) ? 0.5f : 0.4f)*dmin, {dx, dy, dz}, {1240, (int[]){0, 1, 2, 4}[fuzz_int(0, 3)]}};
	dom Dom; sur Sur = {{0}};
	float *dlo[3] = {&Dom.x.Min, &Dom.y.Min, &Dom.z.Min}, *dhi[3] = {&Dom.x.Max, &Dom.y.Max, &Dom.z.Max};
	float *slo[3] = {&Sur.x.MinSur, &Sur.y.MinSur, &Sur.z.MinSur}, *shi[3] = {&Sur.x.MaxSur, &Sur.y.MaxSur, &Sur.z.MaxSur};
	const float d[3] = {dx, dy, dz};
	int sym[3] = {0, 0, 0}, flat[3] = {0, 0, 0}, pml[3] = {0, 0, 0};
	const int big = getenv("FUZZ_BIG") != 0;                  /* 3-D worlds of at least +-180 nm: the fixed probe positions of monitors.c fit */
	const int tfsf_mode = fuzz_int(0, 3) == 0;                /* every fourth world is laid out so that a TFSF plane wave fits */
	const int flat_axis = tfsf_mode || big ? 5 : fuzz_int(0, 5);     /* 0..2: that axis is flat (2-D world), else 3-D */
	for (int a=0; a<3; a++) {
		const float half = big ? d[a]*ceilf(180/d[a] + fuzz_int(0, 3)) : d[a]*fuzz_int(6, 11);
		*dlo[a] = -half - d[a]*fuzz_int(0, 3); *dhi[a] = half;
		switch (tfsf_mode ? (a < 2 ? fuzz_int(0, 1) : 0) : fuzz_int(0, 5)) {
		case 0: *slo[a] = PML; *shi[a] = PML; pml[a] = 1; break;
		case 1: *slo[a] = SYM; *shi[a] = PML; sym[a] = 1; pml[a] = 1; *dlo[a] = half; *dhi[a] = 0; break;
		case 2: *slo[a] = PBC; *shi[a] = 0; break;
		case 3: *slo[a] = BBC; *shi[a] = (*dhi[a] - *dlo[a])*(2 + fuzz_int(0, 3)*0.75f); break;
		case 4: *slo[a] = PEC; *shi[a] = fuzz_int(0, 1) ? PEC : PML; pml[a] = *shi[a] == PML; break;
		default: *slo[a] = PML; *shi[a] = fuzz_int(0, 1) ? PML : PEC; pml[a] = 1; break;
		}
		if (a == flat_axis) { flat[a] = 1; *dlo[a] = *dhi[a] = 0; *slo[a] = PBC; *shi[a] = 0; sym[a] = 0; pml[a] = 0; }
	}
	Sur.pml.Num = (int[]){0, 6, 8, 24}[fuzz_int(0, 3)];
	if (!tfsf_mode && fuzz_int(0, 3) == 0) for (int a=0; a<3; a++) if (pml[a] && !flat[a]) {      /* graded grid outside +-4 cells on one absorbing axis */
		if (a == 0) Res.nug.x.m = 3, Res.nug.x.Min = 4*d[a];
		if (a == 1) Res.nug.y.m = 3, Res.nug.y.Min = 4*d[a];
		if (a == 2) Res.nug.z.m = 2, Res.nug.z.Min = 4*d[a];
		break;
	}
	world W = createWorld(Dom, Res, Sur, "");
	fprintf(stderr, "fuzz: d=%g,%g,%g dt=%g sur=%g/%g %g/%g %g/%g num=%d flat=%d nug=%d%d%d subN=%d\n", dx, dy, dz, Res.dt,
		Sur.x.MinSur, Sur.x.MaxSur, Sur.y.MinSur, Sur.y.MaxSur, Sur.z.MinSur, Sur.z.MaxSur, Sur.pml.Num, flat_axis, Res.nug.x.m, Res.nug.y.m, Res.nug.z.m, Res.spec.subN);

	static object O[4];
	static matter M[4];
	static object kid[3];
	const int nobj = fuzz_int(0, 3);
	matter Lossy = {{2.25}, {{0, 0.35}, {2.1, 0.3, 1.7}}};
	matter Aniso = {{2.0, 3.0, 4.0}, {{8.0, 0.05}, {7.0, 0.06}, {6.0, 0.07}, {2.5, 0.3, 0.4, 0.2}, {2.7, 0.25, 0.5, 0.1}, {2.9, 0.2, 0.6, 0.3}}};
	for (int o=0; o<nobj; o++) {
		const float cx = flat[0] ? 0 : d[0]*fuzz_int(-3, 3), cy = flat[1] ? 0 : d[1]*fuzz_int(-3, 3), cz = flat[2] ? 0 : d[2]*fuzz_int(-3, 3);
		const float rx = flat[0] ? INF : d[0]*(2 + 0.5f*fuzz_int(0, 6)), ry = flat[1] ? INF : d[1]*(2 + 0.5f*fuzz_int(0, 6)), rz = flat[2] ? INF : d[2]*(2 + 0.5f*fuzz_int(0, 6));
		shape *S[6] = {Box, Ball, RodZ, Diamond, DRodX, BiconeY};
		shape *pick = S[fuzz_int(0, 5)];
		if (flat[0] || flat[1] || flat[2]) pick = Box;          /* infinite extents: only the box predicate is meaningful */
		object base = {pick, {{cx-rx, cx+rx}, {cy-ry, cy+ry}, {cz-rz, cz+rz}}};
		if (fuzz_int(0, 4) == 0 && pick != Box) { kid[o] = base; O[o] = (object){Translation, {{d[0], -d[1], 0.5f*d[2]}}, &kid[o]}; }
		else O[o] = base;
		switch (fuzz_int(0, 5)) {
		case 0: M[o] = n(1.5); break;
		case 1: M[o] = n(3.4); break;
		case 2: M[o] = Au; break;
		case 3: M[o] = Ag; break;
		case 4: M[o] = Lossy; break;
		default: M[o] = Aniso; break;
		}
	}
	M[nobj] = fuzz_int(0, 2) ? Air : n(1.3);
	O[nobj] = (object){0};
	if (nobj) putObjectArray(W, M, O);

	/* sometimes a normal-incidence TFSF plane wave (its auxiliary 1-D world is stepped first every pass):
	 * only on 3-D real worlds whose x and y sides absorb or mirror and whose z sides absorb */
	{
		const int ok_xy = (Sur.x.MinSur==PML || Sur.x.MinSur==SYM) && Sur.x.MaxSur==PML && (Sur.y.MinSur==PML || Sur.y.MinSur==SYM) && Sur.y.MaxSur==PML;
		const int want = tfsf_mode;
		if (want && ok_xy && Sur.z.MinSur==PML && Sur.z.MaxSur==PML && flat_axis > 2 && !W->complexField && !Res.nug.x.m && !Res.nug.y.m && !Res.nug.z.m) {
			dom DTF = {{sym[0] ? W->xMax-3*d[0] : W->xMin+3*d[0], sym[0] ? 0 : W->xMax-3*d[0]},
			           {sym[1] ? W->yMax-3*d[1] : W->yMin+3*d[1], sym[1] ? 0 : W->yMax-3*d[1]},
			           {W->zMin+3*d[2], W->zMax-3*d[2]}};
			const float wl = 40*dmin*(1 + 0.25f*fuzz_int(0, 4));
			fprintf(stderr, "fuzz: tfsf\n");
			const int kind = fuzz_int(0, 2);
			if (kind == 0) planewave(W, DTF, fuzz_int(0, 1) ? Ppol : Spol, fuzz_int(0, 1) ? 0 : 180, 0, Pulse, wl, 1.5f*wl);
			else if (kind == 1) planewave(W, DTF, fuzz_int(0, 1) ? Ppol : Spol, 0, 90.0f*fuzz_int(0, 1), Sine, wl, 3);
			else planewave(W, DTF, fuzz_int(0, 1) ? Ppol : Spol, (float[]){25, 40, 55}[fuzz_int(0, 2)], 0, Sine, wl, 4);   /* oblique: Bloch-complex auxiliary world */
		}
	}

	field *Fre[6] = {Ex, Ey, Ez, Hx, Hy, Hz}, *Fim[6] = {iEx, iEy, iEz, iHx, iHy, iHz};
	waveform *Wf[3] = {Pulse, Sine, Band};
	for (int s=0, ns=fuzz_int(1, 2); s<ns; s++) {
		field *F = (W->complexField && fuzz_int(0, 2) == 0) ? Fim[fuzz_int(0, 5)] : Fre[fuzz_int(0, 5)];
		float p[3];
		for (int a=0; a<3; a++) p[a] = flat[a] ? 0 : sym[a] ? d[a]*fuzz_int(0, 3) : d[a]*fuzz_int(-3, 3);
		const int w = fuzz_int(0, 2);
		const float wl = 40*dmin*(1 + 0.25f*fuzz_int(0, 4));
		if (w == 1) pointDipole(W, F, p[0], p[1], p[2], Sine, wl, 3 + fuzz_int(0, 3), 1, 30.0f*fuzz_int(0, 5));
		else pointDipole(W, F, p[0], p[1], p[2], Wf[w], wl, wl*(1.5f + 0.25f*fuzz_int(0, 3)), 1 + fuzz_int(0, 1));
	}
	return W;
}

static const harness_config harness_configs[] = {
	{"dipole_pml",  cfg_dipole_pml,  60, 0},
	{"vacuum_cpml", cfg_vacuum_cpml, 60, 0},
	{"pec_box",     cfg_pec_box,     80, 0},
	{"pbc_slab",    cfg_pbc_slab,    80, 0},
	{"nanoblock",   cfg_nanoblock,   120, 1},
	{"tfsf_normal", cfg_tfsf_normal, 120, 0},
	{"phc",         cfg_phc,         100, 0},
	{"metasurface", cfg_metasurface, 120, 0},
	{"dipole_band", cfg_dipole_band, 80, 0},
	{"aniso_ade",   cfg_aniso_ade,   100, 1},
	{"bloch_wide",  cfg_bloch_wide,  70, 0},
	{"shape_zoo",   cfg_shape_zoo,   90, 0},
	{"focused_beam", cfg_focused_beam, 100, 0},
	{"focused_tfsf", cfg_focused_tfsf, 110, 0},
	{"fuzz",        cfg_fuzz,        40, 0},
	{"guided_y",    cfg_guided_y,    110, 0},
	{"guided_xz",   cfg_guided_xz,   110, 0},
	{"chi2_shg",    cfg_chi2_signal, 80, 0, cfg_chi2_pump1, 0, 0.7f, 0},
	{"chi2_twm",    cfg_chi2_signal, 80, 0, cfg_chi2_pump1, cfg_chi2_pump2, -1.3f, 0},
	{"chi2_alias",  cfg_chi2b_signal, 80, 0, cfg_chi2b_pump1, 0, 0.9f, 1},
	{0, 0, 0, 0}
};

static const harness_config *harness_find(const char *name)
{
	for (const harness_config *c = harness_configs; c->name; c++) if (!strcmp(c->name, name)) return c;
	return 0;
}
#endif
