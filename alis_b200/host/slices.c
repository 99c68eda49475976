/*
 * slices.c -- slice monitors: a `field` evaluated over a line / plane / box of the grid, as a
 * snapshot, a running time average or a single-frequency running DFT, handed to a writer
 * (txt, png, h5).
 *
 * Only the slice region is fetched from the device (view.c); the field function is then applied
 * on the host.  Behaviour follows reference src/slices.c (cited per function).  PNG files are
 * produced by a small built-in encoder (8-bit RGB, stored-deflate zlib stream) because libpng is
 * not available here; pixel values are computed exactly as the reference's png() does.
 */
#include <sys/stat.h>
#include <string.h>
#include <limits.h>
#include "alis_c.h"
#include "host_internal.h"

static char *dup_string(const char *s) { char *d = malloc(strlen(s)+1); strcpy(d, s); return d; }   /* (strdup is not ISO C99) */

/* per (world, writer, name) state (reference slices.c:13-22, 214-239) */
typedef struct sentry {
	world W;
	method M;
	char *name;
	float max, Max, ***F, ***G;
	alisb_accum *dev;             /* the running sums live on the device (F has a device twin): F, G are filled when a frame is due */
	int dev_tried;
	long calls;
	struct sentry *next;
} sentry;

static sentry *lookup(world W, char *name, method M, sentry **list)
{
	sentry *e = *list;
	while (e && (e->W != W || e->M != M || strcmp(name, e->name))) e = e->next;
	if (!e) {
		e = calloc(1, sizeof *e);
		e->W = W; e->M = M; e->name = dup_string(name);
		e->next = *list;
		*list = e;
		char path[1024];
		strcat(strcpy(path, W->ID), e->name);
		if (e->name[0]=='/') mkdir(W->ID, 0755);
		if (path[0] && path[strlen(path)-1]=='/') mkdir(path, 0755);
	}
	char path[1024];
	strcat(strcpy(path, W->ID), e->name);
	strcpy(name, path);
	e->calls++;
	return e;
}

/* bring the slice region (mirror images folded back) into the host views */
static void prefetch(world W, field F, slice S)
{
	int lo[3] = { S->iMin, S->jMin, S->kMin }, hi[3] = { S->iMax, S->jMax, S->kMax };
	const int sym[3] = { W->xMinSur==SYM, W->yMinSur==SYM, W->zMinSur==SYM };
	for (int a=0; a<3; a++) if (sym[a] && lo[a] < 0) {
		const int far = MAX(-lo[a], hi[a]);
		lo[a] = hi[a] < 0 ? -hi[a] : 0;
		hi[a] = far;
	}
	alis_view_need_for(W, F, lo[0], hi[0]+1, lo[1], hi[1]+1, lo[2], hi[2]+1);
}

/* ---- resampling a graded grid onto the uniform raster (reference slices.c:25-64) --------------------- */
static float *cell_at(float ***F, const int at[3]) { return &F[at[0]][at[1]][at[2]]; }

/* one pass of the resampling: along axis `a` every grid interval [c, c+1] is spread linearly over the
 * raster points n*d it covers; axes before `a` are already on the raster, the ones after still on the grid.
 * Intervals go in ascending order: a raster point on the seam between two intervals takes the upper one's value */
static void resample_along(world W, slice S, int a, float ***from, float ***to)
{
	const int grid_lo[3] = {S->iMin, S->jMin, S->kMin}, grid_hi[3] = {S->iMax, S->jMax, S->kMax};
	const int rast_lo[3] = {S->iMIN, S->jMIN, S->kMIN}, rast_hi[3] = {S->iMAX, S->jMAX, S->kMAX};
	const float d = a == 0 ? W->dx : a == 1 ? W->dy : W->dz;
	const int u = (a+1)%3, v = (a+2)%3;
	const int u0 = u < a ? rast_lo[u] : grid_lo[u], u1 = u < a ? rast_hi[u] : grid_hi[u];
	const int v0 = v < a ? rast_lo[v] : grid_lo[v], v1 = v < a ? rast_hi[v] : grid_hi[v];

	for (int c = grid_lo[a]; c < grid_hi[a]; c++) {
		const float p0 = a == 0 ? W->itox(W, c) : a == 1 ? W->jtoy(W, c) : W->ktoz(W, c);
		const float p1 = a == 0 ? W->itox(W, c+1) : a == 1 ? W->jtoy(W, c+1) : W->ktoz(W, c+1);
		const int n0 = p0/d + (c>0 ? 1 : 0), n1 = p1/d - (c<0 ? 1 : 0);
		for (int n = n0; n <= n1; n++) {
			const float w_lower = p1 - n*d, w_upper = n*d - p0;
			int lower[3], upper[3], out[3];
			lower[a] = c, upper[a] = c+1, out[a] = n;
			for (int p = u0; p <= u1; p++) for (int q = v0; q <= v1; q++) {
				lower[u] = upper[u] = out[u] = p;
				lower[v] = upper[v] = out[v] = q;
				*cell_at(to, out) = (*cell_at(from, lower) * w_lower + *cell_at(from, upper) * w_upper) / (p1 - p0);
			}
		}
	}
}

static float ***to_uniform_raster(world W, slice S, float ***F)
{
	const int stretched[3] = { S->iMin != S->iMIN || S->iMax != S->iMAX, S->jMin != S->jMIN || S->jMax != S->jMAX, S->kMin != S->kMIN || S->kMax != S->kMAX };
	float ***spare = makeField(S->iMIN, S->iMAX, S->jMIN, S->jMAX, S->kMIN, S->kMAX);

	for (int a = 0; a < 3; a++) if (stretched[a]) {          /* ping-pong between the slice's buffer and the spare */
		resample_along(W, S, a, F, spare);
		float ***done = spare; spare = F; F = done;
	}
	removeField(spare, S->iMIN, S->jMIN, S->kMIN);
	return F;
}

/* reference slices.c:67-117 */
slice createSlice(world W, float xMin, float xMax, float yMin, float yMax, float zMin, float zMax)
{
	slice S;
	MALLOC(S, 0, 1);
	*S = (struct slice) {
		W->xtoi(W, xMin), W->xtoi(W, xMax),
		W->ytoj(W, yMin), W->ytoj(W, yMax),
		W->ztok(W, zMin), W->ztok(W, zMax),
	};
	if (W->Res.nug.x.m>1 || W->Res.nug.y.m>1 || W->Res.nug.z.m>1) S->interpolate = to_uniform_raster;

	/* clip to the grid (to the mirrored grid on symmetric axes) */
	if (S->iMax > W->iMAX) S->iMax = W->iMAX;
	if (S->jMax > W->jMAX) S->jMax = W->jMAX;
	if (S->kMax > W->kMAX) S->kMax = W->kMAX;
	if (S->iMin < W->iMIN && W->xMinSur!=SYM) S->iMin = W->iMIN;
	if (S->jMin < W->jMIN && W->yMinSur!=SYM) S->jMin = W->jMIN;
	if (S->kMin < W->kMIN && W->zMinSur!=SYM) S->kMin = W->kMIN;
	if (S->iMin <-W->iMAX && W->xMinSur==SYM) S->iMin =-W->iMAX;
	if (S->jMin <-W->jMAX && W->yMinSur==SYM) S->jMin =-W->jMAX;
	if (S->kMin <-W->kMAX && W->zMinSur==SYM) S->kMin =-W->kMAX;
	/* extent of the uniform raster the slice is resampled to on graded grids */
	const int flat_i = S->iMin==S->iMax || !S->interpolate, flat_j = S->jMin==S->jMax || !S->interpolate, flat_k = S->kMin==S->kMax || !S->interpolate;
	S->iMIN = flat_i ? S->iMin : MIN(S->iMin, W->itox(W, S->iMin)/W->dx);
	S->jMIN = flat_j ? S->jMin : MIN(S->jMin, W->jtoy(W, S->jMin)/W->dy);
	S->kMIN = flat_k ? S->kMin : MIN(S->kMin, W->ktoz(W, S->kMin)/W->dz);
	S->iMAX = flat_i ? S->iMax : MAX(S->iMax, W->itox(W, S->iMax)/W->dx);
	S->jMAX = flat_j ? S->jMax : MAX(S->jMax, W->jtoy(W, S->jMax)/W->dy);
	S->kMAX = flat_k ? S->kMax : MAX(S->kMax, W->ktoz(W, S->kMax)/W->dz);
	S->iNUM = 1-S->iMIN+S->iMAX;
	S->jNUM = 1-S->jMIN+S->jMAX;
	S->kNUM = 1-S->kMIN+S->kMAX;
	S->F = makeField(S->iMIN, S->iMAX, S->jMIN, S->jMAX, S->kMIN, S->kMAX);

	/* outline overlay of the structure for 1-D / 2-D slices once objects exist */
	if (!W->CE->x || !W->CE->y || !W->CE->z || (S->iMin!=S->iMax && S->jMin!=S->jMax && S->kMin!=S->kMax)) return S;
	float Max = 0;
	S->C = makeField(S->iMIN, S->iMAX, S->jMIN, S->jMAX, S->kMIN, S->kMAX);
	for (int i=S->iMIN; i<=S->iMAX; i++) for (int j=S->jMIN; j<=S->jMAX; j++) for (int k=S->kMIN; k<=S->kMAX; k++)
		if (Max < (S->C[i][j][k] = ContourRI(W, i, j, k))) Max = S->C[i][j][k];
	if (Max) for (int i=S->iMIN; i<=S->iMAX; i++) for (int j=S->jMIN; j<=S->jMAX; j++) for (int k=S->kMIN; k<=S->kMAX; k++)
		S->C[i][j][k] = 0.5*S->C[i][j][k]/Max;
	return S;
}

void deleteSlice(slice S)
{
	if (S->F) removeField(S->F, S->iMIN, S->jMIN, S->kMIN);
	if (S->C) removeField(S->C, S->iMIN, S->jMIN, S->kMIN);
	free(S);
}

/* ---- writers (reference slices.c:128-211) ------------------------------------------------------------------- */
void (h5)(world W, slice S, float Max, char *F, char *filename, unsigned long int* cmap)
{
	(void)Max; (void)F; (void)cmap;
	int dim[3] = {S->iMin==S->iMax?0:S->iNUM, S->jMin==S->jMax?0:S->jNUM, S->kMin==S->kMax?0:S->kNUM};
	if (S->interpolate) S->F = S->interpolate(W, S, S->F);
	h5write(dim, &S->F[S->iMIN][S->jMIN][S->kMIN], filename);
}

void (txt)(world W, slice S, float Max, char *F, char *filename, unsigned long int* cmap)
{
	(void)Max; (void)F; (void)cmap;
	int d = 0, dim[3] = {0};
	FILE *File = fopen(strcat(filename, ".txt"), "w");
	if (!File) { printf("alis-b200: cannot open %s\n", filename); exit(0); }

	if (S->iMin != S->iMax) dim[d++] = S->iNUM;
	if (S->jMin != S->jMax) dim[d++] = S->jNUM;
	if (S->kMin != S->kMax) dim[d++] = S->kNUM;
	if (d==3) printf("Can't write 3D slice to txt file.\n"), exit(0);

	if (d==1) {                                  /* coordinate, value */
		if (S->iMin != S->iMax) for (int i=S->iMin; i<=S->iMax; i++) fprintf(File, "%f\t%e\n", W->itox(W, i), S->F[i][S->jMin][S->kMin]);
		if (S->jMin != S->jMax) for (int j=S->jMin; j<=S->jMax; j++) fprintf(File, "%f\t%e\n", W->jtoy(W, j), S->F[S->iMin][j][S->kMin]);
		if (S->kMin != S->kMax) for (int k=S->kMin; k<=S->kMax; k++) fprintf(File, "%f\t%e\n", W->ktoz(W, k), S->F[S->iMin][S->jMin][k]);
	}
	else {                                       /* matrix, second slice axis downwards */
		if (S->interpolate) S->F = S->interpolate(W, S, S->F);
		float *flat = &S->F[S->iMIN][S->jMIN][S->kMIN];
		for (int j=dim[1]-1; j>=0; j--) {
			fprintf(File, "%e", flat[j]);
			for (int i=1; i<dim[0]; i++) fprintf(File, "\t%e", flat[dim[1]*i+j]);
			fprintf(File, "\n");
		}
	}
	fclose(File);
}

/* -- minimal PNG encoder: signature, IHDR, one IDAT holding a zlib stream of stored blocks, IEND -- */
static unsigned long crc_table[256];
static void crc_init(void)
{
	if (crc_table[1]) return;
	for (unsigned long n=0; n<256; n++) {
		unsigned long c = n;
		for (int k=0; k<8; k++) c = c & 1 ? 0xedb88320UL ^ (c >> 1) : c >> 1;
		crc_table[n] = c;
	}
}
static unsigned long crc_update(unsigned long c, const unsigned char *buf, size_t len)
{
	for (size_t n=0; n<len; n++) c = crc_table[(c ^ buf[n]) & 0xff] ^ (c >> 8);
	return c;
}
static void put32(unsigned char *p, unsigned long v) { p[0] = v>>24; p[1] = v>>16; p[2] = v>>8; p[3] = v; }
static void write_chunk(FILE *f, const char *type, const unsigned char *data, size_t len)
{
	unsigned char hdr[8], tail[4];
	put32(hdr, len); memcpy(hdr+4, type, 4);
	unsigned long c = crc_update(0xffffffffUL, hdr+4, 4);
	if (len) c = crc_update(c, data, len);
	put32(tail, c ^ 0xffffffffUL);
	fwrite(hdr, 1, 8, f);
	if (len) fwrite(data, 1, len, f);
	fwrite(tail, 1, 4, f);
}
static void write_png_rgb(FILE *f, int width, int height, const unsigned char *rgb)
{
	crc_init();
	static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
	fwrite(sig, 1, 8, f);
	unsigned char ihdr[13];
	put32(ihdr, width); put32(ihdr+4, height);
	ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
	write_chunk(f, "IHDR", ihdr, 13);
	const size_t stride = 1 + 3*(size_t)width, raw = stride*height;
	unsigned char *scan = malloc(raw);
	for (int y=0; y<height; y++) { scan[y*stride] = 0; memcpy(scan + y*stride + 1, rgb + 3*(size_t)width*y, 3*(size_t)width); }
	const size_t blocks = (raw + 65534)/65535;
	unsigned char *z = malloc(2 + raw + 5*blocks + 4), *q = z;
	*q++ = 0x78; *q++ = 0x01;
	unsigned long a = 1, b = 0;
	for (size_t off=0; off<raw || off==0; off+=65535) {
		const size_t n = raw - off < 65535 ? raw - off : 65535;
		*q++ = off + n >= raw;
		*q++ = n & 0xff; *q++ = n >> 8; *q++ = ~n & 0xff; *q++ = (~n >> 8) & 0xff;
		memcpy(q, scan + off, n); q += n;
		for (size_t t=0; t<n; t++) { a = (a + scan[off+t]) % 65521; b = (b + a) % 65521; }
		if (!raw) break;
	}
	put32(q, (b << 16) | a); q += 4;
	write_chunk(f, "IDAT", z, q - z);
	write_chunk(f, "IEND", 0, 0);
	free(z); free(scan);
}

/* colour-mapped image of a 1-D (drawn 64 pixels high) or 2-D slice; the structure outline S->C darkens
 * field plots.  cmap[0] = number of colour stops, negative for maps symmetric about zero. */
void (png)(world W, slice S, float Max, char *F, char *filename, unsigned long int* cmap)
{
	ALIS_PROF("png writer");
	int d = 0, dim[3] = {1,1,1};
	FILE *File = fopen(strcat(filename, ".png"), "wb");
	if (!File) { printf("alis-b200: cannot open %s\n", filename); exit(0); }
	float Min = 0;
	if (cmap[0] > LONG_MAX) Min = -Max;
	if (F[0]=='L' && F[1]=='o' && F[2]=='g' && F[3]!='R' && F[4]!='I') Min = Max-3;

	if (S->iMin != S->iMax) dim[d++] = S->iNUM;
	if (S->jMin != S->jMax) dim[d++] = S->jNUM;
	if (S->kMin != S->kMax) dim[d++] = S->kNUM;
	if (d==1) dim[2] = 64;
	if (d==3) printf("Can't write 3D slice to png file.\n"), exit(0);

	if (S->interpolate) S->F = S->interpolate(W, S, S->F);
	const int width = dim[0], height = dim[1]*dim[2], stops = abs((long)cmap[0]);
	unsigned char *image = malloc(3*(size_t)width*height), *row = image;
	const float *val = &S->F[S->iMIN][S->jMIN][S->kMIN], *out = S->C ? &S->C[S->iMIN][S->jMIN][S->kMIN] : 0;
	const int is_field = F[0]=='E' || F[0]=='H' || F[0]=='J' || F[0]=='U' || F[0]=='S' || F[0]=='i';

	for (int j=dim[1]-1; j>=0; j--) {
		for (int i=0; i<width; i++) {
			float C = !out ? 1 : 1-out[dim[1]*i+j];
			float f = (Max-Min)?MAX(0,MIN(stops-1,(stops-1)*(val[dim[1]*i+j]-Min)/(Max-Min))):0;
			int n = f;
			if (!is_field) C = 1;
			if (n > stops-2) n = stops-2;
			row[3*i]   = C * ((1-f+n) * (cmap[n+1]>>0x10&0xFF) + (f-n) * (cmap[n+2]>>0x10&0xFF));
			row[3*i+1] = C * ((1-f+n) * (cmap[n+1]>>0x08&0xFF) + (f-n) * (cmap[n+2]>>0x08&0xFF));
			row[3*i+2] = C * ((1-f+n) * (cmap[n+1]>>0x00&0xFF) + (f-n) * (cmap[n+2]>>0x00&0xFF));
		}
		for (int r=1; r<dim[2]; r++) memcpy(row + 3*(size_t)width*r, row, 3*(size_t)width);
		row += 3*(size_t)width*dim[2];
	}
	write_png_rgb(File, width, height, image);
	fclose(File);
	free(image);
}

/* ---- common argument handling -------------------------------------------------------------------------------- */
static int is_static_field(field F)
{
	return F==RI || F==LogRI || F==ContourRI || F==Contour || F==raweEffx || F==raweEffy || F==raweEffz || F==eEff;
}

/* "%"-placeholder in the file name -> "<field>-<slice>" (the macro stringifies both arguments) */
static void expand_name(char *filename, const char *f, const char *s)
{
	char rest[1024], *mark = strchr(filename, '%');
	if (!mark) return;
	strcpy(rest, mark+1);
	strcat(strcat(strcat(strcpy(mark, f), "-"), s), rest);
}

/* F over the slice as a dense block, k fastest (evaluated on the device when F has a twin there) */
static float *slice_values(world W, field F, slice S)
{
	return alis_eval_region(W, F, S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);
}

static float sample(world W, field F, slice S, float ***dst)
{
	float max = 0, *v = slice_values(W, F, S);
	size_t q = 0;
	for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++)
		if (max < fabsf(dst[i][j][k] = v[q++])) max = fabsf(dst[i][j][k]);
	free(v);
	return max;
}

/* Running sums of sliceTimeAvg / sliceFreqDom on the device: one asynchronous kernel per call instead of a
 * synchronous evaluation + device-to-host copy of the slice every step; the sums come back when a frame
 * is written.  Same float operations per point in the same order: bit-identical to the host loops. */
static alisb_accum *accum_open(world W, field F, slice S, int with_imag)
{
	alisb_field_fn d;
	int take_log;
	const int box[6] = { S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax };
	if (alis_host_monitors() || !alis_field_device_fn(W, F, &d, &take_log) || take_log || !alis_device_region_ok(W, &d, box)) return 0;
	const int32_t ib[6] = { box[0], box[1], box[2], box[3], box[4], box[5] };
	alisb_accum *a = 0;
	alis_bridge_commit(alis_host_bridge(W));
	if (alisb_accum_create(alis_bridge_engine(alis_host_bridge(W)), &d, ib, with_imag, &a) != ALISB_OK) return 0;
	return a;
}

static void accum_add(world W, alisb_accum *a, float wr, float wi)
{
	alis_bridge_commit(alis_host_bridge(W));
	if (alisb_accum_add(a, wr, wi) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
}

static void accum_fetch(slice S, alisb_accum *a, float ***F, float ***G)
{
	const size_t n = (size_t)(S->iMax-S->iMin+1)*(S->jMax-S->jMin+1)*(S->kMax-S->kMin+1);
	float *re = xmalloc(sizeof(float)*n), *im = G ? xmalloc(sizeof(float)*n) : 0;
	if (alisb_accum_read(a, re, im) != ALISB_OK) { printf("alis-b200: %s\n", alisb_last_error()); exit(1); }
	size_t q = 0;
	for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++, q++) {
		F[i][j][k] = re[q];
		if (G) G[i][j][k] = im[q];
	}
	free(re); free(im);
}

/* sliceSnap(W, F, S, N, writer[, cmap, max], "name"...): N frames per period W->T; max = 0 auto per
 * frame, max < 0 auto-scale lagging by -max periods (reference slices.c:242-279) */
void (sliceSnap)(world W, field F, slice S, char *f, char *s, ...)
{
	ALIS_PROF("sliceSnap");
	static sentry *list;
	char *format, string[1024], filename[1024];
	va_list ap;
	va_start(ap, s);
	int N = is_static_field(F) ? 0 : va_arg(ap, int);
	method M = va_arg(ap, method);
	unsigned long int *cmap = (M == png) ? va_arg(ap, unsigned long int *) : 0;
	float Max = (M == png) ? va_arg(ap, double) : 1;
	int step = (!N || N>W->T) ? 1 : W->T/N;
	int skip = Max<0 ? ((int)(-Max*W->T/step))*step : 0;
	format = va_arg(ap, char*);
	vsprintf(filename, format, ap);
	va_end(ap);
	expand_name(filename, f, s);
	sentry *e = lookup(W, filename, M, &list);
	if (N) sprintf(string, "%05ld", (e->calls-skip)/step), strcat(filename, string);

	if (e->calls % step == 0) {
		float max = sample(W, F, S, S->F);
		if (!Max) Max = max;
		else if (Max<0) {
			Max = e->Max;
			if (e->max < max) e->max = max;
			if (e->calls % skip == 0) e->Max = e->max, e->max = 0;
		}
		if (e->calls > skip) M(W, S, Max?Max:1, f, filename, cmap);
	}
}

/* average over N calls (reference slices.c:282-323) */
void (sliceTimeAvg)(world W, field F, slice S, char *f, char *s, ...)
{
	ALIS_PROF("sliceTimeAvg");
	static sentry *list;
	char *format, string[1024], filename[1024];
	va_list ap;
	va_start(ap, s);
	int N = is_static_field(F) ? 0 : va_arg(ap, int);
	method M = va_arg(ap, method);
	unsigned long int *cmap = (M == png) ? va_arg(ap, unsigned long int *) : 0;
	float Max = (M == png) ? va_arg(ap, double) : 1;
	format = va_arg(ap, char*);
	vsprintf(filename, format, ap);
	va_end(ap);
	expand_name(filename, f, s);
	sentry *e = lookup(W, filename, M, &list);
	if (N) sprintf(string, "%05ld", e->calls / N), strcat(filename, string);

	if (e->calls == 1) e->F = makeField(S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);
	if (!e->dev_tried) { e->dev_tried = 1; e->dev = accum_open(W, F, S, 0); }
	if (e->dev) accum_add(W, e->dev, 1, 0);              /* F*1 is F: the same sum */
	else {
		float *v = slice_values(W, F, S);
		size_t q = 0;
		for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++)
			e->F[i][j][k] += v[q++];
		free(v);
	}

	if (e->calls % N == 0) {
		if (e->dev) { accum_fetch(S, e->dev, e->F, 0); alisb_accum_clear(e->dev); }
		float max = 0;
		for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++)
			if (max < fabsf(S->F[i][j][k] = e->F[i][j][k] / N)) max = fabsf(S->F[i][j][k]);
		if (Max<=0) Max = max;
		M(W, S, Max?Max:1, f, filename, cmap);
		for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++)
			e->F[i][j][k] = 0;
	}
}

/* running DFT at wavelength L over N calls, written as 12 phase frames (reference slices.c:326-379) */
void (sliceFreqDom)(world W, field F, slice S, char *f, char *s, int N, float L, ...)
{
	ALIS_PROF("sliceFreqDom");
	field *linear[30] = {Ex, Ey, Ez, Hx, Hy, Hz, iEx, iEy, iEz, iHx, iHy, iHz, rawEx, rawEy, rawEz, rawHx, rawHy, rawHz,
	                     Jx, Jy, Jz, rawJx, rawJy, rawJz, ScEx, ScEy, ScEz, SoEx, SoEy, SoEz};
	int known = 0;
	for (int q=0; q<30; q++) if (*F == linear[q]) known = 1;
	if (!known) printf("\nThe %s field is not supported in sliceFreqDom().\n", f), exit(0);

	static sentry *list;
	char *format, string[1024], filename[1024];
	va_list ap;
	va_start(ap, L);
	method M = va_arg(ap, method);
	unsigned long int *cmap = (M == png) ? va_arg(ap, unsigned long int *) : 0;
	float Max = (M == png) ? va_arg(ap, double) : 1;
	format = va_arg(ap, char*);
	vsprintf(filename, format, ap);
	va_end(ap);
	expand_name(filename, f, s);
	sentry *e = lookup(W, filename, M, &list);

	if (e->calls == 1) {
		e->F = makeField(S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);
		e->G = makeField(S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax);
	}
	float real = cosf(2 * PI * e->calls * W->dt / L);
	float imag = sinf(2 * PI * e->calls * W->dt / L);
	if (!e->dev_tried) { e->dev_tried = 1; e->dev = accum_open(W, F, S, 1); }
	if (e->dev) accum_add(W, e->dev, real, imag);
	else {
		float *v = slice_values(W, F, S);
		size_t q = 0;
		for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++, q++) {
			e->F[i][j][k] += v[q] * real;
			e->G[i][j][k] += v[q] * imag;
		}
		free(v);
	}

	if (e->calls % N == 0) {
		if (e->dev) accum_fetch(S, e->dev, e->F, e->G);
		if (Max <= 0) {
			float max = 0;
			for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++)
				if (max < sq(e->F[i][j][k])+sq(e->G[i][j][k])) max = sq(e->F[i][j][k])+sq(e->G[i][j][k]);
			if (Max < sqrtf(max) * 2 / e->calls) Max = sqrtf(max) * 2 / e->calls;
		}
		for (int n=0; n<12; n++) {
			for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++)
				S->F[i][j][k] = (e->F[i][j][k] * cosf(n * PI / 6) + e->G[i][j][k] * sinf(n * PI / 6)) * 2 / e->calls;
			sprintf(string, "%s%02d", filename, n+1);
			M(W, S, Max, f, string, cmap);
		}
	}
}

/* ---- reductions over a slice (reference slices.c:382-427) ------------------------------------------------------ */
float sliceMax(world W, field F, slice S)
{
	ALIS_PROF("sliceMax");
	float best = 0, f;
	const int box[6] = { S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax };
	if (alis_device_absmax(W, F, box, &best)) return best;
	prefetch(W, F, S);
	for (int i=S->iMin; i<=S->iMax; i++) for (int j=S->jMin; j<=S->jMax; j++) for (int k=S->kMin; k<=S->kMax; k++)
		if (best < (f=fabsf(F(W, i, j, k)))) best = f;
	return best;
}

static float line_sum(world W, field F, slice S, int i, int j)
{
	float s = 0;
	for (int k=S->kMin+1; k<S->kMax; k++) s += W->CE->dz[k] * F(W, i, j, k);
	s += 0.5 * W->CE->dz[S->kMin] * F(W, i, j, S->kMin);
	s += 0.5 * W->CE->dz[S->kMax] * F(W, i, j, S->kMax);
	return s;
}

static float sheet_sum(world W, field F, slice S, int i)
{
	float s = 0;
	for (int j=S->jMin+1; j<S->jMax; j++) s += W->CE->dy[j] * line_sum(W, F, S, i, j);
	s += 0.5 * W->CE->dy[S->jMin] * line_sum(W, F, S, i, S->jMin);
	s += 0.5 * W->CE->dy[S->jMax] * line_sum(W, F, S, i, S->jMax);
	return s;
}

/* the same with the z column sums already formed on the device: cols[(i-iMin)*nj + (j-jMin)] */
static float sheet_sum_cols(world W, slice S, const float *cols, int i)
{
	const int nj = S->jMax-S->jMin+1;
	const float *c = cols + (size_t)(i-S->iMin)*nj - S->jMin;
	float s = 0;
	for (int j=S->jMin+1; j<S->jMax; j++) s += W->CE->dy[j] * c[j];
	s += 0.5 * W->CE->dy[S->jMin] * c[S->jMin];
	s += 0.5 * W->CE->dy[S->jMax] * c[S->jMax];
	return s;
}

float sliceSum(world W, field F, slice S)
{
	ALIS_PROF("sliceSum");
	float Sum = 0;
	const int box[6] = { S->iMin, S->iMax, S->jMin, S->jMax, S->kMin, S->kMax };
	float *cols = xmalloc(sizeof(float)*(size_t)(S->iMax-S->iMin+1)*(S->jMax-S->jMin+1));
	if (cols && alis_device_column_sums(W, F, box, cols)) {
		for (int i=S->iMin+1; i<S->iMax; i++) Sum += W->CE->dx[i] * sheet_sum_cols(W, S, cols, i);
		Sum += 0.5 * W->CE->dx[S->iMin] * sheet_sum_cols(W, S, cols, S->iMin);
		Sum += 0.5 * W->CE->dx[S->iMax] * sheet_sum_cols(W, S, cols, S->iMax);
		free(cols);
	}
	else {
	free(cols);
	prefetch(W, F, S);
	for (int i=S->iMin+1; i<S->iMax; i++) Sum += W->CE->dx[i] * sheet_sum(W, F, S, i);
	Sum += 0.5 * W->CE->dx[S->iMin] * sheet_sum(W, F, S, S->iMin);
	Sum += 0.5 * W->CE->dx[S->iMax] * sheet_sum(W, F, S, S->iMax);
	}
	if (S->iMin==S->iMax) Sum /= W->CE->dx[S->iMin];
	if (S->jMin==S->jMax) Sum /= W->CE->dy[S->jMin];
	if (S->kMin==S->kMax) Sum /= W->CE->dz[S->kMin];
	return Sum;
}
