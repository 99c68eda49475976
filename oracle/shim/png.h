/* Test-infrastructure shim (NOT product code): lets the reference's slices.c compile without libpng
 * and still WRITE its images, so that the slice pictures of the unmodified reference can be compared
 * pixel by pixel with this repo's (tests/helpers.py:read_png).  Only the dozen libpng calls the
 * reference's png() makes (reference slices.c:167-211) exist; rows are collected and written in
 * png_write_end as an 8-bit RGB image whose zlib stream consists of stored (uncompressed) blocks. */
#ifndef ORACLE_SHIM_PNG_H
#define ORACLE_SHIM_PNG_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
typedef unsigned char png_byte;
typedef struct shim_png { FILE *file; unsigned w, h, rows; png_byte *pix; } *png_structp;
typedef void *png_infop;
typedef struct { int compression; char *key; char *text; } png_text;
#define PNG_LIBPNG_VER_STRING        "oracle-shim"
#define PNG_Z_DEFAULT_COMPRESSION    (-1)
#define PNG_COLOR_TYPE_RGB           2
#define PNG_INTERLACE_NONE           0
#define PNG_COMPRESSION_TYPE_DEFAULT 0
#define PNG_FILTER_TYPE_DEFAULT      0
#define PNG_TEXT_COMPRESSION_NONE    (-1)

static png_structp png_create_write_struct(const char *ver, void *e, void *f, void *w)
{
	(void)ver; (void)e; (void)f; (void)w;
	return (png_structp)calloc(1, sizeof(struct shim_png));
}
#define png_create_info_struct(p)        ((png_infop)(p))
#define png_init_io(p, f)                ((p)->file = (f))
#define png_set_compression_level(p, l)  ((void)0)
#define png_set_text(...)                ((void)0)
#define png_write_info(p, i)             ((void)0)
static void png_set_IHDR(png_structp p, png_infop info, unsigned w, unsigned h, int depth, int colour, int interlace, int comp, int filter)
{
	(void)info; (void)interlace; (void)comp; (void)filter;
	if (depth != 8 || colour != PNG_COLOR_TYPE_RGB) { fprintf(stderr, "png shim: only 8-bit RGB\n"); abort(); }
	p->w = w; p->h = h; p->rows = 0;
	p->pix = (png_byte*)malloc((size_t)3*w*h + 1);
}
static void png_write_row(png_structp p, const png_byte *row)
{
	if (p->rows < p->h) memcpy(p->pix + (size_t)3*p->w*p->rows++, row, (size_t)3*p->w);
}
static unsigned long shim_png_crc(unsigned long c, const png_byte *b, size_t n)
{
	while (n--) { c ^= *b++; for (int k=0; k<8; k++) c = (c >> 1) ^ (0xedb88320UL & (0UL - (c & 1))); }
	return c;
}
static void shim_png_be32(png_byte *q, unsigned long v) { q[0] = (png_byte)(v>>24); q[1] = (png_byte)(v>>16); q[2] = (png_byte)(v>>8); q[3] = (png_byte)v; }
static void shim_png_chunk(FILE *f, const char *tag, const png_byte *d, size_t n)
{
	png_byte b[4];
	unsigned long c = shim_png_crc(0xffffffffUL, (const png_byte*)tag, 4);
	if (n) c = shim_png_crc(c, d, n);
	shim_png_be32(b, (unsigned long)n); fwrite(b, 1, 4, f); fwrite(tag, 1, 4, f);
	if (n) fwrite(d, 1, n, f);
	shim_png_be32(b, c ^ 0xffffffffUL); fwrite(b, 1, 4, f);
}
static void png_write_end(png_structp p, png_infop info)
{
	(void)info;
	static const png_byte magic[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
	png_byte hd[13];
	fwrite(magic, 1, 8, p->file);
	shim_png_be32(hd, p->w); shim_png_be32(hd + 4, p->h); hd[8] = 8; hd[9] = 2; hd[10] = hd[11] = hd[12] = 0;
	shim_png_chunk(p->file, "IHDR", hd, 13);
	/* zlib stream: header, one stored block per scanline (filter byte 0 + RGB), adler32 */
	const size_t line = 1 + (size_t)3*p->w;
	if (line > 65535) { fprintf(stderr, "png shim: image too wide\n"); abort(); }
	png_byte *z = (png_byte*)malloc(2 + (line + 5)*(size_t)p->h + 5 + 4), *q = z;
	unsigned long s1 = 1, s2 = 0;
	*q++ = 0x78; *q++ = 0x01;
	for (unsigned y=0; y<p->h; y++) {
		*q++ = (y + 1 == p->h); *q++ = (png_byte)(line & 255); *q++ = (png_byte)(line >> 8); *q++ = (png_byte)(~line & 255); *q++ = (png_byte)((~line >> 8) & 255);
		png_byte *s = q;
		*q++ = 0; memcpy(q, p->pix + (size_t)3*p->w*y, (size_t)3*p->w); q += (size_t)3*p->w;
		for (; s < q; s++) { s1 = (s1 + *s) % 65521; s2 = (s2 + s1) % 65521; }
	}
	if (!p->h) { *q++ = 1; *q++ = 0; *q++ = 0; *q++ = 0xff; *q++ = 0xff; }
	shim_png_be32(q, (s2 << 16) | s1); q += 4;
	shim_png_chunk(p->file, "IDAT", z, (size_t)(q - z));
	shim_png_chunk(p->file, "IEND", 0, 0);
	free(z);
}
static void png_destroy_write_struct(png_structp *p, png_infop *i)
{
	(void)i;
	if (p && *p) { free((*p)->pix); free(*p); *p = 0; }
}
#endif
