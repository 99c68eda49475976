/* Test-infrastructure stub (NOT product code): lets the reference's slices.c compile
 * without libpng; the png() writer (reference slices.c:167-211) becomes a no-op. */
#ifndef ORACLE_SHIM_PNG_H
#define ORACLE_SHIM_PNG_H
#include <stdio.h>
#include <limits.h>
typedef void *png_structp;
typedef void *png_infop;
typedef unsigned char png_byte;
typedef struct { int compression; char *key; char *text; } png_text;
#define PNG_LIBPNG_VER_STRING        "oracle-shim"
#define PNG_Z_DEFAULT_COMPRESSION    (-1)
#define PNG_COLOR_TYPE_RGB           2
#define PNG_INTERLACE_NONE           0
#define PNG_COMPRESSION_TYPE_DEFAULT 0
#define PNG_FILTER_TYPE_DEFAULT      0
#define PNG_TEXT_COMPRESSION_NONE    (-1)
#define png_create_write_struct(v, e, f, w) ((png_structp)1)
#define png_create_info_struct(p)           ((png_infop)1)
#define png_init_io(p, f)                   ((void)0)
#define png_set_compression_level(p, l)     ((void)0)
#define png_set_IHDR(...)                   ((void)0)
#define png_set_text(...)                   ((void)0)
#define png_write_info(p, i)                ((void)0)
#define png_write_row(p, r)                 ((void)0)
#define png_write_end(p, i)                 ((void)0)
#define png_destroy_write_struct(p, i)      ((void)0)
#endif
