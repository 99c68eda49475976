/* Test-infrastructure shim (NOT product code): lets the UNMODIFIED reference's world.c compile and
 * run its HDF5 paths (h5read/h5write, reference world.c:469-537; through them the `h5` slice
 * writer and guidedWave*) in an image without libhdf5.  The dozen H5* calls the reference makes
 * are mapped onto this repo's minimal HDF5 reader/writer (alis_b200/host/h5lite.c, compiled into
 * oracle/_ref by build_ref.sh), so files written by the reference twin and by our library are
 * interchangeable.  No reference source is modified. */
#ifndef ORACLE_SHIM_HDF5_H
#define ORACLE_SHIM_HDF5_H
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include "../../alis_b200/host/h5lite.h"
typedef long hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;
enum { H5F_ACC_RDONLY = 0, H5F_ACC_TRUNC = 0, H5P_DEFAULT = 0, H5S_ALL = 0, H5T_NATIVE_FLOAT = 0 };

typedef struct { char path[1024]; int rank; unsigned long long dims[3]; float *data; } oracle_h5_handle;
static oracle_h5_handle oracle_h5_table[16];
static inline hid_t oracle_h5_new(void)
{
	for (int q=1; q<16; q++) if (!oracle_h5_table[q].path[0]) { memset(&oracle_h5_table[q], 0, sizeof oracle_h5_table[q]); oracle_h5_table[q].path[0] = '?'; return q; }
	fprintf(stderr, "oracle shim: out of HDF5 handles\n"); abort();
	return 0;
}
static inline hid_t H5Fopen(const char *name, int flags, int plist)
{
	(void)flags; (void)plist;
	char err[200];
	hid_t h = oracle_h5_new();
	oracle_h5_handle *H = &oracle_h5_table[h];
	snprintf(H->path, sizeof H->path, "%s", name);
	H->data = h5lite_read(name, &H->rank, H->dims, err, sizeof err);
	if (!H->data) { fprintf(stderr, "oracle shim: H5Fopen(%s): %s\n", name, err); abort(); }
	return h;
}
static inline hid_t H5Fcreate(const char *name, int flags, int a, int b)
{
	(void)flags; (void)a; (void)b;
	hid_t h = oracle_h5_new();
	snprintf(oracle_h5_table[h].path, sizeof oracle_h5_table[h].path, "%s", name);
	return h;
}
static inline herr_t H5Fclose(hid_t h) { free(oracle_h5_table[h].data); memset(&oracle_h5_table[h], 0, sizeof oracle_h5_table[h]); return 0; }
static inline long H5Gget_objname_by_idx(hid_t file, int idx, char *name, size_t size) { (void)file; (void)idx; snprintf(name, size, "data"); return 4; }
static inline hid_t H5Dopen(hid_t file, const char *name, int plist) { (void)name; (void)plist; return file; }
static inline hid_t H5Dget_space(hid_t dataset) { return dataset; }
static inline int H5Sget_simple_extent_ndims(hid_t space) { return oracle_h5_table[space].rank; }
static inline int H5Sget_simple_extent_dims(hid_t space, hsize_t *dims, hsize_t *maxdims)
{
	for (int q=0; q<oracle_h5_table[space].rank; q++) { dims[q] = oracle_h5_table[space].dims[q]; if (maxdims) maxdims[q] = dims[q]; }
	return oracle_h5_table[space].rank;
}
static inline herr_t H5Dread(hid_t dataset, int type, int ms, int fs, int plist, void *buf)
{
	(void)type; (void)ms; (void)fs; (void)plist;
	oracle_h5_handle *H = &oracle_h5_table[dataset];
	size_t n = 1;
	for (int q=0; q<H->rank; q++) n *= H->dims[q];
	memcpy(buf, H->data, n*sizeof(float));
	return 0;
}
static inline hid_t H5Screate_simple(int rank, const hsize_t *dims, const hsize_t *maxdims)
{
	(void)maxdims;
	hid_t h = oracle_h5_new();
	oracle_h5_table[h].rank = rank;
	for (int q=0; q<rank; q++) oracle_h5_table[h].dims[q] = dims[q];
	return h;
}
static inline hid_t H5Dcreate(hid_t file, const char *name, int type, hid_t space, int a, int b, int c)
{
	(void)name; (void)type; (void)a; (void)b; (void)c;
	oracle_h5_table[file].rank = oracle_h5_table[space].rank;
	memcpy(oracle_h5_table[file].dims, oracle_h5_table[space].dims, sizeof oracle_h5_table[file].dims);
	return file;
}
static inline herr_t H5Dwrite(hid_t dataset, int type, int ms, int fs, int plist, const void *buf)
{
	(void)type; (void)ms; (void)fs; (void)plist;
	oracle_h5_handle *H = &oracle_h5_table[dataset];
	if (h5lite_write(H->path, H->rank, H->dims, (const float*)buf)) { fprintf(stderr, "oracle shim: cannot write %s\n", H->path); abort(); }
	return 0;
}
static inline herr_t H5Dclose(hid_t dataset) { (void)dataset; return 0; }       /* the file handle owns the state */
static inline herr_t H5Sclose(hid_t space) { if (oracle_h5_table[space].path[0] == '?') memset(&oracle_h5_table[space], 0, sizeof oracle_h5_table[space]); return 0; }
#endif
