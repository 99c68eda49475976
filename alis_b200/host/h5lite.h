/*
 * h5lite.h -- minimal HDF5 reader/writer (one float dataset of rank 1..3 per file), see h5lite.c.
 * Internal header of the host layer (also used by the oracle's hdf5.h shim so that the unmodified
 * reference can run its HDF5-dependent paths in an image without libhdf5).
 */
#ifndef ALIS_H5LITE_H
#define ALIS_H5LITE_H
#include <stddef.h>

/* writes dataset "data" (float32, contiguous) into a new file; 0 on success */
int   h5lite_write(const char *path, int rank, const unsigned long long *dims, const float *data);
/* reads the first dataset of the root group as float32 (malloc'd, row-major); NULL + message on error */
void *h5lite_read(const char *path, int *rank, unsigned long long dims[3], char *err, size_t errlen);

#endif
