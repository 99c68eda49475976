/* Test-infrastructure stub (NOT product code): lets the reference's world.c compile
 * without libhdf5.  Only h5read/h5write (reference world.c:469-537) touch these; they
 * abort if ever reached.  No reference source is modified. */
#ifndef ORACLE_SHIM_HDF5_H
#define ORACLE_SHIM_HDF5_H
#include <stdlib.h>
#include <stdio.h>
typedef long hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;
enum { H5F_ACC_RDONLY = 0, H5F_ACC_TRUNC = 0, H5P_DEFAULT = 0, H5S_ALL = 0, H5T_NATIVE_FLOAT = 0 };
static inline hid_t oracle_shim_no_hdf5(void)
{
	fprintf(stderr, "oracle shim: HDF5 is not available in this image\n");
	abort();
	return 0;
}
#define H5Fopen(...)                     oracle_shim_no_hdf5()
#define H5Fcreate(...)                   oracle_shim_no_hdf5()
#define H5Fclose(...)                    oracle_shim_no_hdf5()
#define H5Gget_objname_by_idx(...)       oracle_shim_no_hdf5()
#define H5Dopen(...)                     oracle_shim_no_hdf5()
#define H5Dcreate(...)                   oracle_shim_no_hdf5()
#define H5Dget_space(...)                oracle_shim_no_hdf5()
#define H5Dread(...)                     oracle_shim_no_hdf5()
#define H5Dwrite(...)                    oracle_shim_no_hdf5()
#define H5Dclose(...)                    oracle_shim_no_hdf5()
#define H5Screate_simple(...)            oracle_shim_no_hdf5()
#define H5Sget_simple_extent_ndims(...)  ((int)oracle_shim_no_hdf5())
#define H5Sget_simple_extent_dims(...)   oracle_shim_no_hdf5()
#define H5Sclose(...)                    oracle_shim_no_hdf5()
#endif
