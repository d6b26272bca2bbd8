/* TEST INFRASTRUCTURE — stub of <hdf5.h>, only so that the reference's
 * libvvhd/src/TSpace.cpp (which mixes the HDF5 file open/close calls with the
 * force bookkeeping we need: calc_forces/zero_forces/EnumerateBodies) compiles
 * unmodified in an image that has no HDF5. None of these entry points is ever
 * reached by the oracle drivers; ref_stubs.cpp defines them to abort(). */
#pragma once
#include <stdint.h>
typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
#define H5E_DEFAULT 0
#define H5P_DEFAULT 0
#define H5F_ACC_RDONLY 0u
#define H5F_ACC_TRUNC 2u
#ifdef __cplusplus
extern "C" {
#endif
herr_t H5Eset_auto(hid_t, void*, void*);
htri_t H5Fis_hdf5(const char*);
hid_t H5Fopen(const char*, unsigned, hid_t);
hid_t H5Fcreate(const char*, unsigned, hid_t, hid_t);
herr_t H5Fclose(hid_t);
#ifdef __cplusplus
}
#endif
