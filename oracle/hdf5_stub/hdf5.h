/* Test infrastructure only. Minimal stand-in for <hdf5.h>: libhdf5 is not
 * installed in this image, and the trajectory loop of the reference only needs
 * the three scalar typedefs below to compile (reference simulation.c:5,:444;
 * h5file.h:6-8,:15-46). No HDF5 functionality is provided. */
#ifndef HEROAM_ORACLE_HDF5_STUB_H
#define HEROAM_ORACLE_HDF5_STUB_H
#include <stdint.h>
typedef int64_t hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
#endif
