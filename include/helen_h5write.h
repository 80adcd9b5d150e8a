/*
 * helen_h5write.h -- C ABI of the host-side prediction-file writer of the B200-native HELEN call_consensus path
 * (part of helen_b200/lib/libhelen_feed.so, host C++, no libhdf5).
 *
 * Replaces what h5py does under helen/modules/python/DataStore.py:83-133 (write_prediction: three datasets and two
 * scalars per image, in nested groups): a classic HDF5 file any libhdf5 reads - superblock 0, version-1 object headers,
 * symbol-table groups (local heap + SNOD nodes + version-1 B-tree), contiguous datasets of little-endian integers,
 * IEEE floats or fixed-length byte strings.  Raw data goes to the file as each dataset is handed over; the group
 * structure is kept in memory and written by hw_close.  The bytes are those helen_b200/minih5.py's own writer produces
 * for the same sequence of calls (tests/test_h5write_native.py compares whole files), so the two are interchangeable.
 */
#ifndef HELEN_H5WRITE_H
#define HELEN_H5WRITE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HW_ABI_VERSION 1

enum hw_status {
    HW_OK = 0,
    HW_E_EXISTS = 1,      /* "Unable to create dataset (name already exists)", or a path component that is a dataset */
    HW_E_TYPE = 2,        /* a type the writer does not store */
    HW_E_IO = 3,
    HW_E_ARGUMENT = 4
};

typedef struct hw_file hw_file;

int hw_abi_version(void);
int hw_create(const char *path, hw_file **out, char *err, int errlen);

/* One dataset at `path` ("group/sub/name"; missing groups are created).  kind: 'i' / 'u' integers, 'f' IEEE floats of 4 or
 * 8 bytes, 'S' fixed-length byte strings; little-endian, C-contiguous data of dims[0] x ... x dims[rank-1] elements
 * (rank 0: a scalar). */
int hw_dataset(hw_file *file, const char *path, char kind, int itemsize, int rank, const uint64_t *dims, const void *data,
               char *err, int errlen);
/* n datasets parents[i] + "/" + name, dataset i = row i of a C-contiguous [n, row_dims...] array; the array is written in
 * one piece.  `parents` is n '\0'-terminated strings back to back. */
int hw_rows(hw_file *file, const char *parents, int64_t n, const char *name, char kind, int itemsize, int row_rank,
            const uint64_t *row_dims, const void *data, char *err, int errlen);
/* 1 if `path` names a group or dataset written so far. */
int hw_contains(const hw_file *file, const char *path);
/* Writes the group structure and the superblock, closes the file and frees the handle (also after an error). */
int hw_close(hw_file *file, char *err, int errlen);

#ifdef __cplusplus
}
#endif
#endif
