/*
 * helen_feed.h -- C ABI of the host-side input feed of the B200-native HELEN call_consensus path.
 *
 * Replaces the per-image work of the reference's reader,
 *   helen/modules/python/models/dataloader_predict.py:54-88  (SequenceDataset.__getitem__: open the HDF5 file, read
 *   images/<name>/{contig, contig_start, contig_end, feature_chunk_idx, image, position}, pad to 1000 columns)
 * and the DataLoader's collation of batch_size such items: one call fills the arrays of a whole batch straight from the
 * memory-mapped file, with several host threads.  No libhdf5: the library parses the subset of the HDF5 file format that
 * libhdf5 writes by default for such files (superblock 0-3, version 1 / 2 object headers, symbol-table or compact-link
 * groups, contiguous or compact datasets of little-endian integers and fixed- or variable-length strings).  Anything
 * else (chunked / filtered datasets, dense groups, big-endian data) is answered with HF_UNSUPPORTED and the caller falls
 * back to its general reader (h5py, or helen_b200/minih5.py); results are identical either way (tests/test_feed_native.py).
 *
 * Host code only (no CUDA).  Thread safety: a handle may be read from several threads at once.
 */
#ifndef HELEN_FEED_H
#define HELEN_FEED_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HF_ABI_VERSION 2

enum hf_status {
    HF_OK = 0,
    HF_UNSUPPORTED = 1,   /* a valid file that uses a feature outside the subset: use the general reader */
    HF_E_SIZE = 2,        /* "IMAGE SIZE ERROR": more columns than seq_len, a feature count that differs from the block's,
                             or position rows != image rows (dataloader_predict.py:85-86 raises ValueError) */
    HF_E_FORMAT = 3,      /* not an HDF5 file, truncated, or inconsistent */
    HF_E_ARGUMENT = 4
};

typedef struct hf_file hf_file;

int hf_abi_version(void);

/* Maps the file and lists the members of /images in the order the file stores them (the order h5py's keys() gives for
 * a symbol-table group).  A file without /images opens with zero images (the reference warns and skips it). */
int hf_open(const char *path, hf_file **out, char *err, int errlen);
void hf_close(hf_file *file);

int64_t hf_image_count(const hf_file *file);
/* Names of all images, each terminated by '\0', in file order.  Returns the bytes needed; fills `buf` if it is large enough. */
int64_t hf_image_names(const hf_file *file, char *buf, int64_t buflen);
/* Feature count (second dimension of images/<name>/image) of image i. */
int hf_image_features(const hf_file *file, int64_t i, int *features, char *err, int errlen);

/* Images [first, first + count) of the file:
 *   images    u8 [count, seq_len, features]   rows past an image's own length are zero
 *   position  i64[count, seq_len, 3]          rows past an image's own length are (-1, -1, -1)
 *   contig_start, contig_end, chunk_id  i64[count]
 *   contigs   char[count][contig_stride]      '\0'-terminated contig names (apostrophes removed, as the reference does)
 * `threads` host threads share the images (<= 0: one). */
int hf_read_block(const hf_file *file, int64_t first, int64_t count, int seq_len, int features,
                  uint8_t *images, int64_t *position, int64_t *contig_start, int64_t *contig_end, int64_t *chunk_id,
                  char *contigs, int contig_stride, int threads, char *err, int errlen);

/* Prediction files, for the stitch (helen/modules/python/Stitch.py:214-245: per region, the position / bases / rles rows
 * of all its chunks): the rows of predictions/<contig>/<region>/<chunk>/{position, bases, rles} of every chunk, chunks in
 * string order of their names, back to back: position i64[rows, 3] (the file's uint32 widened), bases u8[rows], rles u8[rows].
 * The region is found by searching the groups' B-trees, not by listing them (a contig has one member per region).
 * *total_rows is always the region's row count; the arrays are filled only if capacity_rows >= *total_rows.
 * HF_UNSUPPORTED: no such region in this file's `predictions` group (e.g. a packed file), or data outside the subset. */
/* The first pass of the stitch (StitchInterface.py:52-66): contig == NULL lists the contigs of `predictions`; otherwise the
 * regions of that contig in file order with their contig_start / contig_end datasets.  Names '\0'-terminated back to back;
 * *count and *names_needed are always set, the arrays filled when they are large enough. */
int hf_list_predictions(const hf_file *file, const char *contig, char *names, int64_t names_len, int64_t *starts, int64_t *ends,
                        int64_t max_entries, int64_t *count, int64_t *names_needed, char *err, int errlen);
int hf_read_prediction_region(const hf_file *file, const char *contig, const char *region, int64_t capacity_rows,
                              int64_t *position, uint8_t *bases, uint8_t *rles, int64_t *total_rows, char *err, int errlen);

#ifdef __cplusplus
}
#endif
#endif
