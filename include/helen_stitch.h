/* helen_stitch.h — C ABI of the host-side stitch library.
 *
 * The step immediately after the GPU hot path: decode the predicted (base, run-length) labels of one
 * region into a sequence, and join adjacent regions on an anchor found by a local alignment.  It is
 * CPU work (inputs are a few hundred bases per call, Stitch.py:122-134), so this is a plain C++17 host
 * library (helen_b200/csrc_host/stitch_host.cpp -> helen_b200/lib/libhelen_stitch.so), no CUDA.
 *
 * What each entry point replaces in the reference (kishwarshafin/helen @ a075e9f):
 *   hs_ssw_align           HELEN.Aligner(...).SetReferenceSequence + Align_cpp with a default Filter and
 *                          maskLen 0 (pybind_api.h:41-46, ssw_cpp.cpp:320-352, ssw.c:801-887), as called
 *                          from Stitch.py:110-135.  Same score, begin/end positions and =/X/I/D/S cigar
 *                          string, including the reference's tie-breaking.
 *   hs_anchor_from_cigar   Stitch.get_confident_positions (Stitch.py:34-94)
 *   hs_decode_region       the position dictionary + label decoding of Stitch.small_chunk_stitch
 *                          (Stitch.py:214-245; label_decoder Options.py:3)
 *   hs_stitcher_*          Stitch.alignment_stitch (Stitch.py:96-193)
 *
 * All functions return 0 on success and a negative HS_E_* code on failure unless stated otherwise;
 * hs_last_error() describes the last failure of the calling thread.
 */
#ifndef HELEN_STITCH_H
#define HELEN_STITCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_ABI_VERSION 1

#define HS_E_ARGUMENT (-1)   /* null pointer / negative size / scoring without gap_open > gap_extend >= 0, match > 0 */
#define HS_E_CAPACITY (-2)   /* output buffer too small */
#define HS_E_RANGE    (-3)   /* score beyond the reference's 16-bit kernel */
#define HS_E_TRACE    (-4)   /* the banded traceback left the band (the reference prints "Trace back error") */
#define HS_E_CIGAR    (-5)   /* cigar operation other than M = X I D S */

/* Options.py:4-7: MATCH_PENALTY 4, MISMATCH_PENALTY 6, GAP_PENALTY 8, GAP_EXTEND_PENALTY 2 */
typedef struct hs_scoring {
    int32_t match;        /* added for equal A/C/G/T */
    int32_t mismatch;     /* subtracted otherwise (any pair involving a non-ACGT letter is a mismatch) */
    int32_t gap_open;     /* subtracted for the first base of a gap */
    int32_t gap_extend;   /* subtracted for each further base */
} hs_scoring;

/* the fields of StripedSmithWaterman::Alignment (ssw_cpp.h:14-24) that Align_cpp fills */
typedef struct hs_alignment {
    int32_t score;        /* best_score */
    int32_t ref_begin;    /* reference_begin, 0-based */
    int32_t ref_end;      /* reference_end, 0-based inclusive */
    int32_t query_begin;
    int32_t query_end;
    int32_t mismatches;   /* mismatching + inserted + deleted bases */
    int32_t cigar_len;    /* strlen of the cigar string */
    int32_t kernel;       /* 8 or 16: which of the reference's two striped kernels the result follows */
} hs_alignment;

int32_t hs_abi_version(void);
const char* hs_last_error(void);

/* Local alignment of `query` against `ref`.  `cigar` receives the NUL-terminated cigar string.
 * An empty ref or query, or a pair without a single matching base, gives score 0 and an empty cigar
 * (Stitch.py:138 tests best_score == 0 before anything else is read). */
int32_t hs_ssw_align(const char* ref, int32_t ref_len, const char* query, int32_t query_len,
                     const hs_scoring* scoring, hs_alignment* out, char* cigar, int32_t cigar_cap);

/* First run of at least `min_run` aligned bases (= and X count alike, adjacent runs are merged).
 * On return *ref_pos / *query_pos are the positions where it starts, or -1 / -1 (Stitch.py:94). */
int32_t hs_anchor_from_cigar(const char* cigar, int32_t ref_begin, int32_t min_run,
                             int32_t* ref_pos, int32_t* query_pos);

/* positions: n rows of (position, insert index, split index) as stored by the prediction writer, in the
 * order the reference visits them (chunks sorted by name, rows in order).  Rows with position < 0 or
 * index < 0 are skipped, the FIRST prediction of a key is kept, keys are sorted, and every kept row
 * contributes label_decoder[base] x rle.  Returns the sequence length (excluding the NUL written after it). */
int64_t hs_decode_region(const int64_t* positions, const uint8_t* bases, const uint8_t* rles, int64_t n,
                         char* out, int64_t out_cap);

/* Stitcher: add (start, end, sequence) pieces in any order, run, read the result. */
typedef struct hs_stitcher hs_stitcher;

#define HS_WARN_NO_ALIGNMENT  0   /* Stitch.py:139 */
#define HS_WARN_NO_ANCHOR     1   /* Stitch.py:156 */
#define HS_WARN_NO_OVERLAP    2   /* Stitch.py:186 */

hs_stitcher* hs_stitcher_create(const hs_scoring* scoring, int32_t overlap_threshold, double base_error_rate);
void hs_stitcher_destroy(hs_stitcher* s);
int32_t hs_stitcher_add(hs_stitcher* s, int64_t start, int64_t end, const char* sequence, int64_t length);
/* Stable-sorts the pieces by (start, end) and joins them.  warnings[3] (may be null) counts the three
 * warning cases above; alignments (may be null) counts the local alignments performed. */
int32_t hs_stitcher_run(hs_stitcher* s, int64_t* start, int64_t* end, int64_t* length,
                        int64_t* warnings, int64_t* alignments);
int64_t hs_stitcher_sequence(const hs_stitcher* s, char* out, int64_t out_cap);

#ifdef __cplusplus
}
#endif
#endif
