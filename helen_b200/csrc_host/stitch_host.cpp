// Host-side stitch library (include/helen_stitch.h).
//
// The reference joins adjacent regions with the Complete-Striped-Smith-Waterman library
// (helen/modules/src/local_reassembly/ssw.c, ssw_cpp.cpp) called from Stitch.py:96-193.  The inputs
// are a few hundred bases, so nothing here needs SIMD striping or a GPU; what matters for a drop-in is
// that the SAME alignment comes out, including which of several equally scoring alignments is reported.
// That depends on three details of the reference that a textbook Smith-Waterman does not share, all
// reproduced below with plain integer dynamic programming:
//
//  1. Cell values.  ssw.c keeps its columns in interleaved segments (16 lanes in the 8-bit kernel, 8 in
//     the 16-bit one, ssw.c:181,409), carries the vertical-gap value F inside a segment in the main loop
//     and across segments in a "lazy F" fix-up (ssw.c:264-293, 486-498), and computes the horizontal-gap
//     value E of the next column BEFORE that fix-up (ssw.c:250-253), so a horizontal gap may directly
//     follow a vertical one only when that one started in the same segment.  For gap_open > gap_extend
//     (required here) this changes no cell: the same two gaps in the other order reach the same cell at
//     the same cost and that order is never restricted.  Every cell therefore equals the textbook
//     affine-gap recurrence, in both kernels, which `sweep_columns` computes directly.
//  2. Ends: the alignment ends at the first column (in sweep order) where the running maximum reaches
//     its final value, and at the smallest read position of that column holding it (ssw.c:295-330);
//     the begin comes from the same sweep over the reversed prefixes, stopped at the first column that
//     reaches the score (ssw.c:327, 846-858).  (The 8-bit kernel is used unless its score + bias would
//     reach 255, ssw.c:305; by detail 1 the choice does not change a result and is only reported.)
//  3. The cigar comes from a banded global-ish pass between the two ends whose band doubles until the
//     score is reproduced (ssw.c:584-786); its band-relative buffers are reused between rows, cells on
//     the band edge read zeros (or, when the reference is shorter than the band, a cleared neighbour),
//     ties prefer the diagonal, then the longer-standing gap.  `BandedTrace` keeps the same buffer
//     indexing so these edge effects are identical.
//
// Every function cites the reference lines whose behaviour it reproduces; no reference code is included.
#include "../../include/helen_stitch.h"

#include <algorithm>
#if defined(__SSE2__)
#include <immintrin.h>
#endif
#include <array>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

int32_t fail(int32_t code, const std::string& what) {
    g_error = what;
    return code;
}

// ssw_cpp.cpp:9-18: A/a/U/u -> 0, C/c -> 1, G/g -> 2, T/t -> 3, everything else -> 4
inline int8_t base_code(char c) {
    switch (c) {
        case 'A': case 'a': case 'U': case 'u': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

std::vector<int8_t> encode(const char* s, int64_t n) {
    std::vector<int8_t> out(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) out[i] = base_code(s[i]);
    return out;
}

struct Scoring {
    int match, mismatch, gap_open, gap_extend;
    // ssw_cpp.cpp:20-40: only equal A/C/G/T score `match`; code 4 never matches, not even itself
    int pair(int8_t a, int8_t b) const { return (a < 4 && a == b) ? match : -mismatch; }
    int bias() const { return mismatch > 0 ? mismatch : 0; }          // ssw.c:772-776
};

// ---------------------------------------------------------------------------------------------------------
// Details 1 + 2: one sweep over the columns of the score matrix.

struct SweepEnd {
    int best = 0;        // highest cell value met
    int ref = 0;         // column where `best` was first reached
    int read = 0;        // smallest read position of that column holding `best`
};

// Cell values H(i, p) of the affine-gap local alignment, one reference column at a time:
//   E(i, p) = max(E(i-1, p) - ge, H(i-1, p) - go)          gap that consumes reference bases
//   F(i, p) = max(F(i, p-1) - ge, H(i, p-1) - go)          gap that consumes read bases
//   H(i, p) = max(0, H(i-1, p-1) + pair, E(i, p), F(i, p))
// With go > ge a cell that owes its value to F never opens a better F below it (H - go = F - go < F - ge), so
// F(i, p+1) = max(F(i, p) - ge, T(i, p) - go) with T = max(0, diagonal, E): the part of a column that does not
// depend on F is computed first (independent per row, the compiler vectorises it), and only a two-operation
// chain runs down the column.  Cell type T16 = int16_t while scores fit, int32_t otherwise.
template <typename Cell>
SweepEnd sweep_columns(const int8_t* ref, int ref_len, bool backwards, const int8_t* read, int read_len,
                       const Scoring& sc, int stop_at) {
    const int rows = read_len;
    std::array<std::vector<Cell>, 5> profile;
    for (int c = 0; c < 5; ++c) {
        profile[c].resize(rows);
        for (int p = 0; p < rows; ++p) profile[c][p] = static_cast<Cell>(sc.pair(static_cast<int8_t>(c), read[p]));
    }
    std::vector<Cell> h_store(rows + 1, 0), e_store(rows, static_cast<Cell>(-sc.gap_open)), t_store(rows);
    const Cell go = static_cast<Cell>(sc.gap_open), ge = static_cast<Cell>(sc.gap_extend);
    SweepEnd out;
    for (int n = 0; n < ref_len; ++n) {
        const int i = backwards ? ref_len - 1 - n : n;
        const Cell* __restrict prof = profile[ref[i]].data();
        Cell* __restrict h = h_store.data();          // h[p + 1] = H(previous column, p); h[0] = 0
        Cell* __restrict e = e_store.data();
        Cell* __restrict t = t_store.data();
        for (int p = 0; p < rows; ++p) {
            const Cell ext = static_cast<Cell>(e[p] - ge), opened = static_cast<Cell>(h[p + 1] - go);
            const Cell e_now = ext > opened ? ext : opened;
            const Cell diag = static_cast<Cell>(h[p] + prof[p]);
            Cell best = diag > e_now ? diag : e_now;
            e[p] = e_now;
            t[p] = best > 0 ? best : Cell(0);
        }
        int f = -sc.gap_open, column_best = 0;
        for (int p = 0; p < rows; ++p) {
            const int tp = t[p];
            const int hv = tp > f ? tp : f;
            h[p + 1] = static_cast<Cell>(hv);
            const int f_ext = f - ge, f_open = tp - go;
            f = f_ext > f_open ? f_ext : f_open;
            column_best = hv > column_best ? hv : column_best;
        }
        if (column_best > out.best) {
            out.best = column_best;
            out.ref = i;
            for (int p = 0; p < rows; ++p)
                if (h[p + 1] == column_best) {
                    out.read = p;
                    break;
                }
        }
        if (column_best == stop_at) break;
    }
    return out;
}

#if defined(__SSE2__)
// The same column sweep with 8 read positions per instruction (SSE2 is part of every x86-64).  The F chain becomes
// a running maximum: F(i, p) = max_{k < p} (T(i, k) + k ge) - go - (p - 1) ge, an exclusive prefix maximum of
// G(k) = T(k) + k ge >= 0, taken inside a vector with three shift-and-max steps and carried between vectors.
SweepEnd sweep_columns_sse2(const int8_t* ref, int ref_len, bool backwards, const int8_t* read, int read_len,
                            const Scoring& sc, int stop_at) {
    const int vectors = (read_len + 7) / 8, rows = vectors * 8;
    std::array<std::vector<int16_t>, 5> profile;
    for (int c = 0; c < 5; ++c) {
        profile[c].assign(rows, 0);
        for (int p = 0; p < read_len; ++p) profile[c][p] = static_cast<int16_t>(sc.pair(static_cast<int8_t>(c), read[p]));
    }
    std::vector<int16_t> h_a(rows + 8, 0), h_b(rows + 8, 0), e_store(rows, 0);   // h[p + 1] = H(column, p); h[0] = 0
    alignas(16) int16_t lanes[8], tail[8];
    for (int k = 0; k < 8; ++k) {
        lanes[k] = static_cast<int16_t>(k * sc.gap_extend);
        tail[k] = (rows - 8 + k) < read_len ? int16_t(-1) : int16_t(0);           // clears the padding rows of the last vector
    }
    const __m128i go = _mm_set1_epi16(static_cast<int16_t>(sc.gap_open)), ge = _mm_set1_epi16(static_cast<int16_t>(sc.gap_extend));
    const __m128i step = _mm_set1_epi16(static_cast<int16_t>(8 * sc.gap_extend));
    const __m128i lane_ge = _mm_load_si128(reinterpret_cast<const __m128i*>(lanes));
    const __m128i tail_mask = _mm_load_si128(reinterpret_cast<const __m128i*>(tail));
    const __m128i zero = _mm_setzero_si128();
    int16_t* prev = h_a.data();
    int16_t* cur = h_b.data();
    int16_t* e = e_store.data();
    SweepEnd out;
    for (int n = 0; n < ref_len; ++n) {
        const int i = backwards ? ref_len - 1 - n : n;
        const int16_t* prof = profile[ref[i]].data();
        __m128i k_ge = lane_ge;                                                  // k * ge for the rows of this vector
        __m128i offset = _mm_add_epi16(lane_ge, _mm_sub_epi16(go, ge));          // go + (p - 1) ge
        __m128i carry = zero, column = zero;
        for (int v = 0; v < vectors; ++v) {
            const int p = v * 8;
            const __m128i h_left = _mm_loadu_si128(reinterpret_cast<const __m128i*>(prev + p + 1));
            const __m128i h_diag = _mm_loadu_si128(reinterpret_cast<const __m128i*>(prev + p));
            __m128i e_now = _mm_loadu_si128(reinterpret_cast<const __m128i*>(e + p));
            e_now = _mm_max_epi16(_mm_subs_epi16(e_now, ge), _mm_subs_epi16(h_left, go));
            _mm_storeu_si128(reinterpret_cast<__m128i*>(e + p), e_now);
            __m128i t = _mm_adds_epi16(h_diag, _mm_loadu_si128(reinterpret_cast<const __m128i*>(prof + p)));
            t = _mm_max_epi16(_mm_max_epi16(t, e_now), zero);
            __m128i g = _mm_adds_epi16(t, k_ge);
            g = _mm_max_epi16(g, _mm_slli_si128(g, 2));
            g = _mm_max_epi16(g, _mm_slli_si128(g, 4));
            g = _mm_max_epi16(g, _mm_slli_si128(g, 8));                           // inclusive prefix maximum inside the vector
            const __m128i before = _mm_max_epi16(_mm_slli_si128(g, 2), carry);    // exclusive, with the rows above
            const __m128i f = _mm_subs_epi16(before, offset);
            __m128i h = _mm_max_epi16(t, f);
            if (v == vectors - 1) h = _mm_and_si128(h, tail_mask);
            _mm_storeu_si128(reinterpret_cast<__m128i*>(cur + p + 1), h);
            column = _mm_max_epi16(column, h);
            const __m128i all = _mm_max_epi16(g, carry);
            carry = _mm_shuffle_epi32(_mm_shufflehi_epi16(all, 0xFF), 0xFF);      // last lane to every lane
            k_ge = _mm_add_epi16(k_ge, step);
            offset = _mm_add_epi16(offset, step);
        }
        column = _mm_max_epi16(column, _mm_srli_si128(column, 8));
        column = _mm_max_epi16(column, _mm_srli_si128(column, 4));
        column = _mm_max_epi16(column, _mm_srli_si128(column, 2));
        const int column_best = static_cast<int16_t>(_mm_extract_epi16(column, 0));
        if (column_best > out.best) {
            out.best = column_best;
            out.ref = i;
            const __m128i wanted = _mm_set1_epi16(static_cast<int16_t>(column_best));
            for (int v = 0; v < vectors; ++v) {
                const int hit = _mm_movemask_epi8(_mm_cmpeq_epi16(_mm_loadu_si128(reinterpret_cast<const __m128i*>(cur + v * 8 + 1)), wanted));
                if (hit) {
                    out.read = v * 8 + __builtin_ctz(static_cast<unsigned>(hit)) / 2;
                    break;
                }
            }
        }
        if (column_best == stop_at) break;
        std::swap(prev, cur);
    }
    return out;
}
#endif

#if defined(__SSE2__) && defined(__GNUC__)
#define HS_HAVE_AVX2_SWEEP 1
// 16 read positions per instruction where the CPU has AVX2 (checked once at run time; the library itself is built
// for baseline x86-64 because it is compiled on one machine and shipped to another).  Same arithmetic as above; the
// byte shifts of AVX2 work inside each 128-bit half, so the prefix maximum is finished by folding the last element
// of the low half into the high half.
#define HS_LAST_OF_EACH_HALF(v) _mm256_shuffle_epi32(_mm256_shufflehi_epi16((v), 0xFF), 0xFF)
__attribute__((target("avx2")))
SweepEnd sweep_columns_avx2(const int8_t* ref, int ref_len, bool backwards, const int8_t* read, int read_len,
                            const Scoring& sc, int stop_at) {
    const int vectors = (read_len + 15) / 16, rows = vectors * 16;
    std::array<std::vector<int16_t>, 5> profile;
    for (int c = 0; c < 5; ++c) {
        profile[c].assign(rows, 0);
        for (int p = 0; p < read_len; ++p) profile[c][p] = static_cast<int16_t>(sc.pair(static_cast<int8_t>(c), read[p]));
    }
    std::vector<int16_t> h_a(rows + 16, 0), h_b(rows + 16, 0), e_store(rows, 0);
    alignas(32) int16_t lanes[16], tail[16];
    for (int k = 0; k < 16; ++k) {
        lanes[k] = static_cast<int16_t>(k * sc.gap_extend);
        tail[k] = (rows - 16 + k) < read_len ? int16_t(-1) : int16_t(0);
    }
    const __m256i go = _mm256_set1_epi16(static_cast<int16_t>(sc.gap_open)), ge = _mm256_set1_epi16(static_cast<int16_t>(sc.gap_extend));
    const __m256i step = _mm256_set1_epi16(static_cast<int16_t>(16 * sc.gap_extend));
    const __m256i lane_ge = _mm256_load_si256(reinterpret_cast<const __m256i*>(lanes));
    const __m256i tail_mask = _mm256_load_si256(reinterpret_cast<const __m256i*>(tail));
    const __m256i zero = _mm256_setzero_si256();
    int16_t* prev = h_a.data();
    int16_t* cur = h_b.data();
    int16_t* e = e_store.data();
    SweepEnd out;
    for (int n = 0; n < ref_len; ++n) {
        const int i = backwards ? ref_len - 1 - n : n;
        const int16_t* prof = profile[ref[i]].data();
        __m256i k_ge = lane_ge;
        __m256i offset = _mm256_add_epi16(lane_ge, _mm256_sub_epi16(go, ge));
        __m256i carry = zero, column = zero;
        for (int v = 0; v < vectors; ++v) {
            const int p = v * 16;
            const __m256i h_left = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(prev + p + 1));
            const __m256i h_diag = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(prev + p));
            __m256i e_now = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(e + p));
            e_now = _mm256_max_epi16(_mm256_subs_epi16(e_now, ge), _mm256_subs_epi16(h_left, go));
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(e + p), e_now);
            __m256i t = _mm256_adds_epi16(h_diag, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(prof + p)));
            t = _mm256_max_epi16(_mm256_max_epi16(t, e_now), zero);
            __m256i g = _mm256_adds_epi16(t, k_ge);
            g = _mm256_max_epi16(g, _mm256_slli_si256(g, 2));
            g = _mm256_max_epi16(g, _mm256_slli_si256(g, 4));
            g = _mm256_max_epi16(g, _mm256_slli_si256(g, 8));                     // prefix maximum inside each half
            const __m256i ends = HS_LAST_OF_EACH_HALF(g);
            g = _mm256_max_epi16(g, _mm256_permute2x128_si256(ends, ends, 0x08));  // low half's last element into the high half
            const __m256i low_up = _mm256_permute2x128_si256(g, g, 0x08);          // [0 | low half]
            const __m256i before = _mm256_max_epi16(_mm256_alignr_epi8(g, low_up, 14), carry);   // one row up, with the rows above
            const __m256i f = _mm256_subs_epi16(before, offset);
            __m256i h = _mm256_max_epi16(t, f);
            if (v == vectors - 1) h = _mm256_and_si256(h, tail_mask);
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(cur + p + 1), h);
            column = _mm256_max_epi16(column, h);
            const __m256i all = HS_LAST_OF_EACH_HALF(_mm256_max_epi16(g, carry));
            carry = _mm256_permute2x128_si256(all, all, 0x11);                     // last row to every lane
            k_ge = _mm256_add_epi16(k_ge, step);
            offset = _mm256_add_epi16(offset, step);
        }
        __m128i m = _mm_max_epi16(_mm256_castsi256_si128(column), _mm256_extracti128_si256(column, 1));
        m = _mm_max_epi16(m, _mm_srli_si128(m, 8));
        m = _mm_max_epi16(m, _mm_srli_si128(m, 4));
        m = _mm_max_epi16(m, _mm_srli_si128(m, 2));
        const int column_best = static_cast<int16_t>(_mm_extract_epi16(m, 0));
        if (column_best > out.best) {
            out.best = column_best;
            out.ref = i;
            const __m256i wanted = _mm256_set1_epi16(static_cast<int16_t>(column_best));
            for (int v = 0; v < vectors; ++v) {
                const unsigned hit = static_cast<unsigned>(_mm256_movemask_epi8(
                    _mm256_cmpeq_epi16(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(cur + v * 16 + 1)), wanted)));
                if (hit) {
                    out.read = v * 16 + __builtin_ctz(hit) / 2;
                    break;
                }
            }
        }
        if (column_best == stop_at) break;
        std::swap(prev, cur);
    }
    return out;
}

bool cpu_has_avx2() {
    static const bool yes = __builtin_cpu_supports("avx2") && !std::getenv("HS_NO_AVX2");
    return yes;
}
#endif

SweepEnd sweep(const int8_t* ref, int ref_len, bool backwards, const int8_t* read, int read_len, const Scoring& sc, int stop_at) {
    const int64_t bound = static_cast<int64_t>(std::min(ref_len, read_len)) * sc.match + sc.match + sc.gap_open + sc.mismatch
                          + static_cast<int64_t>(read_len + 8) * sc.gap_extend;
#if defined(HS_HAVE_AVX2_SWEEP)
    if (bound < 32000 && read_len > 24 && cpu_has_avx2()) return sweep_columns_avx2(ref, ref_len, backwards, read, read_len, sc, stop_at);
#endif
#if defined(__SSE2__)
    if (bound < 32000) return sweep_columns_sse2(ref, ref_len, backwards, read, read_len, sc, stop_at);
#else
    if (bound < 32000) return sweep_columns<int16_t>(ref, ref_len, backwards, read, read_len, sc, stop_at);
#endif
    return sweep_columns<int32_t>(ref, ref_len, backwards, read, read_len, sc, stop_at);
}

// ---------------------------------------------------------------------------------------------------------
// Detail 3: banded pass + traceback (behaviour of ssw.c:584-786).

struct CigarOp {
    int length;
    char op;
};

class BandedTrace {
public:
    BandedTrace(const int8_t* ref, int ref_len, const int8_t* read, int read_len, const Scoring& sc)
        : ref_(ref), read_(read), ref_len_(ref_len), read_len_(read_len), sc_(sc) {}

    // false: the traceback met a cell that was never filled
    bool run(int score, int band, std::vector<CigarOp>& cigar) {
        int best = 0;
        do {
            band_ = band;
            fill(best);
            band *= 2;
        } while (best < score);
        return trace(cigar);
    }

private:
    const int8_t* ref_;
    const int8_t* read_;
    int ref_len_, read_len_;
    const Scoring& sc_;
    int band_ = 0;
    // row buffers addressed by band slot; their contents survive from row to row and from one band width to
    // the next exactly as the reference's reallocated arrays do
    std::vector<int> h_above_, e_above_, h_row_;
    std::vector<int8_t> moves_;      // per cell: [how E was reached, how F was reached, how H was reached]

    int first_col(int i) const { return i > band_ ? i - band_ : 0; }          // leftmost column stored for row i
    int slot(int i, int j) const { return j - first_col(i) + 1; }
    int64_t cell(int i, int j, int which) const { return static_cast<int64_t>(j - first_col(i)) * 3 + which; }
    int64_t row_stride() const { return static_cast<int64_t>(band_ * 2 + 1) * 3; }

    void fill(int& best) {
        const int width = band_ * 2 + 3;
        const size_t slots = static_cast<size_t>(width) + 1;
        if (h_above_.size() < slots) {
            h_above_.resize(slots, 0);
            e_above_.resize(slots, 0);
            h_row_.resize(slots, 0);
        }
        const size_t cells = static_cast<size_t>(row_stride()) * read_len_ + 3;
        if (moves_.size() < cells) moves_.resize(cells, 0);
        const int go = sc_.gap_open, ge = sc_.gap_extend;

        for (int k = 1; k < width - 1; ++k) h_above_[k] = 0;
        for (int i = 0; i < read_len_; ++i) {
            const int beg = std::max(0, i - band_);
            const int end = std::min(ref_len_ - 1, i + band_);
            const int edge = std::min(end + 1, width - 1);
            int f = 0, last_slot = 0;
            h_above_[0] = e_above_[0] = h_above_[edge] = e_above_[edge] = h_row_[0] = 0;
            int8_t* moves = moves_.data() + row_stride() * i;

            for (int j = beg; j <= end; ++j) {
                const int u = slot(i, j);
                const int up = slot(i - 1, j), diag_slot = slot(i - 1, j - 1);
                // E: a gap in the reference (read base i against nothing)
                const int e_open = (i == 0 ? 0 : h_above_[up]) - go;
                const int e_ext = (i == 0 ? 0 : e_above_[up]) - ge;
                const int e = std::max(e_open, e_ext);
                e_above_[u] = e;
                const int8_t e_move = e_open > e_ext ? 3 : 2;
                // F: a gap in the read
                const int f_open = h_row_[u - 1] - go;
                const int f_ext = f - ge;
                f = std::max(f_open, f_ext);
                const int8_t f_move = f_open > f_ext ? 5 : 4;

                const int e1 = std::max(e, 0), f1 = std::max(f, 0);
                const int gap = std::max(e1, f1);
                const int diag = h_above_[diag_slot] + sc_.pair(ref_[j], read_[i]);
                const int h = std::max(gap, diag);
                h_row_[u] = h;
                best = std::max(best, h);

                moves[cell(i, j, 0)] = e_move;
                moves[cell(i, j, 1)] = f_move;
                moves[cell(i, j, 2)] = gap <= diag ? int8_t(1) : (e1 > f1 ? e_move : f_move);
                last_slot = u;
            }
            for (int k = 1; k <= last_slot; ++k) h_above_[k] = h_row_[k];
        }
    }

    bool trace(std::vector<CigarOp>& cigar) const {
        // walk from the last cell to row 0 (ssw.c:677-731); the run lengths are collected back to front
        std::vector<CigarOp> reversed;
        int i = read_len_ - 1, j = ref_len_ - 1;
        int run = 0, state = 2;
        char op = 'M', run_op = 'M';
        int64_t row = row_stride() * (read_len_ - 1);
        while (i > 0) {
            const int64_t at = row + cell(i, j, state);
            if (at < 0 || at >= static_cast<int64_t>(moves_.size())) return false;
            switch (moves_[at]) {
                case 1: --i; --j; state = 2; row -= row_stride(); op = 'M'; break;
                case 2: --i; state = 0; row -= row_stride(); op = 'I'; break;
                case 3: --i; state = 2; row -= row_stride(); op = 'I'; break;
                case 4: --j; state = 1; op = 'D'; break;
                case 5: --j; state = 2; op = 'D'; break;
                default: return false;
            }
            if (op == run_op) {
                ++run;
            } else {
                reversed.push_back({run, run_op});
                run_op = op;
                run = 1;
            }
        }
        if (op == 'M') {
            reversed.push_back({run + 1, 'M'});          // row 0 is always an aligned pair
        } else {
            reversed.push_back({run, op});
            reversed.push_back({1, 'M'});
        }
        cigar.assign(reversed.rbegin(), reversed.rend());
        return true;
    }
};

// ssw_cpp.cpp:43-79 + :104-187: soft clips around the path, M runs split into '=' and 'X'
std::string describe(const std::vector<CigarOp>& path, const int8_t* ref, const int8_t* query, int query_len,
                     const hs_alignment& al, int& mismatches) {
    std::string out;
    auto emit = [&out](int length, char op) { out += std::to_string(length); out += op; };
    if (path.empty()) return out;
    if (al.query_begin > 0) emit(al.query_begin, 'S');
    const int8_t* r = ref + al.ref_begin;
    const int8_t* q = query + al.query_begin;
    int same = 0, diff = 0;
    auto flush = [&]() {
        if (same) emit(same, '=');
        else if (diff) emit(diff, 'X');
        same = diff = 0;
    };
    mismatches = 0;
    for (const CigarOp& c : path) {
        if (c.op == 'M') {
            for (int k = 0; k < c.length; ++k, ++r, ++q) {
                if (*r != *q) {
                    ++mismatches;
                    if (same) { emit(same, '='); same = 0; }
                    ++diff;
                } else {
                    if (diff) { emit(diff, 'X'); diff = 0; }
                    ++same;
                }
            }
        } else if (c.op == 'I') {
            q += c.length;
            mismatches += c.length;
            flush();
            emit(c.length, 'I');
        } else if (c.op == 'D') {
            r += c.length;
            mismatches += c.length;
            flush();
            emit(c.length, 'D');
        }
    }
    flush();
    const int tail = query_len - al.query_end - 1;
    if (tail > 0) emit(tail, 'S');
    return out;
}

// ssw.c:801-887 (ssw_align with flag 0x0f, filters 0 / 32767, maskLen 0) + ssw_cpp.cpp:320-352
int32_t local_align(const std::vector<int8_t>& ref, const std::vector<int8_t>& query, const Scoring& sc,
                    hs_alignment& al, std::string& cigar) {
    al = hs_alignment{};
    cigar.clear();
    // the reference's own fix-up loops are only exact when opening a gap costs more than extending one (with
    // equal costs its traceback can run off the band and crash); identical results are established for that case only
    if (sc.match <= 0 || sc.mismatch < 0 || sc.gap_extend < 0 || sc.gap_open <= sc.gap_extend || sc.match + sc.bias() >= 128)
        return fail(HS_E_ARGUMENT, "scoring must have match > 0, mismatch >= 0, gap_open > gap_extend >= 0");
    const int ref_len = static_cast<int>(ref.size()), query_len = static_cast<int>(query.size());
    if (ref_len == 0 || query_len == 0) return 0;          // Align_cpp returns false, the Alignment stays zero

    const SweepEnd fwd = sweep(ref.data(), ref_len, false, query.data(), query_len, sc, -1);
    if (fwd.best > 32767) return fail(HS_E_RANGE, "alignment score exceeds the 16-bit kernel of the reference");
    al.kernel = fwd.best >= 255 - sc.bias() ? 16 : 8;
    al.score = fwd.best;
    if (fwd.best == 0) return 0;
    al.ref_end = fwd.ref;
    al.query_end = fwd.read;

    // the begin: same sweep over the reversed prefixes, stopped at the first column that reaches the score
    std::vector<int8_t> reversed(query.begin(), query.begin() + fwd.read + 1);
    std::reverse(reversed.begin(), reversed.end());
    const SweepEnd back = sweep(ref.data(), fwd.ref + 1, true, reversed.data(), fwd.read + 1, sc, fwd.best);
    al.ref_begin = back.ref;
    al.query_begin = fwd.read - back.read;
    if (back.best != fwd.best || al.ref_begin > al.ref_end || al.query_begin < 0)
        return fail(HS_E_TRACE, "reverse sweep did not locate the alignment begin");

    const int sub_ref = al.ref_end - al.ref_begin + 1, sub_query = al.query_end - al.query_begin + 1;
    const int band = std::abs(sub_ref - sub_query) + 1;
    std::vector<CigarOp> path;
    BandedTrace tracer(ref.data() + al.ref_begin, sub_ref, query.data() + al.query_begin, sub_query, sc);
    if (!tracer.run(fwd.best, band, path)) return fail(HS_E_TRACE, "banded traceback left the filled band");
    int mismatches = 0;
    cigar = describe(path, ref.data(), query.data(), query_len, al, mismatches);
    al.mismatches = mismatches;
    al.cigar_len = static_cast<int32_t>(cigar.size());
    return 0;
}

// Stitch.py:34-94
int32_t anchor_from_cigar(const char* cigar, int32_t ref_begin, int32_t min_run, int32_t& ref_pos, int32_t& query_pos) {
    std::vector<CigarOp> runs;
    for (const char* p = cigar; *p;) {
        if (*p < '0' || *p > '9') return fail(HS_E_CIGAR, std::string("malformed cigar: ") + cigar);
        int64_t n = 0;
        while (*p >= '0' && *p <= '9') n = n * 10 + (*p++ - '0');
        if (!*p) return fail(HS_E_CIGAR, std::string("malformed cigar: ") + cigar);
        char op = *p++;
        if (op == '=' || op == 'X') op = 'M';
        if (!runs.empty() && runs.back().op == op) runs.back().length += static_cast<int>(n);
        else runs.push_back({static_cast<int>(n), op});
    }
    int ref_index = ref_begin, read_index = 0;
    for (const CigarOp& r : runs) {
        if (r.op == 'M' && r.length >= min_run) {
            ref_pos = ref_index;
            query_pos = read_index;
            return 0;
        }
        switch (r.op) {
            case 'S': case 'I': read_index += r.length; break;
            case 'D': ref_index += r.length; break;
            case 'M': ref_index += r.length; read_index += r.length; break;
            default: return fail(HS_E_CIGAR, std::string("invalid cigar operation encountered while stitching: ") + r.op);
        }
    }
    ref_pos = query_pos = -1;
    return 0;
}

struct Piece {
    int64_t start, end;
    std::string sequence;
};

}  // namespace

struct hs_stitcher {
    Scoring scoring;
    int32_t overlap_threshold;
    double base_error_rate;
    std::vector<Piece> pieces;
    std::string running;
};

extern "C" {

int32_t hs_abi_version(void) { return HS_ABI_VERSION; }

const char* hs_last_error(void) { return g_error.c_str(); }

int32_t hs_ssw_align(const char* ref, int32_t ref_len, const char* query, int32_t query_len,
                     const hs_scoring* scoring, hs_alignment* out, char* cigar, int32_t cigar_cap) {
    if (!scoring || !out || !cigar || cigar_cap < 1 || ref_len < 0 || query_len < 0 || (!ref && ref_len) || (!query && query_len))
        return fail(HS_E_ARGUMENT, "hs_ssw_align: bad argument");
    const Scoring sc{scoring->match, scoring->mismatch, scoring->gap_open, scoring->gap_extend};
    std::string text;
    const int32_t rc = local_align(encode(ref, ref_len), encode(query, query_len), sc, *out, text);
    if (rc) return rc;
    if (static_cast<int64_t>(text.size()) + 1 > cigar_cap) return fail(HS_E_CAPACITY, "hs_ssw_align: cigar buffer too small");
    std::memcpy(cigar, text.c_str(), text.size() + 1);
    return 0;
}

int32_t hs_anchor_from_cigar(const char* cigar, int32_t ref_begin, int32_t min_run, int32_t* ref_pos, int32_t* query_pos) {
    if (!cigar || !ref_pos || !query_pos) return fail(HS_E_ARGUMENT, "hs_anchor_from_cigar: bad argument");
    return anchor_from_cigar(cigar, ref_begin, min_run, *ref_pos, *query_pos);
}

int64_t hs_decode_region(const int64_t* positions, const uint8_t* bases, const uint8_t* rles, int64_t n,
                         char* out, int64_t out_cap) {
    if (n < 0 || !out || out_cap < 1 || (n && (!positions || !bases || !rles)))
        return fail(HS_E_ARGUMENT, "hs_decode_region: bad argument");
    // Stitch.py:227-238: first prediction of a (position, index, split) key wins; keys are then sorted
    std::vector<int64_t> order;
    order.reserve(static_cast<size_t>(n));
    for (int64_t k = 0; k < n; ++k)
        if (positions[3 * k] >= 0 && positions[3 * k + 1] >= 0) order.push_back(k);
    auto key_less = [positions](int64_t a, int64_t b) {
        return std::lexicographical_compare(positions + 3 * a, positions + 3 * a + 3, positions + 3 * b, positions + 3 * b + 3);
    };
    std::stable_sort(order.begin(), order.end(), key_less);
    static const char letters[5] = {0, 'A', 'C', 'G', 'T'};          // Options.py:3
    int64_t length = 0;
    for (size_t k = 0; k < order.size(); ++k) {
        if (k && !key_less(order[k - 1], order[k])) continue;         // same key as the row kept before it
        const uint8_t base = bases[order[k]];
        if (base > 4) return fail(HS_E_ARGUMENT, "hs_decode_region: base label outside 0..4");
        if (base == 0) continue;
        const int64_t repeat = rles[order[k]];
        if (length + repeat + 1 > out_cap) return fail(HS_E_CAPACITY, "hs_decode_region: output buffer too small");
        std::memset(out + length, letters[base], static_cast<size_t>(repeat));
        length += repeat;
    }
    out[length] = 0;
    return length;
}

hs_stitcher* hs_stitcher_create(const hs_scoring* scoring, int32_t overlap_threshold, double base_error_rate) {
    if (!scoring) {
        fail(HS_E_ARGUMENT, "hs_stitcher_create: null scoring");
        return nullptr;
    }
    auto* s = new hs_stitcher;
    s->scoring = Scoring{scoring->match, scoring->mismatch, scoring->gap_open, scoring->gap_extend};
    s->overlap_threshold = overlap_threshold;
    s->base_error_rate = base_error_rate;
    return s;
}

void hs_stitcher_destroy(hs_stitcher* s) { delete s; }

int32_t hs_stitcher_add(hs_stitcher* s, int64_t start, int64_t end, const char* sequence, int64_t length) {
    if (!s || length < 0 || (!sequence && length)) return fail(HS_E_ARGUMENT, "hs_stitcher_add: bad argument");
    s->pieces.push_back(Piece{start, end, std::string(sequence ? sequence : "", static_cast<size_t>(length))});
    return 0;
}

// Stitch.py:96-193
int32_t hs_stitcher_run(hs_stitcher* s, int64_t* start, int64_t* end, int64_t* length, int64_t* warnings, int64_t* alignments) {
    if (!s || s->pieces.empty()) return fail(HS_E_ARGUMENT, "hs_stitcher_run: no pieces");
    std::stable_sort(s->pieces.begin(), s->pieces.end(), [](const Piece& a, const Piece& b) {
        return a.start != b.start ? a.start < b.start : a.end < b.end;
    });
    int64_t warned[3] = {0, 0, 0}, aligned = 0;
    std::string& running = s->running;
    running = s->pieces[0].sequence;
    const int64_t running_start = s->pieces[0].start;
    int64_t running_end = s->pieces[0].end;
    const std::string gap(10, 'N');

    for (size_t k = 1; k < s->pieces.size(); ++k) {
        const Piece& piece = s->pieces[k];
        const std::string& incoming = piece.sequence;
        if (piece.start < running_end) {
            int64_t overlap = running_end - piece.start;
            overlap += static_cast<int64_t>(static_cast<double>(overlap) * s->base_error_rate);
            const size_t keep = running.size() > static_cast<size_t>(overlap) ? running.size() - static_cast<size_t>(overlap) : 0;
            const size_t head = std::min(incoming.size(), static_cast<size_t>(overlap));
            // align the head of the incoming piece to the tail of the running sequence
            hs_alignment al;
            std::string cigar;
            ++aligned;
            const int32_t rc = local_align(encode(running.data() + keep, static_cast<int64_t>(running.size() - keep)),
                                           encode(incoming.data(), static_cast<int64_t>(head)), s->scoring, al, cigar);
            if (rc) return rc;
            if (al.score == 0) {
                ++warned[HS_WARN_NO_ALIGNMENT];
                if (head > 10) {                        // only the overlapping head is appended (Stitch.py:144-147)
                    running += gap;
                    running.append(incoming, 0, head);
                    running_end = piece.end;
                }
                continue;
            }
            int32_t ref_pos = -1, query_pos = -1;
            const int32_t arc = anchor_from_cigar(cigar.c_str(), al.ref_begin, s->overlap_threshold, ref_pos, query_pos);
            if (arc) return arc;
            if (ref_pos == -1 || query_pos == -1) {
                ++warned[HS_WARN_NO_ANCHOR];
                if (incoming.size() > 10) {
                    running += gap;
                    running += incoming;
                    running_end = piece.end;
                }
                continue;
            }
            running.resize(keep + static_cast<size_t>(ref_pos));
            if (static_cast<size_t>(query_pos) < incoming.size()) running.append(incoming, static_cast<size_t>(query_pos), std::string::npos);
            running_end = piece.end;
        } else {
            ++warned[HS_WARN_NO_OVERLAP];
            if (incoming.size() > 10) {
                running += gap;
                running += incoming;
                running_end = piece.end;
            }
        }
    }
    if (start) *start = running_start;
    if (end) *end = running_end;
    if (length) *length = static_cast<int64_t>(running.size());
    if (warnings) std::copy(warned, warned + 3, warnings);
    if (alignments) *alignments = aligned;
    return 0;
}

int64_t hs_stitcher_sequence(const hs_stitcher* s, char* out, int64_t out_cap) {
    if (!s || !out) return fail(HS_E_ARGUMENT, "hs_stitcher_sequence: bad argument");
    if (static_cast<int64_t>(s->running.size()) + 1 > out_cap) return fail(HS_E_CAPACITY, "hs_stitcher_sequence: output buffer too small");
    std::memcpy(out, s->running.c_str(), s->running.size() + 1);
    return static_cast<int64_t>(s->running.size());
}

}  // extern "C"
