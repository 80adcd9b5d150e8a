// Tensor engine: the GRU / head contractions on tcgen05 tensor cores (sm_100a) with
// fp32-equivalent results through a 3-term split of fp16 operands:
//     W = (W_hi + W_lo) * 2^-kw ,  x = (x_hi + x_lo) * 2^-10          (hi, lo in fp16)
//     W.x  ~=  (W_hi.x_hi + W_lo.x_hi + W_hi.x_lo) * 2^-(kw+10)       fp32 accumulate in TMEM
// uint8 pileup pixels are exact in fp16, so the encoder input projection needs only 2 terms.
// (tools/precision_probe.py: max |dP| 2.6e-7 vs fp64, same as plain fp32; bf16 3-term: 1.3e-5.)
//
// Data path of one chunk (reference loop body, predict_gpu.py:114-149):
//   ximg --proj(enc)--> gi --rec(enc)--> yimg1 --proj(dec)--> gi --rec(dec)--> yimg2 --heads--> P +=
//
// Every activation tensor that feeds an MMA is kept in global memory as *operand images*: blocks
// of [8 windows x K] fp16 already in the K-major core-matrix layout tcgen05 reads from shared
// memory, so a tile is staged with a handful of 1-D bulk async copies (TMA engine, cp.async.bulk)
// and no thread ever touches it:
//   ximg : [window group][column t][K/8 k-groups][8 windows][8 k]          (pixels, exact fp16)
//   yimg : [window group][direction][hi, lo][column t][16 k-groups x 144 B] (GRU outputs * 2^10); consecutive columns of
//          one (group, direction, part) are contiguous, so a tile of columns is ONE bulk copy per direction and part
// Weights are the stationary operand: for the GRU contractions they live in tensor memory (TMEM)
// as the A operand (M = 128 gate rows); a thread's TMEM lane is "its" gate row / hidden unit.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <type_traits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/helen_b200.h"
#include "fp32_kernels.cuh"
#include "tc_ptx.cuh"
#include "workspace.cuh"

namespace hb {

constexpr float ACT_SCALE = 1024.0f;              // activations in (-1, 1) are scaled by 2^10 before the split
constexpr float ACT_SCALE_INV = 1.0f / 1024.0f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int WG = 8;                             // windows per window group == rows of one core matrix
// h / y images: [8 windows x 128 k] fp16 blocks in the K-major 128-byte-swizzle layout (tc::sw128_offset): dense, the
// tensor core reads full 128-byte rows, and the gate threads' 2-byte stores still spread over all banks (the XOR moves
// the four k-chunks a warp writes to different bank groups).  (The MMA rate is the same as with the no-swizzle core-matrix
// layout, tools/mma_rate.cu; that layout needed its k-group stride padded to 144 B to avoid the store conflicts.)
constexpr int YBLK = 2048;                        // [8 windows x 128 k] fp16 image of one direction: two swizzle atoms
constexpr int YROW = 2 * YBLK;                    // 4608 B: both directions (K = 256) of one (group, t, part)  [sizes only]
constexpr int GI_ROW_BYTES = G * 4;               // 1536 B: gi of one (window, t, direction)
// gi' lives in global memory as a "gi image": [window group][direction x gate r,z,n = 6 blocks][column][128 units][8 windows]
// fp32.  What one projection job produces (one gate block of 8 windows x 8 consecutive columns, 32 KB) is contiguous, and
// a projection thread (one gate row) writes the 8 windows of a column as two 16-byte stores; what one recurrence step
// needs (the three gate blocks of a window group's column and direction) is three 4 KB bulk copies, and a gate thread
// (one hidden unit, 2 / 4 / 8 windows) reads its values with one or two vector loads.  The two 16-byte halves of a
// unit's 8 windows are swapped for units with bit 2 set (gi_window_index), which makes those 16-byte shared-memory reads
// conflict-free (lane stride 32 bytes).
// (The first layouts were window-major: [..][8 windows][128 units].  A step's 12 KB contiguous made every projection job
// eight separate bulk stores; one block per gate made it one 32 KB bulk store, but bulk stores out of shared memory run
// at ~20 B per cycle next to the MMAs' operand reads and the store warp paced the projection role either way.)
constexpr int GI_BLK_FLOATS = WG * H;             // 1024: one (group, gate block, column)
constexpr int GI_BLK_BYTES = GI_BLK_FLOATS * 4;   // 4096
constexpr int GI_GRP_BYTES = 3 * GI_BLK_BYTES;    // 12288 B: the three gate blocks of one (group, column, direction) in a shared-memory stage
__host__ __device__ constexpr int64_t gi_block(int64_t wg, int64_t cols, int64_t t, int blk) { return ((wg * 6 + blk) * cols + t) * GI_BLK_FLOATS; }
// float index of (unit u, window w of the group) inside a block
__host__ __device__ constexpr int gi_window_index(int u, int w) { return u * WG + ((((w >> 2) ^ (u >> 2)) & 1) << 2) + (w & 3); }
// The shared-memory loads of a stage must have RETURNED before the stage is handed back to the bulk-copy engine: an
// mbarrier arrive does not wait for the warp's outstanding loads (with 8 windows per thread - six 16-byte loads in flight -
// the refill was seen to overtake the last ones once in ~50 launches).  Naming the registers here makes the arrive depend on them.
template <int NW>
__device__ __forceinline__ void gi_loads_done(const float* r, const float* z, const float* n) {
#pragma unroll
    for (int i = 0; i < NW; ++i) asm volatile("" ::"f"(r[i]), "f"(z[i]), "f"(n[i]) : "memory");
}
// a gate thread's NW consecutive windows (starting at w0, a multiple of NW) of unit u: one or two vector loads
template <int NW>
__device__ __forceinline__ void gi_load(const float* blk, int u, int w0, float* out) {
    static_assert(NW == 2 || NW == 4 || NW == 8, "windows per gate thread");

    if constexpr (NW == 2) {
        const float2 v = *reinterpret_cast<const float2*>(blk + gi_window_index(u, w0));
        out[0] = v.x; out[1] = v.y;
    } else if constexpr (NW == 4) {
        const float4 v = *reinterpret_cast<const float4*>(blk + gi_window_index(u, w0));
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else {
        const float4 lo = *reinterpret_cast<const float4*>(blk + gi_window_index(u, 0));
        const float4 hi = *reinterpret_cast<const float4*>(blk + gi_window_index(u, 4));
        out[0] = lo.x; out[1] = lo.y; out[2] = lo.z; out[3] = lo.w; out[4] = hi.x; out[5] = hi.y; out[6] = hi.z; out[7] = hi.w;
    }
}

// -DHB_TIMELINE: the chain from the encoder's last column to the decoder's first step of window group 0 in chunk 2
// (slots 7400..: encoder fwd / rev final publication, scheduler pick, loader issue, MMA commit, counter raised, decoder
// loader saw the counter, first gi row in shared memory)
#ifdef HB_TIMELINE
#define HB_CHAIN(dbgp, cond, k) do { if ((dbgp) != nullptr && (cond)) (dbgp)[7400 + (k)] = (long long)globaltimer_ns(); } while (0)
#else
#define HB_CHAIN(dbgp, cond, k) do { } while (0)
#endif
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Weight images in global memory are stored in the order the TMEM upload reads them: one warp instruction (32 lanes x
// 16 bytes, lane = TMEM lane = matrix row within a 32-row quarter) fetches 512 contiguous bytes.  (Row-major rows made
// every such instruction touch 32 different lines and the upload - once per launch, or per phase of the chunk-loop
// kernel - took ~10 us.)
//   W_ih block image: [gate block 6][hi, lo][quarter 4][16-word unit][v 4][lane 32][4 words]
//   W_hh image      : [dir 2][hi, lo][gate block 3][quarter 4][half 2][v 8][lane 32][4 words]
__host__ __device__ constexpr size_t wih_word_index(int blk, int term, int row, int c, int kwords) {
    return ((((((size_t)blk * 2 + term) * 4 + (row >> 5)) * (kwords >> 4) + (c >> 4)) * 4 + ((c >> 2) & 3)) * 32 + (row & 31)) * 4 + (c & 3);
}
__host__ __device__ constexpr size_t whh_word_index(int dir, int term, int gb, int row, int c) {
    return (((((((size_t)dir * 2 + term) * 3 + gb) * 4 + (row >> 5)) * 2 + (c >> 5)) * 8 + ((c >> 2) & 7)) * 32 + (row & 31)) * 4 + (c & 3);
}

__host__ __device__ constexpr int64_t yimg_block(int64_t wg, int dir, int part, int t, int W) { return (((wg * 2 + dir) * 2 + part) * W + t) * (int64_t)YBLK; }

// ---------------------------------------------------------------------------------------------
// uint8 pileup -> fp16 operand image (once per batch; predict_gpu.py:97 does this cast on the host)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pileup_to_operand_image_kernel(const uint8_t* __restrict__ images, int64_t B, int T, int F, int Kp, __half* __restrict__ ximg, int64_t n_wg)
{
    const int kg = Kp >> 3;
    const int64_t total = n_wg * T * kg * WG;               // one 16-byte chunk (8 k of one window) per thread
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(e % WG);
        const int g8 = (int)((e / WG) % kg);
        const int64_t rest = e / (WG * kg);
        const int t = (int)(rest % T);
        const int64_t wg = rest / T;
        const int64_t b = wg * WG + w;
        __align__(16) __half v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = g8 * 8 + i;
            v[i] = __float2half_rn((b < B && k < F) ? (float)images[(b * T + t) * F + k] : 0.f);
        }
        *reinterpret_cast<int4*>(ximg + e * 8) = *reinterpret_cast<const int4*>(v);
    }
}

// ---------------------------------------------------------------------------------------------
// Input projection  gi'[m, 0:768] = scale_row * (A[m, :] . Wcat^T) + bias_row
// grid = (workers, 6 gate blocks).  The CTA's [128 x Kp] weight block (hi, lo) is TMEM-resident;
// a tile is 64 data rows = 8 windows x 8 consecutive columns, staged by bulk copies.
// Warps 0-3 and 7-10: two epilogue groups (TMEM lane quarter = warp % 4), one per accumulator buffer, so that one group's
// gpu-scope fence (~1700 cycles under load) overlaps the other group's tile.  Warp 4: MMA issuer.  Warp 5: tile loader.
// Warp 6 (chunk-loop kernel only): job scheduler.
// The epilogue writes gi' straight from registers: a thread owns one gate row (its TMEM lane) and the 8 windows of a column
// are contiguous in the gi image, so a column is two 16-byte stores per thread and 1 KB contiguous per warp.  (Staging the
// block in shared memory for a bulk store was the first version: one 32 KB bulk store took ~1700 cycles to issue and drain
// - the tile's MMAs, ~1550 cycles, already keep the shared-memory port busy with their operand reads - and the store warp
// paced the whole role.)  In the chunk-loop kernel every epilogue warp makes its stores visible (gpu-scope fence) and
// raises the tile's counter itself: PROJ_FLAGS_PER_TILE = 3 gate blocks x 4 warps per tile.
// The epilogue folds the bias sums and the -log2(e) factors of the gate nonlinearities into gi'
// (see tc_recurrence_kernel), so gi' is NOT the plain pre-activation of the fp32 engine.
// ---------------------------------------------------------------------------------------------
constexpr int PROJ_THREADS = 352;                // warps 0-3 and 7-10 epilogue (two groups), 4 MMA issuer, 5 tile loader, 6 job scheduler
constexpr int PROJ_NT = 64;
constexpr int PROJ_STAGES = 3;                    // input stages (64 KB each for the decoder's K = 256, hi and lo parts)
constexpr int PROJ_TABLE_MAX = 1024;              // chunk-loop kernel: jobs of one worker per chunk (the scheduler keeps the list in shared memory)
constexpr int PROJ_GROUPS_MAX = 80;               // chunk-loop kernel: window groups of a batch (the scheduler caches their encoder counters)
constexpr int PROJ_RING = 16;                     // job ids in flight between scheduler / loader / MMA issuer / epilogue
constexpr int PROJ_SCHED_LEAD = 4;                // jobs the scheduler may decide ahead of the loader
constexpr int PROJ_FLAGS_PER_TILE = 12;           // counter increments that complete a (group, tile, direction): 3 gate blocks x 4 epilogue warps
constexpr int PROJ_FLAG_BATCH = 4;                // jobs of an epilogue group whose counters may wait for one common fence
constexpr int PROJ_W_COL0 = 128;

// Encoder input projection inside the chunk-loop kernel ("pixel jobs"): while the encoder of chunk k runs, the projection
// role is idle until the first decoder jobs become runnable; it uses that time to project the image columns the encoder of
// chunk k+1 adds (J new columns), so only the first chunk's columns are projected before the kernel starts.
constexpr int PX_R = 8;                          // at most ceil(J / 8) + 1 new column tiles per chunk
constexpr int PX_W_COL0 = 384;                   // TMEM columns of the encoder's W_ih block (after the decoder's 256)
struct ProjPixelArgs {
    int last;                      // 1: a chunk's pixel jobs are scheduled after its decoder tiles, 0: before
    const uint8_t* ximg; int64_t wg_stride; int blk_bytes, Kp;   // pixel operand image (no-swizzle, lbo 128)
    const uint32_t* w_tmem; const float* scale_row; const float* bias_row;
    float* gi; int cols, tiles;                  // gi image of all `cols` image columns (`tiles` column tiles)
    int col_step;                                // J
    unsigned long long* flags;                   // [group][tile][direction], PROJ_FLAGS_PER_TILE when complete
    const int* wgs_of_worker;                    // [workers] window groups whose pixel jobs the worker owns
};

struct ProjArgs {
    // operand image addressing (bytes): block of (group wg, K-slice d, part p, column t) at
    // in_base + wg * in_wg_stride + d * in_dir_stride + p * in_part_stride + t * blk_bytes  (columns contiguous)
    const uint8_t* in_base; int64_t in_wg_stride, in_dir_stride, in_part_stride;
    int blk_bytes;                 // one (group, column) block of ONE K-slice
    int n_dirs;                    // K-slices per column: 2 for the GRU output image (forward | reverse units), 1 for pixels
    int lbo;                       // k-group stride of a no-swizzle image (pixels), or 0: blocks are 128-byte-swizzle images (GRU outputs)
    int Kp; int64_t n_wg; int W;
    const uint32_t* w_tmem;        // packed fp16 pairs, see wih_word_index
    const float* scale_row;        // [768]
    const float* bias_row;         // [768]
    float* gi;                     // gi image with W columns (see gi_block)
    // pair mode: the launch uses clusters of 2 CTAs along the gate-block axis; both CTAs of a pair walk the same
    // tiles, each fetches half of a tile and multicasts it to both, halving the L2 traffic of the activations
    int pair;
    // ---- chunk-loop kernel (producer/consumer mode): the role walks n_chunks chunks.  A job is one column tile of one
    // window group (the whole K = 256 contraction), runnable once BOTH encoder directions have published the tile's
    // columns.  (K-half jobs, runnable per direction with the second half added to the first at the destination, kept
    // the role busier during the encoder phase but doubled its output traffic, which is what bounds it.)
    const int* jobs;               // per worker, in the order the encoder makes them runnable (see pack_proj_job)
    const int* job_offsets;        // [workers + 1]
    const unsigned long long* progress; unsigned long long epoch; int rec_n;   // encoder progress counters, [cta][dir]
    int n_chunks;                  // a chunk's columns count from chunk * W in the progress counters
    unsigned long long* tile_flags;   // [group][tile][decoder direction], PROJ_FLAGS_PER_TILE per chunk
    unsigned long long* tile_reads;   // [group][tile]: gate-block CTAs that are done with the encoder output tile (6 per chunk), or nullptr
    long long* dbg;                   // -DHB_TIMELINE: worker 0 records when it finished each chunk
    ProjPixelArgs px;                 // px.ximg == nullptr: no pixel jobs
    int col_tiles;                    // tile mode: project only the first col_tiles column tiles (0 = all)
};

__host__ __device__ constexpr int pack_proj_job(int wg, int tile) { return wg | (tile << 16); }

struct ProjJob { int64_t wg; int t0, valid; bool pixel; };

// jobs of one worker per chunk, and the idx-th of them in table / tile order
__device__ __forceinline__ int proj_count(const ProjArgs& a, int worker, int n_workers) {
    if (a.jobs != nullptr) return __ldg(a.job_offsets + worker + 1) - __ldg(a.job_offsets + worker);
    const int tiles_t = a.col_tiles > 0 ? a.col_tiles : ((a.W + 7) >> 3);
    const int64_t n_tiles = a.n_wg * tiles_t;
    return n_tiles > worker ? (int)((n_tiles - worker + n_workers - 1) / n_workers) : 0;
}
// column tiles of the pixel image projected before chunk k's encoder starts
__device__ __forceinline__ int px_done(const ProjArgs& a, int k) { return min((a.px.col_step * k + a.W + 7) >> 3, a.px.tiles); }
__device__ __forceinline__ ProjJob proj_decode(const ProjArgs& a, int e) {
    ProjJob j;
    j.wg = e & 0xffff; j.t0 = ((e >> 16) & 0xfff) * 8; j.pixel = (e >> 29) & 1;
    j.valid = min(8, (j.pixel ? a.px.cols : a.W) - j.t0);
    return j;
}
__device__ __forceinline__ ProjJob proj_tile_job(const ProjArgs& a, int worker, int n_workers, int idx) {
    const int64_t tile = worker + (int64_t)idx * n_workers;
    ProjJob j;
    j.wg = tile % a.n_wg; j.t0 = (int)(tile / a.n_wg) * 8; j.pixel = false;
    j.valid = min(8, a.W - j.t0);
    return j;
}
__device__ __forceinline__ int ld_volatile_shared(const volatile int* p) { return *p; }

template <bool kSplitA>
__device__ __forceinline__ void projection_role(const ProjArgs& a, uint8_t* smem, const int blk, const int worker, const int n_workers)
{
    const uint8_t* __restrict__ in_base = a.in_base;
    const int64_t in_wg_stride = a.in_wg_stride, in_dir_stride = a.in_dir_stride, in_part_stride = a.in_part_stride;
    const int lbo = a.lbo, Kp = a.Kp, W = a.W;
    const bool loop = a.jobs != nullptr;                     // chunk-loop kernel: job list, run-time order, flags
    const int blk_bytes = a.blk_bytes;
    const int n_dirs = a.n_dirs;                             // K-slices per stage
    const uint32_t* __restrict__ w_tmem = a.w_tmem;
    const float* __restrict__ scale_row = a.scale_row;
    const float* __restrict__ bias_row = a.bias_row;
    float* __restrict__ gi = a.gi;
    constexpr int PARTS = kSplitA ? 2 : 1;
    // stage: [part hi, lo][K-slice][8 row groups = columns t0..t0+7][blk_bytes]
    const uint32_t slice_bytes = 8u * blk_bytes;
    const uint32_t part_bytes = n_dirs * slice_bytes;
    const uint32_t stage_bytes = PARTS * part_bytes;
    constexpr int n_stages = PROJ_STAGES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + PROJ_STAGES * stage_bytes);
    uint64_t* a_empty = a_full + PROJ_STAGES;
    uint64_t* acc_full = a_empty + PROJ_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    // chunk-loop kernel: the scheduler decides the job order at run time (see below) and passes it on through this ring
    volatile int* job_ring = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [PROJ_RING]
    volatile int* sched_count = job_ring + PROJ_RING;        // jobs the scheduler has published
    volatile int* loader_count = sched_count + 2;            // jobs the loader has taken (+2: keeps what follows 8-byte aligned)
    volatile int* job_tab = loader_count + 1;                // [PROJ_TABLE_MAX], scheduler only: this chunk's job list
    volatile unsigned char* job_done = reinterpret_cast<volatile unsigned char*>(job_tab + PROJ_TABLE_MAX);   // [PROJ_TABLE_MAX], scheduler only
    volatile unsigned long long* enc_count = reinterpret_cast<volatile unsigned long long*>(job_done + PROJ_TABLE_MAX);   // [2 PROJ_GROUPS_MAX], scheduler only

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kwords = Kp >> 1;
    tc::pdl_launch_dependents();
    if (tid == 0) {
        for (int i = 0; i < n_stages; ++i) { tc::mbar_init(a_full + i, 1); tc::mbar_init(a_empty + i, a.pair ? 2 : 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, 4); }
        *sched_count = 0;
        *loader_count = 0;
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == 4) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    tc::named_barrier_sync(1, PROJ_THREADS);
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const bool pixels = loop && a.px.ximg != nullptr;
    const int kwords_px = pixels ? a.px.Kp >> 1 : 0;
    if (warp < 4) {   // weight block(s) -> TMEM, thread = gate row; 32 words in flight per round trip
        const int row = warp * 32 + lane;
        auto upload = [&](const uint32_t* wsrc, int kw, int col0) {
            for (int term = 0; term < 2; ++term) {
                const uint32_t dst = tmem + ((uint32_t)(warp * 32) << 16) + col0 + term * kw;
                int c = 0;
                for (; c + 32 <= kw; c += 32) {
                    uint32_t r[32];
                    const uint4* p = reinterpret_cast<const uint4*>(wsrc + wih_word_index(blk, term, row, c, kw));
#pragma unroll
                    for (int v = 0; v < 8; ++v) { const uint4 x = __ldg(p + v * 32); r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w; }
                    tc::tmem_st16(dst + c, r);
                    tc::tmem_st16(dst + c + 16, r + 16);
                    tc::tmem_st_wait();                      // before r is loaded again (see upload_whh)
                }
                for (; c < kw; c += 16) {
                    uint32_t r[16];
                    const uint4* p = reinterpret_cast<const uint4*>(wsrc + wih_word_index(blk, term, row, c, kw));
#pragma unroll
                    for (int v = 0; v < 4; ++v) { const uint4 x = __ldg(p + v * 32); r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w; }
                    tc::tmem_st16(dst + c, r);
                    tc::tmem_st_wait();
                }
            }
        };
        upload(w_tmem, kwords, PROJ_W_COL0);
        if (pixels) upload(a.px.w_tmem, kwords_px, PX_W_COL0);
        tc::tmem_st_wait();
    }
    tc::pdl_grid_dependency_wait();                          // activations / gi buffers belong to upstream kernels
    tc::tc_fence_before();
    tc::named_barrier_sync(1, PROJ_THREADS);
    tc::tc_fence_after();
    const uint32_t pair_rank = a.pair ? tc::cluster_ctarank() : 0u;
    if (a.pair) tc::cluster_sync_all();                      // the peer's mbarriers exist before anything is multicast at them

    const int n_chunks = a.n_chunks > 0 ? a.n_chunks : 1;
    const int n_table = proj_count(a, worker, n_workers);    // entries of this worker's list
    const int job0 = loop ? __ldg(a.job_offsets + worker) : 0;
    // list = [pixel entries: PX_R relative tiles x my window groups, tile-major][decoder entries in runnable order];
    // chunk k uses the first px_tiles(k) x my_groups pixel entries and all decoder entries
    const int px_wgs = pixels ? __ldg(a.px.wgs_of_worker + worker) : 0;
    const int n_dec = n_table - px_wgs * PX_R;
    auto px_count = [&](int chunk) { return (pixels && chunk + 1 < n_chunks) ? px_wgs * (px_done(a, chunk + 1) - px_done(a, chunk)) : 0; };
    auto table_at = [&](int i, int n_px) { return __ldg(a.jobs + job0 + (i < n_px ? i : px_wgs * PX_R + (i - n_px))); };
    auto jobs_of = [&](int chunk) { return loop ? px_count(chunk) + n_dec : n_table; };
    ProjJob j;
    // -DHB_TIMELINE: worker 0 / block 0 adds up the cycles each role spends at its wait points (slots 7200 + 8 role + k)
#ifdef HB_TIMELINE
    const bool acct = a.dbg != nullptr && worker == 0 && blk == 0 && lane == 0;
    long long t_wait[4] = {0, 0, 0, 0};
    long long t_role0 = acct ? clock64() : 0;
#define HB_TIMED(k, ...) do { const long long t_ = acct ? clock64() : 0; __VA_ARGS__; if (acct) t_wait[k] += clock64() - t_; } while (0)
#define HB_ROLE_REPORT(role) do { if (acct) { for (int k_ = 0; k_ < 4; ++k_) a.dbg[7200 + 8 * (role) + k_] = t_wait[k_]; \
                                              a.dbg[7200 + 8 * (role) + 4] = clock64() - t_role0; } } while (0)
#else
#define HB_TIMED(k, ...) do { __VA_ARGS__; } while (0)
#define HB_ROLE_REPORT(role) do { } while (0)
#endif
    if (warp == 6) {
        // ===================== job scheduler (chunk-loop kernel) =====================
        // Which job next is decided at run time.  A decoder tile is runnable once both encoder directions have published
        // its columns (the middle tiles first, the edge tiles when the encoder ends); the decoder of THIS CTA's direction,
        // however, consumes its tiles from one edge to the other (forward: tile 0 first, reverse: the last tile first), and
        // what it needs first becomes runnable last.  So: among the runnable tiles always take the one this CTA's decoder
        // direction consumes earliest.  The worker's decoder entries are sorted by (tile, group); a forward-direction CTA
        // scans them upwards, a reverse-direction CTA downwards, 32 entries per ballot, against the encoder counters of
        // all window groups (re-read before every decision, all loads in flight at once).  Pixel jobs (the next chunk's
        // encoder projection) wait for nothing and go first: the role is idle at the start of a chunk anyway.
        // The scheduler runs up to PROJ_SCHED_LEAD jobs ahead of the loader, so the global round trips of its counter reads
        // never stall a copy.  (With the decisions in the loader warp itself that warp needed ~1800 cycles per job and
        // starved the MMA issuer, which needs ~1550.  A list walked from both ends - first runnable from the front, last
        // entry as soon as runnable - was the first policy: it served the two decoder directions' first tiles alternately.)
        // The counters are read with relaxed loads: the encoder completed and fenced its bulk stores before it released a
        // counter, and the loader's bulk loads are issued after the value has arrived and read L2 directly.
        if (loop) {
            const int ddir = blk / 3;                        // decoder direction this CTA's gate block belongs to
            const int n_groups = (int)a.n_wg;
            int it = 0;
            for (int chunk = 0; chunk < n_chunks; ++chunk) {
                const int n_px = px_count(chunk);
                const int n_jobs = n_px + n_dec;
                const int px_tile0 = pixels ? px_done(a, chunk) : 0;
                const unsigned long long base = a.epoch + (unsigned long long)chunk * W;
                __syncwarp();
                for (int i = lane; i < n_jobs; i += 32) {    // this chunk's list -> shared memory, pixel entries made absolute
                    int e = table_at(i, n_px);
                    if ((e >> 29) & 1) e += px_tile0 << 16;
                    job_tab[i] = e;
                    job_done[i] = 0;
                }
                __syncwarp();
                int lo = 0, hi = n_dec - 1, px_next = 0;     // undone decoder entries lie in [lo, hi]
                for (int idx = 0; idx < n_jobs; ++idx, ++it) {
                    if (lane == 0) {
                        const long long t_spin = clock64();
                        HB_TIMED(0, while (it - ld_volatile_shared(loader_count) >= PROJ_SCHED_LEAD) { if (clock64() - t_spin > tc::SPIN_LIMIT_CYCLES) __trap(); });
                    }
                    __syncwarp();
                    int e = 0;
                    // pixel jobs (the NEXT chunk's encoder projection) wait for nothing.  Small batches take them after the
                    // chunk's decoder tiles: written half a chunk before their first use instead of a chunk and a half, the gi
                    // rows are still in L2 when the encoder reads them (ncu at B=256: DRAM reads of the launch 1.52 -> 0.89 GB,
                    // same speed).  With fewer than 12 workers, or 16 windows per recurrence CTA, the role has no such slack
                    // (B=320: -5 %, B=512: -3 %): there they go first, into the idle start of the chunk.
                    if (a.px.last ? lo > hi : px_next < n_px) {
                        e = job_tab[px_next++];
                    } else {
#ifdef HB_TIMELINE
                        const long long t_ = acct ? clock64() : 0;
#endif
                        const long long t_spin = clock64();
                        int found = -1;
                        while (true) {
                            for (int i = lane; i < 2 * n_groups; i += 32)
                                enc_count[i] = tc::ld_relaxed_gpu(a.progress + (((i >> 1) * WG) / a.rec_n) * 2 + (i & 1));
                            __syncwarp();
                            for (int k0 = 0; k0 <= hi - lo && found < 0; k0 += 32) {
                                const int k = k0 + lane;
                                const int pos = ddir == 0 ? lo + k : hi - k;
                                bool ok = false;
                                if (k <= hi - lo && !job_done[pos]) {
                                    const ProjJob q = proj_decode(a, job_tab[n_px + pos]);
                                    ok = enc_count[q.wg * 2] >= base + (unsigned long long)(q.t0 + q.valid) &&
                                         enc_count[q.wg * 2 + 1] >= base + (unsigned long long)(W - q.t0);
                                }
                                const unsigned m = __ballot_sync(0xffffffffu, ok);
                                if (m) { const int kk = k0 + __ffs(m) - 1; found = ddir == 0 ? lo + kk : hi - kk; }
                            }
                            if (found >= 0) break;
                            __nanosleep(200);
                            if (clock64() - t_spin > tc::SPIN_LIMIT_CYCLES) __trap();
                        }
#ifdef HB_TIMELINE
                        if (acct) t_wait[1] += clock64() - t_;
#endif
                        e = job_tab[n_px + found];
                        HB_CHAIN(a.dbg, worker == 0 && blk == 0 && chunk == 2 && lane == 0 && proj_decode(a, e).wg == 0 && proj_decode(a, e).t0 == 0, 2);
                        {   // bit 30: one direction of this group's encoder has finished, so the tile is one the decoder starts with (the
                            // forward decoder's first tile is gated by the REVERSE encoder's last columns) or the decoder is
                            // already running: its counters go out at once, not with the next batch.  (Requiring both directions
                            // left the decoder's very first tile in a batch whenever the other encoder direction published its
                            // last columns a microsecond later: 4 us between "counter raised" and "decoder saw it".)
                            const ProjJob q = proj_decode(a, e);
                            if (enc_count[q.wg * 2] >= base + (unsigned long long)W || enc_count[q.wg * 2 + 1] >= base + (unsigned long long)W) e |= 1 << 30;
                        }
                        __syncwarp();
                        if (lane == 0) job_done[found] = 1;
                        __syncwarp();
                        while (lo <= hi && job_done[lo]) ++lo;
                        while (hi >= lo && job_done[hi]) --hi;
                    }
                    if (lane == 0) {
                        job_ring[it & (PROJ_RING - 1)] = e;
                        __threadfence_block();
                        *sched_count = it + 1;
                    }
                }
            }
            HB_ROLE_REPORT(3);
        }
    } else if (warp == 5) {
        // ===================== loader =====================
        int it = 0;
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
        const int n_jobs = jobs_of(chunk);
        for (int idx = 0; idx < n_jobs; ++idx, ++it) {
            const int stage = it % n_stages;
            if (it >= n_stages) HB_TIMED(0, tc::mbar_wait(a_empty + stage, (uint32_t)((it / n_stages - 1) & 1)));
            if (loop) {
                int e = 0;
                if (lane == 0) {
                    const long long t_spin = clock64();
                    HB_TIMED(1, while (ld_volatile_shared(sched_count) <= it) { if (clock64() - t_spin > tc::SPIN_LIMIT_CYCLES) __trap(); });
                    __threadfence_block();
                    e = job_ring[it & (PROJ_RING - 1)];
                }
                e = __shfl_sync(0xffffffffu, e, 0);
                j = proj_decode(a, e);

            } else {
                j = proj_tile_job(a, worker, n_workers, idx);
            }
#ifdef HB_TIMELINE
            const long long t_issue = acct ? clock64() : 0;
#endif
            // geometry of the job's stage: decoder jobs [part hi, lo][K-slice][8 columns][2 KB], pixel jobs [8 columns][xblk].
            // The job's columns are contiguous per (part, K-slice): ONE copy each (issuing a bulk copy blocks the lane for
            // about its size / 48 B per cycle, and the lanes take turns).  A single-slice job is cut in two so that a CTA
            // pair always has a copy each to multicast.
            const int jparts = j.pixel ? 1 : PARTS, jdirs = j.pixel ? 1 : n_dirs, jblk = j.pixel ? a.px.blk_bytes : blk_bytes;
            const int pieces = jparts * jdirs == 1 ? 2 : 1;
            if (lane == 0) tc::mbar_arrive_expect_tx(a_full + stage, (uint32_t)(j.valid * jparts * jdirs * jblk));
            __syncwarp();
            if (lane < pieces * jparts * jdirs) {
                const int piece = lane % pieces, pd = lane / pieces, part = pd / jdirs, d = pd % jdirs;
                const int c0 = piece * 4, cols = pieces == 1 ? j.valid : min(4, j.valid - c0);
                uint8_t* dst = smem + stage * stage_bytes + (j.pixel ? 0 : part * part_bytes + d * slice_bytes) + c0 * jblk;
                const uint8_t* src = j.pixel
                    ? a.px.ximg + j.wg * a.px.wg_stride + (int64_t)(j.t0 + c0) * jblk
                    : in_base + j.wg * in_wg_stride + d * in_dir_stride + part * in_part_stride + (int64_t)(j.t0 + c0) * jblk;
                const uint32_t bytes = (uint32_t)(max(cols, 0) * jblk);
                // pair mode: the two CTAs split the copies and multicast them to both
                const bool mine = cols > 0 && (!a.pair || (uint32_t)(lane & 1) == pair_rank);
                if (mine) {
                    if (a.pair) tc::bulk_g2s_multicast(dst, src, bytes, a_full + stage, (uint16_t)3);
                    else tc::bulk_g2s(dst, src, bytes, a_full + stage);
                }
            }
            __syncwarp();
            HB_CHAIN(a.dbg, loop && worker == 0 && blk == 0 && chunk == 2 && lane == 0 && !j.pixel && j.wg == 0 && j.t0 == 0, 3);
            if (loop && lane == 0) *loader_count = it + 1;
#ifdef HB_TIMELINE
            if (acct) t_wait[3] += clock64() - t_issue;
#endif
        }
        }
        HB_ROLE_REPORT(0);
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = tc::idesc_f16_f32(128, PROJ_NT);
        const int ksteps = Kp >> 4;
        int it = 0;
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
        const int n_jobs = jobs_of(chunk);
        for (int idx = 0; idx < n_jobs; ++idx, ++it) {
            const int stage = it % n_stages, acc = it & 1;
            HB_TIMED(0, tc::mbar_wait(a_full + stage, (uint32_t)((it / n_stages) & 1)));
            j = loop ? proj_decode(a, job_ring[it & (PROJ_RING - 1)]) : proj_tile_job(a, worker, n_workers, idx);
            if (it >= 2) HB_TIMED(1, tc::mbar_wait(acc_empty + acc, (uint32_t)((it / 2 - 1) & 1)));
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t sbase = tc::smem_u32(smem + stage * stage_bytes);
                const uint64_t d_hi = lbo ? tc::smem_desc(sbase, lbo, blk_bytes) : tc::smem_desc_sw128(sbase, blk_bytes);
                const uint64_t d_lo = lbo ? tc::smem_desc(sbase + part_bytes, lbo, blk_bytes) : tc::smem_desc_sw128(sbase + part_bytes, blk_bytes);
                const uint32_t a_hi = tmem + PROJ_W_COL0, a_lo = a_hi + kwords;
                const uint32_t d = tmem + acc * PROJ_NT;
                // Issue loops with compile-time trip counts for the hot shape: the operands of every MMA are then
                // immediates on the uniform datapath (a runtime k loop went through R2UR moves: ~80 cycles per MMA).
                // k-step ks of the stage: K-slice ks / 8 (a slice of the GRU output image is 128 units), then 16 units per step
                auto issue = [&](auto ksteps_c, auto dirs_c) {
                    constexpr int KS = decltype(ksteps_c)::value, ND = decltype(dirs_c)::value;
                    const int ks_n = KS > 0 ? KS : ksteps;
                    auto koff = [&](int ks) {
                        if (KS == 0) return (uint64_t)(ks * 2 * lbo / 16);                         // pixels: no-swizzle image
                        return (uint64_t)((ND == 2 ? (ks >> 3) * slice_bytes : 0) / 16) + tc::sw128_kstep(ks & 7);
                    };
#pragma unroll
                    for (int ks = 0; ks < ks_n; ++ks) tc::mma_f16_ts(d, a_hi + ks * 8, d_hi + koff(ks), idesc, ks != 0);
#pragma unroll
                    for (int ks = 0; ks < ks_n; ++ks) tc::mma_f16_ts(d, a_lo + ks * 8, d_hi + koff(ks), idesc, 1);
                    if (kSplitA) {
#pragma unroll
                        for (int ks = 0; ks < ks_n; ++ks) tc::mma_f16_ts(d, a_hi + ks * 8, d_lo + koff(ks), idesc, 1);
                    }
                };
                if (j.pixel) {                                // pixels: no-swizzle image, 2 terms (uint8 is exact in fp16)
                    const uint64_t dx = tc::smem_desc(sbase, 128, a.px.blk_bytes);
                    const uint32_t ax = tmem + PX_W_COL0;
                    const int ksx = a.px.Kp >> 4;
                    for (int ks = 0; ks < ksx; ++ks) tc::mma_f16_ts(d, ax + ks * 8, dx + (uint64_t)(ks * 2 * 128 / 16), idesc, ks != 0);
                    for (int ks = 0; ks < ksx; ++ks) tc::mma_f16_ts(d, ax + kwords_px + ks * 8, dx + (uint64_t)(ks * 2 * 128 / 16), idesc, 1);
                } else if (n_dirs == 2 && ksteps == 16) issue(std::integral_constant<int, 16>{}, std::integral_constant<int, 2>{});
                else issue(std::integral_constant<int, 0>{}, std::integral_constant<int, 1>{});     // pixels: K = padded feature count
                if (a.pair) tc::mma_commit_multicast(a_empty + stage, (uint16_t)3);   // both loaders wait for both consumers
                else tc::mma_commit(a_empty + stage);
                tc::mma_commit(acc_full + acc);
            }
            __syncwarp();
            HB_CHAIN(a.dbg, loop && worker == 0 && blk == 0 && chunk == 2 && lane == 0 && !j.pixel && j.wg == 0 && j.t0 == 0, 4);
        }
        }
        HB_ROLE_REPORT(1);
    } else {
        // ===================== epilogue: group 0 (warps 0-3) takes the even jobs / accumulator 0, group 1 (warps 7-10) the odd ones
        const int grp = warp < 4 ? 0 : 1;
        const int r = (warp & 3) * 32 + lane;                // gate row within the block == TMEM lane
        const float sc_dec = scale_row[blk * 128 + r], bi_dec = bias_row[blk * 128 + r];
        const float sc_px = pixels ? a.px.scale_row[blk * 128 + r] : 0.f, bi_px = pixels ? a.px.bias_row[blk * 128 + r] : 0.f;
        const int tiles_t = (W + 7) >> 3;
        // this row's 8 windows of a column: 32 bytes at r * 32, the two 16-byte halves swapped for rows with bit 2 set
        // (gi_window_index: the recurrence's 16-byte shared-memory reads are then conflict-free)
        const int swz = (r >> 2) & 1;
        // Counters are raised in batches: the gpu-scope fence that makes a tile's stores visible waits for the newest
        // stores to be acknowledged (~2000 cycles under load), so one fence covers up to PROJ_FLAG_BATCH tiles - except
        // for the tiles the decoder is waiting for (taken from the back of the job list) and the group's last tile of a
        // chunk, which go out at once together with everything pending.
        unsigned long long* pending[PROJ_FLAG_BATCH];
        int n_pending = 0;
        int it = 0;
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
        const int n_jobs = jobs_of(chunk);
        for (int idx = 0; idx < n_jobs; ++idx, ++it) {
            const int acc = it & 1;
            if (acc != grp) continue;
            HB_TIMED(0, tc::mbar_wait(acc_full + acc, (uint32_t)((it / 2) & 1)));
            j = loop ? proj_decode(a, job_ring[it & (PROJ_RING - 1)]) : proj_tile_job(a, worker, n_workers, idx);
            tc::tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + acc * PROJ_NT;
            const float sc = j.pixel ? sc_px : sc_dec;
            const float add = j.pixel ? bi_px : bi_dec;
            const int out_cols = j.pixel ? a.px.cols : W;
            float* dst = (j.pixel ? a.px.gi : gi) + gi_block(j.wg, out_cols, j.t0, blk) + r * WG;
            // 16 accumulator columns (two image columns x 8 windows) per TMEM round trip, the next load in flight while
            // the current values are scaled and stored
            {
                float v[2][16];
                tc::tmem_ld16(taddr, v[0]);
#pragma unroll
                for (int c16 = 0; c16 < PROJ_NT / 16; ++c16) {
                    tc::tmem_ld_wait();
                    if (c16 + 1 < PROJ_NT / 16) tc::tmem_ld16(taddr + (c16 + 1) * 16, v[(c16 + 1) & 1]);
                    else {                                   // every accumulator column is in registers: hand the buffer back
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(acc_empty + acc);
                    }
                    const float* x = v[c16 & 1];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int col = 2 * c16 + h;
                        if (col < j.valid) {
                            float4* p = reinterpret_cast<float4*>(dst + (size_t)col * GI_BLK_FLOATS);
                            // chunk-loop kernel: the decoder reads this tile within the chunk - keep it in L2 until then
                            if (loop && !j.pixel) {
                                tc::st_global_v4_hint(p + swz, make_float4(fmaf(x[8 * h], sc, add), fmaf(x[8 * h + 1], sc, add), fmaf(x[8 * h + 2], sc, add), fmaf(x[8 * h + 3], sc, add)), tc::L2_EVICT_LAST);
                                tc::st_global_v4_hint(p + (swz ^ 1), make_float4(fmaf(x[8 * h + 4], sc, add), fmaf(x[8 * h + 5], sc, add), fmaf(x[8 * h + 6], sc, add), fmaf(x[8 * h + 7], sc, add)), tc::L2_EVICT_LAST);
                                continue;
                            }
                            p[swz] = make_float4(fmaf(x[8 * h], sc, add), fmaf(x[8 * h + 1], sc, add), fmaf(x[8 * h + 2], sc, add), fmaf(x[8 * h + 3], sc, add));
                            p[swz ^ 1] = make_float4(fmaf(x[8 * h + 4], sc, add), fmaf(x[8 * h + 5], sc, add), fmaf(x[8 * h + 6], sc, add), fmaf(x[8 * h + 7], sc, add));
                        }
                    }
                }
            }
            if (a.tile_flags != nullptr) {
                pending[n_pending++] = j.pixel ? a.px.flags + ((j.wg * a.px.tiles + (j.t0 >> 3)) * 2 + blk / 3)
                                               : a.tile_flags + ((j.wg * tiles_t + (j.t0 >> 3)) * 2 + blk / 3);
                const bool urgent = loop && ((job_ring[it & (PROJ_RING - 1)] >> 30) & 1);
                if (urgent || n_pending == PROJ_FLAG_BATCH || idx + 2 >= n_jobs) {
                    // this warp's part of the tiles is visible before their counters move: every lane fences its own
                    // stores, the warp joins, one lane counts
                    HB_TIMED(1, __threadfence());
                    __syncwarp();
                    if (lane == 0)
                        for (int k = 0; k < n_pending; ++k) tc::red_relaxed_gpu_add(pending[k], 1ull);
                    n_pending = 0;
                    HB_CHAIN(a.dbg, worker == 0 && blk == 0 && chunk == 2 && (warp & 3) == 0 && lane == 0 && !j.pixel && j.wg == 0 && j.t0 == 0, 5);
                }
            }
            if (a.tile_reads != nullptr && !j.pixel && (warp & 3) == 0) {
                // the sixth gate-block CTA to finish a tile knows that nobody reads the encoder's output columns of it again
                // before the next chunk's encoder overwrites them: L2 may drop those lines (8 columns x 2 directions x hi, lo
                // x 2 KB) instead of writing them to DRAM.  (This CTA's own copy of the tile left global memory before the MMAs
                // ran; the other CTAs count only after their epilogue, too.)
                unsigned long long before = 0;
                if (lane == 0) before = atomicAdd(a.tile_reads + (j.wg * tiles_t + (j.t0 >> 3)), 1ull);
                before = __shfl_sync(0xffffffffu, before, 0);
                if ((before + 1) % 6 == 0) {
                    for (int dd = 0; dd < 2; ++dd)
                        for (int part = 0; part < 2; ++part) {
                            const uint8_t* tile = in_base + j.wg * in_wg_stride + dd * in_dir_stride + part * in_part_stride + (int64_t)j.t0 * blk_bytes;
                            for (int l = lane; l < j.valid * (blk_bytes / 128); l += 32) tc::discard_l2_line(tile + (int64_t)l * 128);
                        }
                }
            }
#ifdef HB_TIMELINE
            if (a.dbg != nullptr && worker == 0 && blk == 0 && (warp & 3) == 0 && lane == 0) a.dbg[7000 + chunk] = (long long)globaltimer_ns();
#endif
        }
        }
        if (warp == 0) HB_ROLE_REPORT(2);
    }
#undef HB_TIMED
#undef HB_ROLE_REPORT
    tc::tc_fence_before();
    tc::named_barrier_sync(1, PROJ_THREADS);
    if (a.pair) tc::cluster_sync_all();                      // no multicast / remote arrive may target a CTA that has exited
    if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

template <bool kSplitA>
__global__ void __launch_bounds__(PROJ_THREADS, 1)
tc_projection_kernel(const ProjArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem_proj[];
    projection_role<kSplitA>(a, tc::align_smem_1024(smem_proj), (int)blockIdx.y, (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------
// Recurrence.  One CTA = N windows x one direction x W dependent steps of one layer.
//   * W_hh (fp16 hi and lo, 3 gate blocks x 128 rows x 128 k) is loaded ONCE into TMEM and is the
//     A operand of every MMA (384 of the 512 columns); accumulators r | z | n at columns [0, 3N).
//   * the state h lives in registers (fp32 * 2^10, one hidden unit per thread) and is re-published
//     each step as the fp16 hi/lo B operand image in shared memory; that same image is the layer's
//     output for column t and is bulk-stored to yimg by the store warp.
//   * gi' rows of the step are prefetched GI_STAGES steps ahead into shared memory by bulk copies.
// gi' (from tc_projection_kernel) already contains, per gate row,
//     r, z rows:  -log2e * (W_i. x + b_i. + b_h.)        n rows:  -2 log2e * (W_in x + b_in)
// With  er = 2^(gi'_r + acc_r inv_r),  ez = 2^(gi'_z + acc_z inv_z),  e = 2^(gi'_n + r (acc_n inv_n + b_hn')):
//     r = 1 / (1 + er),   z = 1 / (1 + ez),   n = tanh(..) = (1 - e) / (1 + e),   h' = (1 - z) n + z h,
// and the last two reciprocals are taken as ONE:
//     h' 2^10 = [2^10 ez (1 - e) + (h 2^10) (1 + e)] / [(1 + e) (1 + ez)]
// (5 MUFU operations per element and step instead of 6; the exponents are clamped at 2^50 so the products stay finite:
// 1 / (1 + 2^50) is zero to fp32 precision anyway).
// GW gate warps; (warp w, lane l) owns hidden unit j = 32 (w%4) + l (== its TMEM lane) for NW = 4 NLIVE / GW windows
// starting at (w/4) NW.  Warp GW: MMA issuer.  Warp GW+1: gi loader.  Warp GW+2: y store.
// The gate warps are issue-bound (every scheduler runs GW/4 of them plus at most one helper warp, and a step is ~100
// instructions per gate warp): the loop below keeps everything that does not change from step to step in registers
// and the timeline instrumentation is compiled in only with -DHB_TIMELINE.
// ---------------------------------------------------------------------------------------------
constexpr int REC_GATE_WARPS = 16;
constexpr int REC_TC_THREADS = (REC_GATE_WARPS + 3) * 32;
constexpr int REC_W_COL0 = 128;                   // weight columns start here (accumulators below)
constexpr int WHH_TMEM_WORDS = 2 * 3 * 128 * 64;  // per direction: [term][gate block][row][k pair]
constexpr float EXP2_CLAMP = 50.0f;
// gi' stages (one step each, NLIVE x 1536 B): 8 live: 6 x 12 KB (the other roles' traffic delays gi' rows in the
// chunk-loop kernel), 16 live: 4 x 24 KB, 32 live: 3 x 48 KB (227 KB smem limit)
template <int NLIVE> __host__ __device__ constexpr int gi_stages() { return NLIVE <= 8 ? 6 : (NLIVE <= 16 ? 4 : 3); }
// h operand image buffers: 4 give the y store three steps to drain, which hides the global-memory round trips of the
// progress publication (chunk-loop kernel); N = 32 has room for 2 only
template <int N> __host__ __device__ constexpr int h_buffers() { return N <= 16 ? 4 : 2; }
constexpr int PUBLISH_LAG = 4;
// The y-store warp publishes its progress counter when the columns stored (and landed) reach a multiple of 8 counted from
// the edge its direction starts at - the consumers' tile boundaries: projection tiles are 8 columns, heads tiles 16, and the
// reverse direction's count of W - 8 t columns is congruent to W modulo 8.  The release store costs the warp 3-7 thousand
// cycles (its MEMBAR.GPU also waits for the newest bulk stores, still in flight): once per 8 steps, not per 4.
// (A relaxed store after cp.async.bulk.wait_group is NOT enough: tests/test_gpu_stress.py caught batches that differed
// from run to run with it.)  Nothing is published in the last 8 steps of a phase: the final publication follows anyway,
// and a warp that has just spent a release store there hands the phase's last images to the copy engine late - the other
// layer's first tiles wait for exactly those.
#ifndef HB_PUBLISH_EVERY
#define HB_PUBLISH_EVERY 8
#endif
constexpr int PUBLISH_EVERY = HB_PUBLISH_EVERY;
constexpr int REC_STEP_BARRIER = 3;               // named barrier: gate warps -> MMA issuer, once per step (0 = __syncthreads, 1 / 2 = projection / heads roles)

// One GRU layer as the recurrence role sees it.
struct RecLayer {
    const float* gi;               // gi image (see gi_block); the step at column t of chunk k reads column gi_col0 + k * gi_col_step + t
    int gi_cols, gi_col0, gi_col_step;
    const uint32_t* whh_tmem;      // packed fp16 pairs, see whh_word_index
    const float* gate_consts;      // [2 dirs][4][128]: inv_r', inv_z', inv_n', b_hn'  (per unit)
    uint8_t* yimg[2];              // operand image of the layer output: [0] even chunks, [1] odd chunks
    unsigned long long* progress;  // [ctas][2 dirs] columns whose output has landed in yimg (+ epoch + chunk * W), or nullptr
    // chunk-loop kernel only (counters in global memory, see tc_chunkloop_kernel):
    // gi' tiles are announced by the projection role: tile (group, column tile, dir) is ready for chunk k when its counter
    // is >= flag_need_base + flag_need_per_chunk (k + 1).  Decoder: tiles of the chunk's own columns, 6 per chunk (3 gate
    // blocks x 2 K-halves).  Encoder (pixel jobs): tiles of ABSOLUTE image columns, 3 once; tiles below flag_skip_tiles were
    // projected before the kernel started.
    const unsigned long long* tile_flags; int flag_tiles, flag_abs, flag_skip_tiles; unsigned long long flag_need_base, flag_need_per_chunk;
    const unsigned long long* heads_done; int heads_per_chunk;   // decoder: yimg[k & 1] reusable when >= heads_per_chunk (k - 1)
    const unsigned long long* consumed_flags;  // encoder: tile flags (both directions) that tell yimg of chunk k - 1 has been read
    int consumed_per_chunk;                    // ... when they have reached consumed_per_chunk * k
};

// n_layers == 1: one layer, one chunk (per-chunk launches).  n_layers == 2: the whole chunk loop of the reference
// (predict_gpu.py:114-149) in one CTA: encoder and decoder phases alternate, the state h stays in registers from phase
// to phase (decoder h_0 = encoder h_n, next encoder h_0 = decoder h_n, TransducerModel.py:68-78) and only W_hh is
// re-uploaded into TMEM at every phase switch.
struct RecArgs {
    RecLayer layer[2];
    int n_layers, n_chunks;
    const float* h_in;             // [B, 2, 128] fp32 or nullptr (zeros)
    float* h_out;                  // [B, 2, 128] fp32 or nullptr
    int64_t B; int W;
    int tiles_t;                   // 8-column tiles per chunk (tile_flags indexing)
    int n_wg;                      // existing window groups (flags of groups past it are never raised)
    unsigned long long epoch;
    long long* dbg;
    int dbg_layer;                 // -DHB_TIMELINE: which layer's steps are recorded
    long long* phase_times;        // HB_PHASE_TIMES=1 (any build): CTA 0 / forward records [phase][start, first step released, last step done, end] in ns
};

// W_hh of one direction -> TMEM, spread over `n_warps` gate warps (a multiple of 4).  A warp covers TMEM lanes
// 32 (w%4)..+31 (thread = gate row); the warps of a lane quarter split the four (hi | lo image) x (k-pair columns 0-31 |
// 32-63) pieces; 32 words in flight per round trip, two register buffers so that the loads of gate block gb+1 are in
// flight while block gb is stored.  A buffer is loaded again only after tcgen05.wait::st (the TMEM store reads its
// source registers asynchronously).
// Two variants of this routine gave wrong recurrences and are not used: as a separate (not inlined) function, meant to
// keep its 64 staging registers out of the caller's allocation, the chunk-loop kernel miscomputed a step about once in 50
// launches with 8 windows per gate thread (inlined: 0 in 400 launches of every tile / gate-warp combination); with ONE
// staging buffer in a rolled loop over the gate blocks the 32-window per-chunk kernel was wrong in every launch.  Neither
// was root-caused; tests/test_gpu_stress.py repeats every kernel variant a few hundred times to catch this class of fault.
// WORDS = 32 or 16 staging registers per buffer (16: for the 19-warp blocks, whose 96-register budget the 64 staging
// registers of the wide version pushed into local memory - together with loop-invariant values of the step loop).
template <int WORDS = 32>
__device__ __forceinline__ void upload_whh(const uint32_t* __restrict__ whh_tmem, const int dir, const uint32_t tmem, const int warp,
                                           const int lane, const int n_warps)
{
    static_assert(WORDS == 32 || WORDS == 16, "staging registers per buffer");
    constexpr int SUBS = 32 / WORDS;
    const int q = warp & 3, row = q * 32 + lane;
    for (int piece = warp >> 2; piece < 4; piece += n_warps >> 2) {
        const int term = piece & 1, half = piece >> 1;
#pragma unroll 1
        for (int sub = 0; sub < SUBS; ++sub) {
            auto fetch = [&](int gb, uint32_t* r) {
                const uint4* p = reinterpret_cast<const uint4*>(whh_tmem + whh_word_index(dir, term, gb, row, half * 32)) + sub * (WORDS / 4) * 32;
#pragma unroll
                for (int v = 0; v < WORDS / 4; ++v) { const uint4 x = __ldg(p + v * 32); r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w; }
            };
            auto store = [&](int gb, const uint32_t* r) {
                const uint32_t dst = tmem + ((uint32_t)(q * 32) << 16) + REC_W_COL0 + (term * 3 + gb) * 64 + half * 32 + sub * WORDS;
                tc::tmem_st16(dst, r);
                if constexpr (WORDS == 32) tc::tmem_st16(dst + 16, r + 16);
            };
            uint32_t ra0[WORDS], ra1[WORDS];
            fetch(0, ra0);
            fetch(1, ra1);
            store(0, ra0);
            tc::tmem_st_wait();
            fetch(2, ra0);
            store(1, ra1);
            store(2, ra0);
            tc::tmem_st_wait();
        }
    }
}

// One GRU step of one element, split in the three phases in which the accumulators arrive.  All values in the 2^10
// scale of the state.
struct GateR { float c1, c2; };                   // r inv_n,  r b_hn' + gi'_n
__device__ __forceinline__ GateR gate_r(float acc, float inv_r, float gir, float inv_n, float bhn, float gin) {
    const float r = tc::rcp_approx(1.0f + tc::ex2_approx(fmaf(acc, inv_r, gir)));
    return GateR{r * inv_n, fmaf(r, bhn, gin)};
}
struct GateZ { float q, k; };                     // 1 + ez,  2^10 ez
__device__ __forceinline__ GateZ gate_z(float acc, float inv_z, float giz) {
    const float ez = tc::ex2_approx(fminf(fmaf(acc, inv_z, giz), EXP2_CLAMP));
    return GateZ{1.0f + ez, ez * ACT_SCALE};
}
__device__ __forceinline__ float gate_n(float acc, const GateR& gr, const GateZ& gz, float h) {

    const float e = tc::ex2_approx(fminf(fmaf(acc, gr.c1, gr.c2), EXP2_CLAMP));
    const float p = 1.0f + e;
    const float num = fmaf(h, p, fmaf(-e, gz.k, gz.k));           // h (1 + e) + 2^10 ez (1 - e)
    return num * tc::rcp_approx(p * gz.q);
}

// NLIVE <= N windows of the N accumulator columns are real: the MMA shape needs N >= 16, but with 8 live windows
// per CTA a small batch spreads over twice as many SMs and the exposed gate math halves.
//
// STACK: the hi and lo images of h are read as ONE B operand, [h_hi (NLIVE rows) | h_lo (NLIVE rows)], so a step needs
// only the two A terms W_hi and W_lo (48 MMAs instead of 72; the MMA time of a step is set by the instruction count,
// not by N, at these sizes).  Accumulator columns [0, NLIVE) hold W.h_hi, [NLIVE, 2 NLIVE) hold W.h_lo; the gate
// threads add the two.  That is the full 4-term product (W_lo.h_lo included).
// (Measured and dropped: two accumulators per gate block, W_hi and W_lo MMAs issued alternately so that consecutive MMAs
// do not accumulate into the same tile - the MMAs were no faster and the extra TMEM loads cost 90 cycles per step.)
//
// MODE 0: 3-term.  MODE 1: STACK.
template <int N, int NLIVE, int MODE = 0, int GW = REC_GATE_WARPS>
__device__ __forceinline__ void recurrence_role(const RecArgs& ra, uint8_t* smem, const int cta_x, const int dir)
{
    const int64_t B = ra.B;
    const int W = ra.W;
    const int n_layers = ra.n_layers;
    const int n_phases = (ra.n_chunks > 0 ? ra.n_chunks : 1) * n_layers;
    static_assert(N == 16 || N == 32, "N accumulator columns per gate block (3N must stay below REC_W_COL0)");
    static_assert(NLIVE == N || (N == 16 && NLIVE == 8), "live windows per CTA");
    static_assert(GW == 8 || GW == 16, "gate warps");
    constexpr bool STACK = MODE >= 1;
    static_assert(!STACK || N == 16, "stacked operand: 3 x 2 NLIVE accumulator columns must stay below REC_W_COL0");
    constexpr int NTHREADS = (GW + 3) * 32;
    constexpr int NACC = STACK ? 2 * NLIVE : N;              // N of the MMA == accumulator columns per gate block
    constexpr int NBLK = NACC;
    constexpr int NW = NLIVE * 4 / GW;                       // windows per gate thread
    static_assert(NW == 2 || NW == 4 || NW == 8, "a gate thread's windows lie in one window group");
    constexpr int NG = NLIVE / WG;                           // live window groups per CTA
    // one h operand image (hi or lo): all N columns - or, stacked with 8 live windows, just the one 8-row group that exists
    // (rows 8-15 of the B operand are then the lo image, which directly follows the hi image)
    constexpr bool COMPACT = STACK && NLIVE == 8;
    constexpr uint32_t HB_BYTES = COMPACT ? YBLK : (N / WG) * YBLK;
    constexpr uint32_t GI_STAGE_BYTES = NLIVE * GI_ROW_BYTES;
    constexpr int GI_STAGES = gi_stages<NLIVE>();
    // h operand images, NBUF buffers: step s reads buffer s % NBUF (h_s) and writes buffer (s+1) % NBUF (h_{s+1}),
    // so the y store of an image has NBUF steps to drain before the buffer is written again
    // (compact images: twice the buffers in the same shared memory.  The y-store warp is away for 3-7 thousand cycles at
    // every publication of its progress counter - see PUBLISH_EVERY - and hands no buffer back meanwhile)
    constexpr int NBUF = COMPACT ? 2 * h_buffers<N>() : h_buffers<N>();
    uint8_t* h_img = smem;                                   // [NBUF buffers][hi, lo][HB_BYTES]
    uint8_t* gi_s = smem + 2 * NBUF * HB_BYTES;
    // (Measured and dropped: one commit barrier per group of four gate warps - the extra commits delayed every
    // accumulator by ~90 cycles and the last gate warp was as late as before.)
    uint64_t* acc_ready = reinterpret_cast<uint64_t*>(gi_s + GI_STAGES * GI_STAGE_BYTES);   // [3]: r, z, n blocks
    uint64_t* h_ready = acc_ready + 3;
    uint64_t* h_free = h_ready + 1;                          // [NBUF], one per buffer: its y store has drained
    uint64_t* y_ready = h_free + NBUF;                       // [NBUF], one per buffer: image complete (for the store warp;
                                                             // per-buffer so a lagging store warp cannot alias phases)
    uint64_t* gi_full = y_ready + NBUF;
    uint64_t* gi_empty = gi_full + GI_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gi_empty + GI_STAGES);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // (tells the compiler the warp index is uniform)
    const int64_t b0 = (int64_t)cta_x * NLIVE;
    const int t_first = dir ? W - 1 : 0, dt = dir ? -1 : 1;
    // -DHB_TIMELINE: phase-level stamps and the roles' wait accounting (a handful of instructions outside the step loops);
    // -DHB_TIMELINE_STEPS adds per-step stamps in the MMA issuer and the gate warps, which lengthen a step by ~25 %
#ifdef HB_TIMELINE
    long long* __restrict__ dbg = ra.dbg;
#endif
#ifdef HB_TIMELINE_STEPS
#define HB_DBG(role, s, k) do { if (dbg_steps && lane == 0) dbg[(((role) * 128 + (s)) * 8) + (k)] = clock64(); } while (0)
#else
#define HB_DBG(role, s, k) do { } while (0)
#endif

    tc::pdl_launch_dependents();
    if constexpr (NLIVE < N) {                               // dead accumulator columns: keep their operand rows finite
        for (uint32_t i = tid; i < 2 * NBUF * HB_BYTES / 16; i += NTHREADS) reinterpret_cast<int4*>(h_img)[i] = make_int4(0, 0, 0, 0);
        tc::fence_proxy_async_smem();
    }
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) tc::mbar_init(acc_ready + i, 1);
        tc::mbar_init(h_ready, GW);
        for (int i = 0; i < NBUF; ++i) { tc::mbar_init(h_free + i, 1); tc::mbar_init(y_ready + i, GW); }
        for (int i = 0; i < GI_STAGES; ++i) { tc::mbar_init(gi_full + i, 1); tc::mbar_init(gi_empty + i, GW); }
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == GW) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // gate-thread state that lives across phases: (warp w, lane l) owns hidden unit j = 32 (w%4) + l for NW windows
    const int q = warp & 3;
    const int j = q * 32 + lane;
    const int win0 = (warp >> 2) * NW;
    float h_own[NW];                                         // h * 2^10
#pragma unroll
    for (int i = 0; i < NW; ++i) h_own[i] = 0.f;

    for (int phase = 0; phase < n_phases; ++phase) {
    const int chunk = phase / n_layers, li = phase - chunk * n_layers;
    const RecLayer& L = ra.layer[li];
    const float* __restrict__ gi = L.gi;
    const int gi_cols = L.gi_cols;
    const int gi_col0 = L.gi_col0 + chunk * L.gi_col_step;
    uint8_t* __restrict__ yimg = L.yimg[chunk & 1];
    const unsigned long long prog_base = ra.epoch + (unsigned long long)chunk * W;
#ifdef HB_TIMELINE
    const bool dbg_on = dbg != nullptr && cta_x == 0 && dir == 0 && (n_layers == 1 ? li == ra.dbg_layer : chunk == 2);
#endif
#ifdef HB_TIMELINE_STEPS
    const bool dbg_steps = dbg_on && li == ra.dbg_layer;
#endif
    if (phase > 0) {
        // phase boundary: fresh barriers.  The cross-CTA conditions of this phase (heads done with the y image it overwrites,
        // projection done with the previous one) were awaited by the gi loader warp at the end of the previous phase, while
        // the last steps ran.  (One thread re-initialises all barriers: with one thread per barrier the kernel failed with
        // a launch error - not root-caused - and the serial version costs ~0.2 us.)
        if (tid == 0) {
            for (int i = 0; i < 3; ++i) { tc::mbar_inval(acc_ready + i); tc::mbar_init(acc_ready + i, 1); }
            tc::mbar_inval(h_ready); tc::mbar_init(h_ready, GW);
            for (int i = 0; i < NBUF; ++i) {
                tc::mbar_inval(h_free + i); tc::mbar_init(h_free + i, 1);
                tc::mbar_inval(y_ready + i); tc::mbar_init(y_ready + i, GW);
            }
            for (int i = 0; i < GI_STAGES; ++i) {
                tc::mbar_inval(gi_full + i); tc::mbar_init(gi_full + i, 1);
                tc::mbar_inval(gi_empty + i); tc::mbar_init(gi_empty + i, GW);
            }
            tc::mbar_fence_init();
        }
        __syncthreads();
    }
#ifdef HB_TIMELINE
    if (dbg && n_layers > 1 && cta_x == 0 && dir == 0 && tid == 0) dbg[4096 + (li * 64 + chunk) * 2] = (long long)globaltimer_ns();
    // (a clock read right after bar.sync records the ARRIVAL at the barrier - the block is deferred to the next access
    // of barrier-protected state - so the stamps read a shared word first)
#define HB_STAMP(k) do { if (dbg_on && n_layers > 1 && lane == 0) { \
        uint32_t probe; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(probe) : "r"(tc::smem_u32(tmem_slot)) : "memory"); \
        dbg[6144 + li * 8 + (k)] = clock64() + (probe & 0); } } while (0)
#else
#define HB_STAMP(k) do { } while (0)
#endif
    if (warp == 0) HB_STAMP(0);                              // phase start (barriers fresh, cross-CTA conditions met)
    const bool stamp = ra.phase_times != nullptr && cta_x == 0 && dir == 0 && tid == 0 && phase < 64;
    if (stamp) ra.phase_times[phase * 4 + 0] = (long long)globaltimer_ns();

    if (warp == GW + 1) {
        // ===================== gi loader: bulk copies, GI_STAGES steps ahead =====================
        // (the first GI_STAGES rows are requested before the phase's start barrier: the stages are free and the
        // prefetch then overlaps the W_hh upload of the gate warps)
        if (phase == 0) { tc::pdl_grid_dependency_wait(); tc::fence_proxy_async_all(); }   // gi' comes from an upstream kernel (generic-proxy stores)
        bool synced = false;
        if (phase > 0) {
            // later phases: join the phase-start barrier FIRST.  The MMA issuer waits there for every warp, and step 0's MMAs
            // need h_0 and the weights, not gi: with the first loads ahead of the barrier (as in phase 0, where they overlap
            // the W_hh upload) every phase started a global round trip late - the counter this warp polls below, then the
            // copies - ~2.7 us from phase start to the first MMA in the timeline build.
            __syncthreads();
            synced = true;
        }
        const bool discard_gi = n_layers > 1 && li == 1;
        // lane 3 g + gate fetches that gate block (4 KB) of group g's column
        const int lg = min(lane / 3, NG - 1), lgate = lane % 3;
        const float* src0 = gi + gi_block(b0 / WG + lg, gi_cols, gi_col0, dir * 3 + lgate);
        uint8_t* dst0 = gi_s + lg * GI_GRP_BYTES + lgate * GI_BLK_BYTES;
        for (int s = 0, t = t_first; s < W; ++s, t += dt) {
            const int stage = s % GI_STAGES;
            if (s == GI_STAGES && !synced) { __syncthreads(); synced = true; }
            const int col = L.flag_abs ? gi_col0 + t : t;
            if (L.tile_flags != nullptr && (s == 0 || (col & 7) == (dir ? 7 : 0)) && (col >> 3) >= L.flag_skip_tiles) {
                // chunk-loop kernel: the projection CTAs announce finished gi' tiles
                if (lane < NG && cta_x * NG + lane < ra.n_wg)
                    tc::spin_until_ge(L.tile_flags + (((size_t)cta_x * NG + lane) * L.flag_tiles + (col >> 3)) * 2 + dir,
                                      L.flag_need_base + L.flag_need_per_chunk * (unsigned long long)(chunk + 1));
                // the projection role wrote the tile with ordinary (generic-proxy) stores and fenced them before it raised
                // the counter; the bulk loads below belong to the async proxy: order them after what the acquire made visible
                __syncwarp();
                tc::fence_proxy_async_all();
                HB_CHAIN(ra.dbg, n_layers > 1 && cta_x == 0 && dir == 0 && chunk == 2 && li == 1 && s == 0 && lane == 0, 6);
            }
            if (s >= GI_STAGES) tc::mbar_wait(gi_empty + stage, (uint32_t)((s / GI_STAGES - 1) & 1));
            // chunk-loop kernel, decoder: the gi rows of GI_STAGES steps ago have been consumed and nobody reads them again
            // (the next chunk's projection overwrites them): L2 may drop the dirty lines instead of writing them to DRAM
            // (ncu at B=256: DRAM writes of the launch 3.10 -> 1.93 GB; the last GI_STAGES rows of a phase are left alone).
            // 32 lines of 128 bytes per 4 KB block: one per lane and block - with all of a block's lines on one lane the
            // loader fell behind and the launch was 9 % slower.
            if (discard_gi && s >= GI_STAGES) {
#pragma unroll
                for (int gb = 0; gb < 3 * NG; ++gb)
                    tc::discard_l2_line(reinterpret_cast<const char*>(gi + gi_block(b0 / WG + gb / 3, gi_cols, gi_col0 + (t - GI_STAGES * dt), dir * 3 + gb % 3)) + lane * 128);
            }
            if (lane == 0) tc::mbar_arrive_expect_tx(gi_full + stage, NG * GI_GRP_BYTES);
            __syncwarp();
            // (L2 evict-first: a gi block is dead once this copy has read it - the decoder's - or is read once more a chunk
            // later - the encoder's; with the evict-last stores of the projection role the decoder's gi is served from L2:
            // ncu at B=256, DRAM reads of the launch 2.49 -> 1.88 GB)
            if (lane < 3 * NG) tc::bulk_g2s_hint(dst0 + stage * GI_STAGE_BYTES, src0 + (int64_t)t * GI_BLK_FLOATS, GI_BLK_BYTES, gi_full + stage, tc::L2_EVICT_FIRST);
        }
        if (!synced) __syncthreads();
        if (phase + 1 < n_phases) {
            // this warp is GI_STAGES steps ahead of the gate warps: it uses the time to await the cross-CTA conditions of the
            // NEXT phase (one acquire load per lane, ~1 us of global round trip that used to sit between the phases)
            const int next_chunk = (phase + 1) / n_layers;
            const RecLayer& LN = ra.layer[(phase + 1) - next_chunk * n_layers];
            if (LN.heads_done != nullptr && next_chunk >= 2)  // the heads role has read the y image the next phase overwrites
                for (int g = lane; g < NG; g += 32)
                    if (cta_x * NG + g < ra.n_wg)
                        tc::spin_until_ge(LN.heads_done + (size_t)cta_x * NG + g, (unsigned long long)LN.heads_per_chunk * (next_chunk - 1));
            if (LN.consumed_flags != nullptr && next_chunk >= 1)   // every projection CTA has read that layer's previous image
                for (int f = lane; f < NG * ra.tiles_t * 2; f += 32)
                    if (cta_x * NG + f / (ra.tiles_t * 2) < ra.n_wg)
                        tc::spin_until_ge(LN.consumed_flags + (size_t)cta_x * NG * ra.tiles_t * 2 + f, (unsigned long long)LN.consumed_per_chunk * next_chunk);
            __syncwarp();
        }
    } else if (warp == GW + 2) {
        // ===================== y store: the h image of step s is the layer output at column t_s ====
        if (phase == 0) tc::pdl_grid_dependency_wait();      // yimg may still be read by an upstream kernel
        __syncthreads();
        unsigned long long* flag = L.progress ? L.progress + (size_t)cta_x * 2 + dir : nullptr;
        const int pub_off = dir ? (W & (PUBLISH_EVERY - 1)) : 0;
        for (int s = 0, t = t_first; s < W; ++s, t += dt) {
            const int buf = (s + 1) % NBUF;                  // the image written during step s
            tc::mbar_wait(y_ready + buf, (uint32_t)((s / NBUF) & 1));
            if (lane < 2 * NG) {
                const int g = lane >> 1, part = lane & 1;
                tc::bulk_s2g(yimg + yimg_block(b0 / WG + g, dir, part, t, W), h_img + (buf * 2 + part) * HB_BYTES + g * YBLK, YBLK);
                tc::bulk_commit();
                tc::bulk_wait_read0();
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(h_free + buf);
            if (flag != nullptr && s >= PUBLISH_LAG && s + PUBLISH_EVERY < W && ((s + 1 - PUBLISH_LAG - pub_off) & (PUBLISH_EVERY - 1)) == 0) {
                // at the consumers' tile boundaries: all stores but the newest PUBLISH_LAG have landed in global memory, publish
                // that many completed columns to the consumer CTAs.  (Waiting for the newest store, or
                // fencing every step, would put a global round trip on the step's critical path via h_free.)
                if (lane < 2 * NG) tc::bulk_wait_pending<PUBLISH_LAG>();   // writes performed; the release orders them before the counter
                __syncwarp();
                if (lane == 0) tc::st_release_gpu(flag, prog_base + (unsigned long long)(s + 1 - PUBLISH_LAG));
            }
        }
        HB_STAMP(3);                                         // last image handed to the copy engine
        if (lane < 2 * NG) tc::bulk_wait0();
        __syncwarp();
        if (flag != nullptr && lane == 0) tc::st_release_gpu(flag, prog_base + (unsigned long long)W);
        HB_CHAIN(ra.dbg, n_layers > 1 && cta_x == 0 && chunk == 2 && li == 0 && lane == 0, dir);
        HB_STAMP(4);                                         // all columns published
    } else if (warp == GW) {
        // ===================== MMA issuer =====================
        // An M=128, N=16, K=16 MMA occupies the tensor pipe ~9.6 cycles (tools/mma_rate.cu), so the 48 MMAs of a step are
        // ~460 cycles plus ~250 of pipeline fill and commit -> wake-up latency (-DHB_TIMELINE).  A second issuer warp,
        // an issue order rotating over the gate blocks, and starting the r block quarter by quarter as the gate warps
        // publish h (one h_ready barrier per TMEM lane quarter: +85 cycles per step, the extra waits cost more than the
        // overlap gives) were all slower; per-block commits let the gate warps overlap the r and z phases with the
        // remaining MMAs.
        __syncthreads();                                     // weights in TMEM, h_0 in smem
        tc::tc_fence_after();
        HB_STAMP(1);                                         // first step released
        const uint32_t idesc = tc::idesc_f16_f32(128, NACC);
        // stacked operand: NLIVE == 8 takes rows 8..15 from the lo image (group stride = HB_BYTES); with NLIVE == 16
        // the lo image directly follows the two hi groups, so the plain group stride covers all 32 rows
        const uint64_t himg_desc = tc::smem_desc_sw128(tc::smem_u32(h_img), (STACK && NLIVE == 8) ? HB_BYTES : YBLK);
        for (int s = 0; s < W; ++s) {
            if (s > 0) {
                // h_{s} is complete when every gate warp has arrived: a hardware named barrier (the gate warps arrive
                // without blocking, this warp syncs) releases this warp ~50 cycles sooner than an mbarrier wake-up
                tc::named_barrier_sync(REC_STEP_BARRIER, (GW + 1) * 32);
                tc::tc_fence_after();
            }
            const uint64_t hhi_desc = himg_desc + (uint64_t)(((s % NBUF) * 2 + 0) * HB_BYTES / 16);
            const uint64_t hlo_desc = himg_desc + (uint64_t)(((s % NBUF) * 2 + 1) * HB_BYTES / 16);
            if (tc::elect_one()) {
#pragma unroll
                for (int gb = 0; gb < 3; ++gb) {             // gate blocks r, z, n
                    if constexpr (STACK) {
#pragma unroll
                        for (int term = 0; term < 2; ++term) {   // W_hi, W_lo against [h_hi | h_lo]
                            const uint32_t a_col = tmem + REC_W_COL0 + (term * 3 + gb) * 64;
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks)
                                tc::mma_f16_ts(tmem + gb * NBLK, a_col + ks * 8, hhi_desc + tc::sw128_kstep(ks), idesc, (term | ks) != 0);
                        }
                    } else {
#pragma unroll
                        for (int term = 0; term < 3; ++term) {   // (W_hi,h_hi) (W_lo,h_hi) (W_hi,h_lo)
                            const uint32_t a_col = tmem + REC_W_COL0 + ((term == 1 ? 3 : 0) + gb) * 64;
                            const uint64_t bd = term == 2 ? hlo_desc : hhi_desc;
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks)
                                tc::mma_f16_ts(tmem + gb * NBLK, a_col + ks * 8, bd + tc::sw128_kstep(ks), idesc, (term | ks) != 0);
                        }
                    }
                    tc::mma_commit(acc_ready + gb);          // gates start on r while z, n still run
                }
            }
            __syncwarp();
        }
    } else {
        // ===================== gate warps =====================
        if (phase == 0) upload_whh(L.whh_tmem, dir, tmem, warp, lane, GW);   // W_hh of the first phase's layer -> TMEM (later phases: see below)
        const float* gc = L.gate_consts + (size_t)dir * 4 * H + j;
        const float inv_r = gc[0], inv_z = gc[H], inv_n = gc[2 * H], bhn = gc[3 * H];
        uint32_t h_off[NW];
        if (phase == 0) {
            tc::pdl_grid_dependency_wait();                  // h_in comes from an upstream kernel
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const int64_t b = b0 + win0 + i;
                h_own[i] = (ra.h_in != nullptr && b < B) ? __ldcg(ra.h_in + (b * 2 + dir) * H + j) * ACT_SCALE : 0.f;   // L2: may come from another SM
            }
        }
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            h_off[i] = (uint32_t)((win0 + i) / WG) * YBLK + tc::sw128_offset((win0 + i) % WG, j);
            __half hi, lo;
            tc::split_f16(h_own[i], hi, lo);
            *reinterpret_cast<__half*>(h_img + h_off[i]) = hi;                      // h_0 -> buffer 0
            *reinterpret_cast<__half*>(h_img + HB_BYTES + h_off[i]) = lo;
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        __syncthreads();                                     // pairs with the other roles' barrier
        if (stamp) ra.phase_times[phase * 4 + 1] = (long long)globaltimer_ns();

        // everything below that does not change from step to step stays in registers
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)win0;
        // stage: [group][gate][unit][window in group]; this thread's NW windows lie in one group
        const float* gs0 = reinterpret_cast<const float*>(gi_s) + (win0 / WG) * (3 * GI_BLK_FLOATS);
        auto load_acc = [](uint32_t addr, float* a) {
            tc::tmem_ld_n<NW>(addr, a);
            if constexpr (STACK) {
                float a_lo[NW];
                tc::tmem_ld_n<NW>(addr + NLIVE, a_lo);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < NW; ++i) a[i] += a_lo[i];
            } else {
                tc::tmem_ld_wait();
            }
        };
        int stage = 0;
        uint32_t gi_par = 0;
        for (int s = 0; s < W; ++s) {
            const uint32_t par = (uint32_t)(s & 1);
            const int nb = (s + 1) % NBUF;
#ifdef HB_TIMELINE_STEPS
            const int drole = warp == 0 ? 1 : (warp == GW - 1 ? 2 : 3);
#endif
            tc::mbar_wait(gi_full + stage, gi_par);
            HB_CHAIN(ra.dbg, n_layers > 1 && cta_x == 0 && dir == 0 && chunk == 2 && li == 1 && s == 0 && tid == 0, 7);
            // the buffer h_{s+1} goes to: its y store of NBUF steps ago has drained long since - waited for HERE, where
            // the warp would idle anyway, not between the z and n phases
            if (s >= NBUF) tc::mbar_wait(h_free + nb, (uint32_t)(((s - NBUF) / NBUF) & 1));
            HB_DBG(drole, s, 0);
            float gir[NW], giz[NW], gin[NW];
            {
                const float* gs = gs0 + stage * (GI_STAGE_BYTES / 4);
                gi_load<NW>(gs, j, win0 % WG, gir);
                gi_load<NW>(gs + GI_BLK_FLOATS, j, win0 % WG, giz);
                gi_load<NW>(gs + 2 * GI_BLK_FLOATS, j, win0 % WG, gin);
            }
            gi_loads_done<NW>(gir, giz, gin);
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(gi_empty + stage);
            if (++stage == GI_STAGES) { stage = 0; gi_par ^= 1u; }
            float a[NW];
            GateR gr[NW];
            GateZ gz[NW];
            tc::mbar_wait(acc_ready + 0, par);
            HB_DBG(drole, s, 1);
            tc::tc_fence_after();
            load_acc(taddr, a);
#pragma unroll
            for (int i = 0; i < NW; ++i) gr[i] = gate_r(a[i], inv_r, gir[i], inv_n, bhn, gin[i]);
            tc::mbar_wait(acc_ready + 1, par);
            HB_DBG(drole, s, 2);
            tc::tc_fence_after();
            load_acc(taddr + NBLK, a);
#pragma unroll
            for (int i = 0; i < NW; ++i) gz[i] = gate_z(a[i], inv_z, giz[i]);
            uint8_t* h_hi = h_img + (nb * 2) * HB_BYTES;     // image of h_{s+1}
            uint8_t* h_lo = h_hi + HB_BYTES;
            tc::mbar_wait(acc_ready + 2, par);
            HB_DBG(drole, s, 3);
            tc::tc_fence_after();
            load_acc(taddr + 2 * NBLK, a);
            HB_DBG(drole, s, 4);
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                // (2^x as a degree-7 polynomial on the FMA pipe instead of MUFU.EX2 was measured: +50 cycles per step)
                const float hn = gate_n(a[i], gr[i], gz[i], h_own[i]);
                h_own[i] = hn;
                __half hi, lo;
                tc::split_f16(hn, hi, lo);
                *reinterpret_cast<__half*>(h_hi + h_off[i]) = hi;
                *reinterpret_cast<__half*>(h_lo + h_off[i]) = lo;
            }
            HB_DBG(drole, s, 5);
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            __syncwarp();
            if (s + 1 < W) tc::named_barrier_arrive(REC_STEP_BARRIER, (GW + 1) * 32);   // (the last step has no MMA warp waiting)
            if (lane == 0) tc::mbar_arrive(y_ready + nb);
            HB_DBG(drole, s, 6);
#ifdef HB_TIMELINE_STEPS
            if (dbg_steps && lane == 0 && s < 128) dbg[8192 + warp * 128 + s] = clock64();     // every gate warp's arrival
#endif
        }
        if (warp == 0) HB_STAMP(2);                          // last step done
        HB_CHAIN(ra.dbg, n_layers > 1 && cta_x == 0 && chunk == 2 && li == 0 && tid == 0, 8 + dir);
        if (stamp) ra.phase_times[phase * 4 + 2] = (long long)globaltimer_ns();
        // W_hh of the NEXT phase's layer -> TMEM right away, while the y store drains and publishes the last columns: all
        // MMAs of this phase have completed (every gate warp has waited for its last accumulator) and nothing else reads
        // the weight columns.  (Uploading at the start of the next phase instead kept every role waiting ~1.5 us longer.)
        if (n_layers > 1 && phase + 1 < n_phases) upload_whh(ra.layer[(li + 1) % n_layers].whh_tmem, dir, tmem, warp, lane, GW);
        if (phase == n_phases - 1 && ra.h_out != nullptr) {
#pragma unroll
            for (int i = 0; i < NW; ++i)
                if (b0 + win0 + i < B) ra.h_out[((b0 + win0 + i) * 2 + dir) * H + j] = h_own[i] * ACT_SCALE_INV;
        }
    }
    tc::tc_fence_before();
    __syncthreads();                                         // phase end: every role is done with the barriers
    if (stamp) ra.phase_times[phase * 4 + 3] = (long long)globaltimer_ns();
    if (warp == 0) HB_STAMP(5);
#undef HB_STAMP
#ifdef HB_TIMELINE
    if (dbg && n_layers > 1 && cta_x == 0 && dir == 0 && tid == 0) dbg[4096 + (li * 64 + chunk) * 2 + 1] = (long long)globaltimer_ns();
#endif
    }   // phase loop
#undef HB_DBG
    if (warp == GW) tc::tmem_dealloc(tmem, 512);
}

template <int N, int NLIVE, int MODE>
__global__ void __launch_bounds__(REC_TC_THREADS, 1)
tc_recurrence_kernel(const RecArgs ra)
{
    extern __shared__ __align__(1024) uint8_t smem_rec[];
    recurrence_role<N, NLIVE, MODE>(ra, tc::align_smem_1024(smem_rec), (int)blockIdx.x, (int)blockIdx.y);
}

// ---------------------------------------------------------------------------------------------
// Recurrence with TWO window tiles per CTA that share W_hh in TMEM and take turns on the tensor pipe (throughput mode).
// With one tile per CTA the tensor pipe idles during the gate math of a step (~45 % of it) and a second CTA cannot share
// the SM because each needs 384 of the 512 TMEM columns for W_hh.  Here tile A's MMAs of step s+1 are issued while tile
// B's gate warps still work on step s: the pipe stays busy and a step pair costs its 2 x 48 (stacked, 8-window tiles) or
// 2 x 72 (3-term, 16-window tiles) MMAs.
// Warps 0-7: gate warps of tile 0, 8-15: of tile 1 (warp w of a tile: TMEM lane quarter w % 4, window half w / 4),
// 16: MMA issuer, 17: gi' loader, 18: y store.  NLIVE windows per tile; MODE 1 = stacked [h_hi | h_lo] (NLIVE = 8),
// MODE 0 = 3-term (NLIVE = 16).
// Like recurrence_role it runs one layer of one chunk (per-chunk launches, n_layers == 1) or, as a role of the chunk-loop
// kernel, the alternating encoder / decoder phases of the whole chunk loop (n_layers == 2) with the state in registers
// and the cross-CTA counters of RecLayer; a CTA then covers 2 NLIVE / 8 consecutive window groups.
// ---------------------------------------------------------------------------------------------
template <int NLIVE> __host__ __device__ constexpr int rec2_stages() { return NLIVE <= 8 ? 6 : 3; }    // gi' stages per tile
template <int NLIVE> __host__ __device__ constexpr int rec2_hbufs() { return NLIVE <= 8 ? 8 : 2; }     // h image buffers per tile
template <int NLIVE> constexpr size_t recurrence2_smem() {
    return (size_t)2 * (2 * rec2_hbufs<NLIVE>() * (NLIVE <= 8 ? 1 : 2) * YBLK + rec2_stages<NLIVE>() * NLIVE * GI_ROW_BYTES) + 512 +
           (size_t)REC_GATE_WARPS * 32 * (NLIVE / 2) * 4 + 1024;     // tiles, barriers, state parked between phases, alignment
}
constexpr int REC2_PHASE_BARRIER = REC_STEP_BARRIER + 2;     // named barrier: all gate warps of both tiles, once per phase

template <int NLIVE, int MODE>
__device__ __forceinline__ void recurrence2_role(const RecArgs& ra, uint8_t* smem, const int cta_x, const int dir)
{
    static_assert((NLIVE == 8 && MODE == 1) || (NLIVE == 16 && MODE == 0), "8-window stacked tiles or 16-window 3-term tiles");
    constexpr bool STACK = MODE == 1;
    constexpr int NACC = 16;                                 // N of every MMA == accumulator columns per (tile, gate block)
    constexpr int GW = REC_GATE_WARPS / 2;                   // gate warps per tile
    constexpr int NW = NLIVE / 2;                            // windows per gate thread
    constexpr int NG = NLIVE / WG;                           // window groups per tile
    constexpr int NGC = 2 * NG;                              // window groups per CTA
    constexpr uint32_t HB_BYTES = STACK ? YBLK : 2 * YBLK;   // one h operand image (hi or lo): one (stacked) or two 8-row groups
    constexpr int NBUF = rec2_hbufs<NLIVE>(), ST = rec2_stages<NLIVE>();
    constexpr uint32_t GI_STAGE_BYTES = NLIVE * GI_ROW_BYTES;
    constexpr uint32_t TILE_BYTES = 2 * NBUF * HB_BYTES + ST * GI_STAGE_BYTES;
    constexpr int NBAR = 3 + NBUF + NBUF + ST + ST;          // barriers per tile
    static_assert(2 * NBAR * 8 + 8 <= 512, "barrier area");
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TILE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NBAR);
    // the state h is parked here between two phases: kept in registers across the W_hh upload (64 staging registers) the
    // compiler moved it, and the h image offsets, to local memory for the whole step loop
    float* h_park = reinterpret_cast<float*>(smem + 2 * TILE_BYTES + 512) + threadIdx.x;
    auto h_img_of = [&](int tile) { return smem + tile * TILE_BYTES; };
    auto gi_of = [&](int tile) { return smem + tile * TILE_BYTES + 2 * NBUF * HB_BYTES; };
    auto acc_ready = [&](int tile) { return bars + tile * NBAR; };                  // [3]
    auto h_free = [&](int tile) { return bars + tile * NBAR + 3; };                 // [NBUF]
    auto y_ready = [&](int tile) { return bars + tile * NBAR + 3 + NBUF; };         // [NBUF]
    auto gi_full = [&](int tile) { return bars + tile * NBAR + 3 + 2 * NBUF; };     // [ST]
    auto gi_empty = [&](int tile) { return bars + tile * NBAR + 3 + 2 * NBUF + ST; };

    const int64_t B = ra.B;
    const int W = ra.W;
    const int n_layers = ra.n_layers;
    const int n_phases = (ra.n_chunks > 0 ? ra.n_chunks : 1) * n_layers;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int64_t b0 = (int64_t)cta_x * 2 * NLIVE;           // tile t covers windows [b0 + t NLIVE, b0 + (t + 1) NLIVE)
    const int t_first = dir ? W - 1 : 0, dt = dir ? -1 : 1;

    tc::pdl_launch_dependents();
    // (stacked operand: rows 8-15 of the B operand are group 0 of the lo image; the second 8-row group of an image is never read)
    auto init_barriers = [&](bool again) {
        for (int tile = 0; tile < 2; ++tile) {
            auto one = [&](uint64_t* b, int count) { if (again) tc::mbar_inval(b); tc::mbar_init(b, count); };
            for (int i = 0; i < 3; ++i) one(acc_ready(tile) + i, 1);
            for (int i = 0; i < NBUF; ++i) { one(h_free(tile) + i, 1); one(y_ready(tile) + i, GW); }
            for (int i = 0; i < ST; ++i) { one(gi_full(tile) + i, 1); one(gi_empty(tile) + i, GW); }
        }
        tc::mbar_fence_init();
    };
    if (tid == 0) init_barriers(false);
    __syncwarp();
    if (warp == REC_GATE_WARPS) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // gate-thread state that lives across phases
    const int tile_w = (warp / GW) & 1, wi = warp % GW;
    const int q = wi & 3, j = q * 32 + lane, win0 = (wi >> 2) * NW;
    const int64_t bt = b0 + tile_w * NLIVE + win0;           // first window of this gate thread

    for (int phase = 0; phase < n_phases; ++phase) {
    const int chunk = phase / n_layers, li = phase - chunk * n_layers;
    const RecLayer& L = ra.layer[li];
    const int gi_col0 = L.gi_col0 + chunk * L.gi_col_step;
    uint8_t* __restrict__ yimg = L.yimg[chunk & 1];
    const unsigned long long prog_base = ra.epoch + (unsigned long long)chunk * W;
    if (phase > 0) {
        if (tid == 0) init_barriers(true);                   // (one thread: see recurrence_role)
        __syncthreads();
    }
    const bool stamp = ra.phase_times != nullptr && cta_x == 0 && dir == 0 && tid == 0 && phase < 64;
    if (stamp) ra.phase_times[phase * 4 + 0] = (long long)globaltimer_ns();
#ifdef HB_TIMELINE_STEPS
    // per-step stamps of the MMA issuer ([step][tile][released, issued]) and of the first gate warp of each tile
    const bool dbg_steps = ra.dbg != nullptr && cta_x == 0 && dir == 0 && li == ra.dbg_layer && (n_layers == 1 || chunk == 2);
#define HB_DBG2(idx) do { if (dbg_steps && lane == 0 && s < 128) ra.dbg[idx] = clock64(); } while (0)
#else
#define HB_DBG2(idx) do { } while (0)
#endif

    if (warp == REC_GATE_WARPS + 1) {
        // ===================== gi' loader: both tiles, ST steps ahead =====================
        if (phase == 0) { tc::pdl_grid_dependency_wait(); tc::fence_proxy_async_all(); }   // gi' was written with generic-proxy stores, the bulk loads are async-proxy reads
        bool synced = false;
        if (phase > 0) {                                     // (see recurrence_role: the phase-start barrier first)
            __syncthreads();
            synced = true;
        }
        // lanes [8 tile + 3 g, + 3): the three gate blocks (4 KB each) of group g of a tile
        const int tile_l = lane >> 3, g_l = min((lane & 7) / 3, NG - 1), gate_l = (lane & 7) % 3;
        const bool loads = tile_l < 2 && (lane & 7) < 3 * NG;
        const float* src0 = L.gi + gi_block(b0 / WG + min(tile_l, 1) * NG + g_l, L.gi_cols, gi_col0, dir * 3 + gate_l);
        uint8_t* dst0 = gi_of(min(tile_l, 1)) + g_l * GI_GRP_BYTES + gate_l * GI_BLK_BYTES;
        for (int s = 0, t = t_first; s < W; ++s, t += dt) {
            const int stage = s % ST;
            if (s == ST && !synced) { __syncthreads(); synced = true; }
            const int col = L.flag_abs ? gi_col0 + t : t;
            if (L.tile_flags != nullptr && (s == 0 || (col & 7) == (dir ? 7 : 0)) && (col >> 3) >= L.flag_skip_tiles) {
                // chunk-loop kernel: the projection CTAs announce finished gi' tiles (see recurrence_role)
                if (lane < NGC && cta_x * NGC + lane < ra.n_wg)
                    tc::spin_until_ge(L.tile_flags + (((size_t)cta_x * NGC + lane) * L.flag_tiles + (col >> 3)) * 2 + dir,
                                      L.flag_need_base + L.flag_need_per_chunk * (unsigned long long)(chunk + 1));
                __syncwarp();
                tc::fence_proxy_async_all();
            }
            for (int tile = 0; tile < 2; ++tile) {
                if (s >= ST) tc::mbar_wait(gi_empty(tile) + stage, (uint32_t)((s / ST - 1) & 1));
                if (lane == 0) tc::mbar_arrive_expect_tx(gi_full(tile) + stage, NG * GI_GRP_BYTES);
            }
            if (n_layers > 1 && li == 1 && s >= ST) {        // consumed decoder gi rows: see recurrence_role
#pragma unroll
                for (int gb = 0; gb < 3 * NGC; ++gb)
                    tc::discard_l2_line(reinterpret_cast<const char*>(L.gi + gi_block(b0 / WG + gb / 3, L.gi_cols, gi_col0 + (t - ST * dt), dir * 3 + gb % 3)) + lane * 128);
            }
            __syncwarp();
            if (loads) tc::bulk_g2s_hint(dst0 + stage * GI_STAGE_BYTES, src0 + (int64_t)t * GI_BLK_FLOATS, GI_BLK_BYTES, gi_full(tile_l) + stage, tc::L2_EVICT_FIRST);
        }
        if (!synced) __syncthreads();
        if (phase + 1 < n_phases) {
            // cross-CTA conditions of the NEXT phase, awaited while this phase's last steps run (see recurrence_role)
            const int next_chunk = (phase + 1) / n_layers;
            const RecLayer& LN = ra.layer[(phase + 1) - next_chunk * n_layers];
            if (LN.heads_done != nullptr && next_chunk >= 2)
                for (int g = lane; g < NGC; g += 32)
                    if (cta_x * NGC + g < ra.n_wg)
                        tc::spin_until_ge(LN.heads_done + (size_t)cta_x * NGC + g, (unsigned long long)LN.heads_per_chunk * (next_chunk - 1));
            if (LN.consumed_flags != nullptr && next_chunk >= 1)
                for (int f = lane; f < NGC * ra.tiles_t * 2; f += 32)
                    if (cta_x * NGC + f / (ra.tiles_t * 2) < ra.n_wg)
                        tc::spin_until_ge(LN.consumed_flags + (size_t)cta_x * NGC * ra.tiles_t * 2 + f, (unsigned long long)LN.consumed_per_chunk * next_chunk);
            __syncwarp();
        }
    } else if (warp == REC_GATE_WARPS + 2) {
        // ===================== y store: lanes [2 NG tile, + 2 NG) store the (group, part) images of a tile =====================
        if (phase == 0) tc::pdl_grid_dependency_wait();
        __syncthreads();
        unsigned long long* flag = L.progress ? L.progress + (size_t)cta_x * 2 + dir : nullptr;
        const int pub_off = dir ? (W & (PUBLISH_EVERY - 1)) : 0;
        const int tile_y = lane / (2 * NG), g_y = (lane % (2 * NG)) >> 1, part_y = lane & 1;
        const bool stores = lane < 4 * NG;
#ifdef HB_TIMELINE_STEPS
        long long ty[4] = {0, 0, 0, 0};
#define HB_YT(k, ...) do { const long long t_ = clock64(); __VA_ARGS__; if (dbg_steps && s >= 20 && s < 90) ty[k] += clock64() - t_; } while (0)
#else
#define HB_YT(k, ...) do { __VA_ARGS__; } while (0)
#endif
        for (int s = 0, t = t_first; s < W; ++s, t += dt) {
            const int buf = (s + 1) % NBUF;
            for (int tile = 0; tile < 2; ++tile) {
                HB_YT(0, tc::mbar_wait(y_ready(tile) + buf, (uint32_t)((s / NBUF) & 1)));
                HB_YT(1,
                if (stores && tile_y == tile) {
                    tc::bulk_s2g(yimg + yimg_block(b0 / WG + tile * NG + g_y, dir, part_y, t, W), h_img_of(tile) + (buf * 2 + part_y) * HB_BYTES + g_y * YBLK, YBLK);
                    tc::bulk_commit();
                }
                __syncwarp());
            }
            // the buffers of the PREVIOUS step are handed back now: their stores have had a whole step pair to read shared
            // memory (waiting for each store right after issuing it - twice per step pair - made this warp pace the kernel)
            HB_YT(2, if (stores) tc::bulk_wait_read_pending<1>(); __syncwarp());
            if (s > 0 && lane < 2) tc::mbar_arrive(h_free(lane) + s % NBUF);
            if (flag != nullptr && s >= PUBLISH_LAG && s + PUBLISH_EVERY < W && ((s + 1 - PUBLISH_LAG - pub_off) & (PUBLISH_EVERY - 1)) == 0) {
                HB_YT(3, if (stores) tc::bulk_wait_pending<PUBLISH_LAG>(); __syncwarp());   // (one bulk group per lane and step)
#ifndef HB_PUBLISH_LANE
#define HB_PUBLISH_LANE 0
#endif
#ifdef HB_TIMELINE_STEPS
                { const long long t_ = clock64();
                  if (lane == HB_PUBLISH_LANE) tc::st_release_gpu(flag, prog_base + (unsigned long long)(s + 1 - PUBLISH_LAG));
                  __syncwarp();
                  if (dbg_steps && s >= 20 && s < 90 && lane == 0) ra.dbg[3004] += clock64() - t_; }
#else
                if (lane == HB_PUBLISH_LANE) tc::st_release_gpu(flag, prog_base + (unsigned long long)(s + 1 - PUBLISH_LAG));
#endif
            }
        }
#ifdef HB_TIMELINE_STEPS
        if (dbg_steps && lane == 0) for (int k = 0; k < 4; ++k) ra.dbg[3000 + k] = ty[k];
#endif
#undef HB_YT
        if (stores) tc::bulk_wait0();
        __syncwarp();
        if (lane < 2) tc::mbar_arrive(h_free(lane) + W % NBUF);
        if (flag != nullptr && lane == 0) tc::st_release_gpu(flag, prog_base + (unsigned long long)W);
    } else if (warp == REC_GATE_WARPS) {
        // ===================== MMA issuer: tile 0, tile 1, tile 0, ... =====================
        __syncthreads();                                     // weights in TMEM, h_0 of both tiles in smem
        tc::tc_fence_after();
        const uint32_t idesc = tc::idesc_f16_f32(128, NACC);
        const uint64_t base0 = tc::smem_desc_sw128(tc::smem_u32(h_img_of(0)), (STACK && NLIVE == 8) ? HB_BYTES : YBLK);
        for (int s = 0; s < W; ++s) {
#pragma unroll
            for (int tile = 0; tile < 2; ++tile) {
                if (s > 0) {
                    tc::named_barrier_sync(REC_STEP_BARRIER + tile, (GW + 1) * 32);   // this tile's gate warps have published h_s
                    tc::tc_fence_after();
                }
                HB_DBG2((s * 2 + tile) * 2);
                const uint64_t base = base0 + (uint64_t)(tile * (TILE_BYTES / 16));
                const uint64_t hhi_desc = base + (uint64_t)(((s % NBUF) * 2 + 0) * HB_BYTES / 16);
                const uint64_t hlo_desc = base + (uint64_t)(((s % NBUF) * 2 + 1) * HB_BYTES / 16);
                const uint32_t acc0 = tmem + tile * 3 * NACC;
                if (tc::elect_one()) {
#pragma unroll
                    for (int gb = 0; gb < 3; ++gb) {
                        if constexpr (STACK) {
#pragma unroll
                            for (int term = 0; term < 2; ++term)
#pragma unroll
                                for (int ks = 0; ks < 8; ++ks)
                                    tc::mma_f16_ts(acc0 + gb * NACC, tmem + REC_W_COL0 + (term * 3 + gb) * 64 + ks * 8, hhi_desc + tc::sw128_kstep(ks), idesc, (term | ks) != 0);
                        } else {
#pragma unroll
                            for (int term = 0; term < 3; ++term) {   // (W_hi,h_hi) (W_lo,h_hi) (W_hi,h_lo)
                                const uint64_t bd = term == 2 ? hlo_desc : hhi_desc;
#pragma unroll
                                for (int ks = 0; ks < 8; ++ks)
                                    tc::mma_f16_ts(acc0 + gb * NACC, tmem + REC_W_COL0 + ((term == 1 ? 3 : 0) + gb) * 64 + ks * 8, bd + tc::sw128_kstep(ks), idesc, (term | ks) != 0);
                            }
                        }
                        tc::mma_commit(acc_ready(tile) + gb);
                    }
                }
                __syncwarp();
                HB_DBG2((s * 2 + tile) * 2 + 1);
            }
        }
    } else {
        // ===================== gate warps =====================
        if (phase == 0) upload_whh<16>(L.whh_tmem, dir, tmem, warp, lane, REC_GATE_WARPS);   // W_hh -> TMEM, split over all gate warps
        uint8_t* h_img = h_img_of(tile_w);
        const float* gc = L.gate_consts + (size_t)dir * 4 * H + j;
        const float inv_r = gc[0], inv_z = gc[H], inv_n = gc[2 * H], bhn = gc[3 * H];
        uint32_t h_off[NW];
        float h_own[NW];                                     // h * 2^10
        if (phase == 0) {
            tc::pdl_grid_dependency_wait();
#pragma unroll
            for (int i = 0; i < NW; ++i)
                h_own[i] = (ra.h_in != nullptr && bt + i < B) ? __ldcg(ra.h_in + ((bt + i) * 2 + dir) * H + j) * ACT_SCALE : 0.f;
        } else {
#pragma unroll
            for (int i = 0; i < NW; ++i) h_own[i] = h_park[i * (REC_GATE_WARPS * 32)];
        }
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            h_off[i] = (uint32_t)((win0 + i) / WG) * YBLK + tc::sw128_offset((win0 + i) % WG, j);
            __half hi, lo;
            tc::split_f16(h_own[i], hi, lo);
            *reinterpret_cast<__half*>(h_img + h_off[i]) = hi;
            *reinterpret_cast<__half*>(h_img + HB_BYTES + h_off[i]) = lo;
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        __syncthreads();
        if (stamp) ra.phase_times[phase * 4 + 1] = (long long)globaltimer_ns();

        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile_w * 3 * NACC + win0);
        // stage: [group][gate][unit][window in group]; this thread's NW windows lie in one group
        const float* gs0 = reinterpret_cast<const float*>(gi_of(tile_w)) + (win0 / WG) * (3 * GI_BLK_FLOATS);
        uint64_t* const my_acc = acc_ready(tile_w);
        uint64_t* const my_gi_full = gi_full(tile_w);
        uint64_t* const my_gi_empty = gi_empty(tile_w);
        uint64_t* const my_h_free = h_free(tile_w);
        uint64_t* const my_y_ready = y_ready(tile_w);
        auto load_acc = [](uint32_t addr, float* a) {
            tc::tmem_ld_n<NW>(addr, a);
            if constexpr (STACK) {
                float a_lo[NW];
                tc::tmem_ld_n<NW>(addr + NLIVE, a_lo);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < NW; ++i) a[i] += a_lo[i];
            } else {
                tc::tmem_ld_wait();
            }
        };
        int stage = 0;
        uint32_t gi_par = 0;
        for (int s = 0; s < W; ++s) {
            const uint32_t par = (uint32_t)(s & 1);
            const int nb = (s + 1) % NBUF;
            float a[NW], gir[NW], giz[NW], gin[NW];
            GateR gr[NW];
            GateZ gz[NW];
#ifdef HB_TIMELINE_STEPS
#define HB_DBGG(k) do { if (wi == 0) HB_DBG2(1024 + (tile_w * 128 + s) * 8 + (k)); } while (0)
#else
#define HB_DBGG(k) do { } while (0)
#endif
            tc::mbar_wait(my_gi_full + stage, gi_par);
            HB_DBGG(7);
            if (s >= NBUF) tc::mbar_wait(my_h_free + nb, (uint32_t)(((s - NBUF) / NBUF) & 1));
            HB_DBGG(0);
            {
                const float* gs = gs0 + stage * (GI_STAGE_BYTES / 4);
                gi_load<NW>(gs, j, win0 % WG, gir);
                gi_load<NW>(gs + GI_BLK_FLOATS, j, win0 % WG, giz);
                gi_load<NW>(gs + 2 * GI_BLK_FLOATS, j, win0 % WG, gin);
            }
            gi_loads_done<NW>(gir, giz, gin);
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(my_gi_empty + stage);
            if (++stage == ST) { stage = 0; gi_par ^= 1u; }
            tc::mbar_wait(my_acc + 0, par);
            HB_DBGG(1);
            tc::tc_fence_after();
            load_acc(taddr, a);
#pragma unroll
            for (int i = 0; i < NW; ++i) gr[i] = gate_r(a[i], inv_r, gir[i], inv_n, bhn, gin[i]);
            tc::mbar_wait(my_acc + 1, par);
            HB_DBGG(2);
            tc::tc_fence_after();
            load_acc(taddr + NACC, a);
#pragma unroll
            for (int i = 0; i < NW; ++i) gz[i] = gate_z(a[i], inv_z, giz[i]);
            uint8_t* h_hi = h_img + (nb * 2) * HB_BYTES;
            uint8_t* h_lo = h_hi + HB_BYTES;
            tc::mbar_wait(my_acc + 2, par);
            HB_DBGG(3);
            tc::tc_fence_after();
            load_acc(taddr + 2 * NACC, a);
            HB_DBGG(4);
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const float hn = gate_n(a[i], gr[i], gz[i], h_own[i]);
                h_own[i] = hn;
                __half hi, lo;
                tc::split_f16(hn, hi, lo);
                *reinterpret_cast<__half*>(h_hi + h_off[i]) = hi;
                *reinterpret_cast<__half*>(h_lo + h_off[i]) = lo;
            }
            HB_DBGG(5);
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            __syncwarp();
            if (s + 1 < W) tc::named_barrier_arrive(REC_STEP_BARRIER + tile_w, (GW + 1) * 32);
            if (lane == 0) tc::mbar_arrive(my_y_ready + nb);
            HB_DBGG(6);
        }
#undef HB_DBGG
        if (stamp) ra.phase_times[phase * 4 + 2] = (long long)globaltimer_ns();
        if (phase + 1 < n_phases) {
#pragma unroll
            for (int i = 0; i < NW; ++i) h_park[i * (REC_GATE_WARPS * 32)] = h_own[i];
        }
        if (n_layers > 1 && phase + 1 < n_phases) {
            // W_hh of the next phase's layer -> TMEM while the y store drains; the OTHER tile's last MMAs may still be reading
            // the weight columns until its gate warps, too, have seen their last accumulator
            tc::named_barrier_sync(REC2_PHASE_BARRIER, REC_GATE_WARPS * 32);
            upload_whh<16>(ra.layer[(li + 1) % n_layers].whh_tmem, dir, tmem, warp, lane, REC_GATE_WARPS);
        }
        if (phase == n_phases - 1 && ra.h_out != nullptr) {
#pragma unroll
            for (int i = 0; i < NW; ++i)
                if (bt + i < B) ra.h_out[((bt + i) * 2 + dir) * H + j] = h_own[i] * ACT_SCALE_INV;
        }
    }
    tc::tc_fence_before();
    __syncthreads();                                         // phase end: every role is done with the barriers
    if (stamp) ra.phase_times[phase * 4 + 3] = (long long)globaltimer_ns();
#undef HB_DBG2
    }   // phase loop
    if (warp == REC_GATE_WARPS) tc::tmem_dealloc(tmem, 512);
}

template <int NLIVE, int MODE>
__global__ void __launch_bounds__(REC_TC_THREADS, 1)
tc_recurrence2_kernel(const RecArgs ra)
{
    extern __shared__ __align__(1024) uint8_t smem_rec2[];
    recurrence2_role<NLIVE, MODE>(ra, tc::align_smem_1024(smem_rec2), (int)blockIdx.x, (int)blockIdx.y);
}

// ---------------------------------------------------------------------------------------------
// Heads + softmax + accumulate (predict_gpu.py:137-149) on tensor cores.
// Roles are swapped here: A = activations (M = 128 positions = 8 windows x 16 columns, straight
// from yimg), B = the 16 head rows (5 base + 11 rle), so every thread ends up with all 16 logits
// of ONE position in registers and the two softmaxes need no cross-thread traffic.
// ---------------------------------------------------------------------------------------------
constexpr int HEADS_THREADS = 192;
constexpr int HEADS_EPILOGUE_BARRIER = 6;        // named barrier of the four epilogue warps (0 = __syncthreads, 1 / 2 roles, 3-5 recurrence)
constexpr int HEADS_WIMG = NCLS * 2 * H * 2;     // one [16 x 256] fp16 image (dense core matrices): 8192 B

struct HeadsArgs {
    const uint8_t* yimg; const uint8_t* yimg_odd;      // decoder output image (odd chunks may use a second buffer)
    int64_t n_wg, B; int W, T, col0, col_step;         // chunk k accumulates into image columns col0 + k * col_step + t
    const __half* w_img;                               // [hi, lo][16 x 256] core-matrix image, LBO 128 / SBO 4096
    const float* b_head; float inv_scale;
    float* p_base; float* p_rle;                       // accumulated softmax sums in the reference's layouts [B, T, 5] / [B, T, 11] ...
    float* p16;                                        // ... or (when the caller does not ask for them) one [B, T, 16] array: 64 B per position
    int chunk0;                                        // index of the launch's first chunk in the reference loop (first-touch rule of p16)
    // p16 mode: the chunk that covers a column LAST turns the finished sums into the two labels itself (first-index argmax,
    // predict_gpu.py:155-156) and never writes them to p16; total_chunks = chunks of the whole reference loop
    int total_chunks; uint8_t* base_label; uint8_t* rle_label;
    // persistent mode
    int n_chunks; const unsigned long long* progress; int rec_n; const int* tile_order;
    unsigned long long* heads_done;                    // [group] += 4 per finished tile
    long long* dbg;                                    // HB_DEBUG_TIMELINE: worker 0 records when it finished each chunk
};

// job q of a worker -> (chunk, window group, column tile).  Single-chunk launches spread tiles over all workers;
// the persistent kernel pins window groups to workers so that the accumulation of successive chunks into
// the same P rows stays ordered.
__device__ __forceinline__ bool heads_job(const HeadsArgs& a, int worker, int n_workers, int tiles_t, int64_t q,
                                          int& chunk, int64_t& wg, int& t0, bool* last_of_wg = nullptr) {
    if (a.n_chunks <= 0) {
        const int64_t tile = worker + q * n_workers;
        if (tile >= a.n_wg * tiles_t) return false;
        chunk = 0; wg = tile / tiles_t; t0 = (int)(tile % tiles_t) * 16;
        if (last_of_wg) *last_of_wg = false;
        return true;
    }
    const int64_t my_wgs = (a.n_wg - worker + n_workers - 1) / n_workers;
    const int64_t per_chunk = my_wgs * tiles_t;
    if (per_chunk <= 0 || q >= per_chunk * a.n_chunks) return false;
    chunk = (int)(q / per_chunk);
    const int64_t r = q % per_chunk;
    wg = worker + (r / tiles_t) * n_workers;
    const int pos = (int)(r % tiles_t);
    t0 = (a.tile_order ? a.tile_order[pos] : pos) * 16;
    if (last_of_wg) *last_of_wg = pos == tiles_t - 1;         // a worker takes the tiles of a window group back to back
    return true;
}

__device__ __forceinline__ void heads_role(const HeadsArgs& a, uint8_t* smem, const int worker, const int n_workers)
{
    const int W = a.W, T = a.T;
    const int64_t B = a.B;
    constexpr uint32_t SLICE_BYTES = 16 * YBLK;                 // 16 columns of one (window group, direction, part)
    constexpr uint32_t PART_BYTES = 2 * SLICE_BYTES;
    uint8_t* a_img = smem;                                      // [hi, lo][direction][16 row groups][YBLK]
    uint8_t* w_s = smem + 2 * PART_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(w_s + 2 * HEADS_WIMG);   // [2]: hi part, lo part of the activation tile - the
    uint64_t* a_empty = a_full + 2;                            // [2]  loader refills the hi part while the lo part's MMAs still run
    uint64_t* acc_full = a_empty + 2;                          // [2]: two accumulators of 16 columns, so the softmax / P update of
    uint64_t* acc_empty = acc_full + 2;                        // [2]  one tile overlaps the load and the MMAs of the next
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    tc::pdl_launch_dependents();
    for (int i = tid; i < 2 * HEADS_WIMG / 16; i += HEADS_THREADS)
        reinterpret_cast<int4*>(w_s)[i] = reinterpret_cast<const int4*>(a.w_img)[i];
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { tc::mbar_init(a_full + i, 1); tc::mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, 4); }
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == 4) tc::tmem_alloc(tmem_slot, 32);
    tc::pdl_grid_dependency_wait();
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    tc::named_barrier_sync(2, HEADS_THREADS);
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int tiles_t = (W + 15) >> 4;
    int chunk; int64_t wg; int t0;
    // -DHB_TIMELINE: worker 0 adds up the cycles each role spends at its wait points (slots 7300 + 8 role + k)
#ifdef HB_TIMELINE
    const bool acct = a.dbg != nullptr && worker == 0 && lane == 0;
    long long t_wait[4] = {0, 0, 0, 0};
    const long long t_role0 = acct ? clock64() : 0;
#define HB_TIMED(k, ...) do { const long long t_ = acct ? clock64() : 0; __VA_ARGS__; if (acct) t_wait[k] += clock64() - t_; } while (0)
#define HB_ROLE_REPORT(role) do { if (acct) { for (int k_ = 0; k_ < 4; ++k_) a.dbg[7300 + 8 * (role) + k_] = t_wait[k_]; \
                                              a.dbg[7300 + 8 * (role) + 4] = clock64() - t_role0; } } while (0)
#else
#define HB_TIMED(k, ...) do { __VA_ARGS__; } while (0)
#define HB_ROLE_REPORT(role) do { } while (0)
#endif

    if (warp == 5) {
        const uint8_t* prev_img = nullptr;
        int64_t prev_wg = 0;
        int prev_t0 = 0, prev_valid = 0;
        for (int64_t it = 0; heads_job(a, worker, n_workers, tiles_t, it, chunk, wg, t0); ++it) {
            const int valid = min(16, W - t0);
            if (a.progress != nullptr) {                     // both decoder directions must have stored these columns
                if (lane < 2)
                    HB_TIMED(1, tc::spin_until_ge(a.progress + ((wg * WG) / a.rec_n) * 2 + lane,
                                                  (unsigned long long)chunk * W + (unsigned long long)(lane == 0 ? t0 + valid : W - t0)));

                __syncwarp();
            }
            const uint8_t* img = ((chunk & 1) && a.yimg_odd) ? a.yimg_odd : a.yimg;
            // four copies of four columns per (part, direction); columns are contiguous.  The hi part (64 KB) first: its MMAs
            // start while the lo part is still on its way
            const int q4 = lane & 3, d = (lane >> 2) & 1;
            const int cols = min(4, valid - 4 * q4);
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                if (it > 0) HB_TIMED(0, tc::mbar_wait(a_empty + part, (uint32_t)((it - 1) & 1)));
                if (lane == 0) tc::mbar_arrive_expect_tx(a_full + part, (uint32_t)(valid * YROW));
                __syncwarp();
                if ((lane >> 3) == part && cols > 0)
                    tc::bulk_g2s(a_img + part * PART_BYTES + d * SLICE_BYTES + q4 * 4 * YBLK, img + yimg_block(wg, d, part, t0 + 4 * q4, W), (uint32_t)(cols * YBLK), a_full + part);
                // chunk-loop kernel: the PREVIOUS tile's part has been multiplied (a_empty above), so its copy out of global
                // memory completed long ago, and this CTA was its only reader until the decoder overwrites the buffer two
                // chunks later: L2 may drop the lines instead of writing them to DRAM (16 lines of 128 bytes per column block)
                if (a.progress != nullptr && it > 0) {
                    for (int dd = 0; dd < 2; ++dd)
                        for (int l = lane; l < prev_valid * (YBLK / 128); l += 32)
                            tc::discard_l2_line(prev_img + yimg_block(prev_wg, dd, part, prev_t0, W) + (int64_t)l * 128);
                }
            }
            prev_img = img; prev_wg = wg; prev_t0 = t0; prev_valid = valid;
        }
        HB_ROLE_REPORT(0);
    } else if (warp == 4) {
        const uint32_t idesc = tc::idesc_f16_f32(128, NCLS);
        for (int64_t it = 0; heads_job(a, worker, n_workers, tiles_t, it, chunk, wg, t0); ++it) {
            const int acc = (int)(it & 1);
            const uint64_t a_hi = tc::smem_desc_sw128(tc::smem_u32(a_img), YBLK), a_lo = tc::smem_desc_sw128(tc::smem_u32(a_img + PART_BYTES), YBLK);
            const uint64_t w_hi = tc::smem_desc(tc::smem_u32(w_s), 128, 4096), w_lo = tc::smem_desc(tc::smem_u32(w_s + HEADS_WIMG), 128, 4096);
            // k-step ks: direction ks / 8 (forward units are k < 128), then (ks % 8) * 16 units into its slice
            auto koff = [](int ks) { return (uint64_t)(((ks >> 3) * SLICE_BYTES) / 16) + tc::sw128_kstep(ks & 7); };
            const uint32_t d = tmem + acc * NCLS;
            HB_TIMED(0, tc::mbar_wait(a_full + 0, (uint32_t)(it & 1)));
            if (it >= 2) HB_TIMED(1, tc::mbar_wait(acc_empty + acc, (uint32_t)((it / 2 - 1) & 1)));
            tc::tc_fence_after();
            if (tc::elect_one()) {                           // the two terms that read the hi part, then the part is free again
                uint32_t accum = 0;
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) { tc::mma_f16_ss(d, a_hi + koff(ks), w_hi + (uint64_t)(ks * 16), idesc, accum); accum = 1; }
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) tc::mma_f16_ss(d, a_hi + koff(ks), w_lo + (uint64_t)(ks * 16), idesc, 1);
                tc::mma_commit(a_empty + 0);
            }
            __syncwarp();
            HB_TIMED(0, tc::mbar_wait(a_full + 1, (uint32_t)(it & 1)));
            tc::tc_fence_after();
            if (tc::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) tc::mma_f16_ss(d, a_lo + koff(ks), w_hi + (uint64_t)(ks * 16), idesc, 1);
                tc::mma_commit(a_empty + 1);
                tc::mma_commit(acc_full + acc);
            }
            __syncwarp();
        }
        HB_ROLE_REPORT(1);
    } else {
        float bias[NCLS];
#pragma unroll
        for (int c = 0; c < NCLS; ++c) bias[c] = a.b_head[c];
        const int row = warp * 32 + lane;                       // position within the tile: (column, window)
        bool last_of_wg = false;
        for (int64_t it = 0; heads_job(a, worker, n_workers, tiles_t, it, chunk, wg, t0, &last_of_wg); ++it) {
            const int t = t0 + (row >> 3);
            const int64_t b = wg * WG + (row & 7);
            const int col = a.col0 + chunk * a.col_step + t;
            const int acc = (int)(it & 1);
            // p16 mode: the sums the earlier chunks left for this position are fetched BEFORE the wait for the accumulator (a
            // global round trip that used to follow it).  They were written by this CTA - window groups are pinned to workers -
            // possibly in the job just before this one and by another warp: the four epilogue warps meet first.
            const bool live = t < W && b < B;
            const int first_chunk = col < W ? 0 : (col - W) / a.col_step + 1;
            const int last_chunk = min(a.total_chunks - 1, col / a.col_step);
            float4* pp = reinterpret_cast<float4*>(a.p16 + (b * T + col) * NCLS);
            float4 old4[4];
            const bool fetch_old = a.p16 != nullptr && live && a.chunk0 + chunk != first_chunk;
            if (a.p16 != nullptr) {
                tc::named_barrier_sync(HEADS_EPILOGUE_BARRIER, 128);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) old4[c4] = fetch_old ? __ldcg(pp + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            HB_TIMED(0, tc::mbar_wait(acc_full + acc, (uint32_t)((it / 2) & 1)));
            tc::tc_fence_after();
            float v[NCLS];
            tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + acc * NCLS, v);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_empty + acc);
#ifdef HB_TIMELINE
            const long long t_math = acct ? clock64() : 0;
#endif
            if (live) {
                float mb = -INFINITY, mr = -INFINITY;
#pragma unroll
                for (int c = 0; c < NCLS; ++c) {
                    v[c] = fmaf(v[c], a.inv_scale, bias[c]);
                    if (c < NBASE) mb = fmaxf(mb, v[c]); else mr = fmaxf(mr, v[c]);
                }
                float sb = 0.f, sr = 0.f;
#pragma unroll
                for (int c = 0; c < NCLS; ++c) {
                    v[c] = expf(v[c] - (c < NBASE ? mb : mr));
                    if (c < NBASE) sb += v[c]; else sr += v[c];
                }
                if (a.p16 != nullptr) {
                    // one 64-byte record per position: 4 x 16-byte accesses instead of 16 scattered 4-byte ones, and no read
                    // at all when this chunk is the first one that covers the column (predict_gpu.py:137-149 adds into zeros)
#pragma unroll
                    for (int c = 0; c < NBASE; ++c) v[c] /= sb;
#pragma unroll
                    for (int c = NBASE; c < NCLS; ++c) v[c] /= sr;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 o = old4[c4];
                        v[4 * c4] += o.x; v[4 * c4 + 1] += o.y; v[4 * c4 + 2] += o.z; v[4 * c4 + 3] += o.w;
                    }
                    if (a.chunk0 + chunk == last_chunk) {
                        // the sums of this column are final: first-index argmax (strict >, classes in order), labels only
                        int ib = 0, ir = 0;
                        float vb = v[0], vr = v[NBASE];
#pragma unroll
                        for (int c = 1; c < NBASE; ++c) if (v[c] > vb) { vb = v[c]; ib = c; }
#pragma unroll
                        for (int c = 1; c < NRLE; ++c) if (v[NBASE + c] > vr) { vr = v[NBASE + c]; ir = c; }
                        a.base_label[b * T + col] = (uint8_t)ib;
                        a.rle_label[b * T + col] = (uint8_t)ir;
                    } else {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) pp[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
                    }
                } else {
                float* pb = a.p_base + (b * T + col) * NBASE;
                float* pr = a.p_rle + (b * T + col) * NRLE;
                float old[NCLS];                             // all 16 reads in flight at once (one global round trip, not 16)
#pragma unroll
                for (int c = 0; c < NBASE; ++c) old[c] = __ldcg(pb + c);
#pragma unroll
                for (int c = 0; c < NRLE; ++c) old[NBASE + c] = __ldcg(pr + c);
#pragma unroll
                for (int c = 0; c < NBASE; ++c) pb[c] = old[c] + v[c] / sb;
#pragma unroll
                for (int c = 0; c < NRLE; ++c) pr[c] = old[NBASE + c] + v[NBASE + c] / sr;
                }
            }
#ifdef HB_TIMELINE
            if (acct) t_wait[2] += clock64() - t_math;
#endif
            if (a.heads_done != nullptr && last_of_wg) {     // one fence per (chunk, window group), not per tile
                HB_TIMED(1, __threadfence());
                __syncwarp();
                if (lane == 0) tc::red_release_gpu_add(a.heads_done + wg, (unsigned long long)tiles_t);
            }
#ifdef HB_TIMELINE
            if (a.dbg != nullptr && worker == 0 && tid == 0) a.dbg[7100 + chunk] = (long long)globaltimer_ns();
#endif
        }
        if (warp == 0) HB_ROLE_REPORT(2);
    }
#undef HB_TIMED
#undef HB_ROLE_REPORT
    tc::tc_fence_before();
    tc::named_barrier_sync(2, HEADS_THREADS);
    if (warp == 4) tc::tmem_dealloc(tmem, 32);
}

__global__ void __launch_bounds__(HEADS_THREADS, 1)
tc_heads_kernel(const HeadsArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem_heads[];
    heads_role(a, tc::align_smem_1024(smem_heads), (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------
// The whole chunk loop of one batch in ONE launch (every role fits on the chip at once).
//   CTAs [0, 2R)           recurrence (window tile, direction): encoder and decoder phases alternate in the same CTA,
//                          h stays in registers, W_hh of the phase's layer is re-uploaded into TMEM (~1 us per switch)
//   CTAs [2R, 2R + 6P)     decoder input projection, one gate block each, W_ih(dec) block resident in TMEM;
//                          column tiles are projected as soon as both encoder directions have stored them
//   the rest               heads + softmax + accumulate, following the decoder's progress
// Hand-offs go through counters in global memory (release/acquire at gpu scope):
//   encoder columns done -> projection;  projection tiles done -> decoder (gi') and next encoder phase (yimg1 reuse);
//   decoder columns done -> heads;  heads done -> decoder (yimg2 buffer reuse).
// Every CTA is resident (grid <= SM count, one CTA per SM), so spinning is safe; per chunk the critical path is
// 100 encoder steps + the edge tiles of the projection + 100 decoder steps, with no launch boundary in between.
// Launched as clusters of 2 CTAs: the projection role pairs gate blocks (2b, 2b+1) and multicasts activation tiles.
// ---------------------------------------------------------------------------------------------
template <int N, int NLIVE, int MODE, int GW>
__global__ void __launch_bounds__((GW + 3) * 32, 1)
tc_chunkloop_kernel(const RecArgs rec, const ProjArgs proj, const HeadsArgs heads,
                    const int rec_ctas, const int proj_workers, const int heads_workers)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem_all = tc::align_smem_1024(smem_raw);
    const int bid = (int)blockIdx.x;
    static_assert((GW + 3) * 32 >= PROJ_THREADS && (GW + 3) * 32 >= HEADS_THREADS, "the block must hold every role");
    if (bid < 2 * rec_ctas) {
        recurrence_role<N, NLIVE, MODE, GW>(rec, smem_all, bid >> 1, bid & 1);
    } else if (bid < 2 * rec_ctas + 6 * proj_workers) {
        if (threadIdx.x >= PROJ_THREADS) {                   // spare warps only keep the pair's cluster barriers aligned
            if (proj.pair) { tc::cluster_sync_all(); tc::cluster_sync_all(); }
            return;
        }
        const int pb = bid - 2 * rec_ctas;
        projection_role<true>(proj, smem_all, pb % 6, pb / 6, proj_workers);
    } else {
        if (threadIdx.x >= HEADS_THREADS) return;
        heads_role(heads, smem_all, bid - 2 * rec_ctas - 6 * proj_workers, heads_workers);
    }
}

// The same kernel with two 8-window tiles per recurrence CTA (recurrence2_role): for batches of up to ~600 windows the
// recurrence CTAs cover 16 windows each and keep the tensor pipe busy with one tile's MMAs during the other tile's gate math.
__global__ void __launch_bounds__(REC_TC_THREADS, 1)
tc_chunkloop2_kernel(const RecArgs rec, const ProjArgs proj, const HeadsArgs heads,
                     const int rec_ctas, const int proj_workers, const int heads_workers)
{
    extern __shared__ __align__(1024) uint8_t smem_raw2[];
    uint8_t* smem_all = tc::align_smem_1024(smem_raw2);
    const int bid = (int)blockIdx.x;
    static_assert(REC_TC_THREADS >= PROJ_THREADS && REC_TC_THREADS >= HEADS_THREADS, "the block must hold every role");
    if (bid < 2 * rec_ctas) {
        recurrence2_role<8, 1>(rec, smem_all, bid >> 1, bid & 1);
    } else if (bid < 2 * rec_ctas + 6 * proj_workers) {
        if (threadIdx.x >= PROJ_THREADS) {
            if (proj.pair) { tc::cluster_sync_all(); tc::cluster_sync_all(); }
            return;
        }
        const int pb = bid - 2 * rec_ctas;
        projection_role<true>(proj, smem_all, pb % 6, pb / 6, proj_workers);
    } else {
        if (threadIdx.x >= HEADS_THREADS) return;
        heads_role(heads, smem_all, bid - 2 * rec_ctas - 6 * proj_workers, heads_workers);
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
struct TensorLayer {
    uint32_t* wih_tmem = nullptr;  // [6][hi, lo][128][Kp/2]
    float* scale_row = nullptr;    // [768]  projection epilogue scale (2^-kw [2^-10]) * (-log2e | -2 log2e)
    float* bias_row = nullptr;     // [768]  folded biases * (-log2e | -2 log2e)
    uint32_t* whh_tmem = nullptr;  // [2][hi, lo][3][128][64]
    float* gate_consts = nullptr;  // [2][4][128]
    int K = 0, Kp = 0;
};

// Switches read from the environment when a handle is created (A/B measurements and tests of every kernel variant;
// the defaults are the product configuration).
struct TensorTuning {
    bool pdl = true;            // HB_NO_PDL: no programmatic dependent launch
    bool pair = true;           // HB_NO_PAIR: projection without 2-CTA multicast clusters
    bool stack = true;          // HB_NO_STACK: 3-term recurrence MMAs instead of the stacked [h_hi | h_lo] operand
    bool live8 = true;          // HB_NO_LIVE8: never use the 8-live-window recurrence tile
    int windows_per_cta = 0;    // HB_WINDOWS_PER_CTA = 8 | 16 | 32: force the recurrence tile
    bool chunkloop = true;      // HB_NO_CHUNKLOOP: per-chunk launches even when the whole chunk loop fits on the chip
    int heads_workers = 0;      // HB_HEADS_WORKERS: CTAs of the heads role in the chunk-loop kernel (even)
    bool pixel_jobs = true;     // HB_NO_PIXEL_JOBS: project every image column before the chunk-loop kernel starts
    bool pingpong = true;       // HB_NO_PINGPONG: one window tile per recurrence CTA in the per-chunk kernels of large batches
    bool loop_pingpong = true;  // HB_NO_LOOP_PINGPONG: 16-window chunk-loop kernel with one 16-window tile per recurrence CTA instead of two 8-window tiles
    bool pixels_last = true;    // HB_PIXELS_FIRST: pixel jobs before the chunk's decoder tiles also at 8 windows per recurrence CTA
    bool cooperative = true;    // HB_NO_COOPERATIVE: chunk-loop kernel launched without the cooperative attribute
    int gate_warps = 8;         // HB_GATE_WARPS = 8 | 16: gate warps of the chunk-loop kernel's recurrence role (4 or 2 windows per thread at 8-window
                                // tiles).  Measured at B=256: 85.8 k windows/s with 8, 81.7 k with 16 (fewer warps share the per-step waits, TMEM loads,
                                // fences; with 16 the shared-memory fence before the arrive costs 150-190 cycles instead of ~90)
    static TensorTuning from_env() {
        TensorTuning t;
        t.pdl = getenv("HB_NO_PDL") == nullptr;
        t.pair = getenv("HB_NO_PAIR") == nullptr;
        t.stack = getenv("HB_NO_STACK") == nullptr;
        t.live8 = getenv("HB_NO_LIVE8") == nullptr;
        t.chunkloop = getenv("HB_NO_CHUNKLOOP") == nullptr;
        t.pixel_jobs = getenv("HB_NO_PIXEL_JOBS") == nullptr;
        t.pingpong = getenv("HB_NO_PINGPONG") == nullptr;
        t.loop_pingpong = getenv("HB_NO_LOOP_PINGPONG") == nullptr;
        t.pixels_last = getenv("HB_PIXELS_FIRST") == nullptr;
        t.cooperative = getenv("HB_NO_COOPERATIVE") == nullptr;
        if (const char* v = getenv("HB_GATE_WARPS")) { if (atoi(v) == 8 || atoi(v) == 16) t.gate_warps = atoi(v); }
        if (const char* v = getenv("HB_WINDOWS_PER_CTA")) {
            const int n = atoi(v);
            if (n == 8 || n == 16 || n == 32) t.windows_per_cta = n;
        }
        if (const char* v = getenv("HB_HEADS_WORKERS")) t.heads_workers = std::max(2, atoi(v) / 2 * 2);
        return t;
    }
};

struct TensorEngine {
    TensorTuning tune;
    hb_launch_plan last_plan{};
    int* proj_jobs = nullptr;                 // job table of the chunk-loop projection role (see ProjArgs::jobs)
    int* proj_job_offsets = nullptr;
    size_t proj_jobs_capacity = 0;            // fixed at creation: a batch whose table does not fit takes per-chunk launches
    int loop_max_ctas8 = 0, loop_max_ctas16 = 0;   // CTAs of the chunk-loop kernel (8- / 16-window tiles) that can be resident at once
    bool coop_with_pdl = true;                // cleared if the driver refuses cooperative + programmatic serialization together
    long long* phase_times = nullptr;         // HB_PHASE_TIMES=1: [64 phases][4] device stamps of the chunk-loop kernel, printed after the third call
    int phase_calls = 0;
    int jobs_w = -1, jobs_workers = -1, jobs_pixels = -1; int64_t jobs_n_wg = -1;
    int* proj_px_wgs = nullptr;               // [workers] window groups whose pixel jobs a worker owns
    std::vector<int> proj_jobs_host, proj_job_offsets_host, proj_px_wgs_host;
    int tile_order_w = -1;
    int* tile_order16 = nullptr;              // column-tile order of the heads role (earliest complete first)
    std::vector<int> tile_order16_host;
    unsigned long long* flags = nullptr;      // counters of the chunk-loop kernel
    size_t flags_capacity = 0;
    unsigned long long launch_epoch = 0;
    // optional device timing of every recurrence launch (the dominant kernel), bench.py's roofline pass
    bool time_recurrence = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> rec_events;
    size_t rec_events_used = 0;
    cudaStream_t side = nullptr;   // heads + softmax run here, beside the next chunk's encoder
    cudaEvent_t ev_dec[2] = {nullptr, nullptr}, ev_heads[2] = {nullptr, nullptr};
    TensorLayer enc, dec;
    __half* head_img = nullptr;    // [hi, lo][16 x 256]
    float* b_head = nullptr;
    float head_inv = 1.f;
    int features = 0;
    int sm_count = 0;
};

struct TensorWorkspace {
    float* gi_enc;      // [Bp, enc_cols, 768]  encoder projections of every image column, computed once per batch
    float* gi;          // [Bp, W, 768]         decoder projections of the current chunk
    uint8_t* yimg1;
    uint8_t* yimg2[2];  // double buffered: heads(k) runs beside the encoder of chunk k+1
    __half* ximg;
    float* hid_a;
    float* hid_b;
    float* p_base;
    float* p_rle;
    float* p16;         // [B, T, 16]           accumulated softmax sums of both heads, one 64-byte record per position
    size_t bytes;
};

// image columns covered by the chunk loop range(0, T, J) with W-column chunks (predict_gpu.py:114-117)
inline int covered_columns(int T, int W, int J) { return T < W ? 0 : (T - W) / J * J + W; }

inline TensorWorkspace tensor_carve(void* base, int64_t B, int T, int W, int Kp, int enc_cols) {
    TensorWorkspace ws{};
    size_t off = 0;
    auto take = [&](size_t n) {
        size_t o = off;
        off += align_up(n);
        return base ? static_cast<char*>(base) + o : nullptr;
    };
    const size_t Bp = ((size_t)B + 31) / 32 * 32;            // window tiles of the kernels may run past B
    const size_t Wp = (size_t)std::max(W, 0);
    ws.gi_enc = reinterpret_cast<float*>(take(Bp * (size_t)enc_cols * 2 * G * sizeof(float)));
    ws.gi = reinterpret_cast<float*>(take(Bp * Wp * 2 * G * sizeof(float)));
    ws.yimg1 = reinterpret_cast<uint8_t*>(take(Bp / WG * Wp * 2 * YROW));
    ws.yimg2[0] = reinterpret_cast<uint8_t*>(take(Bp / WG * Wp * 2 * YROW));
    ws.yimg2[1] = reinterpret_cast<uint8_t*>(take(Bp / WG * Wp * 2 * YROW));
    ws.ximg = reinterpret_cast<__half*>(take(Bp / WG * (size_t)T * Kp * 16));
    ws.hid_a = reinterpret_cast<float*>(take(Bp * 2 * H * sizeof(float)));
    ws.hid_b = reinterpret_cast<float*>(take(Bp * 2 * H * sizeof(float)));
    ws.p_base = reinterpret_cast<float*>(take((size_t)B * T * NBASE * sizeof(float)));
    ws.p_rle = reinterpret_cast<float*>(take((size_t)B * T * NRLE * sizeof(float)));
    ws.p16 = reinterpret_cast<float*>(take((size_t)B * T * NCLS * sizeof(float)));
    ws.bytes = off;
    return ws;
}

namespace detail {

template <class T> struct ident { using type = T; };

// Launch with (optionally) the programmatic-stream-serialization attribute: the kernel may start
// while its predecessor in the stream is still running and synchronises with griddepcontrol.wait.
// cooperative: the launch is scheduled only when the WHOLE grid can be resident at once (and fails with
// cudaErrorCooperativeLaunchTooLarge when it never can) - what a kernel whose CTAs wait for each other needs.
template <typename... KArgs>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, dim3 cluster,
                             bool cooperative, typename ident<KArgs>::type... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[3];
    int n = 0;
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cooperative) {
        attr[n].id = cudaLaunchAttributeCooperative;
        attr[n].val.cooperative = 1;
        ++n;
    }
    if (cluster.x * cluster.y * cluster.z > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster.x; attr[n].val.clusterDim.y = cluster.y; attr[n].val.clusterDim.z = cluster.z;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    void* params[] = {(void*)&args...};
    return cudaLaunchKernelExC(&cfg, (const void*)kernel, params);
}

template <typename... KArgs>
inline cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, dim3 cluster,
                                  typename ident<KArgs>::type... args) {
    return launch_ex(kernel, grid, block, smem, s, pdl, cluster, false, args...);
}

template <typename... KArgs>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                          typename ident<KArgs>::type... args) {
    return launch_ex(kernel, grid, block, smem, s, pdl, dim3(1, 1, 1), false, args...);
}

inline int pow2_scale_exponent(const float* w, size_t n) {
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(w[i]));
    if (!(mx > 0.f) || !isfinite(mx)) return 0;
    int e = (int)floorf(log2f(16384.0f / mx));      // max |w| 2^e < 2^14  (fp16 max 65504)
    if (e > 24) e = 24;
    if (e < -8) e = -8;
    return e;
}

inline void split_pair_words(float v0, float v1, uint32_t* hi_word, uint32_t* lo_word) {
    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
    const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
    *hi_word = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    *lo_word = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
}

template <typename T>
inline bool to_device(T** dst, const std::vector<T>& src, char* err, size_t errlen) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dst), src.size() * sizeof(T));
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        snprintf(err, errlen, "tensor engine: uploading weights failed: %s", cudaGetErrorString(e));
        return false;
    }
    return true;
}

// TMEM A-operand image of a [rows x K] matrix block: lane = row, 32-bit column c holds k = 2c (low half), 2c+1
inline bool pack_layer(const hb_gru_weights& g, int K, bool activations_scaled, TensorLayer* L, char* err, size_t errlen) {
    const int Kp = (K + 31) / 32 * 32;
    const int kwords = Kp / 2;
    L->K = K;
    L->Kp = Kp;
    std::vector<uint32_t> wih((size_t)6 * 2 * 128 * kwords, 0u), whh((size_t)2 * WHH_TMEM_WORDS, 0u);
    std::vector<float> scale_row(2 * G), bias_row(2 * G), consts((size_t)2 * 4 * H);
    for (int d = 0; d < 2; ++d) {
        const int e_ih = pow2_scale_exponent(g.weight_ih[d], (size_t)G * K);
        const float s_ih = ldexpf(1.f, e_ih);
        for (int r = 0; r < G; ++r) {
            const int blk = d * 3 + r / 128, rr = r % 128;
            for (int c = 0; c < kwords; ++c) {
                const float v0 = 2 * c < K ? g.weight_ih[d][(size_t)r * K + 2 * c] * s_ih : 0.f;
                const float v1 = 2 * c + 1 < K ? g.weight_ih[d][(size_t)r * K + 2 * c + 1] * s_ih : 0.f;
                split_pair_words(v0, v1, &wih[wih_word_index(blk, 0, rr, c, kwords)], &wih[wih_word_index(blk, 1, rr, c, kwords)]);
            }
            // gate-dependent folding: r, z rows carry -log2e and both biases; n rows carry -2 log2e and b_in only
            const bool is_n = r >= 2 * H;
            const float f = is_n ? -2.0f * LOG2E : -LOG2E;
            scale_row[d * G + r] = f * ldexpf(1.f, -e_ih) * (activations_scaled ? ACT_SCALE_INV : 1.f);
            bias_row[d * G + r] = f * (g.bias_ih[d][r] + (is_n ? 0.f : g.bias_hh[d][r]));
        }
        const int e_hh = pow2_scale_exponent(g.weight_hh[d], (size_t)G * H);
        const float s_hh = ldexpf(1.f, e_hh);
        for (int gb = 0; gb < 3; ++gb)
            for (int r = 0; r < 128; ++r)
                for (int c = 0; c < 64; ++c) {
                    const float* row = g.weight_hh[d] + (size_t)(gb * 128 + r) * H;
                    split_pair_words(row[2 * c] * s_hh, row[2 * c + 1] * s_hh, &whh[whh_word_index(d, 0, gb, r, c)], &whh[whh_word_index(d, 1, gb, r, c)]);
                }
        const float inv = ldexpf(1.f, -e_hh) * ACT_SCALE_INV;     // accumulator -> W_hh . h
        for (int j = 0; j < H; ++j) {
            consts[((size_t)d * 4 + 0) * H + j] = -LOG2E * inv;
            consts[((size_t)d * 4 + 1) * H + j] = -LOG2E * inv;
            consts[((size_t)d * 4 + 2) * H + j] = -2.0f * LOG2E * inv;
            consts[((size_t)d * 4 + 3) * H + j] = -2.0f * LOG2E * g.bias_hh[d][2 * H + j];
        }
    }
    return to_device(&L->wih_tmem, wih, err, errlen) && to_device(&L->scale_row, scale_row, err, errlen) &&
           to_device(&L->bias_row, bias_row, err, errlen) && to_device(&L->whh_tmem, whh, err, errlen) &&
           to_device(&L->gate_consts, consts, err, errlen);
}

inline void free_layer(TensorLayer* L) {
    cudaFree(L->wih_tmem); cudaFree(L->scale_row); cudaFree(L->bias_row); cudaFree(L->whh_tmem); cudaFree(L->gate_consts);
}

// (+1024: the kernels align their shared memory to the swizzle atom themselves)
inline size_t projection_smem(int blk_bytes, int parts) {
    return (size_t)PROJ_STAGES * parts * 8 * blk_bytes + 256 + PROJ_TABLE_MAX * 5 + PROJ_GROUPS_MAX * 16 + 1024;
}
template <int N, int NLIVE = N>
constexpr size_t recurrence_smem() { return (size_t)2 * h_buffers<N>() * (N / WG) * YBLK + (size_t)gi_stages<NLIVE>() * NLIVE * GI_ROW_BYTES + 512 + 1024; }
constexpr size_t heads_smem() { return (size_t)2 * 16 * YROW + 2 * HEADS_WIMG + 128 + 1024; }

}  // namespace detail

inline void tensor_engine_destroy(TensorEngine* e) {
    if (!e) return;
    detail::free_layer(&e->enc);
    detail::free_layer(&e->dec);
    cudaFree(e->head_img);
    cudaFree(e->b_head);
    cudaFree(e->proj_jobs);
    cudaFree(e->proj_job_offsets);
    cudaFree(e->proj_px_wgs);
    cudaFree(e->tile_order16);
    cudaFree(e->flags);
    cudaFree(e->phase_times);
    if (e->side) cudaStreamDestroy(e->side);
    for (auto& ev : e->rec_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (int i = 0; i < 2; ++i) {
        if (e->ev_dec[i]) cudaEventDestroy(e->ev_dec[i]);
        if (e->ev_heads[i]) cudaEventDestroy(e->ev_heads[i]);
    }
    delete e;
}

inline TensorEngine* tensor_engine_create(const hb_weights* w, int features, int sm_count, char* err, size_t errlen) {
    TensorEngine* e = new TensorEngine();
    e->tune = TensorTuning::from_env();
    e->features = features;
    e->sm_count = sm_count;
    bool ok = detail::pack_layer(w->encoder, features, /*activations_scaled=*/false, &e->enc, err, errlen) &&
              detail::pack_layer(w->decoder, 2 * H, /*activations_scaled=*/true, &e->dec, err, errlen);
    if (ok) {
        // head rows 0..4 = dense1_base, 5..15 = dense2_rle; B-operand image [16 x 256], dense core matrices
        std::vector<float> wh((size_t)NCLS * 2 * H), bh(NCLS);
        std::memcpy(wh.data(), w->base_weight, (size_t)NBASE * 2 * H * sizeof(float));
        std::memcpy(wh.data() + (size_t)NBASE * 2 * H, w->rle_weight, (size_t)NRLE * 2 * H * sizeof(float));
        std::memcpy(bh.data(), w->base_bias, NBASE * sizeof(float));
        std::memcpy(bh.data() + NBASE, w->rle_bias, NRLE * sizeof(float));
        const int eh = detail::pow2_scale_exponent(wh.data(), wh.size());
        const float sh = ldexpf(1.f, eh);
        std::vector<__half> img((size_t)2 * NCLS * 2 * H);
        for (int r = 0; r < NCLS; ++r)
            for (int k = 0; k < 2 * H; ++k) {
                const float v = wh[(size_t)r * 2 * H + k] * sh;
                const __half hi = __float2half_rn(v);
                const size_t off = tc::core_offset(r, k, 128, 4096) / 2;
                img[off] = hi;
                img[(size_t)NCLS * 2 * H + off] = __float2half_rn(v - __half2float(hi));
            }
        e->head_inv = ldexpf(1.f, -eh) * ACT_SCALE_INV;
        ok = detail::to_device(&e->head_img, img, err, errlen) && detail::to_device(&e->b_head, bh, err, errlen);
    }
    if (ok) {
        cudaError_t ce = cudaSuccess;
        auto set = [&](const void* fn, size_t bytes) {
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        };
        set((const void*)tc_projection_kernel<false>, detail::projection_smem(e->enc.Kp * 16, 1));
        set((const void*)tc_projection_kernel<true>, detail::projection_smem(YROW, 2));
        set((const void*)tc_recurrence_kernel<16, 8, 0>, detail::recurrence_smem<16, 8>());
        set((const void*)tc_recurrence_kernel<16, 16, 0>, detail::recurrence_smem<16>());
        set((const void*)tc_recurrence_kernel<16, 8, 1>, detail::recurrence_smem<16, 8>());
        set((const void*)tc_recurrence_kernel<16, 16, 1>, detail::recurrence_smem<16>());
        set((const void*)tc_recurrence_kernel<32, 32, 0>, detail::recurrence_smem<32>());
        set((const void*)tc_recurrence2_kernel<8, 1>, recurrence2_smem<8>());
        set((const void*)tc_recurrence2_kernel<16, 0>, recurrence2_smem<16>());
        set((const void*)tc_heads_kernel, detail::heads_smem());
        const size_t loop8 = std::max({detail::recurrence_smem<16, 8>(), detail::projection_smem(YROW, 2), detail::heads_smem()});
        const size_t loop16 = std::max({detail::recurrence_smem<16>(), detail::projection_smem(YROW, 2), detail::heads_smem()});
        set((const void*)tc_chunkloop_kernel<16, 8, 1, 16>, loop8);
        set((const void*)tc_chunkloop_kernel<16, 8, 1, 8>, loop8);
        set((const void*)tc_chunkloop_kernel<16, 8, 0, 16>, loop8);
        set((const void*)tc_chunkloop_kernel<16, 16, 1, 16>, loop16);
        set((const void*)tc_chunkloop_kernel<16, 16, 1, 8>, loop16);
        set((const void*)tc_chunkloop_kernel<16, 16, 0, 16>, loop16);
        const size_t loop2 = std::max({recurrence2_smem<8>(), detail::projection_smem(YROW, 2), detail::heads_smem()});
        set((const void*)tc_chunkloop2_kernel, loop2);
        // the chunk-loop kernel's CTAs wait for each other: how many of them fit on the chip at once (one per SM here, but
        // MPS limits, other resident kernels' reservations or a smaller part change that; the plan respects the answer)
        auto resident = [&](const void* fn, size_t bytes, int* out) {
            int per_sm = 0;
            if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, REC_TC_THREADS, bytes);
            *out = per_sm * sm_count;
        };
        resident((const void*)tc_chunkloop_kernel<16, 8, 1, 16>, loop8, &e->loop_max_ctas8);
        resident((const void*)tc_chunkloop_kernel<16, 16, 1, 16>, loop16, &e->loop_max_ctas16);
        int max2 = 0;
        resident((const void*)tc_chunkloop2_kernel, loop2, &max2);
        e->loop_max_ctas16 = std::min(e->loop_max_ctas16, max2);
        // job table of the chunk-loop projection role, sized once for the largest batch the kernel takes (one 8-window
        // group per two SMs) and chunks of up to 1024 columns; anything larger runs as per-chunk launches
        e->proj_jobs_capacity = (size_t)(sm_count / 2 + 1) * (2 * 128 + PX_R);
        if (ce == cudaSuccess) ce = cudaMalloc(reinterpret_cast<void**>(&e->proj_jobs), e->proj_jobs_capacity * sizeof(int));
        if (ce == cudaSuccess) ce = cudaMalloc(reinterpret_cast<void**>(&e->proj_job_offsets), 256 * sizeof(int));
        if (ce == cudaSuccess) ce = cudaMalloc(reinterpret_cast<void**>(&e->proj_px_wgs), 256 * sizeof(int));
        if (ce == cudaSuccess) ce = cudaMalloc(reinterpret_cast<void**>(&e->tile_order16), 4096 * sizeof(int));
        e->flags_capacity = (size_t)1 << 16;
        if (ce == cudaSuccess) ce = cudaMalloc(reinterpret_cast<void**>(&e->flags), e->flags_capacity * sizeof(unsigned long long));
        if (ce == cudaSuccess && getenv("HB_PHASE_TIMES") != nullptr) {
            ce = cudaMalloc(reinterpret_cast<void**>(&e->phase_times), 256 * sizeof(long long));
            if (ce == cudaSuccess) ce = cudaMemset(e->phase_times, 0, 256 * sizeof(long long));
        }
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking);
        for (int i = 0; i < 2 && ce == cudaSuccess; ++i) {
            ce = cudaEventCreateWithFlags(&e->ev_dec[i], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_heads[i], cudaEventDisableTiming);
        }
        if (ce != cudaSuccess) {
            snprintf(err, errlen, "tensor engine: cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(ce));
            ok = false;
        }
    }
    if (!ok) {
        tensor_engine_destroy(e);
        return nullptr;
    }
    return e;
}

inline size_t tensor_engine_workspace_bytes(const TensorEngine* e, int64_t B, int T, int W) {
    return tensor_carve(nullptr, B, T, W, e->enc.Kp, T).bytes;     // J >= 1: at most T covered columns
}

// live windows per recurrence CTA: the smallest tile (lowest step latency) that still fits the batch on the chip
inline int pick_windows_per_cta(const TensorTuning& tune, int64_t B, int sm_count) {
    if (tune.windows_per_cta) return tune.windows_per_cta;
    if (tune.live8 && 2 * B <= (int64_t)8 * sm_count) return 8;
    return 2 * B <= (int64_t)16 * sm_count ? 16 : 32;
}

// Role split of the chunk-loop kernel: recurrence CTAs (8 windows each, or 16 as two 8-window tiles that take turns on the
// tensor pipe), the rest shared between projection workers (6 CTAs each) and heads workers.  Returns false when the
// batch is too large for it: the projection role has to keep pace with the encoder.  Measured (windows/s, one box):
//   B=320  8-window tiles, 10 workers 99.7 k
//   B=384  two tiles 102.5 k | per-chunk launches 98.3 k        B=448  two tiles ~105 k | per-chunk 109 k (before the y-store fix)
//   B=512  two tiles 127.7-128.8 k | one 16-window tile 118 k | per-chunk launches 115.4 k
//   B=576  two tiles, 11 workers 111 k | per-chunk launches 119 k
// so: 8-window tiles while 10 workers fit (B <= 320), two tiles per CTA while 12 workers fit (B <= 512), per-chunk launches
// above.  HB_WINDOWS_PER_CTA forces a tile (tests).
// (Measured and dropped: starting the odd window tiles one phase after the even ones, so that the other roles see two waves
// of work per chunk - the encoder phases got 3 us shorter, the decoder phases 3 us longer, the launch took the same time.)
struct ChunkloopPlan { int tile, rec_ctas, proj_workers, heads_workers; };
inline bool plan_chunkloop_tile(const TensorTuning& tune, int tile, int min_workers, int64_t B, int sm_count, int max_resident8, int max_resident16,
                                ChunkloopPlan* plan) {
    const int sms = sm_count / 2 * 2;
    const int64_t rec = (B + tile - 1) / tile;
    if (2 * rec > sms) return false;
    if (sms > (tile == 8 ? max_resident8 : max_resident16)) return false;   // the grid (one CTA per SM) must be co-resident as a whole
    const int left = sms - (int)(2 * rec);
    // heads: one CTA in seven of what the recurrence leaves, at least 2 (a heads tile is a load -> MMA -> epilogue chain of
    // ~3 us with only the epilogue overlapped; measured at B=256 in one run: 4 or 6 heads CTAs 89.9-90.0 k windows/s,
    // 8 or 12 (both give 12 projection workers + 12 heads CTAs) 91.6-91.7 k - with fewer the heads fall a chunk behind and
    // the kernel ends with a longer tail)
    int heads = tune.heads_workers ? tune.heads_workers : std::max(2, (left / 14) * 2);
    int proj = (left - heads) / 6;
    if (proj < min_workers) return false;
    heads = left - 6 * proj;                                   // whatever the 6-CTA granularity leaves goes to the heads
    heads = heads / 2 * 2;
    if (heads < 2) { --proj; heads += 6; }
    plan->tile = tile; plan->rec_ctas = (int)rec; plan->proj_workers = proj; plan->heads_workers = heads;
    return true;
}
inline bool plan_chunkloop(const TensorTuning& tune, int64_t B, int sm_count, int max_resident8, int max_resident16, ChunkloopPlan* plan) {
    if (tune.windows_per_cta) {
        if (tune.windows_per_cta != 8 && tune.windows_per_cta != 16) return false;     // (32 live windows leave no room for two gi' rows per window)
        return plan_chunkloop_tile(tune, tune.windows_per_cta, 6, B, sm_count, max_resident8, max_resident16, plan);
    }
    if (tune.live8 && plan_chunkloop_tile(tune, 8, 10, B, sm_count, max_resident8, max_resident16, plan)) return true;
    return tune.stack && tune.loop_pingpong && plan_chunkloop_tile(tune, 16, 12, B, sm_count, max_resident8, max_resident16, plan);
}

// Returns the number of kernel launches issued, or a negative hb_status (message in err).
inline int tensor_engine_predict(TensorEngine* e, const uint8_t* images, int64_t B, int T, int W, int J,
                                 uint8_t* base_labels, uint8_t* rle_labels, float* base_prob, float* rle_prob,
                                 void* workspace, cudaStream_t s, char* err, size_t errlen) {
    const int enc_cols = covered_columns(T, W, J);
    TensorWorkspace ws = tensor_carve(workspace, B, T, W, e->enc.Kp, T);
    float* p_base = base_prob ? base_prob : ws.p_base;
    float* p_rle = rle_prob ? rle_prob : ws.p_rle;
    const int F = e->features;
    const int64_t n_wg = (B + WG - 1) / WG;
    int launches = 0;
#define TE_CUDA(expr)                                                                                          \
    do {                                                                                                       \
        cudaError_t te_ = (expr);                                                                              \
        if (te_ != cudaSuccess) {                                                                              \
            snprintf(err, errlen, "tensor engine: %s failed: %s (line %d)", #expr, cudaGetErrorString(te_), __LINE__); \
            return HB_ERR_CUDA;                                                                                \
        }                                                                                                      \
    } while (0)
    // the reference's two arrays only when the caller wants the probabilities back; else one 64-byte record per position,
    // written by the first chunk that covers a column and turned into labels by the last one (no fill, no argmax launch)
    float* p16 = (base_prob || rle_prob) ? nullptr : ws.p16;
    if (p16) {
        if (enc_cols < T) {                                    // columns no chunk covers keep all-zero sums: label 0
            TE_CUDA(cudaMemsetAsync(base_labels, 0, (size_t)B * T, s));
            TE_CUDA(cudaMemsetAsync(rle_labels, 0, (size_t)B * T, s));
        }
    } else {
        TE_CUDA(cudaMemsetAsync(p_base, 0, (size_t)B * T * NBASE * sizeof(float), s));
        TE_CUDA(cudaMemsetAsync(p_rle, 0, (size_t)B * T * NRLE * sizeof(float), s));
    }
    const int xblk = e->enc.Kp * 16;                           // bytes of one (group, column) pixel block
    const int proj_workers = std::max(1, e->sm_count / 6);
    const bool pdl = e->tune.pdl;
    const int pair_mode = e->tune.pair ? 1 : 0;               // 2-CTA clusters share activation tiles by multicast
    const int n_chunks = T < W ? 0 : (T - W) / J + 1;
    const int tiles8 = (W + 7) / 8, tiles16 = (W + 15) / 16;
    // ---- how the batch is laid out on the chip ----
    ChunkloopPlan plan{};
    bool chunkloop = e->tune.chunkloop && n_chunks > 0 && B > 0 && tiles8 <= 128 &&
                     plan_chunkloop(e->tune, B, e->sm_count, e->loop_max_ctas8, e->loop_max_ctas16, &plan);
    // pixel jobs: the projection role also projects the image columns the next chunk's encoder adds
    const int px_tiles = (enc_cols + 7) / 8, px_pre_tiles = std::min(px_tiles, tiles8);
    bool pixels_in_loop = chunkloop && e->tune.pixel_jobs && n_chunks > 1 && e->enc.Kp <= 128 && px_tiles < 4096 && (J + 7) / 8 + 1 <= PX_R;
    size_t n_groups = 0, flags_needed = 0;
    if (chunkloop) {
        n_groups = (size_t)plan.rec_ctas * (plan.tile / WG);       // window groups the recurrence CTAs cover (>= n_wg)
        flags_needed = (size_t)4 * plan.rec_ctas + n_groups + n_groups * tiles8 * 2 + n_groups * tiles8;
        if (pixels_in_loop && flags_needed + n_groups * px_tiles * 2 > e->flags_capacity) pixels_in_loop = false;
        if (pixels_in_loop) flags_needed += n_groups * px_tiles * 2;
        const int64_t worker_tiles = (n_wg * tiles8 + plan.proj_workers - 1) / plan.proj_workers;    // tiles one projection worker owns
        const int64_t worker_jobs = worker_tiles + (n_wg + plan.proj_workers - 1) / plan.proj_workers * PX_R;
        chunkloop = flags_needed <= e->flags_capacity && worker_jobs <= PROJ_TABLE_MAX && n_wg <= PROJ_GROUPS_MAX &&
                    (size_t)n_wg * (tiles8 + (pixels_in_loop ? PX_R : 0)) <= e->proj_jobs_capacity;   // tables sized at creation / in shared memory
        pixels_in_loop = pixels_in_loop && chunkloop;
    }
    if (enc_cols > 0) {
        // once per batch: pixels -> operand image, then the encoder projection of every covered column (chunks overlap
        // by W - J columns and gi of a column does not depend on the chunk) -- or, with pixel jobs in the chunk loop, of
        // the first chunk's columns only
        const int64_t chunks16 = n_wg * T * (e->enc.Kp / 8) * WG;
        const int blocks = (int)std::min<int64_t>((chunks16 + 255) / 256, (int64_t)e->sm_count * 16);
        pileup_to_operand_image_kernel<<<blocks, 256, 0, s>>>(images, B, T, F, e->enc.Kp, ws.ximg, n_wg);
        TE_CUDA(cudaGetLastError());
        const int tiles = (int)std::min<int64_t>(n_wg * (pixels_in_loop ? px_pre_tiles : px_tiles), proj_workers);
        ProjArgs pe{};
        pe.col_tiles = pixels_in_loop ? px_pre_tiles : 0;
        pe.in_base = reinterpret_cast<const uint8_t*>(ws.ximg); pe.in_wg_stride = (int64_t)T * xblk; pe.in_dir_stride = 0; pe.in_part_stride = 0; pe.n_dirs = 1;
        pe.blk_bytes = xblk; pe.lbo = 128; pe.Kp = e->enc.Kp; pe.n_wg = n_wg; pe.W = enc_cols;
        pe.w_tmem = e->enc.wih_tmem; pe.scale_row = e->enc.scale_row; pe.bias_row = e->enc.bias_row; pe.gi = ws.gi_enc;
        pe.pair = pair_mode;
        TE_CUDA(detail::launch_cluster(tc_projection_kernel<false>, dim3(tiles, 6), dim3(PROJ_THREADS), detail::projection_smem(xblk, 1), s, pdl,
                                       dim3(1, pair_mode ? 2 : 1, 1), pe));
        launches += 2;
    }
    // Timeline instrumentation exists only in the -DHB_TIMELINE build of the library (tools/build_variants.py
    // timeline=-DHB_TIMELINE, selected with HB_LIB=...); HB_DEBUG_TIMELINE=1 | e then records decoder | encoder steps.
    static long long* dbg_buf = nullptr;
#ifdef HB_TIMELINE
    static const bool dbg_on = getenv("HB_DEBUG_TIMELINE") != nullptr;
    static const bool dbg_enc = dbg_on && getenv("HB_DEBUG_TIMELINE")[0] == 'e';
    if (dbg_on && !dbg_buf) { cudaMalloc(&dbg_buf, 10240 * sizeof(long long)); cudaMemset(dbg_buf, 0, 10240 * sizeof(long long)); }
#else
    constexpr bool dbg_on = false, dbg_enc = false;
#endif
    // decoder projection arguments (the same for every chunk)
    ProjArgs pd{};
    pd.in_base = ws.yimg1; pd.in_wg_stride = (int64_t)W * 4 * YBLK; pd.in_dir_stride = (int64_t)W * 2 * YBLK; pd.in_part_stride = (int64_t)W * YBLK;
    pd.blk_bytes = YBLK; pd.n_dirs = 2; pd.lbo = 0; pd.Kp = e->dec.Kp; pd.n_wg = n_wg; pd.W = W;
    pd.w_tmem = e->dec.wih_tmem; pd.scale_row = e->dec.scale_row; pd.bias_row = e->dec.bias_row; pd.gi = ws.gi;
    pd.pair = pair_mode;
    auto layer_args = [&](const TensorLayer& L, const float* gi, int gi_cols, int gi_col0, int gi_col_step, uint8_t* y_even, uint8_t* y_odd) {
        RecLayer r{};
        r.gi = gi; r.gi_cols = gi_cols; r.gi_col0 = gi_col0; r.gi_col_step = gi_col_step;
        r.whh_tmem = L.whh_tmem; r.gate_consts = L.gate_consts; r.yimg[0] = y_even; r.yimg[1] = y_odd;
        return r;
    };
    HeadsArgs heads_base{};
    heads_base.n_wg = n_wg; heads_base.B = B; heads_base.W = W; heads_base.T = T; heads_base.col_step = J;
    heads_base.w_img = e->head_img; heads_base.b_head = e->b_head; heads_base.inv_scale = e->head_inv;
    heads_base.p_base = p_base; heads_base.p_rle = p_rle; heads_base.p16 = p16;
    heads_base.total_chunks = n_chunks; heads_base.base_label = base_labels; heads_base.rle_label = rle_labels;

    // hb_enable_kernel_timing(2): CUDA events around every launch of the dominant kernel (no launch overlap then)
    auto dominant_begin = [&]() -> size_t {
        // measurement mode only (hb_enable_kernel_timing(2) created the events); launches beyond them go untimed
        if (!e->time_recurrence || e->rec_events_used == e->rec_events.size()) return (size_t)-1;
        const size_t slot = e->rec_events_used++;
        cudaEventRecord(e->rec_events[slot].first, s);
        return slot;
    };
    auto dominant_end = [&](size_t slot) {
        if (slot != (size_t)-1) cudaEventRecord(e->rec_events[slot].second, s);
    };
    // ---- chunk-loop kernel: every role of the whole chunk loop resident at once ----
    bool two_tiles = false;
    if (chunkloop) {
        if (e->tile_order_w != W) {
            // heads: a column tile is complete once the forward pass is past its last column and the reverse pass past its first
            const int n = tiles16;
            std::vector<int>& order = e->tile_order16_host;            // member vector: outlives the async copy
            order.resize(n);
            for (int i = 0; i < n; ++i) order[i] = i;
            auto ready = [&](int tt) { return std::max(std::min(16 * tt + 16, W), W - 16 * tt); };
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return ready(x) < ready(y); });
            TE_CUDA(cudaMemcpyAsync(e->tile_order16, order.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
            e->tile_order_w = W;
        }
        if (e->jobs_w != W || e->jobs_workers != plan.proj_workers || e->jobs_n_wg != n_wg || e->jobs_pixels != (int)pixels_in_loop) {
            // projection job table: tile (group, tile) belongs to worker (group + tile) % workers, which spreads every tile
            // index (in particular the edge tiles both decoder directions start with) evenly over the workers; a worker's
            // entries are sorted by (tile, group), the order in which its forward-direction CTAs scan them (see the scheduler)
            struct Job { int tile, wg, packed; };
            std::vector<std::vector<Job>> per(plan.proj_workers);
            for (int64_t wg = 0; wg < n_wg; ++wg)
                for (int t = 0; t < tiles8; ++t)
                    per[(wg + t) % plan.proj_workers].push_back({t, (int)wg, pack_proj_job((int)wg, t)});
            e->proj_jobs_host.clear();
            e->proj_job_offsets_host.assign(1, 0);
            e->proj_px_wgs_host.assign(plan.proj_workers, 0);
            for (int wk = 0; wk < plan.proj_workers; ++wk) {
                auto& v = per[wk];
                if (pixels_in_loop) {
                    // pixel entries first (they wait for nothing and fill the role's idle start of the encoder phase):
                    // relative tile r of every window group the worker owns, tile-major so that a chunk that adds
                    // fewer than PX_R tiles uses a prefix of them
                    std::vector<int> mine;
                    for (int64_t wg = wk; wg < n_wg; wg += plan.proj_workers) mine.push_back((int)wg);
                    e->proj_px_wgs_host[wk] = (int)mine.size();
                    for (int r = 0; r < PX_R; ++r)
                        for (int wg : mine) e->proj_jobs_host.push_back(pack_proj_job(wg, r) | (1 << 29));
                }
                std::stable_sort(v.begin(), v.end(), [](const Job& x, const Job& y) { return x.tile != y.tile ? x.tile < y.tile : x.wg < y.wg; });
                for (const Job& jb : v) e->proj_jobs_host.push_back(jb.packed);
                e->proj_job_offsets_host.push_back((int)e->proj_jobs_host.size());
            }
            if (e->proj_jobs_host.size() > e->proj_jobs_capacity) {
                snprintf(err, errlen, "tensor engine: projection job table of %zu entries exceeds its capacity %zu", e->proj_jobs_host.size(), e->proj_jobs_capacity);
                return HB_ERR_CUDA;
            }
            // (only when the batch shape changed since the last call; the tables are a few KB)
            TE_CUDA(cudaMemcpyAsync(e->proj_jobs, e->proj_jobs_host.data(), e->proj_jobs_host.size() * sizeof(int), cudaMemcpyHostToDevice, s));
            TE_CUDA(cudaMemcpyAsync(e->proj_job_offsets, e->proj_job_offsets_host.data(), e->proj_job_offsets_host.size() * sizeof(int), cudaMemcpyHostToDevice, s));
            TE_CUDA(cudaMemcpyAsync(e->proj_px_wgs, e->proj_px_wgs_host.data(), e->proj_px_wgs_host.size() * sizeof(int), cudaMemcpyHostToDevice, s));
            e->jobs_w = W; e->jobs_workers = plan.proj_workers; e->jobs_n_wg = n_wg; e->jobs_pixels = (int)pixels_in_loop;
        }
        TE_CUDA(cudaMemsetAsync(e->flags, 0, flags_needed * sizeof(unsigned long long), s));
        unsigned long long* f = e->flags;
        unsigned long long* enc_prog = f;                 f += 2 * plan.rec_ctas;
        unsigned long long* dec_prog = f;                 f += 2 * plan.rec_ctas;
        unsigned long long* heads_done = f;               f += n_groups;               // one per window group
        unsigned long long* tile_flags = f;               f += n_groups * tiles8 * 2;  // [group][tile][dec direction]
        unsigned long long* tile_reads = f;               f += n_groups * tiles8;      // [group][tile]
        unsigned long long* px_flags = f;                                              // [group][image column tile][enc direction]
        RecArgs ra{};
        ra.layer[0] = layer_args(e->enc, ws.gi_enc, enc_cols, 0, J, ws.yimg1, ws.yimg1);
        ra.layer[0].progress = enc_prog; ra.layer[0].consumed_flags = tile_flags; ra.layer[0].consumed_per_chunk = PROJ_FLAGS_PER_TILE;
        ra.layer[1] = layer_args(e->dec, ws.gi, W, 0, 0, ws.yimg2[0], ws.yimg2[1]);
        ra.layer[1].progress = dec_prog; ra.layer[1].tile_flags = tile_flags;
        ra.layer[1].flag_tiles = tiles8; ra.layer[1].flag_abs = 0; ra.layer[1].flag_skip_tiles = 0;
        ra.layer[1].flag_need_base = 0; ra.layer[1].flag_need_per_chunk = PROJ_FLAGS_PER_TILE;
        if (pixels_in_loop) {
            ra.layer[0].tile_flags = px_flags; ra.layer[0].flag_tiles = px_tiles; ra.layer[0].flag_abs = 1;
            ra.layer[0].flag_skip_tiles = px_pre_tiles; ra.layer[0].flag_need_base = PROJ_FLAGS_PER_TILE; ra.layer[0].flag_need_per_chunk = 0;
        }
        ra.layer[1].heads_done = heads_done; ra.layer[1].heads_per_chunk = 4 * tiles16;
        ra.n_layers = 2; ra.n_chunks = n_chunks; ra.h_in = nullptr; ra.h_out = nullptr; ra.B = B; ra.W = W;
        ra.tiles_t = tiles8; ra.n_wg = (int)n_wg; ra.epoch = 0; ra.dbg = dbg_buf; ra.dbg_layer = dbg_enc ? 0 : 1;
        ra.phase_times = e->phase_times;
        ProjArgs pp = pd;
        pp.progress = enc_prog; pp.epoch = 0; pp.rec_n = plan.tile; pp.n_chunks = n_chunks; pp.tile_flags = tile_flags;
        pp.tile_reads = tile_reads;
        pp.pair = 0;          // every CTA of the role picks its own job order (see the loader): no shared tiles
        pp.dbg = dbg_buf;
        pp.jobs = e->proj_jobs; pp.job_offsets = e->proj_job_offsets; 
        if (pixels_in_loop) {
            pp.px.ximg = reinterpret_cast<const uint8_t*>(ws.ximg); pp.px.wg_stride = (int64_t)T * xblk; pp.px.blk_bytes = xblk; pp.px.Kp = e->enc.Kp;
            pp.px.w_tmem = e->enc.wih_tmem; pp.px.scale_row = e->enc.scale_row; pp.px.bias_row = e->enc.bias_row;
            pp.px.gi = ws.gi_enc; pp.px.cols = enc_cols; pp.px.tiles = px_tiles; pp.px.col_step = J;
            pp.px.flags = px_flags; pp.px.wgs_of_worker = e->proj_px_wgs;
            // (pixel jobs last only while the projection role has slack: measured at B=320, 10 workers: 101.1 k windows/s with
            // them last, 106.9 k first)
            pp.px.last = (plan.tile == 8 && plan.proj_workers >= 12 && e->tune.pixels_last) ? 1 : 0;
        }
        HeadsArgs hp = heads_base;
        hp.yimg = ws.yimg2[0]; hp.yimg_odd = ws.yimg2[1]; hp.col0 = 0; hp.n_chunks = n_chunks;
        hp.progress = dec_prog; hp.rec_n = plan.tile; hp.tile_order = e->tile_order16; hp.heads_done = heads_done;
        hp.dbg = dbg_buf;
        const dim3 grid(2 * plan.rec_ctas + 6 * plan.proj_workers + plan.heads_workers), cluster(1, 1, 1);
        two_tiles = plan.tile == 16 && e->tune.stack && e->tune.loop_pingpong;
        const size_t smem = std::max({plan.tile == 8 ? detail::recurrence_smem<16, 8>() : (two_tiles ? recurrence2_smem<8>() : detail::recurrence_smem<16>()),
                                      detail::projection_smem(YROW, 2), detail::heads_smem()});
        const int gw = (e->tune.stack && !two_tiles) ? e->tune.gate_warps : 16;
        // Cooperative launch: the CTAs of this kernel spin on each other's counters, so the grid must be resident as a
        // whole - with the attribute the launch waits until it can be (another handle's kernel, or any other work on the
        // device, cannot leave it half scheduled) instead of relying on an idle chip.
        cudaError_t le = cudaSuccess;
        auto go = [&](auto kernel) {
            const size_t slot = dominant_begin();
            const bool coop = e->tune.cooperative;
            bool use_pdl = pdl && !e->time_recurrence && (!coop || e->coop_with_pdl);
            le = detail::launch_ex(kernel, grid, dim3((gw + 3) * 32), smem, s, use_pdl, cluster, coop,
                                   ra, pp, hp, plan.rec_ctas, plan.proj_workers, plan.heads_workers);
            if (le != cudaSuccess && coop && use_pdl) {       // some drivers refuse the two attributes together
                cudaGetLastError();
                e->coop_with_pdl = false;
                le = detail::launch_ex(kernel, grid, dim3((gw + 3) * 32), smem, s, false, cluster, coop,
                                       ra, pp, hp, plan.rec_ctas, plan.proj_workers, plan.heads_workers);
            }
            dominant_end(slot);
        };
        if (plan.tile == 8) {
            if (!e->tune.stack) go(tc_chunkloop_kernel<16, 8, 0, 16>);
            else if (gw == 8) go(tc_chunkloop_kernel<16, 8, 1, 8>);
            else go(tc_chunkloop_kernel<16, 8, 1, 16>);
        } else if (two_tiles) {
            go(tc_chunkloop2_kernel);
        } else {
            if (!e->tune.stack) go(tc_chunkloop_kernel<16, 16, 0, 16>);
            else if (gw == 8) go(tc_chunkloop_kernel<16, 16, 1, 8>);
            else go(tc_chunkloop_kernel<16, 16, 1, 16>);
        }
        TE_CUDA(le);
        launches += 1;
    }

    // ---- per-chunk launches (batches too large for the chip, or HB_NO_CHUNKLOOP) ----
    const float* hid = nullptr;
    float* hid_bufs[2] = {ws.hid_a, ws.hid_b};
    int flip = 0;
    const int nrec = pick_windows_per_cta(e->tune, B, e->sm_count);
    const dim3 grid_rec((unsigned)((B + nrec - 1) / nrec), 2);
    const int tiles_proj = (int)std::min<int64_t>(n_wg * ((W + 7) / 8), proj_workers);
    const int tiles_heads = (int)std::min<int64_t>(n_wg * ((W + 15) / 16), e->sm_count);
    auto rec_args = [&](const TensorLayer& L, const float* gi, int gi_cols, int gi_col0, const float* h_in, float* h_out, uint8_t* yimg) {
        RecArgs ra{};
        ra.layer[0] = layer_args(L, gi, gi_cols, gi_col0, 0, yimg, yimg);
        ra.n_layers = 1; ra.n_chunks = 1; ra.h_in = h_in; ra.h_out = h_out; ra.B = B; ra.W = W;
        ra.tiles_t = tiles8; ra.n_wg = (int)n_wg; ra.epoch = 0; ra.dbg_layer = 0;
        ra.dbg = (nrec <= 16 && (dbg_enc == (&L == &e->enc))) ? dbg_buf : nullptr;
        return ra;
    };
    auto recurrence = [&](const RecArgs& ra, bool use_pdl) {
        const size_t slot = dominant_begin();
        if (e->time_recurrence) use_pdl = false;
        const bool stack = e->tune.stack;
        if (e->tune.pingpong && nrec == 16 && stack)          // two 8-window stacked tiles per CTA
            detail::launch(tc_recurrence2_kernel<8, 1>, grid_rec, dim3(REC_TC_THREADS), recurrence2_smem<8>(), s, use_pdl, ra);
        else if (e->tune.pingpong && nrec == 32)              // two 16-window 3-term tiles per CTA
            detail::launch(tc_recurrence2_kernel<16, 0>, grid_rec, dim3(REC_TC_THREADS), recurrence2_smem<16>(), s, use_pdl, ra);
        else if (nrec == 8)
            detail::launch(stack ? tc_recurrence_kernel<16, 8, 1> : tc_recurrence_kernel<16, 8, 0>, grid_rec, dim3(REC_TC_THREADS),
                           detail::recurrence_smem<16, 8>(), s, use_pdl, ra);
        else if (nrec == 16)
            detail::launch(stack ? tc_recurrence_kernel<16, 16, 1> : tc_recurrence_kernel<16, 16, 0>, grid_rec, dim3(REC_TC_THREADS),
                           detail::recurrence_smem<16>(), s, use_pdl, ra);
        else
            detail::launch(tc_recurrence_kernel<32, 32, 0>, grid_rec, dim3(REC_TC_THREADS), detail::recurrence_smem<32>(), s, use_pdl, ra);
        dominant_end(slot);
    };
    e->last_plan = hb_launch_plan{};
    e->last_plan.chunkloop = chunkloop ? 1 : 0;
    e->last_plan.windows_per_cta = chunkloop ? plan.tile : nrec;
    e->last_plan.stacked_operand = (e->tune.stack && e->last_plan.windows_per_cta <= 16) ? 1 : 0;
    e->last_plan.recurrence_ctas = chunkloop ? plan.rec_ctas : (int)grid_rec.x;
    e->last_plan.projection_workers = chunkloop ? plan.proj_workers : tiles_proj;
    e->last_plan.heads_workers = chunkloop ? plan.heads_workers : tiles_heads;
    e->last_plan.cooperative = (chunkloop && e->tune.cooperative) ? 1 : 0;
    int chunk = 0;
    for (int i = 0; !chunkloop && i + W <= T; i += J, ++chunk) {
        float* enc_h = hid_bufs[flip];
        float* dec_h = hid_bufs[flip ^ 1];
        const int buf = chunk & 1;
        if (chunk >= 2) TE_CUDA(cudaStreamWaitEvent(s, e->ev_heads[buf], 0));      // heads(chunk - 2) has read this yimg2 buffer
        recurrence(rec_args(e->enc, ws.gi_enc, enc_cols, i, hid, enc_h, ws.yimg1), pdl && chunk == 0);
        detail::launch_cluster(tc_projection_kernel<true>, dim3(tiles_proj, 6), dim3(PROJ_THREADS), detail::projection_smem(YROW, 2), s, pdl,
                               dim3(1, pair_mode ? 2 : 1, 1), pd);
        recurrence(rec_args(e->dec, ws.gi, W, 0, enc_h, dec_h, ws.yimg2[buf]), pdl);
        TE_CUDA(cudaEventRecord(e->ev_dec[buf], s));
        TE_CUDA(cudaStreamWaitEvent(e->side, e->ev_dec[buf], 0));
        HeadsArgs ha = heads_base;
        ha.yimg = ws.yimg2[buf]; ha.col0 = i; ha.chunk0 = chunk;
        tc_heads_kernel<<<tiles_heads, HEADS_THREADS, detail::heads_smem(), e->side>>>(ha);
        TE_CUDA(cudaGetLastError());
        TE_CUDA(cudaEventRecord(e->ev_heads[buf], e->side));
        launches += 4;
        hid = dec_h;
        flip ^= 1;
    }
    for (int b = 0; b < std::min(chunk, 2); ++b) TE_CUDA(cudaStreamWaitEvent(s, e->ev_heads[b], 0));   // join the side stream
    const int64_t positions = B * T;
    if (!p16) {                                                // probabilities requested: labels from the returned arrays
        argmax_kernel<<<(unsigned)((positions + 255) / 256), 256, 0, s>>>(p_base, p_rle, positions, base_labels, rle_labels);
        launches += 1;
    } else if (n_chunks == 0) {                                // T < W: no chunk at all, labels 0 (memset above covered it)
    }
    if (e->phase_times != nullptr && chunkloop && ++e->phase_calls == 3) {
        // HB_PHASE_TIMES=1 (measurement runs): where the recurrence CTA 0 / forward spent the launch, from device timestamps
        cudaStreamSynchronize(s);
        std::vector<long long> pt(256);
        cudaMemcpy(pt.data(), e->phase_times, pt.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const int n_ph = std::min(2 * n_chunks, 64);
        double start[2] = {0, 0}, steps[2] = {0, 0}, drain[2] = {0, 0}, gap = 0;
        for (int ph = 0; ph < n_ph; ++ph) {
            const long long* q = &pt[ph * 4];
            start[ph & 1] += (q[1] - q[0]) * 1e-3; steps[ph & 1] += (q[2] - q[1]) * 1e-3; drain[ph & 1] += (q[3] - q[2]) * 1e-3;
            if (ph + 1 < n_ph) gap += (pt[(ph + 1) * 4] - q[3]) * 1e-3;
        }
        const double nc = n_ph / 2.0;
        fprintf(stderr, "[phase times, CTA 0 forward, us per chunk over %d chunks; %d recurrence CTAs x 2, %d projection workers, %d heads workers]\n"
                        "  encoder: start %.2f | %d steps %.2f (%.3f us/step) | drain %.2f\n  decoder: start %.2f | %d steps %.2f (%.3f us/step) | drain %.2f\n"
                        "  between phases %.2f | chunk %.2f | first phase start -> last phase end %.1f us\n",
                n_ph / 2, plan.rec_ctas, plan.proj_workers, plan.heads_workers,
                start[0] / nc, W, steps[0] / nc, steps[0] / nc / W, drain[0] / nc, start[1] / nc, W, steps[1] / nc, steps[1] / nc / W, drain[1] / nc,
                gap / nc, (pt[(n_ph - 1) * 4 + 3] - pt[0]) * 1e-3 / nc, (pt[(n_ph - 1) * 4 + 3] - pt[0]) * 1e-3);
    }
    if (dbg_on && dbg_buf) {
        static int printed = 0;
        cudaStreamSynchronize(s);
        if (printed++ == 2) {
            std::vector<long long> hbuf(10240);
            cudaMemcpy(hbuf.data(), dbg_buf, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            if (chunkloop) {
                const long long t0 = hbuf[4096];
                fprintf(stderr, "[chunk-loop kernel: tile %d, %d recurrence CTAs x 2, %d projection workers x 6, %d heads workers]\n",
                        plan.tile, plan.rec_ctas, plan.proj_workers, plan.heads_workers);
                fprintf(stderr, "[CTA 0 fwd, us since encoder phase 0 start]\n");
                for (int k = 0; k < std::min(n_chunks, 6); ++k)
                    fprintf(stderr, "  chunk %d: enc %.1f -> %.1f   dec %.1f -> %.1f   projection worker 0 done %.1f   heads worker 0 done %.1f\n", k,
                            (hbuf[4096 + k * 2] - t0) * 1e-3, (hbuf[4096 + k * 2 + 1] - t0) * 1e-3,
                            (hbuf[4096 + (64 + k) * 2] - t0) * 1e-3, (hbuf[4096 + (64 + k) * 2 + 1] - t0) * 1e-3,
                            (hbuf[7000 + k] - t0) * 1e-3, (hbuf[7100 + k] - t0) * 1e-3);
                {
                    const char* role[4] = {"loader   ", "MMA      ", "epilogue ", "scheduler"};
                    const char* what[4][4] = {{"a_empty", "job ring", "-", "issue copies"}, {"a_full", "acc_empty", "-", "-"},
                                              {"acc_full", "gpu fence", "-", "-"}, {"loader lead", "progress", "-", "-"}};
                    fprintf(stderr, "  projection worker 0 / block 0, cycles waiting over the whole launch (%d chunks):\n", n_chunks);
                    for (int r = 0; r < 4; ++r) {
                        const long long* w = &hbuf[7200 + 8 * r];
                        fprintf(stderr, "    %s total %lld:", role[r], w[4]);
                        for (int k = 0; k < 4; ++k) fprintf(stderr, "  %s %lld", what[r][k], w[k]);
                        fprintf(stderr, "\n");
                    }
                }
                {
                    const char* role[3] = {"loader  ", "MMA     ", "epilogue"};
                    const char* what[3][3] = {{"a_empty", "progress", "-"}, {"a_full", "acc_empty", "-"}, {"acc_full", "fence", "softmax + P update"}};
                    fprintf(stderr, "  heads worker 0, cycles over the whole launch:\n");
                    for (int r = 0; r < 3; ++r) {
                        const long long* w = &hbuf[7300 + 8 * r];
                        fprintf(stderr, "    %s total %lld:", role[r], w[4]);
                        for (int k = 0; k < 3; ++k) fprintf(stderr, "  %s %lld", what[r][k], w[k]);
                        fprintf(stderr, "\n");
                    }
                }
                for (int li = 0; li < 2; ++li) {
                    const long long* st = &hbuf[6144 + li * 8];
                    fprintf(stderr, "  chunk 2 %s phase, cycles after phase start: first step released %lld | last step done %lld | last image "
                            "handed to copy engine %lld | all columns published %lld | phase end %lld\n", li ? "dec" : "enc",
                            st[1] - st[0], st[2] - st[0], st[3] - st[0], st[4] - st[0], st[5] - st[0]);
                }
                if (n_chunks > 0)
                    fprintf(stderr, "  last chunk: dec end %.1f\n", (hbuf[4096 + (64 + n_chunks - 1) * 2 + 1] - t0) * 1e-3);
                if (!two_tiles && n_chunks > 2) {
                    const long long c0 = hbuf[7408];
                    auto us = [&](int k) { return (hbuf[7400 + k] - c0) * 1e-3; };
                    fprintf(stderr, "  chunk 2, window group 0, us after the forward encoder's last step: reverse encoder's last step %.2f | final publication fwd %.2f rev %.2f | "
                            "scheduler picks decoder tile 0 %.2f | copies issued %.2f | MMAs committed %.2f | counter raised %.2f | decoder loader saw it %.2f | first gi row in shared memory %.2f\n",
                            us(9), us(0), us(1), us(2), us(3), us(4), us(5), us(6), us(7));
                }
            }
#ifdef HB_TIMELINE_STEPS
            if (two_tiles) {
                // two tiles per recurrence CTA: the MMA issuer's step pair and each tile's first gate warp, cycles after the
                // issuer's release for that tile's step
                double iss[4] = {0}, g[2][8] = {{0}};
                int n = 0;
                for (int st = 20; st < 90; ++st, ++n) {
                    const long long* q0 = &hbuf[(st * 2) * 2];
                    iss[0] += q0[1] - q0[0]; iss[1] += q0[2] - q0[1]; iss[2] += q0[3] - q0[2]; iss[3] += q0[4] - q0[3];
                    for (int tl = 0; tl < 2; ++tl)
                        for (int k = 0; k < 8; ++k) g[tl][k] += hbuf[1024 + (tl * 128 + st) * 8 + k] - hbuf[(st * 2 + tl) * 2];
                }
                fprintf(stderr, "[two-tile step pair, MMA issuer, cycles] tile 0 issue %.0f | wait for tile 1 %.0f | tile 1 issue %.0f | wait for tile 0 %.0f | pair %.0f\n",
                        iss[0] / n, iss[1] / n, iss[2] / n, iss[3] / n, (iss[0] + iss[1] + iss[2] + iss[3]) / n);
                fprintf(stderr, "  y-store warp, cycles per step pair: y_ready waits %.0f | store issue %.0f | wait for the previous stores' reads %.0f | publication (every 4th) %.0f\n",
                        hbuf[3000] / 70.0, hbuf[3001] / 70.0, hbuf[3002] / 70.0, hbuf[3003] / 70.0);
                fprintf(stderr, "  y-store warp: release store of the progress counter %.0f cycles per step pair\n", hbuf[3004] / 70.0);
                const char* names[7] = {"gi+h_free", "acc_r", "acc_z", "acc_n", "ldtm_n", "math_done", "arrived"};
                for (int tl = 0; tl < 2; ++tl) {
                    fprintf(stderr, "  tile %d gate warp 0, cycles after the issuer's release of the step:", tl);
                    fprintf(stderr, " gi_full=%.0f", g[tl][7] / n);
                    for (int k = 0; k < 7; ++k) fprintf(stderr, " %s=%.0f", names[k], g[tl][k] / n);
                    fprintf(stderr, "\n");
                }
            } else {
            // per-step stamps of gate warps 0 and GW-1 (roles 1, 2): cycles after the SAME warp's arrival of the previous step
            // (the MMA issuer carries no stamps: they changed its code and slowed the issue loop)
            auto at = [&](int role, int st, int k) { return hbuf[((size_t)role * 128 + st) * 8 + k]; };
            for (int role = 1; role <= 2; ++role) {
                double acc[7] = {0};
                int n = 0;
                for (int st = 20; st < 90; ++st, ++n)
                    for (int k = 0; k < 7; ++k) acc[k] += at(role, st, k) - at(role, st - 1, 6);
                const char* names[7] = {"gi+h_free", "acc_r", "acc_z", "acc_n", "ldtm_n", "math_done", "arrived(=step)"};
                fprintf(stderr, "[step timeline, gate warp %s, cycles after its previous arrival]", role == 1 ? "0" : "last");
                for (int k = 0; k < 7; ++k) fprintf(stderr, " %s=%.0f", names[k], acc[k] / n);
                fprintf(stderr, "\n");
            }
            fprintf(stderr, "  arrival of every gate warp relative to warp 0:");
            for (int w = 0; w < 16; ++w) {
                double a_w = 0;
                for (int st = 20; st < 90; ++st) a_w += hbuf[8192 + w * 128 + st] - hbuf[8192 + st];
                if (hbuf[8192 + w * 128 + 20] != 0) fprintf(stderr, " %d:%+.0f", w, a_w / 70);
            }
            fprintf(stderr, "\n");
            }
#endif
        }
    }
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
        snprintf(err, errlen, "tensor engine launch failed: %s", cudaGetErrorString(ce));
        return HB_ERR_CUDA;
    }
    e->last_plan.launches = launches;
    return launches;
#undef TE_CUDA
}

}  // namespace hb
