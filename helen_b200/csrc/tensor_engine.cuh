// Tensor engine: the GRU contractions on tcgen05 tensor cores (sm_100a), fp32-equivalent
// results through a 3-term split of fp16 operands:
//     W = (W_hi + W_lo) * 2^-kw ,  x = (x_hi + x_lo) * 2^-10          (hi, lo in fp16)
//     W.x  ~=  (W_hi.x_hi + W_lo.x_hi + W_hi.x_lo) * 2^-(kw+10)       fp32 accumulate in TMEM
// uint8 pileup pixels are exact in fp16, so the encoder input projection needs only 2 terms.
// (tools/precision_probe.py: max |dP| 2.6e-7 vs fp64, same as plain fp32; bf16 3-term: 1.3e-5.)
//
// Operand orientation (both kernels): A = weights, 128 gate rows per MMA (M=128), stationary in
// shared memory; B = activations, N = data rows (windows or window-columns), K-major; D[gate
// row, data row] in TMEM, so a thread's TMEM lane is "its" gate row / hidden unit.
//
//   tc_projection_kernel   gi[m, 0:768] = A[m, 0:K] . Wcat^T + b_ih         (bulk GEMM)
//   tc_recurrence_kernel   W dependent GRU steps; per step 72 MMAs (3 gate blocks x 8 k-steps x
//                          3 split terms) + gate math on the TMEM accumulators
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/helen_b200.h"
#include "fp32_kernels.cuh"
#include "tc_ptx.cuh"
#include "workspace.cuh"

namespace hb {

constexpr float ACT_SCALE = 1024.0f;              // activations in (-1, 1) are scaled by 2^10 before the split
constexpr float ACT_SCALE_INV = 1.0f / 1024.0f;
constexpr int W_LBO = 128;                        // weight images: dense core matrices
constexpr int H_LBO = 144;                        // h operand: padded so the gate threads' 2-byte stores spread over banks
constexpr int H_SBO = 16 * H_LBO;                 // 8-window group stride of the h operand (K = 128 -> 16 core matrices)
constexpr int WHH_IMG_HALFS = G * H;              // one [384 x 128] fp16 image
constexpr int WHH_SBO = (H / 8) * W_LBO;          // 2048

// ---------------------------------------------------------------------------------------------
// Bulk input projection on tensor cores.
// grid = (row-tile workers, 6 gate blocks); each CTA keeps its [128 x Kp] weight block (hi, lo) in
// shared memory and walks over NT-row activation tiles.
// ---------------------------------------------------------------------------------------------
template <typename TA, int NT>
__global__ void __launch_bounds__(256, 1)
tc_projection_kernel(const TA* __restrict__ a, int64_t a_batch_stride, int64_t a_row_stride, int rows_per_window,
                     int64_t M, int K, int Kp,
                     const __half* __restrict__ w_img,      // [6 blocks][hi, lo][128 * Kp] core-matrix images
                     const float* __restrict__ bias,        // [768]
                     const float* __restrict__ inv_scale,   // [6]  2^-(kw [+10])
                     float* __restrict__ gi)                // [M, 768]
{
    constexpr bool kSplitA = sizeof(TA) == 4;               // fp32 activations need a lo term; uint8 is exact
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int blk = blockIdx.y;
    const int kg = Kp >> 3;                                  // 16-byte k groups per row
    const uint32_t w_bytes = 128u * Kp * 2u;                 // one weight image
    const uint32_t a_bytes = (uint32_t)NT * Kp * 2u;         // one activation image
    const uint32_t sbo = (uint32_t)kg * 128u;
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + w_bytes;
    uint8_t* a_hi = smem + 2 * w_bytes;
    uint8_t* a_lo = a_hi + a_bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(a_lo + (kSplitA ? a_bytes : 0));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    {   // weight block: straight copy of the pre-packed image
        const int4* src = reinterpret_cast<const int4*>(w_img + (size_t)blk * 2 * 128 * Kp);
        int4* dst = reinterpret_cast<int4*>(w_hi);
        for (uint32_t i = tid; i < 2 * w_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(tmem_slot, NT < 32 ? 32 : NT);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = tc::idesc_f16_f32(128, NT);
    const float inv = inv_scale[blk];
    const float my_bias = bias[blk * 128 + (warp & 3) * 32 + lane];
    const int ksteps = Kp >> 4;
    uint32_t phase = 0;

    const int64_t n_tiles = (M + NT - 1) / NT;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * NT;
        // ---- stage the activation tile as fp16 hi (/lo) core matrices ----
        for (int e = tid; e < NT * kg; e += blockDim.x) {
            const int r = e / kg, g8 = e - r * kg;
            const int64_t m = row0 + r;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
            if (m < M) {
                const int64_t b = m / rows_per_window, t = m - b * rows_per_window;
                const TA* src = a + b * a_batch_stride + t * a_row_stride + g8 * 8;
                if constexpr (kSplitA) {
                    if (g8 * 8 + 8 <= K) {
                        float4 p0 = *reinterpret_cast<const float4*>(src), p1 = *reinterpret_cast<const float4*>(src + 4);
                        v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) if (g8 * 8 + i < K) v[i] = (float)src[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (g8 * 8 + i < K) v[i] = (float)src[i];
                }
            }
            const uint32_t off = (uint32_t)(r >> 3) * sbo + (uint32_t)g8 * 128u + (uint32_t)(r & 7) * 16u;
            __half hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if constexpr (kSplitA) tc::split_f16(v[i] * ACT_SCALE, hi[i], lo[i]);
                else hi[i] = __float2half_rn(v[i]);
            }
            *reinterpret_cast<int4*>(a_hi + off) = *reinterpret_cast<int4*>(hi);
            if constexpr (kSplitA) *reinterpret_cast<int4*>(a_lo + off) = *reinterpret_cast<int4*>(lo);
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint64_t whi = tc::smem_desc(tc::smem_u32(w_hi), W_LBO, sbo), wlo = tc::smem_desc(tc::smem_u32(w_lo), W_LBO, sbo);
                const uint64_t ahi = tc::smem_desc(tc::smem_u32(a_hi), W_LBO, sbo), alo = tc::smem_desc(tc::smem_u32(a_lo), W_LBO, sbo);
                uint32_t acc = 0;
                const int terms = kSplitA ? 3 : 2;
                for (int term = 0; term < terms; ++term) {
                    const uint64_t wa = term == 1 ? wlo : whi;
                    const uint64_t aa = term == 2 ? alo : ahi;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        tc::mma_f16_ss(tmem, wa + (uint64_t)(ks * 16), aa + (uint64_t)(ks * 16), idesc, acc);
                        acc = 1;
                    }
                }
                tc::mma_commit(bar);
            }
            __syncwarp();
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::tc_fence_after();
        // ---- epilogue: warp w reads lane quarter w%4, column half w/4 ----
        constexpr int COLS = NT / 2;
        const int c0 = (warp >> 2) * COLS;
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
        float* out = gi + blk * 128 + (warp & 3) * 32 + lane;
#pragma unroll
        for (int c = 0; c < COLS; c += 8) {
            float v[8];
            tc::tmem_ld8(taddr + c, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t m = row0 + c0 + c + i;
                if (m < M) out[m * (2 * G)] = fmaf(v[i], inv, my_bias);
            }
        }
        tc::tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tc::tmem_dealloc(tmem, NT < 32 ? 32 : NT);
}

// ---------------------------------------------------------------------------------------------
// Recurrence on tensor cores.  One CTA = N windows x one direction x W dependent steps of one
// layer.
//   * W_hh (fp16 hi and lo, 3 gate blocks x 128 rows x 128 k) is loaded ONCE into tensor memory
//     and used as the TMEM-resident A operand of every MMA (384 of the 512 columns).  With A in
//     shared memory each M=128,K=16 MMA would re-read 4 KB of weights (32 cycles of smem
//     bandwidth) regardless of N; from TMEM the MMA runs at the tensor rate for small N.
//   * the state h lives in registers (fp32, one hidden unit per thread) and is re-published each
//     step as the fp16 hi/lo B operand in shared memory (K-major core matrices).
//   * accumulators r | z | n : TMEM columns [0, 3N).
// Warps 0..7 are gate warps: (warp w, lane l) owns hidden unit j = 32 (w%4) + l (== its TMEM lane)
// for windows [(w/4) N/2, (w/4+1) N/2).  Warp 8 issues the MMAs.  Handshake per step:
//   gate warps --h_ready (8 warp arrivals)--> MMA warp --tcgen05.commit acc_ready--> gate warps
// ---------------------------------------------------------------------------------------------
constexpr int REC_GATE_WARPS = 16;
constexpr int REC_TC_THREADS = (REC_GATE_WARPS + 1) * 32;
constexpr int TMEM_W_COL0 = 128;                  // weight columns start here (accumulators below)
constexpr int WHH_TMEM_WORDS = 2 * 3 * 128 * 64;  // per direction: [term][gate block][row][k pair]

__device__ __forceinline__ float ld_nc_f32(const float* p) {   // asm volatile: stays where it is written
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int N>
__global__ void __launch_bounds__(REC_TC_THREADS, 1)
tc_recurrence_kernel(const float* __restrict__ gi,          // [B*W, 768] (b_ih already added)
                     const uint32_t* __restrict__ whh_tmem, // [2 dirs][hi, lo][3][128][64] packed fp16 pairs
                     const float* __restrict__ b_hh,        // [2][384]
                     const float* __restrict__ inv_scale,   // [2]  2^-(kw + 10)
                     const float* __restrict__ h_in,        // [B, 2, 128] or nullptr
                     float* __restrict__ h_out,             // [B, 2, 128]
                     float* __restrict__ y,                 // [B*W, 256] fp32, or nullptr
                     __half* __restrict__ y_hi,             // [B*W, 256] fp16 split of y * 2^10, or nullptr
                     __half* __restrict__ y_lo,
                     int64_t B, int W)
{
    static_assert(N == 16 || N == 32, "N windows per CTA (3N accumulator columns must stay below TMEM_W_COL0)");
    constexpr int NW = N / 4;                                // windows per gate thread
    constexpr uint32_t HB_BYTES = (N / 8) * H_SBO;           // one h operand image
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* h_hi = smem;
    uint8_t* h_lo = smem + HB_BYTES;
    uint64_t* acc_ready = reinterpret_cast<uint64_t*>(h_lo + HB_BYTES);   // [3]: r, z, n blocks
    uint64_t* h_ready = acc_ready + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int64_t b0 = (int64_t)blockIdx.x * N;

    if (tid == 0) {
        tc::mbar_init(acc_ready + 0, 1); tc::mbar_init(acc_ready + 1, 1); tc::mbar_init(acc_ready + 2, 1);
        tc::mbar_init(h_ready, REC_GATE_WARPS);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == REC_GATE_WARPS) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == REC_GATE_WARPS) {
        // ===================== MMA issuer =====================
        __syncthreads();                                     // weights in TMEM, h_0 in smem
        tc::tc_fence_after();
        const uint32_t idesc = tc::idesc_f16_f32(128, N);
        const uint64_t hhi_desc = tc::smem_desc(tc::smem_u32(h_hi), H_LBO, H_SBO);
        const uint64_t hlo_desc = tc::smem_desc(tc::smem_u32(h_lo), H_LBO, H_SBO);
        for (int s = 0; s < W; ++s) {
            if (s > 0) {
                tc::mbar_wait(h_ready, (uint32_t)((s - 1) & 1));
                tc::tc_fence_after();
            }
            if (tc::elect_one()) {
#pragma unroll
                for (int gb = 0; gb < 3; ++gb) {             // gate blocks r, z, n
#pragma unroll
                    for (int term = 0; term < 3; ++term) {   // (W_hi,h_hi) (W_lo,h_hi) (W_hi,h_lo)
                        const uint32_t a_col = tmem + TMEM_W_COL0 + ((term == 1 ? 3 : 0) + gb) * 64;
                        const uint64_t bd = term == 2 ? hlo_desc : hhi_desc;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            tc::mma_f16_ts(tmem + gb * N, a_col + ks * 8, bd + (uint64_t)(ks * 2 * H_LBO / 16), idesc,
                                           (term | ks) != 0);
                    }
                    tc::mma_commit(acc_ready + gb);          // gates start on r while z, n still run
                }
            }
            __syncwarp();
        }
    } else {
        // ===================== gate warps =====================
        const int q = warp & 3;
        const int j = q * 32 + lane;
        const int win0 = (warp >> 2) * NW;
        if (warp < 8) {   // W_hh -> TMEM: warps 0-3 store the hi image, warps 4-7 the lo image; thread = row j
            const int term = warp >> 2;
            const uint32_t* src = whh_tmem + (size_t)dir * WHH_TMEM_WORDS + (size_t)term * 3 * 128 * 64;
#pragma unroll 1
            for (int gb = 0; gb < 3; ++gb) {
#pragma unroll
                for (int c = 0; c < 64; c += 16) {
                    uint32_t r[16];
                    const uint4* p = reinterpret_cast<const uint4*>(src + ((size_t)gb * 128 + j) * 64 + c);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const uint4 x = p[v];
                        r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
                    }
                    tc::tmem_st16(tmem + ((uint32_t)(q * 32) << 16) + TMEM_W_COL0 + (term * 3 + gb) * 64 + c, r);
                }
            }
            tc::tmem_st_wait();
        }
        const float inv = inv_scale[dir];
        const float bhr = b_hh[dir * G + j], bhz = b_hh[dir * G + H + j], bhn = b_hh[dir * G + 2 * H + j];
        // gi / y rows of windows past B exist in the (padded) workspace, so only h_in / h_out,
        // which may be caller tensors of exactly B windows, need guarding.
        float h_own[NW];
        uint32_t h_off[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            const int64_t b = b0 + win0 + i;
            h_own[i] = (h_in != nullptr && b < B) ? h_in[(b * 2 + dir) * H + j] : 0.f;
            h_off[i] = tc::core_offset(win0 + i, j, H_LBO, H_SBO);
            __half hi, lo;
            tc::split_f16(h_own[i] * ACT_SCALE, hi, lo);
            *reinterpret_cast<__half*>(h_hi + h_off[i]) = hi;
            *reinterpret_cast<__half*>(h_lo + h_off[i]) = lo;
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        __syncthreads();                                     // pairs with the MMA warp's second barrier

        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)win0;
        const int wstride_gi = W * 2 * G, wstride_y = W * 2 * H;   // elements between consecutive windows
        const float* gi_thr = gi + (b0 + win0) * (int64_t)wstride_gi + dir * G + j;
        const int64_t y_thr = (b0 + win0) * (int64_t)wstride_y + dir * H + j;
        int t = dir ? W - 1 : 0;
        const int dt = dir ? -1 : 1;
        for (int s = 0; s < W; ++s, t += dt) {
            // this step's input projections: issued now, consumed after the MMA waits
            float gir[NW], giz[NW], gin[NW];
            const float* gp = gi_thr + t * (2 * G);
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const float* p = gp + i * wstride_gi;
                gir[i] = ld_nc_f32(p); giz[i] = ld_nc_f32(p + H); gin[i] = ld_nc_f32(p + 2 * H);
            }
            const uint32_t par = (uint32_t)(s & 1);
            float r[NW], z[NW], a[NW];
            tc::mbar_wait(acc_ready + 0, par);
            tc::tc_fence_after();
            if constexpr (NW == 4) tc::tmem_ld4(taddr, a); else tc::tmem_ld8(taddr, a);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < NW; ++i) r[i] = tc::sigmoid_fast(gir[i] + fmaf(a[i], inv, bhr));
            tc::mbar_wait(acc_ready + 1, par);
            tc::tc_fence_after();
            if constexpr (NW == 4) tc::tmem_ld4(taddr + N, a); else tc::tmem_ld8(taddr + N, a);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < NW; ++i) z[i] = tc::sigmoid_fast(giz[i] + fmaf(a[i], inv, bhz));
            tc::mbar_wait(acc_ready + 2, par);
            tc::tc_fence_after();
            if constexpr (NW == 4) tc::tmem_ld4(taddr + 2 * N, a); else tc::tmem_ld8(taddr + 2 * N, a);
            tc::tmem_ld_wait();
            const int64_t yo = y_thr + t * (2 * H);
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const float n = tc::tanh_fast(gin[i] + r[i] * fmaf(a[i], inv, bhn));
                const float hn = fmaf(z[i], h_own[i] - n, n);   // (1 - z) n + z h
                h_own[i] = hn;
                __half hi, lo;
                tc::split_f16(hn * ACT_SCALE, hi, lo);
                *reinterpret_cast<__half*>(h_hi + h_off[i]) = hi;
                *reinterpret_cast<__half*>(h_lo + h_off[i]) = lo;
                if (y != nullptr) y[yo + i * wstride_y] = hn;
                if (y_hi != nullptr) { y_hi[yo + i * wstride_y] = hi; y_lo[yo + i * wstride_y] = lo; }
            }
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(h_ready);
        }
#pragma unroll
        for (int i = 0; i < NW; ++i)
            if (b0 + win0 + i < B) h_out[((b0 + win0 + i) * 2 + dir) * H + j] = h_own[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == REC_GATE_WARPS) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
struct TensorLayer {
    __half* wih_img = nullptr;     // [6][hi, lo][128 * Kp]
    float* wih_inv = nullptr;      // [6]
    float* bih = nullptr;          // [768]
    uint32_t* whh_tmem = nullptr;  // [2][hi, lo][3][128][64] packed fp16 pairs (TMEM A-operand image)
    float* whh_inv = nullptr;      // [2]
    float* bhh = nullptr;          // [2][384]
    int K = 0, Kp = 0;
};

struct TensorEngine {
    TensorLayer enc, dec;
    float* w_head = nullptr;
    float* b_head = nullptr;
    int features = 0;
    int sm_count = 0;
    int stages = 3;                // bit 0: tensor projection, bit 1: tensor recurrence (debug A/B switch)
    // fp32 fallbacks for the A/B switch share the fp32 engine's weights (set by hb_api)
    const float* f32_enc_wcat = nullptr; const float* f32_dec_wcat = nullptr;
    const float* f32_enc_whh = nullptr;  const float* f32_dec_whh = nullptr;
};

namespace detail {

inline int pow2_scale_exponent(const float* w, size_t n) {
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(w[i]));
    if (!(mx > 0.f) || !isfinite(mx)) return 0;
    int e = (int)floorf(log2f(16384.0f / mx));      // max |w| 2^e < 2^14  (fp16 max 65504)
    if (e > 24) e = 24;
    if (e < -8) e = -8;
    return e;
}

// [rows x K] fp32 -> hi and lo fp16 core-matrix images (K-major, LBO 128, SBO (Kp/8)*128), per 128-row block
inline void pack_split_image(const float* w, int rows, int K, int Kp, float scale, __half* hi, __half* lo) {
    const int sbo = (Kp / 8) * 128;
    const size_t block_halfs = (size_t)128 * Kp;
    for (size_t i = 0; i < (size_t)rows * Kp; ++i) hi[i] = lo[i] = __float2half_rn(0.f);
    for (int r = 0; r < rows; ++r) {
        const int blk = r / 128, rr = r % 128;
        for (int k = 0; k < K; ++k) {
            const float v = w[(size_t)r * K + k] * scale;
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            const size_t off = blk * block_halfs + tc::core_offset(rr, k, 128, sbo) / 2;
            hi[off] = h;
            lo[off] = l;
        }
    }
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

inline bool pack_layer(const hb_gru_weights& g, int K, bool activations_scaled, TensorLayer* L, char* err, size_t errlen) {
    const int Kp = (K + 15) / 16 * 16;
    L->K = K;
    L->Kp = Kp;
    const size_t blk_halfs = (size_t)128 * Kp;
    std::vector<__half> wih(6 * 2 * blk_halfs);
    std::vector<float> wih_inv(6), bih(2 * G), whh_inv(2), bhh(2 * G);
    std::vector<uint32_t> whh((size_t)2 * WHH_TMEM_WORDS);
    std::vector<__half> hi, lo;
    for (int d = 0; d < 2; ++d) {
        {   // W_ih: 3 blocks of 128 rows
            const int e = pow2_scale_exponent(g.weight_ih[d], (size_t)G * K);
            hi.assign((size_t)G * Kp, __half());
            lo.assign((size_t)G * Kp, __half());
            pack_split_image(g.weight_ih[d], G, K, Kp, ldexpf(1.f, e), hi.data(), lo.data());
            for (int b = 0; b < 3; ++b) {
                std::memcpy(&wih[(size_t)(d * 3 + b) * 2 * blk_halfs], &hi[b * blk_halfs], blk_halfs * sizeof(__half));
                std::memcpy(&wih[(size_t)(d * 3 + b) * 2 * blk_halfs + blk_halfs], &lo[b * blk_halfs], blk_halfs * sizeof(__half));
                wih_inv[d * 3 + b] = ldexpf(1.f, -e) * (activations_scaled ? ACT_SCALE_INV : 1.f);
            }
        }
        {   // W_hh: TMEM image, lane = row within the gate block, column c holds k = 2c (low half), 2c+1 (high)
            const int e = pow2_scale_exponent(g.weight_hh[d], (size_t)G * H);
            const float scale = ldexpf(1.f, e);
            for (int gb = 0; gb < 3; ++gb)
                for (int r = 0; r < 128; ++r)
                    for (int c = 0; c < 64; ++c) {
                        uint32_t word[2] = {0, 0};
                        for (int half_idx = 0; half_idx < 2; ++half_idx) {
                            const float v = g.weight_hh[d][(size_t)(gb * 128 + r) * H + 2 * c + half_idx] * scale;
                            const __half hi_h = __float2half_rn(v);
                            const __half lo_h = __float2half_rn(v - __half2float(hi_h));
                            word[0] |= (uint32_t)__half_as_ushort(hi_h) << (16 * half_idx);
                            word[1] |= (uint32_t)__half_as_ushort(lo_h) << (16 * half_idx);
                        }
                        for (int term = 0; term < 2; ++term)
                            whh[(size_t)d * WHH_TMEM_WORDS + (((size_t)term * 3 + gb) * 128 + r) * 64 + c] = word[term];
                    }
            whh_inv[d] = ldexpf(1.f, -e) * ACT_SCALE_INV;
        }
        std::memcpy(&bih[d * G], g.bias_ih[d], G * sizeof(float));
        std::memcpy(&bhh[d * G], g.bias_hh[d], G * sizeof(float));
    }
    return to_device(&L->wih_img, wih, err, errlen) && to_device(&L->wih_inv, wih_inv, err, errlen) &&
           to_device(&L->bih, bih, err, errlen) && to_device(&L->whh_tmem, whh, err, errlen) &&
           to_device(&L->whh_inv, whh_inv, err, errlen) && to_device(&L->bhh, bhh, err, errlen);
}

inline void free_layer(TensorLayer* L) {
    cudaFree(L->wih_img); cudaFree(L->wih_inv); cudaFree(L->bih);
    cudaFree(L->whh_tmem); cudaFree(L->whh_inv); cudaFree(L->bhh);
}

constexpr int PROJ_NT = 64;

inline size_t projection_smem(int Kp, bool split_a) {
    return (size_t)2 * 128 * Kp * 2 + (size_t)(split_a ? 2 : 1) * PROJ_NT * Kp * 2 + 64;
}
// the kernel allocates all 512 TMEM columns, so force one CTA per SM through the smem request
template <int N>
constexpr size_t recurrence_smem() { return (size_t)120 * 1024; }

}  // namespace detail

inline void tensor_engine_destroy(TensorEngine* e) {
    if (!e) return;
    detail::free_layer(&e->enc);
    detail::free_layer(&e->dec);
    cudaFree(e->w_head);
    cudaFree(e->b_head);
    delete e;
}

inline TensorEngine* tensor_engine_create(const hb_weights* w, int features, int sm_count, char* err, size_t errlen) {
    TensorEngine* e = new TensorEngine();
    e->features = features;
    e->sm_count = sm_count;
    bool ok = detail::pack_layer(w->encoder, features, /*activations_scaled=*/false, &e->enc, err, errlen) &&
              detail::pack_layer(w->decoder, 2 * H, /*activations_scaled=*/true, &e->dec, err, errlen);
    if (ok) {
        std::vector<float> wh((size_t)NCLS * 2 * H), bh(NCLS);
        std::memcpy(wh.data(), w->base_weight, (size_t)NBASE * 2 * H * sizeof(float));
        std::memcpy(wh.data() + (size_t)NBASE * 2 * H, w->rle_weight, (size_t)NRLE * 2 * H * sizeof(float));
        std::memcpy(bh.data(), w->base_bias, NBASE * sizeof(float));
        std::memcpy(bh.data() + NBASE, w->rle_bias, NRLE * sizeof(float));
        ok = detail::to_device(&e->w_head, wh, err, errlen) && detail::to_device(&e->b_head, bh, err, errlen);
    }
    if (ok) {
        cudaError_t ce = cudaSuccess;
        auto set = [&](const void* fn, size_t bytes) {
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        };
        set((const void*)tc_projection_kernel<uint8_t, detail::PROJ_NT>, detail::projection_smem(e->enc.Kp, false));
        set((const void*)tc_projection_kernel<float, detail::PROJ_NT>, detail::projection_smem(e->dec.Kp, true));
        set((const void*)tc_recurrence_kernel<16>, detail::recurrence_smem<16>());
        set((const void*)tc_recurrence_kernel<32>, detail::recurrence_smem<32>());
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

inline size_t tensor_engine_workspace_bytes(const TensorEngine*, int64_t B, int T, int W) { return carve(nullptr, B, T, W).bytes; }

// windows per recurrence CTA: smallest tile that still gives every SM at most one CTA's worth of work
inline int pick_windows_per_cta(int64_t B, int sm_count) {
    const int64_t dir_windows = 2 * B;
    if (dir_windows <= (int64_t)16 * sm_count) return 16;
    return 32;
}

template <typename TA>
inline void launch_tc_projection(const TensorEngine* e, const TensorLayer& L, const TA* a, int64_t a_batch_stride,
                                 int64_t a_row_stride, int W, int64_t M, float* gi, cudaStream_t s) {
    constexpr int NT = detail::PROJ_NT;
    const int64_t tiles = (M + NT - 1) / NT;
    const int workers = (int)std::min<int64_t>(tiles, std::max(1, e->sm_count / 6 * 2));
    dim3 grid((unsigned)workers, 6);
    tc_projection_kernel<TA, NT><<<grid, 256, detail::projection_smem(L.Kp, sizeof(TA) == 4), s>>>(
        a, a_batch_stride, a_row_stride, W, M, L.K, L.Kp, L.wih_img, L.bih, L.wih_inv, gi);
}

inline void launch_tc_recurrence(const TensorEngine* e, const TensorLayer& L, const float* gi, const float* h_in,
                                 float* h_out, float* y, int64_t B, int W, cudaStream_t s) {
    const int n = pick_windows_per_cta(B, e->sm_count);
    dim3 grid((unsigned)((B + n - 1) / n), 2);
    if (n == 16)
        tc_recurrence_kernel<16><<<grid, REC_TC_THREADS, detail::recurrence_smem<16>(), s>>>(gi, L.whh_tmem, L.bhh, L.whh_inv, h_in, h_out, y, nullptr, nullptr, B, W);
    else
        tc_recurrence_kernel<32><<<grid, REC_TC_THREADS, detail::recurrence_smem<32>(), s>>>(gi, L.whh_tmem, L.bhh, L.whh_inv, h_in, h_out, y, nullptr, nullptr, B, W);
}

// Returns the number of kernel launches issued, or a negative hb_status (message in err).
inline int tensor_engine_predict(TensorEngine* e, const uint8_t* images, int64_t B, int T, int W, int J,
                                 uint8_t* base_labels, uint8_t* rle_labels, float* base_prob, float* rle_prob,
                                 void* workspace, cudaStream_t s, char* err, size_t errlen) {
    Workspace ws = carve(workspace, B, T, W);
    float* p_base = base_prob ? base_prob : ws.p_base;
    float* p_rle = rle_prob ? rle_prob : ws.p_rle;
    const int F = e->features;
    int launches = 0;
    cudaMemsetAsync(p_base, 0, (size_t)B * T * NBASE * sizeof(float), s);
    cudaMemsetAsync(p_rle, 0, (size_t)B * T * NRLE * sizeof(float), s);
    const float* hid = nullptr;
    float* hid_bufs[2] = {ws.hid_a, ws.hid_b};
    int flip = 0;
    const int64_t rows = B * W;
    const bool tc_proj = e->stages & 1, tc_rec = e->stages & 2;
    for (int i = 0; i + W <= T; i += J) {
        float* enc_h = hid_bufs[flip];
        float* dec_h = hid_bufs[flip ^ 1];
        // encoder
        if (tc_proj) launch_tc_projection<uint8_t>(e, e->enc, images + (int64_t)i * F, (int64_t)T * F, F, W, rows, ws.gi, s);
        else {
            dim3 gp((unsigned)((rows + 63) / 64), 2 * G / 64);
            input_projection_kernel<uint8_t><<<gp, 256, 0, s>>>(images + (int64_t)i * F, (int64_t)T * F, F, W, rows, F, e->f32_enc_wcat, e->enc.bih, ws.gi);
        }
        if (tc_rec) launch_tc_recurrence(e, e->enc, ws.gi, hid, enc_h, ws.y1, B, W, s);
        else {
            dim3 gr((unsigned)((B + REC_WINDOWS - 1) / REC_WINDOWS), 2);
            gru_recurrence_kernel<<<gr, REC_THREADS, 0, s>>>(ws.gi, e->f32_enc_whh, e->enc.bhh, hid, enc_h, ws.y1, B, W);
        }
        // decoder
        if (tc_proj) launch_tc_projection<float>(e, e->dec, ws.y1, (int64_t)W * 2 * H, 2 * H, W, rows, ws.gi, s);
        else {
            dim3 gp((unsigned)((rows + 63) / 64), 2 * G / 64);
            input_projection_kernel<float><<<gp, 256, 0, s>>>(ws.y1, (int64_t)W * 2 * H, 2 * H, W, rows, 2 * H, e->f32_dec_wcat, e->dec.bih, ws.gi);
        }
        if (tc_rec) launch_tc_recurrence(e, e->dec, ws.gi, enc_h, dec_h, ws.y2, B, W, s);
        else {
            dim3 gr((unsigned)((B + REC_WINDOWS - 1) / REC_WINDOWS), 2);
            gru_recurrence_kernel<<<gr, REC_THREADS, 0, s>>>(ws.gi, e->f32_dec_whh, e->dec.bhh, enc_h, dec_h, ws.y2, B, W);
        }
        const int blocks = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)e->sm_count * 8);
        heads_kernel<<<blocks, 256, 0, s>>>(ws.y2, e->w_head, e->b_head, rows, W, T, i, p_base, p_rle, nullptr, nullptr, 0);
        launches += 5;
        hid = dec_h;
        flip ^= 1;
    }
    const int64_t positions = B * T;
    argmax_kernel<<<(unsigned)((positions + 255) / 256), 256, 0, s>>>(p_base, p_rle, positions, base_labels, rle_labels);
    launches += 1;
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
        snprintf(err, errlen, "tensor engine launch failed: %s", cudaGetErrorString(ce));
        return HB_ERR_CUDA;
    }
    return launches;
}

}  // namespace hb
