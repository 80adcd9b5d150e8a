// fp32 engine: every contraction on the FMA pipes, bit-for-bit fp32 arithmetic.
//
// Kernels (one chunk of the reference loop, predict_gpu.py:114-149, per call):
//   input_projection_kernel  gi[m, 0:768] = A[m, 0:K] . Wcat^T + b_ih      (both directions)
//   gru_recurrence_kernel    100 dependent GRU steps, W_hh resident in registers
//   heads_kernel             logits = y2 . Whead^T + b ; softmax ; P[:, i:i+W] += .
//   argmax_kernel            first-index argmax of the accumulated sums
//
// GRU cell = torch.nn.GRU (TransducerModel.py:43-53,70-72): rows (r, z, n), b_hn inside
// the r* product.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

constexpr int H = 128;          // Options.py:28
constexpr int G = 3 * H;        // gate rows per direction
constexpr int NBASE = 5;        // Options.py:20
constexpr int NRLE = 11;        // Options.py:21
constexpr int NCLS = 16;        // 5 + 11 head rows, evaluated together

__device__ __forceinline__ float sigmoidf_precise(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// gi[m, n] = sum_k A[m, k] * Wcat[n, k] + bias[n],   m = b * W + t,  n in [0, 768)
// A row m lives at a_base + b * a_batch_stride + t * a_row_stride (elements), so the same
// kernel reads a W-column slice of the uint8 pileup image (predict_gpu.py:122, cast :97) or a
// dense fp32 [B*W, K] matrix (encoder output).
// ---------------------------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(256)
input_projection_kernel(const TA* __restrict__ a, int64_t a_batch_stride, int64_t a_row_stride,
                        int rows_per_window, int64_t M, int K,
                        const float* __restrict__ wcat,   // [768, K]
                        const float* __restrict__ bias,   // [768]
                        float* __restrict__ gi)           // [M, 768]
{
    constexpr int BM = 64, BN = 64, BK = 16, N = 2 * G;
    __shared__ float as[BK][BM + 4];
    __shared__ float ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tx = tid & 15, ty = tid >> 4;           // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4] = {};

    for (int k0 = 0; k0 < K; k0 += BK) {
        // 64 x 16 tile of A and of W: 1024 elements each, 4 per thread
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + i * 256;
            int r = e >> 4, kk = e & 15;
            int64_t m = m0 + r;
            float va = 0.f, vw = 0.f;
            if (k0 + kk < K) {
                if (m < M) {
                    int64_t b = m / rows_per_window, t = m - b * rows_per_window;
                    va = (float)a[b * a_batch_stride + t * a_row_stride + k0 + kk];
                }
                vw = wcat[(int64_t)(n0 + r) * K + k0 + kk];
            }
            as[kk][r] = va;
            ws[kk][r] = vw;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = as[kk][ty * 4 + i]; wv[i] = ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float4 o;
        o.x = acc[i][0] + bias[n0 + tx * 4 + 0];
        o.y = acc[i][1] + bias[n0 + tx * 4 + 1];
        o.z = acc[i][2] + bias[n0 + tx * 4 + 2];
        o.w = acc[i][3] + bias[n0 + tx * 4 + 3];
        *reinterpret_cast<float4*>(gi + m * N + n0 + tx * 4) = o;
    }
}

// ---------------------------------------------------------------------------------------------
// One CTA = 4 windows x one direction, all W steps of one GRU layer.
// Thread (j = tid >> 2, q = tid & 3) keeps rows r/z/n of hidden unit j, columns [32q, 32q+32)
// of W_hh in 96 registers for the whole launch (persistent-RNN).  Each step: partial dot
// products against h (broadcast from shared memory), 2-level xor-shuffle reduction over q,
// then lane q finishes the gates of window q.
// ---------------------------------------------------------------------------------------------
constexpr int REC_WINDOWS = 4;
constexpr int REC_THREADS = 512;
constexpr int HPAD = 36;                          // 32 + 4: the four quarters hit different banks

__global__ void __launch_bounds__(REC_THREADS, 1)
gru_recurrence_kernel(const float* __restrict__ gi,       // [B*W, 768]  (b_ih already added)
                      const float* __restrict__ w_hh,     // [2][384][128]
                      const float* __restrict__ b_hh,     // [2][384]
                      const float* __restrict__ h_in,     // [B, 2, 128] or nullptr (zeros)
                      float* __restrict__ h_out,          // [B, 2, 128]
                      float* __restrict__ y,              // [B*W, 256]
                      int64_t B, int W)
{
    __shared__ __align__(16) float hs[2][REC_WINDOWS][4 * HPAD];
    const int tid = threadIdx.x;
    const int j = tid >> 2, q = tid & 3;
    const int dir = blockIdx.y;
    const int64_t b0 = (int64_t)blockIdx.x * REC_WINDOWS;
    const int64_t my_b = b0 + q;                 // the window whose gates this lane finishes
    const bool live = my_b < B;

    float w[3][32];
    {
        const float* wd = w_hh + (int64_t)dir * G * H;
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
                float4 v = *reinterpret_cast<const float4*>(wd + (int64_t)(g * H + j) * H + q * 32 + k);
                w[g][k] = v.x; w[g][k + 1] = v.y; w[g][k + 2] = v.z; w[g][k + 3] = v.w;
            }
    }
    const float bhr = b_hh[dir * G + j], bhz = b_hh[dir * G + H + j], bhn = b_hh[dir * G + 2 * H + j];

    float h_own = 0.f;
    if (live && h_in) h_own = h_in[(my_b * 2 + dir) * H + j];
    hs[0][q][(j >> 5) * HPAD + (j & 31)] = h_own;
    __syncthreads();

    const int64_t row0 = (live ? my_b : b0) * W;
    const float* gi_dir = gi + dir * G + j;
    int t = dir ? W - 1 : 0;
    const int dt = dir ? -1 : 1;
    float gir = 0.f, giz = 0.f, gin = 0.f;
    if (W > 0) {
        const float* p = gi_dir + (row0 + t) * (2 * G);
        gir = p[0]; giz = p[H]; gin = p[2 * H];
    }

    for (int s = 0; s < W; ++s, t += dt) {
        const int cur = s & 1;
        // prefetch next step's input projection while the matvec runs
        float ngir = 0.f, ngiz = 0.f, ngin = 0.f;
        if (s + 1 < W) {
            const float* p = gi_dir + (row0 + t + dt) * (2 * G);
            ngir = p[0]; ngiz = p[H]; ngin = p[2 * H];
        }
        float acc[REC_WINDOWS][3];
#pragma unroll
        for (int wi = 0; wi < REC_WINDOWS; ++wi) {
            float ar = 0.f, az = 0.f, an = 0.f;
            const float4* hp = reinterpret_cast<const float4*>(&hs[cur][wi][q * HPAD]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float4 hv = hp[k];
                ar = fmaf(w[0][4 * k], hv.x, ar); az = fmaf(w[1][4 * k], hv.x, az); an = fmaf(w[2][4 * k], hv.x, an);
                ar = fmaf(w[0][4 * k + 1], hv.y, ar); az = fmaf(w[1][4 * k + 1], hv.y, az); an = fmaf(w[2][4 * k + 1], hv.y, an);
                ar = fmaf(w[0][4 * k + 2], hv.z, ar); az = fmaf(w[1][4 * k + 2], hv.z, az); an = fmaf(w[2][4 * k + 2], hv.z, an);
                ar = fmaf(w[0][4 * k + 3], hv.w, ar); az = fmaf(w[1][4 * k + 3], hv.w, az); an = fmaf(w[2][4 * k + 3], hv.w, an);
            }
            acc[wi][0] = ar; acc[wi][1] = az; acc[wi][2] = an;
        }
        // reduce over the 4 quarter-lanes; lane q keeps window q's sums
        float sr = 0.f, sz = 0.f, sn = 0.f;
#pragma unroll
        for (int wi = 0; wi < REC_WINDOWS; ++wi) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float v = acc[wi][g];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (wi == q) { if (g == 0) sr = v; else if (g == 1) sz = v; else sn = v; }
            }
        }
        const float r = sigmoidf_precise(gir + sr + bhr);
        const float z = sigmoidf_precise(giz + sz + bhz);
        const float n = tanhf(gin + r * (sn + bhn));
        h_own = (1.0f - z) * n + z * h_own;
        hs[cur ^ 1][q][(j >> 5) * HPAD + (j & 31)] = h_own;
        if (live) y[(row0 + t) * (2 * H) + dir * H + j] = h_own;
        gir = ngir; giz = ngiz; gin = ngin;
        __syncthreads();
    }
    if (live) h_out[(my_b * 2 + dir) * H + j] = h_own;
}

// ---------------------------------------------------------------------------------------------
// Heads.  One warp per (window, column): 16 dot products of length 256, warp reduction,
// then either raw logits out (forward_chunk) or softmax + accumulate into the running sums
// (predict_gpu.py:137-149 without materialising the zero padding).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
heads_kernel(const float* __restrict__ y2,          // [B*W, 256]
             const float* __restrict__ w_head,      // [16, 256]  rows 0..4 base, 5..15 rle
             const float* __restrict__ b_head,      // [16]
             int64_t rows, int W, int T, int col0,
             float* __restrict__ p_base,            // [B, T, 5]  accumulate (mode 0)
             float* __restrict__ p_rle,             // [B, T, 11]
             float* __restrict__ logit_base,        // [B*W, 5]   raw (mode 1)
             float* __restrict__ logit_rle,         // [B*W, 11]
             int mode)
{
    __shared__ float wsm[NCLS][2 * H + 1];
    for (int e = threadIdx.x; e < NCLS * 2 * H; e += blockDim.x) wsm[e / (2 * H)][e % (2 * H)] = w_head[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp_global; row < rows; row += n_warps) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = y2[row * (2 * H) + lane + 32 * i];
        float logit = 0.f;                       // lane c < 16 ends up holding class c
#pragma unroll
        for (int c = 0; c < NCLS; ++c) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s = fmaf(v[i], wsm[c][lane + 32 * i], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == c) logit = s + b_head[c];
        }
        if (mode == 1) {
            if (lane < NBASE) logit_base[row * NBASE + lane] = logit;
            else if (lane < NCLS) logit_rle[row * NRLE + (lane - NBASE)] = logit;
            continue;
        }
        // softmax over lanes [0,5) and [5,16) separately
        const bool is_base = lane < NBASE, is_cls = lane < NCLS;
        float mb = is_base ? logit : -INFINITY, mr = (is_cls && !is_base) ? logit : -INFINITY;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
            mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
        }
        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 16));
        mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, 16));
        float e = is_cls ? expf(logit - (is_base ? mb : mr)) : 0.f;
        float sb = is_base ? e : 0.f, sr = (is_cls && !is_base) ? e : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
        }
        const int64_t b = row / W;
        const int t = col0 + (int)(row - b * W);
        if (is_base) p_base[(b * T + t) * NBASE + lane] += e / sb;
        else if (is_cls) p_rle[(b * T + t) * NRLE + (lane - NBASE)] += e / sr;
    }
}

// torch.max(x, 2) (predict_gpu.py:155-156): first maximal index.
__global__ void argmax_kernel(const float* __restrict__ p_base, const float* __restrict__ p_rle,
                              int64_t positions, uint8_t* __restrict__ base_label,
                              uint8_t* __restrict__ rle_label)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= positions) return;
    const float* pb = p_base + i * NBASE;
    const float* pr = p_rle + i * NRLE;
    int ib = 0, ir = 0;
    float vb = pb[0], vr = pr[0];
#pragma unroll
    for (int c = 1; c < NBASE; ++c) if (pb[c] > vb) { vb = pb[c]; ib = c; }
#pragma unroll
    for (int c = 1; c < NRLE; ++c) if (pr[c] > vr) { vr = pr[c]; ir = c; }
    base_label[i] = (uint8_t)ib;
    rle_label[i] = (uint8_t)ir;
}

}  // namespace hb
