// Training step of one chunk (BASELINE configs[3]): forward of TransducerGRU with the
// activations kept, the two cross-entropy losses, and back-propagation through the heads, the decoder and
// the encoder (BPTT over the chunk's W steps, both directions) -- the per-chunk body of the reference's
// training loop, helen/modules/python/models/train.py:174-206 (forward :189, losses :192-198, backward :201),
// with the losses of train.py:121-126 (CrossEntropyLoss, and CrossEntropyLoss(weight=CLASS_WEIGHTS) for the
// run lengths, Options.py:29).  The optimizer step (:202) stays with the caller.
//
// Everything here is fp32 on the FMA pipes, like the fp32 inference engine whose projection / heads kernels it
// reuses: gradients have to match autograd to 1e-4 and this path is not the north-star path.  The sequential
// part of the backward pass is one persistent kernel per layer that mirrors the forward recurrence kernel with
// W_hh TRANSPOSED in registers (dh_{t-1} = z * dh_t + W_hh^T . dgh_t); every weight gradient is a plain GEMM
// over the step-wise pre-activation gradients it leaves behind.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fp32_kernels.cuh"

namespace hb {
namespace train {

// ---------------------------------------------------------------------------------------------
// C[m, n] (+)= sum_k A(m, k) * B(k, n) (+ bias[n]) with arbitrary element strides, so the same kernel
// serves  X . W^T  (forward projections),  dG . W  (input gradients)  and  dG^T . X  (weight gradients).
// gridDim.z > 1 splits K; the partial sums are added atomically into a zeroed (or accumulated) C.
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
    const float* a; int64_t a_m, a_k;      // strides of A (elements)
    const float* b; int64_t b_k, b_n;      // strides of B
    float* c; int64_t c_m, c_n;
    const float* bias;                     // [N] or nullptr (added by the k-split 0 only)
    int64_t M, N, K;
    int accumulate;                        // add into C instead of overwriting it (always the case when K is split)
};

__global__ void __launch_bounds__(256)
gemm_kernel(const GemmArgs g)
{
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float as[BK][BM + 4];
    __shared__ float bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
    const int64_t k_per = ((g.K + gridDim.z - 1) / gridDim.z + BK - 1) / BK * BK;
    const int64_t k_begin = (int64_t)blockIdx.z * k_per, k_end = min(g.K, k_begin + k_per);
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4] = {};
    for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            // A tile: consecutive threads along the unit-stride dimension of A
            int r, kk;
            if (g.a_k == 1) { kk = e & 15; r = e >> 4; } else { r = e & 63; kk = e >> 6; }
            float va = 0.f;
            if (m0 + r < g.M && k0 + kk < k_end) va = g.a[(m0 + r) * g.a_m + (k0 + kk) * g.a_k];
            as[kk][r] = va;
            int c, kb;
            if (g.b_k == 1) { kb = e & 15; c = e >> 4; } else { c = e & 63; kb = e >> 6; }
            float vb = 0.f;
            if (n0 + c < g.N && k0 + kb < k_end) vb = g.b[(k0 + kb) * g.b_k + (n0 + c) * g.b_n];
            bs[kb][c] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = as[kk][ty * 4 + i]; bv[i] = bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias != nullptr && blockIdx.z == 0) v += g.bias[n];
            float* dst = g.c + m * g.c_m + n * g.c_n;
            if (gridDim.z > 1) atomicAdd(dst, v); else if (g.accumulate) *dst += v; else *dst = v;
        }
    }
}

// out[n] += sum_m a[m, n]   (bias gradients); out must be zeroed
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ a, int64_t M, int N, int64_t lda, float* __restrict__ out)
{
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int slice = threadIdx.x >> 5;                       // 8 row slices per CTA
    const int64_t rows_per = (M + gridDim.y * 8 - 1) / (gridDim.y * 8);
    const int64_t m0 = ((int64_t)blockIdx.y * 8 + slice) * rows_per, m1 = min(M, m0 + rows_per);
    if (n >= N) return;
    float s = 0.f;
    for (int64_t m = m0; m < m1; ++m) s += a[m * lda + n];
    atomicAdd(out + n, s);
}

// ---------------------------------------------------------------------------------------------
// Forward recurrence with the gates kept: same work split as gru_recurrence_kernel (CTA = 4 windows x one
// direction, W_hh rows in registers), plus saved[row][dir][r, z, n, W_hn.h + b_hn][128].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(REC_THREADS, 1)
gru_forward_save_kernel(const float* __restrict__ gi,       // [B*W, 768]  (b_ih already added)
                        const float* __restrict__ w_hh0, const float* __restrict__ w_hh1,   // [384, 128] per direction
                        const float* __restrict__ b_hh0, const float* __restrict__ b_hh1,   // [384]
                        const float* __restrict__ h_in,     // [B, 2, 128] or nullptr (zeros)
                        float* __restrict__ h_out,          // [B, 2, 128]
                        float* __restrict__ y,              // [B*W, 256]
                        float* __restrict__ saved,          // [B*W, 2, 4, 128]
                        int64_t B, int W)
{
    __shared__ __align__(16) float hs[2][REC_WINDOWS][4 * HPAD];
    const int tid = threadIdx.x;
    const int j = tid >> 2, q = tid & 3;
    const int dir = blockIdx.y;
    const int64_t b0 = (int64_t)blockIdx.x * REC_WINDOWS;
    const int64_t my_b = b0 + q;
    const bool live = my_b < B;
    const float* wd = dir ? w_hh1 : w_hh0;
    const float* bd = dir ? b_hh1 : b_hh0;

    float w[3][32];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
            float4 v = *reinterpret_cast<const float4*>(wd + (int64_t)(g * H + j) * H + q * 32 + k);
            w[g][k] = v.x; w[g][k + 1] = v.y; w[g][k + 2] = v.z; w[g][k + 3] = v.w;
        }
    const float bhr = bd[j], bhz = bd[H + j], bhn = bd[2 * H + j];

    float h_own = 0.f;
    if (live && h_in) h_own = h_in[(my_b * 2 + dir) * H + j];
    hs[0][q][(j >> 5) * HPAD + (j & 31)] = h_own;
    __syncthreads();

    const int64_t row0 = (live ? my_b : b0) * W;
    const float* gi_dir = gi + dir * G + j;
    int t = dir ? W - 1 : 0;
    const int dt = dir ? -1 : 1;
    for (int s = 0; s < W; ++s, t += dt) {
        const int cur = s & 1;
        const float* p = gi_dir + (row0 + t) * (2 * G);
        const float gir = p[0], giz = p[H], gin = p[2 * H];
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
        const float ghn = sn + bhn;
        const float n = tanhf(gin + r * ghn);
        h_own = (1.0f - z) * n + z * h_own;
        hs[cur ^ 1][q][(j >> 5) * HPAD + (j & 31)] = h_own;
        if (live) {
            y[(row0 + t) * (2 * H) + dir * H + j] = h_own;
            float* sv = saved + (((row0 + t) * 2 + dir) * 4) * H + j;
            sv[0] = r; sv[H] = z; sv[2 * H] = n; sv[3 * H] = ghn;
        }
        __syncthreads();
    }
    if (live) h_out[(my_b * 2 + dir) * H + j] = h_own;
}

// ---------------------------------------------------------------------------------------------
// Backward recurrence of one layer.  CTA = 4 windows x one direction, walking the forward order backwards.
// Thread (k = tid >> 2, q = tid & 3) keeps COLUMN k of W_hh for gate rows [32q, 32q+32) of r, z, n
// (96 registers): dh_prev[k] = z[k] dh[k] + sum_rows W_hh[row, k] dgh[row].  Lane q finishes window q.
// Leaves behind, per (row = b*W + t, dir): dgi [r, z, n] (gradient of the input-side pre-activations, also the
// b_ih gradient summand), dgh (hidden side: the n part is dn_pre * r), and hprev (the state the step started from).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(REC_THREADS, 1)
gru_backward_kernel(const float* __restrict__ dy,          // [B*W, 256] gradient wrt the layer output
                    const float* __restrict__ dh_n,        // [B, 2, 128] gradient wrt the final state, or nullptr
                    const float* __restrict__ saved,       // [B*W, 2, 4, 128]
                    const float* __restrict__ y,           // [B*W, 256] layer output (h_t)
                    const float* __restrict__ h_in,        // [B, 2, 128] or nullptr
                    const float* __restrict__ w_hh0, const float* __restrict__ w_hh1,
                    float* __restrict__ dgi,               // [B*W, 768]
                    float* __restrict__ dgh,               // [B*W, 768]
                    float* __restrict__ hprev,             // [B*W, 256]
                    float* __restrict__ dh_in,             // [B, 2, 128] gradient wrt the initial state, or nullptr
                    int64_t B, int W)
{
    __shared__ __align__(16) float gs[REC_WINDOWS][3][4 * HPAD];     // dgh of the step: [window][gate][row]
    const int tid = threadIdx.x;
    const int k = tid >> 2, q = tid & 3;
    const int dir = blockIdx.y;
    const int64_t b0 = (int64_t)blockIdx.x * REC_WINDOWS;
    const int64_t my_b = b0 + q;
    const bool live = my_b < B;
    const float* wd = dir ? w_hh1 : w_hh0;

    float wt[3][32];                                          // wt[g][jj] = W_hh[g*128 + 32q + jj][k]
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) wt[g][jj] = wd[(int64_t)(g * H + q * 32 + jj) * H + k];

    float dh = (live && dh_n) ? dh_n[(my_b * 2 + dir) * H + k] : 0.f;
    const int64_t row0 = (live ? my_b : b0) * W;
    // forward visited t_first, t_first + dt, ...; walk it backwards
    int t = dir ? 0 : W - 1;
    const int dt = dir ? -1 : 1;                              // forward increment
    for (int s = W - 1; s >= 0; --s, t -= dt) {
        float dr_pre = 0.f, dz_pre = 0.f, dn_pre = 0.f, dghn = 0.f, z = 0.f;
        if (live) {
            const int64_t row = row0 + t;
            const float* sv = saved + ((row * 2 + dir) * 4) * H + k;
            const float r = sv[0], n = sv[2 * H], ghn = sv[3 * H];
            z = sv[H];
            const float hp = s > 0 ? y[(row - dt) * (2 * H) + dir * H + k] : (h_in ? h_in[(my_b * 2 + dir) * H + k] : 0.f);
            dh += dy[row * (2 * H) + dir * H + k];
            const float dn = dh * (1.0f - z);
            const float dz = dh * (hp - n);
            dn_pre = dn * (1.0f - n * n);
            dghn = dn_pre * r;
            dr_pre = dn_pre * ghn * r * (1.0f - r);
            dz_pre = dz * z * (1.0f - z);
            float* o = dgi + row * (2 * G) + dir * G + k;
            o[0] = dr_pre; o[H] = dz_pre; o[2 * H] = dn_pre;
            float* oh = dgh + row * (2 * G) + dir * G + k;
            oh[0] = dr_pre; oh[H] = dz_pre; oh[2 * H] = dghn;
            hprev[row * (2 * H) + dir * H + k] = hp;
        }
        const int slot = (k >> 5) * HPAD + (k & 31);
        gs[q][0][slot] = dr_pre; gs[q][1][slot] = dz_pre; gs[q][2][slot] = dghn;
        __syncthreads();
        // W_hh^T . dgh for the four windows: partial sums over this lane's 32 rows of each gate
        float part[REC_WINDOWS];
#pragma unroll
        for (int wi = 0; wi < REC_WINDOWS; ++wi) {
            float a = 0.f;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const float4* gp = reinterpret_cast<const float4*>(&gs[wi][g][q * HPAD]);
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const float4 x = gp[v];
                    a = fmaf(wt[g][4 * v], x.x, a); a = fmaf(wt[g][4 * v + 1], x.y, a);
                    a = fmaf(wt[g][4 * v + 2], x.z, a); a = fmaf(wt[g][4 * v + 3], x.w, a);
                }
            }
            part[wi] = a;
        }
        float rec = 0.f;
#pragma unroll
        for (int wi = 0; wi < REC_WINDOWS; ++wi) {
            float v = part[wi];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (wi == q) rec = v;
        }
        dh = dh * z + rec;                                    // gradient wrt the state the step started from
        __syncthreads();
    }
    if (live && dh_in) dh_in[(my_b * 2 + dir) * H + k] = dh;
}

// ---------------------------------------------------------------------------------------------
// Cross entropy of both heads, mean reduction as torch.nn.CrossEntropyLoss: loss = sum_i w[y_i] nll_i / sum_i w[y_i]
// (w = 1 for the base head).  Pass 1 adds the numerators and denominators into sums[4] = {num_base, den_base,
// num_rle, den_rle}; pass 2 writes dlogits [rows, 16] = w[y] (softmax - onehot) / den.
// ---------------------------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(256)
ce_kernel(const float* __restrict__ logit_base, const float* __restrict__ logit_rle,
          const int64_t* __restrict__ label_base, const int64_t* __restrict__ label_rle,
          const float* __restrict__ rle_weight,      // [11]
          int64_t rows, float* __restrict__ sums, float* __restrict__ dlogits)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float nb = 0.f, db = 0.f, nr = 0.f, dr = 0.f;
    if (i < rows) {
        float lb[NBASE], lr[NRLE];
        float mb = -INFINITY, mr = -INFINITY;
#pragma unroll
        for (int c = 0; c < NBASE; ++c) { lb[c] = logit_base[i * NBASE + c]; mb = fmaxf(mb, lb[c]); }
#pragma unroll
        for (int c = 0; c < NRLE; ++c) { lr[c] = logit_rle[i * NRLE + c]; mr = fmaxf(mr, lr[c]); }
        float sb = 0.f, sr = 0.f;
#pragma unroll
        for (int c = 0; c < NBASE; ++c) { lb[c] = expf(lb[c] - mb); sb += lb[c]; }
#pragma unroll
        for (int c = 0; c < NRLE; ++c) { lr[c] = expf(lr[c] - mr); sr += lr[c]; }
        // A label outside its class range poisons the loss with NaN instead of reading past rle_weight
        // (torch's CrossEntropyLoss raises a device assert there; ChunkTrainer.step turns the NaN into a ValueError).
        const int64_t yb64 = label_base[i], yr64 = label_rle[i];
        const bool in_range = yb64 >= 0 && yb64 < NBASE && yr64 >= 0 && yr64 < NRLE;
        const int yb = in_range ? (int)yb64 : 0, yr = in_range ? (int)yr64 : 0;
        const float wr = in_range ? rle_weight[yr] : __int_as_float(0x7fc00000);
        if (PASS == 1) {
            float pb = 0.f, pr = 0.f;
#pragma unroll
            for (int c = 0; c < NBASE; ++c) if (c == yb) pb = lb[c];
#pragma unroll
            for (int c = 0; c < NRLE; ++c) if (c == yr) pr = lr[c];
            nb = -logf(pb / sb); db = 1.f;
            nr = -wr * logf(pr / sr); dr = wr;
        } else {
            const float inv_b = 1.0f / sums[1], inv_r = wr / sums[3];
#pragma unroll
            for (int c = 0; c < NBASE; ++c) dlogits[i * NCLS + c] = (lb[c] / sb - (c == yb ? 1.f : 0.f)) * inv_b;
#pragma unroll
            for (int c = 0; c < NRLE; ++c) dlogits[i * NCLS + NBASE + c] = (lr[c] / sr - (c == yr ? 1.f : 0.f)) * inv_r;
        }
    }
    if (PASS == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nb += __shfl_xor_sync(0xffffffffu, nb, o); db += __shfl_xor_sync(0xffffffffu, db, o);
            nr += __shfl_xor_sync(0xffffffffu, nr, o); dr += __shfl_xor_sync(0xffffffffu, dr, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(sums + 0, nb); atomicAdd(sums + 1, db); atomicAdd(sums + 2, nr); atomicAdd(sums + 3, dr);
        }
    }
}

// loss[0] = total, loss[1] = base, loss[2] = rle   (train.py:192-198)
__global__ void loss_finish_kernel(const float* __restrict__ sums, float* __restrict__ loss)
{
    const float lb = sums[0] / sums[1], lr = sums[2] / sums[3];
    loss[0] = lb + lr; loss[1] = lb; loss[2] = lr;
}

}  // namespace train
}  // namespace hb
