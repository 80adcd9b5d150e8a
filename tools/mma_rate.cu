// Microbenchmark: cycles per tcgen05.mma (kind::f16, fp32 accumulate, cta_group::1, M = 128, K = 16) on one SM as a
// function of N, of where A lives (TMEM / shared memory), of the shared-memory layout of B (no-swizzle core matrices /
// 128-byte swizzle) and of whether consecutive MMAs accumulate into the same TMEM tile.  Backs the statement in
// DESIGN.md 5.3 that these MMAs cost ~N cycles each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_rate tools/mma_rate.cu && tools/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../helen_b200/csrc/tc_ptx.cuh"

using namespace hb;

// BG: what the other three warps do meanwhile: 0 nothing, 1 stream 16-byte shared-memory loads + stores (other region),
// 2 stream tcgen05.ld of other TMEM columns, 3 bulk copies global -> shared (TMA writes into shared memory)
// COMMIT: 0 = one tcgen05.commit after all MMAs; k > 0 = a commit (to a scratch mbarrier nobody waits on) after every k MMAs
template <int N, bool A_TMEM, bool SWIZZLE, int NACC, int BG = 0, int COMMIT = 0>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps, const uint8_t* gsrc = nullptr)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = tc::align_smem_1024(smem_raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;   // zeros: finite operands
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    tc::fence_proxy_async_smem();
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    __shared__ volatile int stop;
    __shared__ uint64_t bg_bar, scratch_bar;
    if (threadIdx.x == 0) { stop = 0; tc::mbar_init(&bg_bar, 1); tc::mbar_init(&scratch_bar, 1000000); tc::mbar_fence_init(); }
    __syncthreads();
    if (warp != 0 && BG != 0) {
        uint8_t* region = smem + 100 * 1024 + (warp - 1) * 16 * 1024;      // 48 KB of background buffers after the operands
        float sink = 0.f;
        uint32_t phase = 0;
        while (!stop) {
            if (BG == 1) {
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    int4* p = reinterpret_cast<int4*>(region) + ((i * 32 + (threadIdx.x & 31)) & 1023);
                    int4 v = *p; v.x += i; *p = v;
                }
            } else if (BG == 2) {
                float v[8];
                tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + 384 + 8 * (warp - 1), v);
                tc::tmem_ld_wait();
                sink += v[0];
            } else if (BG == 3 && warp == 1) {
                if ((threadIdx.x & 31) == 0) {
                    tc::mbar_arrive_expect_tx(&bg_bar, 12288);
                    tc::bulk_g2s(region, gsrc, 12288, &bg_bar);
                }
                tc::mbar_wait(&bg_bar, phase);
                phase ^= 1;
            }
        }
        if (sink == 123.f) out[1] = 1;
    }
    if (warp == 0) {
        const uint32_t idesc = tc::idesc_f16_f32(128, N);
        // B: [N rows x 128 k]; A (smem form): [128 rows x 128 k] after it
        const uint32_t b_addr = tc::smem_u32(smem), a_addr = tc::smem_u32(smem + 64 * 1024);
        const uint64_t b_desc = SWIZZLE ? tc::smem_desc_sw128(b_addr, 2048) : tc::smem_desc(b_addr, 128, 2048);
        const uint64_t a_desc = SWIZZLE ? tc::smem_desc_sw128(a_addr, 2048) : tc::smem_desc(a_addr, 128, 2048);
        long long t0 = 0, t1 = 0;
        for (int pass = 0; pass < 2; ++pass) {                 // pass 0 warms up
            t0 = clock64();
            if (tc::elect_one()) {
                for (int r = 0; r < reps; ++r) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t d = tmem + ((r * 8 + ks) % NACC) * N;
                        const uint64_t bk = b_desc + (SWIZZLE ? tc::sw128_kstep(ks) : (uint64_t)(ks * 2 * 128 / 16));
                        if (A_TMEM) tc::mma_f16_ts(d, tmem + 256 + ks * 8, bk, idesc, 1);
                        else tc::mma_f16_ss(d, a_desc + (SWIZZLE ? tc::sw128_kstep(ks) : (uint64_t)(ks * 2 * 128 / 16)), bk, idesc, 1);
                        if (COMMIT > 0 && ((r * 8 + ks + 1) % COMMIT) == 0) tc::mma_commit(&scratch_bar);
                    }
                }
                tc::mma_commit(&bar);
            }
            __syncwarp();
            tc::mbar_wait(&bar, (uint32_t)pass);
            t1 = clock64();
        }
        if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
        stop = 1;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int N, bool A_TMEM, bool SWIZZLE, int NACC, int BG = 0, int COMMIT = 0>
void run(const char* what, long long* dev)
{
    const int reps = 64;
    static uint8_t* gsrc = nullptr;
    if (!gsrc) { cudaMalloc(&gsrc, 1 << 20); cudaMemset(gsrc, 0, 1 << 20); }
    auto k = rate_kernel<N, A_TMEM, SWIZZLE, NACC, BG, COMMIT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
    k<<<1, 128, 150 * 1024>>>(dev, reps, gsrc);
    long long cyc = 0;
    cudaError_t e = cudaMemcpy(&cyc, dev, sizeof(cyc), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%-58s N=%3d  FAILED: %s\n", what, N, cudaGetErrorString(e)); return; }
    printf("%-58s N=%3d  %7.1f cycles / MMA   (%d MMAs, %lld cycles)\n", what, N, (double)cyc / (reps * 8), reps * 8, cyc);
}

int main()
{
    long long* dev;
    cudaMalloc(&dev, 64);
    printf("# tcgen05.mma kind::f16 cta_group::1 M=128 K=16, one issuing thread, back-to-back, one SM\n");
#define ROW(N) \
    run<N, true, true, 1>("A in TMEM, B 128B-swizzle, one accumulator", dev); \
    run<N, true, true, 2>("A in TMEM, B 128B-swizzle, two accumulators alternating", dev); \
    run<N, true, false, 1>("A in TMEM, B no-swizzle, one accumulator", dev); \
    run<N, false, true, 1>("A in smem, B 128B-swizzle, one accumulator", dev);
    ROW(16) ROW(32) ROW(64) ROW(128)
    run<256, true, true, 1>("A in TMEM, B 128B-swizzle, one accumulator", dev);
    printf("# with background traffic from the other three warps of the CTA\n");
    run<16, true, true, 1, 1>("A in TMEM, + shared-memory load/store stream", dev);
    run<16, true, true, 1, 2>("A in TMEM, + tcgen05.ld stream", dev);
    run<16, true, true, 1, 3>("A in TMEM, + 12 KB bulk copies into shared memory", dev);
    run<64, true, true, 1, 1>("A in TMEM, + shared-memory load/store stream", dev);
    run<64, true, true, 1, 2>("A in TMEM, + tcgen05.ld stream", dev);
    run<64, true, true, 1, 3>("A in TMEM, + 12 KB bulk copies into shared memory", dev);
    printf("# with a tcgen05.commit after every k MMAs\n");
    run<16, true, true, 1, 0, 16>("A in TMEM, commit every 16 MMAs", dev);
    run<16, true, true, 1, 0, 8>("A in TMEM, commit every 8 MMAs", dev);
    run<16, true, true, 1, 0, 1>("A in TMEM, commit every MMA", dev);
    run<64, true, true, 1, 0, 24>("A in TMEM, commit every 24 MMAs", dev);
    run<64, true, true, 1, 0, 8>("A in TMEM, commit every 8 MMAs", dev);
    return 0;
}
