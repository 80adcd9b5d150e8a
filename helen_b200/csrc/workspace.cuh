// Device scratch layout shared by both engines (caller-owned memory, see hb_workspace_bytes).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>

#include "fp32_kernels.cuh"

namespace hb {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct Workspace {
    float* gi;       // [B*W, 768]  input projections of the layer being run (both directions)
    float* y1;       // [B*W, 256]  encoder output of the current chunk
    float* y2;       // [B*W, 256]  decoder output of the current chunk
    float* hid_a;    // [B, 2, 128] hidden carry (ping)
    float* hid_b;    // [B, 2, 128] hidden carry (pong)
    float* p_base;   // [B, T, 5]   accumulated softmax sums when the caller does not ask for them
    float* p_rle;    // [B, T, 11]
    size_t bytes;
};

inline Workspace carve(void* base, int64_t B, int T, int W) {
    Workspace ws{};
    size_t off = 0;
    auto take = [&](size_t n) {
        size_t o = off;
        off += align_up(n);
        return base ? reinterpret_cast<float*>(static_cast<char*>(base) + o) : nullptr;
    };
    // window counts are padded to a multiple of 32 so that window tiles of the tensor kernels may
    // read/write the rows of (non-existent) windows past B without bounds checks
    const size_t Bp = ((size_t)B + 31) / 32 * 32;
    const size_t rows = Bp * (size_t)std::max(W, 0);
    ws.gi = take(rows * 2 * G * sizeof(float));
    ws.y1 = take(rows * 2 * H * sizeof(float));
    ws.y2 = take(rows * 2 * H * sizeof(float));
    ws.hid_a = take(Bp * 2 * H * sizeof(float));
    ws.hid_b = take(Bp * 2 * H * sizeof(float));
    ws.p_base = take((size_t)B * T * NBASE * sizeof(float));
    ws.p_rle = take((size_t)B * T * NRLE * sizeof(float));
    ws.bytes = off;
    return ws;
}

}  // namespace hb
