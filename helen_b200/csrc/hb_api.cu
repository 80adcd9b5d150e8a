// C ABI of the B200-native HELEN predict hot path (see include/helen_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/helen_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "fp32_kernels.cuh"
#include "train_kernels.cuh"
#include "workspace.cuh"
#ifndef HB_NO_TENSOR_ENGINE
#include "tensor_engine.cuh"
#endif

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

#define HB_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(HB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                        __FILE__, __LINE__);                                                  \
    } while (0)

using hb::Workspace;
using hb::carve;
using hb::kAlign;
using hb::align_up;

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(device) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

struct hb_handle {
    int device = 0;
    int features = 0;
    int engine = HB_ENGINE_FP32;
    int sm_count = 0;
    int64_t launches = 0;

    // fp32 engine weights (device)
    float* enc_wcat = nullptr;   // [768, F]   fwd rows 0..383, reverse rows 384..767
    float* enc_bcat = nullptr;   // [768]      b_ih
    float* enc_whh = nullptr;    // [2, 384, 128]
    float* enc_bhh = nullptr;    // [2, 384]
    float* dec_wcat = nullptr;   // [768, 256]
    float* dec_bcat = nullptr;
    float* dec_whh = nullptr;
    float* dec_bhh = nullptr;
    float* w_head = nullptr;     // [16, 256]
    float* b_head = nullptr;     // [16]

#ifndef HB_NO_TENSOR_ENGINE
    hb::TensorEngine* tensor = nullptr;
#endif

    // host-entry staging (owned)
    uint8_t* stage_images_host = nullptr;   // pinned
    uint8_t* stage_labels_host = nullptr;   // pinned, 2 * B * T
    uint8_t* stage_images_dev = nullptr;
    uint8_t* stage_labels_dev = nullptr;
    float* stage_prob_dev = nullptr;
    char* stage_workspace = nullptr;
    size_t cap_images = 0, cap_images_dev = 0, cap_labels = 0, cap_labels_dev = 0, cap_prob = 0, cap_workspace = 0;
    cudaStream_t host_stream = nullptr;
    // The engines keep per-handle device state (chunk-loop counters, job tables) and the chunk-loop kernel wants the whole
    // chip: successive predict calls of one handle are serialised on the device even when they come on different streams.
    cudaEvent_t last_predict = nullptr;
    bool predicted = false;

    // device timing of the predict kernel sequence
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    size_t events_used = 0;
    double timed_ms = 0.0;
    int64_t timed_launches = 0;
};

namespace {

int upload(float** dst, const float* src, size_t n) {
    HB_CUDA(cudaMalloc(dst, n * sizeof(float)));
    HB_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return HB_OK;
}

int pack_gru(const hb_gru_weights& g, int K, float** wcat, float** bcat, float** whh, float** bhh) {
    std::vector<float> wc((size_t)2 * hb::G * K), bc(2 * hb::G), wh((size_t)2 * hb::G * hb::H), bh(2 * hb::G);
    for (int d = 0; d < 2; ++d) {
        if (!g.weight_ih[d] || !g.weight_hh[d] || !g.bias_ih[d] || !g.bias_hh[d])
            return fail(HB_ERR_INVALID_ARGUMENT, "hb_create: null GRU weight pointer (direction %d)", d);
        std::memcpy(&wc[(size_t)d * hb::G * K], g.weight_ih[d], (size_t)hb::G * K * sizeof(float));
        std::memcpy(&bc[(size_t)d * hb::G], g.bias_ih[d], hb::G * sizeof(float));
        std::memcpy(&wh[(size_t)d * hb::G * hb::H], g.weight_hh[d], (size_t)hb::G * hb::H * sizeof(float));
        std::memcpy(&bh[(size_t)d * hb::G], g.bias_hh[d], hb::G * sizeof(float));
    }
    int rc;
    if ((rc = upload(wcat, wc.data(), wc.size()))) return rc;
    if ((rc = upload(bcat, bc.data(), bc.size()))) return rc;
    if ((rc = upload(whh, wh.data(), wh.size()))) return rc;
    if ((rc = upload(bhh, bh.data(), bh.size()))) return rc;
    return HB_OK;
}

// One layer of one chunk on the fp32 engine: projection then recurrence.
template <typename TA>
int run_layer(hb_handle* h, const TA* a, int64_t a_batch_stride, int64_t a_row_stride, int K,
              const float* wcat, const float* bcat, const float* whh, const float* bhh,
              const float* h_in, float* h_out, float* gi, float* y, int64_t B, int W, cudaStream_t s) {
    const int64_t M = B * W;
    dim3 grid_p((unsigned)((M + 63) / 64), 2 * hb::G / 64);
    hb::input_projection_kernel<TA><<<grid_p, 256, 0, s>>>(a, a_batch_stride, a_row_stride, W, M, K, wcat, bcat, gi);
    dim3 grid_r((unsigned)((B + hb::REC_WINDOWS - 1) / hb::REC_WINDOWS), 2);
    hb::gru_recurrence_kernel<<<grid_r, hb::REC_THREADS, 0, s>>>(gi, whh, bhh, h_in, h_out, y, B, W);
    h->launches += 2;
    HB_CUDA(cudaGetLastError());
    return HB_OK;
}

// Measurement mode: the events were created by hb_enable_kernel_timing; calls beyond them go untimed (nothing is created
// inside a predict call).
int begin_timing(hb_handle* h, cudaStream_t s, size_t* slot) {
    *slot = (size_t)-1;
    if (!h->timing || h->events_used == h->events.size()) return HB_OK;
    *slot = h->events_used++;
    HB_CUDA(cudaEventRecord(h->events[*slot].first, s));
    return HB_OK;
}

int end_timing(hb_handle* h, cudaStream_t s, size_t slot) {
    if (slot == (size_t)-1) return HB_OK;
    HB_CUDA(cudaEventRecord(h->events[slot].second, s));
    return HB_OK;
}

int predict_fp32(hb_handle* h, const uint8_t* images, int64_t B, int T, int W, int J,
                 uint8_t* base_labels, uint8_t* rle_labels, float* p_base, float* p_rle,
                 const Workspace& ws, cudaStream_t s) {
    const int F = h->features;
    HB_CUDA(cudaMemsetAsync(p_base, 0, (size_t)B * T * hb::NBASE * sizeof(float), s));
    HB_CUDA(cudaMemsetAsync(p_rle, 0, (size_t)B * T * hb::NRLE * sizeof(float), s));
    const float* hid = nullptr;            // zeros for the first chunk (predict_gpu.py:99)
    float* hid_bufs[2] = {ws.hid_a, ws.hid_b};
    int flip = 0, rc;
    for (int i = 0; i + W <= T; i += J) {  // predict_gpu.py:114-117
        float* enc_h = hid_bufs[flip];
        float* dec_h = hid_bufs[flip ^ 1];
        if ((rc = run_layer<uint8_t>(h, images + (int64_t)i * F, (int64_t)T * F, F, F, h->enc_wcat, h->enc_bcat,
                                     h->enc_whh, h->enc_bhh, hid, enc_h, ws.gi, ws.y1, B, W, s))) return rc;
        if ((rc = run_layer<float>(h, ws.y1, (int64_t)W * 2 * hb::H, 2 * hb::H, 2 * hb::H, h->dec_wcat, h->dec_bcat,
                                   h->dec_whh, h->dec_bhh, enc_h, dec_h, ws.gi, ws.y2, B, W, s))) return rc;
        const int64_t rows = B * W;
        const int blocks = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)h->sm_count * 8);
        hb::heads_kernel<<<blocks, 256, 0, s>>>(ws.y2, h->w_head, h->b_head, rows, W, T, i, p_base, p_rle,
                                               nullptr, nullptr, 0);
        h->launches += 1;
        hid = dec_h;
        flip ^= 1;
    }
    const int64_t positions = B * T;
    if (positions > 0) {
        hb::argmax_kernel<<<(unsigned)((positions + 255) / 256), 256, 0, s>>>(p_base, p_rle, positions, base_labels, rle_labels);
        h->launches += 1;
    }
    HB_CUDA(cudaGetLastError());
    return HB_OK;
}

int check_predict_args(const hb_handle* h, int64_t B, int T, int W, int J) {
    if (!h) return fail(HB_ERR_INVALID_ARGUMENT, "null handle");
    if (B < 0 || T < 0) return fail(HB_ERR_INVALID_ARGUMENT, "negative batch (%lld) or length (%d)", (long long)B, T);
    if (W <= 0 || J <= 0) return fail(HB_ERR_INVALID_ARGUMENT, "chunk width W=%d and jump J=%d must be positive", W, J);
    if ((int64_t)B * std::max(T, 1) > (int64_t)1 << 40) return fail(HB_ERR_INVALID_ARGUMENT, "batch too large");
    return HB_OK;
}

template <typename Tp>
int grow(Tp** ptr, size_t* cap, size_t need, bool pinned) {
    if (need <= *cap) return HB_OK;
    if (*ptr) {
        if (pinned) cudaFreeHost(*ptr); else cudaFree(*ptr);
        *ptr = nullptr;
        *cap = 0;
    }
    size_t want = std::max(need, *cap * 2);
    cudaError_t e = pinned ? cudaMallocHost(reinterpret_cast<void**>(ptr), want) : cudaMalloc(reinterpret_cast<void**>(ptr), want);
    if (e != cudaSuccess) return fail(HB_ERR_OUT_OF_MEMORY, "allocating %zu staging bytes: %s", want, cudaGetErrorString(e));
    *cap = want;
    return HB_OK;
}

}  // namespace

extern "C" {

int hb_abi_version(void) { return HB_ABI_VERSION; }

const char* hb_last_error(void) { return g_error; }

int hb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(HB_ERR_UNSUPPORTED_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int hb_create(const hb_weights* w, int image_features, int hidden, int n_base, int n_rle, int device,
              hb_handle** out) {
    if (!w || !out) return fail(HB_ERR_INVALID_ARGUMENT, "hb_create: null argument");
    *out = nullptr;
    if (hidden != hb::H || n_base != hb::NBASE || n_rle != hb::NRLE)
        return fail(HB_ERR_INVALID_ARGUMENT,
                    "hb_create: only hidden_size=128, 5 base classes, 11 rle classes are supported "
                    "(got %d, %d, %d); the reference ships no other configuration (Options.py:20-28)",
                    hidden, n_base, n_rle);
    if (image_features < 1 || image_features > 256)
        return fail(HB_ERR_INVALID_ARGUMENT, "hb_create: image_features=%d outside [1, 256]", image_features);
    if (!w->base_weight || !w->base_bias || !w->rle_weight || !w->rle_bias)
        return fail(HB_ERR_INVALID_ARGUMENT, "hb_create: null head weight pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(HB_ERR_UNSUPPORTED_DEVICE, "hb_create: no CUDA device (%s); there is no CPU fallback",
                    e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(HB_ERR_INVALID_ARGUMENT, "hb_create: device %d of %d", device, count);
    cudaDeviceProp prop{};
    HB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(HB_ERR_UNSUPPORTED_DEVICE, "hb_create: device %d is sm_%d%d; this library is sm_100a only",
                    device, prop.major, prop.minor);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(HB_ERR_CUDA, "hb_create: cudaSetDevice(%d) failed", device);

    hb_handle* h = new (std::nothrow) hb_handle();
    if (!h) return fail(HB_ERR_OUT_OF_MEMORY, "hb_create: host allocation failed");
    h->device = device;
    h->features = image_features;
    h->sm_count = prop.multiProcessorCount;
    int rc;
    if ((rc = pack_gru(w->encoder, image_features, &h->enc_wcat, &h->enc_bcat, &h->enc_whh, &h->enc_bhh)) ||
        (rc = pack_gru(w->decoder, 2 * hb::H, &h->dec_wcat, &h->dec_bcat, &h->dec_whh, &h->dec_bhh))) {
        hb_destroy(h);
        return rc;
    }
    std::vector<float> wh((size_t)hb::NCLS * 2 * hb::H), bh(hb::NCLS);
    std::memcpy(wh.data(), w->base_weight, (size_t)hb::NBASE * 2 * hb::H * sizeof(float));
    std::memcpy(wh.data() + (size_t)hb::NBASE * 2 * hb::H, w->rle_weight, (size_t)hb::NRLE * 2 * hb::H * sizeof(float));
    std::memcpy(bh.data(), w->base_bias, hb::NBASE * sizeof(float));
    std::memcpy(bh.data() + hb::NBASE, w->rle_bias, hb::NRLE * sizeof(float));
    if ((rc = upload(&h->w_head, wh.data(), wh.size())) || (rc = upload(&h->b_head, bh.data(), bh.size()))) {
        hb_destroy(h);
        return rc;
    }
    {
        cudaError_t ee = cudaEventCreateWithFlags(&h->last_predict, cudaEventDisableTiming);
        if (ee != cudaSuccess) {
            hb_destroy(h);
            return fail(HB_ERR_CUDA, "hb_create: cudaEventCreate failed: %s", cudaGetErrorString(ee));
        }
    }
#ifndef HB_NO_TENSOR_ENGINE
    h->tensor = hb::tensor_engine_create(w, image_features, h->sm_count, g_error, sizeof(g_error));
    if (!h->tensor) {
        hb_destroy(h);
        return HB_ERR_CUDA;
    }
    h->engine = HB_ENGINE_TENSOR;
#endif
    *out = h;
    return HB_OK;
}

void hb_destroy(hb_handle* h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    for (float* p : {h->enc_wcat, h->enc_bcat, h->enc_whh, h->enc_bhh, h->dec_wcat, h->dec_bcat, h->dec_whh,
                     h->dec_bhh, h->w_head, h->b_head})
        if (p) cudaFree(p);
#ifndef HB_NO_TENSOR_ENGINE
    if (h->tensor) hb::tensor_engine_destroy(h->tensor);
#endif
    if (h->stage_images_host) cudaFreeHost(h->stage_images_host);
    if (h->stage_labels_host) cudaFreeHost(h->stage_labels_host);
    if (h->stage_images_dev) cudaFree(h->stage_images_dev);
    if (h->stage_labels_dev) cudaFree(h->stage_labels_dev);
    if (h->stage_prob_dev) cudaFree(h->stage_prob_dev);
    if (h->stage_workspace) cudaFree(h->stage_workspace);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    if (h->last_predict) cudaEventDestroy(h->last_predict);
    for (auto& ev : h->events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    delete h;
}

int hb_set_engine(hb_handle* h, int engine) {
    if (!h) return fail(HB_ERR_INVALID_ARGUMENT, "null handle");
    if (engine == HB_ENGINE_DEFAULT) {
#ifndef HB_NO_TENSOR_ENGINE
        engine = HB_ENGINE_TENSOR;
#else
        engine = HB_ENGINE_FP32;
#endif
    }
#ifdef HB_NO_TENSOR_ENGINE
    if (engine == HB_ENGINE_TENSOR) return fail(HB_ERR_INVALID_ARGUMENT, "tensor engine not built into this library");
#endif
    if (engine != HB_ENGINE_FP32 && engine != HB_ENGINE_TENSOR)
        return fail(HB_ERR_INVALID_ARGUMENT, "unknown engine %d", engine);
    h->engine = engine;
    return HB_OK;
}

int hb_get_engine(const hb_handle* h) { return h ? h->engine : fail(HB_ERR_INVALID_ARGUMENT, "null handle"); }

int hb_workspace_bytes(const hb_handle* h, int64_t B, int T, int W, size_t* out) {
    if (!out) return fail(HB_ERR_INVALID_ARGUMENT, "null out");
    int rc = check_predict_args(h, B, T, W, 1);
    if (rc) return rc;
    size_t bytes = carve(nullptr, B, T, W).bytes;
#ifndef HB_NO_TENSOR_ENGINE
    bytes = std::max(bytes, hb::tensor_engine_workspace_bytes(h->tensor, B, T, W));
#endif
    *out = bytes + kAlign;
    return HB_OK;
}

int hb_predict_windows(hb_handle* h, const uint8_t* images_dev, int64_t B, int T, int W, int J,
                       uint8_t* base_labels_dev, uint8_t* rle_labels_dev, float* base_prob_dev,
                       float* rle_prob_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    int rc = check_predict_args(h, B, T, W, J);
    if (rc) return rc;
    if (B == 0 || T == 0) return HB_OK;
    if (!images_dev || !base_labels_dev || !rle_labels_dev)
        return fail(HB_ERR_INVALID_ARGUMENT, "hb_predict_windows: null image/label pointer");
    size_t need = 0;
    if ((rc = hb_workspace_bytes(h, B, T, W, &need))) return rc;
    if (!workspace_dev || workspace_bytes < need)
        return fail(HB_ERR_WORKSPACE, "hb_predict_windows: workspace %zu bytes < required %zu", workspace_bytes, need);
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    void* base = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace_dev)));
    if (h->predicted) HB_CUDA(cudaStreamWaitEvent(s, h->last_predict, 0));
    h->predicted = true;
    size_t slot;
    if ((rc = begin_timing(h, s, &slot))) return rc;
#ifndef HB_NO_TENSOR_ENGINE
    if (h->engine == HB_ENGINE_TENSOR) {
        int n = hb::tensor_engine_predict(h->tensor, images_dev, B, T, W, J, base_labels_dev, rle_labels_dev,
                                          base_prob_dev, rle_prob_dev, base, s, g_error, sizeof(g_error));
        if (n < 0) return n;
        h->launches += n;
    } else
#endif
    {
        Workspace ws = carve(base, B, T, W);
        float* pb = base_prob_dev ? base_prob_dev : ws.p_base;
        float* pr = rle_prob_dev ? rle_prob_dev : ws.p_rle;
        if ((rc = predict_fp32(h, images_dev, B, T, W, J, base_labels_dev, rle_labels_dev, pb, pr, ws, s))) return rc;
    }
    if ((rc = end_timing(h, s, slot))) return rc;
    h->timed_launches += (slot != (size_t)-1);
    HB_CUDA(cudaEventRecord(h->last_predict, s));
    return HB_OK;
}

int hb_predict_windows_host(hb_handle* h, const uint8_t* images_host, int64_t B, int T, int W, int J,
                            uint8_t* base_labels_host, uint8_t* rle_labels_host, float* base_prob_host,
                            float* rle_prob_host) {
    int rc = check_predict_args(h, B, T, W, J);
    if (rc) return rc;
    if (B == 0 || T == 0) return HB_OK;
    if (!images_host || !base_labels_host || !rle_labels_host)
        return fail(HB_ERR_INVALID_ARGUMENT, "hb_predict_windows_host: null image/label pointer");
    DeviceGuard guard(h->device);
    if (!h->host_stream) HB_CUDA(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    cudaStream_t s = h->host_stream;
    const size_t img_bytes = (size_t)B * T * h->features, lab_bytes = (size_t)B * T;
    const bool want_prob = base_prob_host || rle_prob_host;
    const size_t pb_bytes = lab_bytes * hb::NBASE * sizeof(float), pr_bytes = lab_bytes * hb::NRLE * sizeof(float);
    size_t ws_bytes = 0;
    if ((rc = hb_workspace_bytes(h, B, T, W, &ws_bytes))) return rc;
    if ((rc = grow(&h->stage_images_host, &h->cap_images, img_bytes, true))) return rc;
    if ((rc = grow(&h->stage_images_dev, &h->cap_images_dev, img_bytes, false))) return rc;
    if ((rc = grow(&h->stage_labels_host, &h->cap_labels, 2 * lab_bytes, true))) return rc;
    if ((rc = grow(&h->stage_labels_dev, &h->cap_labels_dev, 2 * lab_bytes, false))) return rc;
    if ((rc = grow(&h->stage_workspace, &h->cap_workspace, ws_bytes, false))) return rc;
    if (want_prob && (rc = grow(&h->stage_prob_dev, &h->cap_prob, pb_bytes + pr_bytes, false))) return rc;

    // caller memory -> device (the reference's images.to(device)).  Page-locked caller memory (torch pin_memory, the
    // DataLoader's pinned batches) is copied from directly; pageable memory goes through the pinned staging buffer in
    // slices, so the host memcpy of one slice overlaps the DMA of the previous one.
    cudaPointerAttributes attr{};
    const bool caller_pinned = cudaPointerGetAttributes(&attr, images_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();                                      // (unregistered host memory reports an error on old drivers)
    if (caller_pinned) {
        HB_CUDA(cudaMemcpyAsync(h->stage_images_dev, images_host, img_bytes, cudaMemcpyHostToDevice, s));
    } else {
        const size_t slice = std::max<size_t>((img_bytes / 8 + 4095) / 4096 * 4096, (size_t)1 << 20);
        for (size_t off = 0; off < img_bytes; off += slice) {
            const size_t n = std::min(slice, img_bytes - off);
            std::memcpy(h->stage_images_host + off, images_host + off, n);
            HB_CUDA(cudaMemcpyAsync(h->stage_images_dev + off, h->stage_images_host + off, n, cudaMemcpyHostToDevice, s));
        }
    }
    float* pb = want_prob ? h->stage_prob_dev : nullptr;
    float* pr = want_prob ? reinterpret_cast<float*>(reinterpret_cast<char*>(h->stage_prob_dev) + pb_bytes) : nullptr;
    rc = hb_predict_windows(h, h->stage_images_dev, B, T, W, J, h->stage_labels_dev, h->stage_labels_dev + lab_bytes,
                            pb, pr, h->stage_workspace, h->cap_workspace, s);
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(h->stage_labels_host, h->stage_labels_dev, 2 * lab_bytes, cudaMemcpyDeviceToHost, s));
    if (base_prob_host) HB_CUDA(cudaMemcpyAsync(base_prob_host, pb, pb_bytes, cudaMemcpyDeviceToHost, s));
    if (rle_prob_host) HB_CUDA(cudaMemcpyAsync(rle_prob_host, pr, pr_bytes, cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    std::memcpy(base_labels_host, h->stage_labels_host, lab_bytes);
    std::memcpy(rle_labels_host, h->stage_labels_host + lab_bytes, lab_bytes);
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
// Training step of one chunk (row a15): see train_kernels.cuh
// ---------------------------------------------------------------------------------------------
namespace {

struct TrainWorkspace {
    float *gi1, *y1, *sv1, *gi2, *y2, *sv2, *h_enc, *logit_b, *logit_r, *sums, *dlogits, *dy, *dgi, *dgh, *hprev, *dh_dec0;
    size_t bytes;
};

TrainWorkspace train_carve(void* base, int64_t B, int W) {
    TrainWorkspace ws{};
    size_t off = 0;
    auto take = [&](size_t n) {
        size_t o = off;
        off += hb::align_up(n * sizeof(float));
        return base ? reinterpret_cast<float*>(static_cast<char*>(base) + o) : nullptr;
    };
    const size_t M = (size_t)B * (size_t)W;
    ws.gi1 = take(M * 2 * hb::G); ws.y1 = take(M * 2 * hb::H); ws.sv1 = take(M * 8 * hb::H);
    ws.gi2 = take(M * 2 * hb::G); ws.y2 = take(M * 2 * hb::H); ws.sv2 = take(M * 8 * hb::H);
    ws.h_enc = take((size_t)B * 2 * hb::H);
    ws.logit_b = take(M * hb::NBASE); ws.logit_r = take(M * hb::NRLE);
    ws.sums = take(4); ws.dlogits = take(M * hb::NCLS);
    ws.dy = take(M * 2 * hb::H); ws.dgi = take(M * 2 * hb::G); ws.dgh = take(M * 2 * hb::G); ws.hprev = take(M * 2 * hb::H);
    ws.dh_dec0 = take((size_t)B * 2 * hb::H);
    ws.bytes = off;
    return ws;
}

// C[M, N] (+)= A . B with element strides; split_k > 1 (or accumulate) adds into C
void gemm(cudaStream_t s, const float* a, int64_t a_m, int64_t a_k, const float* b, int64_t b_k, int64_t b_n,
          float* c, int64_t c_m, int64_t c_n, const float* bias, int64_t M, int64_t N, int64_t K, int split_k) {
    hb::train::GemmArgs g{a, a_m, a_k, b, b_k, b_n, c, c_m, c_n, bias, M, N, K, split_k < 0 ? 1 : 0};   // split_k < 0: accumulate, no split
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)std::max(1, split_k));
    hb::train::gemm_kernel<<<grid, 256, 0, s>>>(g);
}

void colsum(cudaStream_t s, const float* a, int64_t M, int N, int64_t lda, float* out) {
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, M / 256)));
    hb::train::colsum_kernel<<<grid, 256, 0, s>>>(a, M, N, lda, out);
}

bool gru_ptrs_ok(const hb_gru_weights& g) {
    for (int d = 0; d < 2; ++d)
        if (!g.weight_ih[d] || !g.weight_hh[d] || !g.bias_ih[d] || !g.bias_hh[d]) return false;
    return true;
}

}  // namespace

int hb_train_workspace_bytes(const hb_handle* h, int64_t B, int W, size_t* out) {
    if (!h || !out) return fail(HB_ERR_INVALID_ARGUMENT, "null argument");
    if (B < 0 || W <= 0) return fail(HB_ERR_INVALID_ARGUMENT, "hb_train_workspace_bytes: bad shape B=%lld W=%d", (long long)B, W);
    *out = train_carve(nullptr, B, W).bytes + hb::kAlign;
    return HB_OK;
}

int hb_train_step_chunk(hb_handle* h, const hb_weights* w, const hb_weights* gr, const float* x_dev, const float* h_in_dev,
                        const int64_t* label_base_dev, const int64_t* label_rle_dev, const float* rle_class_weights_dev,
                        int64_t B, int W, float* loss_dev, float* h_out_dev, float* base_logits_dev, float* rle_logits_dev,
                        void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!h) return fail(HB_ERR_INVALID_ARGUMENT, "null handle");
    if (B < 0 || W <= 0) return fail(HB_ERR_INVALID_ARGUMENT, "hb_train_step_chunk: bad shape B=%lld W=%d", (long long)B, W);
    if (B == 0) return HB_OK;
    if (!w || !gr || !x_dev || !label_base_dev || !label_rle_dev || !rle_class_weights_dev || !loss_dev || !h_out_dev)
        return fail(HB_ERR_INVALID_ARGUMENT, "hb_train_step_chunk: null pointer");
    if (!gru_ptrs_ok(w->encoder) || !gru_ptrs_ok(w->decoder) || !w->base_weight || !w->base_bias || !w->rle_weight || !w->rle_bias ||
        !gru_ptrs_ok(gr->encoder) || !gru_ptrs_ok(gr->decoder) || !gr->base_weight || !gr->base_bias || !gr->rle_weight || !gr->rle_bias)
        return fail(HB_ERR_INVALID_ARGUMENT, "hb_train_step_chunk: null weight or gradient pointer");
    size_t need = 0;
    int rc;
    if ((rc = hb_train_workspace_bytes(h, B, W, &need))) return rc;
    if (!workspace_dev || workspace_bytes < need)
        return fail(HB_ERR_WORKSPACE, "hb_train_step_chunk: workspace %zu bytes < required %zu", workspace_bytes, need);
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TrainWorkspace ws = train_carve(reinterpret_cast<void*>(hb::align_up(reinterpret_cast<size_t>(workspace_dev))), B, W);
    const int F = h->features, H = hb::H, G = hb::G;
    const int64_t M = B * W;
    const dim3 grid_rec((unsigned)((B + hb::REC_WINDOWS - 1) / hb::REC_WINDOWS), 2);
    float* logit_b = base_logits_dev ? base_logits_dev : ws.logit_b;
    float* logit_r = rle_logits_dev ? rle_logits_dev : ws.logit_r;
    const int split = 16;                                     // k-splits of the weight-gradient GEMMs (K = B * W)

    // ---- forward (TransducerModel.py:60-79), activations kept ----
    for (int d = 0; d < 2; ++d)
        gemm(s, x_dev, F, 1, w->encoder.weight_ih[d], 1, F, ws.gi1 + d * G, 2 * G, 1, w->encoder.bias_ih[d], M, G, F, 1);
    hb::train::gru_forward_save_kernel<<<grid_rec, hb::REC_THREADS, 0, s>>>(ws.gi1, w->encoder.weight_hh[0], w->encoder.weight_hh[1],
        w->encoder.bias_hh[0], w->encoder.bias_hh[1], h_in_dev, ws.h_enc, ws.y1, ws.sv1, B, W);
    for (int d = 0; d < 2; ++d)
        gemm(s, ws.y1, 2 * H, 1, w->decoder.weight_ih[d], 1, 2 * H, ws.gi2 + d * G, 2 * G, 1, w->decoder.bias_ih[d], M, G, 2 * H, 1);
    hb::train::gru_forward_save_kernel<<<grid_rec, hb::REC_THREADS, 0, s>>>(ws.gi2, w->decoder.weight_hh[0], w->decoder.weight_hh[1],
        w->decoder.bias_hh[0], w->decoder.bias_hh[1], ws.h_enc, h_out_dev, ws.y2, ws.sv2, B, W);
    gemm(s, ws.y2, 2 * H, 1, w->base_weight, 1, 2 * H, logit_b, hb::NBASE, 1, w->base_bias, M, hb::NBASE, 2 * H, 1);
    gemm(s, ws.y2, 2 * H, 1, w->rle_weight, 1, 2 * H, logit_r, hb::NRLE, 1, w->rle_bias, M, hb::NRLE, 2 * H, 1);

    // ---- losses (train.py:121-126, 192-198) and their gradient ----
    HB_CUDA(cudaMemsetAsync(ws.sums, 0, 4 * sizeof(float), s));
    const unsigned ce_blocks = (unsigned)((M + 255) / 256);
    hb::train::ce_kernel<1><<<ce_blocks, 256, 0, s>>>(logit_b, logit_r, label_base_dev, label_rle_dev, rle_class_weights_dev, M, ws.sums, nullptr);
    hb::train::loss_finish_kernel<<<1, 1, 0, s>>>(ws.sums, loss_dev);
    hb::train::ce_kernel<2><<<ce_blocks, 256, 0, s>>>(logit_b, logit_r, label_base_dev, label_rle_dev, rle_class_weights_dev, M, ws.sums, ws.dlogits);

    // ---- backward: heads ----
    auto zero = [&](const float* p, size_t n) { return cudaMemsetAsync(const_cast<float*>(p), 0, n * sizeof(float), s); };
    HB_CUDA(zero(gr->base_weight, (size_t)hb::NBASE * 2 * H)); HB_CUDA(zero(gr->base_bias, hb::NBASE));
    HB_CUDA(zero(gr->rle_weight, (size_t)hb::NRLE * 2 * H)); HB_CUDA(zero(gr->rle_bias, hb::NRLE));
    float* g_bw = const_cast<float*>(gr->base_weight); float* g_bb = const_cast<float*>(gr->base_bias);
    float* g_rw = const_cast<float*>(gr->rle_weight); float* g_rb = const_cast<float*>(gr->rle_bias);
    gemm(s, ws.dlogits, 1, hb::NCLS, ws.y2, 2 * H, 1, g_bw, 2 * H, 1, nullptr, hb::NBASE, 2 * H, M, split);
    gemm(s, ws.dlogits + hb::NBASE, 1, hb::NCLS, ws.y2, 2 * H, 1, g_rw, 2 * H, 1, nullptr, hb::NRLE, 2 * H, M, split);
    colsum(s, ws.dlogits, M, hb::NBASE, hb::NCLS, g_bb);
    colsum(s, ws.dlogits + hb::NBASE, M, hb::NRLE, hb::NCLS, g_rb);
    // dy2 = dl_base . W_base + dl_rle . W_rle   (the second product accumulates)
    gemm(s, ws.dlogits, hb::NCLS, 1, w->base_weight, 2 * H, 1, ws.dy, 2 * H, 1, nullptr, M, 2 * H, hb::NBASE, 1);
    gemm(s, ws.dlogits + hb::NBASE, hb::NCLS, 1, w->rle_weight, 2 * H, 1, ws.dy, 2 * H, 1, nullptr, M, 2 * H, hb::NRLE, -1);

    // ---- backward: one GRU layer (recurrence kernel, then the weight-gradient GEMMs) ----
    auto layer_backward = [&](const hb_gru_weights& lw, const hb_gru_weights& lg, const float* in, int K, const float* dh_n,
                              const float* sv, const float* y, const float* h0, float* dh0) -> int {
        hb::train::gru_backward_kernel<<<grid_rec, hb::REC_THREADS, 0, s>>>(ws.dy, dh_n, sv, y, h0, lw.weight_hh[0], lw.weight_hh[1],
                                                                           ws.dgi, ws.dgh, ws.hprev, dh0, B, W);
        for (int d = 0; d < 2; ++d) {
            float* g_wih = const_cast<float*>(lg.weight_ih[d]); float* g_whh = const_cast<float*>(lg.weight_hh[d]);
            float* g_bih = const_cast<float*>(lg.bias_ih[d]); float* g_bhh = const_cast<float*>(lg.bias_hh[d]);
            HB_CUDA(zero(g_wih, (size_t)G * K)); HB_CUDA(zero(g_whh, (size_t)G * H)); HB_CUDA(zero(g_bih, G)); HB_CUDA(zero(g_bhh, G));
            gemm(s, ws.dgi + d * G, 1, 2 * G, in, K, 1, g_wih, K, 1, nullptr, G, K, M, split);                    // dW_ih = dgi^T . x
            gemm(s, ws.dgh + d * G, 1, 2 * G, ws.hprev + d * H, 2 * H, 1, g_whh, H, 1, nullptr, G, H, M, split);  // dW_hh = dgh^T . h_prev
            colsum(s, ws.dgi + d * G, M, G, 2 * G, g_bih);
            colsum(s, ws.dgh + d * G, M, G, 2 * G, g_bhh);
        }
        return HB_OK;
    };
    // decoder: dy = dy2, no gradient through the returned state (the loss does not see it; train.py:206 detaches it)
    if ((rc = layer_backward(w->decoder, gr->decoder, ws.y1, 2 * H, nullptr, ws.sv2, ws.y2, ws.h_enc, ws.dh_dec0))) return rc;
    // gradient wrt the encoder output: dy1 = sum_d dgi2[:, d] . W_ih_dec[d]   (overwrites dy; the second direction accumulates)
    gemm(s, ws.dgi, 2 * G, 1, w->decoder.weight_ih[0], 2 * H, 1, ws.dy, 2 * H, 1, nullptr, M, 2 * H, G, 1);
    gemm(s, ws.dgi + G, 2 * G, 1, w->decoder.weight_ih[1], 2 * H, 1, ws.dy, 2 * H, 1, nullptr, M, 2 * H, G, -1);
    // encoder: its final state fed the decoder's initial state
    if ((rc = layer_backward(w->encoder, gr->encoder, x_dev, F, ws.dh_dec0, ws.sv1, ws.y1, h_in_dev, nullptr))) return rc;
    h->launches += 40;
    HB_CUDA(cudaGetLastError());
    return HB_OK;
}

int hb_forward_chunk(hb_handle* h, const float* x_dev, const float* h_in_dev, int64_t B, int W,
                     float* base_logits_dev, float* rle_logits_dev, float* h_out_dev, void* workspace_dev,
                     size_t workspace_bytes, void* stream) {
    int rc = check_predict_args(h, B, W, W, 1);
    if (rc) return rc;
    if (B == 0) return HB_OK;
    if (!x_dev || !h_in_dev || !base_logits_dev || !rle_logits_dev || !h_out_dev)
        return fail(HB_ERR_INVALID_ARGUMENT, "hb_forward_chunk: null pointer");
    size_t need = 0;
    if ((rc = hb_workspace_bytes(h, B, W, W, &need))) return rc;
    if (!workspace_dev || workspace_bytes < need)
        return fail(HB_ERR_WORKSPACE, "hb_forward_chunk: workspace %zu bytes < required %zu", workspace_bytes, need);
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Workspace ws = carve(reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace_dev))), B, W, W);
    const int F = h->features;
    if ((rc = run_layer<float>(h, x_dev, (int64_t)W * F, F, F, h->enc_wcat, h->enc_bcat, h->enc_whh, h->enc_bhh,
                               h_in_dev, ws.hid_a, ws.gi, ws.y1, B, W, s))) return rc;
    if ((rc = run_layer<float>(h, ws.y1, (int64_t)W * 2 * hb::H, 2 * hb::H, 2 * hb::H, h->dec_wcat, h->dec_bcat,
                               h->dec_whh, h->dec_bhh, ws.hid_a, h_out_dev, ws.gi, ws.y2, B, W, s))) return rc;
    const int64_t rows = B * W;
    const int blocks = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)h->sm_count * 8);
    hb::heads_kernel<<<blocks, 256, 0, s>>>(ws.y2, h->w_head, h->b_head, rows, W, W, 0, nullptr, nullptr,
                                           base_logits_dev, rle_logits_dev, 1);
    h->launches += 1;
    HB_CUDA(cudaGetLastError());
    return HB_OK;
}

int64_t hb_launch_count(const hb_handle* h) { return h ? h->launches : 0; }

int hb_enable_kernel_timing(hb_handle* h, int enable) {
    if (!h) return fail(HB_ERR_INVALID_ARGUMENT, "null handle");
    DeviceGuard guard(h->device);
    // every event the measurement mode uses is created here, never inside a predict call
    auto ensure = [](std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v, size_t n) -> cudaError_t {
        while (v.size() < n) {
            cudaEvent_t a, b;
            cudaError_t e = cudaEventCreate(&a);
            if (e == cudaSuccess) e = cudaEventCreate(&b);
            if (e != cudaSuccess) return e;
            v.emplace_back(a, b);
        }
        return cudaSuccess;
    };
    if (enable) HB_CUDA(ensure(h->events, 4096));
    h->timing = enable != 0;
#ifndef HB_NO_TENSOR_ENGINE
    if (enable == 2) HB_CUDA(ensure(h->tensor->rec_events, 1024));
    h->tensor->time_recurrence = enable == 2;
#endif
    return HB_OK;
}

int hb_last_launch_plan(const hb_handle* h, hb_launch_plan* out) {
    if (!h || !out) return fail(HB_ERR_INVALID_ARGUMENT, "null argument");
    std::memset(out, 0, sizeof(*out));
#ifndef HB_NO_TENSOR_ENGINE
    *out = h->tensor->last_plan;
#endif
    return HB_OK;
}

int hb_dominant_kernel_time_ms(hb_handle* h, double* total_ms, int64_t* launches, int reset) {
    if (!h || !total_ms || !launches) return fail(HB_ERR_INVALID_ARGUMENT, "null argument");
    *total_ms = 0.0;
    *launches = 0;
#ifndef HB_NO_TENSOR_ENGINE
    DeviceGuard guard(h->device);
    hb::TensorEngine* e = h->tensor;
    for (size_t i = 0; i < e->rec_events_used; ++i) {
        HB_CUDA(cudaEventSynchronize(e->rec_events[i].second));
        float ms = 0.f;
        HB_CUDA(cudaEventElapsedTime(&ms, e->rec_events[i].first, e->rec_events[i].second));
        *total_ms += ms;
    }
    *launches = (int64_t)e->rec_events_used;
    if (reset) e->rec_events_used = 0;
#endif
    return HB_OK;
}

int hb_kernel_time_ms(hb_handle* h, double* total_ms, int64_t* launches, int reset) {
    if (!h || !total_ms || !launches) return fail(HB_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard(h->device);
    for (size_t i = 0; i < h->events_used; ++i) {
        HB_CUDA(cudaEventSynchronize(h->events[i].second));
        float ms = 0.f;
        HB_CUDA(cudaEventElapsedTime(&ms, h->events[i].first, h->events[i].second));
        h->timed_ms += ms;
    }
    h->events_used = 0;
    *total_ms = h->timed_ms;
    *launches = h->timed_launches;
    if (reset) {
        h->timed_ms = 0.0;
        h->timed_launches = 0;
    }
    return HB_OK;
}

}  // extern "C"
