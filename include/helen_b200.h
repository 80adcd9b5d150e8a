/*
 * helen_b200.h -- C ABI of the B200-native HELEN call_consensus / predict hot path.
 *
 * The reference (kishwarshafin/helen @ a075e9f) is pure Python on top of torch; the
 * "FFI" a maintainer would bind for this path is therefore a ctypes binding of this
 * header (INTEGRATION.md shows the stub).  Every entry point below names the reference
 * interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross the boundary.
 *  - all functions return HB_OK (0) or a negative hb_status; hb_last_error() returns a
 *    thread-local message for the last failure on the calling thread.  Nothing throws.
 *  - a handle is bound to one CUDA device and is not thread-safe (one handle per
 *    process/GPU, like the reference's one-process-per-GPU predict(), predict_gpu.py:38).
 *  - "_dev" pointers are device memory owned by the caller; the handle owns only its
 *    packed copy of the weights and a few KB of scheduling tables sized in hb_create.
 *    Device entry points allocate nothing, create no events, never synchronise the
 *    host, and are ordered on the cudaStream_t passed as `stream` (successive calls of
 *    one handle are also ordered among themselves, whatever their streams).  The
 *    measurement modes (hb_enable_kernel_timing) create their events when enabled.
 *  - there is no CPU fallback: every call fails loudly without a CUDA sm_100 device.
 */
#ifndef HELEN_B200_H
#define HELEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 4

typedef enum hb_status {
    HB_OK = 0,
    HB_ERR_INVALID_ARGUMENT = -1, /* bad shape / null pointer / unsupported model config   */
    HB_ERR_UNSUPPORTED_DEVICE = -2, /* no CUDA device, or device is not sm_100              */
    HB_ERR_CUDA = -3,             /* a CUDA runtime call failed (message has the detail)    */
    HB_ERR_WORKSPACE = -4,        /* caller-provided workspace too small                    */
    HB_ERR_OUT_OF_MEMORY = -5
} hb_status;

/* Host pointers to the 20 fp32 parameter tensors, in torch state_dict layout, exactly
 * as TransducerGRU owns them (helen/modules/python/models/TransducerModel.py:43-58):
 * GRU rows ordered (r, z, n).  Index 0 = forward direction, 1 = "_reverse". */
typedef struct hb_gru_weights {
    const float *weight_ih[2]; /* [3H, K]  K = image_features (encoder) or 2H (decoder) */
    const float *weight_hh[2]; /* [3H, H] */
    const float *bias_ih[2];   /* [3H] */
    const float *bias_hh[2];   /* [3H] */
} hb_gru_weights;

typedef struct hb_weights {
    hb_gru_weights encoder;   /* gru_encoder.*_l0 / *_l0_reverse */
    hb_gru_weights decoder;   /* gru_decoder.*_l0 / *_l0_reverse */
    const float *base_weight; /* dense1_base.weight [n_base, 2H] */
    const float *base_bias;   /* dense1_base.bias   [n_base]     */
    const float *rle_weight;  /* dense2_rle.weight  [n_rle, 2H]  */
    const float *rle_bias;    /* dense2_rle.bias    [n_rle]      */
} hb_weights;

typedef struct hb_handle hb_handle;

/* Kernel selection for hb_predict_windows*.  Both engines implement the same
 * arithmetic contract; HB_ENGINE_FP32 keeps every contraction on fp32 FMA pipes,
 * HB_ENGINE_TENSOR runs them on tcgen05 tensor cores with split-fp16 (3-term) operands
 * and fp32 accumulation.  Not a backend dispatch: both are sm_100a CUDA in this library. */
typedef enum hb_engine {
    HB_ENGINE_DEFAULT = 0,
    HB_ENGINE_FP32 = 1,
    HB_ENGINE_TENSOR = 2
} hb_engine;

int hb_abi_version(void);
const char *hb_last_error(void);

/* Number of CUDA devices usable by this library (replaces torch.cuda.device_count()
 * at CallConsensusInterface.py:99).  Returns a count >= 0 or a negative hb_status. */
int hb_device_count(void);

/* Replaces TransducerGRU.__init__ + load_state_dict + .to(device)
 * (TransducerModel.py:24-58, ModelHander.py:38-82, predict_gpu.py:58-68): packs the
 * weights into the kernels' layout and uploads them to `device`.
 * Requires hidden == 128, n_base == 5, n_rle == 11, 1 <= image_features <= 256. */
int hb_create(const hb_weights *weights, int image_features, int hidden, int n_base,
              int n_rle, int device, hb_handle **out);
void hb_destroy(hb_handle *handle);

int hb_set_engine(hb_handle *handle, int engine);
int hb_get_engine(const hb_handle *handle);

/* Bytes of device scratch the calls below need for a batch of B windows of T columns
 * (chunk width W).  The caller allocates it once (torch.empty) and passes it in. */
int hb_workspace_bytes(const hb_handle *handle, int64_t B, int T, int W, size_t *out);

/* Replaces the whole per-batch body of predict() (predict_gpu.py:97-159 ==
 * predict.py:90-154): float cast, zero hidden, the chunk loop over
 * range(0, T, J) with W-column chunks, TransducerGRU.forward per chunk with hidden
 * carry, softmax, pad+add, and the first-index argmax.
 *   images_dev      uint8 [B, T, F] row-major (dataloader_predict.py:69 dtype)
 *   base_labels_dev uint8 [B, T]   rle_labels_dev uint8 [B, T]  (DataStore.py:129-133 dtype)
 *   base_prob_dev   float [B, T, 5] or NULL; rle_prob_dev float [B, T, 11] or NULL
 *                   (the accumulated softmax sums of predict.py:150-151, for parity tests)
 * T < W yields zero chunks: labels are 0 and probabilities 0, as in the reference loop. */
int hb_predict_windows(hb_handle *handle, const uint8_t *images_dev, int64_t B, int T,
                       int W, int J, uint8_t *base_labels_dev, uint8_t *rle_labels_dev,
                       float *base_prob_dev, float *rle_prob_dev, void *workspace_dev,
                       size_t workspace_bytes, void *stream);

/* Same contract with HOST buffers: copies the images to the device, runs
 * hb_predict_windows, copies both label arrays back and synchronises before returning
 * (this is the boundary predict() crosses with `images.to(device)` / `.cpu()`,
 * predict_gpu.py:125,152-153).  Uses internal pinned staging + device buffers that
 * grow on demand and are owned by the handle.  Probabilities are optional (may be NULL). */
int hb_predict_windows_host(hb_handle *handle, const uint8_t *images_host, int64_t B,
                            int T, int W, int J, uint8_t *base_labels_host,
                            uint8_t *rle_labels_host, float *base_prob_host,
                            float *rle_prob_host);

/* Replaces TransducerGRU.forward (TransducerModel.py:60-79) for one chunk:
 *   x_dev float [B, W, F], h_in_dev float [B, 2, H]  ->
 *   base_logits_dev float [B, W, 5], rle_logits_dev float [B, W, 11], h_out_dev [B, 2, H].
 * Always runs the fp32 engine. */
int hb_forward_chunk(hb_handle *handle, const float *x_dev, const float *h_in_dev,
                     int64_t B, int W, float *base_logits_dev, float *rle_logits_dev,
                     float *h_out_dev, void *workspace_dev, size_t workspace_bytes,
                     void *stream);

/* Replaces the per-chunk body of the training loop, helen/modules/python/models/train.py:189-201:
 *   output_base, output_rle, hidden = transducer_model(image_chunk, hidden)              (TransducerModel.py:60-79)
 *   loss = CrossEntropyLoss()(output_base, label_base) + CrossEntropyLoss(weight=CLASS_WEIGHTS)(output_rle, label_rle)
 *   loss.backward()
 * for one chunk: forward with the activations kept, both losses (mean reduction; the run-length loss is the weighted
 * mean torch computes, Options.py:29), and back-propagation through heads, decoder and encoder (both directions).
 * `weights_dev` / `grads_dev` hold DEVICE pointers in the hb_weights layout: the caller's own parameter tensors and the
 * tensors that receive d loss / d parameter (overwritten, not accumulated) -- no copy of the parameters is kept, so an
 * optimizer may update them between calls (train.py:202).  The initial state gets no gradient (train.py:206 detaches it).
 *   x_dev float [B, W, F]; h_in_dev float [B, 2, H] or NULL (zeros); labels int64 [B, W]; rle_class_weights_dev float [n_rle]
 *   loss_dev float [3] = {loss, loss_base, loss_rle}; h_out_dev float [B, 2, H];
 *   base_logits_dev float [B, W, 5] / rle_logits_dev float [B, W, 11] or NULL.
 * fp32 on the FMA pipes (gradients match autograd to ~1e-5 relative); image_features comes from the handle. */
int hb_train_workspace_bytes(const hb_handle *handle, int64_t B, int W, size_t *out);
int hb_train_step_chunk(hb_handle *handle, const hb_weights *weights_dev, const hb_weights *grads_dev,
                        const float *x_dev, const float *h_in_dev, const int64_t *label_base_dev,
                        const int64_t *label_rle_dev, const float *rle_class_weights_dev, int64_t B, int W,
                        float *loss_dev, float *h_out_dev, float *base_logits_dev, float *rle_logits_dev,
                        void *workspace_dev, size_t workspace_bytes, void *stream);

/* Number of kernel launches issued by this handle since creation (bench.py reports the
 * per-step delta as "gpu_launches"). */
int64_t hb_launch_count(const hb_handle *handle);

/* Device timing of the dominant kernel: when enabled, hb_predict_windows brackets its
 * kernels with CUDA events on the launching stream; hb_kernel_time_ms returns the
 * accumulated milliseconds and launch count since the last reset (synchronises). */
int hb_enable_kernel_timing(hb_handle *handle, int enable);
int hb_kernel_time_ms(hb_handle *handle, double *total_ms, int64_t *launches, int reset);

/* How the last hb_predict_windows call of the tensor engine was laid out on the chip (reporting only).
 * chunkloop = 1: the whole chunk loop of the batch ran as ONE launch of tc_chunkloop_kernel (windows_per_cta = 8,
 * batches of up to 320 windows) or tc_chunkloop2_kernel (windows_per_cta = 16 as two 8-window tiles, up to 512) with
 * 2 * recurrence_ctas recurrence CTAs, 6 * projection_workers projection CTAs and heads_workers heads
 * CTAs, all resident at once; chunkloop = 0: four launches per chunk (the batch does not fit on the chip). */
typedef struct hb_launch_plan {
    int chunkloop;
    int windows_per_cta;     /* live windows per recurrence CTA: 8, 16 or 32 */
    int stacked_operand;     /* 1: [h_hi | h_lo] stacked B operand (48 MMAs per step), 0: 3-term (72) */
    int recurrence_ctas;     /* per direction */
    int projection_workers;
    int heads_workers;
    int cooperative;         /* 1: the chunk-loop kernel was a cooperative launch (co-residency guaranteed by the driver) */
    int launches;            /* kernel launches the call issued */
} hb_launch_plan;
int hb_last_launch_plan(const hb_handle *handle, hb_launch_plan *out);

/* With hb_enable_kernel_timing(handle, 2) every launch of the dominant kernel is bracketed on its own:
 * tc_chunkloop_kernel (one launch = the whole chunk loop of the batch) when hb_last_launch_plan reports
 * chunkloop = 1, else the per-chunk GRU recurrence kernel (one launch = B windows x W dependent steps x
 * 2 directions of one layer);
 * returns the accumulated milliseconds and the launch count since the last reset (synchronises).
 * This mode disables launch overlap between kernels, so it is for measurement passes only. */
int hb_dominant_kernel_time_ms(hb_handle *handle, double *total_ms, int64_t *launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* HELEN_B200_H */
