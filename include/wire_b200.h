/* wire_b200.h — C ABI of the B200-native WIRE hot path.
 *
 * Drop-in boundary for the forward/backward pass through the complex Gabor stack of
 * Annatk26/wire.  The reference has no native code, so there is no FFI to mirror: these entry
 * points replace the *PyTorch op sequences* below (file:line relative to the reference checkout)
 * and are what a maintainer binds from Python with ctypes:
 *
 *   wire_net_forward          <- wire.INR.forward          modules/wire.py:161-165
 *                                wire2d.INR.forward        modules/wire2d.py:121-125
 *   wire_net_backward         <- loss.backward() through the same modules (PyTorch complex autograd;
 *                                closed form in SURVEY.md appendix A.2)
 *   wire_gabor_layer_forward  <- ComplexGaborLayer.forward   modules/wire.py:88-93
 *                                ComplexGaborLayer2D.forward modules/wire2d.py:56-67
 *   wire_gabor_layer_backward <- autograd of the above (used by model.net[i](x), modules/utils.py:251-252)
 *   wire_final_linear_*       <- nn.Linear(M, out, dtype=cfloat) + .real   modules/wire.py:156-165
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (torch-allocated in the Python host); nothing is allocated, freed or retained;
 *   - complex tensors are interleaved float pairs (torch.view_as_real of a contiguous complex64),
 *     weights are [out, in] row-major exactly as nn.Linear stores them;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and never synchronise;
 *   - every function returns 0 on success, non-zero on failure (wire_b200_last_error() explains);
 *   - there is NO CPU fallback: on a device that is not sm_100 the calls fail.
 */
#ifndef WIRE_B200_H
#define WIRE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WIRE_B200_ABI_VERSION 2
#define WIRE_B200_MAX_LAYERS 16 /* first layer + hidden layers */

/* TF32    : GEMM operands TF32 (10-bit mantissa, fp32 storage), FP32 accumulate, saved pre-activations FP16.
 * FP32    : FP32 FMAs on the CUDA cores (precision yardstick; same GPU, not a fallback).
 * MIXED16 : whole-network path with 16-bit tensors in HBM -- activations and forward weights FP16 (the same 11-bit
 *           significand as TF32), gradients and dgrad/wgrad operands BF16 (FP32's exponent range, so no loss scaling) --
 *           FP32 accumulate in TMEM; half the activation traffic and twice the tensor-core rate of TF32.  The whole-network
 *           calls fall back to the TF32 kernels for shapes it does not cover (in_features > 3, out_features > 4); the
 *           single-layer entry points run the 16-bit kernels for hidden layers (caller tensors stay fp32 and are converted on
 *           the way in and out) and FP32 / TF32 kernels for the first layer and the stand-alone final Linear. */
enum { WIRE_PRECISION_TF32 = 0, WIRE_PRECISION_FP32 = 1, WIRE_PRECISION_MIXED16 = 2 };

typedef struct wire_net_desc {
  int32_t two_d;         /* 0 = wire (modules/wire.py), 1 = wire2d (modules/wire2d.py) */
  int32_t in_features;   /* coordinate dimensions (1..8) */
  int32_t width;         /* M: complex hidden features (already int(h/sqrt2) or h/2) */
  int32_t hidden_layers; /* H: complex Gabor layers after the first one */
  int32_t out_features;
  int32_t precision;     /* WIRE_PRECISION_* */
} wire_net_desc;

typedef struct wire_layer_params {
  const float* weight;  /* first layer: real [M][in]; hidden: complex [M][M] interleaved */
  const float* bias;    /* first layer: real [M];     hidden: complex [M] interleaved   */
  const float* weight2; /* wire2d scale_orth.weight (NULL for wire) */
  const float* bias2;   /* wire2d scale_orth.bias */
  const float* omega0;  /* device scalar, the layer's omega_0 parameter (f32[1]) */
  const float* scale0;  /* device scalar, the layer's scale_0 parameter (f32[1]) */
} wire_layer_params;

typedef struct wire_net_params {
  wire_layer_params layer[WIRE_B200_MAX_LAYERS]; /* [0] first, [1..H] hidden */
  const float* final_weight;                     /* complex [out][M] interleaved */
  const float* final_bias;                       /* complex [out] interleaved */
} wire_net_params;

typedef struct wire_layer_grads {
  float* weight;
  float* bias;
  float* weight2;
  float* bias2;
  /* trainable=True of ComplexGaborLayer(2D) (modules/wire.py:66,80-81): gradients of the layer's own omega_0 / scale_0, one float
   * each, accumulated inside the fused backward kernels of wire_net_backward (MIXED16 only); NULL = the scalars are constants.
   * Ignored by the single-layer entry points (wire_gabor_scalar_grads serves those). */
  float* omega0;
  float* scale0;
} wire_layer_grads;

/* How wire_net_backward clears the accumulation targets before its kernels add into them (split-K partial sums are
 * accumulated with fp32 atomics, so every slot must start at zero):
 *   WIRE_GRADS_CLEAR_SLOTS  one memset per slot (default; the slots may live anywhere)
 *   WIRE_GRADS_CLEAR_FLAT   every slot lies inside [flat_base, flat_base + flat_floats): ONE memset of that range
 *   WIRE_GRADS_PREZEROED    the caller guarantees the slots are zero on entry (e.g. wire_adam_step_dev(..., zero_grad=1) cleared
 *                           them while consuming the previous step's gradients): no memset at all */
enum { WIRE_GRADS_CLEAR_SLOTS = 0, WIRE_GRADS_CLEAR_FLAT = 1, WIRE_GRADS_PREZEROED = 2 };

typedef struct wire_net_grads {
  wire_layer_grads layer[WIRE_B200_MAX_LAYERS];
  float* final_weight;
  float* final_bias;
  int32_t clear_mode;  /* WIRE_GRADS_* */
  float* flat_base;    /* WIRE_GRADS_CLEAR_FLAT: the flat gradient buffer that contains every slot */
  size_t flat_floats;
} wire_net_grads;

int wire_b200_abi_version(void);
const char* wire_b200_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x) */
int wire_b200_device_ok(void);
int wire_b200_sm_count(void);
/* rows per pass of an inference (training == 0) forward: workspaces are sized for min(n, this) rows */
int64_t wire_b200_infer_chunk_rows(void);

/* ---- launch accounting / per-kernel device timing (bench.py) --------------------------- */
/* Every kernel launch is counted per kind. With timing != 0 each launch is also bracketed by CUDA
 * events recorded on the launching stream; wire_b200_prof_get() resolves them. */
int wire_b200_prof_enable(int32_t timing);
int wire_b200_prof_reset(void);
int wire_b200_prof_kinds(void);
const char* wire_b200_prof_name(int32_t kind);
int wire_b200_prof_get(int32_t kind, uint64_t* launches, double* ms);

/* ---- whole network ------------------------------------------------------------------- */

/* Scratch needed for `n` coordinates. training != 0 keeps what backward needs. */
size_t wire_net_workspace_bytes(const wire_net_desc* d, int64_t n, int32_t training);
/* One-time initialisation of a freshly allocated workspace (zero padding, "ones" columns). */
int wire_net_workspace_init(const wire_net_desc* d, int64_t n, int32_t training, void* workspace,
                            size_t workspace_bytes, void* stream);
/* out[n][out_features] = Re(final(gabor...(first(coords)))) ; coords is [n][in_features] f32 */
int wire_net_forward(const wire_net_desc* d, const wire_net_params* p, const float* coords,
                     int64_t n, float* out, void* workspace, size_t workspace_bytes,
                     int32_t training, void* stream);
/* Gradients of sum(out * grad_out) w.r.t. every parameter (OVERWRITES *grads; complex gradients
 * in PyTorch's convention dL/dRe + j dL/dIm) and optionally w.r.t. coords (grad_coords may be NULL).
 * Must follow a wire_net_forward(training=1) on the same workspace, coords and params. */
int wire_net_backward(const wire_net_desc* d, const wire_net_params* p, const float* coords,
                      int64_t n, const float* grad_out, void* workspace, size_t workspace_bytes,
                      const wire_net_grads* grads, float* grad_coords, void* stream);

/* Inspection of a TRAINING workspace (per-layer parity tests of the fused path; saved activations for diagnostics): copies one
 * tensor out as dense fp32, converting from the storage type of the precision mode (FP16 activations / pre-activations, BF16
 * gradients under MIXED16).  which / index:
 *   WIRE_WS_Y   y_index, the output of layer `index` = the input of layer index+1 (0..H-1; H too when the final Linear is not fused)
 *   WIRE_WS_Z   z_index = x W^T + b of hidden layer `index` (1..H)        WIRE_WS_W  the scale_orth pre-activation (wire2d)
 *   WIRE_WS_GZ  gradient buffer `index` (0..1): after wire_net_backward, g_z of hidden layer l (dL/dRe z + j dL/dIm z) is in
 *               buffer (H - l) & 1 -- the two buffers alternate, so only the last two layers written (l = 1, 2) survive
 *   WIRE_WS_GW  same for the scale_orth branch (wire2d)
 *   WIRE_WS_GZ0 real g_z of the first layer [n][M]                         WIRE_WS_GW0 same for its scale_orth branch
 * out: [n][2M] floats (interleaved re, im), or [n][M] for GZ0 / GW0. */
enum { WIRE_WS_Y = 0, WIRE_WS_Z = 1, WIRE_WS_W = 2, WIRE_WS_GZ = 3, WIRE_WS_GW = 4, WIRE_WS_GZ0 = 5, WIRE_WS_GW0 = 6 };
int wire_net_workspace_read(const wire_net_desc* d, int64_t n, const void* workspace, size_t workspace_bytes, int32_t which,
                            int32_t index, float* out, void* stream);

/* wire_net_backward with the MSE loss of the training loops fused into the top of the backward pass
 * (loss = ((pixelvalues - gt)**2).mean(); loss.backward(): wire_image_denoise.py:153-156, wire_occupancy.py:149-153):
 * grad_out is not an input — it is computed on the fly as 2 (pred - target) / count_global, where pred is the output
 * wire_net_forward(training=1) wrote for the same coords and count_global is the element count the mean is taken over
 * (n * out_features on one GPU; the whole batch's count when this rank holds a shard).  The loss is added to the device
 * ring exactly as by wire_mse_loss_grad_ring.  grad_out_scratch ([n][out_features]) is used only by the precisions whose
 * kernels cannot compute the gradient themselves. */
int wire_net_backward_mse(const wire_net_desc* d, const wire_net_params* p, const float* coords, int64_t n, const float* pred,
                          const float* target, int64_t count_global, float* loss_ring, int32_t ring_n, const int64_t* step_dev,
                          float* grad_out_scratch, void* workspace, size_t workspace_bytes, const wire_net_grads* grads,
                          float* grad_coords, void* stream);

/* ---- single layers (model.net[i](x) and per-layer parity) ---------------------------- */

size_t wire_gabor_layer_workspace_bytes(const wire_net_desc* d, int32_t is_first, int32_t in_features,
                                        int64_t n);
/* x: first layer real [n][in_features], else complex [n][in_features]; y: complex [n][width].
 * z_save / w_save (complex [n][width]; first layer: real [n][width]) may be NULL (no_grad). */
int wire_gabor_layer_forward(const wire_net_desc* d, int32_t is_first, int32_t in_features,
                             const wire_layer_params* p, const float* x, int64_t n, float* y,
                             float* z_save, float* w_save, void* workspace, size_t workspace_bytes,
                             void* stream);
/* grad_x may be NULL. grads are OVERWRITTEN. */
int wire_gabor_layer_backward(const wire_net_desc* d, int32_t is_first, int32_t in_features,
                              const wire_layer_params* p, const float* x, const float* z_save,
                              const float* w_save, const float* grad_y, int64_t n, float* grad_x,
                              const wire_layer_grads* grads, void* workspace, size_t workspace_bytes,
                              void* stream);
/* out[n][out] = Re(h Wf^T + bf) ; h complex [n][width] */
int wire_final_linear_forward(const wire_net_desc* d, const float* weight, const float* bias,
                              const float* h, int64_t n, float* out, void* stream);
/* grad_h complex [n][width] (may be NULL); grad_weight/grad_bias OVERWRITTEN */
int wire_final_linear_backward(const wire_net_desc* d, const float* weight, const float* h,
                               const float* grad_out, int64_t n, float* grad_h, float* grad_weight,
                               float* grad_bias, void* stream);

/* ---- fused optimiser step (torch.optim.Adam semantics on view_as_real parameters) ------ */
/* p -= lr * m_hat / (sqrt(v_hat) + eps), over `count` floats; `step` is the 1-based step index. */
int wire_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t count,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                   float grad_scale, void* stream);
/* Same update with the step counter (*step_dev, 0-based count of completed steps, incremented by the kernel) and the
 * learning rate (*lr_dev) on the device, so a captured CUDA graph of a whole training step can be replayed.
 * scratch_dev: one zero-initialised uint32 used by the kernel to detect its last block.
 * zero_grad != 0: every gradient element is set to zero after it has been read (optim.zero_grad() folded into the step), so the
 * next wire_net_backward can run with WIRE_GRADS_PREZEROED and the training step contains no memset. */
int wire_adam_step_dev(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t count,
                       const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                       int64_t* step_dev, float grad_scale, uint32_t* scratch_dev, int32_t zero_grad, void* stream);
/* grad_out[i] = 2*(pred[i]-target[i])/count ; *loss (device scalar, accumulated) += mean sq err */
int wire_mse_loss_grad(const float* pred, const float* target, int64_t count, float* grad_out,
                       float* loss, void* stream);
/* Same for a rank that holds `count` elements of a global batch of `count_global`: grad_out = 2*(pred-target)/count_global,
 * *loss += sum sq err / count_global, so gradients and losses of the ranks ADD UP to those of the global mean
 * (criterion = MSELoss over the whole chunk, wire_occupancy.py:149) whatever the shard sizes. */
int wire_mse_loss_grad_n(const float* pred, const float* target, int64_t count, int64_t count_global, float* grad_out,
                         float* loss, void* stream);
/* Same, with the loss kept in a ring on the device: the loss of training step s = *step_dev (the optimiser-step counter of
 * wire_adam_step_dev / _peer) is accumulated into loss_ring[s % ring_n] and slot (s+1) % ring_n is cleared for the next
 * step; the host may read a step's loss at any time during the following ring_n - 1 steps (the reference reads
 * loss.item() every chunk, wire_occupancy.py:156 — here without a reset kernel and without stalling the stream).
 * loss_ring must start zeroed. */
int wire_mse_loss_grad_ring(const float* pred, const float* target, int64_t count, int64_t count_global, float* grad_out,
                            float* loss_ring, int32_t ring_n, const int64_t* step_dev, void* stream);

/* ---- data-parallel gradient exchange fused with Adam, over NVLink peer memory ----------------
 * The reference is single-GPU; its chunked loops (wire_occupancy.py:137-154) shard by coordinate batch, and the only
 * exchange is a SUM of the flat weight gradient.  Every rank allocates one peer buffer
 *     [ wire_peer_header_bytes() of barrier counters | grad_floats floats ]
 * with wire_peer_alloc (cudaMalloc + cudaIpc handle), exchanges the 64-byte handles through any host channel, maps the
 * others with wire_peer_open, and points wire_net_backward's gradient slots into the float area of its OWN buffer.
 * wire_adam_step_peer then sums all ranks' gradients with P2P loads (rank order, so replicas stay bit-identical) and applies
 * torch.optim.Adam's update; wire_peer_wait_done must be enqueued before the next kernel that overwrites the gradient
 * area (it waits until every peer has finished reading).  peer_bases: host array of `world` pointers, [rank] = own buffer.
 * All ranks must call in lockstep; no NCCL call is involved and the kernels can be captured in a CUDA graph. */
#define WIRE_B200_IPC_HANDLE_BYTES 64
#define WIRE_B200_MAX_PEERS 16
size_t wire_peer_header_bytes(void);
int wire_peer_alloc(size_t grad_floats, void** base, void* ipc_handle /* WIRE_B200_IPC_HANDLE_BYTES out */);
int wire_peer_open(const void* ipc_handle, void** base);
int wire_peer_close(void* base);
int wire_peer_free(void* base);
int wire_adam_step_peer(float* param, void* const* peer_bases, int32_t world, int32_t rank, float* exp_avg, float* exp_avg_sq,
                        int64_t count, const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                        int64_t* step_dev, float grad_scale, uint32_t* scratch_dev, void* stream);
int wire_peer_wait_done(void* const* peer_bases, int32_t world, int32_t rank, const int64_t* step_dev, void* stream);

/* ---- on-device coordinate pipeline and metrics ----------------------------------
 * Replace the host-side batch assembly of the reference's loops: CPU randperm + CPU gather + .cuda() per chunk
 * (wire_image_denoise.py:142-147, wire_occupancy.py:137-144), rec[b_indices] = pixelvalues (wire_image_denoise.py:150-151,
 * wire_occupancy.py:146-147) and the per-epoch metrics (modules/volutils.py:74-91, modules/utils.py:67-82).
 *
 * wire_grid_batch: for r < n, li = idx ? idx[r] : idx_base + r is a linear index into the (H, W[, T]) grid in the
 *   reference's order (np.meshgrid 'xy': li = (i*W + j)*T + k -> coordinate (x_j, y_i, z_k)); coords[r] (n x ndim f32, may
 *   be NULL) is that grid point with linspace_kind 0 = np.linspace float64 cast to f32 (utils.get_coords,
 *   modules/utils.py:163-176) or 1 = torch.linspace float32 (wire_image_denoise.py:63-66), bit-exact; target[r]
 *   (n x out_features, may be NULL) = signal[li].  An out-of-range index sets *err_flag (device int, may be NULL) to 1.
 * wire_scatter_rows: dst[li] = src[r] for rows of `width` floats.
 * wire_iou_counts: counts[0] += |pred AND gt|, counts[1] += |pred OR gt| after thresholding pred at `thres` when use_thres
 *   (binarize_in_place also writes the 0/1 values back, which is what the reference does to its argument).
 * wire_sq_err_stats: stats[0] += sum (x - xhat)^2, stats[1] = max(stats[1], max x)  (float64 on the device). */
int wire_grid_batch(const int32_t* dims, int32_t ndim, int32_t linspace_kind, const int64_t* idx, int64_t idx_base, int64_t n,
                    const float* signal, int32_t out_features, float* coords, float* target, int32_t* err_flag, void* stream);
int wire_scatter_rows(const int64_t* idx, int64_t idx_base, int64_t n, const float* src, int32_t width, float* dst,
                      int64_t dst_rows, int32_t* err_flag, void* stream);
int wire_iou_counts(float* preds, const float* gt, int64_t count, float thres, int32_t use_thres, int32_t binarize_in_place,
                    uint64_t* counts, void* stream);
int wire_sq_err_stats(const float* x, const float* xhat, int64_t count, double* stats, void* stream);
/* wire_avgpool_mse_loss_grad: the super-resolution loss of wire_SISR.py:154-161 — pred is the HR prediction [H*W][channels],
 *   target_lr the LR image [(H/scale)*(W/scale)][channels]; loss = mean((target_lr - AvgPool2d(scale)(pred))^2) is ADDED to
 *   *loss (device scalar, may be NULL) and grad_out [H*W][channels] receives d loss / d pred. */
int wire_avgpool_mse_loss_grad(const float* pred, const float* target_lr, int32_t H, int32_t W, int32_t channels, int32_t scale,
                               float* grad_out, float* loss, void* stream);

/* Radon forward operator of the CT driver and its adjoint (modules/lin_inverse.py:19-40 `radon`, called every iteration by
 * wire_ct.py:126-128): image [nimg][H][W] -> sinogram [nangles][nimg][W], each angle (degrees) rotating the image about its centre
 * like kornia.geometry.rotate (bilinear, zeros outside, align_corners=True) and summing over the rows.  wire_radon_backward
 * OVERWRITES grad_image [nimg][H][W] with the adjoint applied to grad_sinogram (autograd of the above).  kornia is not part of
 * the build image: the rotation convention is restated from its published source, parity against kornia itself is unpinned. */
int wire_radon_forward(const float* image, int32_t nimg, int32_t H, int32_t W, const float* angles_deg, int32_t nangles, float* sinogram,
                       void* stream);
int wire_radon_backward(const float* grad_sinogram, int32_t nimg, int32_t H, int32_t W, const float* angles_deg, int32_t nangles,
                        float* grad_image, void* stream);

/* Gradients of a layer's own omega_0 / scale_0 (trainable=True of ComplexGaborLayer(2D), modules/wire.py:66,80-81,
 * modules/wire2d.py:27,42-43; autograd of wire.py:88-93 w.r.t. the two scalars): out2[0] += sum Im(conj(z) p),
 * out2[1] += -2 s0 sum (|z|^2 + |w|^2) Re p, p = conj(y) grad_y, from the tensors wire_gabor_layer_forward saved.
 * out2: two float64 on the device, accumulated. */
int wire_gabor_scalar_grads(int32_t is_first, int32_t two_d, int32_t width, const float* z_save, const float* w_save,
                            const float* grad_y, int64_t n, const float* omega0, const float* scale0, double* out2, void* stream);

/* RealGaborLayer (modules/wire.py:6-42; not used by INR): the fused activation y = cos(omega_0 f) exp(-(scale_0 s)^2) of the
 * two real Linears' outputs f = freqs(x), s = scale(x) (modules/wire.py:38-42), and its derivative w.r.t. f and s. */
int wire_real_gabor_forward(const float* f, const float* s, int64_t count, float omega0, float scale0, float* y, void* stream);
/* The whole RealGaborLayer.forward (modules/wire.py:29-42) and its autograd on FP32 FMAs: x [n][K], weights [M][K] row-major as
 * nn.Linear stores them, biases [M] (may be NULL).  forward: y [n][M]; f_save / s_save [n][M] = the two Linears' outputs, kept for
 * the backward pass (may be NULL under no_grad).  backward: OVERWRITES the four parameter gradients (g_b_* may be NULL) and, when
 * non-NULL, grad_x [n][K]; scratch_gf / scratch_gs are [n][M] work buffers. */
int wire_real_gabor_layer_forward(const float* x, int64_t n, int32_t K, int32_t M, const float* w_freqs, const float* b_freqs,
                                  const float* w_scale, const float* b_scale, float omega0, float scale0, float* y, float* f_save,
                                  float* s_save, void* stream);
int wire_real_gabor_layer_backward(const float* x, const float* f_save, const float* s_save, const float* grad_y, int64_t n, int32_t K,
                                   int32_t M, const float* w_freqs, const float* w_scale, float omega0, float scale0, float* grad_x,
                                   float* g_w_freqs, float* g_b_freqs, float* g_w_scale, float* g_b_scale, float* scratch_gf,
                                   float* scratch_gs, void* stream);
int wire_real_gabor_backward(const float* f, const float* s, const float* grad_y, int64_t count, float omega0, float scale0,
                             float* grad_f, float* grad_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WIRE_B200_H */
