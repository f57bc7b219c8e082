// data_kernels.cuh — the on-device coordinate pipeline either side of the WIRE hot path.
//
// The reference's training loops build every batch on the HOST: a CPU randperm, a CPU gather of the coordinate rows, a
// host->device copy per chunk (wire_image_denoise.py:142-147, wire_occupancy.py:137-144), then scatter the prediction back
// (rec[:, b_indices] = pixelvalues, wire_image_denoise.py:150-151) and compute PSNR / IoU with more torch ops and host
// syncs (modules/utils.py:67-82, modules/volutils.py:74-91).  Here a batch is assembled from LINEAR INDICES on the device:
// the coordinate of index i is generated, not gathered (the grid is a linspace product), the target row is gathered from the
// resident signal, and the metrics are single-pass reductions.  All integer / index work is bit-exact with the reference;
// the coordinates reproduce numpy's float64 linspace (modules/utils.py:163-176) or torch's float32 CPU linspace
// (wire_image_denoise.py:63-66) bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace wire {

struct GridSpec {
  int ndim;      // 2 or 3
  int dims[3];   // H, W, T  (np.meshgrid 'xy' order: flat index = (i*W + j)*T + k  ->  coordinate (x_j, y_i, z_k))
  int linspace;  // 0 = numpy float64 linspace cast to f32 (utils.get_coords), 1 = torch float32 CPU linspace (the image drivers)
};

// np.linspace(-1, 1, n)[j] cast to float32: y = j*step + start in float64 (two roundings, no FMA), last sample == stop
__device__ __forceinline__ float linspace_np(int j, int n) {
  if (n == 1) return -1.0f;
  if (j == n - 1) return 1.0f;
  const double step = 2.0 / double(n - 1);
  return float(__dadd_rn(__dmul_rn(double(j), step), -1.0));
}
// torch.linspace(-1, 1, n)[j] (float32, CPU kernel): step = 2/(n-1) in f32; first half start + step*j, second half
// end - step*(n-1-j), each with a single rounding (verified against torch for n = 2..1024)
__device__ __forceinline__ float linspace_torch(int j, int n) {
  if (n == 1) return -1.0f;
  const float step = __fdiv_rn(2.0f, float(n - 1));
  if (j < n / 2) return __fmaf_rn(step, float(j), -1.0f);
  return __fmaf_rn(-step, float(n - 1 - j), 1.0f);
}
__device__ __forceinline__ float grid_value(int kind, int j, int n) { return kind ? linspace_torch(j, n) : linspace_np(j, n); }

// coords[r] = grid coordinate of linear index idx[r] (idx == nullptr: r + idx_base), target[r] = signal[idx[r]] (optional)
__global__ void grid_batch_kernel(GridSpec g, const int64_t* __restrict__ idx, int64_t idx_base, int64_t n,
                                  const float* __restrict__ signal, int out_features, float* __restrict__ coords,
                                  float* __restrict__ target, int* __restrict__ err) {
  const int64_t total = g.ndim == 3 ? int64_t(g.dims[0]) * g.dims[1] * g.dims[2] : int64_t(g.dims[0]) * g.dims[1];
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x) {
    int64_t li = idx ? idx[r] : idx_base + r;
    if (li < 0 || li >= total) { if (err) atomicExch(err, 1); li = 0; }
    if (coords) {
      if (g.ndim == 3) {
        const int k = int(li % g.dims[2]);
        const int64_t ij = li / g.dims[2];
        const int j = int(ij % g.dims[1]), i = int(ij / g.dims[1]);
        coords[3 * r] = grid_value(g.linspace, j, g.dims[1]);
        coords[3 * r + 1] = grid_value(g.linspace, i, g.dims[0]);
        coords[3 * r + 2] = grid_value(g.linspace, k, g.dims[2]);
      } else {
        const int j = int(li % g.dims[1]), i = int(li / g.dims[1]);
        coords[2 * r] = grid_value(g.linspace, j, g.dims[1]);
        coords[2 * r + 1] = grid_value(g.linspace, i, g.dims[0]);
      }
    }
    if (target) {
      for (int o = 0; o < out_features; ++o) target[r * out_features + o] = __ldg(signal + li * out_features + o);
    }
  }
}

// dst[idx[r]] = src[r] (rows of `width` floats); idx == nullptr: dst[idx_base + r]
__global__ void scatter_rows_kernel(const int64_t* __restrict__ idx, int64_t idx_base, int64_t n, const float* __restrict__ src,
                                    int width, float* __restrict__ dst, int64_t dst_rows, int* __restrict__ err) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n * width; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = t / width;
    const int o = int(t - r * width);
    const int64_t li = idx ? idx[r] : idx_base + r;
    if (li < 0 || li >= dst_rows) { if (err) atomicExch(err, 1); continue; }
    dst[li * width + o] = src[t];
  }
}

// volutils.get_I_and_U (modules/volutils.py:79-91): predictions are thresholded (IN PLACE in the reference; here only when
// `binarize` is set) and intersection / union are counts of logical_and / logical_or with the ground truth.
// counts[0] += intersection, counts[1] += union   (exact integer counts)
__global__ void iou_counts_kernel(float* __restrict__ preds, const float* __restrict__ gt, int64_t count, float thres, int use_thres,
                                  int binarize, unsigned long long* __restrict__ counts) {
  unsigned int inter = 0, uni = 0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    float p = preds[i];
    if (use_thres) {
      if (p < thres) p = 0.0f;   // the reference's two masked assignments, in order (so thres <= 0 maps everything to 1,
      if (p >= thres) p = 1.0f;  // and a NaN stays NaN)
      if (binarize) preds[i] = p;
    }
    const bool pb = p != 0.0f, gb = gt[i] != 0.0f;  // logical_and / logical_or: non-zero (NaN included) is true
    inter += (pb && gb);
    uni += (pb || gb);
  }
  for (int s = 16; s > 0; s >>= 1) {
    inter += __shfl_xor_sync(0xffffffffu, inter, s);
    uni += __shfl_xor_sync(0xffffffffu, uni, s);
  }
  __shared__ unsigned int si[32], su[32];
  if ((threadIdx.x & 31) == 0) { si[threadIdx.x >> 5] = inter; su[threadIdx.x >> 5] = uni; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0, b = 0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) { a += si[w]; b += su[w]; }
    if (a) atomicAdd(counts, a);
    if (b) atomicAdd(counts + 1, b);
  }
}

// stats[0] += sum (x - xhat)^2 (float64), stats[1] = max(stats[1], max x): utils.psnr = 10 log10(max(x) / mean err^2)
// (modules/utils.py:67-82); the drivers' own -10 log10(mse) (wire_image_denoise.py:161-167) needs stats[0] only.
__global__ void sq_err_stats_kernel(const float* __restrict__ x, const float* __restrict__ xhat, int64_t count, double* __restrict__ stats) {
  double acc = 0.0;
  float mx = -INFINITY;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    const float a = x[i];
    const float d = a - xhat[i];
    acc += double(d) * double(d);
    mx = fmaxf(mx, a);
  }
  for (int s = 16; s > 0; s >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, s);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  }
  __shared__ double sa[32];
  __shared__ float sm[32];
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = acc; sm[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    float m = -INFINITY;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) { a += sa[w]; m = fmaxf(m, sm[w]); }
    atomicAdd(stats, a);
    // double max through atomicCAS on the bit pattern
    unsigned long long* addr = reinterpret_cast<unsigned long long*>(stats + 1);
    unsigned long long old = *addr, assumed;
    do {
      assumed = old;
      if (__longlong_as_double((long long)assumed) >= double(m)) break;
      old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(double(m)));
    } while (assumed != old);
  }
}

// The super-resolution loss of wire_SISR.py:154-161: the HR prediction [H*W, C] is average-pooled by `scale`
// (torch.nn.AvgPool2d(scale): stride = scale, floor mode, so H2 = H / scale, W2 = W / scale and remainder rows / columns are
// ignored) and compared with the LR target [H2*W2, C]:  loss = mean((gt_lr - pool(rec_hr))^2).  One thread per LR element
// computes the pooled value, its squared error and writes the gradient of its scale x scale HR pixels:
// grad[hr, c] = 2 (pool - gt_lr) / (H2*W2*C) / scale^2.   (grad of ignored remainder pixels is zeroed by the caller.)
__global__ void avgpool_mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target, int H, int W, int C, int scale,
                                        float* __restrict__ grad, float* __restrict__ loss) {
  const int H2 = H / scale, W2 = W / scale;
  const int64_t count = int64_t(H2) * W2 * C;
  const float inv = 1.0f / float(count);
  const float inv_pool = 1.0f / float(scale * scale);
  float local = 0.f;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < count; t += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(t % C);
    const int64_t px = t / C;
    const int j = int(px % W2), i = int(px / W2);
    float sum = 0.f;
    for (int a = 0; a < scale; ++a)
      for (int b = 0; b < scale; ++b) sum += pred[(int64_t(i * scale + a) * W + (j * scale + b)) * C + c];
    const float d = sum * inv_pool - target[t];
    local = fmaf(d, d, local);
    const float gv = 2.0f * d * inv * inv_pool;
    for (int a = 0; a < scale; ++a)
      for (int b = 0; b < scale; ++b) grad[(int64_t(i * scale + a) * W + (j * scale + b)) * C + c] = gv;
  }
  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  __shared__ float ws[32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, v * inv);
  }
}

// Gradients of the layer's own scalars omega_0 / scale_0 when they are trainable (``trainable=True`` of
// ComplexGaborLayer / ComplexGaborLayer2D, modules/wire.py:66,80-81, modules/wire2d.py:27,42-43): with p = conj(y) g_y,
//   g_omega0 = sum Im(conj(z) p)          g_scale0 = -2 s0 sum (|z|^2 + |w|^2) Re(p)          
// y is recomputed from the saved pre-activation with the accurate libdevice functions; z (w) are complex [n][M] (real
// [n][M] for the first layer), g_y complex [n][M].  out[0] += g_omega0, out[1] += g_scale0 (float64 accumulators).
__global__ void gabor_scalar_grads_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ gy,
                                          int64_t count, int is_first, const float* __restrict__ omega_p,
                                          const float* __restrict__ scale_p, double* __restrict__ out) {
  const float omega = *omega_p, s0 = *scale_p;
  double a_om = 0.0, a_s = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    float zr, zi, wn = 0.f;
    if (is_first) { zr = z[i]; zi = 0.f; if (w) wn = w[i] * w[i]; }
    else { zr = z[2 * i]; zi = z[2 * i + 1]; if (w) wn = w[2 * i] * w[2 * i] + w[2 * i + 1] * w[2 * i + 1]; }
    const float t = zr * zr + zi * zi + wn;
    const float m = expf(-omega * zi - s0 * s0 * t);
    float sn, cs;
    sincosf(omega * zr, &sn, &cs);
    const float yr = m * cs, yi = m * sn;
    const float gr = gy[2 * i], gi = gy[2 * i + 1];
    const float pr = yr * gr + yi * gi, pi = yr * gi - yi * gr;   // p = conj(y) g
    a_om += double(zr * pi - zi * pr);                              // Im(conj(z) p)
    a_s += double(t * pr);
  }
  for (int s = 16; s > 0; s >>= 1) {
    a_om += __shfl_xor_sync(0xffffffffu, a_om, s);
    a_s += __shfl_xor_sync(0xffffffffu, a_s, s);
  }
  __shared__ double so[32], ss[32];
  if ((threadIdx.x & 31) == 0) { so[threadIdx.x >> 5] = a_om; ss[threadIdx.x >> 5] = a_s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0.0, y = 0.0;
    for (int k = 0; k < int(blockDim.x >> 5); ++k) { x += so[k]; y += ss[k]; }
    atomicAdd(out, x);
    atomicAdd(out + 1, -2.0 * double(s0) * y);
  }
}

// RealGaborLayer (modules/wire.py:6-42, not used by INR): y = cos(omega_0 f) exp(-(scale_0 s)^2) with f = freqs(x), s = scale(x)
// two real Linears (library GEMMs on the host side); this is the fused activation and its derivative:
//   g_f = -omega_0 sin(omega_0 f) exp(-(scale_0 s)^2) g_y        g_s = -2 scale_0^2 s y g_y
__global__ void real_gabor_fwd_kernel(const float* __restrict__ f, const float* __restrict__ sc, int64_t count, float omega, float s0,
                                      float* __restrict__ y) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    const float t = s0 * sc[i];
    y[i] = cosf(omega * f[i]) * expf(-(t * t));
  }
}
__global__ void real_gabor_bwd_kernel(const float* __restrict__ f, const float* __restrict__ sc, const float* __restrict__ gy, int64_t count,
                                      float omega, float s0, float* __restrict__ gf, float* __restrict__ gs) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    const float t = s0 * sc[i];
    const float e = expf(-(t * t));
    float sn, cs;
    sincosf(omega * f[i], &sn, &cs);
    const float g = gy[i];
    gf[i] = -omega * sn * e * g;
    gs[i] = -2.0f * s0 * s0 * sc[i] * cs * e * g;
  }
}

// ---------------------------------------------------------------------------------------------
// RealGaborLayer as a whole (modules/wire.py:29-42): both real Linears and the activation on FP32 FMAs.  The class is not
// part of any INR, so these are plain 64 x 64 shared-memory tiles (16 x 16 threads, 4 x 4 outputs each), not tcgen05 GEMMs.
//   forward : f = x Wf^T + bf, s = x Ws^T + bs, y = cos(omega f) exp(-(scale s)^2)       [n, K] x [M, K]^T -> [n, M]
//   dgrad   : g_x = g_f Wf + g_s Ws                                                        [n, M] x [M, K]   -> [n, K]
//   wgrad   : g_W = g^T x, g_b = sum_n g (split over n, fp32 atomics into zeroed buffers)  [n, M]^T x [n, K] -> [M, K]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) real_gabor_layer_fwd_kernel(const float* __restrict__ x, int n, int K, int M,
                                                                    const float* __restrict__ Wf, const float* __restrict__ bf,
                                                                    const float* __restrict__ Ws, const float* __restrict__ bs,
                                                                    float omega, float s0, float* __restrict__ y,
                                                                    float* __restrict__ f_save, float* __restrict__ s_save) {
  __shared__ float Xs[16][65], Fs[16][65], Ss[16][65];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  float af[4][4] = {}, as[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 256 + tid, rr = idx >> 4, kk = idx & 15;   // 64 rows (or features) x 16 k
      const bool kok = k0 + kk < K;
      Xs[kk][rr] = (kok && r0 + rr < n) ? x[size_t(r0 + rr) * K + k0 + kk] : 0.f;
      Fs[kk][rr] = (kok && j0 + rr < M) ? Wf[size_t(j0 + rr) * K + k0 + kk] : 0.f;
      Ss[kk][rr] = (kok && j0 + rr < M) ? Ws[size_t(j0 + rr) * K + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float xv[4], fv[4], sv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { xv[i] = Xs[kk][ty * 4 + i]; fv[i] = Fs[kk][tx * 4 + i]; sv[i] = Ss[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { af[i][j] = fmaf(xv[i], fv[j], af[i][j]); as[i][j] = fmaf(xv[i], sv[j], as[i][j]); }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = j0 + tx * 4 + j;
      if (c >= M) continue;
      const float f = af[i][j] + (bf ? __ldg(bf + c) : 0.f), sc = as[i][j] + (bs ? __ldg(bs + c) : 0.f);
      const float t = s0 * sc;
      y[size_t(r) * M + c] = cosf(omega * f) * expf(-(t * t));
      if (f_save) f_save[size_t(r) * M + c] = f;
      if (s_save) s_save[size_t(r) * M + c] = sc;
    }
  }
}
// g_x[n][k] = sum_j g_f[n][j] Wf[j][k] + g_s[n][j] Ws[j][k]
__global__ void __launch_bounds__(256) real_gabor_layer_dgrad_kernel(const float* __restrict__ gf, const float* __restrict__ gs, int n, int K, int M,
                                                                      const float* __restrict__ Wf, const float* __restrict__ Ws,
                                                                      float* __restrict__ gx) {
  __shared__ float Gf[16][65], Gs[16][65], Wa[16][65], Wb[16][65];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  float acc[4][4] = {};
  for (int j0 = 0; j0 < M; j0 += 16) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 256 + tid;
      const int rr = idx >> 4, jj = idx & 15;          // g tiles: 64 rows x 16 j (j contiguous in memory)
      const bool jok = j0 + jj < M;
      Gf[jj][rr] = (jok && r0 + rr < n) ? gf[size_t(r0 + rr) * M + j0 + jj] : 0.f;
      Gs[jj][rr] = (jok && r0 + rr < n) ? gs[size_t(r0 + rr) * M + j0 + jj] : 0.f;
      const int j2 = idx >> 6, kk = idx & 63;          // weight tiles: 16 j x 64 k (k contiguous in memory)
      const bool wok = j0 + j2 < M && k0 + kk < K;
      Wa[j2][kk] = wok ? Wf[size_t(j0 + j2) * K + k0 + kk] : 0.f;
      Wb[j2][kk] = wok ? Ws[size_t(j0 + j2) * K + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      float a[4], b[4], wa[4], wb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = Gf[jj][ty * 4 + i]; b[i] = Gs[jj][ty * 4 + i]; wa[i] = Wa[jj][tx * 4 + i]; wb[i] = Wb[jj][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], wa[j], fmaf(b[i], wb[j], acc[i][j]));
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) gx[size_t(r) * K + k] = acc[i][j];
    }
  }
}
// g_W[j][k] += sum_n g[n][j] x[n][k] over this block's row range; g_b[j] += sum_n g[n][j] (by the blocks of the first k tile)
__global__ void __launch_bounds__(256) real_gabor_layer_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ x, int n, int K, int M,
                                                                      int rows_per_split, float* __restrict__ gW, float* __restrict__ gb) {
  __shared__ float Gt[16][65], Xt[16][65];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int j0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int n0 = blockIdx.z * rows_per_split;
  const int n1 = min(n, n0 + rows_per_split);
  float acc[4][4] = {};
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r0 = n0; r0 < n1; r0 += 16) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 256 + tid, rr = idx >> 6, cc = idx & 63;   // 16 rows x 64 columns (columns contiguous in memory)
      const bool rok = r0 + rr < n1;
      Gt[rr][cc] = (rok && j0 + cc < M) ? g[size_t(r0 + rr) * M + j0 + cc] : 0.f;
      Xt[rr][cc] = (rok && k0 + cc < K) ? x[size_t(r0 + rr) * K + k0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      float gv[4], xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { gv[i] = Gt[rr][ty * 4 + i]; xv[i] = Xt[rr][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum[i] += gv[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], xv[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + ty * 4 + i;
    if (j >= M) continue;
    if (gb && blockIdx.y == 0 && tx == 0) atomicAdd(gb + j, bsum[i]);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int k = k0 + tx * 4 + jj;
      if (k < K) atomicAdd(gW + size_t(j) * K + k, acc[i][jj]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Radon forward operator of the CT driver (modules/lin_inverse.py:19-40, wire_ct.py:126-128): every angle rotates the image
// about its centre (kornia.geometry.rotate: bilinear, zeros outside, align_corners=True, centre ((W-1)/2, (H-1)/2), positive
// angle = counter-clockwise with the origin at the top-left pixel) and sums over the rows:
//     sino[a][m][j] = sum_i rot_a(im[m])[i][j],   rot_a(im)[i][j] = bilinear(im, xs, ys)
//     xs = cx + cos(t) (j - cx) - sin(t) (i - cy),   ys = cy + sin(t) (j - cx) + cos(t) (i - cy),   t = angle_a in radians
// One thread per (angle, image, column) walks the rows: neighbouring threads sample neighbouring source pixels (coalesced up
// to the rotation), and the 4 bilinear taps hit L1/L2 (a 512^2 image is 1 MB).  The adjoint scatters w * g_sino[a][m][j] to the
// same 4 taps with atomics (the gradient of the image the network produced).  HBM bytes are negligible next to the network.
// ---------------------------------------------------------------------------------------------
struct RadonTaps { int x0, y0; float w00, w01, w10, w11; };
__device__ __forceinline__ RadonTaps radon_taps(float cs, float sn, float cx, float cy, int i, int j) {
  const float dx = float(j) - cx, dy = float(i) - cy;
  const float xs = fmaf(cs, dx, fmaf(-sn, dy, cx)), ys = fmaf(sn, dx, fmaf(cs, dy, cy));
  const float xf = floorf(xs), yf = floorf(ys);
  const float ax = xs - xf, ay = ys - yf;
  RadonTaps t;
  t.x0 = int(xf); t.y0 = int(yf);
  t.w00 = (1.f - ax) * (1.f - ay); t.w01 = ax * (1.f - ay); t.w10 = (1.f - ax) * ay; t.w11 = ax * ay;
  return t;
}
__global__ void radon_fwd_kernel(const float* __restrict__ im, int nimg, int H, int W, const float* __restrict__ angles, int nangles,
                                 float* __restrict__ sino) {
  const int64_t total = int64_t(nangles) * nimg * W;
  const float cx = 0.5f * float(W - 1), cy = 0.5f * float(H - 1);
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int j = int(t % W);
    const int m = int((t / W) % nimg);
    const int a = int(t / (int64_t(W) * nimg));
    float sn, cs;
    sincosf(angles[a] * 0.017453292519943295f, &sn, &cs);
    const float* src = im + size_t(m) * H * W;
    float acc = 0.f;
    for (int i = 0; i < H; ++i) {
      const RadonTaps k = radon_taps(cs, sn, cx, cy, i, j);
      const bool xa = k.x0 >= 0 && k.x0 < W, xb = k.x0 + 1 >= 0 && k.x0 + 1 < W;
      const bool ya = k.y0 >= 0 && k.y0 < H, yb = k.y0 + 1 >= 0 && k.y0 + 1 < H;
      float v = 0.f;
      if (ya && xa) v = fmaf(k.w00, __ldg(src + size_t(k.y0) * W + k.x0), v);
      if (ya && xb) v = fmaf(k.w01, __ldg(src + size_t(k.y0) * W + k.x0 + 1), v);
      if (yb && xa) v = fmaf(k.w10, __ldg(src + size_t(k.y0 + 1) * W + k.x0), v);
      if (yb && xb) v = fmaf(k.w11, __ldg(src + size_t(k.y0 + 1) * W + k.x0 + 1), v);
      acc += v;
    }
    sino[t] = acc;
  }
}
// g_im (pre-zeroed) += adjoint of the above applied to g_sino; one thread per (angle, image, row block, column)
__global__ void radon_bwd_kernel(const float* __restrict__ g_sino, int nimg, int H, int W, const float* __restrict__ angles, int nangles,
                                 int rows_per_thread, float* __restrict__ g_im) {
  const int rblocks = (H + rows_per_thread - 1) / rows_per_thread;
  const int64_t total = int64_t(nangles) * nimg * rblocks * W;
  const float cx = 0.5f * float(W - 1), cy = 0.5f * float(H - 1);
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int j = int(t % W);
    const int rb = int((t / W) % rblocks);
    const int m = int((t / (int64_t(W) * rblocks)) % nimg);
    const int a = int(t / (int64_t(W) * rblocks * nimg));
    float sn, cs;
    sincosf(angles[a] * 0.017453292519943295f, &sn, &cs);
    const float g = g_sino[(size_t(a) * nimg + m) * W + j];
    float* dst = g_im + size_t(m) * H * W;
    const int i1 = min(H, (rb + 1) * rows_per_thread);
    for (int i = rb * rows_per_thread; i < i1; ++i) {
      const RadonTaps k = radon_taps(cs, sn, cx, cy, i, j);
      const bool xa = k.x0 >= 0 && k.x0 < W, xb = k.x0 + 1 >= 0 && k.x0 + 1 < W;
      const bool ya = k.y0 >= 0 && k.y0 < H, yb = k.y0 + 1 >= 0 && k.y0 + 1 < H;
      if (ya && xa) atomicAdd(dst + size_t(k.y0) * W + k.x0, k.w00 * g);
      if (ya && xb) atomicAdd(dst + size_t(k.y0) * W + k.x0 + 1, k.w01 * g);
      if (yb && xa) atomicAdd(dst + size_t(k.y0 + 1) * W + k.x0, k.w10 * g);
      if (yb && xb) atomicAdd(dst + size_t(k.y0 + 1) * W + k.x0 + 1, k.w11 * g);
    }
  }
}

}  // namespace wire
