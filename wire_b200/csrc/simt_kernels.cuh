// simt_kernels.cuh — the CUDA-core kernels of the WIRE hot path: everything that is too skinny for
// tensor cores (first layer K = 2..3, final layer out = 1..3), weight packing, and the optimiser.
// All are HBM-bound streaming kernels: coalesced along the feature axis, rows blocked per CTA.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gabor_math.cuh"
#include "sm100.cuh"

namespace wire {

constexpr int kSimtMaxIn = 8;
constexpr int kSimtMaxOut = 8;

// ---------------------------------------------------------------------------------------------
// first layer forward (modules/wire.py:88-93 with is_first=True; wire2d.py:56-67):
//   z = c W0^T + b0 (real) ; [w = c W0b^T + b0b] ; y = exp(j w0 z - s0^2 (z^2 [+ w^2]))  (complex)
// y: [n][y_pitch] interleaved complex. Optional z_out/w_out (real [n][zr_pitch]) for the per-layer API.
// ---------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(256) first_fwd_kernel(const float* __restrict__ coords, int n, int in_f, int M,
                                                         const float* __restrict__ W0, const float* __restrict__ b0,
                                                         const float* __restrict__ W0b, const float* __restrict__ b0b,
                                                         const float* __restrict__ omega_p, const float* __restrict__ scale_p,
                                                         float* __restrict__ y, int y_pitch, int round_y,
                                                         float* __restrict__ z_out, float* __restrict__ w_out, int zr_pitch,
                                                         int rows_per_block) {
  __shared__ float sc[64 * kSimtMaxIn];
  const int row0 = blockIdx.x * rows_per_block;
  int rows = n - row0;
  rows = rows > rows_per_block ? rows_per_block : rows;
  for (int i = threadIdx.x; i < rows * in_f; i += blockDim.x) sc[i] = coords[size_t(row0) * in_f + i];
  __syncthreads();
  const float omega = __ldg(omega_p), s = __ldg(scale_p), s2 = s * s;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    float w[kSimtMaxIn], wb[kSimtMaxIn];
#pragma unroll
    for (int d = 0; d < kSimtMaxIn; ++d) {
      w[d] = d < in_f ? W0[size_t(j) * in_f + d] : 0.f;
      wb[d] = (W0b && d < in_f) ? W0b[size_t(j) * in_f + d] : 0.f;
    }
    const float bj = b0[j], bbj = W0b ? b0b[j] : 0.f;
    for (int r = 0; r < rows; ++r) {
      float z = bj, ww = bbj;
#pragma unroll
      for (int d = 0; d < kSimtMaxIn; ++d)
        if (d < in_f) { z = fmaf(sc[r * in_f + d], w[d], z); ww = fmaf(sc[r * in_f + d], wb[d], ww); }
      float yr, yi;
      gabor_fwd<FAST>(z, 0.f, omega, s2, W0b ? s2 * ww * ww : 0.f, yr, yi);
      if (round_y) { yr = sm100::round_tf32(yr); yi = sm100::round_tf32(yi); }
      *reinterpret_cast<float2*>(y + size_t(row0 + r) * y_pitch + 2 * j) = make_float2(yr, yi);
      if (z_out) z_out[size_t(row0 + r) * zr_pitch + j] = z;
      if (w_out) w_out[size_t(row0 + r) * zr_pitch + j] = ww;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing: complex W[M_out][K_in] -> real K-major B matrices for tc_rows (TF32-rounded).
//   mode 0 (forward)  row = output real column (z cols, then w cols per block for wire2d)
//                     z_re: ( Wre, -Wim)   z_im: ( Wim,  Wre)          z = x W^T
//   mode 1 (dgrad)    row = real column of g_x ; K runs over g_z columns (then g_w columns)
//                     gx_re: ( Wre,  Wim)  gx_im: (-Wim,  Wre)         g_x = g_z conj(W)
// B is [n_blocks*nb][k_pad_total], zero padded.
// ---------------------------------------------------------------------------------------------
// One element of a packed matrix (see pack_weights_kernel below for the layouts).
__device__ __forceinline__ float pack_weight_value(const float* __restrict__ W1, const float* __restrict__ W2, int M_out, int K_in, int mode,
                                                   int nb, int nbh, int k0_pad, int k_pad_total, int pair_perm, int idx) {
  const int r = idx / k_pad_total, kk = idx % k_pad_total;
  float v = 0.f;
  if (mode == 0) {
    const int blk = r / nb, c = r % nb;
    const float* W = W1;
    int oc;
    if (nbh < nb) {  // wire2d forward: [z half | w half]
      if (c < nbh) oc = blk * nbh + (pair_perm ? acc_col_perm(c) : c); else { oc = blk * nbh + (pair_perm ? acc_col_perm(c - nbh) : c - nbh); W = W2; }
    } else oc = blk * nb + (pair_perm ? acc_col_perm(c) : c);
    if (oc < 2 * M_out && kk < 2 * K_in && c < 2 * nbh && W) {
      const int j = oc >> 1, part = oc & 1, k = kk >> 1, d = kk & 1;
      const float wr = W[(size_t(j) * K_in + k) * 2], wi = W[(size_t(j) * K_in + k) * 2 + 1];
      v = part == 0 ? (d == 0 ? wr : -wi) : (d == 0 ? wi : wr);
    }
  } else {
    const int oc = pair_perm ? acc_col_perm(r) : r;
    const float* W = W1;
    int kq = kk;
    if (kk >= k0_pad) { W = W2; kq = kk - k0_pad; }
    if (oc < 2 * K_in && kq < 2 * M_out && W) {
      const int k = oc >> 1, part = oc & 1, j = kq >> 1, d = kq & 1;
      const float wr = W[(size_t(j) * K_in + k) * 2], wi = W[(size_t(j) * K_in + k) * 2 + 1];
      v = part == 0 ? (d == 0 ? wr : wi) : (d == 0 ? -wi : wr);
    }
  }
  return v;
}

// All packed matrices of one training step in ONE launch (mixed16 path): job = blockIdx.y.
struct PackJob {
  const float* W1;
  const float* W2;
  void* B;
  int mode, n_blocks, nb, nbh, k0_pad, k_pad_total, elem;
};
constexpr int kMaxPackJobs = 32;
struct PackJobs {
  PackJob job[kMaxPackJobs];
  int n, M_out, K_in;
};
__global__ void pack_all16_kernel(const __grid_constant__ PackJobs J) {
  // First kernel of a step: wait BEFORE letting the dependents go.  The kernels behind it read parameters in their
  // pre-wait prologues, which is only safe once whatever wrote the parameters (the optimiser kernel in front of this one,
  // if it sits directly in front) has completed.
  sm100::pdl_wait();  // (also: the packed matrices are still read by the previous step's backward kernels)
  sm100::pdl_trigger();
  const PackJob& j = J.job[blockIdx.y];
  const int total = j.n_blocks * j.nb * j.k_pad_total;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const float v = pack_weight_value(j.W1, j.W2, J.M_out, J.K_in, j.mode, j.nb, j.nbh, j.k0_pad, j.k_pad_total, 1, idx);
    if (j.elem == 1) reinterpret_cast<__half*>(j.B)[idx] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(j.B)[idx] = __float2bfloat16_rn(v);
  }
}

// ELEM (sm100_host::ElemType): 0 = fp32 (TF32-rounded if do_round), 1 = FP16, 2 = BF16 (mixed16 path)
template <int ELEM>
__global__ void pack_weights_kernel(const float* __restrict__ W1, const float* __restrict__ W2, int M_out, int K_in,
                                    int mode, int n_blocks, int nb, int nbh, int k0_pad, int k_pad_total,
                                    void* __restrict__ Bv, int do_round, int pair_perm) {
  const int total = n_blocks * nb * k_pad_total;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / k_pad_total, kk = idx % k_pad_total;
    float v = 0.f;
    if (mode == 0) {
      const int blk = r / nb, c = r % nb;
      const float* W = W1;
      int oc;
      // pair_perm: accumulator columns in pair-transposed order (acc_col_perm; halves / blocks are multiples of 4 columns)
      if (nbh < nb) {  // wire2d forward: [z half | w half]
        if (c < nbh) oc = blk * nbh + (pair_perm ? acc_col_perm(c) : c); else { oc = blk * nbh + (pair_perm ? acc_col_perm(c - nbh) : c - nbh); W = W2; }
      } else oc = blk * nb + (pair_perm ? acc_col_perm(c) : c);
      if (oc < 2 * M_out && kk < 2 * K_in && c < 2 * nbh && W) {
        const int j = oc >> 1, part = oc & 1, k = kk >> 1, d = kk & 1;
        const float wr = W[(size_t(j) * K_in + k) * 2], wi = W[(size_t(j) * K_in + k) * 2 + 1];
        v = part == 0 ? (d == 0 ? wr : -wi) : (d == 0 ? wi : wr);
      }
    } else {
      const int oc = pair_perm ? acc_col_perm(r) : r;  // real column of g_x, blocks are contiguous
      const float* W = W1;
      int kq = kk;
      if (kk >= k0_pad) { W = W2; kq = kk - k0_pad; }
      if (oc < 2 * K_in && kq < 2 * M_out && W) {
        const int k = oc >> 1, part = oc & 1, j = kq >> 1, d = kq & 1;
        const float wr = W[(size_t(j) * K_in + k) * 2], wi = W[(size_t(j) * K_in + k) * 2 + 1];
        v = part == 0 ? (d == 0 ? wr : wi) : (d == 0 ? -wi : wr);
      }
    }
    if constexpr (ELEM == 0) reinterpret_cast<float*>(Bv)[idx] = do_round ? sm100::round_tf32(v) : v;
    else if constexpr (ELEM == 1) reinterpret_cast<__half*>(Bv)[idx] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(Bv)[idx] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------
// final Linear forward when it is not fused into the last Gabor epilogue:
//   out[n][o] = Re( sum_k h[n,k] Wf[o,k] + bf[o] )           (modules/wire.py:156-165)
// one warp per row, lanes strided over k.
// ---------------------------------------------------------------------------------------------
// H16: h is an FP16 tensor (mixed16 path), h_pitch in elements
template <bool H16 = false>
__global__ void __launch_bounds__(256) final_fwd_kernel(const float* __restrict__ h, int h_pitch, int n, int M, int out_f,
                                                         const float* __restrict__ Wf, const float* __restrict__ bf,
                                                         float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < n; row += nwarps) {
    float acc[kSimtMaxOut];
#pragma unroll
    for (int o = 0; o < kSimtMaxOut; ++o) acc[o] = 0.f;
    for (int k = lane; k < M; k += 32) {
      float2 hv;
      if constexpr (H16) hv = __half22float2(*reinterpret_cast<const __half2*>(reinterpret_cast<const __half*>(h) + size_t(row) * h_pitch + 2 * k));
      else hv = *reinterpret_cast<const float2*>(h + size_t(row) * h_pitch + 2 * k);
#pragma unroll
      for (int o = 0; o < kSimtMaxOut; ++o)
        if (o < out_f) {
          const float2 wv = *reinterpret_cast<const float2*>(Wf + (size_t(o) * M + k) * 2);
          acc[o] = fmaf(hv.x, wv.x, fmaf(-hv.y, wv.y, acc[o]));
        }
    }
#pragma unroll
    for (int o = 0; o < kSimtMaxOut; ++o)
      if (o < out_f) {
        float v = acc[o];
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) out[size_t(row) * out_f + o] = v + bf[2 * o];
      }
  }
}

// ---------------------------------------------------------------------------------------------
// top of the backward pass: final Linear backward + Gabor backward of the last hidden layer.
//   g_h = g_o conj(Wf) ; g_Wf = g_o^T conj(h) ; g_bf = sum g_o (imag = 0)
//   h = gabor(z_H [, w_H]) is RECOMPUTED from the saved pre-activation; g_z (g_w) are written
//   TF32-rounded for the dgrad / wgrad GEMMs that consume them.
// If z == nullptr the kernel is the plain final-linear backward (h given, g_h written).
// block = 256 threads over the feature axis, `rows_per_block` rows per block.
// ---------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(256) top_bwd_kernel(const float* __restrict__ g_out, int n, int M, int out_f,
                                                       const float* __restrict__ Wf,
                                                       const float* __restrict__ z, const float* __restrict__ w, int zw_pitch,
                                                       const float* __restrict__ h_in, int h_pitch,
                                                       const float* __restrict__ omega_p, const float* __restrict__ scale_p,
                                                       float* __restrict__ gz, float* __restrict__ gw, int g_pitch, int round_g,
                                                       float* __restrict__ g_Wf, float* __restrict__ g_bf,
                                                       int rows_per_block) {
  __shared__ float sgo[64 * kSimtMaxOut];
  const int row0 = blockIdx.x * rows_per_block;
  int rows = n - row0;
  rows = rows > rows_per_block ? rows_per_block : rows;
  for (int i = threadIdx.x; i < rows * out_f; i += blockDim.x) sgo[i] = g_out[size_t(row0) * out_f + i];
  __syncthreads();
  const float omega = omega_p ? __ldg(omega_p) : 0.f;
  const float s = scale_p ? __ldg(scale_p) : 0.f, s2 = s * s;
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    float wr[kSimtMaxOut], wi[kSimtMaxOut], ar[kSimtMaxOut], ai[kSimtMaxOut];
#pragma unroll
    for (int o = 0; o < kSimtMaxOut; ++o) {
      wr[o] = o < out_f ? Wf[(size_t(o) * M + k) * 2] : 0.f;
      wi[o] = o < out_f ? Wf[(size_t(o) * M + k) * 2 + 1] : 0.f;
      ar[o] = 0.f; ai[o] = 0.f;
    }
    for (int r = 0; r < rows; ++r) {
      const size_t row = size_t(row0 + r);
      float gyr = 0.f, gyi = 0.f;
#pragma unroll
      for (int o = 0; o < kSimtMaxOut; ++o)
        if (o < out_f) { gyr = fmaf(sgo[r * out_f + o], wr[o], gyr); gyi = fmaf(-sgo[r * out_f + o], wi[o], gyi); }
      float yr, yi;
      if (z) {
        const float2 zv = *reinterpret_cast<const float2*>(z + row * zw_pitch + 2 * k);
        float2 wv = make_float2(0.f, 0.f);
        if (w) wv = *reinterpret_cast<const float2*>(w + row * zw_pitch + 2 * k);
        gabor_fwd<FAST>(zv.x, zv.y, omega, s2, w ? s2 * (wv.x * wv.x + wv.y * wv.y) : 0.f, yr, yi);
        float gzr, gzi;
        const float pr = gabor_bwd(yr, yi, zv.x, zv.y, gyr, gyi, omega, s2, gzr, gzi);
        if (round_g) { gzr = sm100::round_tf32(gzr); gzi = sm100::round_tf32(gzi); }
        *reinterpret_cast<float2*>(gz + row * g_pitch + 2 * k) = make_float2(gzr, gzi);
        if (w) {
          const float t = -2.0f * s2 * pr;
          float gwr = t * wv.x, gwi = t * wv.y;
          if (round_g) { gwr = sm100::round_tf32(gwr); gwi = sm100::round_tf32(gwi); }
          *reinterpret_cast<float2*>(gw + row * g_pitch + 2 * k) = make_float2(gwr, gwi);
        }
      } else {
        const float2 hv = *reinterpret_cast<const float2*>(h_in + row * h_pitch + 2 * k);
        yr = hv.x; yi = hv.y;
        if (gz) *reinterpret_cast<float2*>(gz + row * g_pitch + 2 * k) = make_float2(gyr, gyi);
      }
#pragma unroll
      for (int o = 0; o < kSimtMaxOut; ++o)
        if (o < out_f) { ar[o] = fmaf(sgo[r * out_f + o], yr, ar[o]); ai[o] = fmaf(-sgo[r * out_f + o], yi, ai[o]); }
    }
#pragma unroll
    for (int o = 0; o < kSimtMaxOut; ++o)
      if (o < out_f) {
        atomicAdd(g_Wf + (size_t(o) * M + k) * 2, ar[o]);
        atomicAdd(g_Wf + (size_t(o) * M + k) * 2 + 1, ai[o]);
      }
  }
  if (threadIdx.x < out_f) {
    float sacc = 0.f;
    for (int r = 0; r < rows; ++r) sacc += sgo[r * out_f + threadIdx.x];
    atomicAdd(g_bf + 2 * threadIdx.x, sacc);
  }
}

// ---------------------------------------------------------------------------------------------
// first layer weight gradient: g_W0[j][d] = sum_n gz0[n,j] c[n,d] ; g_b0[j] = sum_n gz0[n,j]
// ---------------------------------------------------------------------------------------------
// G16: g_z0 is a BF16 tensor (mixed16 path, widths the streaming kernel does not cover)
template <bool G16 = false>
__global__ void __launch_bounds__(256) first_wgrad_kernel(const float* __restrict__ gz0, int g_pitch,
                                                           const float* __restrict__ coords, int n, int in_f, int M,
                                                           float* __restrict__ gW0, float* __restrict__ gb0,
                                                           int rows_per_block) {
  __shared__ float sc[64 * kSimtMaxIn];
  const int row0 = blockIdx.x * rows_per_block;
  int rows = n - row0;
  rows = rows > rows_per_block ? rows_per_block : rows;
  for (int i = threadIdx.x; i < rows * in_f; i += blockDim.x) sc[i] = coords[size_t(row0) * in_f + i];
  __syncthreads();
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    float acc[kSimtMaxIn], accb = 0.f;
#pragma unroll
    for (int d = 0; d < kSimtMaxIn; ++d) acc[d] = 0.f;
    for (int r = 0; r < rows; ++r) {
      const float g = G16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gz0)[size_t(row0 + r) * g_pitch + j])
                          : gz0[size_t(row0 + r) * g_pitch + j];
      accb += g;
#pragma unroll
      for (int d = 0; d < kSimtMaxIn; ++d)
        if (d < in_f) acc[d] = fmaf(g, sc[r * in_f + d], acc[d]);
    }
#pragma unroll
    for (int d = 0; d < kSimtMaxIn; ++d)
      if (d < in_f) atomicAdd(gW0 + size_t(j) * in_f + d, acc[d]);
    atomicAdd(gb0 + j, accb);
  }
}

// g_c[n][d] (+)= sum_j gz0[n,j] W0[j,d]   (gradient w.r.t. the coordinates; one warp per row)
__global__ void __launch_bounds__(256) grad_coords_kernel(const float* __restrict__ gz0, int g_pitch, int n, int in_f, int M,
                                                           const float* __restrict__ W0, float* __restrict__ gc, int accumulate) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < n; row += nwarps) {
    float acc[kSimtMaxIn];
#pragma unroll
    for (int d = 0; d < kSimtMaxIn; ++d) acc[d] = 0.f;
    for (int j = lane; j < M; j += 32) {
      const float g = gz0[size_t(row) * g_pitch + j];
#pragma unroll
      for (int d = 0; d < kSimtMaxIn; ++d)
        if (d < in_f) acc[d] = fmaf(g, W0[size_t(j) * in_f + d], acc[d]);
    }
#pragma unroll
    for (int d = 0; d < kSimtMaxIn; ++d)
      if (d < in_f) {
        float v = acc[d];
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) {
          float* dst = gc + size_t(row) * in_f + d;
          *dst = accumulate ? *dst + v : v;
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------
// strided 2-D copy (pad / unpad rows between caller tensors and the 128 B-aligned workspace)
// ---------------------------------------------------------------------------------------------
__global__ void copy2d_kernel(const float* __restrict__ src, int src_pitch, float* __restrict__ dst, int dst_pitch,
                              int64_t n, int cols, int do_round) {
  const int64_t total = n * cols;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int c = int(i % cols);
    float v = src[r * src_pitch + c];
    dst[r * dst_pitch + c] = do_round ? sm100::round_tf32(v) : v;
  }
}
// set one column of a pitched matrix (the "ones" column of activation buffers)
__global__ void set_column_kernel(float* __restrict__ dst, int pitch, int64_t n, int col, float v) {
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x)
    dst[r * pitch + col] = v;
}

__global__ void set_column16_kernel(uint16_t* __restrict__ dst, int pitch, int64_t n, int col, uint16_t bits) {
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x)
    dst[r * pitch + col] = bits;
}

// dense fp32 [n][cols] -> pitched 16-bit rows (the single-layer entry points under mixed16: caller tensors are fp32)
template <int ELEM>  // 1 = f16, 2 = bf16
__global__ void to16_rows_kernel(const float* __restrict__ src, int64_t n, int cols, void* __restrict__ dst, int dst_pitch) {
  const int64_t total = n * cols;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int c = int(i % cols);
    if constexpr (ELEM == 1) reinterpret_cast<__half*>(dst)[r * dst_pitch + c] = __float2half_rn(src[i]);
    else reinterpret_cast<__nv_bfloat16*>(dst)[r * dst_pitch + c] = __float2bfloat16_rn(src[i]);
  }
}

// workspace tensor (fp32, FP16 or BF16, row pitch in elements) -> dense fp32 [n][cols] (wire_net_workspace_read)
template <int ELEM>  // sm100_host::ElemType: 0 = f32, 1 = f16, 2 = bf16
__global__ void read_rows_kernel(const void* __restrict__ src, int src_pitch, int64_t n, int cols, float* __restrict__ dst) {
  const int64_t total = n * cols;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int c = int(i % cols);
    float v;
    if constexpr (ELEM == 1) v = __half2float(reinterpret_cast<const __half*>(src)[r * src_pitch + c]);
    else if constexpr (ELEM == 2) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[r * src_pitch + c]);
    else v = reinterpret_cast<const float*>(src)[r * src_pitch + c];
    dst[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, amsgrad=False, maximize=False) on a flat fp32 view; complex params are
// their view_as_real, which is exactly how torch treats them.
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t count, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            float grad_scale) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// Same update with the step counter and the learning rate living on the device, so a captured CUDA graph can be
// replayed: *step_ptr is read (1-based step = *step_ptr + 1) by every thread and incremented by one thread at the end.
// zero_grad: the gradient is cleared as it is consumed, so the next backward pass can accumulate without a memset.
__global__ void adam_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                int64_t count, const float* __restrict__ lr_ptr, float b1, float b2, float eps, float wd,
                                long long* __restrict__ step_ptr, float grad_scale, unsigned int* __restrict__ done_counter,
                                int zero_grad) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const long long step = *step_ptr + 1;
  const float lr = *lr_ptr;
  const float bc1 = 1.0f - powf(b1, float(step));
  const float bc2_sqrt = sqrtf(1.0f - powf(b2, float(step)));
  auto upd = [&](float gi, float& pp, float& mm, float& vv) {
    gi *= grad_scale;
    if (wd != 0.f) gi = fmaf(wd, pp, gi);
    mm = fmaf(b1, mm, (1.f - b1) * gi);
    vv = fmaf(b2, vv, (1.f - b2) * gi * gi);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp = pp - (lr / bc1) * (mm / denom);
  };
  // 16-byte accesses when the four buffers allow it (the Trainer's flat buffers do), scalar tail / fallback otherwise
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t n4 = vec ? count >> 2 : 0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    upd(g4.x, p4.x, m4.x, v4.x); upd(g4.y, p4.y, m4.y, v4.y); upd(g4.z, p4.z, m4.z, v4.z); upd(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t i = 4 * n4 + blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    upd(g[i], pi, mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (zero_grad) g[i] = 0.f;
  }
  // the last block to finish bumps the step counter (every block has read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) { *step_ptr = step; *done_counter = 0u; }
  }
}

// grad = 2 (pred - target) / count ; loss += sum (pred-target)^2 / count
// count_norm: the element count the mean is taken over (== count on one GPU; the GLOBAL batch's count when this rank holds
// a shard of it, so that the ranks' gradients and losses simply add up to those of the global mean)
// Loss ring (ring != nullptr): the loss of training step s (= *step_ptr, the count of completed optimiser steps) is
// accumulated into ring[s % ring_n] and the NEXT slot is cleared for step s + 1, so the host can read a step's loss any time
// within the following ring_n - 1 steps without a device-side reset kernel and without stalling the compute stream.
__global__ void mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t count,
                                float* __restrict__ grad, float* __restrict__ loss, int64_t count_norm,
                                float* __restrict__ ring = nullptr, int ring_n = 0, const long long* __restrict__ step_ptr = nullptr) {
  sm100::pdl_trigger();
  sm100::pdl_wait();
  if (ring) {
    const long long s = *step_ptr;
    loss = ring + (s % ring_n);
    if (blockIdx.x == 0 && threadIdx.x == 0) ring[(s + 1) % ring_n] = 0.f;
  }
  float local = 0.f;
  const float inv = 1.0f / float(count_norm);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    const float d = pred[i] - target[i];
    grad[i] = 2.0f * d * inv;
    local = fmaf(d, d, local);
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

}  // namespace wire

// =============================================================================================
// v2 streaming kernels: same math as above, restructured for memory-level parallelism
// (16-byte accesses, 4 rows in flight per thread) — these are the ones the whole-network path uses.
// =============================================================================================
namespace wire {

// first layer forward, 2 complex features (one float4) per thread, rows_per_block rows per block
// OUT16: y is an FP16 tensor (mixed16 path; y_pitch in elements): 8-byte stores of two complex features
template <bool FAST, bool OUT16 = false>
__global__ void __launch_bounds__(128) first_fwd2_kernel(const float* __restrict__ coords, int n, int in_f, int M,
                                                          const float* __restrict__ W0, const float* __restrict__ b0,
                                                          const float* __restrict__ W0b, const float* __restrict__ b0b,
                                                          const float* __restrict__ omega_p, const float* __restrict__ scale_p,
                                                          float* __restrict__ y, int y_pitch, int round_y, int rows_per_block) {
  __shared__ float sc[64 * kSimtMaxIn];
  const int row0 = blockIdx.x * rows_per_block;
  int rows = n - row0;
  rows = rows > rows_per_block ? rows_per_block : rows;
  for (int i = threadIdx.x; i < rows * in_f; i += blockDim.x) sc[i] = coords[size_t(row0) * in_f + i];
  __syncthreads();
  const GaborConst G = make_gabor_const(__ldg(omega_p), __ldg(scale_p));
  for (int jp = threadIdx.x; 2 * jp < M; jp += blockDim.x) {
    const int j0 = 2 * jp, j1 = (2 * jp + 1 < M) ? 2 * jp + 1 : j0;
    float w0[kSimtMaxIn], w1[kSimtMaxIn], v0[kSimtMaxIn], v1[kSimtMaxIn];
#pragma unroll
    for (int d = 0; d < kSimtMaxIn; ++d) {
      w0[d] = d < in_f ? W0[size_t(j0) * in_f + d] : 0.f;
      w1[d] = d < in_f ? W0[size_t(j1) * in_f + d] : 0.f;
      v0[d] = (W0b && d < in_f) ? W0b[size_t(j0) * in_f + d] : 0.f;
      v1[d] = (W0b && d < in_f) ? W0b[size_t(j1) * in_f + d] : 0.f;
    }
    const float bj0 = b0[j0], bj1 = b0[j1];
    const float bb0 = W0b ? b0b[j0] : 0.f, bb1 = W0b ? b0b[j1] : 0.f;
    for (int r = 0; r < rows; ++r) {
      float z0 = bj0, z1 = bj1, u0 = bb0, u1 = bb1;
#pragma unroll
      for (int d = 0; d < kSimtMaxIn; ++d)
        if (d < in_f) {
          const float c = sc[r * in_f + d];
          z0 = fmaf(c, w0[d], z0); z1 = fmaf(c, w1[d], z1);
          u0 = fmaf(c, v0[d], u0); u1 = fmaf(c, v1[d], u1);
        }
      float4 o;
      gabor_fwd_c<FAST>(G, z0, 0.f, W0b ? u0 * u0 : 0.f, o.x, o.y);
      gabor_fwd_c<FAST>(G, z1, 0.f, W0b ? u1 * u1 : 0.f, o.z, o.w);
      if (round_y) { o.x = sm100::round_tf32(o.x); o.y = sm100::round_tf32(o.y); o.z = sm100::round_tf32(o.z); o.w = sm100::round_tf32(o.w); }
      if constexpr (OUT16) {
        __half* dst = reinterpret_cast<__half*>(y) + size_t(row0 + r) * y_pitch + 4 * jp;
        const __half2 a = __floats2half2_rn(o.x, o.y), b = __floats2half2_rn(o.z, o.w);
        if (2 * jp + 1 < M) *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
        else *reinterpret_cast<__half2*>(dst) = a;
      } else {
        float* dst = y + size_t(row0 + r) * y_pitch + 4 * jp;
        if (2 * jp + 1 < M) *reinterpret_cast<float4*>(dst) = o;
        else *reinterpret_cast<float2*>(dst) = make_float2(o.x, o.y);
      }
    }
  }
}

// top of the backward pass (training path): z (and w) pitched with 16-byte aligned rows.
// One thread = 2 complex features; 4 rows of loads in flight.
// ZHALF: z / w are FP16 [n][zw_pitch] (pitch in elements)
__device__ __forceinline__ float4 load_zw4(const float* base, size_t row, int pitch, int kp, bool zhalf) {
  if (!zhalf) return __ldcs(reinterpret_cast<const float4*>(base + row * pitch + 4 * kp));
  const uint2 u = __ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(base) + row * pitch + 4 * kp));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// GBF16: g_z / g_w are BF16 tensors (mixed16 path; g_pitch in elements)
__device__ __forceinline__ void store_g4(float* base, size_t row, int pitch, int kp, bool pair, float4 v, bool bf16) {
  if (!bf16) {
    float* d = base + row * pitch + 4 * kp;
    if (pair) *reinterpret_cast<float4*>(d) = v; else *reinterpret_cast<float2*>(d) = make_float2(v.x, v.y);
  } else {
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(base) + row * pitch + 4 * kp;
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    if (pair) *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    else *reinterpret_cast<__nv_bfloat162*>(d) = a;
  }
}

template <bool FAST, bool TWO_D, bool ZHALF = false, bool GBF16 = false>
__global__ void __launch_bounds__(512) top_bwd2_kernel(const float* __restrict__ g_out, int n, int M, int out_f,
                                                        const float* __restrict__ Wf, const float* __restrict__ z,
                                                        const float* __restrict__ w, int zw_pitch,
                                                        const float* __restrict__ omega_p, const float* __restrict__ scale_p,
                                                        float* __restrict__ gz, float* __restrict__ gw, int g_pitch, int round_g,
                                                        float* __restrict__ g_Wf, float* __restrict__ g_bf, int rows_per_block) {
  // Each block owns a contiguous range of `rows_per_block` rows and walks it in chunks of 64: the g_W / g_b
  // partial sums stay in registers for the whole range, so every output address sees only gridDim.x atomics
  // (same-address atomics serialise in L2; thousands of small blocks made this kernel atomics-bound).
  // Requires M <= 2 * blockDim.x (one feature pair per thread).
  __shared__ float sgo[64 * 4];
  const GaborConst G = make_gabor_const(__ldg(omega_p), __ldg(scale_p));
  const float omega = G.omega, s2 = G.s2;
  const int kp = threadIdx.x;
  const bool active = 2 * kp < M;
  const int k0 = active ? 2 * kp : 0, k1 = (2 * kp + 1 < M) ? 2 * kp + 1 : k0;
  const bool pair = 2 * kp + 1 < M;
  float wr0[4], wi0[4], wr1[4], wi1[4], ar0[4], ai0[4], ar1[4], ai1[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    wr0[o] = o < out_f ? Wf[(size_t(o) * M + k0) * 2] : 0.f;
    wi0[o] = o < out_f ? Wf[(size_t(o) * M + k0) * 2 + 1] : 0.f;
    wr1[o] = o < out_f ? Wf[(size_t(o) * M + k1) * 2] : 0.f;
    wi1[o] = o < out_f ? Wf[(size_t(o) * M + k1) * 2 + 1] : 0.f;
    ar0[o] = ai0[o] = ar1[o] = ai1[o] = 0.f;
  }
  float bsum = 0.f;
  const int range0 = blockIdx.x * rows_per_block;
  int range1 = range0 + rows_per_block;
  range1 = range1 > n ? n : range1;
  for (int row0 = range0; row0 < range1; row0 += 64) {
    const int rows = (range1 - row0) > 64 ? 64 : (range1 - row0);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) {
      const int r = i >> 2, o = i & 3;
      sgo[i] = (r < rows && o < out_f) ? g_out[size_t(row0 + r) * out_f + o] : 0.f;
    }
    __syncthreads();
    if (threadIdx.x < 4) for (int r = 0; r < rows; ++r) bsum += sgo[r * 4 + threadIdx.x];
    if (!active) continue;
    for (int rb = 0; rb < rows; rb += 4) {
      float4 zv[4], wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u < rows ? rb + u : rows - 1;
        zv[u] = load_zw4(z, size_t(row0 + r), zw_pitch, kp, ZHALF);
        if (TWO_D) wv[u] = load_zw4(w, size_t(row0 + r), zw_pitch, kp, ZHALF);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rb + u >= rows) break;
        const int r = rb + u;
        const float4 go = *reinterpret_cast<const float4*>(&sgo[r * 4]);
        const float g4[4] = {go.x, go.y, go.z, go.w};
        float gyr0 = 0.f, gyi0 = 0.f, gyr1 = 0.f, gyi1 = 0.f;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          gyr0 = fmaf(g4[o], wr0[o], gyr0); gyi0 = fmaf(-g4[o], wi0[o], gyi0);
          gyr1 = fmaf(g4[o], wr1[o], gyr1); gyi1 = fmaf(-g4[o], wi1[o], gyi1);
        }
        float yr0, yi0, yr1, yi1;
        float e0 = 0.f, e1 = 0.f;
        if (TWO_D) { e0 = wv[u].x * wv[u].x + wv[u].y * wv[u].y; e1 = wv[u].z * wv[u].z + wv[u].w * wv[u].w; }
        gabor_fwd_c<FAST>(G, zv[u].x, zv[u].y, e0, yr0, yi0);
        gabor_fwd_c<FAST>(G, zv[u].z, zv[u].w, e1, yr1, yi1);
        float4 gzo;
        const float pr0 = gabor_bwd(yr0, yi0, zv[u].x, zv[u].y, gyr0, gyi0, omega, s2, gzo.x, gzo.y);
        const float pr1 = gabor_bwd(yr1, yi1, zv[u].z, zv[u].w, gyr1, gyi1, omega, s2, gzo.z, gzo.w);
        if (round_g) { gzo.x = sm100::round_tf32(gzo.x); gzo.y = sm100::round_tf32(gzo.y); gzo.z = sm100::round_tf32(gzo.z); gzo.w = sm100::round_tf32(gzo.w); }
        store_g4(gz, size_t(row0 + r), g_pitch, kp, pair, gzo, GBF16);
        if (TWO_D) {
          const float t0 = -2.0f * s2 * pr0, t1 = -2.0f * s2 * pr1;
          float4 gwo = make_float4(t0 * wv[u].x, t0 * wv[u].y, t1 * wv[u].z, t1 * wv[u].w);
          if (round_g) { gwo.x = sm100::round_tf32(gwo.x); gwo.y = sm100::round_tf32(gwo.y); gwo.z = sm100::round_tf32(gwo.z); gwo.w = sm100::round_tf32(gwo.w); }
          store_g4(gw, size_t(row0 + r), g_pitch, kp, pair, gwo, GBF16);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          ar0[o] = fmaf(g4[o], yr0, ar0[o]); ai0[o] = fmaf(-g4[o], yi0, ai0[o]);
          ar1[o] = fmaf(g4[o], yr1, ar1[o]); ai1[o] = fmaf(-g4[o], yi1, ai1[o]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int o = 0; o < 4; ++o)
      if (o < out_f) {
        atomicAdd(g_Wf + (size_t(o) * M + k0) * 2, ar0[o]);
        atomicAdd(g_Wf + (size_t(o) * M + k0) * 2 + 1, ai0[o]);
        if (pair) {
          atomicAdd(g_Wf + (size_t(o) * M + k1) * 2, ar1[o]);
          atomicAdd(g_Wf + (size_t(o) * M + k1) * 2 + 1, ai1[o]);
        }
      }
  }
  if (threadIdx.x < out_f) atomicAdd(g_bf + 2 * threadIdx.x, bsum);
}

// first layer weight gradient, float4 (4 features) per thread, 4 row groups per block reduced in smem.
// block = (64 feature quads) x (4 row groups); rows_per_block rows per block.
__global__ void __launch_bounds__(256) first_wgrad2_kernel(const float* __restrict__ gz0, int g_pitch,
                                                            const float* __restrict__ coords, int n, int in_f, int M,
                                                            float* __restrict__ gW0, float* __restrict__ gb0, int rows_per_block) {
  __shared__ float sc[128 * 4];
  __shared__ float red[3][64][4][5];  // row groups 1..3 -> [quad][feature][d(0..3 coords, 4 = bias)]
  // Each block owns `rows_per_block` rows and walks them in chunks of 128 (few blocks => few same-address atomics).
  // Requires M <= 256 (one float4 of features per thread column).
  const int tq = threadIdx.x & 63, tg = threadIdx.x >> 6;
  const int range0 = blockIdx.x * rows_per_block;
  int range1 = range0 + rows_per_block;
  range1 = range1 > n ? n : range1;
  {
    const int q = tq;
    float acc[4][5];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int d = 0; d < 5; ++d) acc[f][d] = 0.f;
    for (int row0 = range0; row0 < range1; row0 += 128) {
      const int rows = (range1 - row0) > 128 ? 128 : (range1 - row0);
      __syncthreads();
      for (int i = threadIdx.x; i < 128 * 4; i += blockDim.x) {
        const int r = i >> 2, d = i & 3;
        sc[i] = (r < rows && d < in_f) ? coords[size_t(row0 + r) * in_f + d] : 0.f;
      }
      __syncthreads();
      if (4 * q >= g_pitch) continue;
      for (int rb = tg; rb < rows; rb += 16) {
        float4 g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = rb + 4 * u;
          g[u] = (r < rows) ? __ldcs(reinterpret_cast<const float4*>(gz0 + size_t(row0 + r) * g_pitch + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = rb + 4 * u;
          if (r >= rows) break;
          const float4 c = *reinterpret_cast<const float4*>(&sc[r * 4]);
          const float gv[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            acc[f][0] = fmaf(gv[f], c.x, acc[f][0]);
            acc[f][1] = fmaf(gv[f], c.y, acc[f][1]);
            acc[f][2] = fmaf(gv[f], c.z, acc[f][2]);
            acc[f][3] = fmaf(gv[f], c.w, acc[f][3]);
            acc[f][4] += gv[f];
          }
        }
      }
    }
    if (tg > 0) {
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int d = 0; d < 5; ++d) red[tg - 1][tq][f][d] = acc[f][d];
    }
    __syncthreads();
    if (tg == 0) {
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int j = 4 * q + f;
        if (j < M) {
#pragma unroll
          for (int d = 0; d < 5; ++d) {
            const float v = acc[f][d] + red[0][tq][f][d] + red[1][tq][f][d] + red[2][tq][f][d];
            if (d < 4) { if (d < in_f) atomicAdd(gW0 + size_t(j) * in_f + d, v); }
            else atomicAdd(gb0 + j, v);
          }
        }
      }
    }
  }
}

}  // namespace wire
