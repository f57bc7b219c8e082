// simt_rows.cuh — FP32 CUDA-core versions of the row-tile GEMM and of the weight-gradient GEMM.
//
// WIRE_PRECISION_FP32: bit-for-bit the same data flow, packing and epilogues as the tcgen05 path
// (rows_epilogue.cuh is shared), but every product is a plain FP32 FMA and the transcendentals are
// libdevice expf/sincosf.  It is the on-device numerical yardstick for the TF32 kernels (the
// reference itself runs cgemm in FP32) — a precision mode, not a fallback: it needs the same GPU.
#pragma once
#include "rows_epilogue.cuh"

namespace wire {

struct SimtRowsParams {
  const float* a[2];
  int a_pitch[2];
  int k_cols[2];
  const float* b;  // packed [n_blocks*nb][b_pitch]
  int b_pitch;
  int k0_pad;      // column offset of part 1 inside the packed B
  float* o[3];
  int o_pitch[3];
  int n_blocks, nb, nbh, store_mask;
  RowsEpi e;
};

template <int MODE>
__global__ void __launch_bounds__(128) simt_rows_kernel(const SimtRowsParams P) {
  __shared__ float As[32][129];
  __shared__ __align__(16) float Bs[32][36];
  __shared__ __align__(16) float Bs2[32][36];
  const RowsEpi& E = P.e;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * 128;
  const int row = row0 + tid;
  const bool row_ok = row < E.n_rows;
  const float omega = (MODE != MODE_PLAIN) ? __ldg(E.omega) : 0.f;
  const float sc = (MODE != MODE_PLAIN) ? __ldg(E.scale) : 0.f;
  const float s2 = sc * sc;
  float cin[kMaxIn];
  rows_load_coords<MODE>(E, row, row_ok, cin);
  float facc[kMaxOut];
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) facc[o] = 0.f;

  for (int blk = 0; blk < P.n_blocks; ++blk) {
    const int ncol_blk = (MODE == MODE_GABOR2D_FWD) ? P.nbh : P.nb;
    const int col0 = blk * ncol_blk;
    int valid = E.n_cols - col0;
    valid = valid > ncol_blk ? ncol_blk : valid;
    const int nchunks = (valid + 31) / 32;
    for (int ch = 0; ch < nchunks; ++ch) {
      float v[32], v2[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) { v[i] = 0.f; v2[i] = 0.f; }
      const int brow = blk * P.nb + ch * 32;  // first packed-B row of this chunk (z half)
      for (int part = 0; part < 2; ++part) {
        const int kcols = P.k_cols[part];
        for (int k0 = 0; k0 < kcols; k0 += 32) {
          __syncthreads();
#pragma unroll 4
          for (int i = 0; i < 32; ++i) {
            const int idx = i * 128 + tid, r = idx >> 5, kk = idx & 31;
            float a = 0.f;
            if (row0 + r < E.n_rows && k0 + kk < kcols) a = P.a[part][size_t(row0 + r) * P.a_pitch[part] + k0 + kk];
            As[kk][r] = a;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 128 + tid, cc = idx >> 5, kk = idx & 31;
            const int kb = (part ? P.k0_pad : 0) + k0 + kk;
            const bool kok = (k0 + kk < kcols) && (ch * 32 + cc < ncol_blk);
            Bs[kk][cc] = kok ? P.b[size_t(brow + cc) * P.b_pitch + kb] : 0.f;
            if constexpr (MODE == MODE_GABOR2D_FWD) Bs2[kk][cc] = kok ? P.b[size_t(brow + P.nbh + cc) * P.b_pitch + kb] : 0.f;
          }
          __syncthreads();
#pragma unroll 8
          for (int kk = 0; kk < 32; ++kk) {
            const float a = As[kk][tid];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][4 * j]);
              v[4 * j] = fmaf(a, b.x, v[4 * j]);
              v[4 * j + 1] = fmaf(a, b.y, v[4 * j + 1]);
              v[4 * j + 2] = fmaf(a, b.z, v[4 * j + 2]);
              v[4 * j + 3] = fmaf(a, b.w, v[4 * j + 3]);
              if constexpr (MODE == MODE_GABOR2D_FWD) {
                const float4 b2 = *reinterpret_cast<const float4*>(&Bs2[kk][4 * j]);
                v2[4 * j] = fmaf(a, b2.x, v2[4 * j]);
                v2[4 * j + 1] = fmaf(a, b2.y, v2[4 * j + 1]);
                v2[4 * j + 2] = fmaf(a, b2.z, v2[4 * j + 2]);
                v2[4 * j + 3] = fmaf(a, b2.w, v2[4 * j + 3]);
              }
            }
          }
        }
      }
      const int c = col0 + ch * 32;
      float o0[32], o1[32], o2[32];
      rows_epilogue_chunk<MODE, false>(E, row, row_ok, c, omega, s2, v, v2, cin, facc, o0, o1, o2);
      if (row_ok) {
        int slot = 0;
        if (P.store_mask & 1) {
          float* dst = P.o[slot] + size_t(row) * P.o_pitch[slot] + c;
#pragma unroll
          for (int i = 0; i < 32; ++i) if (c + i < E.n_cols) dst[i] = o0[i];
          ++slot;
        }
        if constexpr (MODE == MODE_GABOR_FWD || MODE == MODE_GABOR2D_FWD || MODE == MODE_GABOR2D_BWD) {
          if (P.store_mask & 2) {
            float* dst = P.o[slot] + size_t(row) * P.o_pitch[slot] + c;
#pragma unroll
            for (int i = 0; i < 32; ++i) if (c + i < E.n_cols) dst[i] = o1[i];
            ++slot;
          }
        }
        if constexpr (MODE == MODE_GABOR2D_FWD) {
          if (P.store_mask & 4) {
            float* dst = P.o[slot] + size_t(row) * P.o_pitch[slot] + c;
#pragma unroll
            for (int i = 0; i < 32; ++i) if (c + i < E.n_cols) dst[i] = o2[i];
            ++slot;
          }
        }
      }
    }
  }
  rows_store_final<MODE>(E, row, row_ok, facc);
}

// FP32 weight gradient: G[c, r] = sum_n X[n,c] g[n,r], folded to complex exactly like tc_wgrad.
// grid = (ceil((2K+1)/64), ceil(2M/64), splits), block = 256 (16x16 threads, 4x4 outputs each).
__global__ void __launch_bounds__(256) simt_wgrad_kernel(const float* __restrict__ x, int x_pitch, int k_in,
                                                          const float* __restrict__ g, int g_pitch, int g_cols, int n,
                                                          int rows_per_split, float* __restrict__ gW, float* __restrict__ gB) {
  __shared__ __align__(16) float Xs[16][64];
  __shared__ __align__(16) float Gs[16][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int cbase = blockIdx.x * 64, rbase = blockIdx.y * 64;
  const int x_cols = 2 * k_in + 1;
  const int n0 = blockIdx.z * rows_per_split;
  int n1 = n0 + rows_per_split;
  n1 = n1 > n ? n : n1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int r0 = n0; r0 < n1; r0 += 16) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 256 + tid, kk = idx >> 6, cc = idx & 63;
      const bool rok = r0 + kk < n1;
      Xs[kk][cc] = (rok && cbase + cc < x_cols) ? x[size_t(r0 + kk) * x_pitch + cbase + cc] : 0.f;
      Gs[kk][cc] = (rok && rbase + cc < g_cols) ? g[size_t(r0 + kk) * g_pitch + rbase + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Gs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  const int c0 = cbase + ty * 4, rr0 = rbase + tx * 4;
#pragma unroll
  for (int ci = 0; ci < 4; ci += 2)
#pragma unroll
    for (int rj = 0; rj < 4; rj += 2) {
      const int c = c0 + ci, r = rr0 + rj;
      if (r >= g_cols) continue;
      if (c < 2 * k_in) {
        const size_t o = (size_t(r >> 1) * k_in + (c >> 1)) * 2;
        atomicAdd(gW + o, acc[ci][rj] + acc[ci + 1][rj + 1]);
        atomicAdd(gW + o + 1, acc[ci][rj + 1] - acc[ci + 1][rj]);
      } else if (c == 2 * k_in) {
        atomicAdd(gB + r, acc[ci][rj]);
        atomicAdd(gB + r + 1, acc[ci][rj + 1]);
      }
    }
}

}  // namespace wire
