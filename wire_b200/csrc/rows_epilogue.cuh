// rows_epilogue.cuh — the point-wise half of every row-tile GEMM, shared by the tcgen05 (TF32) and
// the CUDA-core (FP32) kernels: one thread owns one coordinate (row) and a chunk of 32 consecutive
// real accumulator columns (= 16 complex features).
//
//   MODE_PLAIN        o0 = ACC
//   MODE_GABOR_FWD    z = ACC + b ; y = gabor(z)                 o0 = y, o1 = z      modules/wire.py:88-93
//   MODE_GABOR2D_FWD  z,w = ACC halves + b1,b2 ; y = gabor2d      o0 = y, o1 = z, o2 = w  modules/wire2d.py:56-67
//   MODE_GABOR_BWD    g_y = ACC ; g_z = gabor'(z_saved, g_y)      o0 = g_z            (autograd of the above)
//   MODE_GABOR2D_BWD  same + g_w                                  o0 = g_z, o1 = g_w
//   MODE_FIRST_BWD / MODE_FIRST2D_BWD   g_y0 = ACC ; z0 recomputed from the coordinates;
//                     real g_z0 (g_w0) are stored directly (16 floats = two full sectors per thread)
#pragma once
#include "gabor_math.cuh"
#include "sm100.cuh"

namespace wire {

enum RowsMode : int {
  MODE_PLAIN = 0,
  MODE_GABOR_FWD = 1,
  MODE_GABOR2D_FWD = 2,
  MODE_GABOR_BWD = 3,
  MODE_GABOR2D_BWD = 4,
  MODE_FIRST_BWD = 5,
  MODE_FIRST2D_BWD = 6,
};

constexpr int kMaxIn = 8;   // coordinate dimensions supported by the fused first-layer epilogue
constexpr int kMaxOut = 4;  // output features supported by the fused final Linear

struct RowsEpi {
  int n_rows;
  int n_cols;      // valid real output columns (2M)
  int round_out0;  // round o0 to TF32 (it feeds the next GEMM)
  const float* bias;
  const float* bias2;
  const float* omega;  // device scalars of the layer whose nonlinearity runs in the epilogue
  const float* scale;
  const float* z_src;  // saved pre-activations for the backward epilogues
  const float* w_src;
  int zw_pitch;
  int z_half;  // tcgen05 path only: the saved z / w tensors are FP16 (same 11-bit significand as TF32, half the HBM bytes);
               // forward stores them through FP16 tensor maps, backward loads them the same way
  const float* coords;  // first-layer backward: z0 is recomputed from the coordinates
  int in_features;
  const float* w0;
  const float* b0;
  const float* w0b;
  const float* b0b;
  float* gz0;
  float* gw0;
  int gz0_pitch;
  const float* wf;  // fused final Linear: [out][M] complex interleaved
  const float* bf;
  float* out;
  int out_features;
  int fuse_final;
  // trainable omega_0 / scale_0 of the layer whose nonlinearity backward runs in this epilogue (modules/wire.py:66,80-81):
  // device floats ACCUMULATED by the 16-bit kernels' SCAL instantiations (nullptr = the scalars are constants)
  float* g_omega;
  float* g_scale;
};

__device__ __forceinline__ void load_row32(const float* src, bool ok, float (&v)[32]) {
  const float4* p = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 t = ok ? __ldg(p + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[4 * j] = t.x;
    v[4 * j + 1] = t.y;
    v[4 * j + 2] = t.z;
    v[4 * j + 3] = t.w;
  }
}

// v  : accumulator chunk (32 real columns starting at output column c)
// v2 : second accumulator chunk (w half) for MODE_GABOR2D_FWD
template <int MODE, bool FAST>
__device__ __forceinline__ void rows_epilogue_chunk(const RowsEpi& E, int row, bool row_ok, int c, float omega, float s2,
                                                    const float (&v)[32], const float (&v2)[32], const float (&cin)[kMaxIn],
                                                    float (&facc)[kMaxOut], float (&o0)[32], float (&o1)[32], float (&o2)[32]) {
  if constexpr (MODE == MODE_PLAIN) {
#pragma unroll
    for (int i = 0; i < 32; ++i) o0[i] = v[i];
  } else if constexpr (MODE == MODE_GABOR_FWD || MODE == MODE_GABOR2D_FWD) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int cc = c + 2 * i;
      const bool ok = cc < E.n_cols;
      const float zr = v[2 * i] + (ok ? __ldg(E.bias + cc) : 0.f);
      const float zi = v[2 * i + 1] + (ok ? __ldg(E.bias + cc + 1) : 0.f);
      float extra = 0.f;
      if constexpr (MODE == MODE_GABOR2D_FWD) {
        const float wr = v2[2 * i] + (ok ? __ldg(E.bias2 + cc) : 0.f);
        const float wi = v2[2 * i + 1] + (ok ? __ldg(E.bias2 + cc + 1) : 0.f);
        extra = s2 * (wr * wr + wi * wi);
        o2[2 * i] = wr;
        o2[2 * i + 1] = wi;
      }
      float yr, yi;
      gabor_fwd<FAST>(zr, zi, omega, s2, extra, yr, yi);
      if (E.fuse_final && ok) {
        const int k = cc >> 1;
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
          if (o < E.out_features) {
            const float2 wv = __ldg(reinterpret_cast<const float2*>(E.wf) + size_t(o) * (E.n_cols >> 1) + k);
            facc[o] = fmaf(yr, wv.x, fmaf(-yi, wv.y, facc[o]));
          }
        }
      }
      if (E.round_out0) { yr = sm100::round_tf32(yr); yi = sm100::round_tf32(yi); }
      o0[2 * i] = yr;
      o0[2 * i + 1] = yi;
      o1[2 * i] = zr;
      o1[2 * i + 1] = zi;
    }
  } else if constexpr (MODE == MODE_GABOR_BWD || MODE == MODE_GABOR2D_BWD) {
    float z[32];
    load_row32(E.z_src + size_t(row) * E.zw_pitch + c, row_ok, z);
    float w[32];
    if constexpr (MODE == MODE_GABOR2D_BWD) load_row32(E.w_src + size_t(row) * E.zw_pitch + c, row_ok, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float zr = z[2 * i], zi = z[2 * i + 1];
      float extra = 0.f;
      if constexpr (MODE == MODE_GABOR2D_BWD) extra = s2 * (w[2 * i] * w[2 * i] + w[2 * i + 1] * w[2 * i + 1]);
      float yr, yi, gzr, gzi;
      gabor_fwd<FAST>(zr, zi, omega, s2, extra, yr, yi);
      const float pr = gabor_bwd(yr, yi, zr, zi, v[2 * i], v[2 * i + 1], omega, s2, gzr, gzi);
      if (E.round_out0) { gzr = sm100::round_tf32(gzr); gzi = sm100::round_tf32(gzi); }
      o0[2 * i] = gzr;
      o0[2 * i + 1] = gzi;
      if constexpr (MODE == MODE_GABOR2D_BWD) {
        const float t = -2.0f * s2 * pr;
        float gwr = t * w[2 * i], gwi = t * w[2 * i + 1];
        if (E.round_out0) { gwr = sm100::round_tf32(gwr); gwi = sm100::round_tf32(gwi); }
        o1[2 * i] = gwr;
        o1[2 * i + 1] = gwi;
      }
    }
  } else {  // MODE_FIRST_BWD / MODE_FIRST2D_BWD: real z0 recomputed from coordinates
    float gz[16], gw[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = (c >> 1) + i;  // complex feature index
      const bool ok = (2 * j) < E.n_cols;
      float z0 = ok ? __ldg(E.b0 + j) : 0.f;
      float w0v = 0.f;
#pragma unroll
      for (int d = 0; d < kMaxIn; ++d)
        if (d < E.in_features && ok) z0 = fmaf(cin[d], __ldg(E.w0 + size_t(j) * E.in_features + d), z0);
      float extra = 0.f;
      if constexpr (MODE == MODE_FIRST2D_BWD) {
        w0v = ok ? __ldg(E.b0b + j) : 0.f;
#pragma unroll
        for (int d = 0; d < kMaxIn; ++d)
          if (d < E.in_features && ok) w0v = fmaf(cin[d], __ldg(E.w0b + size_t(j) * E.in_features + d), w0v);
        extra = s2 * w0v * w0v;
      }
      float yr, yi;
      gabor_fwd<FAST>(z0, 0.f, omega, s2, extra, yr, yi);
      const float pr = gabor_first_bwd(yr, yi, z0, v[2 * i], v[2 * i + 1], omega, s2, gz[i]);
      gw[i] = -2.0f * s2 * pr * w0v;
    }
    if (row_ok) {
      float4* dst = reinterpret_cast<float4*>(E.gz0 + size_t(row) * E.gz0_pitch + (c >> 1));
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4)
        if ((c >> 1) + 4 * j4 < E.gz0_pitch) dst[j4] = make_float4(gz[4 * j4], gz[4 * j4 + 1], gz[4 * j4 + 2], gz[4 * j4 + 3]);
      if constexpr (MODE == MODE_FIRST2D_BWD) {
        float4* dw = reinterpret_cast<float4*>(E.gw0 + size_t(row) * E.gz0_pitch + (c >> 1));
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          if ((c >> 1) + 4 * j4 < E.gz0_pitch) dw[j4] = make_float4(gw[4 * j4], gw[4 * j4 + 1], gw[4 * j4 + 2], gw[4 * j4 + 3]);
      }
    }
  }
}

template <int MODE>
__device__ __forceinline__ void rows_load_coords(const RowsEpi& E, int row, bool row_ok, float (&cin)[kMaxIn]) {
#pragma unroll
  for (int d = 0; d < kMaxIn; ++d) cin[d] = 0.f;
  if constexpr (MODE == MODE_FIRST_BWD || MODE == MODE_FIRST2D_BWD) {
#pragma unroll
    for (int d = 0; d < kMaxIn; ++d)
      if (d < E.in_features && row_ok) cin[d] = __ldg(E.coords + size_t(row) * E.in_features + d);
  }
}

template <int MODE>
__device__ __forceinline__ void rows_store_final(const RowsEpi& E, int row, bool row_ok, const float (&facc)[kMaxOut]) {
  if constexpr (MODE == MODE_GABOR_FWD || MODE == MODE_GABOR2D_FWD) {
    if (E.fuse_final && row_ok) {
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o)
        if (o < E.out_features) E.out[size_t(row) * E.out_features + o] = facc[o] + __ldg(E.bf + 2 * o);
    }
  }
}

}  // namespace wire
