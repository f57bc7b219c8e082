// gabor_math.cuh — the complex Gabor wavelet and its Wirtinger derivative, shared by every kernel.
//
// Reference semantics (file:line relative to the reference checkout):
//   forward   modules/wire.py:88-93     y = exp(1j*omega_0*z - |scale_0*z|^2)
//             modules/wire2d.py:56-67   y = exp(1j*omega_0*z) * exp(-scale_0^2*(|z|^2+|w|^2))
//   backward  PyTorch complex autograd of the above:
//             p = conj(y)*g_y ; g_z = -j*omega_0*p - 2*scale_0^2*z*Re(p) ; g_w = -2*scale_0^2*w*Re(p)
//             first layer (real z): g_z = omega_0*Im(p) - 2*scale_0^2*z*Re(p)
#pragma once
#include <cuda_runtime.h>

namespace wire {

// Accumulator column order of the 16-bit GABOR / FIRST row-tile modes (tc_rows16.cuh): position p holds logical column
// acc_col_perm(p) — the two low bits swapped, (re A, im A, re B, im B) -> (re A, re B, im A, im B) for features A = 2j,
// B = 2j+1 — so that packed FP32 math on feature pairs needs no register shuffles.  An involution; applied by
// pack_weights_kernel to the rows of the packed weight matrices.  MODE_PLAIN is unpermuted.
__host__ __device__ __forceinline__ int acc_col_perm(int p) { return (p & ~3) | ((p & 1) << 1) | ((p >> 1) & 1); }

// FAST = true : MUFU ex2/sin/cos with an explicit two-term Cody-Waite reduction (TF32 path)
// FAST = false: libdevice expf/sincosf (FP32 path)
template <bool FAST>
__device__ __forceinline__ void exp_cis(float mag_arg, float phase, float& yr, float& yi) {
  if constexpr (FAST) {
    const float k = rintf(phase * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, phase);       // 2*pi hi
    r = fmaf(-k, -1.7484555314695172e-7f, r);             // 2*pi lo
    float s, c;
    __sincosf(r, &s, &c);
    const float m = __expf(mag_arg);
    yr = m * c;
    yi = m * s;
  } else {
    float s, c;
    sincosf(phase, &s, &c);
    const float m = expf(mag_arg);
    yr = m * c;
    yi = m * s;
  }
}

// ---- branch-free FTZ fast path (TF32 mode): magnitude via ex2, phase via sin / cos.approx ----
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sin_ftz(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cos_ftz(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

struct GaborConst {
  float omega, s2;
  float c_t;     // -s2 * log2(e)
  float c_zi;    // -omega * log2(e)
  float c_turn;  // omega / (2 pi)
};
__device__ __forceinline__ GaborConst make_gabor_const(float omega, float scale) {
  GaborConst g;
  g.omega = omega;
  g.s2 = scale * scale;
  g.c_t = -g.s2 * 1.4426950408889634f;
  g.c_zi = -omega * 1.4426950408889634f;
  g.c_turn = omega * 0.15915494309189535f;
  return g;
}
// y = exp(j w z - s2 (|z|^2 + wnorm)), wnorm = |w|^2 for wire2d
__device__ __forceinline__ void gabor_fast(const GaborConst& g, float zr, float zi, float wnorm, float& yr, float& yi) {
  const float t = fmaf(zi, zi, fmaf(zr, zr, wnorm));
  const float m = ex2_ftz(fmaf(g.c_t, t, g.c_zi * zi));
  // the phase goes in as it is: sin / cos.approx multiply by 1/2pi and the MUFU takes the fraction of a turn (f32x2.cuh:
  // gabor_phase_x2; tools/sincos_probe: 1e-5 up to |w Re z| = 100).  The explicit reduction it replaces cost a FRND on the XU pipe.
#ifdef WIRE_B200_EXPLICIT_TURNS
  float u = zr * g.c_turn;
  u -= rintf(u);
  const float r = u * 6.283185307179586f;
#else
  const float r = zr * g.omega;
#endif
  yr = m * cos_ftz(r);
  yi = m * sin_ftz(r);
}

// hidden layer: complex z; `extra` = scale^2*|w|^2 for wire2d, 0 for wire
template <bool FAST>
__device__ __forceinline__ void gabor_fwd(float zr, float zi, float omega, float s2, float extra,
                                          float& yr, float& yi) {
  const float mag_arg = -omega * zi - s2 * (zr * zr + zi * zi) - extra;
  exp_cis<FAST>(mag_arg, omega * zr, yr, yi);
}
// same with precomputed constants; FAST -> branch-free MUFU path, else libdevice
template <bool FAST>
__device__ __forceinline__ void gabor_fwd_c(const GaborConst& g, float zr, float zi, float wnorm, float& yr, float& yi) {
  if constexpr (FAST) gabor_fast(g, zr, zi, wnorm, yr, yi);
  else gabor_fwd<false>(zr, zi, g.omega, g.s2, g.s2 * wnorm, yr, yi);
}

// g_z for a hidden layer from y, z and the upstream g_y; returns Re(p) for the wire2d g_w term
__device__ __forceinline__ float gabor_bwd(float yr, float yi, float zr, float zi, float gr, float gi,
                                           float omega, float s2, float& gzr, float& gzi) {
  const float pr = yr * gr + yi * gi;   // p = conj(y) * g_y
  const float pi = yr * gi - yi * gr;
  const float t = -2.0f * s2 * pr;
  gzr = fmaf(omega, pi, t * zr);
  gzi = fmaf(-omega, pr, t * zi);
  return pr;
}

// first layer: real z (and real w for wire2d)
__device__ __forceinline__ float gabor_first_bwd(float yr, float yi, float z, float gr, float gi,
                                                 float omega, float s2, float& gz) {
  const float pr = yr * gr + yi * gi;
  const float pi = yr * gi - yi * gr;
  gz = fmaf(omega, pi, -2.0f * s2 * pr * z);
  return pr;
}

}  // namespace wire
