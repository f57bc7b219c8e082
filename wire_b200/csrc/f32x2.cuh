// f32x2.cuh — packed 2 x FP32 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot, two lanes of FP32 math)
// and the complex Gabor wavelet + its Wirtinger derivative written for PAIRS of features.
//
// The 16-bit kernels' epilogues are bound by issue slots, not by HBM or the tensor pipe (ncu: issue-active 62-64 %,
// 23 instructions per complex feature of which 17 are scalar FP32): doing the FP32 math two features at a time halves
// that.  ex2 / sin / cos stay scalar (the 16-lane XU pipe has no packed form).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gabor_math.cuh"

namespace wire {

typedef unsigned long long f2;  // {lo, hi} = two floats in an aligned 64-bit register pair

__device__ __forceinline__ f2 f2_make(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 f2_bcast(float x) { return f2_make(x, x); }
__device__ __forceinline__ f2 f2_bits(uint32_t lo, uint32_t hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ float f2_lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float f2_hi(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

struct GaborConst2 {
  f2 c_t, c_zi, c_turn;   // broadcast copies of GaborConst's constants
  f2 magic, nmagic, none, two_pi;
  f2 omega, nomega, m2s2;  // backward: omega, -omega, -2 s0^2
};
__device__ __forceinline__ GaborConst2 make_gabor_const2(const GaborConst& g) {
  GaborConst2 c;
  c.c_t = f2_bcast(g.c_t);
  c.c_zi = f2_bcast(g.c_zi);
  c.c_turn = f2_bcast(g.c_turn);
  c.magic = f2_bcast(12582912.0f);
  c.nmagic = f2_bcast(-12582912.0f);
  c.none = f2_bcast(-1.0f);
  c.two_pi = f2_bcast(6.283185307179586f);
  c.omega = f2_bcast(g.omega);
  c.nomega = f2_bcast(-g.omega);
  c.m2s2 = f2_bcast(-2.0f * g.s2);
  return c;
}

// y = exp(j w z - s2 (|z|^2 + wnorm)) for two features at once (zr, zi, wnorm, yr, yi are {feature A, feature B} pairs).
// 7 packed FP32 instructions + 6 MUFU (+ 2 FMUL.RZ inside sin / cos.approx) per pair (scalar: 14 + 3 per feature).
// Phase: sin.approx / cos.approx are a multiply by 1/2pi (round toward zero) and a MUFU that takes the fraction of a turn itself, so
// the phase w Re z goes in as it is.  Measured (tools/sincos_probe, profiles/r02_probe_sincos.log): max |error| 1.0e-5 for
// |w Re z| <= 100 (every driver: omega_0 <= 20), 1.4e-4 up to 1000 -- against the 4.9e-4 of the FP16 the result is stored in.  The
// explicit reduction this replaces (u = z w/2pi; k = rint(u) by the magic constant; r = (u - k) 2pi: four more packed
// instructions per pair) was 3x more accurate at every range and bought nothing at 16-bit storage; -DWIRE_B200_EXPLICIT_TURNS
// brings it back.
__device__ __forceinline__ f2 gabor_phase_x2(const GaborConst2& c, f2 zr) {
#ifdef WIRE_B200_EXPLICIT_TURNS
  const f2 u = f2_mul(zr, c.c_turn);
  const f2 k = f2_add(f2_add(u, c.magic), c.nmagic);  // rint(u) for |u| < 2^22
  return f2_mul(f2_fma(k, c.none, u), c.two_pi);      // (u - rint(u)) is exact
#else
  return f2_mul(zr, c.omega);
#endif
}
__device__ __forceinline__ void gabor_x2(const GaborConst2& c, f2 zr, f2 zi, f2 wnorm, f2& yr, f2& yi) {
  const f2 t = f2_fma(zi, zi, f2_fma(zr, zr, wnorm));
  const f2 arg = f2_fma(c.c_t, t, f2_mul(c.c_zi, zi));
  const f2 r = gabor_phase_x2(c, zr);
  const f2 m = f2_make(ex2_ftz(f2_lo(arg)), ex2_ftz(f2_hi(arg)));
  const f2 cs = f2_make(cos_ftz(f2_lo(r)), cos_ftz(f2_hi(r)));
  const f2 sn = f2_make(sin_ftz(f2_lo(r)), sin_ftz(f2_hi(r)));
  yr = f2_mul(m, cs);
  yi = f2_mul(m, sn);
}
// first layer / real z: zi = 0
__device__ __forceinline__ void gabor_real_x2(const GaborConst2& c, f2 z, f2 wnorm, f2& yr, f2& yi) {
  const f2 t = f2_fma(z, z, wnorm);
  const f2 arg = f2_mul(c.c_t, t);
  const f2 r = gabor_phase_x2(c, z);
  const f2 m = f2_make(ex2_ftz(f2_lo(arg)), ex2_ftz(f2_hi(arg)));
  const f2 cs = f2_make(cos_ftz(f2_lo(r)), cos_ftz(f2_hi(r)));
  const f2 sn = f2_make(sin_ftz(f2_lo(r)), sin_ftz(f2_hi(r)));
  yr = f2_mul(m, cs);
  yi = f2_mul(m, sn);
}

// hidden-layer Wirtinger backward for a pair: p = conj(y) g ; g_z = -j w p - 2 s2 z Re p ; returns Re p (for g_w)
__device__ __forceinline__ f2 gabor_bwd_x2(const GaborConst2& c, f2 yr, f2 yi, f2 zr, f2 zi, f2 gr, f2 gi, f2& gzr, f2& gzi) {
  const f2 pr = f2_fma(yi, gi, f2_mul(yr, gr));
  const f2 pi = f2_fma(f2_mul(yi, c.none), gr, f2_mul(yr, gi));
  const f2 t = f2_mul(c.m2s2, pr);
  gzr = f2_fma(c.omega, pi, f2_mul(t, zr));
  gzi = f2_fma(c.nomega, pr, f2_mul(t, zi));
  return pr;
}
// The same with Im p handed out as well (trainable omega_0 / scale_0: g_omega0 = sum Im(conj(z) p), g_scale0 = -2 s0 sum (|z|^2 + |w|^2) Re p)
__device__ __forceinline__ f2 gabor_bwd_x2_p(const GaborConst2& c, f2 yr, f2 yi, f2 zr, f2 zi, f2 gr, f2 gi, f2& gzr, f2& gzi, f2& pi) {
  const f2 pr = f2_fma(yi, gi, f2_mul(yr, gr));
  pi = f2_fma(f2_mul(yi, c.none), gr, f2_mul(yr, gi));
  const f2 t = f2_mul(c.m2s2, pr);
  gzr = f2_fma(c.omega, pi, f2_mul(t, zr));
  gzi = f2_fma(c.nomega, pr, f2_mul(t, zi));
  return pr;
}
__device__ __forceinline__ f2 gabor_first_bwd_x2_p(const GaborConst2& c, f2 yr, f2 yi, f2 z, f2 gr, f2 gi, f2& gz, f2& pi) {
  const f2 pr = f2_fma(yi, gi, f2_mul(yr, gr));
  pi = f2_fma(f2_mul(yi, c.none), gr, f2_mul(yr, gi));
  gz = f2_fma(c.omega, pi, f2_mul(f2_mul(c.m2s2, pr), z));
  return pr;
}
// first layer (real z): g_z = w Im p - 2 s2 z Re p ; returns Re p
__device__ __forceinline__ f2 gabor_first_bwd_x2(const GaborConst2& c, f2 yr, f2 yi, f2 z, f2 gr, f2 gi, f2& gz) {
  const f2 pr = f2_fma(yi, gi, f2_mul(yr, gr));
  const f2 pi = f2_fma(f2_mul(yi, c.none), gr, f2_mul(yr, gi));
  gz = f2_fma(c.omega, pi, f2_mul(f2_mul(c.m2s2, pr), z));
  return pr;
}

}  // namespace wire
