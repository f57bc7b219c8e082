// tc_rows.cuh — the row-tile GEMM of the WIRE hot path on tcgen05 (sm_100a).
//
//   ACC[128 coords, nb cols] = A[128 coords, K] (K-major, TMA) x Bpacked[nb cols, K] (K-major, TMA)
//
// with TF32 inputs, FP32 accumulation in TMEM, and the layer's point-wise work fused in the
// epilogue straight out of TMEM.  One kernel template, several epilogues:
//
//   MODE_PLAIN        out0 = ACC                                         (probe / tests)
//   MODE_GABOR_FWD    z = ACC + b ; y = gabor(z)          -> y, z        modules/wire.py:88-93
//   MODE_GABOR2D_FWD  z,w = ACC halves + b1,b2 ; y = gabor2d(z,w) -> y,z,w   modules/wire2d.py:56-67
//   MODE_GABOR_BWD    g_y = ACC ; g_z = gabor'(z_saved, g_y)  -> g_z     (autograd of wire.py:88-93)
//   MODE_GABOR2D_BWD  same + g_w                               -> g_z, g_w
//   MODE_FIRST_BWD / MODE_FIRST2D_BWD   g_y0 = ACC ; z0 recomputed from coords -> real g_z0 (g_w0)
//
// A is the complex activation tensor seen as real [N, 2M] (interleaved re,im = torch complex64),
// Bpacked is the real 2x2-block expansion of the complex weight (see pack kernels), so one real
// GEMM of width 2M x 2M *is* the complex GEMM (8*M^2 flop/coord, no de-interleave anywhere).
//
// Warp roles (192 threads): warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc),
// warps 2..5 = epilogue (TMEM sub-partition = warp_idx & 3).
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

constexpr int kMaxIn = 8;    // coordinate dimensions supported by the fused first-layer epilogue
constexpr int kMaxOut = 4;   // output features supported by the fused final Linear
constexpr int kRowsThreads = 192;
constexpr int kTileRows = 128;
constexpr int kChunk = 32;   // fp32 columns per 128-byte swizzle row

struct RowsParams {
  CUtensorMap a_map[2];  // A parts, box {32 cols, 128 rows}
  CUtensorMap b_map;     // packed B [n_blocks*nb, Kpad], box {32 cols, b_box_rows}
  CUtensorMap o_map[3];  // outputs, box {32 cols, 32 rows}
  int n_rows;
  int k_cols[2];   // valid K columns in each A part (part 1 may be 0)
  int n_blocks;    // column blocks (work item = row tile x block)
  int nb;          // accumulator columns per block (multiple of 16, <= 512)
  int nbh;         // 2D fwd: columns of the z half (w half follows); otherwise == nb
  int n_cols;      // valid real output columns (2M)
  int b_box_rows, b_boxes;
  int stages;
  int store_mask;  // which epilogue results are TMA-stored: bit0 = y / g_z / plain, bit1 = z / g_w, bit2 = w;
                   // o_map[] slots are consumed in bit order
  int round_out0;  // round output 0 to TF32 (it feeds the next GEMM)
  const float* bias;
  const float* bias2;
  const float* omega;  // device scalars of the layer whose nonlinearity runs in the epilogue
  const float* scale;
  const float* z_src;  // saved pre-activations for the backward epilogues
  const float* w_src;
  int zw_pitch;
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
};

__device__ __forceinline__ void stage_row(uint32_t buf, int lane, const float (&v)[32]) {
  const uint32_t row = buf + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t addr = row + ((j ^ sw) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j]),
                 "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                 : "memory");
  }
}

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

template <int MODE>
__global__ void __launch_bounds__(kRowsThreads, 1) tc_rows_kernel(const __grid_constant__ RowsParams P) {
  using namespace sm100;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8];
  __shared__ __align__(8) uint64_t bar_empty[8];
  __shared__ __align__(8) uint64_t bar_tmem_full;
  __shared__ __align__(8) uint64_t bar_tmem_empty;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = kTileRows * 128;
  const uint32_t b_bytes = uint32_t(P.nb) * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t staging_base = smem_base + P.stages * stage_bytes;

  const int kc0 = (P.k_cols[0] + kChunk - 1) / kChunk;
  const int kc1 = (P.k_cols[1] + kChunk - 1) / kChunk;
  const int kc_total = kc0 + kc1;
  const int row_tiles = (P.n_rows + kTileRows - 1) / kTileRows;
  const int n_items = row_tiles * P.n_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_tmem_full), 1);
    mbar_init(smem_u32(&bar_tmem_empty), 4);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.a_map[0]);
    tma_prefetch_desc(&P.b_map);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int row0 = (item / P.n_blocks) * kTileRows;
        const int n0 = (item % P.n_blocks) * P.nb;
        for (int kc = 0; kc < kc_total; ++kc) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t full = smem_u32(&bar_full[stage]);
          mbar_expect_tx(full, stage_bytes);
          const uint32_t a_dst = smem_base + stage * stage_bytes;
          const int part = kc < kc0 ? 0 : 1;
          const int kcol = (part ? kc - kc0 : kc) * kChunk;
          tma_load_2d_hint(a_dst, &P.a_map[part], full, kcol, row0, kEvictFirst);
          for (int bx = 0; bx < P.b_boxes; ++bx)
            tma_load_2d_hint(a_dst + a_bytes + bx * P.b_box_rows * 128, &P.b_map, full, kc * kChunk,
                             n0 + bx * P.b_box_rows, kEvictLast);
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const int n1 = P.nb > 256 ? 256 : P.nb;
      const int n2 = P.nb - n1;
      const uint32_t idesc1 = make_idesc_tf32(128, n1, false, false);
      const uint32_t idesc2 = make_idesc_tf32(128, n2 > 0 ? n2 : 16, false, false);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tphase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        if (it > 0) {
          mbar_wait(smem_u32(&bar_tmem_empty), tphase);
          tphase ^= 1;
        }
        tc_fence_after();
        for (int kc = 0; kc < kc_total; ++kc) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_base = smem_base + stage * stage_bytes;
          const uint32_t b_base = a_base + a_bytes;
          const int part = kc < kc0 ? 0 : 1;
          const int kcol = (part ? kc - kc0 : kc) * kChunk;
          int steps = (P.k_cols[part] - kcol + 7) >> 3;
          steps = steps > 4 ? 4 : steps;
          for (int ks = 0; ks < steps; ++ks) {
            const uint32_t acc = (kc | ks) ? 1u : 0u;
            const uint64_t adesc = make_sdesc_sw128(a_base + ks * 32, 16, 1024);
            const uint64_t bdesc = make_sdesc_sw128(b_base + ks * 32, 16, 1024);
            umma_tf32(tmem_base, adesc, bdesc, idesc1, acc);
            if (n2 > 0) {
              const uint64_t bdesc2 = make_sdesc_sw128(b_base + n1 * 128 + ks * 32, 16, 1024);
              umma_tf32(tmem_base + n1, adesc, bdesc2, idesc2, acc);
            }
          }
          umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&bar_tmem_full));
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int ew = warp - 2;
    const int n_out = __popc(P.store_mask);
    const uint32_t wbuf = staging_base + ew * (n_out * 2 * 4096);
    const float omega = (MODE != MODE_PLAIN) ? __ldg(P.omega) : 0.f;
    const float sc = (MODE != MODE_PLAIN) ? __ldg(P.scale) : 0.f;
    const float s2 = sc * sc;
    uint32_t tphase = 0;
    uint32_t bufsel = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int row0 = (item / P.n_blocks) * kTileRows;
      const int blk = item % P.n_blocks;
      const int row = row0 + q * 32 + lane;
      const bool row_ok = row < P.n_rows;
      // columns of this block in output space
      const int ncol_blk = (MODE == MODE_GABOR2D_FWD) ? P.nbh : P.nb;
      const int col0 = blk * ncol_blk;
      int valid = P.n_cols - col0;
      valid = valid > ncol_blk ? ncol_blk : valid;
      const int nchunks = (valid + kChunk - 1) / kChunk;

      float cin[kMaxIn];
      if constexpr (MODE == MODE_FIRST_BWD || MODE == MODE_FIRST2D_BWD) {
#pragma unroll
        for (int d = 0; d < kMaxIn; ++d)
          cin[d] = (d < P.in_features && row_ok) ? __ldg(P.coords + size_t(row) * P.in_features + d) : 0.f;
      }
      float facc[kMaxOut];
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) facc[o] = 0.f;

      mbar_wait(smem_u32(&bar_tmem_full), tphase);
      tphase ^= 1;
      tc_fence_after();

      for (int ch = 0; ch < nchunks; ++ch) {
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + ch * kChunk;
        uint32_t raw[32];
        float v2[32];
        tmem_ld32(taddr, raw);
        if constexpr (MODE == MODE_GABOR2D_FWD) {
          uint32_t raw2[32];
          tmem_ld32(taddr + P.nbh, raw2);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v2[i] = __uint_as_float(raw2[i]);
        } else {
          tmem_wait_ld();
        }
        if (ch == nchunks - 1) {
          // all TMEM reads of this tile are done: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty));
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        const int c = col0 + ch * kChunk;  // first real output column of this chunk

        float o0[32], o1[32], o2[32];
        if constexpr (MODE == MODE_PLAIN) {
#pragma unroll
          for (int i = 0; i < 32; ++i) o0[i] = v[i];
        } else if constexpr (MODE == MODE_GABOR_FWD || MODE == MODE_GABOR2D_FWD) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int cc = c + 2 * i;
            const bool ok = cc < P.n_cols;
            const float zr = v[2 * i] + (ok ? __ldg(P.bias + cc) : 0.f);
            const float zi = v[2 * i + 1] + (ok ? __ldg(P.bias + cc + 1) : 0.f);
            float extra = 0.f;
            if constexpr (MODE == MODE_GABOR2D_FWD) {
              const float wr = v2[2 * i] + (ok ? __ldg(P.bias2 + cc) : 0.f);
              const float wi = v2[2 * i + 1] + (ok ? __ldg(P.bias2 + cc + 1) : 0.f);
              extra = s2 * (wr * wr + wi * wi);
              o2[2 * i] = wr;
              o2[2 * i + 1] = wi;
            }
            float yr, yi;
            gabor_fwd<true>(zr, zi, omega, s2, extra, yr, yi);
            if (P.fuse_final && ok) {
              const int k = cc >> 1;
#pragma unroll
              for (int o = 0; o < kMaxOut; ++o) {
                if (o < P.out_features) {
                  const float2 wv = __ldg(reinterpret_cast<const float2*>(P.wf) + size_t(o) * (P.n_cols >> 1) + k);
                  facc[o] = fmaf(yr, wv.x, fmaf(-yi, wv.y, facc[o]));
                }
              }
            }
            if (P.round_out0) { yr = round_tf32(yr); yi = round_tf32(yi); }
            o0[2 * i] = yr;
            o0[2 * i + 1] = yi;
            o1[2 * i] = zr;
            o1[2 * i + 1] = zi;
          }
        } else if constexpr (MODE == MODE_GABOR_BWD || MODE == MODE_GABOR2D_BWD) {
          float z[32];
          load_row32(P.z_src + size_t(row) * P.zw_pitch + c, row_ok, z);
          float w[32];
          if constexpr (MODE == MODE_GABOR2D_BWD) load_row32(P.w_src + size_t(row) * P.zw_pitch + c, row_ok, w);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float zr = z[2 * i], zi = z[2 * i + 1];
            float extra = 0.f;
            if constexpr (MODE == MODE_GABOR2D_BWD) extra = s2 * (w[2 * i] * w[2 * i] + w[2 * i + 1] * w[2 * i + 1]);
            float yr, yi, gzr, gzi;
            gabor_fwd<true>(zr, zi, omega, s2, extra, yr, yi);
            const float pr = gabor_bwd(yr, yi, zr, zi, v[2 * i], v[2 * i + 1], omega, s2, gzr, gzi);
            o0[2 * i] = round_tf32(gzr);
            o0[2 * i + 1] = round_tf32(gzi);
            if constexpr (MODE == MODE_GABOR2D_BWD) {
              const float t = -2.0f * s2 * pr;
              o1[2 * i] = round_tf32(t * w[2 * i]);
              o1[2 * i + 1] = round_tf32(t * w[2 * i + 1]);
            }
          }
        } else {  // MODE_FIRST_BWD / MODE_FIRST2D_BWD: real z0 recomputed from coordinates
          float gz[16], gw[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int j = (c >> 1) + i;  // complex feature index
            const bool ok = (2 * j) < P.n_cols;
            float z0 = ok ? __ldg(P.b0 + j) : 0.f;
            float w0v = 0.f;
#pragma unroll
            for (int d = 0; d < kMaxIn; ++d)
              if (d < P.in_features && ok) z0 = fmaf(cin[d], __ldg(P.w0 + size_t(j) * P.in_features + d), z0);
            float extra = 0.f;
            if constexpr (MODE == MODE_FIRST2D_BWD) {
              w0v = ok ? __ldg(P.b0b + j) : 0.f;
#pragma unroll
              for (int d = 0; d < kMaxIn; ++d)
                if (d < P.in_features && ok) w0v = fmaf(cin[d], __ldg(P.w0b + size_t(j) * P.in_features + d), w0v);
              extra = s2 * w0v * w0v;
            }
            float yr, yi;
            gabor_fwd<true>(z0, 0.f, omega, s2, extra, yr, yi);
            const float pr = gabor_first_bwd(yr, yi, z0, v[2 * i], v[2 * i + 1], omega, s2, gz[i]);
            gw[i] = -2.0f * s2 * pr * w0v;
          }
          if (row_ok) {
            // 16 real outputs per thread = 64 contiguous bytes (two full sectors): direct stores
            float4* dst = reinterpret_cast<float4*>(P.gz0 + size_t(row) * P.gz0_pitch + (c >> 1));
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              if ((c >> 1) + 4 * j4 < P.gz0_pitch) dst[j4] = make_float4(gz[4 * j4], gz[4 * j4 + 1], gz[4 * j4 + 2], gz[4 * j4 + 3]);
            if constexpr (MODE == MODE_FIRST2D_BWD) {
              float4* dw = reinterpret_cast<float4*>(P.gw0 + size_t(row) * P.gz0_pitch + (c >> 1));
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4)
                if ((c >> 1) + 4 * j4 < P.gz0_pitch) dw[j4] = make_float4(gw[4 * j4], gw[4 * j4 + 1], gw[4 * j4 + 2], gw[4 * j4 + 3]);
            }
          }
        }

        if (n_out > 0) {
          // staging (double-buffered per warp) -> TMA store; OOB rows/cols are clipped by TMA
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          const uint32_t sb = wbuf + bufsel * (n_out * 4096);
          int slot = 0;
          if (P.store_mask & 1) { stage_row(sb, lane, o0); ++slot; }
          if constexpr (MODE == MODE_GABOR_FWD || MODE == MODE_GABOR2D_FWD || MODE == MODE_GABOR2D_BWD) {
            if (P.store_mask & 2) { stage_row(sb + slot * 4096, lane, o1); ++slot; }
          }
          if constexpr (MODE == MODE_GABOR2D_FWD) {
            if (P.store_mask & 4) { stage_row(sb + slot * 4096, lane, o2); ++slot; }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            for (int s = 0; s < n_out; ++s) tma_store_2d(&P.o_map[s], sb + s * 4096, c, row0 + q * 32);
            tma_store_commit();
          }
          bufsel ^= 1;
        }
      }

      if constexpr (MODE == MODE_GABOR_FWD || MODE == MODE_GABOR2D_FWD) {
        if (P.fuse_final && row_ok) {
#pragma unroll
          for (int o = 0; o < kMaxOut; ++o)
            if (o < P.out_features) P.out[size_t(row) * P.out_features + o] = facc[o] + __ldg(P.bf + 2 * o);
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  // ===================== teardown =====================
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace wire
