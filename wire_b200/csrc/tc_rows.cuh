// tc_rows.cuh — the row-tile GEMM of the WIRE hot path on tcgen05 (sm_100a).
//
//   ACC[128 coords, nb cols] = A[128 coords, K] (K-major, TMA) x Bpacked[nb cols, K] (K-major, TMA)
//
// with TF32 inputs, FP32 accumulation in TMEM, and the layer's point-wise work (rows_epilogue.cuh)
// fused in the epilogue straight out of TMEM; results leave through swizzled staging + TMA stores.
//
// A is the complex activation tensor seen as real [N, 2M] (interleaved re,im = torch complex64),
// Bpacked is the real 2x2-block expansion of the complex weight (pack_weights_kernel), so one real
// GEMM of width 2M x 2M *is* the complex GEMM (8*M^2 flop/coord, no de-interleave anywhere).
//
// Warp roles (192 threads): warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc),
// warps 2..5 = epilogue (TMEM sub-partition = warp_idx & 3).
#pragma once
#include "rows_epilogue.cuh"
#include "sm100.cuh"

namespace wire {

constexpr int kRowsThreads = 192;
constexpr int kTileRows = 128;
constexpr int kChunk = 32;  // fp32 columns per 128-byte swizzle row

struct RowsParams {
  CUtensorMap a_map[2];  // A parts, box {32 cols, 128 rows}
  CUtensorMap b_map;     // packed B [n_blocks*nb, Kpad], box {32 cols, b_box_rows}
  CUtensorMap o_map[3];  // outputs, box {32 cols, 32 rows}
  int k_cols[2];         // valid K columns in each A part (part 1 may be 0)
  int n_blocks;          // column blocks (work item = row tile x block)
  int nb;                // accumulator columns per block (multiple of 16, <= 512)
  int nbh;               // 2D fwd: columns of the z half (w half follows); otherwise == nb
  int b_box_rows, b_boxes;
  int stages;
  int store_mask;  // which epilogue results are TMA-stored: bit0 = o0, bit1 = o1, bit2 = o2;
                   // o_map[] slots are consumed in bit order
  RowsEpi e;
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
  const int row_tiles = (P.e.n_rows + kTileRows - 1) / kTileRows;
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
    const RowsEpi& E = P.e;
    const int q = warp & 3;
    const int ew = warp - 2;
    const int n_out = __popc(P.store_mask);
    const uint32_t wbuf = staging_base + ew * (n_out * 2 * 4096);
    const float omega = (MODE != MODE_PLAIN) ? __ldg(E.omega) : 0.f;
    const float sc = (MODE != MODE_PLAIN) ? __ldg(E.scale) : 0.f;
    const float s2 = sc * sc;
    uint32_t tphase = 0;
    uint32_t bufsel = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int row0 = (item / P.n_blocks) * kTileRows;
      const int blk = item % P.n_blocks;
      const int row = row0 + q * 32 + lane;
      const bool row_ok = row < E.n_rows;
      const int ncol_blk = (MODE == MODE_GABOR2D_FWD) ? P.nbh : P.nb;  // output columns per block
      const int col0 = blk * ncol_blk;
      int valid = E.n_cols - col0;
      valid = valid > ncol_blk ? ncol_blk : valid;
      const int nchunks = (valid + kChunk - 1) / kChunk;

      float cin[kMaxIn];
      rows_load_coords<MODE>(E, row, row_ok, cin);
      float facc[kMaxOut];
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) facc[o] = 0.f;

      mbar_wait(smem_u32(&bar_tmem_full), tphase);
      tphase ^= 1;
      tc_fence_after();

      for (int ch = 0; ch < nchunks; ++ch) {
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + ch * kChunk;
        uint32_t raw[32];
        float v[32], v2[32];
        tmem_ld32(taddr, raw);
        if constexpr (MODE == MODE_GABOR2D_FWD) {
          uint32_t raw2[32];
          tmem_ld32(taddr + P.nbh, raw2);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v2[i] = __uint_as_float(raw2[i]);
        } else {
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v2[i] = 0.f;
        }
        if (ch == nchunks - 1) {
          // all TMEM reads of this tile are done: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        const int c = col0 + ch * kChunk;  // first real output column of this chunk

        float o0[32], o1[32], o2[32];
        rows_epilogue_chunk<MODE, true>(E, row, row_ok, c, omega, s2, v, v2, cin, facc, o0, o1, o2);

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
      rows_store_final<MODE>(E, row, row_ok, facc);
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
