// tc_rows16.cuh — the row-tile GEMM of the WIRE hot path for 16-bit tensors (mixed16 precision) on tcgen05.
//
//   ACC[128 coords, ns cols] = A[128 coords, K] (FP16 or BF16, K-major, TMA) x Bpacked[ns cols, K] (same format)
//
// MMA kind::f16, FP32 accumulation in TMEM.  Producer / MMA-issue roles and the TMEM double buffering are those of
// tc_rows.cuh (a pipeline stage covers 64 K columns instead of 32).  The epilogue is rebuilt for the regime the 16-bit
// path lives in: with half the operand bytes and twice the MMA rate, the point-wise Gabor work — not HBM, not the
// tensor pipe — bounded the TF32-style epilogue (profiles/r01_ncu_rows16_v1_*: 8 epilogue warps, issue slots 28 % busy,
// MMA thread waiting for a free accumulator 52 % of the time).  So here:
//   * 16 epilogue warps (four per TMEM sub-partition, 576 threads; registers are granted per 4 warps, so 96 per thread), chunks dealt round-robin with a
//     per-job rotation so the 7 chunks of a 224-column slice balance over 4 warps in the long run;
//   * results are converted and staged 4 complex features at a time (one 16-byte st.shared per output), so a thread
//     never holds more than the 32 accumulator registers of its chunk;
//   * 16-bit staging / saved-z tiles use the 64-byte TMA swizzle (conflict-free 16-byte accesses at a 64-byte pitch);
//   * layer flags are template parameters (no per-element branches), rint() is the add-magic-constant trick (FRND
//     shares the 16-lane XU pipe with ex2/sin/cos);
//   * the FP32 math runs on PAIRS of features with the packed FFMA2 / FMUL2 / FADD2 instructions (f32x2.cuh).  To make
//     that free of register shuffles the packed weight matrices of the 16-bit path put their rows in "pair-transposed"
//     order (acc_col_perm): accumulator columns 4j..4j+3 hold (re A, re B, im A, im B) of features A = 2j, B = 2j+1,
//     so tcgen05.ld delivers aligned register pairs; tensors in HBM keep torch's interleaved (re, im) order.
//
// Modes and math are exactly those of tc_rows.cuh / rows_epilogue.cuh (same reference lines).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "f32x2.cuh"
#include "tc_rows.cuh"

namespace wire {


constexpr int kEpi16Warps = 16;
constexpr int kEpi16Parts = kEpi16Warps / 4;  // warps per TMEM sub-partition
constexpr int kRows16Threads = 64 + 32 * kEpi16Warps;
constexpr int kTile16Bytes = 2048;  // one 32 x 32 tile of 16-bit elements

// 32x32 16-bit tile, TMA SWIZZLE_64B: row r at r*64, 16-byte chunk j stored at chunk (j ^ ((r >> 1) & 3))
__device__ __forceinline__ uint32_t sw64_addr(uint32_t tile, int lane, int j) {
  return tile + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  const __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_f16(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

// y = exp(j w z - s2 (|z|^2 + wnorm)) with the phase reduced in turns; rint(u) = (u + 1.5*2^23) - 1.5*2^23 (|u| < 2^22)
__device__ __forceinline__ void gabor16(const GaborConst& g, float zr, float zi, float wnorm, float& yr, float& yi) {
  const float t = fmaf(zi, zi, fmaf(zr, zr, wnorm));
  const float m = ex2_ftz(fmaf(g.c_t, t, g.c_zi * zi));
  const float u = zr * g.c_turn;
  const float k = __fadd_rn(__fadd_rn(u, 12582912.0f), -12582912.0f);
  const float r = (u - k) * 6.283185307179586f;
  yr = m * cos_ftz(r);
  yi = m * sin_ftz(r);
}

// FUSE: the final Linear (.real) is accumulated in this epilogue (GABOR_FWD / GABOR2D_FWD only)
// SCAL: backward modes only -- the gradients of the epilogue layer's own omega_0 / scale_0 are accumulated on the way
//       (per-thread partial sums over everything the thread touches, one warp reduction + two atomics per warp at the end)
// FUSE: 0 = the layer's y is stored; 3 / 4 = the final Linear (up to 3 / 4 real outputs) is applied in the epilogue instead: with three
// outputs (every image driver) the weight table is 12 floats per feature pair and a pair costs 6 packed FMAs + 3 LDS.128, not 8 + 4
template <int MODE, bool PAIR, int FUSE = 0, bool SCAL = false>
__global__ void __launch_bounds__(kRows16Threads, 1) tc_rows16_kernel(const __grid_constant__ RowsParams P) {
  using namespace sm100;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8];
  __shared__ __align__(8) uint64_t bar_empty[8];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ __align__(8) uint64_t bar_in[kEpi16Warps];
  __shared__ __align__(8) uint64_t bar_b;  // B-stationary schedule: the resident weight slice has landed
  __shared__ uint32_t tmem_slot;

  constexpr bool kFwd = (MODE == MODE_GABOR_FWD || MODE == MODE_GABOR2D_FWD);
  constexpr bool kBwd = (MODE == MODE_GABOR_BWD || MODE == MODE_GABOR2D_BWD);
  constexpr bool kFirst = (MODE == MODE_FIRST_BWD || MODE == MODE_FIRST2D_BWD);
  constexpr bool k2D = (MODE == MODE_GABOR2D_FWD || MODE == MODE_GABOR2D_BWD || MODE == MODE_FIRST2D_BWD);
  static_assert(!FUSE || kFwd, "only the forward modes fuse the final Linear");
  static_assert(!SCAL || kBwd || kFirst, "omega_0 / scale_0 gradients belong to the backward modes");
  constexpr int kKC = 64;  // K columns per pipeline stage (one 128 B swizzle row of 16-bit elements)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_bytes = kTileRows * 128;
  const uint32_t b_bytes = uint32_t(P.b_box_rows) * 128;
  const bool bstat = P.bstat != 0;
  const uint32_t stage_bytes = bstat ? a_bytes : a_bytes + b_bytes;   // B-stationary: the pipeline stages hold A tiles only
  const uint32_t bres_base = smem_base + P.bres_off;
  const uint32_t staging_base = smem_base + P.staging_off;
  float* params = reinterpret_cast<float*>(smem_gen + P.param_off);
  const RowsEpi& E = P.e;

  const int kc0 = (P.k_cols[0] + kKC - 1) / kKC;
  const int kc1 = (P.k_cols[1] + kKC - 1) / kKC;
  const int kc_total = kc0 + kc1;
  const int row_tiles = (E.n_rows + kTileRows - 1) / kTileRows;
  const int n_feat = E.n_cols >> 1;
  const int C = P.cluster;
  const int crank = int(cluster_ctarank());
  const int n_clusters = gridDim.x / C;
  const int my_cluster = blockIdx.x / C;
  const int row_groups = (row_tiles + C - 1) / C;
  const int n_units = row_groups * P.n_blocks;
  // B-stationary: cluster c owns slice c % slices and walks the row groups c / slices, c / slices + cps, ... (n_blocks == 1, so a
  // unit is a row group); the launcher makes the cluster count a multiple of the slice count
  const int cps = n_clusters / P.slices;
  const int my_sl = bstat ? my_cluster % P.slices : 0;
  const int my_seq = my_cluster / P.slices;
  const int n_iters = bstat ? (cps > 0 && my_seq < cps ? (row_groups + cps - 1) / cps : 0) : (n_units + n_clusters - 1) / n_clusters;
  const int n_jobs = bstat ? n_iters : n_iters * P.slices;
  auto job_it = [&](int jb) { return bstat ? jb : jb / P.slices; };
  auto job_sl = [&](int jb) { return bstat ? my_sl : jb % P.slices; };
  constexpr bool pair = PAIR;
  const bool leader = crank == 0;
  unsigned long long* dbg = P.dbg ? P.dbg + size_t(blockIdx.x) * 16 : nullptr;  // [0..7] stall counters, [8..11] phase stamps
  const long long t_entry = WIRE_CLK();
  // work unit of iteration `it`; reversed sweeps mirror the valid units (phantom units past the end stay phantom)
  auto unit_of = [&](int it) {
    if (bstat) return it * cps + my_seq;
    const int u = it * n_clusters + my_cluster;
    return (P.reverse && u < n_units) ? n_units - 1 - u : u;
  };

  // ---- shared parameter tables (zero padded: the epilogue needs no column checks) ----
  //  fwd  : bias[param_cols] | bias2[param_cols] (2D) | wf[(param_cols/2)][8] (wr[4], wi[4]) | fin exchange (FUSE)
  //  first: tab[(param_cols/2)][4] = {w0[0..2], b0} | tab2 (2D)
  //  fwd  : bias[param_cols] (pair-transposed order) | bias2 (2D) | wf[pairs][16] | fin exchange (FUSE)
  //         wf pair P = features (2P, 2P+1): float4 {wr0A,wr0B,wr1A,wr1B} {wr2A,wr2B,wr3A,wr3B} {-wi0A,-wi0B,-wi1A,-wi1B} {-wi2..}
  //  first: tab[pairs][8] = float4 {w0A,w0B,w1A,w1B} {w2A,w2B,bA,bB} | tab2 (2D)
  if constexpr (kFwd) {
    for (int i = threadIdx.x; i < P.param_cols; i += blockDim.x) {
      const int l = acc_col_perm(i);
      params[i] = l < E.n_cols ? E.bias[l] : 0.f;
      if constexpr (k2D) params[P.param_cols + i] = l < E.n_cols ? E.bias2[l] : 0.f;
    }
    if constexpr (FUSE == 3) {
      // pair P = features (2P, 2P+1): float4 {wr0A,wr0B,wr1A,wr1B} {wr2A,wr2B,-wi0A,-wi0B} {-wi1A,-wi1B,-wi2A,-wi2B}
      float* wf = params + (k2D ? 2 : 1) * P.param_cols;
      for (int i = threadIdx.x; i < (P.param_cols >> 2) * 12; i += blockDim.x) {
        const int pr = i / 12, j = i - pr * 12;
        const int slot = j >> 1, o = slot % 3, im = slot / 3, k = 2 * pr + (j & 1);
        float v = 0.f;
        if (k < n_feat && o < E.out_features) v = E.wf[(size_t(o) * n_feat + k) * 2 + im];
        wf[i] = im ? -v : v;
      }
    } else if constexpr (FUSE != 0) {
      float* wf = params + (k2D ? 2 : 1) * P.param_cols;
      for (int i = threadIdx.x; i < (P.param_cols >> 2) * 16; i += blockDim.x) {
        const int pr = i >> 4, qd = (i >> 2) & 3, e = i & 3;
        const int o = (qd & 1) * 2 + (e >> 1), k = 2 * pr + (e & 1), im = qd >> 1;
        float v = 0.f;
        if (k < n_feat && o < E.out_features) v = E.wf[(size_t(o) * n_feat + k) * 2 + im];
        wf[i] = im ? -v : v;
      }
    }
  }
  if constexpr (kFirst) {
    for (int i = threadIdx.x; i < (P.param_cols >> 2) * 8; i += blockDim.x) {
      const int pr = i >> 3, qd = (i >> 2) & 1, e = i & 3;
      const int k = 2 * pr + (e & 1);
      const int d = qd * 2 + (e >> 1);  // 0..2 = coordinate weight, 3 = bias
      float v = 0.f, v2 = 0.f;
      if (k < n_feat) {
        if (d < 3) {
          if (d < E.in_features) { v = E.w0[size_t(k) * E.in_features + d]; if constexpr (k2D) v2 = E.w0b[size_t(k) * E.in_features + d]; }
        } else { v = E.b0[k]; if constexpr (k2D) v2 = E.b0b[k]; }
      }
      params[i] = v;
      if constexpr (k2D) params[(P.param_cols >> 2) * 8 + i] = v2;
    }
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tmem_full[b]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[b]), kEpi16Warps * C);
    }
    for (int w = 0; w < kEpi16Warps; ++w) mbar_init(smem_u32(&bar_in[w]), 1);
    mbar_init(smem_u32(&bar_b), 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.a_map[0]);
    tma_prefetch_desc(&P.b_map);
  }
  if (warp == 1) {
    if (pair) { tmem_alloc_2cta(smem_u32(&tmem_slot), 512); tmem_relinquish_2cta(); }
    else { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();  // the next kernel of the step may be scheduled (it becomes resident only as these CTAs exit)
  pdl_wait();     // everything above read parameters only; the tensors of earlier kernels are read below
  if (dbg && threadIdx.x == 0) dbg[8] = (unsigned long long)(WIRE_CLK() - t_entry);  // prologue

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long t_wait = 0;
      const long long t_begin = WIRE_CLK();
      if (bstat && n_jobs > 0) {  // the resident weight slice: every K stage, once
        const uint32_t bb = smem_u32(&bar_b);
        const int brow_res = my_sl * P.ns + crank * P.b_box_rows;
        if (!pair) {
          mbar_expect_tx(bb, kc_total * b_bytes);
          for (int kc = 0; kc < kc_total; ++kc) tma_load_2d_hint(bres_base + kc * b_bytes, &P.b_map, bb, kc * kKC, brow_res, kEvictLast);
        } else {
          if (leader) mbar_expect_tx(bb, 2 * kc_total * b_bytes);
          for (int kc = 0; kc < kc_total; ++kc) tma_load_2d_2cta(bres_base + kc * b_bytes, &P.b_map, bb & kPeerBitMask, kc * kKC, brow_res, kEvictLast);
        }
      }
      for (int jb = 0; jb < n_jobs; ++jb) {
        const int it = job_it(jb), sl = job_sl(jb);
        const int unit = unit_of(it);
        const int row0 = ((unit / P.n_blocks) * C + crank) * kTileRows;
        const int brow = (unit % P.n_blocks) * P.nb + sl * P.ns + crank * P.b_box_rows;
        const uint64_t a_policy = (!bstat && sl == P.slices - 1 && (unit % P.n_blocks) == P.n_blocks - 1) ? kEvictFirst : kEvictLast;
        if (!bstat && sl == 0 && P.l2_prefetch && it + 1 < n_iters) {
          // pull the NEXT work unit's A rows into L2 now: their demand loads then see L2 latency instead of HBM latency
          // (the pipeline holds only ~5 stages; HBM latency under load starved the MMA thread, mma_wait_full 52 %)
          const int unit_n = unit_of(it + 1);
          if ((unit_n % P.n_blocks) == 0) {
            const int row_n = ((unit_n / P.n_blocks) * C + crank) * kTileRows;
            for (int kc = 0; kc < kc0; ++kc) tma_prefetch_l2_2d(&P.a_map[0], kc * kKC, row_n);
            for (int kc = 0; kc < kc1; ++kc) tma_prefetch_l2_2d(&P.a_map[1], kc * kKC, row_n);
          }
        }
        for (int kc = 0; kc < kc_total; ++kc) {
          const long long t0 = WIRE_CLK();
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          t_wait += WIRE_CLK() - t0;
          const uint32_t full_own = smem_u32(&bar_full[stage]);
          const uint32_t a_dst = smem_base + stage * stage_bytes;
          const int part = kc < kc0 ? 0 : 1;
          const int kcol = (part ? kc - kc0 : kc) * kKC;
          if (!pair) {
            mbar_expect_tx(full_own, stage_bytes);
            tma_load_2d_hint(a_dst, &P.a_map[part], full_own, kcol, row0, a_policy);
            if (!bstat) tma_load_2d_hint(a_dst + a_bytes, &P.b_map, full_own, kc * kKC, brow, kEvictLast);
          } else {
            const uint32_t full_leader = full_own & kPeerBitMask;
            if (leader) mbar_expect_tx(full_own, 2 * stage_bytes);
            tma_load_2d_2cta(a_dst, &P.a_map[part], full_leader, kcol, row0, a_policy);
            if (!bstat) tma_load_2d_2cta(a_dst + a_bytes, &P.b_map, full_leader, kc * kKC, brow, kEvictLast);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg) { dbg[kDbgProdWaitEmpty] = t_wait; dbg[kDbgProdTotal] = WIRE_CLK() - t_begin; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && (!pair || leader)) {
      const uint32_t idesc = make_idesc_f16(pair ? 256 : 128, P.ns, false, false, uint32_t(P.a_fmt), uint32_t(P.b_fmt));
      const uint32_t desc_hi = uint32_t(make_sdesc_sw128(0, 16, 1024) >> 32);
      const uint32_t a_lo0 = uint32_t(make_sdesc_sw128(smem_base, 16, 1024));
      const uint32_t stage_units = stage_bytes >> 4, b_units = a_bytes >> 4;
      const uint32_t bres_lo0 = a_lo0 + (P.bres_off >> 4), bres_units = b_bytes >> 4;
      if (bstat && n_jobs > 0) { mbar_wait(smem_u32(&bar_b), 0); tc_fence_after(); }
      auto tail_steps = [](int cols, int kc) {
        const int st = (cols - (kc - 1) * kKC + 15) / 16;
        return st > 4 ? 4 : st;
      };
      const int last0 = tail_steps(P.k_cols[0], kc0);
      const int last1 = kc1 > 0 ? tail_steps(P.k_cols[1], kc1) : 4;
      int stage = 0;
      uint32_t phase = 0;
      long long w_full = 0, w_tmem = 0;
      const long long t_begin = WIRE_CLK();
      for (int jb = 0; jb < n_jobs; ++jb) {
        const int buf = jb & 1;
        if (jb >= 2) {
          const long long t0 = WIRE_CLK();
          mbar_wait(smem_u32(&bar_tmem_empty[buf]), ((jb >> 1) & 1) ^ 1);
          w_tmem += WIRE_CLK() - t0;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * P.buf_cols;
        for (int kc = 0; kc < kc_total; ++kc) {
          const long long t1 = WIRE_CLK();
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          w_full += WIRE_CLK() - t1;
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * stage_units;
          const uint32_t b_lo = bstat ? bres_lo0 + kc * bres_units : a_lo + b_units;
          const int steps = (kc == kc0 - 1) ? last0 : ((kc == kc_total - 1) ? last1 : 4);
          auto mma = [&](int ks, uint32_t acc) {
            const uint64_t adesc = (uint64_t(desc_hi) << 32) | (a_lo + 2 * ks);
            const uint64_t bdesc = (uint64_t(desc_hi) << 32) | (b_lo + 2 * ks);
            if (pair) umma_f16_2cta(d_tmem, adesc, bdesc, idesc, acc);
            else umma_f16(d_tmem, adesc, bdesc, idesc, acc);
          };
          if (steps == 4) {
            mma(0, kc ? 1u : 0u);
            mma(1, 1u);
            mma(2, 1u);
            mma(3, 1u);
          } else {
            for (int ks = 0; ks < steps; ++ks) mma(ks, (kc | ks) ? 1u : 0u);
          }
          if (pair) umma_commit_2cta_mcast(smem_u32(&bar_empty[stage]), 3);
          else umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        if (pair) umma_commit_2cta_mcast(smem_u32(&bar_tmem_full[buf]), 3);
        else umma_commit(smem_u32(&bar_tmem_full[buf]));
      }
      if (dbg) { dbg[kDbgMmaWaitFull] = w_full; dbg[kDbgMmaWaitTmem] = w_tmem; dbg[kDbgMmaTotal] = WIRE_CLK() - t_begin; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;   // 0..15
    const int q = warp & 3;    // TMEM sub-partition (lanes 32q..32q+31)
    const int part = ew >> 2;  // 0..3: which of the sub-partition's four warps
    // (a layer whose epilogue applies the final Linear never stores y: known at compile time, so its packed y is never kept)
    const bool st0 = FUSE ? false : bool(P.store_mask & 1), st1 = P.store_mask & 2, st2 = P.store_mask & 4;
    const int n_out = int(st0) + int(st1) + int(st2);
    const int slots = n_out + P.n_in;
    const uint32_t wbuf = staging_base + ew * (slots * kTile16Bytes);
    const uint32_t wbuf1 = wbuf + (st0 ? kTile16Bytes : 0);                      // slot of o1
    const uint32_t wbuf2 = wbuf1 + (st1 ? kTile16Bytes : 0);                     // slot of o2
    const uint32_t inbuf = wbuf + n_out * kTile16Bytes;                          // saved z (and w) tiles
    const GaborConst G = (MODE != MODE_PLAIN) ? make_gabor_const(__ldg(E.omega), __ldg(E.scale)) : make_gabor_const(0.f, 0.f);
    const float4* s_bias4 = reinterpret_cast<const float4*>(params);
    const float4* s_bias24 = reinterpret_cast<const float4*>(params + P.param_cols);
    const float4* s_wf = reinterpret_cast<const float4*>(params + (k2D ? 2 : 1) * P.param_cols);
    // FUSE: partial sums of parts 1..3 travel through [slot 2][part-1][128 rows] float4 after the wf table
    float4* s_fin = reinterpret_cast<float4*>(params + (k2D ? 2 : 1) * P.param_cols + (P.param_cols >> 1) * 8);
    const float4* s_tab = reinterpret_cast<const float4*>(params);
    const float4* s_tab2 = reinterpret_cast<const float4*>(params + (P.param_cols >> 2) * 8);
    const GaborConst2 G2 = make_gabor_const2(G);
    const uint32_t empty_addr0 = pair ? (smem_u32(&bar_tmem_empty[0]) & kPeerBitMask) : smem_u32(&bar_tmem_empty[0]);
    const uint32_t empty_addr1 = pair ? (smem_u32(&bar_tmem_empty[1]) & kPeerBitMask) : smem_u32(&bar_tmem_empty[1]);
    const int out_blk = (MODE == MODE_GABOR2D_FWD) ? P.nbh : P.nb;
    const int out_ns = (MODE == MODE_GABOR2D_FWD) ? P.nbh : P.ns;
    uint32_t in_phase = 0;
    long long e_wait = 0, e_wait_in = 0;
    const long long e_begin = WIRE_CLK();
    f2 cin2[3] = {0ull, 0ull, 0ull};          // first-layer modes: this row's coordinates, broadcast pairs
    f2 facc2[kMaxOut] = {0ull, 0ull, 0ull, 0ull};  // FUSE: partial sums of the final Linear over even / odd features
    f2 s_om = 0ull, s_sc = 0ull;               // SCAL: sum Im(conj(z) p) and sum (|z|^2 + |w|^2) Re p over this thread's elements
    auto release = [&](int buf) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (pair) mbar_arrive_cluster(buf ? empty_addr1 : empty_addr0); else mbar_arrive(buf ? empty_addr1 : empty_addr0); }
    };
    for (int jb = 0; jb < n_jobs; ++jb) {
      const int it = job_it(jb), sl = job_sl(jb);
      const int buf = jb & 1;
      const int unit = unit_of(it);
      const int row0 = ((unit / P.n_blocks) * C + crank) * kTileRows;
      const int blk = unit % P.n_blocks;
      const int row = row0 + q * 32 + lane;
      const bool row_ok = row < E.n_rows;
      const int col0 = blk * out_blk + sl * out_ns;
      int valid = E.n_cols - col0;
      valid = valid > out_ns ? out_ns : valid;
      const int nchunks = valid > 0 ? (valid + kChunk - 1) / kChunk : 0;
      // chunks of this warp: ch0, ch0 + 4, ... with a per-job rotation ((ch + jb) % 4 == part)
      const int ch0 = (part - jb) & (kEpi16Parts - 1);
      int my_last = -1;
      if (nchunks > ch0) my_last = ch0 + kEpi16Parts * ((nchunks - 1 - ch0) / kEpi16Parts);

      if (bstat || sl == 0) {  // a new row tile
        if constexpr (kFirst) {
          float c0 = 0.f, c1 = 0.f, c2 = 0.f;
          if (row_ok) {
            c0 = __ldg(E.coords + size_t(row) * E.in_features);
            if (E.in_features > 1) c1 = __ldg(E.coords + size_t(row) * E.in_features + 1);
            if (E.in_features > 2) c2 = __ldg(E.coords + size_t(row) * E.in_features + 2);
          }
          cin2[0] = f2_bcast(c0); cin2[1] = f2_bcast(c1); cin2[2] = f2_bcast(c2);
        }
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) facc2[o] = 0ull;
      }

      if constexpr (kBwd) {  // prefetch the first saved-z tile of this job while its MMAs are still running
        if (ch0 < nchunks && lane == 0) {
          const uint32_t bar = smem_u32(&bar_in[ew]);
          mbar_expect_tx(bar, P.n_in * kTile16Bytes);
          for (int s = 0; s < P.n_in; ++s)
            tma_load_2d(inbuf + s * kTile16Bytes, &P.z_map[s], bar, col0 + ch0 * kChunk, row0 + q * 32);
        }
      }

      {
        const long long t0 = WIRE_CLK();
        mbar_wait(smem_u32(&bar_tmem_full[buf]), (jb >> 1) & 1);
        e_wait += WIRE_CLK() - t0;
      }
      tc_fence_after();
      if (my_last < 0) release(buf);

      for (int ch = ch0; ch < nchunks; ch += kEpi16Parts) {
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * P.buf_cols + ch * kChunk;
        const int c = col0 + ch * kChunk;  // first real output column of this chunk (also the smem table index)
        uint32_t raw[32];
        f2 wn[8];          // 2D fwd: |w|^2 per feature pair
        uint32_t pkw[16];  // 2D fwd: the packed w half, staged together with y and z below
        if constexpr (MODE == MODE_GABOR2D_FWD) {
          // w half first: |w|^2 per feature is all the Gabor needs; w itself goes straight to its staging tile
          tmem_ld32(taddr + P.nbh, raw);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 b = s_bias24[(c >> 2) + 2 * g + h];
              const f2 wr = f2_add(f2_bits(raw[8 * g + 4 * h], raw[8 * g + 4 * h + 1]), f2_make(b.x, b.y));
              const f2 wi = f2_add(f2_bits(raw[8 * g + 4 * h + 2], raw[8 * g + 4 * h + 3]), f2_make(b.z, b.w));
              wn[2 * g + h] = f2_fma(wr, wr, f2_mul(wi, wi));
              pkw[4 * g + 2 * h] = pack_f16(f2_lo(wr), f2_lo(wi));
              pkw[4 * g + 2 * h + 1] = pack_f16(f2_hi(wr), f2_hi(wi));
            }
          }
        }
        tmem_ld32(taddr, raw);
        tmem_wait_ld();
        if (ch == my_last) release(buf);  // all TMEM reads of this warp for this job are done

        if constexpr (MODE == MODE_PLAIN) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a = __uint_as_float(raw[8 * g + 2 * i]), b = __uint_as_float(raw[8 * g + 2 * i + 1]);
              pk[i] = P.o_fmt[0] == 2 ? pack_bf16(a, b) : pack_f16(a, b);
            }
            sts128(sw64_addr(wbuf, lane, g), pk[0], pk[1], pk[2], pk[3]);
          }
        } else if constexpr (kFwd) {
          // all the math of the chunk first (results packed into the registers the accumulators leave behind), THEN the wait
          // for the previous chunk's TMA stores to have read the staging tiles: the stores drain behind ~500 cycles of math
          // instead of in front of it (ncu r01 v15: long_scoreboard 6.9 warps per issue cycle in the two-store forward layer)
          uint32_t py[16], pz[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // two feature pairs per group of 8 accumulator columns
              const float4 b = s_bias4[(c >> 2) + 2 * g + h];
              const f2 zr = f2_add(f2_bits(raw[8 * g + 4 * h], raw[8 * g + 4 * h + 1]), f2_make(b.x, b.y));
              const f2 zi = f2_add(f2_bits(raw[8 * g + 4 * h + 2], raw[8 * g + 4 * h + 3]), f2_make(b.z, b.w));
              f2 yr, yi;
              gabor_x2(G2, zr, zi, k2D ? wn[2 * g + h] : 0ull, yr, yi);
              if constexpr (FUSE == 3) {
                const float4* wq = s_wf + ((c >> 2) + 2 * g + h) * 3;
                const float4 w0 = wq[0], w1 = wq[1], w2 = wq[2];
                facc2[0] = f2_fma(yr, f2_make(w0.x, w0.y), f2_fma(yi, f2_make(w1.z, w1.w), facc2[0]));
                facc2[1] = f2_fma(yr, f2_make(w0.z, w0.w), f2_fma(yi, f2_make(w2.x, w2.y), facc2[1]));
                facc2[2] = f2_fma(yr, f2_make(w1.x, w1.y), f2_fma(yi, f2_make(w2.z, w2.w), facc2[2]));
              } else if constexpr (FUSE != 0) {
                const float4* wq = s_wf + ((c >> 2) + 2 * g + h) * 4;
                const float4 w0 = wq[0], w1 = wq[1], w2 = wq[2], w3 = wq[3];
                facc2[0] = f2_fma(yr, f2_make(w0.x, w0.y), f2_fma(yi, f2_make(w2.x, w2.y), facc2[0]));
                facc2[1] = f2_fma(yr, f2_make(w0.z, w0.w), f2_fma(yi, f2_make(w2.z, w2.w), facc2[1]));
                facc2[2] = f2_fma(yr, f2_make(w1.x, w1.y), f2_fma(yi, f2_make(w3.x, w3.y), facc2[2]));
                facc2[3] = f2_fma(yr, f2_make(w1.z, w1.w), f2_fma(yi, f2_make(w3.z, w3.w), facc2[3]));
              }
              py[4 * g + 2 * h] = pack_f16(f2_lo(yr), f2_lo(yi));
              py[4 * g + 2 * h + 1] = pack_f16(f2_hi(yr), f2_hi(yi));
              pz[4 * g + 2 * h] = pack_f16(f2_lo(zr), f2_lo(zi));
              pz[4 * g + 2 * h + 1] = pack_f16(f2_hi(zr), f2_hi(zi));
            }
          }
          if (st0 && c <= E.n_cols && c + kChunk > E.n_cols) {
            // the chunk that holds column 2M: the stored width is rounded up to a whole 32-byte sector (api.cu: run_rows_job), so this
            // store overwrites the "ones" column of the y tensor.  The zero-padded feature M gives y = gabor(0) = 1 + 0j there
            // anyway; written explicitly so that the wgrad's bias row does not hang on ex2(0) * cos(0) being exactly 1
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c + 2 * i == E.n_cols) py[i] = 0x00003C00u;   // FP16 (1.0, 0.0)
          }
          if (n_out > 0) { if (lane == 0) tma_store_wait_read<0>(); __syncwarp(); }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (st0) sts128(sw64_addr(wbuf, lane, g), py[4 * g], py[4 * g + 1], py[4 * g + 2], py[4 * g + 3]);
            if (st1) sts128(sw64_addr(wbuf1, lane, g), pz[4 * g], pz[4 * g + 1], pz[4 * g + 2], pz[4 * g + 3]);
            if constexpr (MODE == MODE_GABOR2D_FWD) {
              if (st2) sts128(sw64_addr(wbuf2, lane, g), pkw[4 * g], pkw[4 * g + 1], pkw[4 * g + 2], pkw[4 * g + 3]);
            }
          }
        } else if constexpr (kBwd) {
          // this chunk's saved z (w) tile: copy the packed halves to registers, then prefetch the next tile into the same buffer
          uint32_t zp[16], wp[16];
          {
            const long long t0 = WIRE_CLK();
            mbar_wait(smem_u32(&bar_in[ew]), in_phase);
            e_wait_in += WIRE_CLK() - t0;
          }
          in_phase ^= 1;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            lds128(sw64_addr(inbuf, lane, g), zp[4 * g], zp[4 * g + 1], zp[4 * g + 2], zp[4 * g + 3]);
            if constexpr (k2D) lds128(sw64_addr(inbuf + kTile16Bytes, lane, g), wp[4 * g], wp[4 * g + 1], wp[4 * g + 2], wp[4 * g + 3]);
          }
          __syncwarp();
          if (ch + kEpi16Parts < nchunks && lane == 0) {
            const uint32_t bar = smem_u32(&bar_in[ew]);
            mbar_expect_tx(bar, P.n_in * kTile16Bytes);
            for (int s = 0; s < P.n_in; ++s)
              tma_load_2d(inbuf + s * kTile16Bytes, &P.z_map[s], bar, c + kEpi16Parts * kChunk, row0 + q * 32);
          }
          uint32_t pz[16], pw[16];  // math first, store-read wait after it (see the forward branch)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float2 za = unpack_f16(zp[4 * g + 2 * h]), zb = unpack_f16(zp[4 * g + 2 * h + 1]);
              const f2 zr = f2_make(za.x, zb.x), zi = f2_make(za.y, zb.y);
              f2 wr = 0ull, wi = 0ull, wnorm = 0ull;
              if constexpr (k2D) {
                const float2 wa = unpack_f16(wp[4 * g + 2 * h]), wb = unpack_f16(wp[4 * g + 2 * h + 1]);
                wr = f2_make(wa.x, wb.x); wi = f2_make(wa.y, wb.y);
                wnorm = f2_fma(wr, wr, f2_mul(wi, wi));
              }
              const f2 gr = f2_bits(raw[8 * g + 4 * h], raw[8 * g + 4 * h + 1]);
              const f2 gi = f2_bits(raw[8 * g + 4 * h + 2], raw[8 * g + 4 * h + 3]);
              f2 yr, yi, gzr, gzi;
              gabor_x2(G2, zr, zi, wnorm, yr, yi);
              f2 pr;
              if constexpr (SCAL) {
                f2 pi;
                pr = gabor_bwd_x2_p(G2, yr, yi, zr, zi, gr, gi, gzr, gzi, pi);
                s_om = f2_fma(zr, pi, f2_fma(f2_mul(zi, G2.none), pr, s_om));
                s_sc = f2_fma(f2_fma(zi, zi, f2_fma(zr, zr, wnorm)), pr, s_sc);
              } else {
                pr = gabor_bwd_x2(G2, yr, yi, zr, zi, gr, gi, gzr, gzi);
              }
              pz[4 * g + 2 * h] = pack_bf16(f2_lo(gzr), f2_lo(gzi));
              pz[4 * g + 2 * h + 1] = pack_bf16(f2_hi(gzr), f2_hi(gzi));
              if constexpr (k2D) {
                const f2 t = f2_mul(G2.m2s2, pr);
                const f2 gwr = f2_mul(t, wr), gwi = f2_mul(t, wi);
                pw[4 * g + 2 * h] = pack_bf16(f2_lo(gwr), f2_lo(gwi));
                pw[4 * g + 2 * h + 1] = pack_bf16(f2_hi(gwr), f2_hi(gwi));
              }
            }
          }
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (st0) sts128(sw64_addr(wbuf, lane, g), pz[4 * g], pz[4 * g + 1], pz[4 * g + 2], pz[4 * g + 3]);
            if constexpr (k2D) { if (st1) sts128(sw64_addr(wbuf1, lane, g), pw[4 * g], pw[4 * g + 1], pw[4 * g + 2], pw[4 * g + 3]); }
          }
        } else {  // kFirst: real z0 recomputed from the coordinates and the smem weight table; BF16 direct stores
          uint32_t gzp[8], gwp[8];  // 16 real outputs as packed BF16 pairs (features A, B of each pair are adjacent)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 ta = s_tab[((c >> 2) + j) * 2], tb = s_tab[((c >> 2) + j) * 2 + 1];
            const f2 z0 = f2_fma(cin2[0], f2_make(ta.x, ta.y), f2_fma(cin2[1], f2_make(ta.z, ta.w), f2_fma(cin2[2], f2_make(tb.x, tb.y), f2_make(tb.z, tb.w))));
            f2 w0v = 0ull, wnorm = 0ull;
            if constexpr (k2D) {
              const float4 ua = s_tab2[((c >> 2) + j) * 2], ub = s_tab2[((c >> 2) + j) * 2 + 1];
              w0v = f2_fma(cin2[0], f2_make(ua.x, ua.y), f2_fma(cin2[1], f2_make(ua.z, ua.w), f2_fma(cin2[2], f2_make(ub.x, ub.y), f2_make(ub.z, ub.w))));
              wnorm = f2_mul(w0v, w0v);
            }
            f2 yr, yi, gz;
            gabor_real_x2(G2, z0, wnorm, yr, yi);
            f2 pr;
            if constexpr (SCAL) {
              f2 pi;
              pr = gabor_first_bwd_x2_p(G2, yr, yi, z0, f2_bits(raw[4 * j], raw[4 * j + 1]), f2_bits(raw[4 * j + 2], raw[4 * j + 3]), gz, pi);
              s_om = f2_fma(z0, pi, s_om);
              s_sc = f2_fma(f2_fma(z0, z0, wnorm), pr, s_sc);
            } else {
              pr = gabor_first_bwd_x2(G2, yr, yi, z0, f2_bits(raw[4 * j], raw[4 * j + 1]), f2_bits(raw[4 * j + 2], raw[4 * j + 3]), gz);
            }
            gzp[j] = pack_bf16(f2_lo(gz), f2_hi(gz));
            if constexpr (k2D) {
              const f2 gw = f2_mul(f2_mul(G2.m2s2, pr), w0v);
              gwp[j] = pack_bf16(f2_lo(gw), f2_hi(gw));
            }
          }
          if (row_ok) {
            // 16 real outputs per thread as BF16 = 32 contiguous bytes (one full sector): two direct 16-byte stores
            // (gz0_pitch is a multiple of 8 elements; E.gz0 / E.gw0 point at BF16 tensors on this path)
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(E.gz0) + size_t(row) * E.gz0_pitch + (c >> 1);
#pragma unroll
            for (int j8 = 0; j8 < 2; ++j8)
              if ((c >> 1) + 8 * j8 < E.gz0_pitch)
                *reinterpret_cast<uint4*>(dst + 8 * j8) = make_uint4(gzp[4 * j8], gzp[4 * j8 + 1], gzp[4 * j8 + 2], gzp[4 * j8 + 3]);
            if constexpr (k2D) {
              __nv_bfloat16* dw = reinterpret_cast<__nv_bfloat16*>(E.gw0) + size_t(row) * E.gz0_pitch + (c >> 1);
#pragma unroll
              for (int j8 = 0; j8 < 2; ++j8)
                if ((c >> 1) + 8 * j8 < E.gz0_pitch)
                  *reinterpret_cast<uint4*>(dw + 8 * j8) = make_uint4(gwp[4 * j8], gwp[4 * j8 + 1], gwp[4 * j8 + 2], gwp[4 * j8 + 3]);
            }
          }
        }

        if (n_out > 0) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            for (int s = 0; s < n_out; ++s) tma_store_2d(&P.o_map[s], wbuf + s * kTile16Bytes, c, row0 + q * 32);
            tma_store_commit();
          }
        }
      }

      if constexpr (FUSE) {
        if (sl == P.slices - 1) {
          // the four warps of a sub-partition hold partial sums over their chunks (of every slice)
          const int slot = it & 1;
          const float facc[4] = {f2_lo(facc2[0]) + f2_hi(facc2[0]), f2_lo(facc2[1]) + f2_hi(facc2[1]),
                                 f2_lo(facc2[2]) + f2_hi(facc2[2]), f2_lo(facc2[3]) + f2_hi(facc2[3])};
          if (part > 0) s_fin[(slot * (kEpi16Parts - 1) + (part - 1)) * kTileRows + q * 32 + lane] = make_float4(facc[0], facc[1], facc[2], facc[3]);
          asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * kEpi16Parts) : "memory");
          if (part == 0 && row_ok) {
            float r[4] = {facc[0], facc[1], facc[2], facc[3]};
#pragma unroll
            for (int pp = 0; pp < kEpi16Parts - 1; ++pp) {
              const float4 o = s_fin[(slot * (kEpi16Parts - 1) + pp) * kTileRows + q * 32 + lane];
              r[0] += o.x; r[1] += o.y; r[2] += o.z; r[3] += o.w;
            }
#pragma unroll
            for (int oo = 0; oo < kMaxOut; ++oo)
              if (oo < E.out_features) E.out[size_t(row) * E.out_features + oo] = r[oo] + __ldg(E.bf + 2 * oo);
          }
        }
      }
    }
    if constexpr (SCAL) {
      // rows past the end and padded feature columns contribute exact zeros (their accumulators / z tiles are TMA zero fill)
      float a = f2_lo(s_om) + f2_hi(s_om), b = f2_lo(s_sc) + f2_hi(s_sc);
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, sft); b += __shfl_xor_sync(0xffffffffu, b, sft); }
      if (lane == 0) {
        if (E.g_omega) atomicAdd(E.g_omega, a);
        if (E.g_scale) atomicAdd(E.g_scale, -2.0f * __ldg(E.scale) * b);
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
    if (dbg && ew == 0 && lane == 0) { dbg[kDbgEpiWaitAcc] = e_wait; dbg[kDbgEpiTotal] = WIRE_CLK() - e_begin; dbg[kDbgEpiWaitIn] = e_wait_in; }
  }

  // ===================== teardown =====================
  if (dbg && threadIdx.x == 0) dbg[9] = (unsigned long long)(WIRE_CLK() - t_entry);   // producer done
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[10] = (unsigned long long)(WIRE_CLK() - t_entry);  // every warp of this CTA done
  if (pair) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (pair) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
  if (dbg && threadIdx.x == 0) dbg[11] = (unsigned long long)(WIRE_CLK() - t_entry);  // exit
}

}  // namespace wire
