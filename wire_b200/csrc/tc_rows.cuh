// tc_rows.cuh — the row-tile GEMM of the WIRE hot path on tcgen05 (sm_100a).
//
//   ACC[128 coords, nb cols] = A[128 coords, K] (K-major, TMA) x Bpacked[nb cols, K] (K-major, TMA)
//
// with TF32 inputs, FP32 accumulation in TMEM, and the layer's point-wise work fused in the
// epilogue straight out of TMEM; results leave through swizzled staging + TMA stores.
//
//   MODE_PLAIN        out0 = ACC                                         (probe / per-layer dgrad)
//   MODE_GABOR_FWD    z = ACC + b ; y = gabor(z)          -> y, z        modules/wire.py:88-93
//   MODE_GABOR2D_FWD  z,w = ACC halves + b1,b2 ; y = gabor2d(z,w) -> y,z,w   modules/wire2d.py:56-67
//   MODE_GABOR_BWD    g_y = ACC ; g_z = gabor'(z_saved, g_y)  -> g_z     (autograd of wire.py:88-93)
//   MODE_GABOR2D_BWD  same + g_w                               -> g_z, g_w
//   MODE_FIRST_BWD / MODE_FIRST2D_BWD   g_y0 = ACC ; z0 recomputed from coords -> real g_z0 (g_w0)
//
// A is the complex activation tensor seen as real [N, 2M] (interleaved re,im = torch complex64),
// Bpacked is the real 2x2-block expansion of the complex weight (pack_weights_kernel), so one real
// GEMM of width 2M x 2M *is* the complex GEMM (8*M^2 flop/coord, no de-interleave anywhere).
//
// Warp roles (320 threads): warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc),
// warps 2..9 = epilogue: TMEM sub-partition = warp & 3, the two warps of a sub-partition take the
// even / odd 32-column chunks.  Layer constants (bias, final-Linear weights, first-layer weights)
// live in shared memory; the saved pre-activations of the backward modes are TMA-prefetched.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "rows_epilogue.cuh"
#include "sm100.cuh"

namespace wire {

constexpr int kEpiWarps = 8;
constexpr int kRowsThreads = 64 + 32 * kEpiWarps;
constexpr int kTileRows = 128;
constexpr int kChunk = 32;  // fp32 columns per 128-byte swizzle row

struct RowsParams {
  CUtensorMap a_map[2];  // A parts, box {32 cols, 128 rows}
  CUtensorMap b_map;     // packed B [n_blocks*nb, Kpad], box {32 cols, b_box_rows}
  CUtensorMap o_map[3];  // outputs, box {32 cols, 32 rows}
  CUtensorMap z_map[2];  // saved z (and w) for the backward epilogues, box {32 cols, 32 rows}
  int k_cols[2];         // valid K columns in each A part (part 1 may be 0)
  int n_blocks;          // column blocks (work item = row tile x block)
  int nb;                // accumulator columns per block (multiple of 16, <= 512)
  int nbh;               // 2D fwd: columns of the z half (w half follows); otherwise == nb
  int slices;               // N-slices per tile (1 or 2): each slice has its own K loop and TMEM accumulator buffer, so
                            // the MMAs of one slice overlap the epilogue of the previous one (TMEM is double-buffered)
  int ns;                   // accumulator columns per slice (<= 256; slices * ns == nb)
  int buf_cols;             // TMEM columns between the two accumulator buffers
  int b_box_rows;           // rows of the B tile this CTA loads per stage (= ns / cluster)
  int cluster;              // 1 = one CTA per tile (cta_group::1); 2 = CTA pair (cta_group::2): a 256-row tile, each
                            // CTA stages its own 128 rows of A and HALF of the B tile, halving the smem fill per flop
  int stages;
  int bstat;             // tc_rows16, two N-slices, one column block: B-STATIONARY schedule — every cluster owns ONE slice for the
                         // whole launch, loads that slice's packed weights (all K stages) into shared memory once and then
                         // streams only A tiles; the other slice of the same rows runs on the neighbouring cluster (its A reads
                         // hit L2).  Per row tile the SM's TMA unit moves 112 KB of operands instead of 420 KB.
  uint32_t bres_off;     // byte offset of the resident B region inside dynamic smem
  int store_mask;  // which epilogue results are TMA-stored: bit0 = o0, bit1 = o1, bit2 = o2
  int o_fmt[3];    // element type of each STORED output slot (sm100_host::ElemType): f32 tiles are 32x32x4 B with the
                   // 128 B swizzle, 16-bit tiles (FP16 saved z / y, BF16 gradients) are dense 32x32x2 B
  int l2_prefetch;   // tc_rows16: the producer prefetches the next work unit's A rows into L2
  int reverse;       // tc_rows16: walk the work units from the last row tile to the first (see api.cu: sweep directions)
  int a_fmt, b_fmt;  // OP16 kernels: operand formats of the kind::f16 MMA (sm100::kFmtF16 / kFmtBF16)
  int n_in;        // TMA-prefetched epilogue inputs (0, 1 = z, 2 = z and w)
  uint32_t staging_off;  // byte offsets inside dynamic smem (from the 1 KB aligned base)
  uint32_t param_off;
  int param_cols;        // padded column count of the bias tables (multiple of 32)
  // GEN kernels: the A operand is not loaded but COMPUTED by four generator warps straight into the swizzled
  // smem tile: A = y0 = gabor(coords W0^T + b0), the first WIRE layer (K = 2..3), so y0 never exists in HBM.
  // (coords / w0 / b0 / w0b / b0b / in_features of `e` describe that layer; these are its omega_0, scale_0.)
  const float* gen_omega;
  const float* gen_scale;
  int gen_two_d;
  uint32_t gen_tab_off;     // float offset of the generator's {w0[0..2], b0} table inside the param area
  unsigned long long* dbg;  // optional per-CTA stall counters [8] (tools/umma_probe): see kDbg* below
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
__device__ __forceinline__ void unstage_row(uint32_t buf, int lane, float (&v)[32]) {
  const uint32_t row = buf + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t addr = row + ((j ^ sw) << 4);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[4 * j]), "=f"(v[4 * j + 1]), "=f"(v[4 * j + 2]), "=f"(v[4 * j + 3])
                 : "r"(addr)
                 : "memory");
  }
}

// FP16 tiles (saved z / w): 32 rows x 32 halves = 64 B per row, dense (TMA SWIZZLE_NONE)
__device__ __forceinline__ void stage_row_half(uint32_t buf, int lane, const float (&v)[32]) {
  const uint32_t row = buf + lane * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2 t = __floats2half2_rn(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
      h[k] = *reinterpret_cast<const uint32_t*>(&t);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + j * 16), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  }
}
__device__ __forceinline__ void stage_row_bf16(uint32_t buf, int lane, const float (&v)[32]) {
  const uint32_t row = buf + lane * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
      h[k] = *reinterpret_cast<const uint32_t*>(&t);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + j * 16), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  }
}
// stage one output chunk in the element type of its tensor (fmt: sm100_host::ElemType)
__device__ __forceinline__ void stage_row_fmt(uint32_t buf, int lane, const float (&v)[32], int fmt) {
  if (fmt == 0) stage_row(buf, lane, v);
  else if (fmt == 1) stage_row_half(buf, lane, v);
  else stage_row_bf16(buf, lane, v);
}
__device__ __forceinline__ void unstage_row_half(uint32_t buf, int lane, float (&v)[32]) {
  const uint32_t row = buf + lane * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t h[4];
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]) : "r"(row + j * 16) : "memory");
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[k]));
      v[8 * j + 2 * k] = f.x;
      v[8 * j + 2 * k + 1] = f.y;
    }
  }
}

// waits with back-off for the single-thread producer / MMA roles (they share schedulers with epilogue warps)
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  using namespace sm100;
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);
    if ((++spins & 0xff) == 0 && (clock64() - t0) > 4000000000LL) __trap();
  }
}

// stall counters (clock64 cycles), one row of 8 per CTA; compiled in only with -DWIRE_B200_STALL_COUNTERS
#ifdef WIRE_B200_STALL_COUNTERS
#define WIRE_CLK() clock64()
#else
#define WIRE_CLK() 0ll
#endif
enum { kDbgMmaWaitFull = 0, kDbgMmaWaitTmem = 1, kDbgMmaTotal = 2, kDbgEpiWaitAcc = 3, kDbgEpiTotal = 4, kDbgProdWaitEmpty = 5,
       kDbgProdTotal = 6, kDbgEpiWaitIn = 7 };

constexpr int kGenWarps = 4;

// OP16: A and B are 16-bit (FP16 activations / BF16 gradients, formats in P.a_fmt / P.b_fmt), MMA kind::f16.
// The smem tiles keep their byte geometry (128 B swizzled rows, 32 B per K-step), so a stage covers 64 K columns.
template <int MODE, bool PAIR, bool GEN = false, bool OP16 = false>
__global__ void __launch_bounds__(kRowsThreads + (GEN ? 32 * kGenWarps : 0), 1) tc_rows_kernel(const __grid_constant__ RowsParams P) {
  using namespace sm100;
  static_assert(!(GEN && OP16), "the in-place generator writes TF32 tiles");
  constexpr int kKC = OP16 ? 64 : 32;    // K columns per pipeline stage (one 128 B swizzle row)
  constexpr int kKStep = OP16 ? 16 : 8;  // K columns per tcgen05.mma
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8];
  __shared__ __align__(8) uint64_t bar_empty[8];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ __align__(8) uint64_t bar_in[kEpiWarps];
  __shared__ __align__(16) float fin_xchg[2][kTileRows][kMaxOut];
  __shared__ uint32_t tmem_slot;

  constexpr bool kFwd = (MODE == MODE_GABOR_FWD || MODE == MODE_GABOR2D_FWD);
  constexpr bool kBwd = (MODE == MODE_GABOR_BWD || MODE == MODE_GABOR2D_BWD);
  constexpr bool kFirst = (MODE == MODE_FIRST_BWD || MODE == MODE_FIRST2D_BWD);
  constexpr bool k2D = (MODE == MODE_GABOR2D_FWD || MODE == MODE_GABOR2D_BWD || MODE == MODE_FIRST2D_BWD);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_bytes = kTileRows * 128;
  const uint32_t b_bytes = uint32_t(P.b_box_rows) * 128;  // this CTA's share of one B slice chunk
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t staging_base = smem_base + P.staging_off;
  float* params = reinterpret_cast<float*>(smem_gen + P.param_off);
  const RowsEpi& E = P.e;

  const int kc0 = (P.k_cols[0] + kKC - 1) / kKC;
  const int kc1 = (P.k_cols[1] + kKC - 1) / kKC;
  const int kc_total = kc0 + kc1;
  const int row_tiles = (E.n_rows + kTileRows - 1) / kTileRows;
  const int n_feat = E.n_cols >> 1;  // complex features M
  // Work decomposition.  The two CTAs of a pair work on the SAME column block and on two consecutive
  // row tiles, and every CTA of the grid runs the same number of iterations (tiles past the end are
  // computed on zero-filled rows and their stores are clipped by TMA).  A tile is processed as
  // `slices` jobs; job jb uses TMEM accumulator buffer jb & 1.
  const int C = P.cluster;
  const int crank = int(cluster_ctarank());
  const int n_clusters = gridDim.x / C;
  const int my_cluster = blockIdx.x / C;
  const int row_groups = (row_tiles + C - 1) / C;
  const int n_units = row_groups * P.n_blocks;                 // cluster-level work units
  const int n_iters = (n_units + n_clusters - 1) / n_clusters;
  const int n_jobs = n_iters * P.slices;
  constexpr bool pair = PAIR;
  const bool leader = crank == 0;
  unsigned long long* dbg = P.dbg ? P.dbg + size_t(blockIdx.x) * 8 : nullptr;

  // ---- shared parameter tables (zero padded so the epilogue needs no column checks) ----
  //  fwd : bias[param_cols] | bias2[param_cols] (2D) | wf[(param_cols/2)][8]  (wr[4], wi[4]) if fused
  //  first: tab[(param_cols/2)][4] = {w0[0..2], b0}  | tab2 (2D)              (in_features <= 3)
  if constexpr (kFwd) {
    for (int i = threadIdx.x; i < P.param_cols; i += blockDim.x) {
      params[i] = i < E.n_cols ? E.bias[i] : 0.f;
      if constexpr (k2D) params[P.param_cols + i] = i < E.n_cols ? E.bias2[i] : 0.f;
    }
    if (E.fuse_final) {
      float* wf = params + (k2D ? 2 : 1) * P.param_cols;
      for (int i = threadIdx.x; i < (P.param_cols >> 1) * 8; i += blockDim.x) {
        const int k = i >> 3, j = i & 7, o = j & 3;
        float v = 0.f;
        if (k < n_feat && o < E.out_features) v = E.wf[(size_t(o) * n_feat + k) * 2 + (j >> 2)];
        wf[i] = v;
      }
    }
  }
  if constexpr (kFirst) {
    for (int i = threadIdx.x; i < (P.param_cols >> 1) * 4; i += blockDim.x) {
      const int k = i >> 2, j = i & 3;
      float v = 0.f, v2 = 0.f;
      if (k < n_feat) {
        if (j < 3) {
          if (j < E.in_features) { v = E.w0[size_t(k) * E.in_features + j]; if constexpr (k2D) v2 = E.w0b[size_t(k) * E.in_features + j]; }
        } else { v = E.b0[k]; if constexpr (k2D) v2 = E.b0b[k]; }
      }
      params[i] = v;
      if constexpr (k2D) params[(P.param_cols >> 1) * 4 + i] = v2;
    }
  }

  if constexpr (GEN) {
    float* gt = params + P.gen_tab_off;
    for (int i = threadIdx.x; i < (P.param_cols >> 1) * 4; i += blockDim.x) {
      const int k = i >> 2, j = i & 3;
      const int n_in_feat = P.k_cols[0] >> 1;  // complex features of the generated operand
      float v = 0.f, v2 = 0.f;
      if (k < n_in_feat) {
        if (j < 3) {
          if (j < E.in_features) { v = E.w0[size_t(k) * E.in_features + j]; if (P.gen_two_d) v2 = E.w0b[size_t(k) * E.in_features + j]; }
        } else { v = E.b0[k]; if (P.gen_two_d) v2 = E.b0b[k]; }
      }
      gt[i] = v;
      gt[(P.param_cols >> 1) * 4 + i] = v2;
    }
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      // GEN: one arrival from the TMA producer (B tile bytes) + one per generator warp of every CTA of the pair
      mbar_init(smem_u32(&bar_full[s]), GEN ? 1 + kGenWarps * C : 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tmem_full[b]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[b]), kEpiWarps * C);  // pair: both CTAs' epilogues release the leader
    }
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(smem_u32(&bar_in[w]), 1);
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
  if (pair) cluster_sync_all();  // the peer's barriers must exist before any remote arrive / complete_tx
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long t_wait = 0;
      const long long t_begin = WIRE_CLK();
      for (int jb = 0; jb < n_jobs; ++jb) {
        const int it = jb / P.slices, sl = jb % P.slices;
        const int unit = it * n_clusters + my_cluster;
        const int row0 = ((unit / P.n_blocks) * C + crank) * kTileRows;
        const int brow = (unit % P.n_blocks) * P.nb + sl * P.ns + crank * P.b_box_rows;  // this CTA's B rows
        // the A tile is read once per slice: keep it in L2 until its last use, then let it go first
        const uint64_t a_policy = (sl == P.slices - 1 && (unit % P.n_blocks) == P.n_blocks - 1) ? kEvictFirst : kEvictLast;
        for (int kc = 0; kc < kc_total; ++kc) {
          const long long t0 = WIRE_CLK();
          mbar_wait_backoff(smem_u32(&bar_empty[stage]), phase ^ 1);
          t_wait += WIRE_CLK() - t0;
          const uint32_t full_own = smem_u32(&bar_full[stage]);
          const uint32_t a_dst = smem_base + stage * stage_bytes;
          const int part = kc < kc0 ? 0 : 1;
          const int kcol = (part ? kc - kc0 : kc) * kKC;
          const uint32_t tx_bytes = GEN ? b_bytes : stage_bytes;  // GEN: the A tile is written by the generator warps
          if (!pair) {
            mbar_expect_tx(full_own, tx_bytes);
            if (!GEN) tma_load_2d_hint(a_dst, &P.a_map[part], full_own, kcol, row0, a_policy);
            tma_load_2d_hint(a_dst + a_bytes, &P.b_map, full_own, kc * kKC, brow, kEvictLast);
          } else {
            // both CTAs load into their own smem; all bytes complete on the LEADER's full barrier
            const uint32_t full_leader = full_own & kPeerBitMask;
            if (leader) mbar_expect_tx(full_own, 2 * tx_bytes);
            if (!GEN) tma_load_2d_2cta(a_dst, &P.a_map[part], full_leader, kcol, row0, a_policy);
            tma_load_2d_2cta(a_dst + a_bytes, &P.b_map, full_leader, kc * kKC, brow, kEvictLast);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg) { dbg[kDbgProdWaitEmpty] = t_wait; dbg[kDbgProdTotal] = WIRE_CLK() - t_begin; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && (!pair || leader)) {
      const uint32_t idesc = OP16 ? make_idesc_f16(pair ? 256 : 128, P.ns, false, false, uint32_t(P.a_fmt), uint32_t(P.b_fmt))
                                  : make_idesc_tf32(pair ? 256 : 128, P.ns, false, false);
      // descriptor words are precomputed: only the 14-bit start-address field changes (stage, K-step)
      const uint32_t desc_hi = uint32_t(make_sdesc_sw128(0, 16, 1024) >> 32);
      const uint32_t a_lo0 = uint32_t(make_sdesc_sw128(smem_base, 16, 1024));
      const uint32_t stage_units = stage_bytes >> 4, b_units = a_bytes >> 4;
      auto tail_steps = [](int cols, int kc) {  // K-steps of the last (possibly partial) stage of an A part
        const int st = (cols - (kc - 1) * kKC + kKStep - 1) / kKStep;
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
          mbar_wait_backoff(smem_u32(&bar_tmem_empty[buf]), ((jb >> 1) & 1) ^ 1);
          w_tmem += WIRE_CLK() - t0;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * P.buf_cols;
        for (int kc = 0; kc < kc_total; ++kc) {
          const long long t1 = WIRE_CLK();
          mbar_wait_backoff(smem_u32(&bar_full[stage]), phase);
          w_full += WIRE_CLK() - t1;
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * stage_units;
          const uint32_t b_lo = a_lo + b_units;
          const int steps = (kc == kc0 - 1) ? last0 : ((kc == kc_total - 1) ? last1 : 4);
          auto mma = [&](int ks, uint32_t acc) {
            const uint64_t adesc = (uint64_t(desc_hi) << 32) | (a_lo + 2 * ks);
            const uint64_t bdesc = (uint64_t(desc_hi) << 32) | (b_lo + 2 * ks);
            if constexpr (OP16) {
              if (pair) umma_f16_2cta(d_tmem, adesc, bdesc, idesc, acc);
              else umma_f16(d_tmem, adesc, bdesc, idesc, acc);
            } else {
              if (pair) umma_tf32_2cta(d_tmem, adesc, bdesc, idesc, acc);
              else umma_tf32(d_tmem, adesc, bdesc, idesc, acc);
            }
          };
          if (steps == 4) {
            mma(0, kc ? 1u : 0u);
            mma(1, 1u);
            mma(2, 1u);
            mma(3, 1u);
          } else {
            for (int ks = 0; ks < steps; ++ks) mma(ks, (kc | ks) ? 1u : 0u);
          }
          // release the smem stage (in both CTAs of a pair) once these MMAs have completed
          if (pair) umma_commit_2cta_mcast(smem_u32(&bar_empty[stage]), 3);
          else umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        if (pair) umma_commit_2cta_mcast(smem_u32(&bar_tmem_full[buf]), 3);
        else umma_commit(smem_u32(&bar_tmem_full[buf]));
      }
      if (dbg) { dbg[kDbgMmaWaitFull] = w_full; dbg[kDbgMmaWaitTmem] = w_tmem; dbg[kDbgMmaTotal] = WIRE_CLK() - t_begin; }
    }
  } else if (GEN && warp >= 2 + kEpiWarps) {
    // ===================== A-operand generator warps (first WIRE layer computed in place) =====================
    const int gw = warp - (2 + kEpiWarps);  // rows 32*gw .. 32*gw+31 of the tile
    const GaborConst G0 = make_gabor_const(__ldg(P.gen_omega), __ldg(P.gen_scale));
    const float4* tab = reinterpret_cast<const float4*>(params + P.gen_tab_off);
    const float4* tab2 = tab + (P.param_cols >> 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int jb = 0; jb < n_jobs; ++jb) {
      const int it = jb / P.slices;
      const int unit = it * n_clusters + my_cluster;
      const int row = ((unit / P.n_blocks) * C + crank) * kTileRows + gw * 32 + lane;
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (row < E.n_rows) {
        c0 = __ldg(E.coords + size_t(row) * E.in_features);
        if (E.in_features > 1) c1 = __ldg(E.coords + size_t(row) * E.in_features + 1);
        if (E.in_features > 2) c2 = __ldg(E.coords + size_t(row) * E.in_features + 2);
      }
      for (int kc = 0; kc < kc_total; ++kc) {
        if (lane == 0) mbar_wait_backoff(smem_u32(&bar_empty[stage]), phase ^ 1);
        __syncwarp();
        float v[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 t = tab[kc * 16 + i];
          const float z0 = fmaf(c0, t.x, fmaf(c1, t.y, fmaf(c2, t.z, t.w)));
          float wn = 0.f;
          if (P.gen_two_d) {
            const float4 t2 = tab2[kc * 16 + i];
            const float w0v = fmaf(c0, t2.x, fmaf(c1, t2.y, fmaf(c2, t2.z, t2.w)));
            wn = w0v * w0v;
          }
          float yr, yi;
          gabor_fast(G0, z0, 0.f, wn, yr, yi);
          v[2 * i] = round_tf32(yr);
          v[2 * i + 1] = round_tf32(yi);
        }
        stage_row(smem_base + stage * stage_bytes + gw * 4096, lane, v);
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          const uint32_t full_own = smem_u32(&bar_full[stage]);
          if (pair) mbar_arrive_cluster(full_own & kPeerBitMask); else mbar_arrive(full_own);
        }
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;       // 0..7
    const int q = warp & 3;        // TMEM sub-partition (lanes 32q..32q+31)
    const int half = ew >> 2;      // takes chunks ch = half, half+2, ...
    const int n_out = __popc(P.store_mask);
    const int slots = n_out + P.n_in;
    const uint32_t wbuf = staging_base + ew * (slots * 4096);
    const uint32_t inbuf = wbuf + n_out * 4096;  // n_in x 4 KB (a tile is copied to registers at once, so one buffer
                                                  // is enough to keep the next chunk's TMA load in flight)
    const GaborConst G = (MODE != MODE_PLAIN) ? make_gabor_const(__ldg(E.omega), __ldg(E.scale)) : make_gabor_const(0.f, 0.f);
    const float* s_bias = params;
    const float* s_bias2 = params + P.param_cols;
    const float4* s_wf = reinterpret_cast<const float4*>(params + (k2D ? 2 : 1) * P.param_cols);
    const float4* s_tab = reinterpret_cast<const float4*>(params);
    const float4* s_tab2 = reinterpret_cast<const float4*>(params + (P.param_cols >> 1) * 4);
    const uint32_t empty_addr0 = pair ? (smem_u32(&bar_tmem_empty[0]) & kPeerBitMask) : smem_u32(&bar_tmem_empty[0]);
    const uint32_t empty_addr1 = pair ? (smem_u32(&bar_tmem_empty[1]) & kPeerBitMask) : smem_u32(&bar_tmem_empty[1]);
    const int out_blk = (MODE == MODE_GABOR2D_FWD) ? P.nbh : P.nb;  // output columns per column block
    const int out_ns = (MODE == MODE_GABOR2D_FWD) ? P.nbh : P.ns;   // output columns per slice
    uint32_t in_phase = 0;
    long long e_wait = 0, e_wait_in = 0;
    const long long e_begin = WIRE_CLK();
    float cin[3] = {0.f, 0.f, 0.f};
    float facc[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
    for (int jb = 0; jb < n_jobs; ++jb) {
      const int it = jb / P.slices, sl = jb % P.slices;
      const int buf = jb & 1;
      const int unit = it * n_clusters + my_cluster;
      const int row0 = ((unit / P.n_blocks) * C + crank) * kTileRows;
      const int blk = unit % P.n_blocks;
      const int row = row0 + q * 32 + lane;
      const bool row_ok = row < E.n_rows;
      const int col0 = blk * out_blk + sl * out_ns;  // first output column of this slice
      int valid = E.n_cols - col0;
      valid = valid > out_ns ? out_ns : valid;
      const int nchunks = valid > 0 ? (valid + kChunk - 1) / kChunk : 0;
      // chunks of this warp: half, half+2, ... ; last one it owns:
      int my_last = -1;
      if (nchunks > half) my_last = half + 2 * ((nchunks - 1 - half) >> 1);

      if (sl == 0) {
        if constexpr (kFirst) {
          cin[0] = cin[1] = cin[2] = 0.f;
          if (row_ok) {
            cin[0] = __ldg(E.coords + size_t(row) * E.in_features);
            if (E.in_features > 1) cin[1] = __ldg(E.coords + size_t(row) * E.in_features + 1);
            if (E.in_features > 2) cin[2] = __ldg(E.coords + size_t(row) * E.in_features + 2);
          }
        }
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) facc[o] = 0.f;
      }

      // prefetch the first saved-activation chunk of this job while its MMAs are still running
      if constexpr (kBwd) {
        if (half < nchunks && lane == 0) {
          const uint32_t bar = smem_u32(&bar_in[ew]);
          mbar_expect_tx(bar, P.n_in * (E.z_half ? 2048 : 4096));
          for (int s = 0; s < P.n_in; ++s)
            tma_load_2d(inbuf + s * 4096, &P.z_map[s], bar, col0 + half * kChunk, row0 + q * 32);
        }
      }

      {
        const long long t0 = WIRE_CLK();
        mbar_wait(smem_u32(&bar_tmem_full[buf]), (jb >> 1) & 1);
        e_wait += WIRE_CLK() - t0;
      }
      tc_fence_after();
      if (my_last < 0) {  // nothing to read in this job: just release the accumulator buffer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (pair) mbar_arrive_cluster(buf ? empty_addr1 : empty_addr0); else mbar_arrive(buf ? empty_addr1 : empty_addr0); }
      }

      for (int ch = half; ch < nchunks; ch += 2) {
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * P.buf_cols + ch * kChunk;
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
        if (ch == my_last) {
          // all TMEM reads of this warp for this job are done: hand the accumulator buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (pair) mbar_arrive_cluster(buf ? empty_addr1 : empty_addr0); else mbar_arrive(buf ? empty_addr1 : empty_addr0); }
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        const int c = col0 + ch * kChunk;  // first real output column of this chunk (also the smem table index)

        float o0[32], o1[32], o2[32];
        if constexpr (MODE == MODE_PLAIN) {
#pragma unroll
          for (int i = 0; i < 32; ++i) o0[i] = v[i];
        } else if constexpr (kFwd) {
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c);
          const float4* b24 = reinterpret_cast<const float4*>(s_bias2 + c);
#pragma unroll
          for (int i2 = 0; i2 < 8; ++i2) {
            const float4 bb = b4[i2];
            float4 bb2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (k2D) bb2 = b24[i2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int i = 2 * i2 + h;
              const float zr = v[2 * i] + (h ? bb.z : bb.x);
              const float zi = v[2 * i + 1] + (h ? bb.w : bb.y);
              float wn = 0.f;
              if constexpr (k2D) {
                const float wr = v2[2 * i] + (h ? bb2.z : bb2.x);
                const float wi = v2[2 * i + 1] + (h ? bb2.w : bb2.y);
                wn = fmaf(wr, wr, wi * wi);
                o2[2 * i] = wr;
                o2[2 * i + 1] = wi;
              }
              float yr, yi;
              gabor_fast(G, zr, zi, wn, yr, yi);
              if (E.fuse_final) {
                const float4 wr4 = s_wf[((c >> 1) + i) * 2];
                const float4 wi4 = s_wf[((c >> 1) + i) * 2 + 1];
                facc[0] = fmaf(yr, wr4.x, fmaf(-yi, wi4.x, facc[0]));
                facc[1] = fmaf(yr, wr4.y, fmaf(-yi, wi4.y, facc[1]));
                facc[2] = fmaf(yr, wr4.z, fmaf(-yi, wi4.z, facc[2]));
                facc[3] = fmaf(yr, wr4.w, fmaf(-yi, wi4.w, facc[3]));
              }
              if (E.round_out0) { yr = round_tf32(yr); yi = round_tf32(yi); }
              o0[2 * i] = yr;
              o0[2 * i + 1] = yi;
              o1[2 * i] = zr;
              o1[2 * i + 1] = zi;
            }
          }
        } else if constexpr (kBwd) {
          // wait for this chunk's z (w) tile, then immediately prefetch the next one into the same buffer
          float z[32], w[32];
          {
            const long long t0 = WIRE_CLK();
            mbar_wait(smem_u32(&bar_in[ew]), in_phase);
            e_wait_in += WIRE_CLK() - t0;
          }
          in_phase ^= 1;
          if (E.z_half) {
            unstage_row_half(inbuf, lane, z);
            if constexpr (k2D) unstage_row_half(inbuf + 4096, lane, w);
          } else {
            unstage_row(inbuf, lane, z);
            if constexpr (k2D) unstage_row(inbuf + 4096, lane, w);
          }
          __syncwarp();
          if (ch + 2 < nchunks && lane == 0) {
            const uint32_t bar = smem_u32(&bar_in[ew]);
            mbar_expect_tx(bar, P.n_in * (E.z_half ? 2048 : 4096));
            for (int s = 0; s < P.n_in; ++s)
              tma_load_2d(inbuf + s * 4096, &P.z_map[s], bar, c + 2 * kChunk, row0 + q * 32);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float zr = z[2 * i], zi = z[2 * i + 1];
            float wn = 0.f;
            if constexpr (k2D) wn = fmaf(w[2 * i], w[2 * i], w[2 * i + 1] * w[2 * i + 1]);
            float yr, yi, gzr, gzi;
            gabor_fast(G, zr, zi, wn, yr, yi);
            const float pr = gabor_bwd(yr, yi, zr, zi, v[2 * i], v[2 * i + 1], G.omega, G.s2, gzr, gzi);
            if (E.round_out0) { gzr = round_tf32(gzr); gzi = round_tf32(gzi); }
            o0[2 * i] = gzr;
            o0[2 * i + 1] = gzi;
            if constexpr (k2D) {
              const float t = -2.0f * G.s2 * pr;
              float gwr = t * w[2 * i], gwi = t * w[2 * i + 1];
              if (E.round_out0) { gwr = round_tf32(gwr); gwi = round_tf32(gwi); }
              o1[2 * i] = gwr;
              o1[2 * i + 1] = gwi;
            }
          }
        } else {  // kFirst: real z0 recomputed from the coordinates and the smem weight table
          float gz[16], gw[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 t = s_tab[(c >> 1) + i];
            const float z0 = fmaf(cin[0], t.x, fmaf(cin[1], t.y, fmaf(cin[2], t.z, t.w)));
            float w0v = 0.f;
            if constexpr (k2D) {
              const float4 t2 = s_tab2[(c >> 1) + i];
              w0v = fmaf(cin[0], t2.x, fmaf(cin[1], t2.y, fmaf(cin[2], t2.z, t2.w)));
            }
            float yr, yi;
            gabor_fast(G, z0, 0.f, w0v * w0v, yr, yi);
            const float pr = gabor_first_bwd(yr, yi, z0, v[2 * i], v[2 * i + 1], G.omega, G.s2, gz[i]);
            gw[i] = -2.0f * G.s2 * pr * w0v;
          }
          if (row_ok) {
            // 16 real outputs per thread = 64 contiguous bytes (two full sectors): direct stores
            float4* dst = reinterpret_cast<float4*>(E.gz0 + size_t(row) * E.gz0_pitch + (c >> 1));
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              if ((c >> 1) + 4 * j4 < E.gz0_pitch) dst[j4] = make_float4(gz[4 * j4], gz[4 * j4 + 1], gz[4 * j4 + 2], gz[4 * j4 + 3]);
            if constexpr (k2D) {
              float4* dw = reinterpret_cast<float4*>(E.gw0 + size_t(row) * E.gz0_pitch + (c >> 1));
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4)
                if ((c >> 1) + 4 * j4 < E.gz0_pitch) dw[j4] = make_float4(gw[4 * j4], gw[4 * j4 + 1], gw[4 * j4 + 2], gw[4 * j4 + 3]);
            }
          }
        }

        if (n_out > 0) {
          // staging -> TMA store; the previous store of this warp must have finished READING smem
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          int slot = 0;
          if (P.store_mask & 1) { stage_row_fmt(wbuf, lane, o0, P.o_fmt[0]); ++slot; }
          if constexpr (kFwd || MODE == MODE_GABOR2D_BWD) {
            if (P.store_mask & 2) { stage_row_fmt(wbuf + slot * 4096, lane, o1, P.o_fmt[slot]); ++slot; }
          }
          if constexpr (MODE == MODE_GABOR2D_FWD) {
            if (P.store_mask & 4) { stage_row_fmt(wbuf + slot * 4096, lane, o2, P.o_fmt[slot]); ++slot; }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            for (int s = 0; s < n_out; ++s) tma_store_2d(&P.o_map[s], wbuf + s * 4096, c, row0 + q * 32);
            tma_store_commit();
          }
        }
      }

      if constexpr (kFwd) {
        if (E.fuse_final && sl == P.slices - 1) {
          // the two warps of a sub-partition hold partial sums over even / odd chunks (of every slice)
          const int slot = it & 1;
          if (half == 1) {
            *reinterpret_cast<float4*>(&fin_xchg[slot][q * 32 + lane][0]) = make_float4(facc[0], facc[1], facc[2], facc[3]);
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          if (half == 0 && row_ok) {
            const float4 o = *reinterpret_cast<const float4*>(&fin_xchg[slot][q * 32 + lane][0]);
            const float r[4] = {facc[0] + o.x, facc[1] + o.y, facc[2] + o.z, facc[3] + o.w};
#pragma unroll
            for (int oo = 0; oo < kMaxOut; ++oo)
              if (oo < E.out_features) E.out[size_t(row) * E.out_features + oo] = r[oo] + __ldg(E.bf + 2 * oo);
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
    if (dbg && ew == 0 && lane == 0) { dbg[kDbgEpiWaitAcc] = e_wait; dbg[kDbgEpiTotal] = WIRE_CLK() - e_begin; dbg[kDbgEpiWaitIn] = e_wait_in; }
  }

  // ===================== teardown =====================
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer still uses the pair's resources
  if (warp == 1) {
    tc_fence_after();
    if (pair) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace wire
