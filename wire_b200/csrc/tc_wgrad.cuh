// tc_wgrad.cuh — weight/bias gradient of a complex hidden Linear on tcgen05 (sm_100a).
//
// PyTorch's complex Linear backward (autograd of modules/wire.py:89 / wire2d.py:57-60):
//     g_W[j,k] = sum_n g_z[n,j] * conj(x[n,k])        g_b[j] = sum_n g_z[n,j]
// is computed as ONE real GEMM over the coordinate axis with both operands MN-major (the
// coordinate index is the slow axis of both tensors in HBM, so no transposed copies exist):
//     G[c, r] = sum_n X[n, c] * Gz[n, r]      c in [0, 2K+1), r in [0, 2M)
// X is the interleaved-complex input with a trailing "ones" column (column 2K) so the bias
// gradient falls out of the same GEMM as row 2K.  The epilogue folds the real 2x2 blocks back
// into complex numbers with one lane-pair shuffle:
//     Re g_W[j,k] = G[2k,2j] + G[2k+1,2j+1]       Im g_W[j,k] = G[2k,2j+1] - G[2k+1,2j]
// and accumulates split-K partials with fp32 red.global.add into the (pre-zeroed) flat gradient.
//
// Work item = (K split, g matrix, column block, M tile of 128 x-columns per CTA); one item per CTA, or per CTA
// pair (cta_group::2: 256 x-columns, each CTA stages its own x tile and HALF of the g tile).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "f32x2.cuh"
#include "gabor_math.cuh"
#include "sm100.cuh"

namespace wire {

constexpr int kWgradThreads = 192;      // TF32 kernels: producer + MMA issuer + 4 epilogue warps
constexpr int kWgradEpi16Warps = 8;     // 16-bit kernels: 8 epilogue / converter warps (two per TMEM sub-partition)
constexpr int kWgradThreads16 = 64 + 32 * kWgradEpi16Warps;
constexpr int wgrad_threads(bool gen, bool op16) { return op16 ? kWgradThreads16 : kWgradThreads + (gen ? 32 * 4 : 0); }
constexpr int kWgradGenWarps = 4;  // GEN kernels: x = y0 = gabor(coords W0^T + b0) is computed in place (first hidden layer)
constexpr int kWgradKC = 32;    // coordinates per pipeline stage (TF32 operands: 32 rows x 128 B per 32-column block)
constexpr int kWgradKC16 = 64;  // 16-bit operands: 64 rows x 128 B per 64-column block (8 KB blocks, 128 B swizzle)

struct WgradParams {
  CUtensorMap x_map;     // x [N, 2K+1(+pad)], box {32 cols, 32 rows}, SWIZZLE_128B_ATOM_32B
  CUtensorMap g_map[2];  // g_z (and g_w for wire2d) [N, 2M], box {32 cols, 32 rows}
  int n_rows;
  int k_in;      // complex input features K (x has 2K+1 meaningful columns)
  int g_cols;    // 2M
  int n_g;       // 1 = wire, 2 = wire2d (linear + scale_orth)
  int cluster;   // 1 = cta_group::1 (128 x-columns per tile); 2 = CTA pair (256 x-columns per tile, g tile split)
  int m_tiles;   // ceil((2K+1) / (128 * cluster))
  int n_blocks;  // column blocks per g matrix
  int nb;        // columns per block (multiple of 32, <= 448)
  int splits;
  int stages;
  float* gW[2];  // [M][K][2] fp32, accumulated
  float* gB[2];  // [M][2]
  int x_fmt, g_fmt;  // OP16 kernels: operand formats of x and g in HBM (sm100::kFmtF16 / kFmtBF16)
  int bias_sum;      // OP16 + x_conv: the x tiles do NOT include the "ones" column (2K is a multiple of the tile width, so column 2K
                     // would cost a whole extra tile: wire2d at M = 128 runs 3 x-tiles of 128 instead of one pair tile of 256 for
                     // it); the bias gradient g_b = sum_n g is summed by the converter warps from the g tiles in shared memory
  int dual;          // bias_sum + pair + two g tensors of <= 256 padded columns (wire2d: g_z and g_w): ONE work item carries both
                     // (MMA piece 1 = g_map[0], piece 2 = g_map[1], all 512 TMEM columns), so the x tile is staged and converted
                     // once for both and a CTA ingests 48 KB per 1024 cycles of MMA work instead of 32 KB per 512
  int x_conv;        // OP16: x is stored FP16 but g is BF16.  kind::f16 cannot mix the two (illegal instruction on
                     // sm_100a, profiles/r01_probe16.log), so the four epilogue warps -- idle during the K loop --
                     // convert each landed x tile FP16 -> BF16 in place in shared memory (element-wise, so the swizzle
                     // does not matter); activations are O(1), so BF16's range is not an issue, only its 8-bit mantissa.
  // GEN: first-layer description (x is generated, never loaded)
  const float* coords;
  int in_features;
  const float* w0;
  const float* b0;
  const float* w0b;
  const float* b0b;
  const float* gen_omega;
  const float* gen_scale;
  int gen_two_d;
  uint32_t gen_tab_off;  // byte offset of the {w0[0..2], b0} tables inside dynamic smem
  int gen_tab_feats;     // padded feature count of the tables
  unsigned long long* dbg;  // optional per-CTA time stamps [8] (tools/umma_probe): entry, prologue done, K loop done, epilogue done
};

// OP16: x is FP16 and g is BF16 (formats in P.x_fmt / P.g_fmt), both still MN-major; tiles are 64-column blocks of
// 128-byte rows with the 128 B swizzle (UMMA layout SW128; 8 K-rows = one 1024 B atom, SBO = 1024, LBO = block stride),
// 64 coordinates per stage, MMA kind::f16 (K = 16 = 2048 B per step).  (A first version used 32-column blocks of
// 64-byte rows: every TMA row request then moved only two sectors and the loads, not the MMAs, set the pace.)
template <bool PAIR, bool GEN = false, bool OP16 = false>
__global__ void __launch_bounds__(wgrad_threads(GEN, OP16), 1) tc_wgrad_kernel(const __grid_constant__ WgradParams P) {
  using namespace sm100;
  static_assert(!(GEN && OP16), "the in-place generator writes TF32 tiles");
  constexpr int kKC = OP16 ? kWgradKC16 : kWgradKC;
  constexpr uint32_t kLayout = OP16 ? kLayoutSW128 : kLayoutSW128Base32;
  constexpr int kBlkCols = OP16 ? 64 : 32;       // columns per smem block (one 128-byte swizzle row)
  constexpr int kXBlocks = 128 / kBlkCols;       // x blocks per CTA (128 x-columns)
  constexpr uint32_t kSBO = OP16 ? 1024 : 512;
  constexpr uint32_t kStepUnits = OP16 ? 128 : 64;  // descriptor start-address advance per K-step (bytes >> 4)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8];
  __shared__ __align__(8) uint64_t bar_empty[8];
  __shared__ __align__(8) uint64_t bar_tmem_full;
  __shared__ __align__(8) uint64_t bar_x[8];  // x_conv: this CTA's x tile has landed (own barrier; converters wait on it)
  __shared__ __align__(8) uint64_t bar_g[8];  // bias_sum: this CTA's g tile has landed (own barrier; the converters sum its columns)
  __shared__ uint32_t tmem_slot;

  constexpr int C = PAIR ? 2 : 1;
  // epilogue warps (they also convert the x tiles during the K loop): the 16-bit kernels run eight — the FP16 -> BF16
  // conversion sits on the critical path of every pipeline stage and the red.add stream of the epilogue is bound by how many
  // warps issue it (18 us fixed per launch with four, profiles/r01: 12.9 k cycles)
  constexpr int kEpiW = OP16 ? kWgradEpi16Warps : 4;
  const bool conv = OP16 && P.x_conv;
  const bool bsum = conv && P.bias_sum;
  const bool dual = PAIR && bsum && P.dual;
  unsigned long long* dbg = P.dbg ? P.dbg + size_t(blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 64) dbg[0] = clock64();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = PAIR ? int(cluster_ctarank()) : 0;
  const bool leader = crank == 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t blk_bytes = uint32_t(kKC) * 128;  // one block of one stage: kKC rows of 128 B
  const uint32_t a_bytes = kXBlocks * blk_bytes;
  // MMA pieces along N (g columns): n1 + n2 = nb ; a pair splits each piece in halves (multiples of 32 columns)
  int nvalid = P.g_cols;  // per column block below
  // g blocks this CTA stages per chunk: its half of each MMA piece (n1 <= 256, n2), each rounded up to whole blocks
  const int n1_all = P.nb > 256 ? 256 : P.nb, n2_all = P.nb - n1_all;
  const int pb1 = (n1_all / C + kBlkCols - 1) / kBlkCols, pb2 = (n2_all / C + kBlkCols - 1) / kBlkCols;
  const int nbb_cta = PAIR ? pb1 + pb2 : P.nb / kBlkCols;
  const uint32_t b_bytes = uint32_t(nbb_cta) * blk_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  // decode the work item (all CTAs of a pair share it)
  int item = blockIdx.x / C;
  const int mt = item % P.m_tiles;  item /= P.m_tiles;
  const int nblk = item % P.n_blocks;  item /= P.n_blocks;
  const int gi = item % P.n_g;  item /= P.n_g;
  const int split = item;
  const int total_chunks = (P.n_rows + kKC - 1) / kKC;
  const int cps = (total_chunks + P.splits - 1) / P.splits;
  const int ch_begin = split * cps;
  int ch_end = ch_begin + cps;
  ch_end = ch_end > total_chunks ? total_chunks : ch_end;
  const int n_chunks = ch_end > ch_begin ? ch_end - ch_begin : 0;
  nvalid = P.g_cols - nblk * P.nb;
  nvalid = nvalid > P.nb ? P.nb : nvalid;
  // single CTA: pieces rounded to the UMMA N granularity; pair: fixed 64-column-aligned pieces of the padded block
  const int np = PAIR ? P.nb : ((nvalid + 15) & ~15);
  const int n1 = np > 256 ? 256 : np;
  const int n2 = np - n1;

  float4* gtab = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + P.gen_tab_off);
  if constexpr (GEN) {
    float* gt = reinterpret_cast<float*>(gtab);
    for (int i = threadIdx.x; i < P.gen_tab_feats * 4; i += blockDim.x) {
      const int k = i >> 2, j = i & 3;
      float v = 0.f, v2 = 0.f;
      if (k < P.k_in) {
        if (j < 3) {
          if (j < P.in_features) { v = P.w0[size_t(k) * P.in_features + j]; if (P.gen_two_d) v2 = P.w0b[size_t(k) * P.in_features + j]; }
        } else { v = P.b0[k]; if (P.gen_two_d) v2 = P.b0b[k]; }
      }
      gt[i] = v;
      gt[P.gen_tab_feats * 4 + i] = v2;
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      // bias_sum: x and g both land on this CTA's own barriers; the MMA's barrier only counts the converter warps' arrivals
      mbar_init(smem_u32(&bar_full[s]), GEN ? 1 + kWgradGenWarps * C : (bsum ? kEpiW * C : (conv ? 1 + kEpiW * C : 1)));
      mbar_init(smem_u32(&bar_empty[s]), 1);
      mbar_init(smem_u32(&bar_x[s]), 1);
      mbar_init(smem_u32(&bar_g[s]), 1);
    }
    mbar_init(smem_u32(&bar_tmem_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_2cta(smem_u32(&tmem_slot), 512); tmem_relinquish_2cta(); }
    else { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();
  pdl_wait();  // x / g_z tiles (and the gradient buffers this kernel accumulates into) belong to earlier kernels
  if (dbg && threadIdx.x == 64) dbg[1] = clock64();

  if (n_chunks > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const int x_col0 = (mt * C + crank) * 128;
        for (int ch = ch_begin; ch < ch_end; ++ch) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t full_own = smem_u32(&bar_full[stage]);
          const uint32_t a_dst = smem_base + stage * stage_bytes;
          const int r0 = ch * kKC;
          const uint32_t tx_bytes = (GEN || conv) ? b_bytes : stage_bytes;
          if (conv) {  // the x tile completes on this CTA's own barrier: its converter warps pick it up
            const uint32_t xb = smem_u32(&bar_x[stage]);
            mbar_expect_tx(xb, a_bytes);
            for (int b = 0; b < kXBlocks; ++b) tma_load_2d(a_dst + b * blk_bytes, &P.x_map, xb, x_col0 + b * kBlkCols, r0);
          }
          if (bsum) {  // g tile on this CTA's own barrier (the same blocks as below)
            const uint32_t gb = smem_u32(&bar_g[stage]);
            mbar_expect_tx(gb, b_bytes);
            if (!PAIR) {
              for (int b = 0; b < nbb_cta; ++b) tma_load_2d(a_dst + a_bytes + b * blk_bytes, &P.g_map[gi], gb, nblk * P.nb + b * kBlkCols, r0);
            } else {
              for (int b = 0; b < pb1; ++b)
                tma_load_2d(a_dst + a_bytes + b * blk_bytes, &P.g_map[gi], gb, nblk * P.nb + crank * (n1 / 2) + b * kBlkCols, r0);
              // piece 2: the second half of the same tensor, or (dual) this CTA's half of the second tensor
              for (int b = 0; b < pb2; ++b)
                tma_load_2d(a_dst + a_bytes + (pb1 + b) * blk_bytes, &P.g_map[dual ? 1 : gi], gb,
                            (dual ? 0 : nblk * P.nb + n1) + crank * (n2 / 2) + b * kBlkCols, r0);
            }
          } else if (!PAIR) {
            mbar_expect_tx(full_own, tx_bytes);
            if (!GEN && !conv) for (int b = 0; b < kXBlocks; ++b) tma_load_2d(a_dst + b * blk_bytes, &P.x_map, full_own, x_col0 + b * kBlkCols, r0);
            for (int b = 0; b < nbb_cta; ++b)
              tma_load_2d(a_dst + a_bytes + b * blk_bytes, &P.g_map[gi], full_own, nblk * P.nb + b * kBlkCols, r0);
          } else {
            const uint32_t full_leader = full_own & kPeerBitMask;
            if (leader) mbar_expect_tx(full_own, 2 * tx_bytes);
            if (!GEN && !conv)
              for (int b = 0; b < kXBlocks; ++b)
                tma_load_2d_2cta(a_dst + b * blk_bytes, &P.x_map, full_leader, x_col0 + b * kBlkCols, r0, kEvictNormal);
            // piece 1: columns [crank*n1/2, +n1/2) ; piece 2: columns [n1 + crank*n2/2, +n2/2)
            const int p1 = pb1, p2 = pb2;  // (a half that ends inside a block still loads the whole block; the MMA ignores the rest)
            for (int b = 0; b < p1; ++b)
              tma_load_2d_2cta(a_dst + a_bytes + b * blk_bytes, &P.g_map[gi], full_leader,
                               nblk * P.nb + crank * (n1 / 2) + b * kBlkCols, r0, kEvictNormal);
            for (int b = 0; b < p2; ++b)
              tma_load_2d_2cta(a_dst + a_bytes + (p1 + b) * blk_bytes, &P.g_map[gi], full_leader,
                               nblk * P.nb + n1 + crank * (n2 / 2) + b * kBlkCols, r0, kEvictNormal);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && (!PAIR || leader)) {
        const uint32_t xf = conv ? uint32_t(P.g_fmt) : uint32_t(P.x_fmt);
        const uint32_t idesc1 = OP16 ? make_idesc_f16(PAIR ? 256 : 128, n1, true, true, xf, uint32_t(P.g_fmt))
                                     : make_idesc_tf32(PAIR ? 256 : 128, n1, true, true);
        const uint32_t idesc2 = OP16 ? make_idesc_f16(PAIR ? 256 : 128, n2 > 0 ? n2 : 16, true, true, xf, uint32_t(P.g_fmt))
                                     : make_idesc_tf32(PAIR ? 256 : 128, n2 > 0 ? n2 : 16, true, true);
        // descriptor words precomputed; only the start-address field moves (stage, K-step = 1024 B: 8 rows of 128 B,
        // or 16 rows of 64 B for 16-bit operands)
        const uint32_t desc_hi = uint32_t(make_sdesc(0, blk_bytes, kSBO, kLayout) >> 32);
        const uint32_t a_lo0 = uint32_t(make_sdesc(smem_base, blk_bytes, kSBO, kLayout));
        const uint32_t stage_units = stage_bytes >> 4, b_units = a_bytes >> 4;
        const uint32_t b2_units = (uint32_t(PAIR ? pb1 : n1 / kBlkCols) * blk_bytes) >> 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_chunks; ++i) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * stage_units;
          const uint32_t b_lo = a_lo + b_units;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acc = (ks > 0) ? 1u : (i ? 1u : 0u);
            const uint64_t adesc = (uint64_t(desc_hi) << 32) | (a_lo + kStepUnits * ks);
            const uint64_t bdesc = (uint64_t(desc_hi) << 32) | (b_lo + kStepUnits * ks);
            const uint64_t bdesc2 = (uint64_t(desc_hi) << 32) | (b_lo + b2_units + kStepUnits * ks);
            if constexpr (OP16) {
              if (PAIR) {
                umma_f16_2cta(tmem_base, adesc, bdesc, idesc1, acc);
                if (n2 > 0) umma_f16_2cta(tmem_base + n1, adesc, bdesc2, idesc2, acc);
              } else {
                umma_f16(tmem_base, adesc, bdesc, idesc1, acc);
                if (n2 > 0) umma_f16(tmem_base + n1, adesc, bdesc2, idesc2, acc);
              }
            } else if (PAIR) {
              umma_tf32_2cta(tmem_base, adesc, bdesc, idesc1, acc);
              if (n2 > 0) umma_tf32_2cta(tmem_base + n1, adesc, bdesc2, idesc2, acc);
            } else {
              umma_tf32(tmem_base, adesc, bdesc, idesc1, acc);
              if (n2 > 0) umma_tf32(tmem_base + n1, adesc, bdesc2, idesc2, acc);
            }
          }
          if (PAIR) umma_commit_2cta_mcast(smem_u32(&bar_empty[stage]), 3);
          else umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma_commit_2cta_mcast(smem_u32(&bar_tmem_full), 3);
        else umma_commit(smem_u32(&bar_tmem_full));
      }
    } else if (GEN && warp >= 2 + kEpiW) {
      // ===================== x-operand generator warps =====================
      // warp b writes column block b (16 complex features) of every stage; lane = coordinate row of the chunk.
      // MN-major tile, 128B swizzle with 32B atoms: 32-byte chunk index ^= (row & 3).
      const int b = warp - (2 + kEpiW);
      const GaborConst G0 = make_gabor_const(__ldg(P.gen_omega), __ldg(P.gen_scale));
      const float4* tab = gtab;
      const float4* tab2 = gtab + P.gen_tab_feats;
      const int feat0 = ((mt * C + crank) * 128 + b * 32) >> 1;  // first complex feature of this block
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = ch_begin; ch < ch_end; ++ch) {
        const int row = ch * kWgradKC + lane;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (row < P.n_rows) {
          c0 = __ldg(P.coords + size_t(row) * P.in_features);
          if (P.in_features > 1) c1 = __ldg(P.coords + size_t(row) * P.in_features + 1);
          if (P.in_features > 2) c2 = __ldg(P.coords + size_t(row) * P.in_features + 2);
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = feat0 + i;
          float yr = 0.f, yi = 0.f;
          if (row < P.n_rows) {
            if (j < P.k_in) {
              const float4 t = tab[j];
              const float z0 = fmaf(c0, t.x, fmaf(c1, t.y, fmaf(c2, t.z, t.w)));
              float wn = 0.f;
              if (P.gen_two_d) {
                const float4 t2 = tab2[j];
                const float w0v = fmaf(c0, t2.x, fmaf(c1, t2.y, fmaf(c2, t2.z, t2.w)));
                wn = w0v * w0v;
              }
              gabor_fast(G0, z0, 0.f, wn, yr, yi);
              yr = round_tf32(yr);
              yi = round_tf32(yi);
            } else if (j == P.k_in) {
              yr = 1.0f;  // the "ones" column: row 2K of G is the bias gradient
            }
          }
          v[2 * i] = yr;
          v[2 * i + 1] = yi;
        }
        if (lane == 0) mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
        __syncwarp();
        const uint32_t rowaddr = smem_base + stage * stage_bytes + b * blk_bytes + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = rowaddr + ((((j >> 1) ^ (lane & 3)) << 5) | ((j & 1) << 4));
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                       "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const uint32_t full_own = smem_u32(&bar_full[stage]);
          if (PAIR) mbar_arrive_cluster(full_own & kPeerBitMask); else mbar_arrive(full_own);
        }
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    } else {
      if (conv) {
        // ===================== x converters (epilogue warps, during the K loop) =====================
        // warp b converts slice b (16 KB / kEpiW = 16-byte pieces per lane) of every landed x tile FP16 -> BF16 in place.
        // bias_sum: the same warps add up the columns of this CTA's g tile (only the CTAs of the first x tile: every x tile
        // sees the same g).  Thread -> one 16-byte piece (8 BF16 columns) `pc` of the tile's rows rb, rb + rpp, ...: four
        // shared loads in flight, then two integer ops and one packed add per column pair, FP32 sums in 8 registers for the
        // whole kernel.  (A first version summed 32-bit words with per-item index arithmetic: 2 440 cycles per chunk against
        // 1 540 without it, profiles/r02_probe_wgrad_variants.log.)
        const int ctid = (warp - 2) * 32 + lane;
        const bool do_sum = bsum && mt == 0;
        const int ppr = nbb_cta * 8;                 // 16-byte pieces per tile row
        const int rpp = (32 * kEpiW) / ppr;          // rows per pass of the converter threads
        const bool sum_thread = do_sum && ctid < rpp * ppr;
        const int pc = ctid % ppr, rb = ctid / ppr;
        const uint32_t piece_off = uint32_t(pc >> 3) * blk_bytes;
        f2 bacc[4] = {0ull, 0ull, 0ull, 0ull};
        auto x_done = [&](int stage, uint32_t phase) {
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          if (bsum) {
            mbar_wait(smem_u32(&bar_g[stage]), phase);   // every converter warp waits: its arrival below also vouches for the g tile
            if (sum_thread) {
              const uint32_t gbase = smem_base + stage * stage_bytes + a_bytes + piece_off;
              for (int r4 = rb; r4 < kKC; r4 += 4 * rpp) {
                uint32_t v[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t r = uint32_t(r4 + j * rpp);
                  if (r < uint32_t(kKC)) {
                    // SW128: 16-byte chunk (pc % 8) of row r sits at chunk position (pc % 8) ^ (r % 8)
                    const uint32_t addr = gbase + r * 128 + (((uint32_t(pc) & 7u) ^ (r & 7u)) << 4);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[j][0]), "=r"(v[j][1]), "=r"(v[j][2]), "=r"(v[j][3]) : "r"(addr) : "memory");
                  } else {
                    v[j][0] = v[j][1] = v[j][2] = v[j][3] = 0u;
                  }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                  for (int k = 0; k < 4; ++k) bacc[k] = f2_add(bacc[k], f2_bits(v[j][k] << 16, v[j][k] & 0xffff0000u));
              }
            }
          }
          __syncwarp();
          if (lane == 0) {
            const uint32_t full_own = smem_u32(&bar_full[stage]);
            if (PAIR) mbar_arrive_cluster(full_own & kPeerBitMask); else mbar_arrive(full_own);
          }
        };
        int stage = 0;
        uint32_t phase = 0;
        const int b = warp - 2;
        constexpr int kPieces = 16384 / kEpiW / 512;
        for (int i = 0; i < n_chunks; ++i) {
          mbar_wait(smem_u32(&bar_x[stage]), phase);
          const uint32_t blk = smem_base + stage * stage_bytes + b * (16384 / kEpiW);
          uint32_t h[kPieces][4];
#pragma unroll
          for (int j = 0; j < kPieces; ++j) {
            const uint32_t addr = blk + (j * 32 + lane) * 16;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h[j][0]), "=r"(h[j][1]), "=r"(h[j][2]), "=r"(h[j][3]) : "r"(addr) : "memory");
          }
#pragma unroll
          for (int j = 0; j < kPieces; ++j) {
            const uint32_t addr = blk + (j * 32 + lane) * 16;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[j][k]));
              const __nv_bfloat162 t = __floats2bfloat162_rn(f.x, f.y);
              h[j][k] = *reinterpret_cast<const uint32_t*>(&t);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[j][0]), "r"(h[j][1]), "r"(h[j][2]), "r"(h[j][3]) : "memory");
          }
          x_done(stage, phase);
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        if (sum_thread) {
          // column 8 pc + e of this CTA's staged blocks -> real column of g.  Single CTA: block-contiguous; pair: this CTA's half of
          // each MMA piece (a half that ends inside a block loaded the whole block: only the columns of the half are this CTA's to
          // count); dual: piece 2 is the second tensor
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float val = (e & 1) ? f2_hi(bacc[e >> 1]) : f2_lo(bacc[e >> 1]);
            int col = 8 * pc + e, gcol;
            bool own = true;
            float* gB = P.gB[gi];
            int limit = nblk * P.nb + nvalid;
            if (!PAIR) gcol = nblk * P.nb + col;
            else if (col < pb1 * kBlkCols) {
              own = col < n1 / 2; gcol = nblk * P.nb + crank * (n1 / 2) + col;
              if (dual) limit = P.g_cols;
            } else {
              col -= pb1 * kBlkCols; own = col < n2 / 2;
              if (dual) { gB = P.gB[1]; gcol = crank * (n2 / 2) + col; limit = P.g_cols; }
              else gcol = nblk * P.nb + n1 + crank * (n2 / 2) + col;
            }
            if (own && gcol < limit) atomicAdd(gB + gcol, val);
          }
        }
      }
      const int q = warp & 3;
      const int part = (warp - 2) >> 2;                      // which of the sub-partition's kEpiW / 4 warps
      const int c = (mt * C + crank) * 128 + q * 32 + lane;  // row of G = real column of x
      const int two_k = 2 * P.k_in;
      float* gW = P.gW[gi];
      float* gB = P.gB[gi];
      const int nchunks = dual ? 16 : (nvalid + 31) / 32;   // dual: TMEM columns [0, 256) = first tensor, [256, 512) = second
      mbar_wait(smem_u32(&bar_tmem_full), 0);
      tc_fence_after();
      if (dbg && threadIdx.x == 64) dbg[2] = clock64();
      // (Starting every K split at a different chunk, so that the ~37 CTAs adding into the same gradient entries do not walk
      // the same L2 lines in step, changed nothing: the epilogue was bound by its own instruction stream, see below.)
      for (int ch = part; ch < nchunks; ch += kEpiW / 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + ch * 32, raw);
        tmem_wait_ld();
        const bool second = dual && ch >= 8;
        const int r0 = second ? (ch - 8) * 32 : nblk * P.nb + ch * 32;
        // all 16 lane-pair exchanges first (independent shuffles in flight together), then straight-line reductions:
        // the per-element branches of the first version serialised shuffle -> branch -> address -> red (80 cycles each)
        float other[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) other[i] = __shfl_xor_sync(0xffffffffu, __uint_as_float(raw[2 * i + 1]), 1);
        if (c < two_k) {
          float* dst = (second ? P.gW[1] : gW) + size_t(r0 >> 1) * two_k + c;
          const int n_ok = (P.g_cols - r0 + 1) >> 1;  // complex outputs of this chunk inside the matrix (warp-uniform)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a0 = __uint_as_float(raw[2 * i]);
            const float val = (lane & 1) ? (other[i] - a0) : (a0 + other[i]);
            if (i < n_ok) atomicAdd(dst + size_t(i) * two_k, val);
          }
        } else if (c == two_k) {  // the "ones" column: one lane of one CTA per tile
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = r0 + 2 * i;
            if (r < P.g_cols) {
              atomicAdd(gB + r, __uint_as_float(raw[2 * i]));
              atomicAdd(gB + r + 1, __uint_as_float(raw[2 * i + 1]));
            }
          }
        }
      }
    }
  }

  if (dbg && threadIdx.x == 64) dbg[3] = clock64();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
  if (dbg && threadIdx.x == 64) dbg[4] = clock64();
}

}  // namespace wire
