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
// Work item = (K split, g matrix, column block, 128-row M tile); one item per CTA.
#pragma once
#include "sm100.cuh"

namespace wire {

constexpr int kWgradThreads = 192;
constexpr int kWgradKC = 32;  // coordinates per pipeline stage

struct WgradParams {
  CUtensorMap x_map;     // x [N, 2K+1(+pad)], box {32 cols, 32 rows}, SWIZZLE_128B_ATOM_32B
  CUtensorMap g_map[2];  // g_z (and g_w for wire2d) [N, 2M], box {32 cols, 32 rows}
  int n_rows;
  int k_in;      // complex input features K (x has 2K+1 meaningful columns)
  int g_cols;    // 2M
  int n_g;       // 1 = wire, 2 = wire2d (linear + scale_orth)
  int m_tiles;   // ceil((2K+1)/128)
  int n_blocks;  // column blocks per g matrix
  int nb;        // columns per block (multiple of 32, <= 448)
  int splits;
  int stages;
  float* gW[2];  // [M][K][2] fp32, accumulated
  float* gB[2];  // [M][2]
};

__global__ void __launch_bounds__(kWgradThreads, 1) tc_wgrad_kernel(const __grid_constant__ WgradParams P) {
  using namespace sm100;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8];
  __shared__ __align__(8) uint64_t bar_empty[8];
  __shared__ __align__(8) uint64_t bar_tmem_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t blk_bytes = kWgradKC * 128;  // one 32-column block of one stage
  const uint32_t a_bytes = 4 * blk_bytes;
  const uint32_t b_bytes = uint32_t(P.nb / 32) * blk_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  // decode the work item
  int item = blockIdx.x;
  const int mt = item % P.m_tiles;  item /= P.m_tiles;
  const int nblk = item % P.n_blocks;  item /= P.n_blocks;
  const int gi = item % P.n_g;  item /= P.n_g;
  const int split = item;
  const int total_chunks = (P.n_rows + kWgradKC - 1) / kWgradKC;
  const int cps = (total_chunks + P.splits - 1) / P.splits;
  const int ch_begin = split * cps;
  int ch_end = ch_begin + cps;
  ch_end = ch_end > total_chunks ? total_chunks : ch_end;
  const int n_chunks = ch_end > ch_begin ? ch_end - ch_begin : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_tmem_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (n_chunks > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const int nbb = P.nb / 32;
        for (int ch = ch_begin; ch < ch_end; ++ch) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t full = smem_u32(&bar_full[stage]);
          mbar_expect_tx(full, stage_bytes);
          const uint32_t a_dst = smem_base + stage * stage_bytes;
          const int r0 = ch * kWgradKC;
          for (int b = 0; b < 4; ++b)
            tma_load_2d(a_dst + b * blk_bytes, &P.x_map, full, mt * 128 + b * 32, r0);
          for (int b = 0; b < nbb; ++b)
            tma_load_2d(a_dst + a_bytes + b * blk_bytes, &P.g_map[gi], full, nblk * P.nb + b * 32, r0);
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // valid accumulator columns of this block, rounded up to the UMMA N granularity
        int nvalid = P.g_cols - nblk * P.nb;
        nvalid = nvalid > P.nb ? P.nb : nvalid;
        const int np = (nvalid + 15) & ~15;
        const int n1 = np > 256 ? 256 : np;
        const int n2 = np - n1;
        const uint32_t idesc1 = make_idesc_tf32(128, n1, true, true);
        const uint32_t idesc2 = make_idesc_tf32(128, n2 > 0 ? n2 : 16, true, true);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_chunks; ++i) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_base = smem_base + stage * stage_bytes;
          const uint32_t b_base = a_base + a_bytes;
#pragma unroll
          for (int ks = 0; ks < kWgradKC / 8; ++ks) {
            const uint32_t acc = (i | ks) ? 1u : 0u;
            const uint64_t adesc = make_sdesc(a_base + ks * 1024, blk_bytes, 512, kLayoutSW128Base32);
            const uint64_t bdesc = make_sdesc(b_base + ks * 1024, blk_bytes, 512, kLayoutSW128Base32);
            umma_tf32(tmem_base, adesc, bdesc, idesc1, acc);
            if (n2 > 0) {
              const uint64_t bdesc2 = make_sdesc(b_base + 8 * blk_bytes + ks * 1024, blk_bytes, 512, kLayoutSW128Base32);
              umma_tf32(tmem_base + n1, adesc, bdesc2, idesc2, acc);
            }
          }
          umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&bar_tmem_full));
      }
    } else {
      const int q = warp & 3;
      const int c = mt * 128 + q * 32 + lane;  // row of G = real column of x
      const int two_k = 2 * P.k_in;
      float* gW = P.gW[gi];
      float* gB = P.gB[gi];
      int nvalid = P.g_cols - nblk * P.nb;
      nvalid = nvalid > P.nb ? P.nb : nvalid;
      const int nchunks = (nvalid + 31) / 32;
      mbar_wait(smem_u32(&bar_tmem_full), 0);
      tc_fence_after();
      for (int ch = 0; ch < nchunks; ++ch) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + ch * 32, raw);
        tmem_wait_ld();
        const int r0 = nblk * P.nb + ch * 32;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a0 = __uint_as_float(raw[2 * i]);
          const float a1 = __uint_as_float(raw[2 * i + 1]);
          const float other = __shfl_xor_sync(0xffffffffu, a1, 1);
          const int r = r0 + 2 * i;
          if (r < P.g_cols) {
            if (c < two_k) {
              const float val = (lane & 1) ? (other - a0) : (a0 + other);
              atomicAdd(gW + size_t(r >> 1) * two_k + c, val);
            } else if (c == two_k) {
              atomicAdd(gB + r, a0);
              atomicAdd(gB + r + 1, a1);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace wire
