// simt16_kernels.cuh — the CUDA-core kernels of the mixed16 whole-network path: the stages that are too skinny for
// tensor cores, rebuilt for 16-bit tensors.  With the activation bytes halved these stages stopped being HBM-bound
// (profiles/r01_bench_v10_mixed16.json: first_fwd 0.09 ms, top_bwd 0.24 ms, first_wgrad 0.08 ms at 30-40 % of HBM
// peak): they are bound by the XU pipe (ex2 / sin / cos, 16 lanes per SM per clock), by issue slots and by how many
// bytes each SM keeps in flight, so each is organised around those three limits.
//
//   first_fwd16_kernel    first ComplexGaborLayer (real z, K = 2..3) -> FP16 y         modules/wire.py:88-93 (is_first)
//   top_bwd16_kernel      final Linear backward + Gabor backward of the last hidden layer (autograd of wire.py:156-165
//                         and :88-93): FP16 z tiles in by TMA, BF16 g_z tiles out by TMA, g_Wf / g_bf in registers
//   first_wgrad16_kernel  g_W0 = g_z0^T c, g_b0 = sum g_z0 from the BF16 g_z0 the first-layer dgrad epilogue writes
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gabor_math.cuh"
#include "sm100.cuh"
#include "tc_rows16.cuh"  // gabor16, pack helpers

namespace wire {

// ---------------------------------------------------------------------------------------------
// first layer forward.  Work item = (row, quad of 4 complex features) dealt round-robin to the threads of a block
// (every lane busy whatever M is); the per-feature {w0[0..2], b0} table and the block's coordinates sit in smem;
// one 16-byte store (8 halves) per item.  y: FP16 [n][y_pitch]; columns >= 2M (the "ones" column) are not touched.
// ---------------------------------------------------------------------------------------------
// Persistent blocks (a few per SM) own a contiguous row range, so the table is built once per block.
template <bool TWO_D>
__global__ void __launch_bounds__(256) first_fwd16_kernel(const float* __restrict__ coords, int n, int in_f, int M,
                                                           const float* __restrict__ W0, const float* __restrict__ b0,
                                                           const float* __restrict__ W0b, const float* __restrict__ b0b,
                                                           const float* __restrict__ omega_p, const float* __restrict__ scale_p,
                                                           __half* __restrict__ y, int y_pitch, int rows_per_block) {
  extern __shared__ __align__(16) float4 fsm[];
  const int nq = (M + 3) >> 2;
  // table entry of feature k = 4q + f sits at [f * nq + q]: the lanes of a warp (consecutive q) then read consecutive
  // float4 (the [q][f] order made every LDS.128 a 4-way bank conflict: 21.6 M conflicts, smem wavefronts bound the kernel)
  float4* tab = fsm;            // [4][nq]
  float4* tab2 = fsm + 4 * nq;  // [4][nq] (wire2d scale_orth)
  const int row0 = blockIdx.x * rows_per_block;
  int rows = n - row0;
  rows = rows > rows_per_block ? rows_per_block : rows;
  if (rows <= 0) return;
  for (int k = threadIdx.x; k < 4 * nq; k += blockDim.x) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f), t2 = t;
    if (k < M) {
      t.x = W0[size_t(k) * in_f];
      if (in_f > 1) t.y = W0[size_t(k) * in_f + 1];
      if (in_f > 2) t.z = W0[size_t(k) * in_f + 2];
      t.w = b0[k];
      if constexpr (TWO_D) {
        t2.x = W0b[size_t(k) * in_f];
        if (in_f > 1) t2.y = W0b[size_t(k) * in_f + 1];
        if (in_f > 2) t2.z = W0b[size_t(k) * in_f + 2];
        t2.w = b0b[k];
      }
    }
    tab[(k & 3) * nq + (k >> 2)] = t;
    if constexpr (TWO_D) tab2[(k & 3) * nq + (k >> 2)] = t2;
  }
  __syncthreads();
  const GaborConst2 G2 = make_gabor_const2(make_gabor_const(__ldg(omega_p), __ldg(scale_p)));
  const int items = rows * nq;
  // (r, q) of this thread's first item and the per-step increment, without a division in the loop
  int r = int(threadIdx.x) / nq, q = int(threadIdx.x) - r * nq;
  const int dr = int(blockDim.x) / nq, dq = int(blockDim.x) - dr * nq;
  for (int item = threadIdx.x; item < items; item += blockDim.x) {
    // the ~nq threads that share a row hit the same L1 line
    const float* cp = coords + size_t(row0 + r) * in_f;
    const f2 c0p = f2_bcast(__ldg(cp)), c1p = f2_bcast(in_f > 1 ? __ldg(cp + 1) : 0.f), c2p = f2_bcast(in_f > 2 ? __ldg(cp + 2) : 0.f);
    // two feature pairs (f = 0,1 and 2,3), packed FP32 math (f32x2.cuh)
    uint32_t pp[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 ta = tab[(2 * h) * nq + q], tb = tab[(2 * h + 1) * nq + q];
      const f2 z = f2_fma(c0p, f2_make(ta.x, tb.x), f2_fma(c1p, f2_make(ta.y, tb.y), f2_fma(c2p, f2_make(ta.z, tb.z), f2_make(ta.w, tb.w))));
      f2 wn = 0ull;
      if constexpr (TWO_D) {
        const float4 ua = tab2[(2 * h) * nq + q], ub = tab2[(2 * h + 1) * nq + q];
        const f2 w = f2_fma(c0p, f2_make(ua.x, ub.x), f2_fma(c1p, f2_make(ua.y, ub.y), f2_fma(c2p, f2_make(ua.z, ub.z), f2_make(ua.w, ub.w))));
        wn = f2_mul(w, w);
      }
      f2 yr, yi;
      gabor_real_x2(G2, z, wn, yr, yi);
      pp[2 * h] = pack_f16(f2_lo(yr), f2_lo(yi));
      pp[2 * h + 1] = pack_f16(f2_hi(yr), f2_hi(yi));
    }
    __half* dst = y + size_t(row0 + r) * y_pitch + 8 * q;
    if (4 * q + 3 < M) {
      __stcs(reinterpret_cast<uint4*>(dst), make_uint4(pp[0], pp[1], pp[2], pp[3]));
    } else {  // ragged last quad: only the valid features (the ones column follows them)
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
      for (int f = 0; f < 4; ++f)
        if (4 * q + f < M) d32[f] = pp[f];
    }
    r += dr; q += dq;
    if (q >= nq) { q -= nq; ++r; }
  }
}

// ---------------------------------------------------------------------------------------------
// top of the backward pass (training path, 16-bit tensors):
//   g_h = g_o conj(Wf) ; h = gabor(z_H [, w_H]) recomputed ; g_z = gabor'(z, g_h) [, g_w] ; g_Wf += g_o^T conj(h) ; g_bf += sum g_o
// Persistent CTAs stream tiles of kTopRows coordinates: the FP16 z (w) tile arrives by TMA (full row pitch, boxes of <= 256
// columns, double buffered), one thread owns one complex feature for the CTA's whole row range (so g_Wf stays in
// registers), results are written to a BF16 smem tile that leaves by TMA (clipped to the 2M valid columns).
// ---------------------------------------------------------------------------------------------
constexpr int kTopRows = 8;
struct TopBwd16Params {
  CUtensorMap z_map[2];   // z, w: [n][pitch] FP16, box {bw cols, kTopRows}, no swizzle
  CUtensorMap g_map[2];   // g_z, g_w: [n][2M] BF16 (row pitch = pitch), same box
  const float* g_out;     // [n][out_f]
  const float* Wf;        // [out_f][M] complex
  const float* omega;
  const float* scale;
  float* g_Wf;            // accumulated
  float* g_bf;
  int n, M, out_f, pitch, two_d;
  int bw, n_box;          // a row of `pitch` columns is moved as n_box boxes of bw columns (bw <= 256, bw % 8 == 0)
};

// Block = W compute warps (thread k owns complex feature k) + one I/O warp.  No block-wide barrier in the loop: tiles
// flow through two-deep in / out rings guarded by mbarriers (in_full: TMA bytes + g_out rows staged by the I/O warp;
// in_empty / out_full: one arrival per compute warp; out_empty: the I/O warp, once the TMA store has read the tile).
// MAXT: launch bound (512 leaves 128 registers per thread for the usual widths; 1024 covers M up to 992)
template <bool TWO_D, int OUTF, int MAXT>
__global__ void __launch_bounds__(MAXT) top_bwd16_kernel(const __grid_constant__ TopBwd16Params P) {
  using namespace sm100;
  extern __shared__ __align__(1024) uint8_t tsm[];
  __shared__ __align__(8) uint64_t in_full[2], in_empty[2], out_full[2], out_empty[2];
  __shared__ __align__(16) float s_go[2][kTopRows][4];
  const uint32_t pad = ((smem_u32(tsm) + 127u) & ~127u) - smem_u32(tsm);
  uint8_t* sm = tsm + pad;
  const uint32_t base = smem_u32(sm);
  const uint32_t tile_bytes = uint32_t(P.pitch) * kTopRows * 2;   // one tensor, one stage
  constexpr int n_t = TWO_D ? 2 : 1;
  // layout: in[2 stages][n_t tensors] | out[2 buffers][n_t tensors]
  const uint32_t out_off = 2 * n_t * tile_bytes;
  const uint32_t box_bytes = uint32_t(P.bw) * kTopRows * 2;

  const int n_tiles = (P.n + kTopRows - 1) / kTopRows;
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per;
  int t_end = t_begin + per;
  t_end = t_end > n_tiles ? n_tiles : t_end;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cw = (blockDim.x >> 5) - 1;  // compute warps

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&in_full[s]), 2);
      mbar_init(smem_u32(&in_empty[s]), n_cw);
      mbar_init(smem_u32(&out_full[s]), n_cw);
      mbar_init(smem_u32(&out_empty[s]), 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == n_cw) {
    // ===================== I/O warp =====================
    auto issue_load = [&](int tile, int stage) {
      if (lane == 0) {
        const uint32_t bar = smem_u32(&in_full[stage]);
        mbar_expect_tx(bar, n_t * tile_bytes);
        for (int t = 0; t < n_t; ++t) {
          const uint32_t dst = base + (stage * n_t + t) * tile_bytes;
          for (int b = 0; b < P.n_box; ++b) tma_load_2d_hint(dst + b * box_bytes, &P.z_map[t], bar, b * P.bw, tile * kTopRows, kEvictFirst);
        }
      }
      {  // this tile's g_out rows (zero beyond n)
        const int rr = lane >> 2, o = lane & 3;
        const int row = tile * kTopRows + rr;
        s_go[stage][rr][o] = (rr < kTopRows && row < P.n && o < OUTF) ? __ldg(P.g_out + size_t(row) * OUTF + o) : 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&in_full[stage]));
    };
    if (t_begin < t_end) issue_load(t_begin, 0);
    if (t_begin + 1 < t_end) issue_load(t_begin + 1, 1);
    uint32_t ph[2] = {0, 0};
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int stage = (tile - t_begin) & 1;
      if (tile + 2 < t_end) {  // refill the in-stage as soon as every compute warp has read it
        if (lane == 0) mbar_wait(smem_u32(&in_empty[stage]), ph[stage]);
        __syncwarp();
        issue_load(tile + 2, stage);
      }
      if (lane == 0) {
        mbar_wait(smem_u32(&out_full[stage]), ph[stage]);
        for (int t = 0; t < n_t; ++t) {
          const uint32_t src = base + out_off + (stage * n_t + t) * tile_bytes;
          for (int b = 0; b < P.n_box; ++b)
            if (b * P.bw < 2 * P.M) tma_store_2d(&P.g_map[t], src + b * box_bytes, b * P.bw, tile * kTopRows);
        }
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(smem_u32(&out_empty[stage]));
      }
      ph[stage] ^= 1;
    }
    if (lane == 0) tma_store_wait_all<0>();
    return;
  }

  // ===================== compute warps =====================
  const int k = threadIdx.x;  // complex feature owned by this thread
  const bool active = k < P.M;
  const GaborConst G = make_gabor_const(__ldg(P.omega), __ldg(P.scale));
  float wr[OUTF], wi[OUTF], ar[OUTF], ai[OUTF];
#pragma unroll
  for (int o = 0; o < OUTF; ++o) {
    wr[o] = active ? __ldg(P.Wf + (size_t(o) * P.M + k) * 2) : 0.f;
    wi[o] = active ? __ldg(P.Wf + (size_t(o) * P.M + k) * 2 + 1) : 0.f;
    ar[o] = ai[o] = 0.f;
  }
  float bsum = 0.f;
  // byte offset of this thread's (re, im) pair inside a tile: box b holds columns [b*bw, (b+1)*bw) as a dense [rows][bw] block
  const int col = active ? 2 * k : 0;
  const int bi = col / P.bw;
  const uint32_t row_bytes = uint32_t(P.bw) * 2;
  const uint32_t off0 = uint32_t(bi) * box_bytes + uint32_t(col - bi * P.bw) * 2;

  uint32_t ph[2] = {0, 0};
  for (int tile = t_begin; tile < t_end; ++tile) {
    const int stage = (tile - t_begin) & 1;
    mbar_wait(smem_u32(&in_full[stage]), ph[stage]);
    mbar_wait(smem_u32(&out_empty[stage]), ph[stage] ^ 1);  // passes at once the first time a buffer is used
    ph[stage] ^= 1;
    if (threadIdx.x < OUTF) {
#pragma unroll
      for (int rr = 0; rr < kTopRows; ++rr) bsum += s_go[stage][rr][threadIdx.x];
    }
    const uint8_t* zin = sm + (stage * n_t) * tile_bytes + off0;
    uint8_t* zout = sm + out_off + (stage * n_t) * tile_bytes + off0;
#pragma unroll 4
    for (int rr = 0; rr < kTopRows; ++rr) {
      const uint32_t zp = *reinterpret_cast<const uint32_t*>(zin + rr * row_bytes);
      uint32_t wp = 0;
      if constexpr (TWO_D) wp = *reinterpret_cast<const uint32_t*>(zin + tile_bytes + rr * row_bytes);
      const float4 go = *reinterpret_cast<const float4*>(&s_go[stage][rr][0]);
      const float g4[4] = {go.x, go.y, go.z, go.w};
      float gyr = 0.f, gyi = 0.f;
#pragma unroll
      for (int o = 0; o < OUTF; ++o) { gyr = fmaf(g4[o], wr[o], gyr); gyi = fmaf(-g4[o], wi[o], gyi); }
      const float2 z = unpack_f16(zp);
      float2 w = make_float2(0.f, 0.f);
      if constexpr (TWO_D) w = unpack_f16(wp);
      const float wnorm = TWO_D ? fmaf(w.x, w.x, w.y * w.y) : 0.f;
      float yr, yi, gzr, gzi;
      gabor16(G, z.x, z.y, wnorm, yr, yi);
      const float pr = gabor_bwd(yr, yi, z.x, z.y, gyr, gyi, G.omega, G.s2, gzr, gzi);
      // (lanes past the last feature alias feature 0's slot: they compute on it but must not store)
      if (active) *reinterpret_cast<uint32_t*>(zout + rr * row_bytes) = pack_bf16(gzr, gzi);
      if constexpr (TWO_D) {
        const float t = -2.0f * G.s2 * pr;
        if (active) *reinterpret_cast<uint32_t*>(zout + tile_bytes + rr * row_bytes) = pack_bf16(t * w.x, t * w.y);
      }
#pragma unroll
      for (int o = 0; o < OUTF; ++o) { ar[o] = fmaf(g4[o], yr, ar[o]); ai[o] = fmaf(-g4[o], yi, ai[o]); }
    }
    fence_proxy_async_smem();  // out tile written through the generic proxy -> visible to the TMA store
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(smem_u32(&in_empty[stage]));
      mbar_arrive(smem_u32(&out_full[stage]));
    }
  }
  if (active) {
#pragma unroll
    for (int o = 0; o < OUTF; ++o) {
      atomicAdd(P.g_Wf + (size_t(o) * P.M + k) * 2, ar[o]);
      atomicAdd(P.g_Wf + (size_t(o) * P.M + k) * 2 + 1, ai[o]);
    }
  }
  if (threadIdx.x < OUTF) atomicAdd(P.g_bf + 2 * threadIdx.x, bsum);
}

// ---------------------------------------------------------------------------------------------
// first layer weight gradient from BF16 g_z0 [n][g_pitch] (g_pitch a multiple of 8 elements):
//   g_W0[j][d] = sum_n g_z0[n,j] c[n,d] ; g_b0[j] = sum_n g_z0[n,j]
// lane = octet of 8 features (one 16-byte load), warp = row group, 4 rows in flight per thread.  One or two 512-thread
// blocks per SM: every output address then sees only ~2 x SM-count global atomics (same-address atomics serialise in
// L2; with 4 blocks per SM they, not the loads, set this kernel's time).  Requires M <= 256, in_f <= 3.
// ---------------------------------------------------------------------------------------------
constexpr int kFirstWgrad16Threads = 512;
__global__ void __launch_bounds__(kFirstWgrad16Threads) first_wgrad16_kernel(const __nv_bfloat16* __restrict__ gz0, int g_pitch,
                                                                              const float* __restrict__ coords, int n, int in_f, int M,
                                                                              float* __restrict__ gW0, float* __restrict__ gb0,
                                                                              int rows_per_block) {
  __shared__ float red[32][33];  // [lane][8 features x 4]
  constexpr int kWarps = kFirstWgrad16Threads / 32;
  const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const int range0 = blockIdx.x * rows_per_block;
  int range1 = range0 + rows_per_block;
  range1 = range1 > n ? n : range1;
  const bool col_ok = 8 * lane < g_pitch;
  float acc[8][4];
#pragma unroll
  for (int f = 0; f < 8; ++f)
#pragma unroll
    for (int d = 0; d < 4; ++d) acc[f][d] = 0.f;
  for (int rb = range0 + wg; rb < range1; rb += 4 * kWarps) {
    uint4 g[4];
    float c[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = rb + kWarps * u;
      const bool ok = row < range1 && col_ok;
      g[u] = ok ? __ldcs(reinterpret_cast<const uint4*>(gz0 + size_t(row) * g_pitch + 8 * lane)) : make_uint4(0u, 0u, 0u, 0u);
      const int rc = row < range1 ? row : range0;
      c[u][0] = __ldg(coords + size_t(rc) * in_f);
      c[u][1] = in_f > 1 ? __ldg(coords + size_t(rc) * in_f + 1) : 0.f;
      c[u][2] = in_f > 2 ? __ldg(coords + size_t(rc) * in_f + 2) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t pk[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[h]));
        acc[2 * h][0] = fmaf(v.x, c[u][0], acc[2 * h][0]);
        acc[2 * h][1] = fmaf(v.x, c[u][1], acc[2 * h][1]);
        acc[2 * h][2] = fmaf(v.x, c[u][2], acc[2 * h][2]);
        acc[2 * h][3] += v.x;
        acc[2 * h + 1][0] = fmaf(v.y, c[u][0], acc[2 * h + 1][0]);
        acc[2 * h + 1][1] = fmaf(v.y, c[u][1], acc[2 * h + 1][1]);
        acc[2 * h + 1][2] = fmaf(v.y, c[u][2], acc[2 * h + 1][2]);
        acc[2 * h + 1][3] += v.y;
      }
    }
  }
#pragma unroll
  for (int f = 0; f < 8; ++f)
#pragma unroll
    for (int d = 0; d < 4; ++d) atomicAdd(&red[lane][4 * f + d], acc[f][d]);
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int l = i >> 5, v = i & 31, f = v >> 2, d = v & 3;
    const int j = 8 * l + f;
    if (j < M) {
      if (d < 3) { if (d < in_f) atomicAdd(gW0 + size_t(j) * in_f + d, red[l][v]); }
      else atomicAdd(gb0 + j, red[l][v]);
    }
  }
}

// g_c[n][d] (+)= sum_j g_z0[n,j] W0[j,d] from BF16 g_z0 (gradient w.r.t. the coordinates; one warp per row)
__global__ void __launch_bounds__(256) grad_coords16_kernel(const __nv_bfloat16* __restrict__ gz0, int g_pitch, int n, int in_f, int M,
                                                             const float* __restrict__ W0, float* __restrict__ gc, int accumulate) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < n; row += nwarps) {
    float acc[3] = {0.f, 0.f, 0.f};
    for (int j = lane; j < M; j += 32) {
      const float g = __bfloat162float(gz0[size_t(row) * g_pitch + j]);
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (d < in_f) acc[d] = fmaf(g, W0[size_t(j) * in_f + d], acc[d]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (d < in_f) {
        float v = acc[d];
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) {
          float* dst = gc + size_t(row) * in_f + d;
          *dst = accumulate ? *dst + v : v;
        }
      }
  }
}

}  // namespace wire
