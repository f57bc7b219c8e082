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
// first layer forward.  y: FP16 [n][y_pitch]; columns >= 2M (the "ones" column) are not touched.
// A thread OWNS one quad of 4 complex features (its {w0[0..2], b0} rows live in registers as packed pairs) and walks the
// rows of its block's range: slot s = threadIdx / nq takes rows s, s + slots, ...; the lanes of a warp are consecutive quads
// of (mostly) one row, so the 16-byte stores of a warp are contiguous.  The block's coordinates are staged once in shared
// memory as float4.  Per item that leaves one broadcast LDS.128, the packed math, 12 MUFU and one STG.128.
// (The previous version dealt (row, quad) items round-robin and re-read the quad's table from shared memory for every item:
// 107 instructions per item, 12 of them register moves pairing the halves of two LDS.128, 3 dependent global loads for the
// coordinates; ncu r02: issue slots 66 % busy, XU 60 %, 73 us.  The XU floor of this stage is 37 us.)
// ---------------------------------------------------------------------------------------------
constexpr int kFirstFwd16Threads = 256;
constexpr int kFirstFwd16MaxRows = 512;   // rows per block (coordinates staged in 8 KB of shared memory)
template <bool TWO_D>
__global__ void __launch_bounds__(kFirstFwd16Threads) first_fwd16_kernel(const float* __restrict__ coords, int n, int in_f, int M,
                                                           const float* __restrict__ W0, const float* __restrict__ b0,
                                                           const float* __restrict__ W0b, const float* __restrict__ b0b,
                                                           const float* __restrict__ omega_p, const float* __restrict__ scale_p,
                                                           __half* __restrict__ y, int y_pitch, int rows_per_block, int sector_tail) {
  __shared__ __align__(16) float4 cs[kFirstFwd16MaxRows];
  const int nq = (M + 3) >> 2;
  const int row0 = blockIdx.x * rows_per_block;
  int rows = n - row0;
  rows = rows > rows_per_block ? rows_per_block : rows;
  sm100::pdl_trigger();
  if (rows <= 0) return;
  // quads per pass and row slots: nq <= blockDim (every driver): one pass covers all quads, blockDim / nq rows at a time
  const int qpp = nq < int(blockDim.x) ? nq : int(blockDim.x);
  const int slots = int(blockDim.x) / qpp;
  const int qi = int(threadIdx.x) % qpp, slot = int(threadIdx.x) / qpp;
  const GaborConst2 G2 = make_gabor_const2(make_gabor_const(__ldg(omega_p), __ldg(scale_p)));
  bool waited = false;
  for (int qb = 0; qb < nq; qb += qpp) {   // (block-uniform trip count: the loop body synchronises)
    const int q = qb + qi;
    const bool active = q < nq && slot < slots;   // threads beyond slots * qpp idle (blockDim is not a multiple of nq)
    // this quad's table: pair h = features (4q + 2h, 4q + 2h + 1); parameters only, so before the dependency wait
    f2 wx[2], wy[2], wz[2], wb[2], vx[2], vy[2], vz[2], vb[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float t[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, u[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 4 * q + 2 * h + e;
        if (k < M) {
          t[e][0] = __ldg(W0 + size_t(k) * in_f);
          if (in_f > 1) t[e][1] = __ldg(W0 + size_t(k) * in_f + 1);
          if (in_f > 2) t[e][2] = __ldg(W0 + size_t(k) * in_f + 2);
          t[e][3] = __ldg(b0 + k);
          if constexpr (TWO_D) {
            u[e][0] = __ldg(W0b + size_t(k) * in_f);
            if (in_f > 1) u[e][1] = __ldg(W0b + size_t(k) * in_f + 1);
            if (in_f > 2) u[e][2] = __ldg(W0b + size_t(k) * in_f + 2);
            u[e][3] = __ldg(b0b + k);
          }
        }
      }
      wx[h] = f2_make(t[0][0], t[1][0]); wy[h] = f2_make(t[0][1], t[1][1]); wz[h] = f2_make(t[0][2], t[1][2]); wb[h] = f2_make(t[0][3], t[1][3]);
      vx[h] = f2_make(u[0][0], u[1][0]); vy[h] = f2_make(u[0][1], u[1][1]); vz[h] = f2_make(u[0][2], u[1][2]); vb[h] = f2_make(u[0][3], u[1][3]);
    }
    if (!waited) { sm100::pdl_wait(); waited = true; }   // coords and the y buffer (still read by last step's kernels) below
    const bool full = 4 * q + 3 < M;
    for (int c0r = 0; c0r < rows; c0r += kFirstFwd16MaxRows) {   // the block's rows, kFirstFwd16MaxRows staged at a time
      const int crow = rows - c0r < kFirstFwd16MaxRows ? rows - c0r : kFirstFwd16MaxRows;
      __syncthreads();   // everyone is done with the previous staging
      for (int r = threadIdx.x; r < crow; r += blockDim.x) {
        const float* cp = coords + size_t(row0 + c0r + r) * in_f;
        cs[r] = make_float4(__ldg(cp), in_f > 1 ? __ldg(cp + 1) : 0.f, in_f > 2 ? __ldg(cp + 2) : 0.f, 0.f);
      }
      __syncthreads();
      if (!active) continue;
      __half* dst = y + size_t(row0 + c0r + slot) * y_pitch + 8 * q;
      const size_t dstep = size_t(slots) * y_pitch;
      // A row of 2M columns that ends inside a 32-byte sector makes that sector a read-modify-write in DRAM for every row (see
      // run_rows_job in api.cu): the thread of the row's last quad completes it with column 2M = 1.0 (what that column holds
      // anyway: the wgrad's "ones" column) and zeros.  tail_words: 32-bit words from column 2M to the sector boundary.
      const int tail_cols = ((2 * M + 15) & ~15) - 2 * M;
      const int tail_words = (sector_tail && q == nq - 1 && 2 * M + tail_cols <= y_pitch) ? tail_cols / 2 : 0;
      const bool tail_vec = tail_words == 4 && (M & 3) == 0;
      auto item = [&](const float4 c, uint32_t (&pp)[4]) {
        const f2 c0 = f2_bcast(c.x), c1 = f2_bcast(c.y), c2 = f2_bcast(c.z);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const f2 z = f2_fma(c0, wx[h], f2_fma(c1, wy[h], f2_fma(c2, wz[h], wb[h])));
          f2 wn = 0ull;
          if constexpr (TWO_D) {
            const f2 w = f2_fma(c0, vx[h], f2_fma(c1, vy[h], f2_fma(c2, vz[h], vb[h])));
            wn = f2_mul(w, w);
          }
          f2 yr, yi;
          gabor_real_x2(G2, z, wn, yr, yi);
          pp[2 * h] = pack_f16(f2_lo(yr), f2_lo(yi));
          pp[2 * h + 1] = pack_f16(f2_hi(yr), f2_hi(yi));
        }
      };
      auto store = [&](__half* d, const uint32_t (&pp)[4]) {
        if (full) {
          __stcs(reinterpret_cast<uint4*>(d), make_uint4(pp[0], pp[1], pp[2], pp[3]));
        } else {  // ragged last quad: only the valid features (the ones column follows them)
          uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
#pragma unroll
          for (int f = 0; f < 4; ++f)
            if (4 * q + f < M) d32[f] = pp[f];
        }
      };
      // (two rows per iteration -- 24 MUFU in flight per thread -- measured equal: 85.8 vs 81.8 us on a 3 % slower box)
      int r = slot;
      for (; r < crow; r += slots, dst += dstep) {
        uint32_t pp[4];
        item(cs[r], pp);
        store(dst, pp);
        if (tail_words > 0) {   // last quad of the row: the "ones" column and zeros up to the end of the 32-byte sector
          uint32_t* t32 = reinterpret_cast<uint32_t*>(dst + (2 * M - 8 * q));
          if (tail_vec) {
            __stcs(reinterpret_cast<uint4*>(t32), make_uint4(0x00003C00u, 0u, 0u, 0u));
          } else {
            t32[0] = 0x00003C00u;
            for (int i = 1; i < tail_words; ++i) t32[i] = 0u;
          }
        }
      }
    }
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
constexpr int kTopOut = 4; // maximum depth of the output ring
constexpr int kTopIn = 4;  // maximum depth of the input ring (ncu: with 2 the compute warps waited on TMA latency, long_scoreboard 3.0)
// The depths actually used are launch parameters (TopBwd16Params::in_depth / out_depth <= the maxima): wire2d doubles the bytes
// per stage, and at the SISR width (M = 128) a CTA is only four compute warps, so shallower rings that let more CTAs share an
// SM beat deep ones.
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
  // Fused MSE (wire_net_backward_mse): g_out is not read but computed by the I/O warp as 2 (pred - target) / count_norm
  // (wire_image_denoise.py:153 / criterion of wire_occupancy.py:149), and the loss goes to the device ring (see mse_grad_kernel).
  const float* pred;      // [n][out_f] or nullptr
  const float* target;
  float g_scale, loss_scale;   // 2 / count_norm, 1 / count_norm
  float* ring;
  int ring_n;
  const long long* step_ptr;
  // trainable omega_0 / scale_0 of the last hidden layer (SCAL instantiations): accumulated device floats, or nullptr
  float* gs_omega;
  float* gs_scale;
  int in_depth, out_depth;  // ring depths (<= kTopIn / kTopOut)
  int store_cols;           // columns of g_z / g_w that leave (2M rounded up to whole 32-byte sectors when the pitch allows)
};

// Block = W compute warps (thread k owns complex feature k) + one I/O warp.  No block-wide barrier in the loop: tiles
// flow through two-deep in / out rings guarded by mbarriers (in_full: TMA bytes + g_out rows staged by the I/O warp;
// in_empty / out_full: one arrival per compute warp; out_empty: the I/O warp, once the TMA store has read the tile).
// MAXT: launch bound (512 leaves 128 registers per thread for the usual widths; 1024 covers M up to 992)
template <bool TWO_D, int OUTF, int MAXT, bool SCAL = false>
__global__ void __launch_bounds__(MAXT) top_bwd16_kernel(const __grid_constant__ TopBwd16Params P) {
  using namespace sm100;
  extern __shared__ __align__(1024) uint8_t tsm[];
  __shared__ __align__(8) uint64_t in_full[kTopIn], in_empty[kTopIn], out_full[kTopOut], out_empty[kTopOut];
  __shared__ __align__(16) float s_go[kTopIn][kTopRows / 2][4][2];  // [stage][row pair][output][row of the pair]
  const uint32_t pad = ((smem_u32(tsm) + 127u) & ~127u) - smem_u32(tsm);
  uint8_t* sm = tsm + pad;
  const uint32_t base = smem_u32(sm);
  const uint32_t tile_bytes = uint32_t(P.pitch) * kTopRows * 2;   // one tensor, one stage
  constexpr int n_t = TWO_D ? 2 : 1;
  // layout: in[kTopIn stages][n_t tensors] | out[kTopOut buffers][n_t tensors]
  const int kIn = P.in_depth, kOut = P.out_depth;
  const uint32_t out_off = kIn * n_t * tile_bytes;
  const uint32_t box_bytes = uint32_t(P.bw) * kTopRows * 2;

  const int n_tiles = (P.n + kTopRows - 1) / kTopRows;
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per;
  int t_end = t_begin + per;
  t_end = t_end > n_tiles ? n_tiles : t_end;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cw = (blockDim.x >> 5) - 1;  // compute warps

  if (threadIdx.x == 0) {
    for (int s = 0; s < kIn; ++s) {
      mbar_init(smem_u32(&in_full[s]), 2);
      mbar_init(smem_u32(&in_empty[s]), n_cw);
    }
    for (int s = 0; s < kOut; ++s) {
      mbar_init(smem_u32(&out_full[s]), n_cw);
      mbar_init(smem_u32(&out_empty[s]), 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();

  if (warp == n_cw) {
    // ===================== I/O warp =====================
    // One latency chain per tile would make this warp the bottleneck (measured: 4 CTAs/SM x one tile per ~2 us), so nothing
    // here waits for a round trip: the g_out rows of the tile loaded NEXT iteration are prefetched into a register now,
    // and a TMA store is only waited for (wait_read<1>) one iteration after it was issued.
    const int rr = lane >> 2, o = lane & 3;
    float loss_acc = 0.f;
    long long ring_step = 0;
    if (P.pred) {
      ring_step = *P.step_ptr;
      if (blockIdx.x == 0 && lane == 0) P.ring[(ring_step + 1) % P.ring_n] = 0.f;   // slot of the next step
    }
    auto load_go = [&](int tile) -> float {
      const int row = tile * kTopRows + rr;
      if (!(tile < t_end && row < P.n && o < OUTF)) return 0.f;
      if (P.pred) {
        const float dlt = __ldg(P.pred + size_t(row) * OUTF + o) - __ldg(P.target + size_t(row) * OUTF + o);
        loss_acc = fmaf(dlt, dlt, loss_acc);
        return dlt * P.g_scale;
      }
      return __ldg(P.g_out + size_t(row) * OUTF + o);
    };
    auto issue_load = [&](int tile, int stage, float go) {
      if (lane == 0) {
        const uint32_t bar = smem_u32(&in_full[stage]);
        mbar_expect_tx(bar, n_t * tile_bytes);
        for (int t = 0; t < n_t; ++t) {
          const uint32_t dst = base + (stage * n_t + t) * tile_bytes;
          for (int b = 0; b < P.n_box; ++b) tma_load_2d_hint(dst + b * box_bytes, &P.z_map[t], bar, b * P.bw, tile * kTopRows, kEvictFirst);
        }
      }
      s_go[stage][rr >> 1][o][rr & 1] = go;  // this tile's g_out rows (zero beyond n)
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&in_full[stage]));
    };
    for (int s = 0; s < kIn; ++s)
      if (t_begin + s < t_end) issue_load(t_begin + s, s, load_go(t_begin + s));
    float go_next = load_go(t_begin + kIn);
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int it = tile - t_begin;
      const int istage = it % kIn, ostage = it % kOut;
      const uint32_t iph = (it / kIn) & 1, oph = (it / kOut) & 1;
      if (tile + kIn < t_end) {  // refill the in-stage as soon as every compute warp has read it
        if (lane == 0) mbar_wait(smem_u32(&in_empty[istage]), iph);
        __syncwarp();
        issue_load(tile + kIn, istage, go_next);
        go_next = load_go(tile + 1 + kIn);
      }
      if (lane == 0) {
        mbar_wait(smem_u32(&out_full[ostage]), oph);
        for (int t = 0; t < n_t; ++t) {
          const uint32_t src = base + out_off + (ostage * n_t + t) * tile_bytes;
          for (int b = 0; b < P.n_box; ++b)
            if (b * P.bw < 2 * P.M) tma_store_2d(&P.g_map[t], src + b * box_bytes, b * P.bw, tile * kTopRows);
        }
        tma_store_commit();
        tma_store_wait_read<1>();  // the PREVIOUS tile's store has read its buffer
        if (it >= 1) mbar_arrive(smem_u32(&out_empty[(it - 1) % kOut]));
      }
    }
    if (P.pred) {
      for (int sft = 16; sft > 0; sft >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, sft);
      if (lane == 0 && t_begin < t_end) atomicAdd(P.ring + (ring_step % P.ring_n), loss_acc * P.loss_scale);
    }
    if (lane == 0) tma_store_wait_all<0>();
    return;
  }

  // ===================== compute warps =====================
  // Thread k owns complex feature k and processes TWO ROWS at a time with the packed FP32 instructions (f32x2.cuh):
  // every quantity below is a {row 2j, row 2j+1} pair, the feature's weights are broadcast pairs.
  const int k = threadIdx.x;  // complex feature owned by this thread
  const bool active = k < P.M;
  const GaborConst2 G2 = make_gabor_const2(make_gabor_const(__ldg(P.omega), __ldg(P.scale)));
  f2 wr2[OUTF], nwi2[OUTF], ar2[OUTF], ai2[OUTF];
#pragma unroll
  for (int o = 0; o < OUTF; ++o) {
    wr2[o] = f2_bcast(active ? __ldg(P.Wf + (size_t(o) * P.M + k) * 2) : 0.f);
    nwi2[o] = f2_bcast(active ? -__ldg(P.Wf + (size_t(o) * P.M + k) * 2 + 1) : 0.f);
    ar2[o] = ai2[o] = 0ull;
  }
  float bsum = 0.f;
  f2 s_om = 0ull, s_sc = 0ull;  // SCAL: sum Im(conj(z) p), sum (|z|^2 + |w|^2) Re p over this thread's feature
  // byte offset of this thread's (re, im) pair inside a tile: box b holds columns [b*bw, (b+1)*bw) as a dense [rows][bw] block
  // threads of the zero-padded features up to the stored width write their (zero) results; the rest alias feature 0's slot and never store
  const bool st_ok = 2 * k < P.store_cols;
  const int col = st_ok ? 2 * k : 0;
  const int bi = col / P.bw;
  const uint32_t row_bytes = uint32_t(P.bw) * 2;
  const uint32_t off0 = uint32_t(bi) * box_bytes + uint32_t(col - bi * P.bw) * 2;

  // ring positions and parities are carried along instead of being re-derived from the tile index (it % 3, it / 3 ... were a
  // fifth of this loop's instructions; ncu r02: the kernel issues 460 instructions per tile and warp, 220 of them math)
  int istage = 0, stage = 0;
  uint32_t iphase = 0, ophase = 1;
  const uint32_t in_stride = n_t * tile_bytes;
  const uint8_t* zin = sm + off0;
  uint8_t* zout = sm + out_off + off0;
  for (int tile = t_begin; tile < t_end; ++tile) {
    mbar_wait(smem_u32(&in_full[istage]), iphase);
    mbar_wait(smem_u32(&out_empty[stage]), ophase);  // passes at once the first time a buffer is used
    if (threadIdx.x < OUTF) {
#pragma unroll
      for (int rp = 0; rp < kTopRows / 2; ++rp) bsum += s_go[istage][rp][threadIdx.x][0] + s_go[istage][rp][threadIdx.x][1];
    }
#pragma unroll
    for (int rp = 0; rp < kTopRows / 2; ++rp) {
      const uint32_t za = *reinterpret_cast<const uint32_t*>(zin + (2 * rp) * row_bytes);
      const uint32_t zb = *reinterpret_cast<const uint32_t*>(zin + (2 * rp + 1) * row_bytes);
      // {re row a, re row b} and {im row a, im row b} as half2, then one conversion each to an aligned float pair
      const float2 zre = unpack_f16(__byte_perm(za, zb, 0x5410)), zim = unpack_f16(__byte_perm(za, zb, 0x7632));
      const f2 zr = f2_make(zre.x, zre.y), zi = f2_make(zim.x, zim.y);
      f2 wre = 0ull, wim = 0ull, wnorm = 0ull;
      if constexpr (TWO_D) {
        const uint32_t wa = *reinterpret_cast<const uint32_t*>(zin + tile_bytes + (2 * rp) * row_bytes);
        const uint32_t wb = *reinterpret_cast<const uint32_t*>(zin + tile_bytes + (2 * rp + 1) * row_bytes);
        const float2 w_re = unpack_f16(__byte_perm(wa, wb, 0x5410)), w_im = unpack_f16(__byte_perm(wa, wb, 0x7632));
        wre = f2_make(w_re.x, w_re.y); wim = f2_make(w_im.x, w_im.y);
        wnorm = f2_fma(wre, wre, f2_mul(wim, wim));
      }
      const f2* go2 = reinterpret_cast<const f2*>(&s_go[istage][rp][0][0]);  // [o] -> {g_o row a, g_o row b}
      f2 g2[OUTF];
      f2 gyr = 0ull, gyi = 0ull;
#pragma unroll
      for (int o = 0; o < OUTF; ++o) {
        g2[o] = go2[o];
        gyr = f2_fma(g2[o], wr2[o], gyr);
        gyi = f2_fma(g2[o], nwi2[o], gyi);
      }
      f2 yr, yi, gzr, gzi;
      gabor_x2(G2, zr, zi, wnorm, yr, yi);
      f2 pr;
      if constexpr (SCAL) {
        f2 pi;
        pr = gabor_bwd_x2_p(G2, yr, yi, zr, zi, gyr, gyi, gzr, gzi, pi);
        s_om = f2_fma(zr, pi, f2_fma(f2_mul(zi, G2.none), pr, s_om));
        s_sc = f2_fma(f2_fma(zi, zi, f2_fma(zr, zr, wnorm)), pr, s_sc);
      } else {
        pr = gabor_bwd_x2(G2, yr, yi, zr, zi, gyr, gyi, gzr, gzi);
      }
      // (lanes past the stored width alias feature 0's slot: they compute on it but must not store)
      // (padded features up to the stored width: their weights are zero, so p = 0 and g_z = 0 * z; z there is the exact 0 the
      // forward kernel stored for the zero-padded feature -- the same stored-width rule on both sides, api.cu: run_rows_job)
      if (st_ok) {
        *reinterpret_cast<uint32_t*>(zout + (2 * rp) * row_bytes) = pack_bf16(f2_lo(gzr), f2_lo(gzi));
        *reinterpret_cast<uint32_t*>(zout + (2 * rp + 1) * row_bytes) = pack_bf16(f2_hi(gzr), f2_hi(gzi));
      }
      if constexpr (TWO_D) {
        const f2 t = f2_mul(G2.m2s2, pr);
        const f2 gwr = f2_mul(t, wre), gwi = f2_mul(t, wim);
        if (st_ok) {
          *reinterpret_cast<uint32_t*>(zout + tile_bytes + (2 * rp) * row_bytes) = pack_bf16(f2_lo(gwr), f2_lo(gwi));
          *reinterpret_cast<uint32_t*>(zout + tile_bytes + (2 * rp + 1) * row_bytes) = pack_bf16(f2_hi(gwr), f2_hi(gwi));
        }
      }
#pragma unroll
      for (int o = 0; o < OUTF; ++o) { ar2[o] = f2_fma(g2[o], yr, ar2[o]); ai2[o] = f2_fma(g2[o], yi, ai2[o]); }
    }
    fence_proxy_async_smem();  // out tile written through the generic proxy -> visible to the TMA store
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(smem_u32(&in_empty[istage]));
      mbar_arrive(smem_u32(&out_full[stage]));
    }
    zin += in_stride; zout += in_stride;
    if (++istage == kIn) { istage = 0; iphase ^= 1; zin = sm + off0; }
    if (++stage == kOut) { stage = 0; ophase ^= 1; zout = sm + out_off + off0; }
  }
  if (active) {
#pragma unroll
    for (int o = 0; o < OUTF; ++o) {  // g_Wf = g_o^T conj(h): even-row + odd-row partial sums; the imaginary part is negated
      atomicAdd(P.g_Wf + (size_t(o) * P.M + k) * 2, f2_lo(ar2[o]) + f2_hi(ar2[o]));
      atomicAdd(P.g_Wf + (size_t(o) * P.M + k) * 2 + 1, -(f2_lo(ai2[o]) + f2_hi(ai2[o])));
    }
  }
  if (threadIdx.x < OUTF) atomicAdd(P.g_bf + 2 * threadIdx.x, bsum);
  if constexpr (SCAL) {
    // (rows past n arrive as TMA zero fill with g_o = 0: p = 0; threads past the last feature are masked here)
    float a = active ? f2_lo(s_om) + f2_hi(s_om) : 0.f, b = active ? f2_lo(s_sc) + f2_hi(s_sc) : 0.f;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, sft); b += __shfl_xor_sync(0xffffffffu, b, sft); }
    if (lane == 0) {
      if (P.gs_omega) atomicAdd(P.gs_omega, a);
      if (P.gs_scale) atomicAdd(P.gs_scale, -2.0f * __ldg(P.scale) * b);
    }
  }
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
  sm100::pdl_trigger();
  sm100::pdl_wait();
  const int range0 = blockIdx.x * rows_per_block;
  int range1 = range0 + rows_per_block;
  range1 = range1 > n ? n : range1;
  // narrow rows (wire2d at the SISR width: 128 features = 16 lanes) are packed two or four to a warp, so every lane loads
  const int lpr = g_pitch <= 64 ? 8 : (g_pitch <= 128 ? 16 : 32);  // lanes per row
  const int rpw = 32 / lpr;                                         // rows per warp and load
  const int oct = lane % lpr, rsub = lane / lpr;
  const bool col_ok = 8 * oct < g_pitch;
  float acc[8][4];
#pragma unroll
  for (int f = 0; f < 8; ++f)
#pragma unroll
    for (int d = 0; d < 4; ++d) acc[f][d] = 0.f;
  for (int rb = range0 + wg * rpw; rb < range1; rb += 4 * kWarps * rpw) {
    uint4 g[4];
    float c[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = rb + kWarps * rpw * u + rsub;
      const bool ok = row < range1 && col_ok;
      g[u] = ok ? __ldcs(reinterpret_cast<const uint4*>(gz0 + size_t(row) * g_pitch + 8 * oct)) : make_uint4(0u, 0u, 0u, 0u);
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
    for (int d = 0; d < 4; ++d) atomicAdd(&red[oct][4 * f + d], acc[f][d]);
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

// ---------------------------------------------------------------------------------------------
// The same sums, streamed: first_wgrad16_kernel keeps its loads in registers (4 rows per thread, load -> use in the same
// iteration), so every block iteration pays one full memory round trip and the kernel ran at 2.2 TB/s (ncu r02:
// long_scoreboard 9.7 warps per issue cycle).  g_z0 rows are contiguous (row pitch = g_pitch), so here ONE producer lane moves
// 64-row chunks (27 KB at M = 212) with 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) into a ring of `stages`
// shared-memory buffers — up to ~190 KB in flight per SM whatever the compute warps do — and 16 compute warps sum them from
// shared memory with packed FFMA2.  One block per SM; chunks are dealt round-robin FROM THE LAST ROW TO THE FIRST: the
// first-layer dgrad kernel in front of this one wrote g_z0 (113 MB, less than the 126 MB L2) in ascending row order, so the
// rows it wrote last are still in L2 when this kernel starts there.
// Requires M <= 256, in_f <= 3, g_pitch % 8 == 0, 16-byte aligned gz0 / coords.
// ---------------------------------------------------------------------------------------------
// rows per chunk: 128 for full-width rows (lanes per row = 32), 256 for narrow rows (wire2d at the SISR width: 256-byte rows).  With 64
// the per-chunk bookkeeping (barrier wait, arrive, loop) was half of the instructions and the kernel, not the copies, set the time:
// 43 -> 37.5 us at 512^2, 199 -> 138 us for the two launches of the SISR step; the copy skeleton alone streams at 4.8-5.6 TB/s
// (tools/stream_probe, profiles/r02_probe_stream.log)
__host__ __device__ constexpr int fw16s_rows(int lpr) { return lpr == 32 ? 128 : 256; }
constexpr int kFw16sWarps = 16;                    // compute warps
constexpr int kFw16sThreads = 32 * (kFw16sWarps + 1);
__host__ __device__ inline uint32_t fw16s_stage_bytes(int g_pitch, int rows) {
  return (uint32_t(rows) * uint32_t(g_pitch) * 2u + uint32_t(rows) * 16u + 127u) & ~127u;  // g rows | coords (<= 4 floats a row)
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(sm100::kEvictFirst) : "memory");
}
template <int IN_F, int LPR>   // input features (1..3); lanes per g_z0 row (8 / 16 / 32: narrow rows are packed 4 / 2 to a warp)
__global__ void __launch_bounds__(kFw16sThreads, 1) first_wgrad16s_kernel(const __nv_bfloat16* __restrict__ gz0, int g_pitch,
                                                                           const float* __restrict__ coords, int n, int M,
                                                                           float* __restrict__ gW0, float* __restrict__ gb0, int stages,
                                                                           const __grid_constant__ CUtensorMap g_map, int use_tma) {
  // use_tma: the g_z0 chunk arrives as ONE 2-D tensor box {g_pitch columns, 64 rows} (rows past n are zero fill) instead of a 1-D
  // bulk copy: the 1-D copies streamed at 11.4 B/clk per SM whatever was in flight (ncu, cold L2: 3.2 TB/s at both row widths)
  constexpr int in_f = IN_F;
  using namespace sm100;
  extern __shared__ __align__(128) uint8_t fwsm[];
  __shared__ __align__(8) uint64_t bar_full[8];
  __shared__ __align__(8) uint64_t bar_empty[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t pad = ((smem_u32(fwsm) + 127u) & ~127u) - smem_u32(fwsm);
  uint8_t* sm = fwsm + pad;
  constexpr int kFw16sRows = fw16s_rows(LPR);
  const uint32_t stage_bytes = fw16s_stage_bytes(g_pitch, kFw16sRows);
  const uint32_t g_bytes_row = uint32_t(g_pitch) * 2u;
  const uint32_t c_off = uint32_t(kFw16sRows) * g_bytes_row;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), kFw16sWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  const int n_chunks = (n + kFw16sRows - 1) / kFw16sRows;
  // chunk of iteration i: the sweep runs from the last chunk to the first, all blocks side by side
  const int n_mine = (n_chunks - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  auto chunk_of = [&](int i) { return n_chunks - 1 - (i * int(gridDim.x) + int(blockIdx.x)); };

  if (warp == kFw16sWarps) {
    // ===================== producer warp =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < n_mine; ++i) {
      const int r0 = chunk_of(i) * kFw16sRows;
      const int rows = n - r0 < kFw16sRows ? n - r0 : kFw16sRows;
      if (lane == 0) mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
      __syncwarp();
      uint8_t* dst = sm + size_t(stage) * stage_bytes;
      const uint32_t cbytes = uint32_t(rows) * uint32_t(in_f) * 4u;
      const bool c_bulk = (cbytes & 15u) == 0;  // r0 is a multiple of 64 rows, so the source offset is 16-byte aligned
      if (!c_bulk) {  // ragged last chunk: plain copies, published by lane 0's arrive below
        float* cs = reinterpret_cast<float*>(dst + c_off);
        for (int k = lane; k < rows * in_f; k += 32) cs[k] = __ldg(coords + size_t(r0) * in_f + k);
        __syncwarp();
      }
      if (lane == 0) {
        const uint32_t bar = smem_u32(&bar_full[stage]);
        const uint32_t gbytes = uint32_t(rows) * g_bytes_row;
        if (use_tma) {
          mbar_expect_tx(bar, uint32_t(kFw16sRows) * g_bytes_row + (c_bulk ? cbytes : 0u));   // a box counts whole, zero fill included
          tma_load_2d_hint(smem_u32(dst), &g_map, bar, 0, r0, kEvictFirst);
        } else {
          mbar_expect_tx(bar, gbytes + (c_bulk ? cbytes : 0u));
          bulk_load_1d(smem_u32(dst), gz0 + size_t(r0) * g_pitch, gbytes, bar);
        }
        if (c_bulk) bulk_load_1d(smem_u32(dst + c_off), coords + size_t(r0) * in_f, cbytes, bar);
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    return;
  }

  // ===================== compute warps =====================
  // lane = octet of 8 features (one 16-byte shared load), narrow rows packed two or four to a warp; acc[h][d] is the
  // {feature 2h, feature 2h+1} pair of sums against (c0, c1, c2, 1)
  constexpr int lpr = LPR, rpw = 32 / LPR;
  const int oct = lane % lpr, rsub = lane / lpr;
  // lanes past the end of a row (g_pitch = 216: octets 27..31) read octet 0 instead and their sums are dropped by the j < M test of
  // the reduction: no predicate inside the loop.  (A first version guarded the accumulation with `if (col_ok)`: the compiler then
  // kept two copies of the 32 accumulators and the loop was 66 instructions per row, 24 of them moves; ncu: issue-bound, 22 M
  // warp instructions for 113 MB.)
  const uint32_t oct_off = 16u * uint32_t(8 * oct < g_pitch ? oct : 0);
  f2 acc[4][4];
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int d = 0; d < 4; ++d) acc[h][d] = 0ull;
  auto add_row = [&](const uint8_t* src, const float* cs, int r) {
    const uint4 g = *reinterpret_cast<const uint4*>(src + uint32_t(r) * g_bytes_row + oct_off);
    const f2 c0 = f2_bcast(cs[r * in_f]);
    const f2 c1 = f2_bcast(in_f > 1 ? cs[r * in_f + (in_f > 1 ? 1 : 0)] : 0.f);
    const f2 c2 = f2_bcast(in_f > 2 ? cs[r * in_f + (in_f > 2 ? 2 : 0)] : 0.f);
    const uint32_t pk[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const f2 v = f2_bits(pk[h] << 16, pk[h] & 0xffff0000u);  // BF16 pair -> {float, float}
      acc[h][0] = f2_fma(v, c0, acc[h][0]);
      if (in_f > 1) acc[h][1] = f2_fma(v, c1, acc[h][1]);
      if (in_f > 2) acc[h][2] = f2_fma(v, c2, acc[h][2]);
      acc[h][3] = f2_add(v, acc[h][3]);
    }
  };
  int stage = 0;
  uint32_t phase = 0;
  for (int i = 0; i < n_mine; ++i) {
    const int r0 = chunk_of(i) * kFw16sRows;
    const int rows = n - r0 < kFw16sRows ? n - r0 : kFw16sRows;
    const uint8_t* src = sm + size_t(stage) * stage_bytes;
    const float* cs = reinterpret_cast<const float*>(src + c_off);
    mbar_wait(smem_u32(&bar_full[stage]), phase);
    if (rows == kFw16sRows) {   // whole chunk: straight-line, all rows of this warp in flight together
#pragma unroll
      for (int u = 0; u < kFw16sRows / (kFw16sWarps * rpw); ++u) add_row(src, cs, warp * rpw + rsub + u * kFw16sWarps * rpw);
    } else {                    // ragged last chunk
      for (int r = warp * rpw + rsub; r < rows; r += kFw16sWarps * rpw) add_row(src, cs, r);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_empty[stage]));
    if (++stage == stages) { stage = 0; phase ^= 1; }
  }
  // Block reduction without shared-memory atomics (float atomicAdd on shared memory is a compare-and-swap loop: 32 of them per
  // thread with 16 warps contending cost more than the whole streaming part).  Every compute warp has consumed all its chunks,
  // so every bulk copy has landed and the ring is free: warp w parks its 32 x 32 partial sums there as part[w][v][lane].
  float* part = reinterpret_cast<float*>(sm);
  asm volatile("bar.sync 1, %0;" ::"n"(32 * kFw16sWarps) : "memory");
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      part[(warp * 32 + 4 * (2 * h) + d) * 32 + lane] = f2_lo(acc[h][d]);
      part[(warp * 32 + 4 * (2 * h + 1) + d) * 32 + lane] = f2_hi(acc[h][d]);
    }
  asm volatile("bar.sync 1, %0;" ::"n"(32 * kFw16sWarps) : "memory");
  for (int i = threadIdx.x; i < 32 * 32; i += 32 * kFw16sWarps) {
    const int v = i >> 5, l = i & 31, f = v >> 2, d = v & 3;   // l = octet, v = (feature of the octet, which sum)
    const int j = 8 * l + f;
    if (l < lpr && j < M && (d == 3 || d < in_f)) {
      float sum = 0.f;
      for (int w = 0; w < kFw16sWarps; ++w)
        for (int rs = 0; rs < rpw; ++rs) sum += part[(w * 32 + v) * 32 + l + rs * lpr];
      if (d < 3) atomicAdd(gW0 + size_t(j) * in_f + d, sum);
      else atomicAdd(gb0 + j, sum);
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
