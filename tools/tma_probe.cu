// tma_probe.cu — TMA traffic of the two-store forward layer (tc_rows16<GABOR_FWD>) without its math: how fast can one CTA per
// SM load its operand stages and store its output tiles, with 64-byte-row store boxes (32 FP16 columns, what the kernel does) and
// with 128-byte-row boxes (64 columns)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I wire_b200/csrc -o tools/tma_probe tools/tma_probe.cu -lcuda
// Per 128-row tile and CTA: 14 operand stages of 16 KB (A box 64 cols x 128 rows, from a [N][448] FP16 tensor) + 14 KB (B box, from a
// [448][448] weight tensor that stays in L2); 14 chunks x 4 row quarters of output for each of two [N][448] FP16 tensors.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "sm100.cuh"
using namespace sm100;

struct Params {
  CUtensorMap a_map, b_map;       // loads
  CUtensorMap o32[2], o64[2];     // stores: 32-column (SW64) and 64-column (SW128) boxes
  int n_tiles, do_load, store_mode;   // store_mode: 0 none, 1 = 64-byte rows (2 x 14 x 4 boxes of 2 KB), 2 = 128-byte rows (2 x 7 x 4 boxes of 4 KB)
  int n_store;                    // tensors stored (1 or 2)
  int vc;                         // valid columns
  int tile_major;
  uint8_t* o_ptr[2];              // store modes 3 / 4: the same tiles leave through LDS.128 + coalesced st.global.v4 (3: 64-byte rows, 8 rows per
                                  // warp instruction; 4: 128-byte rows, 4 rows per instruction) instead of TMA stores
};

constexpr int kStages = 5;
constexpr uint32_t kStageBytes = 16384 + 14336;

__global__ void __launch_bounds__(64 + 512, 1) probe_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kStages], bar_empty[kStages];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging = base + kStages * kStageBytes;   // 16 warps x 4 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0 && P.do_load) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x)
        for (int s = 0; s < 14; ++s) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t bar = smem_u32(&bar_full[stage]);
          mbar_expect_tx(bar, kStageBytes);
          if (P.tile_major) tma_load_2d(base + stage * kStageBytes, &P.a_map, bar, 0, (t * 7 + (s % 7)) * 128);
          else tma_load_2d(base + stage * kStageBytes, &P.a_map, bar, (s % 7) * 64, t * 128);
          tma_load_2d(base + stage * kStageBytes + 16384, &P.b_map, bar, (s % 7) * 64, (s / 7) * 224);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (lane == 0 && P.do_load) {   // stands in for the MMA issuer: consumes every stage as soon as it has landed
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x)
        for (int s = 0; s < 14; ++s) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          mbar_arrive(smem_u32(&bar_empty[stage]));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (P.store_mode) {
    const int ew = warp - 2, q = ew & 3, part = ew >> 2;
    const uint32_t tile = staging + ew * 4096;
    if (P.store_mode >= 3) {
      const int wide = P.store_mode == 4;
      const int rpi = wide ? 4 : 8, ppr = wide ? 8 : 4;       // rows per instruction, 16-byte pieces per row
      const int r_in = lane / ppr, piece = lane % ppr;
      for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const int row = t * 128 + q * 32;
        const int n_units = wide ? 7 : 14, cols = wide ? 64 : 32;
        for (int u = part; u < n_units; u += 4) {
          for (int s = 0; s < P.n_store; ++s) {
            uint4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i < 32 / rpi) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                                             : "r"(tile + (wide ? 0 : s * 2048) + (i * rpi + r_in) * (wide ? 128 : 64) + piece * 16));
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i < 32 / rpi) {
                const int c = u * cols + piece * 8;
                if (c < P.vc) *reinterpret_cast<uint4*>(P.o_ptr[s] + (size_t(row + i * rpi + r_in) * 448 + c) * 2) = v[i];
              }
          }
        }
      }
    } else if (lane == 0) {
      for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const int row = t * 128 + q * 32;
        if (P.store_mode == 1) {
          for (int ch = part; ch < 14; ch += 4) {
            tma_store_wait_read<0>();
            for (int s = 0; s < P.n_store; ++s) {
              if (P.tile_major) tma_store_2d(&P.o32[s], tile + s * 2048, (ch & 1) * 32, (t * 7 + (ch >> 1)) * 128 + q * 32);
              else tma_store_2d(&P.o32[s], tile + s * 2048, ch * 32, row);
            }
            tma_store_commit();
          }
        } else {
          for (int cp = part; cp < 7; cp += 4) {
            for (int s = 0; s < P.n_store; ++s) {   // the 4 KB tile is reused for the second tensor: wait in between
              tma_store_wait_read<0>();
              tma_store_2d(&P.o64[s], tile, cp * 64, row);
              tma_store_commit();
            }
          }
        }
      }
      tma_store_wait_all<0>();
    }
  }
}

int main(int argc, char** argv) {
  const int N = 262144, W = 448;
  const int VC = argc > 1 ? atoi(argv[1]) : 424;   // valid (written) columns per row
  const int only = argc > 2 ? atoi(argv[2]) : -1;  // run only this store mode
  const int tile_major = argc > 3 ? atoi(argv[3]) : 0;   // 1: activations stored as [row tile][k block of 64 cols][128 rows][64 cols]: every box is contiguous
  void *a, *b, *o0, *o1;
  cudaMalloc(&a, size_t(N) * W * 2); cudaMalloc(&o0, size_t(N) * W * 2); cudaMalloc(&o1, size_t(N) * W * 2); cudaMalloc(&b, size_t(W) * W * 2);
  cudaMemset(a, 0, size_t(N) * W * 2); cudaMemset(b, 0, size_t(W) * W * 2);
  Params P;
  P.tile_major = tile_major;
  bool ok = tile_major ? sm100_host::make_tmap_2d_t(&P.a_map, a, size_t(N) * 7, 64, 64, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, sm100_host::kElemF16)
                       : sm100_host::make_tmap_2d_t(&P.a_map, a, N, W, W, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, sm100_host::kElemF16);
  ok &= sm100_host::make_tmap_2d_t(&P.b_map, b, W, W, W, 112, 64, CU_TENSOR_MAP_SWIZZLE_128B, sm100_host::kElemF16);
  void* o[2] = {o0, o1};
  for (int s = 0; s < 2; ++s) {
    if (tile_major) ok &= sm100_host::make_tmap_2d_t(&P.o32[s], o[s], size_t(N) * 7, 64, 64, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, sm100_host::kElemF16);
    else ok &= sm100_host::make_tmap_2d_t(&P.o32[s], o[s], N, VC, W, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, sm100_host::kElemF16);
    ok &= sm100_host::make_tmap_2d_t(&P.o64[s], o[s], N, VC, W, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B, sm100_host::kElemF16);
  }
  if (!ok) { printf("tensor map failed\n"); return 1; }
  P.n_tiles = N / 128; P.vc = VC;
  printf("valid columns %d of %d%s\n", VC, W, tile_major ? "  TILE-MAJOR activations (mode 1 only)" : "");
  P.o_ptr[0] = (uint8_t*)o0; P.o_ptr[1] = (uint8_t*)o1;
  const size_t smem = kStages * kStageBytes + 16 * 4096 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("per launch: %d row tiles of 128 x 448 FP16 on 148 CTAs; loads 420 KB / tile, stores 112 KB / tile and tensor\n", P.n_tiles);
  for (int load = 0; load <= 1; ++load)
    for (int mode = 0; mode <= 4; ++mode)
      for (int ns = 1; ns <= 2; ++ns) {
        if (!load && !mode) continue;
        if (only >= 0 && mode && mode != only) continue;
        if (!mode && ns == 2) continue;
        P.do_load = load; P.store_mode = mode; P.n_store = ns;
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
          cudaEventRecord(e0);
          probe_kernel<<<148, 64 + 512, smem>>>(P);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (it && ms < best) best = ms;
        }
        printf("loads %d  stores: %-22s tensors %d   %.1f us\n", load, mode == 0 ? "none" : (mode == 1 ? "TMA 64-byte rows" : (mode == 2 ? "TMA 128-byte rows" : (mode == 3 ? "st.global 64-byte rows" : "st.global 128-byte rows"))), mode ? ns : 0, best * 1e3);
      }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
