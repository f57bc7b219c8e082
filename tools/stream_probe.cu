// stream_probe.cu — how fast can one producer lane per SM stream a [N][pitch] BF16 tensor through a shared-memory ring with
// cp.async.bulk (1-D) or 2-D tensor boxes?  (The first_wgrad16s kernel streams at 3.2 TB/s; is that the copy engine or the consumer?)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I wire_b200/csrc -o tools/stream_probe tools/stream_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "sm100.cuh"
using namespace sm100;

struct Params {
  CUtensorMap map;
  const uint8_t* src;
  int n_chunks, rows, pitch_bytes, stages, mode, descending, consumers, hint;   // mode 0: 1-D bulk, 1: 2-D box
};

__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, int hint) {
  if (hint)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(kEvictFirst) : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(544, 1) stream_kernel(const __grid_constant__ Params P, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t raw[];
  __shared__ __align__(8) uint64_t full[16], empty[16];
  const uint32_t base = (smem_u32(raw) + 127u) & ~127u;
  const uint32_t chunk_bytes = uint32_t(P.rows) * P.pitch_bytes;
  const uint32_t stage_bytes = (chunk_bytes + 127u) & ~127u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), P.consumers); }
    fence_barrier_init();
  }
  __syncthreads();
  const int n_mine = (P.n_chunks - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  auto chunk_of = [&](int i) { const int c = i * int(gridDim.x) + int(blockIdx.x); return P.descending ? P.n_chunks - 1 - c : c; };
  if (warp == 16) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < n_mine; ++i) {
        mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
        const uint32_t bar = smem_u32(&full[stage]);
        mbar_expect_tx(bar, chunk_bytes);
        const int c = chunk_of(i);
        if (P.mode == 0) bulk_1d(base + stage * stage_bytes, P.src + size_t(c) * chunk_bytes, chunk_bytes, bar, P.hint);
        else if (P.hint) tma_load_2d_hint(base + stage * stage_bytes, &P.map, bar, 0, c * P.rows, kEvictFirst);
        else tma_load_2d(base + stage * stage_bytes, &P.map, bar, 0, c * P.rows);
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < P.consumers) {
    int stage = 0; uint32_t phase = 0; unsigned acc = 0;
    for (int i = 0; i < n_mine; ++i) {
      mbar_wait(smem_u32(&full[stage]), phase);
      // touch 4 rows like the real kernel (one 16-byte piece per lane and row)
      for (int r = warp; r < P.rows; r += 16) {
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(base + stage * stage_bytes + r * P.pitch_bytes + (lane * 16) % P.pitch_bytes));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[stage]));
      if (++stage == P.stages) { stage = 0; phase ^= 1; }
    }
    if (acc == 0x9e3779b9u) *sink = acc;
  }
}

int main() {
  unsigned* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int shapes[2][2] = {{262144, 216}, {1048576, 128}};   // rows, pitch in BF16 elements
  for (auto& sh : shapes) {
    const int N = sh[0], pitch = sh[1];
    void* src; cudaMalloc(&src, size_t(N) * pitch * 2); cudaMemset(src, 1, size_t(N) * pitch * 2);
    printf("[%d][%d] BF16 = %.0f MB\n", N, pitch, N * double(pitch) * 2 / 1e6);
    for (int rows : {64, 128}) for (int mode = 0; mode <= 1; ++mode) for (int desc = 0; desc <= 1; ++desc) for (int hint = 0; hint <= 1; ++hint) {
      if (rows == 128 && (desc || hint)) continue;
      Params P;
      P.src = (const uint8_t*)src; P.rows = rows; P.pitch_bytes = pitch * 2; P.n_chunks = N / rows; P.mode = mode; P.descending = desc;
      P.consumers = 16; P.hint = hint;
      const uint32_t stage_bytes = (uint32_t(rows) * pitch * 2 + 127u) & ~127u;
      P.stages = int(200u * 1024u / stage_bytes); if (P.stages > 8) P.stages = 8;
      if (!sm100_host::make_tmap_2d_t(&P.map, src, N, pitch, pitch, rows, pitch, CU_TENSOR_MAP_SWIZZLE_NONE, sm100_host::kElemBF16)) { printf("tmap failed\n"); return 1; }
      const size_t smem = size_t(P.stages) * stage_bytes + 128;
      cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      float best = 1e30f;
      for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        stream_kernel<<<148, 544, smem>>>(P, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
      }
      printf("  rows/chunk %3d  stages %d  %-9s %-10s hint %d   %.1f us  %.2f TB/s\n", rows, P.stages, mode ? "2-D box" : "1-D bulk", desc ? "descending" : "ascending", hint,
             best * 1e3, N * double(pitch) * 2 / (best * 1e-3) / 1e12);
    }
    cudaFree(src);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
