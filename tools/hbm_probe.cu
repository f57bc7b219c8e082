// hbm_probe.cu — HBM bandwidth of a B200 by read : write mix (roofline denominators for the activation-streaming kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_probe tools/hbm_probe.cu && tools/hbm_probe
// Every thread moves 16-byte pieces in a grid-stride loop; `R` source buffers are read and `W` destination buffers written per
// piece (R, W in 0..2), buffers of 1 GiB each (far beyond the 126 MB L2).  Prints GB/s of (read + written) bytes, best of 5.
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int W, bool STREAM>
__global__ void __launch_bounds__(256) mix_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ c,
                                                  uint4* __restrict__ d, size_t n, unsigned* sink) {
  unsigned acc = 0;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    uint4 v = make_uint4(unsigned(i), 1u, 2u, 3u);
    if (R >= 1) { const uint4 t = STREAM ? __ldcs(a + i) : a[i]; v.x ^= t.x; v.y += t.y; v.z ^= t.z; v.w += t.w; }
    if (R >= 2) { const uint4 t = STREAM ? __ldcs(b + i) : b[i]; v.x ^= t.x; v.y += t.y; v.z ^= t.z; v.w += t.w; }
    if (W >= 1) { if (STREAM) __stcs(c + i, v); else c[i] = v; }
    if (W >= 2) { if (STREAM) __stcs(d + i, v); else d[i] = v; }
    if (W == 0) acc ^= v.x + v.y + v.z + v.w;
  }
  if (W == 0 && acc == 0x12345678u) *sink = acc;
}

template <int R, int W, bool STREAM>
double run(const uint4* a, const uint4* b, uint4* c, uint4* d, size_t n, unsigned* sink, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 6; ++it) {
    cudaEventRecord(e0);
    mix_kernel<R, W, STREAM><<<blocks, 256>>>(a, b, c, d, n, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it && ms < best) best = ms;
  }
  return double(n) * 16.0 * (R + W) / (best * 1e-3) / 1e9;
}

int main() {
  const size_t bytes = size_t(1) << 30, n = bytes / 16;
  uint4 *a, *b, *c, *d; unsigned* sink;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&c, bytes); cudaMalloc(&d, bytes); cudaMalloc(&sink, 4);
  cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes); cudaMemset(c, 0, bytes); cudaMemset(d, 0, bytes);
  int sm = 148; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  for (int per = 8; per <= 16; per += 8) {
    const int blocks = sm * per;
    printf("blocks/SM %d  (GB/s of read+written bytes; plain | streaming hints)\n", per);
    printf("  read only      1R0W  %7.0f | %7.0f\n", run<1, 0, false>(a, b, c, d, n, sink, blocks), run<1, 0, true>(a, b, c, d, n, sink, blocks));
    printf("  2 reads        2R0W  %7.0f | %7.0f\n", run<2, 0, false>(a, b, c, d, n, sink, blocks), run<2, 0, true>(a, b, c, d, n, sink, blocks));
    printf("  write only     0R1W  %7.0f | %7.0f\n", run<0, 1, false>(a, b, c, d, n, sink, blocks), run<0, 1, true>(a, b, c, d, n, sink, blocks));
    printf("  2 writes       0R2W  %7.0f | %7.0f\n", run<0, 2, false>(a, b, c, d, n, sink, blocks), run<0, 2, true>(a, b, c, d, n, sink, blocks));
    printf("  copy           1R1W  %7.0f | %7.0f\n", run<1, 1, false>(a, b, c, d, n, sink, blocks), run<1, 1, true>(a, b, c, d, n, sink, blocks));
    printf("  2 reads 1 wr   2R1W  %7.0f | %7.0f\n", run<2, 1, false>(a, b, c, d, n, sink, blocks), run<2, 1, true>(a, b, c, d, n, sink, blocks));
    printf("  1 read 2 wr    1R2W  %7.0f | %7.0f\n", run<1, 2, false>(a, b, c, d, n, sink, blocks), run<1, 2, true>(a, b, c, d, n, sink, blocks));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
