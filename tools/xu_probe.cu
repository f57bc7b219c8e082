// xu_probe.cu — throughput of the XU (MUFU) pipe on sm_100a: ex2 / sin / cos / rsqrt, lanes per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_probe xu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i * 0.1f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678f) out[0] = s;
}
template <int OP>
void run(const char* name, int sms) {
  float* d; cudaMalloc(&d, 4);
  const int iters = 2000;
  k<OP><<<sms * 2, 1024>>>(d, 10, 0.3f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<sms * 2, 1024>>>(d, iters, 0.3f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = double(sms) * 2 * 1024 * iters * 8;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-6s %.3f ms  %.1f Gop/s  = %.2f lanes/clk/SM at %d MHz (nominal)\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(d);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  run<0>("ex2", p.multiProcessorCount);
  run<1>("sin", p.multiProcessorCount);
  run<2>("cos", p.multiProcessorCount);
  run<3>("rsqrt", p.multiProcessorCount);
  run<4>("ffma", p.multiProcessorCount);
  return 0;
}
