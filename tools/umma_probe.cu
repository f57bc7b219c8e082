// umma_probe.cu — hardware bring-up probe for the tcgen05 kernels (run on a B200 via gpurun).
//
//   1. tc_rows_kernel<MODE_PLAIN> (K-major A and B, SW128 TMA, TMEM epilogue, TMA store) vs a CPU
//      double-precision GEMM, at the WIRE shapes (2M = 424 -> nb = 432 split 256+176, K tail).
//   2. tc_wgrad_kernel (both operands MN-major, lane-pair complex fold, split-K atomics) vs CPU.
//   3. micro-benchmarks: TF32 tcgen05 issue-rate peak, and tc_rows / tc_wgrad at the 512^2 shape.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o umma_probe umma_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../wire_b200/csrc/tc_launch.cuh"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static float tf32_round_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;  // round-to-nearest (ties away) on the 13 dropped bits
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
static float frand() { return float(rand()) / float(RAND_MAX) * 2.f - 1.f; }

// ---- MMA issue-rate peak: every CTA loops tcgen05.mma on resident (zero) smem tiles ----
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters, int n_cols) {
  using namespace sm100;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (16384 + 256 * 128) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(128, n_cols, false, false);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = make_sdesc_sw128(base + ks * 32, 16, 1024);
        const uint64_t bd = make_sdesc_sw128(base + 16384 + ks * 32, 16, 1024);
        umma_tf32(tm, ad, bd, idesc, 1);
        umma_tf32(tm + 256, ad, bd, idesc, 1);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

static int g_sms = 148;
static bool g_skip_check = false;
static int g_cluster = 1;

static bool test_rows(int n_rows, int two_m, bool time_it) {
  const int K = two_m, pitch = wire::round_up(two_m + 1, 32);
  const int nb = two_m > 256 ? wire::round_up(two_m, 64) : wire::round_up(two_m, 16);
  std::vector<float> A(size_t(n_rows) * pitch, 0.f), B(size_t(nb) * pitch, 0.f), C(size_t(n_rows) * pitch, -7.f);
  for (int r = 0; r < n_rows; ++r) {
    for (int c = 0; c < K; ++c) A[size_t(r) * pitch + c] = tf32_round_host(frand());
    A[size_t(r) * pitch + K] = 1.0f;  // ones column must NOT leak into the K loop
  }
  for (int r = 0; r < two_m; ++r)
    for (int c = 0; c < K; ++c) B[size_t(r) * pitch + c] = tf32_round_host(frand() * 0.1f);
  float *dA, *dB, *dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice));

  wire::RowsParams P;
  memset(&P, 0, sizeof(P));
  P.e.n_rows = n_rows; P.k_cols[0] = K; P.k_cols[1] = 0; P.n_blocks = 1; P.e.n_cols = two_m;
  size_t smem = wire::rows_configure(P, nb, nb, 1, 0, two_m, wire::MODE_PLAIN, false, g_cluster);
  if (!smem) { printf("rows_configure failed\n"); return false; }
  bool ok = sm100_host::make_tmap_2d(&P.a_map[0], dA, n_rows, K, pitch, 128, 32);
  ok &= sm100_host::make_tmap_2d(&P.a_map[1], dA, n_rows, K, pitch, 128, 32);
  ok &= sm100_host::make_tmap_2d(&P.b_map, dB, nb, pitch, pitch, P.b_box_rows, 32);
  ok &= sm100_host::make_tmap_2d(&P.o_map[0], dC, n_rows, two_m, pitch, 32, 32);
  P.o_map[1] = P.o_map[0]; P.o_map[2] = P.o_map[0];
  if (!ok) { printf("tensor map creation failed\n"); return false; }
  printf("[rows] cluster=%d n_rows=%d 2M=%d nb=%d stages=%d smem=%zu\n", g_cluster, n_rows, two_m, nb, P.stages, smem);
  CK(wire::launch_rows(wire::MODE_PLAIN, P, smem, g_sms, 0));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  long bad = 0;
  const int check_rows = n_rows < 1024 ? n_rows : 1024;
  for (int rr = 0; rr < check_rows; ++rr) {
    const int r = (n_rows <= 1024) ? rr : int((long long)rr * 9973 % n_rows);
    for (int c = 0; c < two_m; ++c) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += double(A[size_t(r) * pitch + k]) * double(B[size_t(c) * pitch + k]);
      const double err = fabs(acc - double(C[size_t(r) * pitch + c]));
      if (err > max_err) max_err = err;
      if (fabs(acc) > max_ref) max_ref = fabs(acc);
      if (err > 1e-3) ++bad;
    }
    // pad columns must be untouched
    if (C[size_t(r) * pitch + two_m] != -7.f) ++bad;
  }
  printf("[rows] max_abs_err=%.3e (max |ref| %.3f) bad=%ld -> %s\n", max_err, max_ref, bad, bad ? "FAIL" : "PASS");
  if (time_it && !bad) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    {  // stall counters
      unsigned long long* dd; CK(cudaMalloc(&dd, 8 * 8 * 1024)); CK(cudaMemset(dd, 0, 8 * 8 * 1024));
      wire::RowsParams Q = P; Q.dbg = dd;
      CK(wire::launch_rows(wire::MODE_PLAIN, Q, smem, g_sms, 0));
      CK(cudaDeviceSynchronize());
      std::vector<unsigned long long> hd(8 * 1024);
      CK(cudaMemcpy(hd.data(), dd, hd.size() * 8, cudaMemcpyDeviceToHost));
      const char* names[8] = {"mma_wait_full", "mma_wait_tmem", "mma_total", "epi_wait_acc", "epi_total", "prod_wait_empty", "prod_total", "epi_wait_in"};
      for (int k = 0; k < 8; ++k) {
        double sum = 0; int cnt = 0;
        for (int b = 0; b < 1024; ++b) if (hd[b * 8 + k]) { sum += double(hd[b * 8 + k]); ++cnt; }
        printf("   [dbg] %-16s avg %.0f cycles over %d CTAs\n", names[k], cnt ? sum / cnt : 0.0, cnt);
      }
      cudaFree(dd);
    }
    for (int i = 0; i < 3; ++i) CK(wire::launch_rows(wire::MODE_PLAIN, P, smem, g_sms, 0));
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; ++i) CK(wire::launch_rows(wire::MODE_PLAIN, P, smem, g_sms, 0));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    const double flop = 2.0 * n_rows * double(two_m) * K;
    printf("[rows] %.3f ms  %.1f TFLOP/s (useful)  A+C traffic %.1f GB/s\n", ms, flop / ms * 1e-9,
           (2.0 * n_rows * two_m * 4) / ms * 1e-6);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return bad == 0;
}

// timing + stall counters of the fused forward epilogue (no correctness check here: tests/ cover that)
static void time_rows_gabor(int n_rows, int two_m) {
  const int K = two_m, pitch = wire::round_up(two_m + 1, 32);
  const int nb = two_m > 256 ? wire::round_up(two_m, 64) : wire::round_up(two_m, 16);
  std::vector<float> A(size_t(n_rows) * pitch, 0.f), B(size_t(nb) * pitch, 0.f), bias(two_m, 0.01f);
  for (auto& v : A) v = tf32_round_host(frand() * 0.5f);
  for (auto& v : B) v = tf32_round_host(frand() * 0.07f);
  float *dA, *dB, *dY, *dZ, *dbias, *dom;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dY, A.size() * 4)); CK(cudaMalloc(&dZ, A.size() * 4));
  CK(cudaMalloc(&dbias, two_m * 4)); CK(cudaMalloc(&dom, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, bias.data(), two_m * 4, cudaMemcpyHostToDevice));
  const float os[2] = {7.f, 6.f};
  CK(cudaMemcpy(dom, os, 8, cudaMemcpyHostToDevice));
  wire::RowsParams P;
  memset(&P, 0, sizeof(P));
  P.e.n_rows = n_rows; P.k_cols[0] = K; P.n_blocks = 1; P.e.n_cols = two_m; P.e.round_out0 = 1;
  P.e.bias = dbias; P.e.omega = dom; P.e.scale = dom + 1;
  size_t smem = wire::rows_configure(P, nb, nb, 3, 0, two_m, wire::MODE_GABOR_FWD, false, g_cluster);
  bool ok = sm100_host::make_tmap_2d(&P.a_map[0], dA, n_rows, K, pitch, 128, 32);
  P.a_map[1] = P.a_map[0];
  ok &= sm100_host::make_tmap_2d(&P.b_map, dB, nb, pitch, pitch, P.b_box_rows, 32);
  ok &= sm100_host::make_tmap_2d(&P.o_map[0], dY, n_rows, two_m, pitch, 32, 32);
  ok &= sm100_host::make_tmap_2d(&P.o_map[1], dZ, n_rows, two_m, pitch, 32, 32);
  P.o_map[2] = P.o_map[0]; P.z_map[0] = P.a_map[0]; P.z_map[1] = P.a_map[0];
  if (!ok || !smem) { printf("gabor setup failed\n"); return; }
  unsigned long long* dd; CK(cudaMalloc(&dd, 8 * 8 * 1024)); CK(cudaMemset(dd, 0, 8 * 8 * 1024));
  P.dbg = dd;
  for (int i = 0; i < 3; ++i) CK(wire::launch_rows(wire::MODE_GABOR_FWD, P, smem, g_sms, 0));
  CK(cudaDeviceSynchronize());
  std::vector<unsigned long long> hd(8 * 1024);
  CK(cudaMemcpy(hd.data(), dd, hd.size() * 8, cudaMemcpyDeviceToHost));
  printf("[gabor_fwd] cluster=%d n_rows=%d 2M=%d nb=%d slices=%d stages=%d smem=%zu\n", g_cluster, n_rows, two_m, nb, P.slices, P.stages, smem);
  const char* names[12] = {"mma_wait_full", "mma_wait_tmem", "mma_total", "epi_wait_acc", "epi_total", "prod_wait_empty", "prod_total", "epi_wait_in",
                           "t_prologue", "t_producer_done", "t_cta_done", "t_exit"};
  for (int k = 0; k < 12; ++k) {
    double sum = 0; int cnt = 0;
    for (int b = 0; b < 1024; ++b) if (hd[b * 16 + k]) { sum += double(hd[b * 16 + k]); ++cnt; }
    printf("   [dbg] %-16s avg %.0f cycles over %d CTAs\n", names[k], cnt ? sum / cnt : 0.0, cnt);
  }
  P.dbg = nullptr;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int i = 0; i < reps; ++i) CK(wire::launch_rows(wire::MODE_GABOR_FWD, P, smem, g_sms, 0));
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  printf("[gabor_fwd] %.3f ms  %.1f TFLOP/s (useful)\n", ms, 2.0 * n_rows * double(two_m) * K / ms * 1e-9);
  cudaFree(dA); cudaFree(dB); cudaFree(dY); cudaFree(dZ); cudaFree(dbias); cudaFree(dom); cudaFree(dd);
}

static bool test_wgrad(int n_rows, int k_in, int m_out, bool time_it) {
  const int xc = 2 * k_in + 1, gc = 2 * m_out;
  const int xp = wire::round_up(xc, 32), gp = wire::round_up(gc + 1, 32);
  std::vector<float> X(size_t(n_rows) * xp, 0.f), G(size_t(n_rows) * gp, 0.f);
  for (int r = 0; r < n_rows; ++r) {
    for (int c = 0; c < 2 * k_in; ++c) X[size_t(r) * xp + c] = tf32_round_host(frand());
    X[size_t(r) * xp + 2 * k_in] = 1.0f;
    for (int c = 0; c < gc; ++c) G[size_t(r) * gp + c] = tf32_round_host(frand() * 0.1f);
  }
  float *dX, *dG, *dW, *dBias;
  CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dG, G.size() * 4));
  CK(cudaMalloc(&dW, size_t(m_out) * k_in * 2 * 4)); CK(cudaMalloc(&dBias, size_t(m_out) * 2 * 4));
  CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dW, 0, size_t(m_out) * k_in * 2 * 4)); CK(cudaMemset(dBias, 0, size_t(m_out) * 2 * 4));
  wire::WgradParams P;
  memset(&P, 0, sizeof(P));
  P.n_rows = n_rows; P.k_in = k_in; P.g_cols = gc; P.n_g = 1;
  P.gW[0] = dW; P.gB[0] = dBias;
  size_t smem = wire::wgrad_configure(P, g_sms, g_cluster);
  bool ok = sm100_host::make_tmap_2d(&P.x_map, dX, n_rows, xc, xp, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  ok &= sm100_host::make_tmap_2d(&P.g_map[0], dG, n_rows, gc, gp, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  P.g_map[1] = P.g_map[0];
  if (!ok || !smem) { printf("wgrad setup failed\n"); return false; }
  printf("[wgrad] cluster=%d ", g_cluster);
  printf("[wgrad] n=%d K=%d M=%d m_tiles=%d n_blocks=%d nb=%d splits=%d stages=%d smem=%zu\n", n_rows, k_in, m_out,
         P.m_tiles, P.n_blocks, P.nb, P.splits, P.stages, smem);
  CK(wire::launch_wgrad(P, smem, 0));
  CK(cudaDeviceSynchronize());
  std::vector<float> W(size_t(m_out) * k_in * 2), Bv(size_t(m_out) * 2);
  CK(cudaMemcpy(W.data(), dW, W.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Bv.data(), dBias, Bv.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0; long bad = 0;
  const int jstep = m_out > 64 ? 7 : 1, kstep = k_in > 64 ? 5 : 1;
  for (int j = 0; j < (g_skip_check ? 0 : m_out); j += jstep) {
    for (int k = 0; k < k_in; k += kstep) {
      double re = 0, im = 0;
      for (int n = 0; n < n_rows; ++n) {
        const double gr = G[size_t(n) * gp + 2 * j], gi = G[size_t(n) * gp + 2 * j + 1];
        const double xr = X[size_t(n) * xp + 2 * k], xi = X[size_t(n) * xp + 2 * k + 1];
        re += gr * xr + gi * xi;   // g * conj(x)
        im += gi * xr - gr * xi;
      }
      const double e = fmax(fabs(re - W[(size_t(j) * k_in + k) * 2]), fabs(im - W[(size_t(j) * k_in + k) * 2 + 1]));
      if (e > max_err) max_err = e;
      if (fabs(re) > max_ref) max_ref = fabs(re);
      if (e > 1e-3 * (1.0 + sqrt(double(n_rows)) * 0.01)) ++bad;
    }
    double br = 0, bi = 0;
    for (int n = 0; n < n_rows; ++n) { br += G[size_t(n) * gp + 2 * j]; bi += G[size_t(n) * gp + 2 * j + 1]; }
    const double e = fmax(fabs(br - Bv[2 * j]), fabs(bi - Bv[2 * j + 1]));
    if (e > max_err) max_err = e;
    if (e > 1e-3 * (1.0 + sqrt(double(n_rows)) * 0.01)) ++bad;
  }
  printf("[wgrad] max_abs_err=%.3e (max |ref| %.3f) bad=%ld -> %s\n", max_err, max_ref, bad, bad ? "FAIL" : "PASS");
  if (time_it && !bad) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(wire::launch_wgrad(P, smem, 0));
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; ++i) CK(wire::launch_wgrad(P, smem, 0));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    const double flop = 8.0 * n_rows * double(m_out) * k_in;
    printf("[wgrad] %.3f ms  %.1f TFLOP/s (algorithmic 8MK)  read traffic %.1f GB/s\n", ms, flop / ms * 1e-9,
           (double(n_rows) * (xc + gc) * 4) / ms * 1e-6);
  }
  cudaFree(dX); cudaFree(dG); cudaFree(dW); cudaFree(dBias);
  return bad == 0;
}

// ---- 16-bit operand kernels (kind::f16; FP16 x BF16 mixing) ----
static uint16_t to16(float v, int fmt) {
  if (fmt == 1) { __half h = __float2half_rn(v); uint16_t u; memcpy(&u, &h, 2); return u; }
  __nv_bfloat16 h = __float2bfloat16_rn(v); uint16_t u; memcpy(&u, &h, 2); return u;
}
static float from16(uint16_t u, int fmt) {
  if (fmt == 1) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
  __nv_bfloat16 h; memcpy(&h, &u, 2); return __bfloat162float(h);
}
// a_elem / b_elem / o_elem: sm100_host::ElemType (1 = f16, 2 = bf16; o_elem may be 0 = f32)
static bool test_rows16(int n_rows, int two_m, int a_elem, int b_elem, int o_elem, bool time_it) {
  const int K = two_m, pitch = wire::round_up(two_m + 1, 32), kpad = wire::round_up(two_m, 64);
  const int nb = two_m > 256 ? wire::round_up(two_m, 64) : wire::round_up(two_m, 16);
  std::vector<uint16_t> A(size_t(n_rows) * pitch, 0), B(size_t(nb) * kpad, 0);
  std::vector<float> Af(A.size(), 0.f), Bf(B.size(), 0.f);
  for (int r = 0; r < n_rows; ++r) {
    for (int c = 0; c < K; ++c) { A[size_t(r) * pitch + c] = to16(frand(), a_elem); Af[size_t(r) * pitch + c] = from16(A[size_t(r) * pitch + c], a_elem); }
    A[size_t(r) * pitch + K] = to16(1.0f, a_elem);  // ones column must NOT leak into the K loop
  }
  for (int r = 0; r < two_m; ++r)
    for (int c = 0; c < K; ++c) { B[size_t(r) * kpad + c] = to16(frand() * 0.1f, b_elem); Bf[size_t(r) * kpad + c] = from16(B[size_t(r) * kpad + c], b_elem); }
  const size_t osz = o_elem == 0 ? 4 : 2;
  void *dA, *dB, *dC;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dC, size_t(n_rows) * pitch * osz));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0, size_t(n_rows) * pitch * osz));
  wire::RowsParams P;
  memset(&P, 0, sizeof(P));
  P.e.n_rows = n_rows; P.k_cols[0] = K; P.k_cols[1] = 0; P.n_blocks = 1; P.e.n_cols = two_m;
  size_t smem = wire::rows16_configure(P, nb, nb, 1, 0, two_m, wire::MODE_PLAIN, false, g_cluster);
  if (!smem) { printf("rows16_configure failed\n"); return false; }
  P.a_fmt = a_elem - 1; P.b_fmt = b_elem - 1; P.o_fmt[0] = o_elem;
  bool ok = sm100_host::make_tmap_2d_t(&P.a_map[0], dA, n_rows, K, pitch, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, a_elem);
  P.a_map[1] = P.a_map[0];
  ok &= sm100_host::make_tmap_2d_t(&P.b_map, dB, nb, kpad, kpad, P.b_box_rows, 64, CU_TENSOR_MAP_SWIZZLE_128B, b_elem);
  if (o_elem == 0) { printf("rows16 stores 16-bit tensors only\n"); return false; }
  ok &= sm100_host::make_tmap_2d_t(&P.o_map[0], dC, n_rows, two_m, pitch, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, o_elem);
  P.o_map[1] = P.o_map[0]; P.o_map[2] = P.o_map[0]; P.z_map[0] = P.a_map[0]; P.z_map[1] = P.a_map[0];
  if (!ok) { printf("tensor map creation failed\n"); return false; }
  printf("[rows16] cluster=%d n_rows=%d 2M=%d nb=%d a=%d b=%d o=%d stages=%d\n", g_cluster, n_rows, two_m, nb, a_elem, b_elem, o_elem, P.stages);
  CK(wire::launch_rows16(wire::MODE_PLAIN, P, smem, g_sms, 0));
  CK(cudaDeviceSynchronize());
  std::vector<float> C(size_t(n_rows) * pitch);
  if (o_elem == 0) CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  else {
    std::vector<uint16_t> C16(C.size());
    CK(cudaMemcpy(C16.data(), dC, C16.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < C.size(); ++i) C[i] = from16(C16[i], o_elem);
  }
  double max_err = 0, max_ref = 0; long bad = 0;
  const int check_rows = n_rows < 1024 ? n_rows : 1024;
  const double tol = o_elem == 0 ? 1e-3 : (o_elem == 1 ? 4e-3 : 3e-2);
  for (int rr = 0; rr < check_rows; ++rr) {
    const int r = (n_rows <= 1024) ? rr : int((long long)rr * 9973 % n_rows);
    for (int c = 0; c < two_m; ++c) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += double(Af[size_t(r) * pitch + k]) * double(Bf[size_t(c) * kpad + k]);
      const double err = fabs(acc - double(C[size_t(r) * pitch + c]));
      if (err > max_err) max_err = err;
      if (fabs(acc) > max_ref) max_ref = fabs(acc);
      if (err > tol) ++bad;
    }
    if (C[size_t(r) * pitch + two_m] != 0.f) ++bad;  // pad column untouched
  }
  printf("[rows16] max_abs_err=%.3e (max |ref| %.3f) bad=%ld -> %s\n", max_err, max_ref, bad, bad ? "FAIL" : "PASS");
  if (time_it && !bad) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(wire::launch_rows16(wire::MODE_PLAIN, P, smem, g_sms, 0));
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; ++i) CK(wire::launch_rows16(wire::MODE_PLAIN, P, smem, g_sms, 0));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf("[rows16] %.3f ms  %.1f TFLOP/s (useful)  A+C traffic %.1f GB/s\n", ms, 2.0 * n_rows * double(two_m) * K / ms * 1e-9,
           (double(n_rows) * two_m * (2 + osz)) / ms * 1e-6);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return bad == 0;
}

static bool test_wgrad16(int n_rows, int k_in, int m_out, int x_elem, int g_elem, bool time_it) {
  const int xc = 2 * k_in + 1, gc = 2 * m_out;
  const int xp = wire::round_up(xc, 32), gp = wire::round_up(gc + 1, 32);
  std::vector<uint16_t> X(size_t(n_rows) * xp, 0), G(size_t(n_rows) * gp, 0);
  std::vector<float> Xf(X.size(), 0.f), Gf(G.size(), 0.f);
  for (int r = 0; r < n_rows; ++r) {
    for (int c = 0; c < 2 * k_in; ++c) {
      X[size_t(r) * xp + c] = to16(frand(), x_elem);
      Xf[size_t(r) * xp + c] = from16(X[size_t(r) * xp + c], x_elem);
      if (x_elem != g_elem) Xf[size_t(r) * xp + c] = from16(to16(Xf[size_t(r) * xp + c], g_elem), g_elem);  // converted in smem
    }
    X[size_t(r) * xp + 2 * k_in] = to16(1.0f, x_elem); Xf[size_t(r) * xp + 2 * k_in] = 1.0f;
    for (int c = 0; c < gc; ++c) { G[size_t(r) * gp + c] = to16(frand() * 0.1f, g_elem); Gf[size_t(r) * gp + c] = from16(G[size_t(r) * gp + c], g_elem); }
  }
  void *dX, *dG; float *dW, *dBias;
  CK(cudaMalloc(&dX, X.size() * 2)); CK(cudaMalloc(&dG, G.size() * 2));
  CK(cudaMalloc(&dW, size_t(m_out) * k_in * 2 * 4)); CK(cudaMalloc(&dBias, size_t(m_out) * 2 * 4));
  CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dG, G.data(), G.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dW, 0, size_t(m_out) * k_in * 2 * 4)); CK(cudaMemset(dBias, 0, size_t(m_out) * 2 * 4));
  wire::WgradParams P;
  memset(&P, 0, sizeof(P));
  P.n_rows = n_rows; P.k_in = k_in; P.g_cols = gc; P.n_g = 1;
  P.gW[0] = dW; P.gB[0] = dBias; P.x_fmt = x_elem - 1; P.g_fmt = g_elem - 1; P.x_conv = x_elem != g_elem;
  size_t smem = wire::wgrad_configure(P, g_sms, g_cluster, false, true);
  bool ok = sm100_host::make_tmap_2d_t(&P.x_map, dX, n_rows, xc, xp, wire::kWgradKC16, 64, CU_TENSOR_MAP_SWIZZLE_128B, x_elem);
  ok &= sm100_host::make_tmap_2d_t(&P.g_map[0], dG, n_rows, gc, gp, wire::kWgradKC16, 64, CU_TENSOR_MAP_SWIZZLE_128B, g_elem);
  P.g_map[1] = P.g_map[0];
  if (!ok || !smem) { printf("wgrad16 setup failed\n"); return false; }
  printf("[wgrad16] cluster=%d n=%d K=%d M=%d x=%d g=%d m_tiles=%d n_blocks=%d nb=%d splits=%d stages=%d\n", g_cluster, n_rows, k_in, m_out,
         x_elem, g_elem, P.m_tiles, P.n_blocks, P.nb, P.splits, P.stages);
  CK(wire::launch_wgrad(P, smem, 0, false, true));
  CK(cudaDeviceSynchronize());
  std::vector<float> W(size_t(m_out) * k_in * 2), Bv(size_t(m_out) * 2);
  CK(cudaMemcpy(W.data(), dW, W.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Bv.data(), dBias, Bv.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0; long bad = 0;
  const int jstep = m_out > 64 ? 7 : 1, kstep = k_in > 64 ? 5 : 1;
  for (int j = 0; j < (g_skip_check ? 0 : m_out); j += jstep) {
    for (int k = 0; k < k_in; k += kstep) {
      double re = 0, im = 0;
      for (int n = 0; n < n_rows; ++n) {
        const double gr = Gf[size_t(n) * gp + 2 * j], gi = Gf[size_t(n) * gp + 2 * j + 1];
        const double xr = Xf[size_t(n) * xp + 2 * k], xi = Xf[size_t(n) * xp + 2 * k + 1];
        re += gr * xr + gi * xi;
        im += gi * xr - gr * xi;
      }
      const double e = fmax(fabs(re - W[(size_t(j) * k_in + k) * 2]), fabs(im - W[(size_t(j) * k_in + k) * 2 + 1]));
      if (e > max_err) max_err = e;
      if (fabs(re) > max_ref) max_ref = fabs(re);
      if (e > 1e-3 * (1.0 + sqrt(double(n_rows)) * 0.01)) ++bad;
    }
    double br = 0, bi = 0;
    for (int n = 0; n < n_rows; ++n) { br += Gf[size_t(n) * gp + 2 * j]; bi += Gf[size_t(n) * gp + 2 * j + 1]; }
    const double e = fmax(fabs(br - Bv[2 * j]), fabs(bi - Bv[2 * j + 1]));
    if (e > max_err) max_err = e;
    if (e > 1e-3 * (1.0 + sqrt(double(n_rows)) * 0.01)) ++bad;
  }
  printf("[wgrad16] max_abs_err=%.3e (max |ref| %.3f) bad=%ld -> %s\n", max_err, max_ref, bad, bad ? "FAIL" : "PASS");
  if (time_it && !bad) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(wire::launch_wgrad(P, smem, 0, false, true));
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; ++i) CK(wire::launch_wgrad(P, smem, 0, false, true));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf("[wgrad16] %.3f ms  %.1f TFLOP/s (algorithmic 8MK)  read traffic %.1f GB/s\n", ms, 8.0 * n_rows * double(m_out) * k_in / ms * 1e-9,
           (double(n_rows) * (xc + gc) * 2) / ms * 1e-6);
    if (g_skip_check) {  // per-CTA phase times (cycles): prologue, K loop, epilogue, teardown
      unsigned long long* dd; CK(cudaMalloc(&dd, 8 * 8 * 1024)); CK(cudaMemset(dd, 0, 8 * 8 * 1024));
      P.dbg = dd;
      CK(wire::launch_wgrad(P, smem, 0, false, true));
      CK(cudaDeviceSynchronize());
      std::vector<unsigned long long> hd(8 * 1024);
      CK(cudaMemcpy(hd.data(), dd, hd.size() * 8, cudaMemcpyDeviceToHost));
      double ph[4] = {0, 0, 0, 0}; int cnt = 0;
      unsigned long long first = ~0ull, last = 0;
      for (int b = 0; b < 1024; ++b) if (hd[b * 8]) {
        for (int k = 0; k < 4; ++k) ph[k] += double(hd[b * 8 + k + 1] - hd[b * 8 + k]);
        ++cnt; if (hd[b * 8] < first) first = hd[b * 8]; if (hd[b * 8 + 4] > last) last = hd[b * 8 + 4];
      }
      printf("   [dbg] CTAs %d: prologue %.0f  K loop %.0f  epilogue %.0f  teardown %.0f cycles (avg); first entry -> last exit %.0f\n", cnt,
             ph[0] / cnt, ph[1] / cnt, ph[2] / cnt, ph[3] / cnt, double(last - first));
      P.dbg = nullptr; cudaFree(dd);
    }
  }
  cudaFree(dX); cudaFree(dG); cudaFree(dW); cudaFree(dBias);
  return bad == 0;
}

static void time_rows_gabor16(int n_rows, int two_m, bool fuse_final, int mask_override = -1) {
  const int K = two_m, pitch = wire::round_up(two_m + 1, 32), kpad = wire::round_up(two_m, 64);
  const int nb = two_m > 256 ? wire::round_up(two_m, 64) : wire::round_up(two_m, 16);
  std::vector<uint16_t> A(size_t(n_rows) * pitch, 0), B(size_t(nb) * kpad, 0);
  for (auto& v : A) v = to16(frand() * 0.5f, 1);
  for (auto& v : B) v = to16(frand() * 0.07f, 1);
  std::vector<float> bias(two_m, 0.01f), wf(size_t(3) * two_m, 0.05f);
  void *dA, *dB, *dY, *dZ; float *dbias, *dom, *dwf, *dout;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dY, A.size() * 2)); CK(cudaMalloc(&dZ, A.size() * 2));
  CK(cudaMalloc(&dbias, two_m * 4)); CK(cudaMalloc(&dom, 8)); CK(cudaMalloc(&dwf, wf.size() * 4)); CK(cudaMalloc(&dout, size_t(n_rows) * 3 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, bias.data(), two_m * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dwf, wf.data(), wf.size() * 4, cudaMemcpyHostToDevice));
  const float os[2] = {7.f, 6.f};
  CK(cudaMemcpy(dom, os, 8, cudaMemcpyHostToDevice));
  wire::RowsParams P;
  memset(&P, 0, sizeof(P));
  P.e.n_rows = n_rows; P.k_cols[0] = K; P.n_blocks = 1; P.e.n_cols = two_m; P.e.z_half = 1;
  P.e.bias = dbias; P.e.omega = dom; P.e.scale = dom + 1;
  if (fuse_final) { P.e.fuse_final = 1; P.e.wf = dwf; P.e.bf = dbias; P.e.out = dout; P.e.out_features = 3; }
  const int mask = mask_override >= 0 ? mask_override : (fuse_final ? 2 : 3);
  size_t smem = wire::rows16_configure(P, nb, nb, mask, 0, two_m, wire::MODE_GABOR_FWD, fuse_final, g_cluster);
  P.a_fmt = 0; P.b_fmt = 0; P.o_fmt[0] = 1; P.o_fmt[1] = 1;
  bool ok = sm100_host::make_tmap_2d_t(&P.a_map[0], dA, n_rows, K, pitch, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, 1);
  P.a_map[1] = P.a_map[0];
  ok &= sm100_host::make_tmap_2d_t(&P.b_map, dB, nb, kpad, kpad, P.b_box_rows, 64, CU_TENSOR_MAP_SWIZZLE_128B, 1);
  ok &= sm100_host::make_tmap_2d_t(&P.o_map[0], (mask & 1) ? dY : dZ, n_rows, two_m, pitch, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, 1);
  ok &= sm100_host::make_tmap_2d_t(&P.o_map[1], dZ, n_rows, two_m, pitch, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, 1);
  P.o_map[2] = P.o_map[0]; P.z_map[0] = P.a_map[0]; P.z_map[1] = P.a_map[0];
  if (!ok || !smem) { printf("gabor16 setup failed\n"); return; }
  unsigned long long* dd; CK(cudaMalloc(&dd, 16 * 8 * 1024)); CK(cudaMemset(dd, 0, 16 * 8 * 1024));
  P.dbg = dd;
  for (int i = 0; i < 3; ++i) CK(wire::launch_rows16(wire::MODE_GABOR_FWD, P, smem, g_sms, 0));
  CK(cudaDeviceSynchronize());
  std::vector<unsigned long long> hd(16 * 1024);
  CK(cudaMemcpy(hd.data(), dd, hd.size() * 8, cudaMemcpyDeviceToHost));
  printf("[gabor_fwd16] cluster=%d n_rows=%d 2M=%d nb=%d slices=%d stages=%d fuse_final=%d mask=%d smem=%zu\n", g_cluster, n_rows, two_m, nb, P.slices, P.stages, int(fuse_final), mask, smem);
  const char* names[12] = {"mma_wait_full", "mma_wait_tmem", "mma_total", "epi_wait_acc", "epi_total", "prod_wait_empty", "prod_total", "epi_wait_in",
                           "t_prologue", "t_producer_done", "t_cta_done", "t_exit"};
  for (int k = 0; k < 12; ++k) {
    double sum = 0; int cnt = 0;
    for (int b = 0; b < 1024; ++b) if (hd[b * 16 + k]) { sum += double(hd[b * 16 + k]); ++cnt; }
    printf("   [dbg] %-16s avg %.0f cycles over %d CTAs\n", names[k], cnt ? sum / cnt : 0.0, cnt);
  }
  {
    double d[2] = {0, 0}, c10[2] = {0, 0}; int cn[2] = {0, 0};
    for (int b = 0; b < 1024; ++b) if (hd[b * 16 + 11]) { d[b & 1] += double(hd[b * 16 + 11] - hd[b * 16 + 10]); c10[b & 1] += double(hd[b * 16 + 10]); ++cn[b & 1]; }
    printf("   [dbg] exit - cta_done: leader CTAs %.0f, peer CTAs %.0f cycles; cta_done leader %.0f peer %.0f\n", d[0] / (cn[0] ? cn[0] : 1), d[1] / (cn[1] ? cn[1] : 1),
           c10[0] / (cn[0] ? cn[0] : 1), c10[1] / (cn[1] ? cn[1] : 1));
  }
  P.dbg = nullptr;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int i = 0; i < reps; ++i) CK(wire::launch_rows16(wire::MODE_GABOR_FWD, P, smem, g_sms, 0));
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  printf("[gabor_fwd16] %.3f ms  %.1f TFLOP/s (useful)\n", ms, 2.0 * n_rows * double(two_m) * K / ms * 1e-9);
  cudaFree(dA); cudaFree(dB); cudaFree(dY); cudaFree(dZ); cudaFree(dbias); cudaFree(dom); cudaFree(dd); cudaFree(dwf); cudaFree(dout);
}

static int main16() {
  bool ok = true;
  for (int c : {1, 2}) {
    g_cluster = c;
    ok &= test_rows16(256, 64, 1, 1, 1, false);      // f16 x f16
    ok &= test_rows16(300, 424, 1, 1, 1, false);     // WIRE width, ragged rows, K tail (424 = 6*64 + 40)
    ok &= test_rows16(300, 424, 2, 2, 2, false);     // bf16 x bf16
    // test_rows16(300, 424, 2, 1, 0, false): bf16 A x f16 B in one kind::f16 MMA -> "illegal instruction" on sm_100a
    // (measured, profiles/r01_probe16.log): the two operands must share one format.
    ok &= test_rows16(1000, 180, 1, 1, 1, false);    // f16 output tiles
    ok &= test_rows16(1000, 180, 2, 2, 2, false);    // bf16 output tiles
    ok &= test_wgrad16(256, 32, 32, 1, 1, false);
    ok &= test_wgrad16(256, 32, 32, 1, 2, false);    // f16 x, bf16 g (the training configuration)
    ok &= test_wgrad16(5000, 212, 212, 1, 2, false);
    ok &= test_wgrad16(777, 90, 90, 1, 2, false);
  }
  for (int c : {1, 2}) {
    g_cluster = c;
    test_rows16(262144, 424, 1, 1, 1, true);
    test_rows16(262144, 424, 2, 2, 2, true);
    test_wgrad16(262144, 212, 212, 1, 2, true);
  }
  printf("PROBE16 %s\n", ok ? "PASS" : "FAIL");
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  printf("device %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, g_sms);
  srand(1234);
  if (argc > 1 && !strcmp(argv[1], "16")) return main16();
  if (argc > 1 && !strcmp(argv[1], "wn")) {  // wgrad time against the row count: fixed cost (prologue, split-K atomics epilogue) vs per-row cost
    g_cluster = 2; g_skip_check = true;
    for (int k : {1, 2, 4, 8, 16, 32, 64, 111}) test_wgrad16(37 * 64 * k, 212, 212, 1, 2, true);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "gp")) {  // phase stamps of the forward kernels (build with -DWIRE_B200_STALL_COUNTERS)
    g_cluster = 2;
    time_rows_gabor16(262144, 424, false);
    time_rows_gabor16(262144, 424, true);
    time_rows_gabor16(25000, 424, true);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "g16")) {
    g_cluster = 2;
    time_rows_gabor16(262144, 424, false);
    time_rows_gabor16(262144, 424, false, 1);   // y only
    time_rows_gabor16(262144, 424, false, 2);   // z only
    time_rows_gabor16(262144, 424, false, 0);   // no stores at all
    time_rows_gabor16(262144, 424, true);
    time_rows_gabor16(262144, 424, true, 0);    // fused final, no stores (inference)
    test_rows16(262144, 424, 1, 1, 1, true);
    return 0;
  }
  bool ok = true;
  ok &= test_rows(256, 64, false);       // single MMA piece, 2 K chunks
  ok &= test_rows(300, 424, false);      // WIRE width: nb=432 (256+176), K tail of 8, ragged rows
  ok &= test_rows(1000, 180, false);     // M=90: K tail of 20 columns (not a multiple of 8)
  ok &= test_wgrad(256, 32, 32, false);
  ok &= test_wgrad(5000, 212, 212, false);
  ok &= test_wgrad(777, 90, 90, false);
  {
    // ---- peak probe ----
    CK(cudaFuncSetAttribute(mma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 256 * 128 + 2048));
    for (int n : {256, 224, 128}) {
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      const int iters = 4000;
      mma_peak_kernel<<<g_sms, 128, 16384 + 256 * 128 + 2048>>>(100, n);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      mma_peak_kernel<<<g_sms, 128, 16384 + 256 * 128 + 2048>>>(iters, n);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double flop = double(g_sms) * iters * 8 * (2.0 * 128 * n * 8);
      printf("[peak] tf32 M=128 N=%d K=8 cta_group::1: %.3f ms -> %.1f TFLOP/s\n", n, ms, flop / ms * 1e-9);
    }
    test_rows(262144, 424, true);
    test_wgrad(262144, 212, 212, true);
    test_rows(262144, 256, true);
    time_rows_gabor(262144, 424);
    g_cluster = 2;
    ok &= test_rows(300, 424, false);
    ok &= test_rows(1000, 180, false);
    test_rows(262144, 424, true);
    test_rows(262144, 256, true);
    time_rows_gabor(262144, 424);
    time_rows_gabor(262144, 256);
    ok &= test_wgrad(256, 32, 32, false);
    ok &= test_wgrad(5000, 212, 212, false);
    ok &= test_wgrad(777, 90, 90, false);
    test_wgrad(262144, 212, 212, true);

  }
  printf("PROBE %s\n", ok ? "PASS" : "FAIL");
  return ok ? 0 : 1;
}
