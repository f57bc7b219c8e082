// peer_kernels.cuh — the data-parallel exchange step of the WIRE hot path, fused with the optimiser, over NVLink peer memory.
//
// coordinate-sharded data parallelism needs exactly one exchange per step, a SUM of the flat fp32
// weight-gradient buffer (0.73 MB for the denoise net).  Instead of a library all-reduce followed by Adam, every rank's
// Adam kernel READS ITS PEERS' GRADIENT BUFFERS DIRECTLY (P2P loads through NVSwitch), sums them in rank order — the same
// order on every rank, so the replicas' parameters stay bit-identical — and applies the update.  No NCCL call on the data
// path, nothing to un-flatten, and the step (spin barriers included) is ordinary kernels, so it is captured in the CUDA
// graph of the training step.
//
// Each rank owns one peer buffer (cudaMalloc + cudaIpc handle, mapped by every other rank):
//     uint32 arrive[kMaxPeers] | uint32 done[kMaxPeers] | pad to 256 B | float grad[count]
// Protocol of step e (1-based Adam step, the same on every rank):
//   in-barrier : rank r stores arrive[r] = e into EVERY rank's header (release, system scope) after its gradient kernels
//                (stream order + __threadfence_system); every block then waits until its LOCAL arrive[p] >= e for all p.
//   sum + Adam : g = grad_scale * sum_p grad_p[i] (system-scope loads, rank order), then torch.optim.Adam's update.
//   out-signal : the last block to finish stores done[r] = e into every rank's header.
//   peer_wait  : before the NEXT step's backward pass overwrites grad, a one-block kernel waits until the local done[p] >=
//                (completed steps) for all p, i.e. every peer has finished reading this rank's gradients.
// Waits are bounded (spin_limit SM cycles: WIRE_B200_PEER_TIMEOUT_S, default 600 s) and trap instead of hanging the GPU forever
// if a peer died.  All ranks must issue their steps in lock-step: a rank that pauses between steps for longer than the bound
// (rank-0-only evaluation, checkpointing) makes its peers trap -- use the all-reduce exchange (Trainer(peer_exchange=False))
// for such loops, or raise the bound.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace wire {

constexpr int kMaxPeers = 16;
constexpr int kPeerHeaderBytes = 256;

struct PeerTable {
  void* base[kMaxPeers];  // base[p] = rank p's peer buffer as mapped in THIS process (base[rank] = the local allocation)
  int world;
  int rank;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// spin until *flag >= want (monotonic counters); traps after spin_limit cycles so a dead peer cannot hang the box
__device__ __forceinline__ void peer_spin(const uint32_t* flag, uint32_t want, long long spin_limit) {
  const long long t0 = clock64();
  while (int32_t(ld_acquire_sys(flag) - want) < 0) {
    __nanosleep(64);
    if (clock64() - t0 > spin_limit) __trap();
  }
}

// count must be a multiple of 4 floats (the flat buffers are laid out in 16-byte slots)
__global__ void __launch_bounds__(256) adam_peer_kernel(float* __restrict__ p, const PeerTable T, float* __restrict__ m, float* __restrict__ v,
                                                        int64_t count, const float* __restrict__ lr_ptr, float b1, float b2, float eps, float wd,
                                                        long long* __restrict__ step_ptr, float grad_scale,
                                                        unsigned int* __restrict__ done_counter, long long spin_limit) {
  const long long step = *step_ptr + 1;
  const uint32_t e = uint32_t(step);
  uint32_t* local = static_cast<uint32_t*>(T.base[T.rank]);
  // ---- in-barrier ----
  if (blockIdx.x == 0 && threadIdx.x < T.world) {
    __threadfence_system();
    st_release_sys(static_cast<uint32_t*>(T.base[threadIdx.x]) + T.rank, e);
  }
  if (threadIdx.x < T.world) peer_spin(local + threadIdx.x, e, spin_limit);
  __syncthreads();

  const float lr = *lr_ptr;
  const float bc1 = 1.0f - powf(b1, float(step));
  const float bc2_sqrt = sqrtf(1.0f - powf(b2, float(step)));
  const float step_size = lr / bc1;
  const int64_t n4 = count >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < T.world; ++r) {  // rank order: identical summation order on every replica
      const float4 t = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(static_cast<const char*>(T.base[r]) + kPeerHeaderBytes) + i);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    float4 pi = p4[i], mi = m4[i], vi = v4[i];
    auto upd = [&](float gi, float& pp, float& mm, float& vv) {
      gi *= grad_scale;
      if (wd != 0.f) gi = fmaf(wd, pp, gi);
      mm = fmaf(b1, mm, (1.f - b1) * gi);
      vv = fmaf(b2, vv, (1.f - b2) * gi * gi);
      const float denom = sqrtf(vv) / bc2_sqrt + eps;
      pp = pp - step_size * (mm / denom);
    };
    upd(g.x, pi.x, mi.x, vi.x); upd(g.y, pi.y, mi.y, vi.y); upd(g.z, pi.z, mi.z, vi.z); upd(g.w, pi.w, mi.w, vi.w);
    p4[i] = pi; m4[i] = mi; v4[i] = vi;
  }
  // ---- out-signal: the last block bumps the step counter and tells every peer this rank is done reading ----
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(done_counter, 1u);
    s_last = (prev == gridDim.x - 1);
    if (s_last) { *step_ptr = step; *done_counter = 0u; }
  }
  __syncthreads();
  if (s_last && threadIdx.x < T.world) {
    __threadfence_system();
    st_release_sys(static_cast<uint32_t*>(T.base[threadIdx.x]) + kMaxPeers + T.rank, e);
  }
}

// waits until every peer has finished reading this rank's gradient buffer of the last completed step
__global__ void peer_wait_kernel(const PeerTable T, const long long* __restrict__ step_ptr, long long spin_limit) {
  const uint32_t e = uint32_t(*step_ptr);
  const uint32_t* local = static_cast<const uint32_t*>(T.base[T.rank]) + kMaxPeers;
  if (threadIdx.x < T.world) peer_spin(local + threadIdx.x, e, spin_limit);
}

}  // namespace wire
