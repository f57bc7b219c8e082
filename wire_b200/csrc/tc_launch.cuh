// tc_launch.cuh — host-side configuration + launch of the tcgen05 kernels (tensor maps, smem budget).
#pragma once
#include <cstring>

#include "tc_rows.cuh"
#include "tc_rows16.cuh"
#include "tc_wgrad.cuh"

namespace wire {

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

constexpr size_t kMaxDynSmem = 232448 - 6144;  // 227 KB minus the static barriers / exchange buffers

// cudaFuncAttributeMaxDynamicSharedMemorySize and cluster occupancy are per device: the launchers below cache them per
// (kernel instantiation, device), so a process that drives several GPUs sets the attribute on each of them.
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}

// Fill in nb-dependent fields and pick the deepest pipeline that fits. Returns dynamic smem bytes
// (0 if the configuration does not fit).
// `n_in`   : TMA-prefetched epilogue inputs (z / w tiles of the backward modes)
// `out_cols`: real output columns (2M) -> size of the shared-memory parameter tables
// `mode`    : decides which tables exist (bias / fused-final weights / first-layer table)
// `cluster` : CTAs that share each B tile through TMA multicast; nb must split into 8-row aligned shares
inline size_t rows_configure(RowsParams& P, int nb, int nbh, int store_mask, int n_in = 0, int out_cols = 0,
                             int mode = MODE_PLAIN, bool fuse_final = false, int cluster = 1, bool gen = false) {
  P.nb = nb;
  P.nbh = nbh;
  P.store_mask = store_mask;
  P.n_in = n_in;
  P.cluster = cluster;
  if (cluster != 1 && cluster != 2) return 0;
  if (nb <= 256) { P.slices = 1; P.ns = nb; P.buf_cols = 256; }
  else {
    // two N-slices, each with its own K loop and TMEM buffer (the MMAs of one overlap the epilogue of the other)
    if (nb % 64 || nb > 512) return 0;
    // wire2d forward: a slice is one column block [z half | w half] of nbh + nbh columns, so two slices need nb == 4 * nbh
    if (mode == MODE_GABOR2D_FWD && nb != 4 * nbh) return 0;
    P.slices = 2; P.ns = nb / 2; P.buf_cols = nb / 2;
  }
  if (P.ns % 16 || (P.ns / cluster) % 8) return 0;
  P.b_box_rows = P.ns / cluster;
  const int n_out = __builtin_popcount(store_mask);
  const size_t stage = size_t(kTileRows) * 128 + size_t(P.b_box_rows) * 128;
  const size_t staging = size_t(kEpiWarps) * (n_out + n_in) * 4096;
  const int pcols = round_up(out_cols > 0 ? out_cols : 32, 32) + 32;  // +1 chunk: the tail chunk may over-read
  const bool two_d = (mode == MODE_GABOR2D_FWD || mode == MODE_GABOR2D_BWD || mode == MODE_FIRST2D_BWD);
  size_t pfloats = 0;
  if (mode == MODE_GABOR_FWD || mode == MODE_GABOR2D_FWD) pfloats = size_t(two_d ? 2 : 1) * pcols + (fuse_final ? size_t(pcols / 2) * 8 : 0);
  if (mode == MODE_FIRST_BWD || mode == MODE_FIRST2D_BWD) pfloats = size_t(two_d ? 2 : 1) * (pcols / 2) * 4;
  P.gen_tab_off = uint32_t((pfloats + 3) / 4 * 4);
  if (gen) pfloats = P.gen_tab_off + size_t(2) * (pcols / 2) * 4;  // generator tables {w0[0..2], b0} (+ scale_orth)
  const size_t pbytes = (pfloats * sizeof(float) + 127) / 128 * 128;
  const size_t budget = kMaxDynSmem - 1024;
  if (staging + pbytes + 2 * stage > budget) return 0;
  int stages = int((budget - staging - pbytes) / stage);
  if (stages > 8) stages = 8;
  P.stages = stages;
  P.staging_off = uint32_t(stages * stage);
  P.param_off = uint32_t(stages * stage + staging);
  P.param_cols = pcols;
  return stages * stage + staging + pbytes + 1024;
}

template <int MODE, bool PAIR, bool GEN = false, bool OP16 = false>
inline cudaError_t launch_rows_mode_p(const RowsParams& P, size_t smem, int sm_count, cudaStream_t st) {
  static bool attr_set_dev[kMaxDevices] = {};
  static int max_clusters_dev[kMaxDevices][5] = {};
  const int dev = current_device_slot();
  bool& attr_set = attr_set_dev[dev];
  int* max_clusters = max_clusters_dev[dev];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_rows_kernel<MODE, PAIR, GEN, OP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxDynSmem));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int C = P.cluster;
  const int row_tiles = (P.e.n_rows + kTileRows - 1) / kTileRows;
  const int units = ((row_tiles + C - 1) / C) * P.n_blocks;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kRowsThreads + (GEN ? 32 * kGenWarps : 0));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = C > 1 ? 1 : 0;
  if (C > 1 && max_clusters[C] == 0) {
    // how many clusters of this size are co-resident (GPC boundaries strand SMs for larger clusters)
    cfg.gridDim = dim3(sm_count / C * C);
    cfg.dynamicSmemBytes = kMaxDynSmem;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, tc_rows_kernel<MODE, PAIR, GEN, OP16>, &cfg) != cudaSuccess || nc <= 0) nc = sm_count / C / 2;
    (void)cudaGetLastError();
    max_clusters[C] = nc;
    cfg.dynamicSmemBytes = smem;
  }
  int clusters = C > 1 ? max_clusters[C] : sm_count;
  if (clusters > units) clusters = units;
  if (clusters <= 0) return cudaSuccess;
  cfg.gridDim = dim3(clusters * C);
  return cudaLaunchKernelEx(&cfg, tc_rows_kernel<MODE, PAIR, GEN, OP16>, P);
}
template <int MODE>
inline cudaError_t launch_rows_mode(const RowsParams& P, size_t smem, int sm_count, cudaStream_t st, bool op16) {
  if (op16) return cudaErrorInvalidValue;  // 16-bit operands run tc_rows16_kernel (launch_rows16)
  return P.cluster == 2 ? launch_rows_mode_p<MODE, true>(P, smem, sm_count, st) : launch_rows_mode_p<MODE, false>(P, smem, sm_count, st);
}

// forward kernels whose A operand (the first layer's output) is generated in place
inline cudaError_t launch_rows_gen(int mode, const RowsParams& P, size_t smem, int sm_count, cudaStream_t st) {
  if (P.e.n_rows <= 0) return cudaSuccess;
  const bool pair = P.cluster == 2;
  if (mode == MODE_GABOR_FWD)
    return pair ? launch_rows_mode_p<MODE_GABOR_FWD, true, true>(P, smem, sm_count, st) : launch_rows_mode_p<MODE_GABOR_FWD, false, true>(P, smem, sm_count, st);
  if (mode == MODE_GABOR2D_FWD)
    return pair ? launch_rows_mode_p<MODE_GABOR2D_FWD, true, true>(P, smem, sm_count, st) : launch_rows_mode_p<MODE_GABOR2D_FWD, false, true>(P, smem, sm_count, st);
  return cudaErrorInvalidValue;
}

// op16: A / B are 16-bit tensors (P.a_fmt / P.b_fmt), tensor maps built with 64-column boxes
inline cudaError_t launch_rows(int mode, const RowsParams& P, size_t smem, int sm_count, cudaStream_t st, bool op16 = false) {
  if (P.e.n_rows <= 0) return cudaSuccess;
  switch (mode) {
    case MODE_PLAIN: return launch_rows_mode<MODE_PLAIN>(P, smem, sm_count, st, op16);
    case MODE_GABOR_FWD: return launch_rows_mode<MODE_GABOR_FWD>(P, smem, sm_count, st, op16);
    case MODE_GABOR2D_FWD: return launch_rows_mode<MODE_GABOR2D_FWD>(P, smem, sm_count, st, op16);
    case MODE_GABOR_BWD: return launch_rows_mode<MODE_GABOR_BWD>(P, smem, sm_count, st, op16);
    case MODE_GABOR2D_BWD: return launch_rows_mode<MODE_GABOR2D_BWD>(P, smem, sm_count, st, op16);
    case MODE_FIRST_BWD: return launch_rows_mode<MODE_FIRST_BWD>(P, smem, sm_count, st, op16);
    case MODE_FIRST2D_BWD: return launch_rows_mode<MODE_FIRST2D_BWD>(P, smem, sm_count, st, op16);
  }
  return cudaErrorInvalidValue;
}

// programmatic dependent launch for the kernels that call sm100::pdl_wait() (WIRE_B200_PDL=0 turns it off)
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("WIRE_B200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
inline void add_pdl_attr(cudaLaunchAttribute* attr, unsigned& n) {
  if (!pdl_enabled()) return;
  attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[n].val.programmaticStreamSerializationAllowed = 1;
  ++n;
}

// ---- 16-bit row-tile kernels (tc_rows16.cuh): 16 epilogue warps, 2 KB staging tiles, final-Linear exchange in dynamic smem ----
// B-stationary schedule (RowsParams::bstat): WIRE_B200_BSTAT = 0 off, 1 (default) where it pays, 2 wherever it fits
inline int bstat_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("WIRE_B200_BSTAT"); v = e ? atoi(e) : 1; if (v < 0 || v > 2) v = 1; }
  return v;
}
// k_stages: 64-column K stages of the GEMM (0 = unknown); n_blocks: column blocks of the job (0 = unknown)
inline size_t rows16_configure(RowsParams& P, int nb, int nbh, int store_mask, int n_in = 0, int out_cols = 0,
                               int mode = MODE_PLAIN, bool fuse_final = false, int cluster = 1, int k_stages = 0, int n_blocks = 0) {
  P.nb = nb;
  P.nbh = nbh;
  P.store_mask = store_mask;
  P.n_in = n_in;
  P.cluster = cluster;
  if (cluster != 1 && cluster != 2) return 0;
  if (nb <= 256) { P.slices = 1; P.ns = nb; P.buf_cols = 256; }
  else {
    if (nb % 64 || nb > 512) return 0;
    // wire2d forward: a slice is one column block [z half | w half] of nbh + nbh columns, so two slices need nb == 4 * nbh
    if (mode == MODE_GABOR2D_FWD && nb != 4 * nbh) return 0;
    P.slices = 2; P.ns = nb / 2; P.buf_cols = nb / 2;
  }
  if (P.ns % 16 || (P.ns / cluster) % 8) return 0;
  P.b_box_rows = P.ns / cluster;
  const int n_out = __builtin_popcount(store_mask);
  const size_t stage = size_t(kTileRows) * 128 + size_t(P.b_box_rows) * 128;
  const size_t staging = size_t(kEpi16Warps) * (n_out + n_in) * kTile16Bytes;
  const int pcols = round_up(out_cols > 0 ? out_cols : 32, 32) + 32;
  const bool two_d = (mode == MODE_GABOR2D_FWD || mode == MODE_GABOR2D_BWD || mode == MODE_FIRST2D_BWD);
  size_t pfloats = 0;
  if (mode == MODE_GABOR_FWD || mode == MODE_GABOR2D_FWD)
    pfloats = size_t(two_d ? 2 : 1) * pcols + (fuse_final ? size_t(pcols / 2) * 8 + size_t(2) * (kEpi16Parts - 1) * kTileRows * 4 : 0);
  if (mode == MODE_FIRST_BWD || mode == MODE_FIRST2D_BWD) pfloats = size_t(two_d ? 2 : 1) * (pcols / 2) * 4;
  const size_t pbytes = (pfloats * sizeof(float) + 127) / 128 * 128;
  const size_t budget = kMaxDynSmem - 1024;
  P.bstat = 0;
  P.bres_off = 0;
  // Measured (profiles/r02_probe_bstat.log): with staging tiles in the budget only 3 A stages fit and the schedule is a wash
  // (forward 154.6 vs 156.4 us, dgrad + Gabor backward 146 vs 140 us); without staging (the first-layer dgrad: 7 A stages)
  // it wins 7 us.  WIRE_B200_BSTAT=2 forces it for every eligible job.
  const bool bstat_ok = bstat_mode() == 2 || (bstat_mode() == 1 && n_out + n_in == 0);
  if (P.slices == 2 && !fuse_final && n_blocks == 1 && k_stages > 0 && bstat_ok) {
    // B-stationary: the slice's packed weights stay resident (k_stages tiles of b_box_rows x 128 B), the stages hold A only
    const size_t a_stage = size_t(kTileRows) * 128;
    const size_t bres = size_t(k_stages) * P.b_box_rows * 128;
    if (staging + pbytes + bres + 3 * a_stage <= budget) {
      int stages = int((budget - staging - pbytes - bres) / a_stage);
      if (stages > 8) stages = 8;
      P.bstat = 1;
      P.stages = stages;
      P.bres_off = uint32_t(stages * a_stage);
      P.staging_off = uint32_t(stages * a_stage + bres);
      P.param_off = uint32_t(stages * a_stage + bres + staging);
      P.param_cols = pcols;
      return stages * a_stage + bres + staging + pbytes + 1024;
    }
  }
  if (staging + pbytes + 2 * stage > budget) return 0;
  int stages = int((budget - staging - pbytes) / stage);
  if (stages > 8) stages = 8;
  P.stages = stages;
  P.staging_off = uint32_t(stages * stage);
  P.param_off = uint32_t(stages * stage + staging);
  P.param_cols = pcols;
  return stages * stage + staging + pbytes + 1024;
}

template <int MODE, bool PAIR, int FUSE, bool SCAL = false>
inline cudaError_t launch_rows16_p(const RowsParams& P, size_t smem, int sm_count, cudaStream_t st) {
  static bool attr_set_dev[kMaxDevices] = {};
  static int max_clusters_dev[kMaxDevices] = {};
  const int dev = current_device_slot();
  bool& attr_set = attr_set_dev[dev];
  int& max_clusters = max_clusters_dev[dev];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_rows16_kernel<MODE, PAIR, FUSE, SCAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxDynSmem));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  constexpr int C = PAIR ? 2 : 1;
  const int row_tiles = (P.e.n_rows + kTileRows - 1) / kTileRows;
  const int units = ((row_tiles + C - 1) / C) * P.n_blocks;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kRows16Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = C > 1 ? 1 : 0;
  if (C > 1 && max_clusters == 0) {
    cfg.gridDim = dim3(sm_count / C * C);
    cfg.dynamicSmemBytes = kMaxDynSmem;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, tc_rows16_kernel<MODE, PAIR, FUSE, SCAL>, &cfg) != cudaSuccess || nc <= 0) nc = sm_count / C / 2;
    (void)cudaGetLastError();
    max_clusters = nc;
    cfg.dynamicSmemBytes = smem;
  }
  int clusters = C > 1 ? max_clusters : sm_count;
  if (P.bstat) {  // every cluster owns one slice: a multiple of the slice count, at most one cluster per (row group, slice)
    if (clusters > units * P.slices) clusters = units * P.slices;
    clusters -= clusters % P.slices;
    if (clusters < P.slices) return cudaErrorInvalidConfiguration;
  } else if (clusters > units) clusters = units;
  if (clusters <= 0) return cudaSuccess;
  cfg.gridDim = dim3(clusters * C);
  add_pdl_attr(attr, cfg.numAttrs);
  return cudaLaunchKernelEx(&cfg, tc_rows16_kernel<MODE, PAIR, FUSE, SCAL>, P);
}
template <int MODE, int FUSE = 0>
inline cudaError_t launch_rows16_mode(const RowsParams& P, size_t smem, int sm_count, cudaStream_t st) {
  return P.cluster == 2 ? launch_rows16_p<MODE, true, FUSE>(P, smem, sm_count, st) : launch_rows16_p<MODE, false, FUSE>(P, smem, sm_count, st);
}
// backward modes: with gradient slots for the epilogue layer's own omega_0 / scale_0 the SCAL instantiation runs (CTA pairs only)
template <int MODE>
inline cudaError_t launch_rows16_bwd(const RowsParams& P, size_t smem, int sm_count, cudaStream_t st) {
  if (P.e.g_omega || P.e.g_scale) {
    if (P.cluster != 2) return cudaErrorNotSupported;
    return launch_rows16_p<MODE, true, 0, true>(P, smem, sm_count, st);
  }
  return launch_rows16_mode<MODE>(P, smem, sm_count, st);
}
// P configured by rows16_configure; 16-bit tensor maps: A/B boxes of 64 columns (SWIZZLE_128B), output / saved-z tiles
// of 32x32 elements with SWIZZLE_64B
inline cudaError_t launch_rows16(int mode, const RowsParams& P, size_t smem, int sm_count, cudaStream_t st) {
  if (P.e.n_rows <= 0) return cudaSuccess;
  const bool fuse = P.e.fuse_final != 0;
  const char* f4 = getenv("WIRE_B200_FUSE4");   // =1: the four-output epilogue whatever the output count (A/B runs)
  const bool three = P.e.out_features <= 3 && !(f4 && f4[0] == '1');
  switch (mode) {
    case MODE_PLAIN: return launch_rows16_mode<MODE_PLAIN>(P, smem, sm_count, st);
    case MODE_GABOR_FWD:
      if (!fuse) return launch_rows16_mode<MODE_GABOR_FWD>(P, smem, sm_count, st);
      return three ? launch_rows16_mode<MODE_GABOR_FWD, 3>(P, smem, sm_count, st) : launch_rows16_mode<MODE_GABOR_FWD, 4>(P, smem, sm_count, st);
    case MODE_GABOR2D_FWD:
      if (!fuse) return launch_rows16_mode<MODE_GABOR2D_FWD>(P, smem, sm_count, st);
      return three ? launch_rows16_mode<MODE_GABOR2D_FWD, 3>(P, smem, sm_count, st) : launch_rows16_mode<MODE_GABOR2D_FWD, 4>(P, smem, sm_count, st);
    case MODE_GABOR_BWD: return launch_rows16_bwd<MODE_GABOR_BWD>(P, smem, sm_count, st);
    case MODE_GABOR2D_BWD: return launch_rows16_bwd<MODE_GABOR2D_BWD>(P, smem, sm_count, st);
    case MODE_FIRST_BWD: return launch_rows16_bwd<MODE_FIRST_BWD>(P, smem, sm_count, st);
    case MODE_FIRST2D_BWD: return launch_rows16_bwd<MODE_FIRST2D_BWD>(P, smem, sm_count, st);
  }
  return cudaErrorInvalidValue;
}

// wgrad: choose column blocking, K splits (to fill the machine) and pipeline depth
inline size_t wgrad_configure(WgradParams& P, int sm_count, int cluster = 2, bool gen = false, bool op16 = false) {
  const int x_cols = 2 * P.k_in + (P.bias_sum ? 0 : 1);   // bias_sum: no "ones" column in the x tiles
  P.cluster = cluster;
  P.m_tiles = (x_cols + 128 * cluster - 1) / (128 * cluster);
  // a pair splits every MMA piece in block-aligned halves (blocks: 32 columns, or 64 for 16-bit operands); the
  // accumulator of one work item may use all 512 TMEM columns
  const int blk_cols = op16 ? 64 : 32;
  // 16-bit pair: a piece's per-CTA half may end in half a block (32 columns): 2M = 424 then runs as 256 + 192 = 448 accumulator
  // columns instead of 512 (ncu: the tensor pipe is 70 % busy in this kernel, so padded columns are real time)
  const int gran = cluster == 2 ? (op16 ? blk_cols : 2 * blk_cols) : blk_cols;
  const int nb_max = op16 ? 512 : 448;
  const int gpad = round_up(P.g_cols, gran);
  // dual (tc_wgrad.cuh): both g tensors of a wire2d layer in one work item when each pads to exactly one 256-column MMA piece
  P.dual = (op16 && cluster == 2 && P.bias_sum && P.n_g == 2 && round_up(P.g_cols, 64) == 256 && P.dual) ? 1 : 0;
  if (P.dual) {
    P.n_g = 1; P.n_blocks = 1; P.nb = 512;
  } else {
    P.n_blocks = (gpad + nb_max - 1) / nb_max;
    P.nb = round_up((gpad + P.n_blocks - 1) / P.n_blocks, gran);
    if (P.nb > nb_max) { P.n_blocks += 1; P.nb = round_up((gpad + P.n_blocks - 1) / P.n_blocks, gran); }
  }
  const int base = P.m_tiles * P.n_blocks * P.n_g;
  int splits = (sm_count / cluster) / base;
  if (splits < 1) splits = 1;
  const int kc = op16 ? kWgradKC16 : kWgradKC;
  const int total_chunks = (P.n_rows + kc - 1) / kc;
  if (splits > total_chunks) splits = total_chunks > 0 ? total_chunks : 1;
  P.splits = splits;
  // x: 128 columns per CTA; g: nb / cluster columns per CTA; one block = blk_cols columns x (32 | 64) rows x (4 | 2) bytes
  const int n1c = (P.nb > 256 ? 256 : P.nb) / cluster, n2c = (P.nb > 256 ? P.nb - 256 : 0) / cluster;   // per-CTA halves of the MMA pieces
  const size_t stage = size_t(128 / blk_cols + (n1c + blk_cols - 1) / blk_cols + (n2c + blk_cols - 1) / blk_cols) * (op16 ? 8192 : 4096);
  P.gen_tab_feats = round_up(P.k_in + 1, 16) + 64 * 4;  // covers every feature index a generator warp may touch
  const size_t tab_bytes = gen ? size_t(2) * P.gen_tab_feats * 16 : 0;
  int stages = int((kMaxDynSmem - 1024 - tab_bytes) / stage);
  if (stages > 8) stages = 8;
  if (stages < 2) return 0;
  P.stages = stages;
  P.gen_tab_off = uint32_t(stages * stage);
  return stages * stage + tab_bytes + 1024;
}

template <bool PAIR, bool GEN = false, bool OP16 = false>
inline cudaError_t launch_wgrad_p(const WgradParams& P, size_t smem, cudaStream_t st) {
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<PAIR, GEN, OP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxDynSmem));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int C = PAIR ? 2 : 1;
  const int grid = P.m_tiles * P.n_blocks * P.n_g * P.splits * C;
  if (grid <= 0 || P.n_rows <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(wgrad_threads(GEN, OP16));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = PAIR ? 1 : 0;
  add_pdl_attr(attr, cfg.numAttrs);
  return cudaLaunchKernelEx(&cfg, tc_wgrad_kernel<PAIR, GEN, OP16>, P);
}
inline cudaError_t launch_wgrad(const WgradParams& P, size_t smem, cudaStream_t st, bool gen = false, bool op16 = false) {
  if (op16) return P.cluster == 2 ? launch_wgrad_p<true, false, true>(P, smem, st) : launch_wgrad_p<false, false, true>(P, smem, st);
  if (gen) return P.cluster == 2 ? launch_wgrad_p<true, true>(P, smem, st) : launch_wgrad_p<false, true>(P, smem, st);
  return P.cluster == 2 ? launch_wgrad_p<true>(P, smem, st) : launch_wgrad_p<false>(P, smem, st);
}

}  // namespace wire
