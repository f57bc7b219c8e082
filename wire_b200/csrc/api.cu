// api.cu — the extern "C" boundary (include/wire_b200.h): workspace layout and kernel sequencing
// for the WIRE forward/backward pass.  No torch, no allocation, no synchronisation.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>

#include "../../include/wire_b200.h"
#include "simt_kernels.cuh"
#include "simt_rows.cuh"
#include "tc_launch.cuh"
#include "simt16_kernels.cuh"
#include "peer_kernels.cuh"
#include "data_kernels.cuh"

namespace {

using namespace wire;

thread_local char g_err[512] = "";
int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
#define CU_OK(expr)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define TRY(expr)          \
  do {                     \
    int r_ = (expr);       \
    if (r_) return r_;     \
  } while (0)


// <<<>>> with the programmatic-dependent-launch attribute (tc_launch.cuh: add_pdl_attr) for kernels that call sm100::pdl_wait()
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr; cfg.numAttrs = 0;
  add_pdl_attr(attr, cfg.numAttrs);
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// Per-device facts, cached per device (a process may drive several GPUs: `with torch.cuda.device(...)`); g_sm_count / g_cc_major
// are the CURRENT device's values for the calling thread, refreshed by every entry point through require_device().
struct DevInfo { int sm_count = 0; int cc_major = -1; };
DevInfo g_dev[kMaxDevices];
thread_local int g_sm_count = 0;
thread_local int g_cc_major = -1;
int device_info() {
  int dev = 0;
  CU_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail("device index %d outside 0..%d", dev, kMaxDevices - 1);
  if (g_dev[dev].cc_major < 0) {
    int sm = 0, cc = 0;
    CU_OK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    CU_OK(cudaDeviceGetAttribute(&cc, cudaDevAttrComputeCapabilityMajor, dev));
    g_dev[dev].sm_count = sm;
    g_dev[dev].cc_major = cc;
  }
  g_sm_count = g_dev[dev].sm_count;
  g_cc_major = g_dev[dev].cc_major;
  return 0;
}
// bound of the in-kernel peer spins in SM cycles (peer_kernels.cuh): WIRE_B200_PEER_TIMEOUT_S seconds at ~2 GHz, default 600 s
long long peer_spin_limit() {
  static long long v = 0;
  if (v == 0) {
    const char* e = getenv("WIRE_B200_PEER_TIMEOUT_S");
    double s = e ? atof(e) : 600.0;
    if (!(s > 0.0)) s = 600.0;
    v = (long long)(s * 2.0e9);
  }
  return v;
}
int require_device() {
  TRY(device_info());
  if (g_cc_major != 10) return fail("wire_b200 needs an sm_100 (Blackwell B200) device, found compute capability %d.x; there is no fallback path", g_cc_major);
  return 0;
}

// ------------------------------------------------------------------------------------------
// launch accounting: every kernel launch is counted; optionally bracketed by CUDA events recorded
// on the launching stream (bench.py's per-kernel roofline numbers come from here).
// ------------------------------------------------------------------------------------------
enum ProfKind { K_FIRST_FWD = 0, K_PACK, K_ROWS_FWD, K_FINAL_FWD, K_TOP_BWD, K_WGRAD, K_ROWS_BWD, K_ROWS_FIRST_BWD,
                K_FIRST_WGRAD, K_GRAD_COORDS, K_LAYER_MISC, K_ADAM, K_MSE, K_PEER_WAIT, K_DATA, K_COUNT };
const char* kProfNames[K_COUNT] = {"first_fwd", "pack_weights", "tc_rows_gabor_fwd", "final_fwd", "top_bwd", "tc_wgrad",
                                   "tc_rows_dgrad_gabor_bwd", "tc_rows_dgrad_first_bwd", "first_wgrad", "grad_coords",
                                   "layer_misc", "adam", "mse_grad", "peer_wait", "data_pipeline"};
// (one process-wide table guarded by a mutex: the entry points may be called from several host threads)
struct ProfPending { int kind; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
struct Prof {
  int timing = 0;
  unsigned long long launches[K_COUNT] = {};
  double ms[K_COUNT] = {};
  ProfPending pending[4096];
  int n_pending = 0;
} g_prof;
struct ProfScope {
  int kind; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr; bool on = false;
  ProfScope(int k, cudaStream_t s, int n_launches = 1) : kind(k), st(s) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.launches[k] += n_launches;
    if (g_prof.timing) {
      if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) {
        cudaEventRecord(e0, st);
        on = true;
      }
    }
  }
  ~ProfScope() {  // the pair is published only once both events are recorded (scopes of several threads may interleave)
    if (!on) return;
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof.n_pending < 4096) g_prof.pending[g_prof.n_pending++] = {kind, e0, e1};
    else { cudaEventDestroy(e0); cudaEventDestroy(e1); }
  }
};
int rows_kind(int mode) {
  switch (mode) {
    case MODE_GABOR_FWD: case MODE_GABOR2D_FWD: return K_ROWS_FWD;
    case MODE_GABOR_BWD: case MODE_GABOR2D_BWD: return K_ROWS_BWD;
    case MODE_FIRST_BWD: case MODE_FIRST2D_BWD: return K_ROWS_FIRST_BWD;
  }
  return K_LAYER_MISC;
}

// CTAs per cluster sharing weight tiles through TMA multicast (WIRE_B200_CLUSTER=1|2|4 overrides)
int cluster_size() {
  static int c = 0;
  if (c == 0) {
    const char* e = getenv("WIRE_B200_CLUSTER");
    c = e ? atoi(e) : 2;
    if (c != 1 && c != 2) c = 2;
  }
  return c;
}

// WIRE_B200_SECTOR_ALIGN=0: stored tensors end at column 2M again (A/B runs; see run_rows_job)
bool store_sector_align() {
  const char* e = getenv("WIRE_B200_SECTOR_ALIGN");
  return !(e && e[0] == '0');
}

constexpr int64_t kInferChunk = 1 << 19;  // rows per pass when nothing has to be kept for backward
constexpr int kRowsPerBlock = 64;

// ------------------------------------------------------------------------------------------
// column blocking of a row-tile GEMM
// ------------------------------------------------------------------------------------------
struct Blocking {
  int n_blocks, nb, nbh;
};
// out_cols = real output columns (2M). two_d_fwd: accumulator holds [z half | w half].
bool choose_blocking(int out_cols, bool two_d_fwd, int store_mask, Blocking& b, int n_in = 0, int mode = MODE_PLAIN,
                     bool fuse_final = false, bool gen = false, bool op16 = false) {
  for (int nblk = 1; nblk <= 64; ++nblk) {
    const int per = (out_cols + nblk - 1) / nblk;
    const int C = cluster_size();
    int nbh, nb;
    if (two_d_fwd) { nbh = round_up(per, 32); nb = 2 * nbh; if (nb > 256) continue; }
    else {
      nb = round_up(per, nblk == 1 ? 16 : 32);
      if (nb > 256) nb = round_up(per, 64);
      nbh = nb;
    }
    RowsParams tmp;
    // (hidden layers are square: the GEMM's K is out_cols per A part; the wire2d dgrad has two parts)
    const int k_stages = ((mode == MODE_GABOR2D_BWD || mode == MODE_FIRST2D_BWD) ? 2 : 1) * ((out_cols + 63) / 64);
    if (op16 ? rows16_configure(tmp, nb, nbh, store_mask, n_in, out_cols, mode, fuse_final, C, k_stages, nblk) == 0
             : rows_configure(tmp, nb, nbh, store_mask, n_in, out_cols, mode, fuse_final, C, gen) == 0) continue;
    b.n_blocks = nblk; b.nb = nb; b.nbh = nbh;
    return true;
  }
  return false;
}

// wire2d forward on the 16-bit path: two column blocks of [z half | w half] = 128 + 128 accumulator columns (M = 128, the
// SISR width) are run as the two N-SLICES of one work unit (same packed-weight layout, TMEM 2 x 256 columns), so the whole
// output row stays in one CTA and the final Linear can be fused into the last layer's epilogue like for wire.
bool merge_2d_blocks(const Blocking& b, bool two_d, bool op16) { return two_d && op16 && b.n_blocks == 2 && b.nb == 256 && b.nbh == 128; }

// ------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------
struct Layout {
  int M, H, two_m, P, PR, k_pad, two_d;
  int prec;                 // effective precision (net_precision)
  int y_elem, z_elem, g_elem;  // element types of the y / saved z,w / g_z,g_w tensors (sm100_host::ElemType)
  int64_t rows;
  int n_act;  // y buffers
  size_t off_y[WIRE_B200_MAX_LAYERS + 1], off_z[WIRE_B200_MAX_LAYERS + 1], off_w[WIRE_B200_MAX_LAYERS + 1];
  size_t off_gz[2], off_gw[2], off_gz0, off_gw0;
  size_t off_bf[WIRE_B200_MAX_LAYERS + 1], off_bd[WIRE_B200_MAX_LAYERS + 1];
  size_t pack_floats;
  size_t total;
  bool fuse_final;
};

int check_desc(const wire_net_desc* d) {
  if (!d) return fail("null descriptor");
  if (d->width < 1 || d->width > 1024) return fail("width %d out of range [1,1024]", d->width);
  if (d->hidden_layers < 1 || d->hidden_layers >= WIRE_B200_MAX_LAYERS) return fail("hidden_layers %d out of range [1,%d)", d->hidden_layers, WIRE_B200_MAX_LAYERS);
  if (d->in_features < 1 || d->in_features > kMaxIn) return fail("in_features %d out of range [1,%d]", d->in_features, kMaxIn);
  if (d->out_features < 1 || d->out_features > kSimtMaxOut) return fail("out_features %d out of range [1,%d]", d->out_features, kSimtMaxOut);
  if (d->precision != WIRE_PRECISION_TF32 && d->precision != WIRE_PRECISION_FP32 && d->precision != WIRE_PRECISION_MIXED16)
    return fail("unknown precision %d", d->precision);
  return 0;
}

// MIXED16 covers the whole-network path for the shapes every reference driver uses (coordinates <= 3-D, <= 4 outputs);
// anything else, and the single-layer entry points, run the TF32 kernels (same GPU, same data flow, 32-bit operands).
int net_precision(const wire_net_desc* d) {
  if (d->precision == WIRE_PRECISION_MIXED16 && (d->in_features > 3 || d->out_features > 4)) return WIRE_PRECISION_TF32;
  return d->precision;
}
// single-layer entry points: a hidden layer under MIXED16 runs the 16-bit kernels (layer16_ok); everything else that asks for
// MIXED16 (the first layer, the stand-alone final Linear) runs the 32-bit-operand kernels
wire_net_desc layer_desc(const wire_net_desc* d) {
  wire_net_desc dd = *d;
  if (dd.precision == WIRE_PRECISION_MIXED16) dd.precision = WIRE_PRECISION_TF32;
  return dd;
}
bool layer16_ok(const wire_net_desc* d, int is_first) { return d->precision == WIRE_PRECISION_MIXED16 && !is_first; }
using sm100_host::kElemBF16;
using sm100_host::kElemF16;
using sm100_host::kElemF32;

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int make_layout(const wire_net_desc* d, int64_t n, int training, Layout& L) {
  TRY(check_desc(d));
  memset(&L, 0, sizeof(L));
  L.M = d->width; L.H = d->hidden_layers; L.two_m = 2 * d->width; L.two_d = d->two_d;
  L.prec = net_precision(d);
  const bool mixed = L.prec == WIRE_PRECISION_MIXED16;
  L.y_elem = mixed ? kElemF16 : kElemF32;                               // activations: FP16 (TF32's 11-bit significand)
  L.z_elem = L.prec == WIRE_PRECISION_FP32 ? kElemF32 : kElemF16;       // saved pre-activations
  L.g_elem = mixed ? kElemBF16 : kElemF32;                              // gradients: BF16 (FP32's range, no loss scaling)
  L.P = round_up(L.two_m + 1, 32);
  L.PR = round_up(L.M, mixed ? 8 : 4);   // real g_z0 rows: fp32 (16-byte aligned), or BF16 on the mixed16 path
  L.k_pad = round_up(L.two_m, mixed ? 64 : 32);
  L.rows = training ? n : (n < kInferChunk ? n : kInferChunk);
  if (L.rows < 1) L.rows = 1;
  Blocking bf;
  if (!choose_blocking(L.two_m, d->two_d != 0, d->two_d ? 6 : 2, bf, 0, d->two_d ? MODE_GABOR2D_FWD : MODE_GABOR_FWD, true, false, mixed))
    return fail("no tile configuration for width %d", d->width);
  L.fuse_final = (bf.n_blocks == 1 || merge_2d_blocks(bf, d->two_d != 0, mixed)) && d->out_features <= kMaxOut;
  size_t off = 0;
  auto act = [&](int elem) { return align_up(size_t(L.rows) * L.P * sm100_host::elem_bytes(elem), 1024); };
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  if (training) {
    const int ny = L.fuse_final ? L.H : L.H + 1;
    for (int l = 0; l < ny; ++l) L.off_y[l] = take(act(L.y_elem));
    L.n_act = ny;
    for (int l = 1; l <= L.H; ++l) { L.off_z[l] = take(act(L.z_elem)); if (d->two_d) L.off_w[l] = take(act(L.z_elem)); }
    for (int i = 0; i < 2; ++i) { L.off_gz[i] = take(act(L.g_elem)); if (d->two_d) L.off_gw[i] = take(act(L.g_elem)); }
    L.off_gz0 = take(size_t(L.rows) * L.PR * sm100_host::elem_bytes(L.g_elem));
    if (d->two_d) L.off_gw0 = take(size_t(L.rows) * L.PR * sm100_host::elem_bytes(L.g_elem));
  } else {
    L.off_y[0] = take(act(L.y_elem)); L.off_y[1] = take(act(L.y_elem)); L.n_act = 2;
  }
  // packed weights: generous bound on rows (blocks are padded to 32) x both K parts
  L.pack_floats = size_t(2 * L.k_pad + 1024) * size_t(2 * L.k_pad);
  for (int l = 1; l <= L.H; ++l) {
    L.off_bf[l] = take(L.pack_floats * sizeof(float));
    if (training) L.off_bd[l] = take(L.pack_floats * sizeof(float));
  }
  L.total = off;
  return 0;
}

inline float* at(void* ws, size_t off) { return reinterpret_cast<float*>(static_cast<char*>(ws) + off); }

// ------------------------------------------------------------------------------------------
// generic row-tile GEMM job -> tcgen05 (TF32) or CUDA-core (FP32) kernel
// ------------------------------------------------------------------------------------------
struct RowsJob {
  int mode;
  const float* a[2];
  int a_pitch[2];
  int k_cols[2];
  const float* b;
  int b_rows, b_pitch, k0_pad;
  Blocking blk;
  float* o[3];
  int o_pitch[3];
  int o_half[3];  // element type of each output tensor (sm100_host::ElemType; non-zero = 16-bit: FP16 z / y, BF16 g)
  int a_elem, b_elem;  // element type of the A / packed-B operands (0 = TF32 kernels, FP16 or BF16 = kind::f16 kernels)
  int store_mask;
  int gen;        // A operand = first-layer output, generated in place (e.coords / e.w0 ... describe that layer)
  const float* gen_omega;
  const float* gen_scale;
  int gen_two_d;
  RowsEpi e;
};
int job_n_in(int mode) { return mode == MODE_GABOR_BWD ? 1 : (mode == MODE_GABOR2D_BWD ? 2 : 0); }
bool job_blocking(int mode, int out_cols, int store_mask, bool fuse_final, Blocking& b, bool gen = false, bool op16 = false) {
  return choose_blocking(out_cols, mode == MODE_GABOR2D_FWD, store_mask, b, job_n_in(mode), mode, fuse_final, gen, op16);
}

template <int MODE>
int launch_simt_rows(const SimtRowsParams& S, cudaStream_t st) {
  const int grid = (S.e.n_rows + 127) / 128;
  if (grid <= 0) return 0;
  simt_rows_kernel<MODE><<<grid, 128, 0, st>>>(S);
  CU_OK(cudaGetLastError());
  return 0;
}

int run_rows(const RowsJob& J, int precision, cudaStream_t st) {
  if (J.e.n_rows <= 0) return 0;
  ProfScope prof(rows_kind(J.mode), st);
  // the tcgen05 first-layer epilogue keeps a {w0[0..2], b0} table: wider coordinates use the FP32 kernel
  if ((J.mode == MODE_FIRST_BWD || J.mode == MODE_FIRST2D_BWD) && J.e.in_features > 3) precision = WIRE_PRECISION_FP32;
  if (precision == WIRE_PRECISION_FP32) {
    SimtRowsParams S;
    memset(&S, 0, sizeof(S));
    for (int i = 0; i < 2; ++i) { S.a[i] = J.a[i]; S.a_pitch[i] = J.a_pitch[i]; S.k_cols[i] = J.k_cols[i]; }
    S.b = J.b; S.b_pitch = J.b_pitch; S.k0_pad = J.k0_pad;
    for (int i = 0; i < 3; ++i) { S.o[i] = J.o[i]; S.o_pitch[i] = J.o_pitch[i]; }
    S.n_blocks = J.blk.n_blocks; S.nb = J.blk.nb; S.nbh = J.blk.nbh; S.store_mask = J.store_mask;
    S.e = J.e;
    switch (J.mode) {
      case MODE_PLAIN: return launch_simt_rows<MODE_PLAIN>(S, st);
      case MODE_GABOR_FWD: return launch_simt_rows<MODE_GABOR_FWD>(S, st);
      case MODE_GABOR2D_FWD: return launch_simt_rows<MODE_GABOR2D_FWD>(S, st);
      case MODE_GABOR_BWD: return launch_simt_rows<MODE_GABOR_BWD>(S, st);
      case MODE_GABOR2D_BWD: return launch_simt_rows<MODE_GABOR2D_BWD>(S, st);
      case MODE_FIRST_BWD: return launch_simt_rows<MODE_FIRST_BWD>(S, st);
      case MODE_FIRST2D_BWD: return launch_simt_rows<MODE_FIRST2D_BWD>(S, st);
    }
    return fail("bad rows mode %d", J.mode);
  }
  RowsParams P;
  memset(&P, 0, sizeof(P));
  P.k_cols[0] = J.k_cols[0]; P.k_cols[1] = J.k_cols[1];
  P.n_blocks = J.blk.n_blocks;
  P.e = J.e;
  const bool op16 = J.a_elem != kElemF32;
  int nb = J.blk.nb;
  if (J.mode == MODE_GABOR2D_FWD && J.e.fuse_final && merge_2d_blocks(J.blk, true, op16)) { P.n_blocks = 1; nb = 2 * J.blk.nb; }
  const size_t smem = op16 ? rows16_configure(P, nb, J.blk.nbh, J.store_mask, job_n_in(J.mode), J.e.n_cols, J.mode, J.e.fuse_final != 0,
                                              cluster_size(), (J.k_cols[0] + 63) / 64 + (J.k_cols[1] + 63) / 64, P.n_blocks)
                           : rows_configure(P, J.blk.nb, J.blk.nbh, J.store_mask, job_n_in(J.mode), J.e.n_cols, J.mode, J.e.fuse_final != 0,
                                            cluster_size(), J.gen != 0);
  P.gen_omega = J.gen_omega; P.gen_scale = J.gen_scale; P.gen_two_d = J.gen_two_d;
  if (!smem) return fail("row-tile configuration does not fit shared memory (nb=%d)", J.blk.nb);
  bool ok = true;
  if (op16 && (J.b_elem != J.a_elem || J.gen)) return fail("16-bit row-tile GEMM needs A and B in the same format");
  P.a_fmt = J.a_elem == kElemBF16 ? int(sm100::kFmtBF16) : int(sm100::kFmtF16);
  P.l2_prefetch = getenv("WIRE_B200_L2PF") ? 1 : 0;  // measured: no gain (0.165 vs 0.154 ms), the fill path is L2->SM bound
  P.b_fmt = P.a_fmt;
  // sweep direction: reversing the last forward layer so that it starts on the rows the previous layer wrote last (still in
  // L2) took 7 us off that launch under per-kernel timing but nothing off the captured step (1.0657 vs 1.0658 ms): off.
  P.reverse = (op16 && getenv("WIRE_B200_REV") && J.e.fuse_final) ? 1 : 0;
  const int kbox = op16 ? 64 : 32;  // one 128-byte swizzle row of K columns
  for (int i = 0; i < 2; ++i) {
    const int src = (J.k_cols[i] > 0) ? i : 0;
    ok &= sm100_host::make_tmap_2d_t(&P.a_map[i], J.a[src], J.e.n_rows, J.k_cols[src], J.a_pitch[src], 128, kbox, CU_TENSOR_MAP_SWIZZLE_128B, J.a_elem);
  }
  ok &= sm100_host::make_tmap_2d_t(&P.b_map, J.b, J.b_rows, J.b_pitch, J.b_pitch, P.b_box_rows, kbox, CU_TENSOR_MAP_SWIZZLE_128B, J.b_elem);
  int nslot = 0;
  for (int bit = 0; bit < 3; ++bit)
    if (J.store_mask & (1 << bit)) {
      P.o_fmt[nslot] = J.o_half[nslot];
      if (op16 && !J.o_half[nslot]) return fail("16-bit row-tile kernels store 16-bit tensors only");
      if (J.o_half[nslot]) {  // 16-bit tiles: dense for the TF32 kernels' saved z, 64 B swizzle for the 16-bit kernels
        // 16-bit kernels: the stored width is rounded up to whole 32-byte sectors (16 columns) when the row pitch has room.  A row
        // that ends inside a sector (2M = 424 columns = 848 B = 26.5 sectors) makes every row's last sector a read-modify-write in
        // DRAM: measured +45 % time on the store stream (tools/tma_probe: 73.8 us per 222 MB tensor against 49.6-51.3 us with 416 /
        // 448 columns, profiles/r02_probe_tma_valid_cols.log).  The extra columns carry the epilogue's values of the zero-padded
        // features (z = 0, y = gabor(0) = 1 + 0j -- column 2M of a y tensor is the wgrad's "ones" column and is forced to 1.0).
        int store_cols = J.e.n_cols;
        // (the TF32 kernels' FP16 saved-z tiles follow the same rule: their parameter tables are zero padded as well)
        if (store_sector_align()) { const int r = round_up(J.e.n_cols, 16); if (r <= J.o_pitch[nslot]) store_cols = r; }
        ok &= sm100_host::make_tmap_2d_t(&P.o_map[nslot], J.o[nslot], J.e.n_rows, store_cols, J.o_pitch[nslot], 32, 32,
                                         op16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, J.o_half[nslot]);
      }
      else
        ok &= sm100_host::make_tmap_2d(&P.o_map[nslot], J.o[nslot], J.e.n_rows, J.e.n_cols, J.o_pitch[nslot], 32, 32);
      ++nslot;
    }
  for (int s = nslot; s < 3; ++s) P.o_map[s] = P.a_map[0];
  P.z_map[0] = P.a_map[0]; P.z_map[1] = P.a_map[0];
  if (op16 && P.n_in && !J.e.z_half) return fail("16-bit row-tile kernels read FP16 saved pre-activations");
  const CUtensorMapSwizzle zsw = op16 ? CU_TENSOR_MAP_SWIZZLE_64B : (J.e.z_half ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B);
  if (P.n_in >= 1) ok &= sm100_host::make_tmap_2d(&P.z_map[0], J.e.z_src, J.e.n_rows, J.e.n_cols, J.e.zw_pitch, 32, 32, zsw, J.e.z_half != 0);
  if (P.n_in >= 2) ok &= sm100_host::make_tmap_2d(&P.z_map[1], J.e.w_src, J.e.n_rows, J.e.n_cols, J.e.zw_pitch, 32, 32, zsw, J.e.z_half != 0);
  if (!ok) return fail("cuTensorMapEncodeTiled failed (pointer/pitch alignment?)");
  if (J.gen) CU_OK(launch_rows_gen(J.mode, P, smem, g_sm_count, st));
  else if (op16) CU_OK(launch_rows16(J.mode, P, smem, g_sm_count, st));
  else CU_OK(launch_rows(J.mode, P, smem, g_sm_count, st, false));
  return 0;
}

// elem: element type of the packed matrix (FP32/TF32 kernels: kElemF32; mixed16: FP16 forward, BF16 dgrad)
int run_pack(const float* W1, const float* W2, int M_out, int K_in, int mode, const Blocking& blk, int k0_pad,
             int k_pad_total, float* B, int precision, cudaStream_t st, int elem = kElemF32, int pair_perm = 1) {
  const int total = blk.n_blocks * blk.nb * k_pad_total;
  const int grid = (total + 255) / 256;
  ProfScope prof(K_PACK, st);
  // 16-bit matrices feed tc_rows16_kernel, whose Gabor epilogues expect pair-transposed accumulator columns (pair_perm; the
  // plain epilogue of the single-layer dgrad keeps the natural order)
  if (elem == kElemF16)
    pack_weights_kernel<1><<<grid, 256, 0, st>>>(W1, W2, M_out, K_in, mode, blk.n_blocks, blk.nb, blk.nbh, k0_pad, k_pad_total, B, 0, pair_perm);
  else if (elem == kElemBF16)
    pack_weights_kernel<2><<<grid, 256, 0, st>>>(W1, W2, M_out, K_in, mode, blk.n_blocks, blk.nb, blk.nbh, k0_pad, k_pad_total, B, 0, pair_perm);
  else
    pack_weights_kernel<0><<<grid, 256, 0, st>>>(W1, W2, M_out, K_in, mode, blk.n_blocks, blk.nb, blk.nbh, k0_pad, k_pad_total, B,
                                                precision == WIRE_PRECISION_TF32, 0);
  CU_OK(cudaGetLastError());
  return 0;
}

// weight gradient of one complex Linear (x has a ones column at 2*k_in)
struct WgradGen {  // first-layer description when x = y0 is generated in place (nullptr coords = load x)
  const float* coords = nullptr;
  int in_features = 0;
  const float *w0 = nullptr, *b0 = nullptr, *w0b = nullptr, *b0b = nullptr, *omega = nullptr, *scale = nullptr;
  int two_d = 0;
};
int run_wgrad(const float* x, int x_pitch, int k_in, const float* g1, const float* g2, int g_pitch, int m_out, int64_t n,
              float* gW1, float* gB1, float* gW2, float* gB2, int precision, cudaStream_t st, const WgradGen* gen = nullptr,
              int x_elem = kElemF32, int g_elem = kElemF32) {
  if (n <= 0) return 0;
  const int n_g = g2 ? 2 : 1;
  ProfScope prof(K_WGRAD, st, (precision == WIRE_PRECISION_FP32 && g2) ? 2 : 1);
  if (precision == WIRE_PRECISION_FP32) {
    const int x_cols = 2 * k_in + 1, g_cols = 2 * m_out;
    dim3 grid((x_cols + 63) / 64, (g_cols + 63) / 64, 1);
    int splits = (4 * g_sm_count) / int(grid.x * grid.y);
    if (splits < 1) splits = 1;
    int rps = int((n + splits - 1) / splits);
    rps = round_up(rps, 16);
    grid.z = unsigned((n + rps - 1) / rps);
    simt_wgrad_kernel<<<grid, 256, 0, st>>>(x, x_pitch, k_in, g1, g_pitch, g_cols, int(n), rps, gW1, gB1);
    if (g2) simt_wgrad_kernel<<<grid, 256, 0, st>>>(x, x_pitch, k_in, g2, g_pitch, g_cols, int(n), rps, gW2, gB2);
    CU_OK(cudaGetLastError());
    return 0;
  }
  WgradParams P;
  memset(&P, 0, sizeof(P));
  P.n_rows = int(n); P.k_in = k_in; P.g_cols = 2 * m_out; P.n_g = n_g;
  P.gW[0] = gW1; P.gB[0] = gB1; P.gW[1] = gW2; P.gB[1] = gB2;
  const bool op16 = g_elem != kElemF32;
  const bool use_gen = gen && gen->coords && !op16;
  if (op16) {
    if (x_elem == kElemF32) return fail("16-bit wgrad needs a 16-bit x operand");
    P.x_fmt = x_elem == kElemBF16 ? int(sm100::kFmtBF16) : int(sm100::kFmtF16);
    P.g_fmt = g_elem == kElemBF16 ? int(sm100::kFmtBF16) : int(sm100::kFmtF16);
    P.x_conv = x_elem != g_elem;
    if (P.x_conv && !(x_elem == kElemF16 && g_elem == kElemBF16)) return fail("unsupported wgrad operand formats");
  }
  // CTA pair (256 x-columns per tile) or single CTAs (128)?  Measured equal at M = 212 (0.108 vs 0.110 ms), so on the 16-bit
  // path take whichever pads the 2K+1 x-columns less: at the SISR width (2K+1 = 257) pairs would spend half of their tiles
  // on the single "ones" column (2 x 256 column slots against 3 x 128).
  // The "ones" column (bias gradient for free) costs a whole extra x tile when 2K is a multiple of the tile width (wire2d at the
  // SISR width: 257 columns = 3 tiles of 128 instead of 2): the converter warps then sum the bias gradient from the g tiles
  // in shared memory instead (WgradParams::bias_sum; WIRE_B200_BIAS_SUM=0 keeps the ones column)
  const char* bs_env = getenv("WIRE_B200_BIAS_SUM");
  const bool bias_sum_on = !(bs_env && bs_env[0] == '0');
  P.bias_sum = (op16 && P.x_conv && bias_sum_on && (2 * k_in) % 128 == 0) ? 1 : 0;
  {
    const char* e = getenv("WIRE_B200_WGRAD_DUAL");   // =0: one work item per g tensor (A/B runs)
    P.dual = (e && e[0] == '0') ? 0 : 1;               // a request; wgrad_configure decides (pair, bias_sum, two 256-column tensors)
  }
  int cl = cluster_size();
  if (op16 && cl == 2) {
    const int xc = 2 * k_in + (P.bias_sum ? 0 : 1);
    if ((xc + 127) / 128 * 128 < (xc + 255) / 256 * 256) cl = 1;
  }
  const size_t smem = wgrad_configure(P, g_sm_count, cl, use_gen, op16);
  if (!smem) return fail("wgrad configuration does not fit shared memory");
  if (const char* e = getenv("WIRE_B200_WGRAD_SPLITS")) P.splits = atoi(e) > 0 ? atoi(e) : P.splits;   // experiments
  if (use_gen) {
    P.coords = gen->coords; P.in_features = gen->in_features; P.w0 = gen->w0; P.b0 = gen->b0; P.w0b = gen->w0b; P.b0b = gen->b0b;
    P.gen_omega = gen->omega; P.gen_scale = gen->scale; P.gen_two_d = gen->two_d;
    x = g1; x_pitch = g_pitch; /* x_map is unused by the GEN kernel but must be a valid descriptor */
  }
  bool ok;
  if (op16) {
    ok = sm100_host::make_tmap_2d_t(&P.x_map, x, n, 2 * k_in + (P.bias_sum ? 0 : 1), x_pitch, kWgradKC16, 64, CU_TENSOR_MAP_SWIZZLE_128B, x_elem);
    ok &= sm100_host::make_tmap_2d_t(&P.g_map[0], g1, n, 2 * m_out, g_pitch, kWgradKC16, 64, CU_TENSOR_MAP_SWIZZLE_128B, g_elem);
    ok &= sm100_host::make_tmap_2d_t(&P.g_map[1], g2 ? g2 : g1, n, 2 * m_out, g_pitch, kWgradKC16, 64, CU_TENSOR_MAP_SWIZZLE_128B, g_elem);
  } else {
    ok = sm100_host::make_tmap_2d(&P.x_map, x, n, use_gen ? 2 * m_out : 2 * k_in + 1, x_pitch, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    ok &= sm100_host::make_tmap_2d(&P.g_map[0], g1, n, 2 * m_out, g_pitch, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    ok &= sm100_host::make_tmap_2d(&P.g_map[1], g2 ? g2 : g1, n, 2 * m_out, g_pitch, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  }
  if (!ok) return fail("cuTensorMapEncodeTiled failed for wgrad");
  // WIRE_B200_WGRAD_CLK=1 (eager launches only, never under graph capture): per-CTA clock64 stamps of the kernel's phases
  static unsigned long long* clk_dev = nullptr;
  const bool clk = getenv("WIRE_B200_WGRAD_CLK") != nullptr;
  if (clk) {
    if (!clk_dev) CU_OK(cudaMalloc(&clk_dev, 1024 * 8 * sizeof(unsigned long long)));
    CU_OK(cudaMemsetAsync(clk_dev, 0, 1024 * 8 * sizeof(unsigned long long), st));
    P.dbg = clk_dev;
  }
  CU_OK(launch_wgrad(P, smem, st, use_gen, op16));
  if (clk) {
    static unsigned long long h[1024 * 8];
    CU_OK(cudaStreamSynchronize(st));
    CU_OK(cudaMemcpy(h, clk_dev, sizeof(h), cudaMemcpyDeviceToHost));
    double s[4] = {0, 0, 0, 0}; int nb = 0; unsigned long long t_min = ~0ull, t_max = 0;
    for (int b = 0; b < 1024; ++b) {
      const unsigned long long* t = h + b * 8;
      if (!t[0] || !t[4]) continue;
      ++nb;
      for (int k = 0; k < 4; ++k) s[k] += double(t[k + 1] - t[k]);
      t_min = t[0] < t_min ? t[0] : t_min; t_max = t[4] > t_max ? t[4] : t_max;
    }
    const int cps_h = ((int(n) + 63) / 64 + P.splits - 1) / P.splits;
    if (nb) fprintf(stderr, "[wgrad clk] splits %d chunks/split %d k-loop cycles/chunk %.0f stage bytes/cta %zu stages %d\n", P.splits, cps_h, s[1] / nb / cps_h,
                    (smem - 1024) / P.stages, P.stages);
    if (nb) fprintf(stderr, "[wgrad clk] ctas %d  setup+wait %.0f  k-loop %.0f  epilogue %.0f  teardown %.0f  (mean cycles; per-SM clocks, first start to last end %llu)\n",
                    nb, s[0] / nb, s[1] / nb, s[2] / nb, s[3] / nb, t_max - t_min);
  }
  return 0;
}

int run_first_fwd(const wire_net_desc* d, const wire_layer_params& p, const float* coords, int64_t n, int in_f, float* y,
                  int y_pitch, float* z_out, float* w_out, int zr_pitch, cudaStream_t st, int y_elem = kElemF32) {
  if (n <= 0) return 0;
  if (y_elem == kElemF16) {  // mixed16 whole-network path: FP16 activations
    if (z_out || w_out || (y_pitch % 4)) return fail("FP16 first layer: unsupported call");
    if (in_f > 3 || (y_pitch % 8)) return fail("FP16 first layer: unsupported shape");
    // exactly one wave of persistent blocks, each owns a contiguous row range; a block is as many whole passes over the quads
    // of a row as fit 256 threads (M = 212: 53 quads x 4 row slots = 212 threads -> 224)
    const int nq = (d->width + 3) / 4;
    const int qpp = nq < kFirstFwd16Threads ? nq : kFirstFwd16Threads;
    const int threads = round_up((kFirstFwd16Threads / qpp) * qpp, 32);
    const size_t smem = 0;
    int per_sm = 1;
    if (d->two_d) CU_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, first_fwd16_kernel<true>, threads, smem));
    else CU_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, first_fwd16_kernel<false>, threads, smem));
    int nblk = (per_sm < 1 ? 1 : per_sm) * g_sm_count;
    if (int64_t(nblk) * 32 > n) nblk = int((n + 31) / 32);
    const int rpb = int((n + nblk - 1) / nblk);
    const int grid = int((n + rpb - 1) / rpb);
    ProfScope prof(K_FIRST_FWD, st);
    const float* nul = nullptr;
    if (d->two_d)
      CU_OK(launch_pdl(first_fwd16_kernel<true>, dim3(grid), dim3(threads), smem, st, coords, int(n), in_f, int(d->width), p.weight, p.bias, p.weight2,
                       p.bias2, p.omega0, p.scale0, reinterpret_cast<__half*>(y), y_pitch, rpb, store_sector_align() ? 1 : 0));
    else
      CU_OK(launch_pdl(first_fwd16_kernel<false>, dim3(grid), dim3(threads), smem, st, coords, int(n), in_f, int(d->width), p.weight, p.bias, nul, nul,
                       p.omega0, p.scale0, reinterpret_cast<__half*>(y), y_pitch, rpb, store_sector_align() ? 1 : 0));
    return 0;
  }
  const int grid = int((n + kRowsPerBlock - 1) / kRowsPerBlock);
  const int round_y = d->precision == WIRE_PRECISION_TF32;
  ProfScope prof(K_FIRST_FWD, st);
  const float* w2 = d->two_d ? p.weight2 : nullptr;
  const float* b2 = d->two_d ? p.bias2 : nullptr;
  if (!z_out && !w_out && (y_pitch % 4) == 0) {  // whole-network path: streaming kernel, 16-byte stores
    if (d->precision == WIRE_PRECISION_TF32)
      first_fwd2_kernel<true><<<grid, 128, 0, st>>>(coords, int(n), in_f, d->width, p.weight, p.bias, w2, b2, p.omega0, p.scale0, y, y_pitch,
                                                    round_y, kRowsPerBlock);
    else
      first_fwd2_kernel<false><<<grid, 128, 0, st>>>(coords, int(n), in_f, d->width, p.weight, p.bias, w2, b2, p.omega0, p.scale0, y, y_pitch,
                                                     round_y, kRowsPerBlock);
  } else if (d->precision == WIRE_PRECISION_TF32) {
    first_fwd_kernel<true><<<grid, 256, 0, st>>>(coords, int(n), in_f, d->width, p.weight, p.bias, w2, b2, p.omega0, p.scale0, y, y_pitch,
                                                 round_y, z_out, w_out, zr_pitch, kRowsPerBlock);
  } else {
    first_fwd_kernel<false><<<grid, 256, 0, st>>>(coords, int(n), in_f, d->width, p.weight, p.bias, w2, b2, p.omega0, p.scale0, y, y_pitch,
                                                  round_y, z_out, w_out, zr_pitch, kRowsPerBlock);
  }
  CU_OK(cudaGetLastError());
  return 0;
}

// the loss fused into the top of the backward pass (wire_net_backward_mse)
struct MseFuse {
  const float* pred; const float* target; int64_t count_norm; float* ring; int ring_n; const long long* step_ptr; float* scratch;
};
int run_mse_ring(const MseFuse& m, int64_t count, cudaStream_t st) {
  int64_t g64 = (count + 255) / 256;
  const int grid = int(g64 > 1184 ? 1184 : g64);
  ProfScope prof(K_MSE, st);
  CU_OK(launch_pdl(mse_grad_kernel, dim3(grid), dim3(256), 0, st, m.pred, m.target, count, m.scratch, (float*)nullptr, m.count_norm, m.ring,
                   m.ring_n, m.step_ptr));
  return 0;
}

int run_top_bwd(const wire_net_desc* d, const float* g_out, int64_t n, const float* Wf, const float* z, const float* w, int zw_pitch,
                int z_half,
                const float* h, int h_pitch, const float* omega, const float* scale, float* gz, float* gw, int g_pitch, float* g_Wf,
                float* g_bf, cudaStream_t st, int g_elem = kElemF32, const MseFuse* mse = nullptr, float* g_omega = nullptr,
                float* g_scale = nullptr) {
  if (n <= 0) return 0;
  if ((g_omega || g_scale) && g_elem != kElemBF16) return fail("omega_0 / scale_0 gradients in the fused path need the mixed16 kernels");
  if (mse && g_elem != kElemBF16) {  // only the 16-bit TMA kernel computes the loss gradient itself
    TRY(run_mse_ring(*mse, n * d->out_features, st));
    g_out = mse->scratch;
    mse = nullptr;
  }
  if (g_elem == kElemBF16) {  // mixed16 whole-network path: FP16 z in, BF16 g_z out
    if (!(z && z_half && d->out_features <= 4 && d->width <= 1024 && (zw_pitch % 4) == 0 && (g_pitch % 4) == 0))
      return fail("BF16 top backward: unsupported shape");
    ProfScope prof(K_TOP_BWD, st);
    // TMA-streamed kernel: a row of `g_pitch` columns moves as n_box equal boxes of <= 256 columns (multiples of 8)
    int n_box = 0;
    for (int nb = (g_pitch + 255) / 256; nb <= 16 && zw_pitch == g_pitch; ++nb)
      if (g_pitch % nb == 0 && (g_pitch / nb) % 8 == 0 && g_pitch / nb <= 256) { n_box = nb; break; }
    const size_t tile_bytes = size_t(g_pitch) * kTopRows * 2;
    // ring depths: the maxima for wire; wire2d (two tensors per stage) at narrow widths runs 3 + 2 so that a fourth CTA fits per SM
    // (WIRE_B200_TOP_DEPTH="in,out" overrides)
    int in_depth = 4, out_depth = 3;
    if (w && d->width <= 160) { in_depth = 3; out_depth = 3; }   // measured at M = 128, 1 M rows: 0.48 ms against 0.53 (4 + 3), 0.66 (3 + 2)
    if (const char* e = getenv("WIRE_B200_TOP_DEPTH")) {
      int a = 0, b = 0;
      if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 2 && a <= kTopIn && b >= 2 && b <= kTopOut) { in_depth = a; out_depth = b; }
    }
    const size_t smem16 = size_t(in_depth + out_depth) * (w ? 2 : 1) * tile_bytes + 256;
    if (n_box > 0 && smem16 <= 200 * 1024 && d->width <= 992 && !getenv("WIRE_B200_TOP_SIMT")) {
      TopBwd16Params T;
      memset(&T, 0, sizeof(T));
      T.g_out = g_out; T.Wf = Wf; T.omega = omega; T.scale = scale; T.g_Wf = g_Wf; T.g_bf = g_bf;
      T.n = int(n); T.M = d->width; T.out_f = d->out_features; T.pitch = g_pitch; T.two_d = w ? 1 : 0;
      T.bw = g_pitch / n_box; T.n_box = n_box;
      T.gs_omega = g_omega; T.gs_scale = g_scale;
      T.in_depth = in_depth; T.out_depth = out_depth;
      if (mse) {
        T.pred = mse->pred; T.target = mse->target; T.g_scale = 2.0f / float(mse->count_norm); T.loss_scale = 1.0f / float(mse->count_norm);
        T.ring = mse->ring; T.ring_n = mse->ring_n; T.step_ptr = mse->step_ptr;
      }
      bool ok = sm100_host::make_tmap_2d_t(&T.z_map[0], z, n, g_pitch, zw_pitch, kTopRows, T.bw, CU_TENSOR_MAP_SWIZZLE_NONE, kElemF16);
      ok &= sm100_host::make_tmap_2d_t(&T.z_map[1], w ? w : z, n, g_pitch, zw_pitch, kTopRows, T.bw, CU_TENSOR_MAP_SWIZZLE_NONE, kElemF16);
      // stored width: whole 32-byte sectors (see run_rows_job); the threads of the padded features store zeros there
      T.store_cols = 2 * d->width;
      if (store_sector_align() && round_up(2 * d->width, 16) <= g_pitch) T.store_cols = round_up(2 * d->width, 16);
      ok &= sm100_host::make_tmap_2d_t(&T.g_map[0], gz, n, T.store_cols, g_pitch, kTopRows, T.bw, CU_TENSOR_MAP_SWIZZLE_NONE, kElemBF16);
      ok &= sm100_host::make_tmap_2d_t(&T.g_map[1], gw ? gw : gz, n, T.store_cols, g_pitch, kTopRows, T.bw, CU_TENSOR_MAP_SWIZZLE_NONE, kElemBF16);
      if (!ok) return fail("cuTensorMapEncodeTiled failed for top_bwd16");
      const int threads = round_up(d->width, 32) + 32;  // compute warps + the I/O warp
      const int n_tiles = int((n + kTopRows - 1) / kTopRows);
      auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 256);
        if (e != cudaSuccess) return e;
        // persistent CTAs: exactly one wave of what is really resident (registers limit it to 4 CTAs of 256 threads per SM;
        // a grid sized from shared memory alone ran 1.75 waves with a 75 %-occupied tail)
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem16);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        int grid16 = g_sm_count * per_sm;
        if (grid16 > n_tiles) grid16 = n_tiles;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid16); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem16; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        cfg.attrs = attr; cfg.numAttrs = 0;
        add_pdl_attr(attr, cfg.numAttrs);
        return cudaLaunchKernelEx(&cfg, kern, T);
      };
      cudaError_t e = cudaErrorInvalidValue;
      const bool small = threads <= 512;
      if (g_omega || g_scale) {
        switch (d->out_features * 2 + (w ? 1 : 0)) {
          case 2: e = small ? launch(top_bwd16_kernel<false, 1, 512, true>) : launch(top_bwd16_kernel<false, 1, 1024, true>); break;
          case 3: e = small ? launch(top_bwd16_kernel<true, 1, 512, true>) : launch(top_bwd16_kernel<true, 1, 1024, true>); break;
          case 4: e = small ? launch(top_bwd16_kernel<false, 2, 512, true>) : launch(top_bwd16_kernel<false, 2, 1024, true>); break;
          case 5: e = small ? launch(top_bwd16_kernel<true, 2, 512, true>) : launch(top_bwd16_kernel<true, 2, 1024, true>); break;
          case 6: e = small ? launch(top_bwd16_kernel<false, 3, 512, true>) : launch(top_bwd16_kernel<false, 3, 1024, true>); break;
          case 7: e = small ? launch(top_bwd16_kernel<true, 3, 512, true>) : launch(top_bwd16_kernel<true, 3, 1024, true>); break;
          case 8: e = small ? launch(top_bwd16_kernel<false, 4, 512, true>) : launch(top_bwd16_kernel<false, 4, 1024, true>); break;
          case 9: e = small ? launch(top_bwd16_kernel<true, 4, 512, true>) : launch(top_bwd16_kernel<true, 4, 1024, true>); break;
        }
        CU_OK(e);
        return 0;
      }
      switch (d->out_features * 2 + (w ? 1 : 0)) {
        case 2: e = small ? launch(top_bwd16_kernel<false, 1, 512>) : launch(top_bwd16_kernel<false, 1, 1024>); break;
        case 3: e = small ? launch(top_bwd16_kernel<true, 1, 512>) : launch(top_bwd16_kernel<true, 1, 1024>); break;
        case 4: e = small ? launch(top_bwd16_kernel<false, 2, 512>) : launch(top_bwd16_kernel<false, 2, 1024>); break;
        case 5: e = small ? launch(top_bwd16_kernel<true, 2, 512>) : launch(top_bwd16_kernel<true, 2, 1024>); break;
        case 6: e = small ? launch(top_bwd16_kernel<false, 3, 512>) : launch(top_bwd16_kernel<false, 3, 1024>); break;
        case 7: e = small ? launch(top_bwd16_kernel<true, 3, 512>) : launch(top_bwd16_kernel<true, 3, 1024>); break;
        case 8: e = small ? launch(top_bwd16_kernel<false, 4, 512>) : launch(top_bwd16_kernel<false, 4, 1024>); break;
        case 9: e = small ? launch(top_bwd16_kernel<true, 4, 512>) : launch(top_bwd16_kernel<true, 4, 1024>); break;
      }
      CU_OK(e);
      return 0;
    }
    if (g_omega || g_scale) return fail("omega_0 / scale_0 gradients: this width is outside the TMA-streamed top backward kernel");
    if (mse) { TRY(run_mse_ring(*mse, n * d->out_features, st)); g_out = mse->scratch; mse = nullptr; }
    const int thr = round_up((d->width + 1) / 2, 32) < 128 ? 128 : round_up((d->width + 1) / 2, 32);
    const int nblk = int(n < int64_t(10 * g_sm_count) * 64 ? (n + 63) / 64 : 10 * g_sm_count);
    const int rpb = int(((n + nblk - 1) / nblk + 63) / 64 * 64);
    const int grid = int((n + rpb - 1) / rpb);
    if (w) top_bwd2_kernel<true, true, true, true><<<grid, thr, 0, st>>>(g_out, int(n), d->width, d->out_features, Wf, z, w, zw_pitch, omega, scale, gz, gw, g_pitch, 0, g_Wf, g_bf, rpb);
    else top_bwd2_kernel<true, false, true, true><<<grid, thr, 0, st>>>(g_out, int(n), d->width, d->out_features, Wf, z, w, zw_pitch, omega, scale, gz, gw, g_pitch, 0, g_Wf, g_bf, rpb);
    CU_OK(cudaGetLastError());
    return 0;
  }
  const int grid = int((n + kRowsPerBlock - 1) / kRowsPerBlock);
  const int round_g = (d->precision == WIRE_PRECISION_TF32) && z;
  ProfScope prof(K_TOP_BWD, st);
  const bool tf = d->precision == WIRE_PRECISION_TF32;
  if (z && d->out_features <= 4 && d->width <= 1024 && (zw_pitch % 4) == 0 && (g_pitch % 4) == 0) {  // training path: streaming kernel
    if (mse) { TRY(run_mse_ring(*mse, n * d->out_features, st)); g_out = mse->scratch; mse = nullptr; }
    const int thr = round_up((d->width + 1) / 2, 32) < 128 ? 128 : round_up((d->width + 1) / 2, 32);  // one feature pair per thread
    const int M = d->width, of = d->out_features;
    // few, long-lived blocks: every g_Wf address then sees only `nblk` atomics
    const int nblk = int(n < int64_t(10 * g_sm_count) * 64 ? (n + 63) / 64 : 10 * g_sm_count);
    const int kRowsPerBlock = int(((n + nblk - 1) / nblk + 63) / 64 * 64);
    const int grid = int((n + kRowsPerBlock - 1) / kRowsPerBlock);
    if (tf && w && z_half) top_bwd2_kernel<true, true, true><<<grid, thr, 0, st>>>(g_out, int(n), M, of, Wf, z, w, zw_pitch, omega, scale, gz, gw, g_pitch, round_g, g_Wf, g_bf, kRowsPerBlock);
    else if (tf && z_half) top_bwd2_kernel<true, false, true><<<grid, thr, 0, st>>>(g_out, int(n), M, of, Wf, z, w, zw_pitch, omega, scale, gz, gw, g_pitch, round_g, g_Wf, g_bf, kRowsPerBlock);
    else if (tf && w) top_bwd2_kernel<true, true><<<grid, thr, 0, st>>>(g_out, int(n), M, of, Wf, z, w, zw_pitch, omega, scale, gz, gw, g_pitch, round_g, g_Wf, g_bf, kRowsPerBlock);
    else if (tf) top_bwd2_kernel<true, false><<<grid, thr, 0, st>>>(g_out, int(n), M, of, Wf, z, w, zw_pitch, omega, scale, gz, gw, g_pitch, round_g, g_Wf, g_bf, kRowsPerBlock);
    else if (w) top_bwd2_kernel<false, true><<<grid, thr, 0, st>>>(g_out, int(n), M, of, Wf, z, w, zw_pitch, omega, scale, gz, gw, g_pitch, round_g, g_Wf, g_bf, kRowsPerBlock);
    else top_bwd2_kernel<false, false><<<grid, thr, 0, st>>>(g_out, int(n), M, of, Wf, z, w, zw_pitch, omega, scale, gz, gw, g_pitch, round_g, g_Wf, g_bf, kRowsPerBlock);
  } else if (tf) {
    top_bwd_kernel<true><<<grid, 256, 0, st>>>(g_out, int(n), d->width, d->out_features, Wf, z, w, zw_pitch, h, h_pitch, omega, scale, gz,
                                               gw, g_pitch, round_g, g_Wf, g_bf, kRowsPerBlock);
  } else {
    top_bwd_kernel<false><<<grid, 256, 0, st>>>(g_out, int(n), d->width, d->out_features, Wf, z, w, zw_pitch, h, h_pitch, omega, scale, gz,
                                                gw, g_pitch, round_g, g_Wf, g_bf, kRowsPerBlock);
  }
  CU_OK(cudaGetLastError());
  return 0;
}

int run_first_wgrad(const float* gz0, int g_pitch, const float* coords, int64_t n, int in_f, int M, float* gW, float* gb, cudaStream_t st,
                    int g_elem = kElemF32) {
  if (n <= 0 || !gW) return 0;
  const int grid = int((n + kRowsPerBlock - 1) / kRowsPerBlock);
  ProfScope prof(K_FIRST_WGRAD, st);
  if (g_elem == kElemBF16) {
    const char* fs_env = getenv("WIRE_B200_FWGRAD_STREAM");   // =0: the register-pipelined kernel (A/B runs)
    const bool stream_on = !(fs_env && fs_env[0] == '0');
    const bool aligned = ((reinterpret_cast<uintptr_t>(gz0) | reinterpret_cast<uintptr_t>(coords)) & 15) == 0;
    // (below ~64 k rows the ring's set-up and the block reduction cost more than the streaming saves: 21 vs 15 us at 25 k rows)
    if (stream_on && n >= 65536 && aligned && in_f >= 1 && in_f <= 3 && M <= 256 && (g_pitch % 8) == 0 && g_pitch <= 256) {
      // streamed variant: bulk copies into a shared-memory ring, one block per SM, reversed sweep (simt16_kernels.cuh)
      const int lpr = g_pitch <= 64 ? 8 : (g_pitch <= 128 ? 16 : 32);   // lanes per row (16-byte octets)
      const int chunk_rows = fw16s_rows(lpr);
      const uint32_t stage_bytes = fw16s_stage_bytes(g_pitch, chunk_rows);
      int stages = int((210u * 1024u) / stage_bytes);
      stages = stages > 8 ? 8 : stages;
      size_t smem = size_t(stages) * stage_bytes;
      smem = (smem < 65536 ? 65536 : smem) + 128;   // the block reduction parks 16 x 32 x 32 partial sums in the ring
      const int n_chunks = int((n + chunk_rows - 1) / chunk_rows);
      CUtensorMap g_map;
      const char* b1 = getenv("WIRE_B200_FWGRAD_BULK1D");   // =1: 1-D bulk copies instead of the tensor box (A/B runs)
      const int use_tma = !(b1 && b1[0] == '1');
      if (!sm100_host::make_tmap_2d_t(&g_map, gz0, n, g_pitch, g_pitch, chunk_rows, g_pitch, CU_TENSOR_MAP_SWIZZLE_NONE, kElemBF16))
        return fail("cuTensorMapEncodeTiled failed for first_wgrad16s");
      auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
        if (e != cudaSuccess) return e;
        return launch_pdl(kern, dim3(n_chunks < g_sm_count ? n_chunks : g_sm_count), dim3(kFw16sThreads), smem, st,
                          reinterpret_cast<const __nv_bfloat16*>(gz0), g_pitch, coords, int(n), M, gW, gb, stages, g_map, use_tma);
      };
      cudaError_t e = cudaErrorInvalidValue;
      switch (in_f * 100 + lpr) {
        case 108: e = launch(first_wgrad16s_kernel<1, 8>); break;
        case 116: e = launch(first_wgrad16s_kernel<1, 16>); break;
        case 132: e = launch(first_wgrad16s_kernel<1, 32>); break;
        case 208: e = launch(first_wgrad16s_kernel<2, 8>); break;
        case 216: e = launch(first_wgrad16s_kernel<2, 16>); break;
        case 232: e = launch(first_wgrad16s_kernel<2, 32>); break;
        case 308: e = launch(first_wgrad16s_kernel<3, 8>); break;
        case 316: e = launch(first_wgrad16s_kernel<3, 16>); break;
        case 332: e = launch(first_wgrad16s_kernel<3, 32>); break;
      }
      CU_OK(e);
    } else if (in_f <= 3 && M <= 256 && (g_pitch % 8) == 0) {
      const int nblk = int(n < int64_t(2 * g_sm_count) * 256 ? (n + 255) / 256 : 2 * g_sm_count);
      const int rpb = int((n + nblk - 1) / nblk);
      CU_OK(launch_pdl(first_wgrad16_kernel, dim3(int((n + rpb - 1) / rpb)), dim3(kFirstWgrad16Threads), 0, st,
                       reinterpret_cast<const __nv_bfloat16*>(gz0), g_pitch, coords, int(n), in_f, M, gW, gb, rpb));
    } else {
      first_wgrad_kernel<true><<<grid, 256, 0, st>>>(gz0, g_pitch, coords, int(n), in_f, M, gW, gb, kRowsPerBlock);
    }
    CU_OK(cudaGetLastError());
    return 0;
  }
  if (in_f <= 4 && M <= 256 && (g_pitch % 4) == 0) {
    const int nblk = int(n < int64_t(4 * g_sm_count) * 128 ? (n + 127) / 128 : 4 * g_sm_count);
    const int rpb = int(((n + nblk - 1) / nblk + 127) / 128 * 128);
    first_wgrad2_kernel<<<int((n + rpb - 1) / rpb), 256, 0, st>>>(gz0, g_pitch, coords, int(n), in_f, M, gW, gb, rpb);
  } else {
    first_wgrad_kernel<false><<<grid, 256, 0, st>>>(gz0, g_pitch, coords, int(n), in_f, M, gW, gb, kRowsPerBlock);
  }
  CU_OK(cudaGetLastError());
  return 0;
}

int zero(float* p, size_t floats, cudaStream_t st) {
  if (p) CU_OK(cudaMemsetAsync(p, 0, floats * sizeof(float), st));
  return 0;
}

RowsEpi base_epi(int64_t n, int n_cols, int precision) {
  RowsEpi e;
  memset(&e, 0, sizeof(e));
  e.n_rows = int(n);
  e.n_cols = n_cols;
  e.round_out0 = precision == WIRE_PRECISION_TF32;
  return e;
}

// tile configuration of hidden layer l's forward / backward row-tile GEMM (the same everywhere it is needed)
bool fwd_job_blocking(const wire_net_desc* d, const Layout& L, int l, int training, Blocking& blk, int& mask, bool& fuse, bool gen = false) {
  fuse = (l == L.H) && L.fuse_final;
  mask = fuse ? 0 : 1;
  if (training) { mask |= 2; if (d->two_d) mask |= 4; }
  return job_blocking(d->two_d ? MODE_GABOR2D_FWD : MODE_GABOR_FWD, L.two_m, mask, fuse, blk, gen, d->precision == WIRE_PRECISION_MIXED16);
}
bool bwd_job_blocking(const wire_net_desc* d, const Layout& L, int l, Blocking& blk, int& mask, int& bmode) {
  const bool to_first = (l == 1);
  mask = to_first ? 0 : (d->two_d ? 3 : 1);
  bmode = to_first ? (d->two_d ? MODE_FIRST2D_BWD : MODE_FIRST_BWD) : (d->two_d ? MODE_GABOR2D_BWD : MODE_GABOR_BWD);
  return job_blocking(bmode, L.two_m, mask, false, blk, false, L.g_elem != kElemF32);
}

// mixed16: every packed weight matrix of the step (forward FP16, and when training the dgrad BF16 ones the backward
// pass will read from the same workspace) in one launch
int pack_all16(const wire_net_desc* d, const wire_net_params* p, const Layout& L, void* ws, int training, cudaStream_t st) {
  PackJobs J;
  memset(&J, 0, sizeof(J));
  J.M_out = L.M; J.K_in = L.M;
  int max_total = 0;
  for (int l = 1; l <= L.H; ++l) {
    Blocking blk; int mask; bool fuse;
    if (!fwd_job_blocking(d, L, l, training, blk, mask, fuse)) return fail("no tile configuration");
    if (J.n + 2 > kMaxPackJobs) return fail("too many layers for the fused weight packing");
    PackJob& a = J.job[J.n++];
    a.W1 = p->layer[l].weight; a.W2 = d->two_d ? p->layer[l].weight2 : nullptr; a.B = at(ws, L.off_bf[l]);
    a.mode = 0; a.n_blocks = blk.n_blocks; a.nb = blk.nb; a.nbh = blk.nbh; a.k0_pad = L.k_pad; a.k_pad_total = L.k_pad; a.elem = L.y_elem;
    max_total = std::max(max_total, a.n_blocks * a.nb * a.k_pad_total);
    if (training) {
      int bmode;
      if (!bwd_job_blocking(d, L, l, blk, mask, bmode)) return fail("no tile configuration");
      PackJob& b = J.job[J.n++];
      b.W1 = p->layer[l].weight; b.W2 = d->two_d ? p->layer[l].weight2 : nullptr; b.B = at(ws, L.off_bd[l]);
      b.mode = 1; b.n_blocks = blk.n_blocks; b.nb = blk.nb; b.nbh = blk.nbh; b.k0_pad = L.k_pad; b.k_pad_total = (d->two_d ? 2 : 1) * L.k_pad;
      b.elem = L.g_elem;
      max_total = std::max(max_total, b.n_blocks * b.nb * b.k_pad_total);
    }
  }
  ProfScope prof(K_PACK, st);
  int gx = (max_total + 255) / 256;
  if (gx > 8 * g_sm_count) gx = 8 * g_sm_count;
  CU_OK(launch_pdl(pack_all16_kernel, dim3(gx, J.n), dim3(256), 0, st, J));
  return 0;
}

// ------------------------------------------------------------------------------------------
// whole-network forward for one chunk of rows
// ------------------------------------------------------------------------------------------
int forward_chunk(const wire_net_desc* d, const wire_net_params* p, const Layout& L, const float* coords, int64_t n, float* out,
                  void* ws, int training, cudaStream_t st) {
  const int M = L.M, H = L.H;
  float* y_prev = at(ws, L.off_y[0]);
  // WIRE_B200_GEN=1 (experiment): the first layer's output is generated inside its consumers and never written to HBM.
  // Correct (parity-green) but slower on B200: four generator warps cannot hide the MUFU latency (fwd 0.25 -> 0.36 ms,
  // wgrad 0.22 -> 0.55 ms, profiles/r01_bench_v9_gen.json), so the default keeps first_fwd2_kernel.
  const bool gen0 = d->precision == WIRE_PRECISION_TF32 && d->in_features <= 3 && getenv("WIRE_B200_GEN") != nullptr;
  const bool mixed = d->precision == WIRE_PRECISION_MIXED16;
  if (mixed) TRY(pack_all16(d, p, L, ws, training, st));
  if (!gen0) TRY(run_first_fwd(d, p->layer[0], coords, n, d->in_features, y_prev, L.P, nullptr, nullptr, 0, st, L.y_elem));
  for (int l = 1; l <= H; ++l) {
    Blocking blk;
    int mask;
    bool fuse;
    const bool gen = gen0 && l == 1;
    if (!fwd_job_blocking(d, L, l, training, blk, mask, fuse, gen)) return fail("no tile configuration");
    const bool store_y = !fuse;
    float* y_out = nullptr;
    if (store_y) y_out = training ? at(ws, L.off_y[l]) : at(ws, L.off_y[l & 1]);
    float* Bf = at(ws, L.off_bf[l]);
    if (!mixed) TRY(run_pack(p->layer[l].weight, d->two_d ? p->layer[l].weight2 : nullptr, M, M, 0, blk, L.k_pad, L.k_pad, Bf, d->precision, st, L.y_elem));
    RowsJob J;
    memset(&J, 0, sizeof(J));
    J.mode = d->two_d ? MODE_GABOR2D_FWD : MODE_GABOR_FWD;
    J.a_elem = L.y_elem; J.b_elem = L.y_elem;
    J.a[0] = y_prev; J.a_pitch[0] = L.P; J.k_cols[0] = L.two_m;
    J.b = Bf; J.b_rows = blk.n_blocks * blk.nb; J.b_pitch = L.k_pad; J.k0_pad = L.k_pad;
    J.blk = blk;
    int slot = 0;
    if (mask & 1) { J.o[slot] = y_out; J.o_half[slot] = L.y_elem; J.o_pitch[slot++] = L.P; }
    const int zh = L.z_elem != kElemF32;  // saved z / w in FP16 on the tensor-core paths
    if (mask & 2) { J.o[slot] = at(ws, L.off_z[l]); J.o_half[slot] = L.z_elem; J.o_pitch[slot++] = L.P; }
    if (mask & 4) { J.o[slot] = at(ws, L.off_w[l]); J.o_half[slot] = L.z_elem; J.o_pitch[slot++] = L.P; }
    J.store_mask = mask;
    J.e = base_epi(n, L.two_m, d->precision);
    J.e.z_half = zh;
    J.e.bias = p->layer[l].bias; J.e.bias2 = p->layer[l].bias2;
    J.e.omega = p->layer[l].omega0; J.e.scale = p->layer[l].scale0;
    if (gen) {
      J.gen = 1; J.gen_omega = p->layer[0].omega0; J.gen_scale = p->layer[0].scale0; J.gen_two_d = d->two_d;
      J.e.coords = coords; J.e.in_features = d->in_features;
      J.e.w0 = p->layer[0].weight; J.e.b0 = p->layer[0].bias; J.e.w0b = p->layer[0].weight2; J.e.b0b = p->layer[0].bias2;
      J.a[0] = Bf; J.a_pitch[0] = L.k_pad;  // unused by the kernel, but the descriptor must be valid
    }
    if (fuse) {
      J.e.fuse_final = 1; J.e.wf = p->final_weight; J.e.bf = p->final_bias; J.e.out = out; J.e.out_features = d->out_features;
    }
    TRY(run_rows(J, d->precision, st));
    if (store_y) y_prev = y_out;
  }
  if (!L.fuse_final) {
    const int64_t g64 = (n * 32 + 255) / 256;
    const int grid = int(g64 > 65535 * 16 ? 65535 * 16 : g64);
    ProfScope prof(K_FINAL_FWD, st);
    if (mixed) final_fwd_kernel<true><<<grid, 256, 0, st>>>(y_prev, L.P, int(n), M, d->out_features, p->final_weight, p->final_bias, out);
    else final_fwd_kernel<false><<<grid, 256, 0, st>>>(y_prev, L.P, int(n), M, d->out_features, p->final_weight, p->final_bias, out);
    CU_OK(cudaGetLastError());
  }
  return 0;
}

}  // namespace

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

int wire_b200_abi_version(void) { return WIRE_B200_ABI_VERSION; }
const char* wire_b200_last_error(void) { return g_err; }
int wire_b200_device_ok(void) { return require_device(); }
int wire_b200_sm_count(void) { return device_info() ? 0 : g_sm_count; }
int64_t wire_b200_infer_chunk_rows(void) { return kInferChunk; }

int wire_b200_prof_enable(int32_t timing) { std::lock_guard<std::mutex> lk(g_prof_mu); g_prof.timing = timing; return 0; }
int wire_b200_prof_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < g_prof.n_pending; ++i) { cudaEventDestroy(g_prof.pending[i].e0); cudaEventDestroy(g_prof.pending[i].e1); }
  g_prof.n_pending = 0;
  for (int k = 0; k < K_COUNT; ++k) { g_prof.launches[k] = 0; g_prof.ms[k] = 0.0; }
  return 0;
}
int wire_b200_prof_kinds(void) { return K_COUNT; }
const char* wire_b200_prof_name(int32_t kind) { return (kind >= 0 && kind < K_COUNT) ? kProfNames[kind] : ""; }
/* Resolves pending events (synchronises on them) and returns launches + accumulated device ms of one kind. */
int wire_b200_prof_get(int32_t kind, uint64_t* launches, double* ms) {
  if (kind < 0 || kind >= K_COUNT) return fail("bad profile kind %d", kind);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < g_prof.n_pending; ++i) {
    float t = 0.f;
    if (cudaEventSynchronize(g_prof.pending[i].e1) == cudaSuccess &&
        cudaEventElapsedTime(&t, g_prof.pending[i].e0, g_prof.pending[i].e1) == cudaSuccess)
      g_prof.ms[g_prof.pending[i].kind] += t;
    cudaEventDestroy(g_prof.pending[i].e0);
    cudaEventDestroy(g_prof.pending[i].e1);
  }
  g_prof.n_pending = 0;
  if (launches) *launches = g_prof.launches[kind];
  if (ms) *ms = g_prof.ms[kind];
  return 0;
}

size_t wire_net_workspace_bytes(const wire_net_desc* d, int64_t n, int32_t training) {
  Layout L;
  if (!d || make_layout(d, n, training, L)) return 0;
  return L.total;
}

int wire_net_workspace_init(const wire_net_desc* d, int64_t n, int32_t training, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Layout L;
  TRY(make_layout(d, n, training, L));
  if (!workspace || workspace_bytes < L.total) return fail("workspace too small: %zu < %zu", workspace_bytes, L.total);
  CU_OK(cudaMemsetAsync(workspace, 0, L.total, st));
  // "ones" column (col 2M) of every activation buffer: the bias-gradient row of the wgrad GEMM
  for (int l = 0; l < L.n_act; ++l) {
    if (L.y_elem == kElemF16)
      set_column16_kernel<<<int((L.rows + 255) / 256), 256, 0, st>>>(reinterpret_cast<uint16_t*>(at(workspace, L.off_y[l])), L.P, L.rows,
                                                                      L.two_m, uint16_t(0x3C00));  // FP16 1.0
    else
      set_column_kernel<<<int((L.rows + 255) / 256), 256, 0, st>>>(at(workspace, L.off_y[l]), L.P, L.rows, L.two_m, 1.0f);
  }
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_net_forward(const wire_net_desc* d_in, const wire_net_params* p, const float* coords, int64_t n, float* out, void* workspace,
                     size_t workspace_bytes, int32_t training, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (!d_in || !p || !coords || !out) return fail("null argument");
  if (n <= 0) return 0;
  TRY(check_desc(d_in));
  wire_net_desc de = *d_in;
  de.precision = net_precision(d_in);  // from here on `precision` is the one that actually runs
  const wire_net_desc* d = &de;
  Layout L;
  TRY(make_layout(d, n, training, L));
  if (!workspace || workspace_bytes < L.total) return fail("workspace too small: %zu < %zu", workspace_bytes, L.total);
  if (training) return forward_chunk(d, p, L, coords, n, out, workspace, 1, st);
  for (int64_t off = 0; off < n; off += L.rows) {
    const int64_t m = (n - off) < L.rows ? (n - off) : L.rows;
    TRY(forward_chunk(d, p, L, coords + off * d->in_features, m, out + off * d->out_features, workspace, 0, st));
  }
  return 0;
}


/* Copies one tensor of a training workspace out as dense fp32 (converting FP16 / BF16): what the forward pass saved and the
 * backward pass left behind -- activations y_l, pre-activations z_l / w_l, gradient buffers g_z / g_w / g_z0 / g_w0. */
int wire_net_workspace_read(const wire_net_desc* d_in, int64_t n, const void* workspace, size_t workspace_bytes, int32_t which,
                            int32_t index, float* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!d_in || !workspace || !out) return fail("null argument");
  TRY(check_desc(d_in));
  wire_net_desc de = *d_in;
  de.precision = net_precision(d_in);
  Layout L;
  TRY(make_layout(&de, n, 1, L));
  if (workspace_bytes < L.total) return fail("workspace too small: %zu < %zu", workspace_bytes, L.total);
  if (n <= 0) return 0;
  size_t off = 0;
  int elem = kElemF32, pitch = L.P, cols = L.two_m;
  switch (which) {
    case WIRE_WS_Y:
      if (index < 0 || index >= L.n_act) return fail("activation y_%d is not kept (layers 0..%d are)", index, L.n_act - 1);
      off = L.off_y[index]; elem = L.y_elem; break;
    case WIRE_WS_Z: case WIRE_WS_W:
      if (index < 1 || index > L.H) return fail("pre-activation index %d outside 1..%d", index, L.H);
      if (which == WIRE_WS_W && !de.two_d) return fail("w tensors exist for wire2d only");
      off = which == WIRE_WS_Z ? L.off_z[index] : L.off_w[index]; elem = L.z_elem; break;
    case WIRE_WS_GZ: case WIRE_WS_GW:
      if (index < 0 || index > 1) return fail("gradient slot %d outside 0..1", index);
      if (which == WIRE_WS_GW && !de.two_d) return fail("g_w tensors exist for wire2d only");
      off = which == WIRE_WS_GZ ? L.off_gz[index] : L.off_gw[index]; elem = L.g_elem; break;
    case WIRE_WS_GZ0: case WIRE_WS_GW0:
      if (which == WIRE_WS_GW0 && !de.two_d) return fail("g_w0 exists for wire2d only");
      off = which == WIRE_WS_GZ0 ? L.off_gz0 : L.off_gw0; elem = L.g_elem; pitch = L.PR; cols = L.M; break;
    default: return fail("unknown workspace tensor %d", which);
  }
  const void* src = static_cast<const char*>(workspace) + off;
  int64_t g64 = (n * cols + 255) / 256;
  const int grid = int(g64 > 1184 * 8 ? 1184 * 8 : g64);
  if (elem == kElemF16) read_rows_kernel<1><<<grid, 256, 0, st>>>(src, pitch, n, cols, out);
  else if (elem == kElemBF16) read_rows_kernel<2><<<grid, 256, 0, st>>>(src, pitch, n, cols, out);
  else read_rows_kernel<0><<<grid, 256, 0, st>>>(src, pitch, n, cols, out);
  CU_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
namespace {
int net_backward_impl(const wire_net_desc* d_in, const wire_net_params* p, const float* coords, int64_t n, const float* grad_out,
                      const MseFuse* mse, void* workspace, size_t workspace_bytes, const wire_net_grads* g, float* grad_coords, void* stream);
}
extern "C" {
int wire_net_backward(const wire_net_desc* d_in, const wire_net_params* p, const float* coords, int64_t n, const float* grad_out,
                      void* workspace, size_t workspace_bytes, const wire_net_grads* g, float* grad_coords, void* stream) {
  if (!grad_out) return fail("null argument");
  return net_backward_impl(d_in, p, coords, n, grad_out, nullptr, workspace, workspace_bytes, g, grad_coords, stream);
}

int wire_net_backward_mse(const wire_net_desc* d_in, const wire_net_params* p, const float* coords, int64_t n, const float* pred,
                          const float* target, int64_t count_global, float* loss_ring, int32_t ring_n, const int64_t* step_dev,
                          float* grad_out_scratch, void* workspace, size_t workspace_bytes, const wire_net_grads* g, float* grad_coords,
                          void* stream) {
  if (!pred || !target || !loss_ring || !step_dev || !grad_out_scratch) return fail("null argument");
  if (ring_n < 2) return fail("loss ring needs at least 2 slots");
  if (!d_in) return fail("null descriptor");
  if (n <= 0) return fail("empty batch: clear the next ring slot on the host side instead");
  if (count_global < n * int64_t(d_in->out_features)) return fail("count_global smaller than this rank's element count");
  MseFuse m{pred, target, count_global, loss_ring, int(ring_n), reinterpret_cast<const long long*>(step_dev), grad_out_scratch};
  return net_backward_impl(d_in, p, coords, n, grad_out_scratch, &m, workspace, workspace_bytes, g, grad_coords, stream);
}
}  // extern "C"

namespace {
int net_backward_impl(const wire_net_desc* d_in, const wire_net_params* p, const float* coords, int64_t n, const float* grad_out,
                      const MseFuse* mse, void* workspace, size_t workspace_bytes, const wire_net_grads* g, float* grad_coords, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (!d_in || !p || !coords || !grad_out || !g) return fail("null argument");
  TRY(check_desc(d_in));
  wire_net_desc de = *d_in;
  de.precision = net_precision(d_in);
  const wire_net_desc* d = &de;
  Layout L;
  TRY(make_layout(d, n, 1, L));
  if (!workspace || workspace_bytes < L.total) return fail("workspace too small: %zu < %zu", workspace_bytes, L.total);
  const int M = L.M, H = L.H, in_f = d->in_features;
  // gradients are overwritten: clear the accumulation targets (one memset per slot, one for a flat buffer, or none)
  if (g->clear_mode == WIRE_GRADS_CLEAR_FLAT) {
    if (!g->flat_base || g->flat_floats == 0) return fail("WIRE_GRADS_CLEAR_FLAT without a flat buffer");
    TRY(zero(g->flat_base, g->flat_floats, st));
  } else if (g->clear_mode != WIRE_GRADS_PREZEROED && g->clear_mode != WIRE_GRADS_CLEAR_SLOTS) {
    return fail("unknown clear_mode %d", g->clear_mode);
  }
  if (g->clear_mode == WIRE_GRADS_CLEAR_SLOTS) {
    for (int l = 0; l <= H; ++l) { TRY(zero(g->layer[l].omega0, 1, st)); TRY(zero(g->layer[l].scale0, 1, st)); }
    TRY(zero(g->final_weight, size_t(d->out_features) * M * 2, st));
    TRY(zero(g->final_bias, size_t(d->out_features) * 2, st));
    TRY(zero(g->layer[0].weight, size_t(M) * in_f, st));
    TRY(zero(g->layer[0].bias, size_t(M), st));
    if (d->two_d) { TRY(zero(g->layer[0].weight2, size_t(M) * in_f, st)); TRY(zero(g->layer[0].bias2, size_t(M), st)); }
    for (int l = 1; l <= H; ++l) {
      TRY(zero(g->layer[l].weight, size_t(M) * M * 2, st));
      TRY(zero(g->layer[l].bias, size_t(M) * 2, st));
      if (d->two_d) { TRY(zero(g->layer[l].weight2, size_t(M) * M * 2, st)); TRY(zero(g->layer[l].bias2, size_t(M) * 2, st)); }
    }
  }
  if (n <= 0) return 0;
  if (!g->final_weight || !g->final_bias) return fail("final layer gradient buffers are required");

  int cur = 0;
  // final Linear backward + Gabor backward of the last hidden layer (h recomputed from z_H)
  TRY(run_top_bwd(d, grad_out, n, p->final_weight, at(workspace, L.off_z[H]), d->two_d ? at(workspace, L.off_w[H]) : nullptr, L.P,
                  L.z_elem != kElemF32, nullptr,
                  0, p->layer[H].omega0, p->layer[H].scale0, at(workspace, L.off_gz[cur]), d->two_d ? at(workspace, L.off_gw[cur]) : nullptr,
                  L.P, g->final_weight, g->final_bias, st, L.g_elem, mse, g->layer[H].omega0, g->layer[H].scale0));
  for (int l = H; l >= 1; --l) {
    const float* gz = at(workspace, L.off_gz[cur]);
    const float* gw = d->two_d ? at(workspace, L.off_gw[cur]) : nullptr;
    const float* x = at(workspace, L.off_y[l - 1]);
    WgradGen wg;
    if (l == 1 && d->precision == WIRE_PRECISION_TF32 && in_f <= 3 && getenv("WIRE_B200_GEN") != nullptr) {
      wg.coords = coords; wg.in_features = in_f; wg.w0 = p->layer[0].weight; wg.b0 = p->layer[0].bias;
      wg.w0b = p->layer[0].weight2; wg.b0b = p->layer[0].bias2; wg.omega = p->layer[0].omega0; wg.scale = p->layer[0].scale0;
      wg.two_d = d->two_d;
    }
    if (g->layer[l].weight)
      TRY(run_wgrad(x, L.P, M, gz, gw, L.P, M, n, g->layer[l].weight, g->layer[l].bias, g->layer[l].weight2, g->layer[l].bias2, d->precision, st,
                    &wg, L.y_elem, L.g_elem));
    // dgrad of layer l fused with the nonlinearity backward of layer l-1
    const bool to_first = (l == 1);
    int mask, bmode;
    Blocking blk;
    if (!bwd_job_blocking(d, L, l, blk, mask, bmode)) return fail("no tile configuration");
    float* Bd = at(workspace, L.off_bd[l]);
    const int kparts = d->two_d ? 2 : 1;
    // mixed16: the dgrad matrices were packed by the forward pass of this step (pack_all16); autograd semantics forbid
    // changing the weights between a forward and its backward
    if (d->precision != WIRE_PRECISION_MIXED16)
      TRY(run_pack(p->layer[l].weight, d->two_d ? p->layer[l].weight2 : nullptr, M, M, 1, blk, L.k_pad, kparts * L.k_pad, Bd, d->precision, st,
                   L.g_elem));
    RowsJob J;
    memset(&J, 0, sizeof(J));
    J.a_elem = L.g_elem; J.b_elem = L.g_elem;
    J.a[0] = gz; J.a_pitch[0] = L.P; J.k_cols[0] = L.two_m;
    if (d->two_d) { J.a[1] = gw; J.a_pitch[1] = L.P; J.k_cols[1] = L.two_m; }
    J.b = Bd; J.b_rows = blk.n_blocks * blk.nb; J.b_pitch = kparts * L.k_pad; J.k0_pad = L.k_pad;
    J.blk = blk;
    J.store_mask = mask;
    J.e = base_epi(n, L.two_m, d->precision);
    J.e.omega = p->layer[l - 1].omega0; J.e.scale = p->layer[l - 1].scale0;
    J.e.g_omega = g->layer[l - 1].omega0; J.e.g_scale = g->layer[l - 1].scale0;   // trainable scalars of the layer below
    if ((J.e.g_omega || J.e.g_scale) && d->precision != WIRE_PRECISION_MIXED16)
      return fail("omega_0 / scale_0 gradients in the fused path need the mixed16 precision (use the per-layer entry points otherwise)");
    if (!to_first) {
      J.mode = d->two_d ? MODE_GABOR2D_BWD : MODE_GABOR_BWD;
      J.e.z_src = at(workspace, L.off_z[l - 1]); J.e.w_src = d->two_d ? at(workspace, L.off_w[l - 1]) : nullptr; J.e.zw_pitch = L.P;
      J.e.z_half = L.z_elem != kElemF32;
      J.o[0] = at(workspace, L.off_gz[1 - cur]); J.o_pitch[0] = L.P; J.o_half[0] = L.g_elem;
      if (d->two_d) { J.o[1] = at(workspace, L.off_gw[1 - cur]); J.o_pitch[1] = L.P; J.o_half[1] = L.g_elem; }
    } else {
      J.mode = d->two_d ? MODE_FIRST2D_BWD : MODE_FIRST_BWD;
      J.e.coords = coords; J.e.in_features = in_f;
      J.e.w0 = p->layer[0].weight; J.e.b0 = p->layer[0].bias; J.e.w0b = p->layer[0].weight2; J.e.b0b = p->layer[0].bias2;
      J.e.gz0 = at(workspace, L.off_gz0); J.e.gw0 = d->two_d ? at(workspace, L.off_gw0) : nullptr; J.e.gz0_pitch = L.PR;
    }
    TRY(run_rows(J, d->precision, st));
    cur = 1 - cur;
  }
  TRY(run_first_wgrad(at(workspace, L.off_gz0), L.PR, coords, n, in_f, M, g->layer[0].weight, g->layer[0].bias, st, L.g_elem));
  if (d->two_d) TRY(run_first_wgrad(at(workspace, L.off_gw0), L.PR, coords, n, in_f, M, g->layer[0].weight2, g->layer[0].bias2, st, L.g_elem));
  if (grad_coords) {
    const int grid = int((n * 32 + 255) / 256);
    ProfScope prof(K_GRAD_COORDS, st, d->two_d ? 2 : 1);
    if (L.g_elem == kElemBF16) {
      grad_coords16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(at(workspace, L.off_gz0)), L.PR, int(n), in_f, M,
                                                 p->layer[0].weight, grad_coords, 0);
      if (d->two_d)
        grad_coords16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(at(workspace, L.off_gw0)), L.PR, int(n), in_f, M,
                                                   p->layer[0].weight2, grad_coords, 1);
    } else {
      grad_coords_kernel<<<grid, 256, 0, st>>>(at(workspace, L.off_gz0), L.PR, int(n), in_f, M, p->layer[0].weight, grad_coords, 0);
      if (d->two_d) grad_coords_kernel<<<grid, 256, 0, st>>>(at(workspace, L.off_gw0), L.PR, int(n), in_f, M, p->layer[0].weight2, grad_coords, 1);
    }
    CU_OK(cudaGetLastError());
  }
  return 0;
}

}  // namespace (net_backward_impl)

// ---------------------------------------------------------------------------------------------
// single layers
// ---------------------------------------------------------------------------------------------
namespace {
struct LayerLayout {
  int P_in, P_out, k_pad_in;
  size_t off_x, off_y, off_z, off_w, off_g1, off_g2, off_gx, off_b, total;
};
void make_layer_layout(const wire_net_desc* d, int in_f, int64_t n, LayerLayout& L) {
  memset(&L, 0, sizeof(L));
  L.P_in = round_up(2 * in_f + 1, 32);
  L.P_out = round_up(2 * d->width + 1, 32);
  L.k_pad_in = round_up(2 * in_f, 32);
  const int k_pad_out = round_up(2 * d->width, 32);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t rows = size_t(n < 1 ? 1 : n);
  L.off_x = take(rows * L.P_in * 4);
  L.off_y = take(rows * L.P_out * 4);
  L.off_z = take(rows * L.P_out * 4);
  L.off_w = take(rows * L.P_out * 4);
  L.off_g1 = take(rows * L.P_out * 4);
  L.off_g2 = take(rows * L.P_out * 4);
  L.off_gx = take(rows * L.P_in * 4);
  const size_t kp = size_t(L.k_pad_in > k_pad_out ? L.k_pad_in : k_pad_out);
  L.off_b = take(size_t(2 * kp + 1024) * (2 * kp) * 4);
  L.total = off;
}
int copy2d(const float* src, int sp, float* dst, int dp, int64_t n, int cols, int do_round, cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t total = n * cols;
  int grid = int((total + 255) / 256 > 1184 * 8 ? 1184 * 8 : (total + 255) / 256);
  ProfScope prof(K_LAYER_MISC, st);
  copy2d_kernel<<<grid, 256, 0, st>>>(src, sp, dst, dp, n, cols, do_round);
  CU_OK(cudaGetLastError());
  return 0;
}
// element-wise Gabor backward on caller tensors (per-layer API)
// hidden-layer Gabor backward on caller tensors with BF16 outputs (single-layer API under mixed16: g_z feeds the 16-bit GEMMs)
__global__ void gabor_bwd_ew16_kernel(const float* __restrict__ gy, const float* __restrict__ z, const float* __restrict__ w, int64_t n,
                                      int M, const float* __restrict__ omega_p, const float* __restrict__ scale_p,
                                      __nv_bfloat16* __restrict__ g1, __nv_bfloat16* __restrict__ g2, int g_pitch) {
  const float omega = __ldg(omega_p), s = __ldg(scale_p), s2 = s * s;
  const int64_t total = n * M;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / M;
    const int k = int(i % M);
    const float gr = gy[i * 2], gi = gy[i * 2 + 1];
    const float zr = z[i * 2], zi = z[i * 2 + 1];
    const float wr = w ? w[i * 2] : 0.f, wi = w ? w[i * 2 + 1] : 0.f;
    float yr, yi, gzr, gzi;
    gabor_fwd<true>(zr, zi, omega, s2, s2 * (wr * wr + wi * wi), yr, yi);
    const float pr = gabor_bwd(yr, yi, zr, zi, gr, gi, omega, s2, gzr, gzi);
    *reinterpret_cast<__nv_bfloat162*>(g1 + r * g_pitch + 2 * k) = __floats2bfloat162_rn(gzr, gzi);
    if (w) *reinterpret_cast<__nv_bfloat162*>(g2 + r * g_pitch + 2 * k) = __floats2bfloat162_rn(-2.0f * s2 * pr * wr, -2.0f * s2 * pr * wi);
  }
}
template <bool FAST>
__global__ void gabor_bwd_ew_kernel(const float* __restrict__ gy, const float* __restrict__ z, const float* __restrict__ w, int64_t n,
                                    int M, int is_first, const float* __restrict__ omega_p, const float* __restrict__ scale_p,
                                    float* __restrict__ g1, float* __restrict__ g2, int g_pitch, int do_round) {
  const float omega = __ldg(omega_p), s = __ldg(scale_p), s2 = s * s;
  const int64_t total = n * M;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / M;
    const int k = int(i % M);
    const float gr = gy[i * 2], gi = gy[i * 2 + 1];
    if (is_first) {
      const float zv = z[i], wv = w ? w[i] : 0.f;
      float yr, yi, gz;
      gabor_fwd<FAST>(zv, 0.f, omega, s2, s2 * wv * wv, yr, yi);
      const float pr = gabor_first_bwd(yr, yi, zv, gr, gi, omega, s2, gz);
      g1[r * g_pitch + k] = gz;
      if (w) g2[r * g_pitch + k] = -2.0f * s2 * pr * wv;
    } else {
      const float zr = z[i * 2], zi = z[i * 2 + 1];
      const float wr = w ? w[i * 2] : 0.f, wi = w ? w[i * 2 + 1] : 0.f;
      float yr, yi, gzr, gzi;
      gabor_fwd<FAST>(zr, zi, omega, s2, s2 * (wr * wr + wi * wi), yr, yi);
      const float pr = gabor_bwd(yr, yi, zr, zi, gr, gi, omega, s2, gzr, gzi);
      if (do_round) { gzr = sm100::round_tf32(gzr); gzi = sm100::round_tf32(gzi); }
      g1[r * g_pitch + 2 * k] = gzr;
      g1[r * g_pitch + 2 * k + 1] = gzi;
      if (w) {
        float a = -2.0f * s2 * pr * wr, b = -2.0f * s2 * pr * wi;
        if (do_round) { a = sm100::round_tf32(a); b = sm100::round_tf32(b); }
        g2[r * g_pitch + 2 * k] = a;
        g2[r * g_pitch + 2 * k + 1] = b;
      }
    }
  }
}
}  // namespace

namespace {
int grid_rows(int64_t total) { const int64_t g = (total + 255) / 256; return int(g > 1184 * 8 ? 1184 * 8 : (g < 1 ? 1 : g)); }
int to16(const float* src, int64_t n, int cols, void* dst, int pitch, int elem, cudaStream_t st) {
  ProfScope prof(K_LAYER_MISC, st);
  if (elem == kElemF16) to16_rows_kernel<1><<<grid_rows(n * cols), 256, 0, st>>>(src, n, cols, dst, pitch);
  else to16_rows_kernel<2><<<grid_rows(n * cols), 256, 0, st>>>(src, n, cols, dst, pitch);
  CU_OK(cudaGetLastError());
  return 0;
}
int from16(const void* src, int pitch, int64_t n, int cols, float* dst, int elem, cudaStream_t st) {
  ProfScope prof(K_LAYER_MISC, st);
  if (elem == kElemF16) read_rows_kernel<1><<<grid_rows(n * cols), 256, 0, st>>>(src, pitch, n, cols, dst);
  else read_rows_kernel<2><<<grid_rows(n * cols), 256, 0, st>>>(src, pitch, n, cols, dst);
  CU_OK(cudaGetLastError());
  return 0;
}

// One hidden layer on the 16-bit kernels of the mixed16 path (tc_rows16 GABOR_FWD; FP16 operands, FP16 y / z / w tiles): the
// caller's fp32 tensors are converted on the way in and out.  Same workspace layout as the 32-bit route (16-bit rows fit).
int layer_forward16(const wire_net_desc* d, int K, const wire_layer_params* p, const float* x, int64_t n, float* y, float* z_save,
                    float* w_save, void* workspace, const LayerLayout& L, cudaStream_t st) {
  const int M = d->width;
  const bool training = z_save != nullptr;
  int mask = 1;
  if (training) { mask |= 2; if (d->two_d) mask |= 4; }
  Blocking blk;
  if (!job_blocking(d->two_d ? MODE_GABOR2D_FWD : MODE_GABOR_FWD, 2 * M, mask, false, blk, false, true)) return fail("no tile configuration for width %d", M);
  const int k_pad = round_up(2 * K, 64);
  void* x16 = at(workspace, L.off_x);
  TRY(to16(x, n, 2 * K, x16, L.P_in, kElemF16, st));
  float* B = at(workspace, L.off_b);
  TRY(run_pack(p->weight, d->two_d ? p->weight2 : nullptr, M, K, 0, blk, k_pad, k_pad, B, WIRE_PRECISION_MIXED16, st, kElemF16));
  RowsJob J;
  memset(&J, 0, sizeof(J));
  J.mode = d->two_d ? MODE_GABOR2D_FWD : MODE_GABOR_FWD;
  J.a_elem = kElemF16; J.b_elem = kElemF16;
  J.a[0] = static_cast<const float*>(x16); J.a_pitch[0] = L.P_in; J.k_cols[0] = 2 * K;
  J.b = B; J.b_rows = blk.n_blocks * blk.nb; J.b_pitch = k_pad; J.k0_pad = k_pad;
  J.blk = blk;
  int slot = 0;
  J.o[slot] = at(workspace, L.off_y); J.o_half[slot] = kElemF16; J.o_pitch[slot++] = L.P_out;
  if (mask & 2) { J.o[slot] = at(workspace, L.off_z); J.o_half[slot] = kElemF16; J.o_pitch[slot++] = L.P_out; }
  if (mask & 4) { J.o[slot] = at(workspace, L.off_w); J.o_half[slot] = kElemF16; J.o_pitch[slot++] = L.P_out; }
  J.store_mask = mask;
  J.e = base_epi(n, 2 * M, WIRE_PRECISION_MIXED16);
  J.e.round_out0 = 0;
  J.e.z_half = 1;
  J.e.bias = p->bias; J.e.bias2 = p->bias2; J.e.omega = p->omega0; J.e.scale = p->scale0;
  TRY(run_rows(J, WIRE_PRECISION_MIXED16, st));
  TRY(from16(at(workspace, L.off_y), L.P_out, n, 2 * M, y, kElemF16, st));
  if (z_save) TRY(from16(at(workspace, L.off_z), L.P_out, n, 2 * M, z_save, kElemF16, st));
  if (w_save && d->two_d) TRY(from16(at(workspace, L.off_w), L.P_out, n, 2 * M, w_save, kElemF16, st));
  return 0;
}

// ... and its backward: Gabor backward of the caller's saved z / w (fp32 math, BF16 g_z / g_w), OP16 tc_wgrad (FP16 x converted
// to BF16 in shared memory, BF16 g), tc_rows16 PLAIN dgrad with BF16 operands
int layer_backward16(const wire_net_desc* d, int K, const wire_layer_params* p, const float* x, const float* z_save, const float* w_save,
                     const float* grad_y, int64_t n, float* grad_x, const wire_layer_grads* g, void* workspace, const LayerLayout& L,
                     cudaStream_t st) {
  const int M = d->width;
  __nv_bfloat16* g1 = reinterpret_cast<__nv_bfloat16*>(at(workspace, L.off_g1));
  __nv_bfloat16* g2 = d->two_d ? reinterpret_cast<__nv_bfloat16*>(at(workspace, L.off_g2)) : nullptr;
  const int g_pitch = L.P_out;
  {
    ProfScope prof(K_LAYER_MISC, st);
    gabor_bwd_ew16_kernel<<<grid_rows(n * M), 256, 0, st>>>(grad_y, z_save, d->two_d ? w_save : nullptr, n, M, p->omega0, p->scale0, g1, g2, g_pitch);
    CU_OK(cudaGetLastError());
  }
  void* x16 = at(workspace, L.off_x);
  TRY(to16(x, n, 2 * K, x16, L.P_in, kElemF16, st));
  set_column16_kernel<<<int((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<uint16_t*>(x16), L.P_in, n, 2 * K, uint16_t(0x3C00));  // FP16 1.0
  CU_OK(cudaGetLastError());
  if (g->weight)
    TRY(run_wgrad(static_cast<const float*>(x16), L.P_in, K, reinterpret_cast<const float*>(g1), reinterpret_cast<const float*>(g2), g_pitch, M, n,
                  g->weight, g->bias, g->weight2, g->bias2, WIRE_PRECISION_MIXED16, st, nullptr, kElemF16, kElemBF16));
  if (grad_x) {
    Blocking blk;
    if (!job_blocking(MODE_PLAIN, 2 * K, 1, false, blk, false, true)) return fail("no tile configuration");
    float* B = at(workspace, L.off_b);
    const int k_pad_out = round_up(2 * M, 64);
    const int kparts = d->two_d ? 2 : 1;
    TRY(run_pack(p->weight, d->two_d ? p->weight2 : nullptr, M, K, 1, blk, k_pad_out, kparts * k_pad_out, B, WIRE_PRECISION_MIXED16, st,
                 kElemBF16, /*pair_perm=*/0));
    RowsJob J;
    memset(&J, 0, sizeof(J));
    J.mode = MODE_PLAIN;
    J.a_elem = kElemBF16; J.b_elem = kElemBF16;
    J.a[0] = reinterpret_cast<const float*>(g1); J.a_pitch[0] = g_pitch; J.k_cols[0] = 2 * M;
    if (d->two_d) { J.a[1] = reinterpret_cast<const float*>(g2); J.a_pitch[1] = g_pitch; J.k_cols[1] = 2 * M; }
    J.b = B; J.b_rows = blk.n_blocks * blk.nb; J.b_pitch = kparts * k_pad_out; J.k0_pad = k_pad_out;
    J.blk = blk;
    J.o[0] = at(workspace, L.off_gx); J.o_pitch[0] = L.P_in; J.o_half[0] = kElemBF16;
    J.store_mask = 1;
    J.e = base_epi(n, 2 * K, WIRE_PRECISION_MIXED16);
    J.e.round_out0 = 0;
    TRY(run_rows(J, WIRE_PRECISION_MIXED16, st));
    TRY(from16(at(workspace, L.off_gx), L.P_in, n, 2 * K, grad_x, kElemBF16, st));
  }
  return 0;
}
}  // namespace

extern "C" {

size_t wire_gabor_layer_workspace_bytes(const wire_net_desc* d, int32_t is_first, int32_t in_features, int64_t n) {
  if (!d) return 0;
  LayerLayout L;
  make_layer_layout(d, in_features, n, L);
  return L.total;
}

int wire_gabor_layer_forward(const wire_net_desc* d_in, int32_t is_first, int32_t in_features, const wire_layer_params* p, const float* x,
                             int64_t n, float* y, float* z_save, float* w_save, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (!d_in || !p || !x || !y) return fail("null argument");
  const wire_net_desc de = layer_desc(d_in);
  const wire_net_desc* d = &de;
  if (n <= 0) return 0;
  const int M = d->width;
  if (is_first) {
    if (in_features > kMaxIn) return fail("first layer supports up to %d input features", kMaxIn);
    wire_net_desc dd = *d;
    return run_first_fwd(&dd, *p, x, n, in_features, y, 2 * M, z_save, d->two_d ? w_save : nullptr, M, st);
  }
  LayerLayout L;
  make_layer_layout(d, in_features, n, L);
  if (!workspace || workspace_bytes < L.total) return fail("layer workspace too small: %zu < %zu", workspace_bytes, L.total);
  if (layer16_ok(d_in, is_first)) return layer_forward16(d_in, in_features, p, x, n, y, z_save, w_save, workspace, L, st);
  const int tf32 = d->precision == WIRE_PRECISION_TF32;
  float* xp = at(workspace, L.off_x);
  TRY(copy2d(x, 2 * in_features, xp, L.P_in, n, 2 * in_features, tf32, st));
  const bool training = z_save != nullptr;
  int mask = 1;
  if (training) { mask |= 2; if (d->two_d) mask |= 4; }
  Blocking blk;
  if (!job_blocking(d->two_d ? MODE_GABOR2D_FWD : MODE_GABOR_FWD, 2 * M, mask, false, blk)) return fail("no tile configuration for width %d", M);
  float* B = at(workspace, L.off_b);
  TRY(run_pack(p->weight, d->two_d ? p->weight2 : nullptr, M, in_features, 0, blk, L.k_pad_in, L.k_pad_in, B, d->precision, st));
  RowsJob J;
  memset(&J, 0, sizeof(J));
  J.mode = d->two_d ? MODE_GABOR2D_FWD : MODE_GABOR_FWD;
  J.a[0] = xp; J.a_pitch[0] = L.P_in; J.k_cols[0] = 2 * in_features;
  J.b = B; J.b_rows = blk.n_blocks * blk.nb; J.b_pitch = L.k_pad_in; J.k0_pad = L.k_pad_in;
  J.blk = blk;
  int slot = 0;
  J.o[slot] = at(workspace, L.off_y); J.o_pitch[slot++] = L.P_out;
  if (mask & 2) { J.o[slot] = at(workspace, L.off_z); J.o_pitch[slot++] = L.P_out; }
  if (mask & 4) { J.o[slot] = at(workspace, L.off_w); J.o_pitch[slot++] = L.P_out; }
  J.store_mask = mask;
  J.e = base_epi(n, 2 * M, d->precision);
  J.e.round_out0 = 0;  // the caller sees y: keep full precision here
  J.e.bias = p->bias; J.e.bias2 = p->bias2; J.e.omega = p->omega0; J.e.scale = p->scale0;
  TRY(run_rows(J, d->precision, st));
  TRY(copy2d(at(workspace, L.off_y), L.P_out, y, 2 * M, n, 2 * M, 0, st));
  if (z_save) TRY(copy2d(at(workspace, L.off_z), L.P_out, z_save, 2 * M, n, 2 * M, 0, st));
  if (w_save && d->two_d) TRY(copy2d(at(workspace, L.off_w), L.P_out, w_save, 2 * M, n, 2 * M, 0, st));
  return 0;
}

int wire_gabor_layer_backward(const wire_net_desc* d_in, int32_t is_first, int32_t in_features, const wire_layer_params* p, const float* x,
                              const float* z_save, const float* w_save, const float* grad_y, int64_t n, float* grad_x,
                              const wire_layer_grads* g, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (!d_in || !p || !x || !z_save || !grad_y || !g) return fail("null argument");
  const wire_net_desc de = layer_desc(d_in);
  const wire_net_desc* d = &de;
  const int M = d->width, K = in_features;
  const size_t wsz = is_first ? size_t(M) * K : size_t(M) * K * 2;
  const size_t bsz = is_first ? size_t(M) : size_t(M) * 2;
  TRY(zero(g->weight, wsz, st)); TRY(zero(g->bias, bsz, st));
  if (d->two_d) { TRY(zero(g->weight2, wsz, st)); TRY(zero(g->bias2, bsz, st)); }
  if (n <= 0) return 0;
  LayerLayout L;
  make_layer_layout(d, in_features, n, L);
  if (!workspace || workspace_bytes < L.total) return fail("layer workspace too small: %zu < %zu", workspace_bytes, L.total);
  if (layer16_ok(d_in, is_first)) {
    if (d->two_d && !w_save) return fail("wire2d layer without its saved scale_orth pre-activation");
    return layer_backward16(d_in, in_features, p, x, z_save, w_save, grad_y, n, grad_x, g, workspace, L, st);
  }
  const int tf32 = d->precision == WIRE_PRECISION_TF32;
  float* g1 = at(workspace, L.off_g1);
  float* g2 = d->two_d ? at(workspace, L.off_g2) : nullptr;
  const float* wsv = d->two_d ? w_save : nullptr;
  const int g_pitch = L.P_out;
  {
    int64_t total = n * M;
    int grid = int((total + 255) / 256 > 1184 * 8 ? 1184 * 8 : (total + 255) / 256);
    if (tf32) gabor_bwd_ew_kernel<true><<<grid, 256, 0, st>>>(grad_y, z_save, wsv, n, M, is_first, p->omega0, p->scale0, g1, g2, g_pitch, !is_first);
    else gabor_bwd_ew_kernel<false><<<grid, 256, 0, st>>>(grad_y, z_save, wsv, n, M, is_first, p->omega0, p->scale0, g1, g2, g_pitch, 0);
    CU_OK(cudaGetLastError());
  }
  if (is_first) {
    TRY(run_first_wgrad(g1, g_pitch, x, n, K, M, g->weight, g->bias, st));
    if (d->two_d) TRY(run_first_wgrad(g2, g_pitch, x, n, K, M, g->weight2, g->bias2, st));
    if (grad_x) {
      const int grid = int((n * 32 + 255) / 256);
      grad_coords_kernel<<<grid, 256, 0, st>>>(g1, g_pitch, int(n), K, M, p->weight, grad_x, 0);
      if (d->two_d) grad_coords_kernel<<<grid, 256, 0, st>>>(g2, g_pitch, int(n), K, M, p->weight2, grad_x, 1);
      CU_OK(cudaGetLastError());
    }
    return 0;
  }
  // wgrad: x needs the ones column
  float* xp = at(workspace, L.off_x);
  TRY(copy2d(x, 2 * K, xp, L.P_in, n, 2 * K, tf32, st));
  set_column_kernel<<<int((n + 255) / 256), 256, 0, st>>>(xp, L.P_in, n, 2 * K, 1.0f);
  CU_OK(cudaGetLastError());
  if (g->weight) TRY(run_wgrad(xp, L.P_in, K, g1, g2, g_pitch, M, n, g->weight, g->bias, g->weight2, g->bias2, d->precision, st));
  if (grad_x) {
    Blocking blk;
    if (!job_blocking(MODE_PLAIN, 2 * K, 1, false, blk)) return fail("no tile configuration");
    float* B = at(workspace, L.off_b);
    const int k_pad_out = round_up(2 * M, 32);
    const int kparts = d->two_d ? 2 : 1;
    TRY(run_pack(p->weight, d->two_d ? p->weight2 : nullptr, M, K, 1, blk, k_pad_out, kparts * k_pad_out, B, d->precision, st));
    RowsJob J;
    memset(&J, 0, sizeof(J));
    J.mode = MODE_PLAIN;
    J.a[0] = g1; J.a_pitch[0] = g_pitch; J.k_cols[0] = 2 * M;
    if (d->two_d) { J.a[1] = g2; J.a_pitch[1] = g_pitch; J.k_cols[1] = 2 * M; }
    J.b = B; J.b_rows = blk.n_blocks * blk.nb; J.b_pitch = kparts * k_pad_out; J.k0_pad = k_pad_out;
    J.blk = blk;
    J.o[0] = at(workspace, L.off_gx); J.o_pitch[0] = L.P_in;
    J.store_mask = 1;
    J.e = base_epi(n, 2 * K, d->precision);
    J.e.round_out0 = 0;
    TRY(run_rows(J, d->precision, st));
    TRY(copy2d(at(workspace, L.off_gx), L.P_in, grad_x, 2 * K, n, 2 * K, 0, st));
  }
  return 0;
}

int wire_final_linear_forward(const wire_net_desc* d, const float* weight, const float* bias, const float* h, int64_t n, float* out,
                              void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (!d || !weight || !bias || !h || !out) return fail("null argument");
  if (d->out_features > kSimtMaxOut) return fail("out_features %d > %d", d->out_features, kSimtMaxOut);
  if (n <= 0) return 0;
  int64_t g64 = (n * 32 + 255) / 256;
  const int grid = int(g64 > 65535 * 16 ? 65535 * 16 : g64);
  final_fwd_kernel<false><<<grid, 256, 0, st>>>(h, 2 * d->width, int(n), d->width, d->out_features, weight, bias, out);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_final_linear_backward(const wire_net_desc* d_in, const float* weight, const float* h, const float* grad_out, int64_t n, float* grad_h,
                               float* grad_weight, float* grad_bias, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (!d_in || !weight || !h || !grad_out || !grad_weight || !grad_bias) return fail("null argument");
  const wire_net_desc de = layer_desc(d_in);
  const wire_net_desc* d = &de;
  TRY(zero(grad_weight, size_t(d->out_features) * d->width * 2, st));
  TRY(zero(grad_bias, size_t(d->out_features) * 2, st));
  return run_top_bwd(d, grad_out, n, weight, nullptr, nullptr, 0, 0, h, 2 * d->width, nullptr, nullptr, grad_h, nullptr, 2 * d->width,
                     grad_weight, grad_bias, st);
}

int wire_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t count, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (count <= 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return fail("null argument");
  const double bc1 = 1.0 - pow(double(beta1), double(step));
  const double bc2 = 1.0 - pow(double(beta2), double(step));
  int64_t g64 = (count + 255) / 256;
  const int grid = int(g64 > 1184 ? 1184 : g64);
  ProfScope prof(K_ADAM, st);
  adam_kernel<<<grid, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, count, lr, beta1, beta2, eps, weight_decay, float(bc1),
                                    float(sqrt(bc2)), grad_scale);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_adam_step_dev(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t count, const float* lr_dev,
                       float beta1, float beta2, float eps, float weight_decay, int64_t* step_dev, float grad_scale,
                       uint32_t* scratch_dev, int32_t zero_grad, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (count <= 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev || !scratch_dev) return fail("null argument");
  int64_t g64 = (count / 4 + 255) / 256;
  const int grid = int(g64 > 592 ? 592 : (g64 < 1 ? 1 : g64));
  ProfScope prof(K_ADAM, st);
  CU_OK(launch_pdl(adam_dev_kernel, dim3(grid), dim3(256), 0, st, param, grad, exp_avg, exp_avg_sq, count, lr_dev, beta1, beta2, eps, weight_decay,
                   reinterpret_cast<long long*>(step_dev), grad_scale, scratch_dev, int(zero_grad)));
  return 0;
}

int wire_mse_loss_grad(const float* pred, const float* target, int64_t count, float* grad_out, float* loss, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (count <= 0) return 0;
  if (!pred || !target || !grad_out) return fail("null argument");
  int64_t g64 = (count + 255) / 256;
  const int grid = int(g64 > 1184 ? 1184 : g64);
  ProfScope prof(K_MSE, st);
  CU_OK(launch_pdl(mse_grad_kernel, dim3(grid), dim3(256), 0, st, pred, target, count, grad_out, loss, count, (float*)nullptr, 0,
                   (const long long*)nullptr));
  return 0;
}

int wire_mse_loss_grad_n(const float* pred, const float* target, int64_t count, int64_t count_global, float* grad_out, float* loss,
                         void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (count <= 0) return 0;
  if (!pred || !target || !grad_out) return fail("null argument");
  if (count_global < count) return fail("count_global %lld < count %lld", (long long)count_global, (long long)count);
  int64_t g64 = (count + 255) / 256;
  const int grid = int(g64 > 1184 ? 1184 : g64);
  ProfScope prof(K_MSE, st);
  CU_OK(launch_pdl(mse_grad_kernel, dim3(grid), dim3(256), 0, st, pred, target, count, grad_out, loss, count_global, (float*)nullptr, 0,
                   (const long long*)nullptr));
  return 0;
}

int wire_mse_loss_grad_ring(const float* pred, const float* target, int64_t count, int64_t count_global, float* grad_out, float* loss_ring,
                            int32_t ring_n, const int64_t* step_dev, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!pred || !target || !grad_out || !loss_ring || !step_dev) return fail("null argument");
  if (ring_n < 2) return fail("loss ring needs at least 2 slots");
  if (count <= 0) return fail("empty batch: clear the next ring slot on the host side instead");
  if (count_global < count) return fail("count_global %lld < count %lld", (long long)count_global, (long long)count);
  int64_t g64 = (count + 255) / 256;
  const int grid = int(g64 > 1184 ? 1184 : g64);
  ProfScope prof(K_MSE, st);
  CU_OK(launch_pdl(mse_grad_kernel, dim3(grid), dim3(256), 0, st, pred, target, count, grad_out, (float*)nullptr, count_global, loss_ring,
                   int(ring_n), reinterpret_cast<const long long*>(step_dev)));
  return 0;
}

// ---- data-parallel exchange over NVLink peer memory (peer_kernels.cuh) ----------------------------------------------
size_t wire_peer_header_bytes(void) { return size_t(kPeerHeaderBytes); }

int wire_peer_alloc(size_t grad_floats, void** base, void* ipc_handle) {
  if (!base || !ipc_handle) return fail("null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == WIRE_B200_IPC_HANDLE_BYTES, "ipc handle size");
  void* ptr = nullptr;
  const size_t bytes = kPeerHeaderBytes + ((grad_floats + 3) / 4) * 16;
  CU_OK(cudaMalloc(&ptr, bytes));  // plain cudaMalloc: memory from the stream-ordered / VMM pools cannot be exported through cudaIpc
  CU_OK(cudaMemset(ptr, 0, bytes));
  CU_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { cudaFree(ptr); return fail("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
  memcpy(ipc_handle, &h, sizeof(h));
  *base = ptr;
  return 0;
}

int wire_peer_open(const void* ipc_handle, void** base) {
  if (!base || !ipc_handle) return fail("null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* ptr = nullptr;
  CU_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
  *base = ptr;
  return 0;
}

int wire_peer_close(void* base) {
  if (base) CU_OK(cudaIpcCloseMemHandle(base));
  return 0;
}

int wire_peer_free(void* base) {
  if (base) CU_OK(cudaFree(base));
  return 0;
}

static int make_peer_table(PeerTable& T, void* const* peer_bases, int32_t world, int32_t rank) {
  if (!peer_bases) return fail("null argument");
  if (world < 1 || world > kMaxPeers) return fail("world %d outside 1..%d", world, kMaxPeers);
  if (rank < 0 || rank >= world) return fail("rank %d outside 0..%d", rank, world - 1);
  memset(&T, 0, sizeof(T));
  for (int r = 0; r < world; ++r) {
    if (!peer_bases[r]) return fail("peer buffer %d is null", r);
    T.base[r] = peer_bases[r];
  }
  T.world = world;
  T.rank = rank;
  return 0;
}

int wire_adam_step_peer(float* param, void* const* peer_bases, int32_t world, int32_t rank, float* exp_avg, float* exp_avg_sq,
                        int64_t count, const float* lr_dev, float beta1, float beta2, float eps, float weight_decay, int64_t* step_dev,
                        float grad_scale, uint32_t* scratch_dev, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (count <= 0) return 0;
  if (!param || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev || !scratch_dev) return fail("null argument");
  if (count & 3) return fail("count must be a multiple of 4 floats");
  PeerTable T;
  TRY(make_peer_table(T, peer_bases, world, rank));
  int64_t g64 = (count / 4 + 255) / 256;
  const int grid = int(g64 > 296 ? 296 : g64);  // every block spins in the in-barrier: keep the grid co-resident
  ProfScope prof(K_ADAM, st);
  adam_peer_kernel<<<grid, 256, 0, st>>>(param, T, exp_avg, exp_avg_sq, count, lr_dev, beta1, beta2, eps, weight_decay,
                                         reinterpret_cast<long long*>(step_dev), grad_scale, scratch_dev, peer_spin_limit());
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_peer_wait_done(void* const* peer_bases, int32_t world, int32_t rank, const int64_t* step_dev, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!step_dev) return fail("null argument");
  PeerTable T;
  TRY(make_peer_table(T, peer_bases, world, rank));
  ProfScope prof(K_PEER_WAIT, st);
  peer_wait_kernel<<<1, 32, 0, st>>>(T, reinterpret_cast<const long long*>(step_dev), peer_spin_limit());
  CU_OK(cudaGetLastError());
  return 0;
}

// ---- on-device coordinate pipeline and metrics (data_kernels.cuh) ----------------------------------------------------
static int make_grid(GridSpec& g, const int32_t* dims, int32_t ndim, int32_t linspace_kind) {
  if (!dims) return fail("null argument");
  if (ndim != 2 && ndim != 3) return fail("grid must be 2-D (H, W) or 3-D (H, W, T), got %d dims", ndim);
  if (linspace_kind != 0 && linspace_kind != 1) return fail("linspace_kind must be 0 (numpy float64) or 1 (torch float32)");
  g.ndim = ndim;
  g.linspace = linspace_kind;
  g.dims[2] = 1;
  for (int i = 0; i < ndim; ++i) {
    if (dims[i] < 1) return fail("grid dimension %d is %d", i, dims[i]);
    g.dims[i] = dims[i];
  }
  return 0;
}
static int grid_for(int64_t work) {
  int64_t g64 = (work + 255) / 256;
  return int(g64 > 2368 ? 2368 : (g64 < 1 ? 1 : g64));
}

int wire_grid_batch(const int32_t* dims, int32_t ndim, int32_t linspace_kind, const int64_t* idx, int64_t idx_base, int64_t n,
                    const float* signal, int32_t out_features, float* coords, float* target, int32_t* err_flag, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GridSpec g;
  TRY(make_grid(g, dims, ndim, linspace_kind));
  if (n <= 0) return 0;
  if (!coords && !target) return fail("nothing to produce: coords and target are both null");
  if (target && (!signal || out_features < 1)) return fail("target requested without a signal");
  ProfScope prof(K_DATA, st);
  grid_batch_kernel<<<grid_for(n), 256, 0, st>>>(g, idx, idx_base, n, signal, out_features, coords, target, err_flag);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_scatter_rows(const int64_t* idx, int64_t idx_base, int64_t n, const float* src, int32_t width, float* dst, int64_t dst_rows,
                      int32_t* err_flag, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n <= 0) return 0;
  if (!src || !dst || width < 1) return fail("null argument");
  ProfScope prof(K_DATA, st);
  scatter_rows_kernel<<<grid_for(n * width), 256, 0, st>>>(idx, idx_base, n, src, width, dst, dst_rows, err_flag);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_iou_counts(float* preds, const float* gt, int64_t count, float thres, int32_t use_thres, int32_t binarize_in_place,
                    uint64_t* counts, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (count <= 0) return 0;
  if (!preds || !gt || !counts) return fail("null argument");
  ProfScope prof(K_DATA, st);
  iou_counts_kernel<<<grid_for(count), 256, 0, st>>>(preds, gt, count, thres, use_thres, binarize_in_place,
                                                     reinterpret_cast<unsigned long long*>(counts));
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_sq_err_stats(const float* x, const float* xhat, int64_t count, double* stats, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (count <= 0) return 0;
  if (!x || !xhat || !stats) return fail("null argument");
  ProfScope prof(K_DATA, st);
  sq_err_stats_kernel<<<grid_for(count), 256, 0, st>>>(x, xhat, count, stats);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_avgpool_mse_loss_grad(const float* pred, const float* target_lr, int32_t H, int32_t W, int32_t channels, int32_t scale,
                               float* grad_out, float* loss, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!pred || !target_lr || !grad_out) return fail("null argument");
  if (H < 1 || W < 1 || channels < 1 || scale < 1) return fail("bad shape H=%d W=%d C=%d scale=%d", H, W, channels, scale);
  if (H / scale < 1 || W / scale < 1) return fail("scale %d larger than the image %dx%d", scale, H, W);
  if (H % scale || W % scale) CU_OK(cudaMemsetAsync(grad_out, 0, size_t(H) * W * channels * sizeof(float), st));
  ProfScope prof(K_MSE, st);
  avgpool_mse_grad_kernel<<<grid_for(int64_t(H / scale) * (W / scale) * channels), 256, 0, st>>>(pred, target_lr, H, W, channels, scale,
                                                                                                grad_out, loss);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_gabor_scalar_grads(int32_t is_first, int32_t two_d, int32_t width, const float* z_save, const float* w_save, const float* grad_y,
                            int64_t n, const float* omega0, const float* scale0, double* out2, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!z_save || !grad_y || !omega0 || !scale0 || !out2) return fail("null argument");
  if (two_d && !w_save) return fail("wire2d layer without its saved scale_orth pre-activation");
  if (n <= 0 || width <= 0) return 0;
  ProfScope prof(K_LAYER_MISC, st);
  gabor_scalar_grads_kernel<<<grid_for(n * width), 256, 0, st>>>(z_save, two_d ? w_save : nullptr, grad_y, n * int64_t(width), is_first,
                                                                 omega0, scale0, out2);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_real_gabor_layer_forward(const float* x, int64_t n, int32_t K, int32_t M, const float* w_freqs, const float* b_freqs,
                                  const float* w_scale, const float* b_scale, float omega0, float scale0, float* y, float* f_save,
                                  float* s_save, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (n <= 0) return 0;
  if (!x || !w_freqs || !w_scale || !y) return fail("null argument");
  if (K < 1 || M < 1) return fail("bad shape K=%d M=%d", K, M);
  ProfScope prof(K_LAYER_MISC, st);
  dim3 grid(unsigned((n + 63) / 64), unsigned((M + 63) / 64));
  real_gabor_layer_fwd_kernel<<<grid, 256, 0, st>>>(x, int(n), K, M, w_freqs, b_freqs, w_scale, b_scale, omega0, scale0, y, f_save, s_save);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_real_gabor_layer_backward(const float* x, const float* f_save, const float* s_save, const float* grad_y, int64_t n, int32_t K,
                                   int32_t M, const float* w_freqs, const float* w_scale, float omega0, float scale0, float* grad_x,
                                   float* g_w_freqs, float* g_b_freqs, float* g_w_scale, float* g_b_scale, float* scratch_gf,
                                   float* scratch_gs, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (!x || !f_save || !s_save || !grad_y || !w_freqs || !w_scale || !g_w_freqs || !g_w_scale || !scratch_gf || !scratch_gs)
    return fail("null argument");
  if (K < 1 || M < 1) return fail("bad shape K=%d M=%d", K, M);
  TRY(zero(g_w_freqs, size_t(M) * K, st)); TRY(zero(g_w_scale, size_t(M) * K, st));
  TRY(zero(g_b_freqs, size_t(M), st)); TRY(zero(g_b_scale, size_t(M), st));
  if (n <= 0) return 0;
  ProfScope prof(K_LAYER_MISC, st, grad_x ? 4 : 3);
  real_gabor_bwd_kernel<<<grid_for(n * M), 256, 0, st>>>(f_save, s_save, grad_y, n * int64_t(M), omega0, scale0, scratch_gf, scratch_gs);
  if (grad_x) {
    dim3 gd(unsigned((n + 63) / 64), unsigned((K + 63) / 64));
    real_gabor_layer_dgrad_kernel<<<gd, 256, 0, st>>>(scratch_gf, scratch_gs, int(n), K, M, w_freqs, w_scale, grad_x);
  }
  dim3 gw(unsigned((M + 63) / 64), unsigned((K + 63) / 64), 1);
  int splits = (4 * g_sm_count) / int(gw.x * gw.y);
  if (splits < 1) splits = 1;
  int rps = int((n + splits - 1) / splits);
  rps = (rps + 15) / 16 * 16;
  gw.z = unsigned((n + rps - 1) / rps);
  real_gabor_layer_wgrad_kernel<<<gw, 256, 0, st>>>(scratch_gf, x, int(n), K, M, rps, g_w_freqs, g_b_freqs);
  real_gabor_layer_wgrad_kernel<<<gw, 256, 0, st>>>(scratch_gs, x, int(n), K, M, rps, g_w_scale, g_b_scale);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_radon_forward(const float* image, int32_t nimg, int32_t H, int32_t W, const float* angles_deg, int32_t nangles, float* sinogram,
                       void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (!image || !angles_deg || !sinogram) return fail("null argument");
  if (nimg < 1 || H < 1 || W < 1 || nangles < 1) return fail("bad shape nimg=%d H=%d W=%d nangles=%d", nimg, H, W, nangles);
  ProfScope prof(K_DATA, st);
  radon_fwd_kernel<<<grid_for(int64_t(nangles) * nimg * W), 256, 0, st>>>(image, nimg, H, W, angles_deg, nangles, sinogram);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_radon_backward(const float* grad_sinogram, int32_t nimg, int32_t H, int32_t W, const float* angles_deg, int32_t nangles,
                        float* grad_image, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (!grad_sinogram || !angles_deg || !grad_image) return fail("null argument");
  if (nimg < 1 || H < 1 || W < 1 || nangles < 1) return fail("bad shape nimg=%d H=%d W=%d nangles=%d", nimg, H, W, nangles);
  CU_OK(cudaMemsetAsync(grad_image, 0, size_t(nimg) * H * W * sizeof(float), st));
  const int rpt = 16;  // rows per thread: enough threads to fill the machine at 100 angles x 256 columns
  ProfScope prof(K_DATA, st);
  radon_bwd_kernel<<<grid_for(int64_t(nangles) * nimg * ((H + rpt - 1) / rpt) * W), 256, 0, st>>>(grad_sinogram, nimg, H, W, angles_deg, nangles,
                                                                                                 rpt, grad_image);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_real_gabor_forward(const float* f, const float* s, int64_t count, float omega0, float scale0, float* y, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (count <= 0) return 0;
  if (!f || !s || !y) return fail("null argument");
  ProfScope prof(K_LAYER_MISC, st);
  real_gabor_fwd_kernel<<<grid_for(count), 256, 0, st>>>(f, s, count, omega0, scale0, y);
  CU_OK(cudaGetLastError());
  return 0;
}

int wire_real_gabor_backward(const float* f, const float* s, const float* grad_y, int64_t count, float omega0, float scale0, float* grad_f,
                             float* grad_s, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TRY(require_device());
  if (count <= 0) return 0;
  if (!f || !s || !grad_y || !grad_f || !grad_s) return fail("null argument");
  ProfScope prof(K_LAYER_MISC, st);
  real_gabor_bwd_kernel<<<grid_for(count), 256, 0, st>>>(f, s, grad_y, count, omega0, scale0, grad_f, grad_s);
  CU_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
