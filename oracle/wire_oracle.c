/* wire_oracle.c — plain-C restatement of the WIRE hot path (TEST INFRASTRUCTURE, not product code).
 *
 * Double precision, scalar loops, no dependencies.  Follows
 *   modules/wire.py:88-93,161-165     (ComplexGaborLayer.forward, INR.forward)
 *   modules/wire2d.py:56-67,121-125   (ComplexGaborLayer2D.forward, INR.forward)
 * and, for the backward pass, the closed form of PyTorch's complex autograd:
 *   p = conj(y) g_y ;  g_z = -j w0 p - 2 s0^2 z Re p ;  g_w = -2 s0^2 w Re p
 *   first layer (real z): g_z = w0 Im p - 2 s0^2 z Re p
 *   Linear: g_x = g_z conj(W) ; g_W = g_z^T conj(x) ; g_b = sum_n g_z ; final: upstream grad is real.
 *
 * Pinned by tests/test_oracle.py against tests/golden/*.npz (outputs of the reference itself).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use it.
 *
 * Parameter / gradient layout (flat double arrays, complex = (re, im) pairs), per layer l = 0..H:
 *   W[M][K] b[M] (then W2[M][K] b2[M] for wire2d)   with K = in (real) for l = 0, K = M (complex) else
 * followed by Wf[out][M] (complex), bf[out] (complex).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cplx;

static size_t layer_doubles(int two_d, int l, int in_f, int M) {
  size_t w = (l == 0) ? (size_t)M * in_f + M : 2 * ((size_t)M * M + M);
  return two_d ? 2 * w : w;
}

size_t wire_oracle_param_doubles(int two_d, int in_f, int M, int H, int out_f) {
  size_t t = 0;
  for (int l = 0; l <= H; ++l) t += layer_doubles(two_d, l, in_f, M);
  return t + 2 * ((size_t)out_f * M + out_f);
}

/* out[n][out_f]; if grad_out != NULL also fills grad_params (same layout as params) and grad_coords[n][in_f]
 * (may be NULL).  Returns 0 on success. */
int wire_oracle_run(int two_d, int n, int in_f, int M, int H, int out_f, const double* coords, const double* params,
                    const double* omega, const double* scale, const double* grad_out, double* out, double* grad_params,
                    double* grad_coords) {
  const size_t NM = (size_t)n * M;
  cplx* y = (cplx*)malloc(sizeof(cplx) * NM * (H + 1));
  cplx* z = (cplx*)malloc(sizeof(cplx) * NM * (H + 1));
  cplx* w = (cplx*)calloc(NM * (H + 1), sizeof(cplx));
  cplx* gy = (cplx*)malloc(sizeof(cplx) * NM);
  cplx* gx = (cplx*)malloc(sizeof(cplx) * NM);
  if (!y || !z || !w || !gy || !gx) return 1;
  const double* lp[32];
  const double* p = params;
  for (int l = 0; l <= H; ++l) { lp[l] = p; p += layer_doubles(two_d, l, in_f, M); }
  const double* Wf = p;
  const double* bf = p + 2 * (size_t)out_f * M;

  /* ---------------- forward ---------------- */
  for (int l = 0; l <= H; ++l) {
    const int K = l == 0 ? in_f : M;
    const int cw = l == 0 ? 1 : 2; /* doubles per weight entry */
    const double* W = lp[l];
    const double* b = W + (size_t)M * K * cw;
    const double* W2 = b + (size_t)M * cw;
    const double* b2 = W2 + (size_t)M * K * cw;
    const double s2 = scale[l] * scale[l], om = omega[l];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < M; ++j) {
        cplx acc = {0, 0}, acc2 = {0, 0};
        if (l == 0) {
          acc.re = b[j];
          if (two_d) acc2.re = b2[j];
          for (int k = 0; k < K; ++k) {
            acc.re += coords[(size_t)i * in_f + k] * W[(size_t)j * K + k];
            if (two_d) acc2.re += coords[(size_t)i * in_f + k] * W2[(size_t)j * K + k];
          }
        } else {
          acc.re = b[2 * j]; acc.im = b[2 * j + 1];
          if (two_d) { acc2.re = b2[2 * j]; acc2.im = b2[2 * j + 1]; }
          const cplx* x = y + (size_t)(l - 1) * NM + (size_t)i * M;
          for (int k = 0; k < K; ++k) {
            const double wr = W[2 * ((size_t)j * K + k)], wi = W[2 * ((size_t)j * K + k) + 1];
            acc.re += x[k].re * wr - x[k].im * wi;
            acc.im += x[k].re * wi + x[k].im * wr;
            if (two_d) {
              const double vr = W2[2 * ((size_t)j * K + k)], vi = W2[2 * ((size_t)j * K + k) + 1];
              acc2.re += x[k].re * vr - x[k].im * vi;
              acc2.im += x[k].re * vi + x[k].im * vr;
            }
          }
        }
        const size_t o = (size_t)l * NM + (size_t)i * M + j;
        z[o] = acc;
        w[o] = acc2;
        const double mag = exp(-om * acc.im - s2 * (acc.re * acc.re + acc.im * acc.im + acc2.re * acc2.re + acc2.im * acc2.im));
        y[o].re = mag * cos(om * acc.re);
        y[o].im = mag * sin(om * acc.re);
      }
  }
  const cplx* h = y + (size_t)H * NM;
  for (int i = 0; i < n; ++i)
    for (int o = 0; o < out_f; ++o) {
      double acc = bf[2 * o];
      for (int k = 0; k < M; ++k)
        acc += h[(size_t)i * M + k].re * Wf[2 * ((size_t)o * M + k)] - h[(size_t)i * M + k].im * Wf[2 * ((size_t)o * M + k) + 1];
      out[(size_t)i * out_f + o] = acc;
    }
  if (!grad_out) { free(y); free(z); free(w); free(gy); free(gx); return 0; }

  /* ---------------- backward ---------------- */
  memset(grad_params, 0, sizeof(double) * wire_oracle_param_doubles(two_d, in_f, M, H, out_f));
  double* gp[32];
  double* q = grad_params;
  for (int l = 0; l <= H; ++l) { gp[l] = q; q += layer_doubles(two_d, l, in_f, M); }
  double* gWf = q;
  double* gbf = q + 2 * (size_t)out_f * M;
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < M; ++k) { gy[(size_t)i * M + k].re = 0; gy[(size_t)i * M + k].im = 0; }
    for (int o = 0; o < out_f; ++o) {
      const double g = grad_out[(size_t)i * out_f + o];
      gbf[2 * o] += g;
      for (int k = 0; k < M; ++k) {
        const cplx hv = h[(size_t)i * M + k];
        gWf[2 * ((size_t)o * M + k)] += g * hv.re;      /* g * conj(h) */
        gWf[2 * ((size_t)o * M + k) + 1] += -g * hv.im;
        gy[(size_t)i * M + k].re += g * Wf[2 * ((size_t)o * M + k)];       /* g * conj(Wf) */
        gy[(size_t)i * M + k].im += -g * Wf[2 * ((size_t)o * M + k) + 1];
      }
    }
  }
  for (int l = H; l >= 0; --l) {
    const int K = l == 0 ? in_f : M;
    const int cw = l == 0 ? 1 : 2;
    const double* W = lp[l];
    const double* W2 = W + (size_t)M * K * cw + (size_t)M * cw;
    double* gW = gp[l];
    double* gb = gW + (size_t)M * K * cw;
    double* gW2 = gb + (size_t)M * cw;
    double* gb2 = gW2 + (size_t)M * K * cw;
    const double s2 = scale[l] * scale[l], om = omega[l];
    if (l > 0) for (size_t t = 0; t < NM; ++t) { gx[t].re = 0; gx[t].im = 0; }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < M; ++j) {
        const size_t o = (size_t)l * NM + (size_t)i * M + j;
        const cplx yv = y[o], zv = z[o], wv = w[o], g = gy[(size_t)i * M + j];
        const double pr = yv.re * g.re + yv.im * g.im, pi = yv.re * g.im - yv.im * g.re;
        if (l == 0) {
          const double gz = om * pi - 2.0 * s2 * zv.re * pr;
          const double gw = -2.0 * s2 * wv.re * pr;
          gb[j] += gz;
          if (two_d) gb2[j] += gw;
          for (int k = 0; k < K; ++k) {
            const double c = coords[(size_t)i * in_f + k];
            gW[(size_t)j * K + k] += gz * c;
            if (two_d) gW2[(size_t)j * K + k] += gw * c;
            if (grad_coords) grad_coords[(size_t)i * in_f + k] += gz * W[(size_t)j * K + k] + (two_d ? gw * W2[(size_t)j * K + k] : 0.0);
          }
        } else {
          const cplx gz = {om * pi - 2.0 * s2 * zv.re * pr, -om * pr - 2.0 * s2 * zv.im * pr};
          const cplx gw = {-2.0 * s2 * wv.re * pr, -2.0 * s2 * wv.im * pr};
          gb[2 * j] += gz.re; gb[2 * j + 1] += gz.im;
          if (two_d) { gb2[2 * j] += gw.re; gb2[2 * j + 1] += gw.im; }
          const cplx* x = y + (size_t)(l - 1) * NM + (size_t)i * M;
          for (int k = 0; k < K; ++k) {
            const size_t wi = 2 * ((size_t)j * K + k);
            gW[wi] += gz.re * x[k].re + gz.im * x[k].im;        /* g_z * conj(x) */
            gW[wi + 1] += gz.im * x[k].re - gz.re * x[k].im;
            gx[(size_t)i * M + k].re += gz.re * W[wi] + gz.im * W[wi + 1];   /* g_z * conj(W) */
            gx[(size_t)i * M + k].im += gz.im * W[wi] - gz.re * W[wi + 1];
            if (two_d) {
              gW2[wi] += gw.re * x[k].re + gw.im * x[k].im;
              gW2[wi + 1] += gw.im * x[k].re - gw.re * x[k].im;
              gx[(size_t)i * M + k].re += gw.re * W2[wi] + gw.im * W2[wi + 1];
              gx[(size_t)i * M + k].im += gw.im * W2[wi] - gw.re * W2[wi + 1];
            }
          }
        }
      }
    if (l > 0) memcpy(gy, gx, sizeof(cplx) * NM);
  }
  free(y); free(z); free(w); free(gy); free(gx);
  return 0;
}
