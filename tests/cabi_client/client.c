/* client.c -- a torch-free, plain-C client of libtsff's C ABI (include/tsff.h), driving the calls exactly as an XLA FFI
 * handler would: caller-owned device buffers (cudaMalloc), results pre-allocated, everything enqueued on
 * a caller-created non-blocking stream, no host synchronisation between the calls:
 *
 *     tsff_ctx_create -> tsff_ff_fwd -> tsff_loss_fwd_bwd -> tsff_ff_bwd          (one stream, one sync at the end)
 *
 * Inputs and expected outputs come from tests/golden/cabi_fixture.bin, written by the float64 ORACLE
 * (tools/make_cabi_fixture.py).  Bars: spectrum <= 1e-5, gradients <= 1e-4 (BASELINE.json north_star).
 *
 *     gcc -O2 -o client client.c -I<repo>/include -I/usr/local/cuda/include -L/usr/local/cuda/lib64 -lcudart -ldl -lm
 *     ./client <libtsff.so> <cabi_fixture.bin>            exit code 0 = parity holds
 */
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tsff.h"

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } \
  } while (0)

typedef int (*ctx_create_t)(int, const tsff_static_cfg*, tsff_ctx**);
typedef void (*ctx_destroy_t)(tsff_ctx*);
typedef const char* (*last_error_t)(void);
typedef size_t (*bytes_t)(const tsff_ctx*, int64_t);
typedef int (*ff_fwd_t)(tsff_ctx*, int64_t, const double*, const void*, int, double*, double*, void*, void*, void*);
typedef int (*ff_bwd_t)(tsff_ctx*, int64_t, const double*, const void*, int, const void*, const double*, const double*, double*, void*,
                        void*, void*);
typedef int (*loss_t)(int64_t, int32_t, const double*, const double*, const double*, double, double, int, double*, double*, void*);

static void* rd(FILE* f, size_t n) {
  void* p = malloc(n);
  if (!p || fread(p, 1, n, f) != n) { fprintf(stderr, "short fixture\n"); exit(2); }
  return p;
}
static double maxrel(const double* got, const double* ref, size_t n) {
  double mx = 0.0, d = 0.0;
  for (size_t i = 0; i < n; i++) {
    if (fabs(ref[i]) > mx) mx = fabs(ref[i]);
    if (fabs(got[i] - ref[i]) > d) d = fabs(got[i] - ref[i]);
  }
  return d / mx;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s libtsff.so fixture.bin\n", argv[0]); return 2; }
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  ctx_create_t ctx_create = (ctx_create_t)dlsym(h, "tsff_ctx_create");
  ctx_destroy_t ctx_destroy = (ctx_destroy_t)dlsym(h, "tsff_ctx_destroy");
  last_error_t last_error = (last_error_t)dlsym(h, "tsff_last_error");
  bytes_t saved_bytes = (bytes_t)dlsym(h, "tsff_ff_saved_bytes"), ws_bytes = (bytes_t)dlsym(h, "tsff_ff_workspace_bytes");
  ff_fwd_t ff_fwd = (ff_fwd_t)dlsym(h, "tsff_ff_fwd");
  ff_bwd_t ff_bwd = (ff_bwd_t)dlsym(h, "tsff_ff_bwd");
  loss_t loss_fwd_bwd = (loss_t)dlsym(h, "tsff_loss_fwd_bwd");
  if (!ctx_create || !ctx_destroy || !last_error || !saved_bytes || !ws_bytes || !ff_fwd || !ff_bwd || !loss_fwd_bwd) {
    fprintf(stderr, "missing symbol\n");
    return 2;
  }

  FILE* f = fopen(argv[2], "rb");
  if (!f) { perror(argv[2]); return 2; }
  char magic[8];
  int32_t hd[8];
  double sc[8];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "TSFFFIX1", 8) || fread(hd, 4, 8, f) != 8 || fread(sc, 8, 8, f) != 8) {
    fprintf(stderr, "bad fixture header\n");
    return 2;
  }
  const int B = hd[0], W = hd[1], V = hd[2], NP = hd[3], zn = hd[4];
  double* zx = rd(f, 8 * (size_t)zn); double* zr = rd(f, 8 * (size_t)zn); double* zi = rd(f, 8 * (size_t)zn);
  double* params = rd(f, 8 * (size_t)B * NP);
  float* fe = rd(f, 4 * (size_t)B * V);
  double* target = rd(f, 8 * (size_t)B * W);
  double* wq = rd(f, 8 * (size_t)W);
  double* x_modl = rd(f, 8 * (size_t)B * W);
  double* x_loss = rd(f, 8);
  double* x_pbar = rd(f, 8 * (size_t)B * NP);
  double* x_fbar = rd(f, 8 * (size_t)B * V);
  fclose(f);

  /* FormFactor.__init__ (form_factor.py:120-161) */
  double sa = sc[4], wt = sc[5];
  tsff_static_cfg cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = TSFF_ABI_VERSION; cfg.mode = TSFF_MODE_DIRECT;
  cfg.W = W; cfg.A = 1; cfg.G = 1; cfg.I = 1; cfg.V = V; cfg.pv_precision = TSFF_PV_FP32;
  cfg.lam_min = sc[0]; cfg.lam_max = sc[1]; cfg.lam_shift = 0.0; cfg.v0 = sc[2]; cfg.dv = sc[3];
  cfg.sa_deg = &sa; cfg.weights = &wt; cfg.jmul = NULL;
  cfg.zp_x = zx; cfg.zp_re = zr; cfg.zp_im = zi; cfg.zp_n = zn;
  tsff_ctx* ctx = NULL;
  int dev = 0;
  CK(cudaSetDevice(dev));
  if (ctx_create(dev, &cfg, &ctx)) { fprintf(stderr, "tsff_ctx_create: %s\n", last_error()); return 2; }

  /* caller-owned buffers, as XLA would hand them to an FFI handler */
  double *d_params, *d_target, *d_wq, *d_modl, *d_loss, *d_tbar, *d_pbar;
  float *d_fe, *d_fbar;
  void *d_saved, *d_ws;
  CK(cudaMalloc((void**)&d_params, 8 * (size_t)B * NP)); CK(cudaMalloc((void**)&d_fe, 4 * (size_t)B * V));
  CK(cudaMalloc((void**)&d_target, 8 * (size_t)B * W)); CK(cudaMalloc((void**)&d_wq, 8 * (size_t)W));
  CK(cudaMalloc((void**)&d_modl, 8 * (size_t)B * W)); CK(cudaMalloc((void**)&d_loss, 8));
  CK(cudaMalloc((void**)&d_tbar, 8 * (size_t)B * W)); CK(cudaMalloc((void**)&d_pbar, 8 * (size_t)B * NP));
  CK(cudaMalloc((void**)&d_fbar, 4 * (size_t)B * V));
  CK(cudaMalloc(&d_saved, saved_bytes(ctx, B))); CK(cudaMalloc(&d_ws, ws_bytes(ctx, B)));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaMemcpyAsync(d_params, params, 8 * (size_t)B * NP, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_fe, fe, 4 * (size_t)B * V, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_target, target, 8 * (size_t)B * W, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_wq, wq, 8 * (size_t)W, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d_loss, 0, 8, st));

  /* the three calls, back to back on the user stream, no synchronisation in between */
  if (ff_fwd(ctx, B, d_params, d_fe, TSFF_F32, d_modl, NULL, d_saved, d_ws, st)) { fprintf(stderr, "tsff_ff_fwd: %s\n", last_error()); return 2; }
  if (loss_fwd_bwd(B, W, d_modl, d_target, d_wq, sc[6], sc[7], 0, d_loss, d_tbar, st)) { fprintf(stderr, "tsff_loss_fwd_bwd: %s\n", last_error()); return 2; }
  if (ff_bwd(ctx, B, d_params, d_fe, TSFF_F32, d_saved, d_tbar, NULL, d_pbar, d_fbar, d_ws, st)) { fprintf(stderr, "tsff_ff_bwd: %s\n", last_error()); return 2; }

  double* modl = malloc(8 * (size_t)B * W); double* pbar = malloc(8 * (size_t)B * NP);
  float* fbar32 = malloc(4 * (size_t)B * V); double* fbar = malloc(8 * (size_t)B * V);
  double loss = 0.0;
  CK(cudaMemcpyAsync(modl, d_modl, 8 * (size_t)B * W, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(pbar, d_pbar, 8 * (size_t)B * NP, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(fbar32, d_fbar, 4 * (size_t)B * V, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&loss, d_loss, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (size_t i = 0; i < (size_t)B * V; i++) fbar[i] = (double)fbar32[i];

  int bad = 0;
  for (int b = 0; b < B; b++) {
    const double es = maxrel(modl + (size_t)b * W, x_modl + (size_t)b * W, W);
    const double ef = maxrel(fbar + (size_t)b * V, x_fbar + (size_t)b * V, V);
    printf("lineout %d: spectrum %.2e  fe_bar %.2e  params_bar", b, es, ef);
    if (!(es <= 1e-5) || !(ef <= 1e-4)) bad = 1;
    const int act[] = {TSFF_P_TE, TSFF_P_NE, TSFF_P_LAM, TSFF_P_ION0 + TSFF_ION_Z, TSFF_P_ION0 + TSFF_ION_TI};
    double rmax = 0.0;   /* a gradient component is compared relative to its own size, floored at 1e-6 of the row's largest */
    for (int k = 0; k < 5; k++) rmax = fmax(rmax, fabs(x_pbar[b * NP + act[k]]));
    for (int k = 0; k < 5; k++) {
      const double g = pbar[b * NP + act[k]], r = x_pbar[b * NP + act[k]];
      const double e = fabs(g - r) / fmax(fabs(r), 1e-6 * rmax);
      printf(" %.1e", e);
      if (!(e <= 1e-4)) { bad = 1; printf(" [got %.6e, oracle %.6e]", g, r); }
    }
    printf("\n");
  }
  const double el = fabs(loss - x_loss[0]) / fabs(x_loss[0]);
  printf("loss %.12g (oracle %.12g, rel %.1e)\n", loss, x_loss[0], el);
  if (!(el <= 1e-5)) bad = 1;
  ctx_destroy(ctx);
  cudaStreamDestroy(st);
  printf(bad ? "FAIL\n" : "PASS: torch-free C client matches the oracle fixture\n");
  return bad;
}
