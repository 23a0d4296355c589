// tsff_ctx.cu -- context creation (static tables of FormFactor.__init__, form_factor.py:120-161), error text,
// roofline microbenchmarks.
#include <stdarg.h>

#include <vector>

#include <stdlib.h>

#include "tsff_pv_kernels.cuh"

namespace tsff {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace tsff

using namespace tsff;

extern "C" const char* tsff_last_error(void) { return tsff::g_err; }
extern "C" int tsff_abi_version(void) { return TSFF_ABI_VERSION; }

// scipy.interpolate.interp1d(x, y, "linear")(xq) on an ascending table (form_factor.py:39-42)
static double table_lerp(const double* x, const double* y, int n, double xq) {
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2;
    if (x[mid] <= xq) lo = mid; else hi = mid;
  }
  double t = (xq - x[lo]) / (x[lo + 1] - x[lo]);
  return y[lo] + t * (y[lo + 1] - y[lo]);
}

extern "C" int tsff_ctx_create(int device, const tsff_static_cfg* cfg, tsff_ctx** out) {
  if (!cfg || !out) { set_error("null argument"); return TSFF_E_INVALID; }
  if (cfg->abi_version != TSFF_ABI_VERSION) { set_error("ABI version mismatch: %d vs %d", cfg->abi_version, TSFF_ABI_VERSION); return TSFF_E_INVALID; }
  if (cfg->W < 2 || cfg->A < 1 || cfg->G < 1 || cfg->I < 1 || cfg->I > TSFF_MAX_IONS || cfg->V < 4) {
    set_error("unsupported shape W=%d A=%d G=%d I=%d V=%d", cfg->W, cfg->A, cfg->G, cfg->I, cfg->V);
    return TSFF_E_INVALID;
  }
  if (cfg->mode != TSFF_MODE_TABLE && cfg->mode != TSFF_MODE_DIRECT && cfg->mode != TSFF_MODE_2V) { set_error("unknown mode %d", cfg->mode); return TSFF_E_INVALID; }
  if (!cfg->sa_deg || !cfg->weights || !cfg->zp_x || !cfg->zp_re || !cfg->zp_im || cfg->zp_n < 2) {
    set_error("missing static table pointer"); return TSFF_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) {
    set_error("no CUDA device %d (libtsff has no CPU fallback)", device);
    return TSFF_E_NODEVICE;
  }
  cudaDeviceProp prop;
  TSFF_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { set_error("device %d is sm_%d%d; libtsff is built for sm_100a only", device, prop.major, prop.minor); return TSFF_E_NODEVICE; }
  DeviceGuard guard(device);   // the caller's current device is restored on every return path
  if (!guard.ok) { set_error("cannot make device %d current", device); return TSFF_E_CUDA; }

  tsff_ctx* c = new tsff_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->tune_fwd_r4 = getenv("TSFF_FWD_R4") != nullptr;
  c->sm_count = prop.multiProcessorCount;
  c->mode = cfg->mode; c->W = cfg->W; c->A = cfg->A; c->G = cfg->G; c->I = cfg->I; c->V = cfg->V;
  c->NP = TSFF_P_ION0 + TSFF_ION_STRIDE * cfg->I;
  c->pv_precision = cfg->pv_precision;
  c->lam_min = cfg->lam_min; c->lam_max = cfg->lam_max; c->lam_shift = cfg->lam_shift;
  c->v0 = cfg->v0; c->dv = cfg->dv;
  c->ud_angle_deg = cfg->ud_angle_deg; c->va_angle_deg = cfg->va_angle_deg;

  const int W = c->W, A = c->A;
  std::vector<double> h_omgs(W), h_lam(W), h_cos(A), h_sin(A), h_w(A), h_jmul(W), h_zr(kXi2N), h_zi(kXi2N), h_xi2(kXi2N);
  // jnp.linspace(l0, l1, W): start + i*step, endpoint exact
  const int Wt = cfg->W_total > 0 ? cfg->W_total : W, w0 = cfg->W_total > 0 ? cfg->w_offset : 0;
  if (w0 < 0 || w0 + W > Wt) { delete c; set_error("wavelength shard [%d, %d) outside the axis of %d points", w0, w0 + W, Wt); return TSFF_E_INVALID; }
  const double step = (c->lam_max - c->lam_min) / (double)(Wt - 1);
  for (int j = 0; j < W; j++) {
    const int jg = w0 + j;
    double lam = (jg == Wt - 1) ? c->lam_max : c->lam_min + (double)jg * step;
    h_omgs[j] = 2e7 * kPi * kC / lam;                 // form_factor.py:134
    h_lam[j] = (2.0 * kPi * kC / h_omgs[j]) * 1e7;    // lams (form_factor.py:293) * 1e7 (generate_spectra.py:163,191)
    h_jmul[j] = cfg->jmul ? cfg->jmul[j] : 1.0;
  }
  for (int a = 0; a < A; a++) {
    h_cos[a] = cos(cfg->sa_deg[a] * kPi / 180.0);    // form_factor.py:210,220
    h_sin[a] = sin(cfg->sa_deg[a] * kPi / 180.0);    // form_factor.py:514
    h_w[a] = cfg->weights[a];
  }
  for (int i = 0; i < kXi2N; i++) {
    h_xi2[i] = -kXiMinMax + (double)i * 0.01;         // jnp.arange(-8.2, 8.2, 0.01)  form_factor.py:138
    h_zr[i] = table_lerp(cfg->zp_x, cfg->zp_re, cfg->zp_n, h_xi2[i]);
    h_zi[i] = table_lerp(cfg->zp_x, cfg->zp_im, cfg->zp_n, h_xi2[i]);
  }
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += align_up(n * sizeof(double)); return o; };
  size_t o_omgs = take(W), o_lam = take(W), o_cos = take(A), o_sin = take(A), o_w = take(A), o_jmul = take(W), o_zr = take(kXi2N),
         o_zi = take(kXi2N), o_xi2 = take(kXi2N);
  if (cudaMalloc(&c->dev_blob, off) != cudaSuccess) { delete c; set_error("cudaMalloc(%zu) failed", off); return TSFF_E_NOMEM; }
  char* base = static_cast<char*>(c->dev_blob);
  auto up = [&](size_t o, const std::vector<double>& v) { return cudaMemcpy(base + o, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice); };
  if (up(o_omgs, h_omgs) || up(o_lam, h_lam) || up(o_cos, h_cos) || up(o_sin, h_sin) || up(o_w, h_w) || up(o_jmul, h_jmul) || up(o_zr, h_zr) ||
      up(o_zi, h_zi) || up(o_xi2, h_xi2)) {
    cudaFree(c->dev_blob); delete c; set_error("table upload failed"); return TSFF_E_CUDA;
  }
  c->omgs = (double*)(base + o_omgs); c->lam_nm = (double*)(base + o_lam); c->costh = (double*)(base + o_cos);
  c->sinth = (double*)(base + o_sin);
  c->wts = (double*)(base + o_w); c->jmul = (double*)(base + o_jmul); c->zr = (double*)(base + o_zr);
  c->zi = (double*)(base + o_zi); c->xi2 = (double*)(base + o_xi2);
  c->zt.zr = c->zr; c->zt.zi = c->zi; c->zt.n = kXi2N; c->zt.x0 = h_xi2[0]; c->zt.h = 0.01; c->zt.xlast = h_xi2[kXi2N - 1];
  // xi1 = linspace(-(8.2 + sqrt2/1024), +(8.2 + sqrt2/1024), 1024)   form_factor.py:137
  const double e = kXiMinMax + sqrt(2.0) / (double)kXi1N;
  c->xi1_0 = -e;
  c->xi1_h = (2.0 * e) / (double)(kXi1N - 1);
  if (c->mode == TSFF_MODE_TABLE) {
    c->pv_nodes = kXi1N - 1; c->pv_z0 = c->xi1_0; c->pv_h = c->xi1_h;
  } else {
    c->pv_nodes = c->V - 1; c->pv_z0 = c->v0; c->pv_h = c->dv;
  }
  c->pv_npad = tree_npad(c->pv_nodes);
  if (c->pv_npad > kTreeMaxNpad) {
    cudaFree(c->dev_blob); delete c; set_error("f-table too long (%d nodes; limit %d)", cfg->V, kTreeMaxNpad); return TSFF_E_INVALID;
  }
  if (cudaMalloc(&c->tstat, kTreeStaticDoubles * sizeof(double)) != cudaSuccess) {
    cudaFree(c->dev_blob); delete c; set_error("cudaMalloc failed"); return TSFF_E_NOMEM;
  }
  k_tree_static<<<kTreeStaticGrid, 256>>>(c->pv_nodes - 1, c->tstat);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(c->tstat); cudaFree(c->dev_blob); delete c; set_error("k_tree_static failed: %s", cudaGetErrorString(cudaGetLastError()));
    return TSFF_E_CUDA;
  }
  *out = c;
  return TSFF_OK;
}

extern "C" int tsff_ctx_set_profile_events(tsff_ctx* ctx, void* e0, void* e1, void* e2, void* e3) {
  if (!ctx) { set_error("null ctx"); return TSFF_E_INVALID; }
  ctx->ev[0] = static_cast<cudaEvent_t>(e0); ctx->ev[1] = static_cast<cudaEvent_t>(e1);
  ctx->ev[2] = static_cast<cudaEvent_t>(e2); ctx->ev[3] = static_cast<cudaEvent_t>(e3);
  return TSFF_OK;
}

extern "C" void tsff_ctx_destroy(tsff_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard guard(ctx->device);
  cudaFree(ctx->tstat);
  cudaFree(ctx->dev_blob);
  delete ctx;
}

// ---- roofline microbenchmarks (FFMA / MUFU.LG2 issue peaks), SURVEY.md 8(d) ------------------------------
__global__ void __launch_bounds__(256) k_micro_ffma(long long iters, float* sink) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 1e-3f;
#pragma unroll 16
  for (long long i = 0; i < iters; i++) {
    a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
    a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
  }
  float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678f) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_micro_lg2(long long iters, float* sink) {
  float a0 = 1.5f + threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
#pragma unroll 16
  for (long long i = 0; i < iters; i++) {
    a0 = lg2_approx(a0) + 4.f; a1 = lg2_approx(a1) + 4.f; a2 = lg2_approx(a2) + 4.f; a3 = lg2_approx(a3) + 4.f;
  }
  float s = a0 + a1 + a2 + a3;
  if (s == 12345.678f) sink[0] = s;
}

// mixed loop: one MUFU.RCP feeding NF independent FFMAs per step (machine model for the pole/node sweeps)
template <int NF>
__global__ void __launch_bounds__(256) k_micro_mix(long long iters, float* sink) {
  float x0 = 1.5f + threadIdx.x * 1e-3f, x1 = x0 + 0.25f;
  float a[12];
#pragma unroll
  for (int k = 0; k < 12; k++) a[k] = 0.1f * k;
  const float c = 1.0009765625f;
#pragma unroll 4
  for (long long i = 0; i < iters; i++) {
    const float r0 = rcp_approx(x0), r1 = rcp_approx(x1);
#pragma unroll
    for (int k = 0; k < NF; k++) a[k] = fmaf(a[k], c, (k & 1) ? r1 : r0);
#pragma unroll
    for (int k = 0; k < NF; k++) a[(k + 1) % 12] = fmaf(a[(k + 1) % 12], c, (k & 1) ? r0 : r1);
    x0 += 0.001f; x1 += 0.001f;
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 12; k++) s += a[k];
  if (s == 12345.678f) sink[0] = s;
}

// packed variant: one MUFU.RCP per NF/2 FFMA2 (= NF FMAs)
template <int NF2>
__global__ void __launch_bounds__(256) k_micro_mix2(long long iters, float* sink) {
  float x0 = 1.5f + threadIdx.x * 1e-3f, x1 = x0 + 0.25f;
  float2 a[6];
#pragma unroll
  for (int k = 0; k < 6; k++) a[k] = make_float2(0.1f * k, 0.2f * k);
  const float2 c = make_float2(1.0009765625f, 1.0009765625f);
#pragma unroll 4
  for (long long i = 0; i < iters; i++) {
    const float r0 = rcp_approx(x0), r1 = rcp_approx(x1);
    const float2 rr = make_float2(r0, r1), rs = make_float2(r1, r0);
#pragma unroll
    for (int k = 0; k < NF2; k++) a[k] = ffma2(a[k], c, (k & 1) ? rs : rr);
#pragma unroll
    for (int k = 0; k < NF2; k++) a[(k + 1) % 6] = ffma2(a[(k + 1) % 6], c, (k & 1) ? rr : rs);
    x0 += 0.001f; x1 += 0.001f;
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 6; k++) s += a[k].x + a[k].y;
  if (s == 12345.678f) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_micro_ffma2(long long iters, float* sink) {
  float2 a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = make_float2(threadIdx.x * 1e-3f + k, 0.5f * k);
  const float2 m = make_float2(0.999f, 0.998f), c = make_float2(1e-3f, 2e-3f);
#pragma unroll 16
  for (long long i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = ffma2(a[k], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k].x + a[k].y;
  if (s == 12345.678f) sink[0] = s;
}

extern "C" int tsff_microbench(int kind, int64_t iters, double* ops, float* sink, void* stream) {
  int dev = 0, sms = 0;
  TSFF_CUDA_OK(cudaGetDevice(&dev));
  TSFF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kind == 0) {
    k_micro_ffma<<<blocks, threads, 0, st>>>(iters, sink);
    if (ops) *ops = (double)blocks * threads * (double)iters * 8.0;
  } else if (kind == 1) {
    k_micro_lg2<<<blocks, threads, 0, st>>>(iters, sink);
    if (ops) *ops = (double)blocks * threads * (double)iters * 4.0;
  } else if (kind == 3) {
    k_micro_ffma2<<<blocks, threads, 0, st>>>(iters, sink);
    if (ops) *ops = (double)blocks * threads * (double)iters * 16.0;  // FMAs
  } else if (kind >= 200 && kind <= 206) {
    switch (kind - 200) {  // NF2 FFMA2 per MUFU
      case 1: k_micro_mix2<1><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 2: k_micro_mix2<2><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 3: k_micro_mix2<3><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 4: k_micro_mix2<4><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 5: k_micro_mix2<5><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 6: k_micro_mix2<6><<<blocks, threads, 0, st>>>(iters, sink); break;
      default: set_error("unsupported mix2"); return TSFF_E_INVALID;
    }
    if (ops) *ops = (double)blocks * threads * (double)iters * 2.0;
  } else if (kind >= 100 && kind <= 112) {
    // kind = 100 + NF: per step 2 MUFU.RCP + 2*NF FFMA + 2 FADD; *ops counts MUFU ops
    switch (kind - 100) {
      case 0: k_micro_mix<0><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 2: k_micro_mix<2><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 4: k_micro_mix<4><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 6: k_micro_mix<6><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 8: k_micro_mix<8><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 10: k_micro_mix<10><<<blocks, threads, 0, st>>>(iters, sink); break;
      case 12: k_micro_mix<12><<<blocks, threads, 0, st>>>(iters, sink); break;
      default: set_error("unsupported mix"); return TSFF_E_INVALID;
    }
    if (ops) *ops = (double)blocks * threads * (double)iters * 2.0;
  } else {
    set_error("unknown microbench kind %d", kind);
    return TSFF_E_INVALID;
  }
  TSFF_LAUNCH_OK("microbench");
  return TSFF_OK;
}
