// tsff_params.cu -- the stage directly upstream of the form factor on every fit step, as kernels:
//
//   tsff_params_fwd / _bwd   ThomsonParams.__call__ (tsadar/core/modules/ts_params.py:583-603): normalised leaves -> physical
//                            parameter block [B][NP] (sigmoid for active leaves, affine de-normalisation :93-104, 202-218,
//                            308-326, 459-495; ion-fraction renormalisation and tied Ti :543-563) fused with the DLM1V f(v)
//                            producer (distribution_functions/base.py:277-294: lerp of the projected super-Gaussian table in
//                            m, normalisation sum f dv = 1) -> fe [B][V]; and the reverse of both.
//   tsff_adam_step           optax.adam on the active leaves of every lineout in one launch (inverse/loops.py:59-95, 225-250;
//                            optax defaults b1 = 0.9, b2 = 0.999, eps = 1e-8, eps_root = 0), per-lineout step counters on
//                            the device so that a captured CUDA graph replays unchanged.
//
// Leaf columns (NL = 10 + 3 I + 1): Te ne | lam Va ud ne_gradient Te_gradient amp1 amp2 amp3 | per ion: Z Ti fract | m.
// A leaf is either ACTIVE (trainable: its normalised value lives in x_active[b][slot], physical = sigmoid(x) scale + shift)
// or static (value in x_static[b][k], physical = x scale + shift).  All FP64 except the f table output (fe_dtype).
// Bound: HBM writes of fe (16 KB per lineout at V = 4096); the table (31 x V doubles) is L2-resident.
#include "tsff_common.cuh"

using namespace tsff;

namespace {
constexpr int kThreads = 256;

struct PCfg {   // tsff_params_cfg by value (kernel argument)
  int I, V, nm, fe_f32, NL, NLA, NP;
  double dv, m_offset, m0, dm;
  int slot[TSFF_MAX_LEAVES];
  double scale[TSFF_MAX_LEAVES], shift[TSFF_MAX_LEAVES];
  double ionA[TSFF_MAX_IONS];
  int ti_same[TSFF_MAX_IONS];
  const double* tab;   // [nm][V]
};

__device__ __forceinline__ double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

// physical value of leaf k of lineout b and d physical / d x (0 for a static leaf)
__device__ __forceinline__ double leaf_phys(const PCfg& c, const double* xa, const double* xs, long long b, int k, double& dphys) {
  const int s = c.slot[k];
  if (s >= 0) {
    const double sg = sigmoid(xa[b * c.NLA + s]);
    dphys = sg * (1.0 - sg) * c.scale[k];
    return sg * c.scale[k] + c.shift[k];
  }
  dphys = 0.0;
  return xs[b * c.NL + k] * c.scale[k] + c.shift[k];
}

// block-wide sums of up to 4 doubles (result valid in every thread)
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* sred) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; k++) {
    const double s = warp_sum(v[k]);
    if (lane == 0) sred[k * (kThreads / 32) + wid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) s += sred[k * (kThreads / 32) + w];
    v[k] = s;
  }
  __syncthreads();
}

// m -> (table row i0, lerp weight w): jnp.interp(m, m_ax, table) with edge clamping (base.py:292)
__device__ __forceinline__ void m_cell(const PCfg& c, double m, int& i0, double& w, bool& inside) {
  double u = (m - c.m0) / c.dm;
  inside = u > 0.0 && u < (double)(c.nm - 1);
  if (!(u > 0.0)) { i0 = 0; w = 0.0; return; }
  if (u >= (double)(c.nm - 1)) { i0 = c.nm - 2; w = 1.0; return; }
  i0 = (int)u;
  if (i0 > c.nm - 2) i0 = c.nm - 2;
  w = u - (double)i0;
}

// one CTA per lineout: thread 0 writes the parameter block, all threads build fe
template <typename T>
__global__ void __launch_bounds__(kThreads) k_params_fwd(const PCfg c, const double* __restrict__ xa, const double* __restrict__ xs,
                                                         double* __restrict__ params, T* __restrict__ fe) {
  __shared__ double sred[4 * (kThreads / 32)];
  const long long b = blockIdx.x;
  if (threadIdx.x == 0) {
    double* p = params + b * c.NP;
    double d;
    static_assert(TSFF_P_TE == 0 && TSFF_P_NE == 1 && TSFF_P_LAM == 2 && TSFF_P_VA == 3 && TSFF_P_UD == 4 && TSFF_P_NE_GRAD == 5 &&
                      TSFF_P_TE_GRAD == 6 && TSFF_P_AMP1 == 7 && TSFF_P_AMP2 == 8 && TSFF_P_AMP3 == 9 && TSFF_P_ION0 == 10,
                  "the first ten leaf columns are the first ten columns of the parameter block");
    for (int k = 0; k < 10; k++) p[k] = leaf_phys(c, xa, xs, b, k, d);
    double fsum = 0.0;
    for (int i = 0; i < c.I; i++) fsum += leaf_phys(c, xa, xs, b, 10 + 3 * i + 2, d);
    for (int i = 0; i < c.I; i++) {
      double* q = p + TSFF_P_ION0 + TSFF_ION_STRIDE * i;
      q[TSFF_ION_A] = c.ionA[i];
      q[TSFF_ION_Z] = leaf_phys(c, xa, xs, b, 10 + 3 * i, d);
      q[TSFF_ION_TI] = leaf_phys(c, xa, xs, b, 10 + 3 * ((i > 0 && c.ti_same[i]) ? 0 : i) + 1, d);   // ts_params.py:555-557
      q[TSFF_ION_FRACT] = leaf_phys(c, xa, xs, b, 10 + 3 * i + 2, d) / fsum;                            // :559-561
    }
  }
  if (!fe || c.V <= 0) return;
  int i0 = 0;
  double w = 0.0;
  bool inside = false;
  if (c.nm > 1) {
    double d;
    m_cell(c, leaf_phys(c, xa, xs, b, c.NL - 1, d) + c.m_offset, i0, w, inside);
  }
  const double* t0 = c.tab + (long long)i0 * c.V;
  const double* t1 = c.nm > 1 ? t0 + c.V : t0;
  double s[1] = {0.0};
  for (int i = threadIdx.x; i < c.V; i += kThreads) s[0] += t0[i] * (1.0 - w) + t1[i] * w;
  block_sum<1>(s, sred);
  const double inv = 1.0 / (s[0] * c.dv);                                                        // base.py:294
  T* out = fe + b * c.V;
  for (int i = threadIdx.x; i < c.V; i += kThreads) out[i] = (T)((t0[i] * (1.0 - w) + t1[i] * w) * inv);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_params_bwd(const PCfg c, const double* __restrict__ xa, const double* __restrict__ xs,
                                                         const double* __restrict__ params_bar, const T* __restrict__ fe_bar,
                                                         double* __restrict__ xa_bar) {
  __shared__ double sred[4 * (kThreads / 32)];
  const long long b = blockIdx.x;
  // ---- m: fe_i = u_i / (S dv), u = (1 - w) T0 + w T1, S = sum u;  w_bar = sum_k u_bar_k (T1_k - T0_k)
  double m_bar = 0.0;
  const int ms = c.nm > 1 ? c.slot[c.NL - 1] : -1;
  if (fe_bar && c.V > 0 && ms >= 0) {
    int i0;
    double w, d;
    bool inside;
    m_cell(c, leaf_phys(c, xa, xs, b, c.NL - 1, d) + c.m_offset, i0, w, inside);
    const double* t0 = c.tab + (long long)i0 * c.V;
    const double* t1 = t0 + c.V;
    const T* fb = fe_bar + b * c.V;
    double v[4] = {0.0, 0.0, 0.0, 0.0};   // S, sum fb u, sum fb d, sum d
    for (int i = threadIdx.x; i < c.V; i += kThreads) {
      const double a0 = t0[i], dd = t1[i] - a0, u = a0 + w * dd, g = (double)fb[i];
      v[0] += u; v[1] += g * u; v[2] += g * dd; v[3] += dd;
    }
    block_sum<4>(v, sred);
    const double iS = 1.0 / v[0];
    const double w_bar = (v[2] - v[1] * v[3] * iS) * iS / c.dv;
    m_bar = inside ? w_bar / c.dm : 0.0;                                                         // clamped outside the m axis
  }
  if (threadIdx.x != 0) return;
  double* out = xa_bar + b * c.NLA;
  for (int s = 0; s < c.NLA; s++) out[s] = 0.0;
  const double* pb = params_bar + b * c.NP;
  double d;
  for (int k = 0; k < 10; k++) {
    leaf_phys(c, xa, xs, b, k, d);
    if (c.slot[k] >= 0) out[c.slot[k]] += pb[k] * d;
  }
  double fsum = 0.0, fdot = 0.0;   // fract_n = r_n / S:  r_bar_n = (fbar_n - sum_k fbar_k r_k / S) / S
  for (int i = 0; i < c.I; i++) {
    const double r = leaf_phys(c, xa, xs, b, 10 + 3 * i + 2, d);
    fsum += r;
    fdot += pb[TSFF_P_ION0 + TSFF_ION_STRIDE * i + TSFF_ION_FRACT] * r;
  }
  for (int i = 0; i < c.I; i++) {
    const double* q = pb + TSFF_P_ION0 + TSFF_ION_STRIDE * i;
    int k = 10 + 3 * i;
    leaf_phys(c, xa, xs, b, k, d);
    if (c.slot[k] >= 0) out[c.slot[k]] += q[TSFF_ION_Z] * d;
    k = 10 + 3 * ((i > 0 && c.ti_same[i]) ? 0 : i) + 1;
    leaf_phys(c, xa, xs, b, k, d);
    if (c.slot[k] >= 0) out[c.slot[k]] += q[TSFF_ION_TI] * d;
    k = 10 + 3 * i + 2;
    leaf_phys(c, xa, xs, b, k, d);
    if (c.slot[k] >= 0) out[c.slot[k]] += (q[TSFF_ION_FRACT] - fdot / fsum) / fsum * d;
  }
  if (ms >= 0) {
    leaf_phys(c, xa, xs, b, c.NL - 1, d);
    out[ms] += m_bar * d;
  }
}

// optax.adam, one thread per (lineout, active leaf); a block owns whole lineouts, so the per-lineout step counter is read by
// every thread of its lineout before the barrier and advanced by one of them after it.
__global__ void __launch_bounds__(kThreads) k_adam(long long B, int n, int lpb, double* __restrict__ x, const double* __restrict__ g,
                                                   double* __restrict__ mu, double* __restrict__ nu, double* __restrict__ count, double lr,
                                                   double b1, double b2, double eps) {
  const int l = threadIdx.x / n, k = threadIdx.x % n;
  const long long b = (long long)blockIdx.x * lpb + l;
  const bool on = l < lpb && b < B;
  double t = 0.0;
  if (on) t = count[b] + 1.0;
  __syncthreads();
  if (!on) return;
  const long long i = b * n + k;
  const double gi = g[i];
  const double m = b1 * mu[i] + (1.0 - b1) * gi;
  const double v = b2 * nu[i] + (1.0 - b2) * gi * gi;
  mu[i] = m;
  nu[i] = v;
  const double mh = m / (1.0 - pow(b1, t)), vh = v / (1.0 - pow(b2, t));
  x[i] -= lr * mh / (sqrt(vh) + eps);
  if (k == 0) count[b] = t;
}

int to_pcfg(const tsff_params_cfg* c, PCfg& p) {
  if (!c || c->I < 1 || c->I > TSFF_MAX_IONS) { set_error("tsff_params: bad ion count"); return TSFF_E_INVALID; }
  const int NL = 10 + 3 * c->I + 1;
  if (NL > TSFF_MAX_LEAVES || c->NLA < 0 || c->NLA > NL) { set_error("tsff_params: bad leaf counts"); return TSFF_E_INVALID; }
  if (c->V > 0 && (!c->f_vx_m || c->nm < 1 || !(c->dv > 0.0) || (c->nm > 1 && !(c->dm > 0.0)))) { set_error("tsff_params: bad f-table description"); return TSFF_E_INVALID; }
  p.I = c->I; p.V = c->V; p.nm = c->nm; p.fe_f32 = c->fe_dtype == TSFF_F32; p.NL = NL; p.NLA = c->NLA;
  p.NP = TSFF_P_ION0 + TSFF_ION_STRIDE * c->I;
  p.dv = c->dv; p.m_offset = c->m_offset; p.m0 = c->m0; p.dm = c->dm; p.tab = c->f_vx_m;
  for (int k = 0; k < TSFF_MAX_LEAVES; k++) { p.slot[k] = -1; p.scale[k] = 1.0; p.shift[k] = 0.0; }
  for (int k = 0; k < NL; k++) {
    p.slot[k] = c->active_slot[k];
    if (p.slot[k] >= c->NLA) { set_error("tsff_params: active slot %d out of range", p.slot[k]); return TSFF_E_INVALID; }
    p.scale[k] = c->scale[k]; p.shift[k] = c->shift[k];
  }
  for (int i = 0; i < TSFF_MAX_IONS; i++) { p.ionA[i] = i < c->I ? c->ionA[i] : 0.0; p.ti_same[i] = i < c->I ? c->ti_same[i] : 0; }
  return TSFF_OK;
}
}  // namespace

extern "C" int tsff_params_fwd(const tsff_params_cfg* cfg, int64_t B, const double* x_active, const double* x_static, double* params,
                               void* fe, void* stream) {
  PCfg p;
  int rc = to_pcfg(cfg, p);
  if (rc) return rc;
  if (B == 0) return TSFF_OK;
  if (B < 0 || !x_static || !params || (p.NLA > 0 && !x_active)) { set_error("tsff_params_fwd: null argument"); return TSFF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p.fe_f32) k_params_fwd<float><<<(unsigned)B, kThreads, 0, st>>>(p, x_active, x_static, params, static_cast<float*>(fe));
  else k_params_fwd<double><<<(unsigned)B, kThreads, 0, st>>>(p, x_active, x_static, params, static_cast<double*>(fe));
  TSFF_LAUNCH_OK("k_params_fwd");
  return TSFF_OK;
}

extern "C" int tsff_params_bwd(const tsff_params_cfg* cfg, int64_t B, const double* x_active, const double* x_static,
                               const double* params_bar, const void* fe_bar, double* x_active_bar, void* stream) {
  PCfg p;
  int rc = to_pcfg(cfg, p);
  if (rc) return rc;
  if (B == 0 || p.NLA == 0) return TSFF_OK;
  if (B < 0 || !x_static || !x_active || !params_bar || !x_active_bar) { set_error("tsff_params_bwd: null argument"); return TSFF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p.fe_f32) k_params_bwd<float><<<(unsigned)B, kThreads, 0, st>>>(p, x_active, x_static, params_bar, static_cast<const float*>(fe_bar), x_active_bar);
  else k_params_bwd<double><<<(unsigned)B, kThreads, 0, st>>>(p, x_active, x_static, params_bar, static_cast<const double*>(fe_bar), x_active_bar);
  TSFF_LAUNCH_OK("k_params_bwd");
  return TSFF_OK;
}

extern "C" int tsff_adam_step(int64_t B, int32_t n_active, double* x, const double* grad, double* mu, double* nu, double* count,
                              double lr, double b1, double b2, double eps, void* stream) {
  if (B == 0 || n_active == 0) return TSFF_OK;
  if (B < 0 || n_active < 0 || n_active > kThreads || !x || !grad || !mu || !nu || !count) { set_error("tsff_adam_step: bad argument"); return TSFF_E_INVALID; }
  const int lpb = kThreads / n_active;
  k_adam<<<(unsigned)((B + lpb - 1) / lpb), kThreads, 0, static_cast<cudaStream_t>(stream)>>>((long long)B, n_active, lpb, x, grad, mu, nu, count, lr,
                                                                                           b1, b2, eps);
  TSFF_LAUNCH_OK("k_adam");
  return TSFF_OK;
}
