// tsff_ats.cu -- angularly resolved Thomson scattering (ARTS) instrument stage, forward + adjoint:
//   irf.add_ATS_IRF                 tsadar/core/physics/irf.py:5-47        (norm == 0 decks)
//   reduce_ATS_to_resunit + noise   tsadar/core/thomson_diagnostic.py:78-107, 139
//
//   X [NA][W]  (= weights @ formfactor^T, generate_spectra.py:194-195)
//   Y1 = conv_same(X, g_ang) along the angle axis   (irf.py:34)       taps g_ang[k] = Gaussian(angAxis[k] - mid)
//   Y2 = conv_same(Y1, g_lam) along wavelength      (irf.py:36)       taps g_lam[k] = Gaussian(lamAxis[k] - mid)
//   Z  = rowmax(X) / rowmax(Y2) * Y2                (irf.py:39)
//   R  = box average over lam_step columns, then ang_step rows  (thomson_diagnostic.py:93-97; short last groups as the
//        reference's slices produce them), rows [row0, row1) kept (:101)
//   T  = e_amps * R / rowmax(R) * (amp1 where lam_unit < lam else amp2) + noise     (:102-106, 139)
//
// The reference convolves with full-length tap vectors; beyond ~10 sigma the taps are below 1e-22 of the peak, so the
// host passes the tap arrays together with their support [t0, t1] and the kernels visit only that range.
// All FP64.  Everything here is O(NA*W*support) multiply-adds on a 16 MB array that lives in L2: HBM/L2-bound, tiny
// next to the form factor (493 k poles per image).
#include "tsff_common.cuh"

using namespace tsff;

namespace {
constexpr int kThreads = 256;

struct AtsGeom {
  int NA, W, na, nl, lam_step, ang_step, row0, nrows;   // na = ceil(NA/ang_step) angle units, nl = ceil(W/lam_step)
  int ta0, ta1, tl0, tl1;                               // tap supports (inclusive)
  double lam_min, dlam;
};

struct AtsLayout {   // saved: Y2 [NA][W], R [nrows][nl], stats; ws: Y1 / cotangent scratch
  size_t s_Y2, s_R, s_stats, saved_bytes, w_Y1, w_Zbar, w_Rbar, w_xmax, ws_bytes;
};
struct RowStat { double mx, my; int jx, jy; };     // per angle row: max/argmax of X and of Y2
struct UnitStat { double mr; int jr, pad; };       // per kept unit row: max/argmax of R

AtsGeom ats_geom(const tsff_ats_cfg* c) {
  AtsGeom g;
  g.NA = c->NA; g.W = c->W; g.lam_step = c->lam_step; g.ang_step = c->ang_step;
  g.nl = (c->W + c->lam_step - 1) / c->lam_step;
  g.na = (c->NA + c->ang_step - 1) / c->ang_step;
  g.row0 = c->row_start; g.nrows = c->row_end - c->row_start;
  g.ta0 = c->ang_t0; g.ta1 = c->ang_t1; g.tl0 = c->lam_t0; g.tl1 = c->lam_t1;
  g.lam_min = c->lam_min; g.dlam = (c->lam_max - c->lam_min) / (double)(c->W - 1);
  return g;
}
AtsLayout ats_layout(const AtsGeom& g) {
  AtsLayout L;
  size_t o = 0;
  L.s_Y2 = o; o += align_up((size_t)g.NA * g.W * 8);
  L.s_R = o; o += align_up((size_t)g.nrows * g.nl * 8);
  L.s_stats = o; o += align_up((size_t)g.NA * sizeof(RowStat) + (size_t)g.nrows * sizeof(UnitStat));
  L.saved_bytes = o;
  o = 0;
  L.w_Y1 = o; o += align_up((size_t)g.NA * g.W * 8);
  L.w_Zbar = o; o += align_up((size_t)g.NA * g.W * 8);
  L.w_Rbar = o; o += align_up((size_t)g.nrows * g.nl * 8);
  L.w_xmax = o; o += align_up((size_t)g.NA * 8);
  L.ws_bytes = o;
  return L;
}

// 'same' convolution with equal-length taps (jnp.convolve(x, v, "same")): y[n] = sum_t v[t] x[n + c - t], c = (N-1)/2.
// ADJ: the transpose  xbar[m] = sum_t v[t] ybar[m - c + t].   Axis 0: the conv index is the row (stride W), axis 1: column.
// Register-tiled: a thread computes kR = 8 consecutive outputs along the convolution axis; their eight-sample window lives
// in registers and rotates by one per tap (fully unrolled: no moves), so one load of a sample and one (broadcast) load of a
// tap feed eight DFMAs.  Every output is still summed in tap order t0 .. t1, out-of-range samples contribute 0.
// Axis 0: adjacent threads take adjacent columns (coalesced row loads); axis 1: adjacent threads take adjacent 8-column
// blocks of one row (the lines stay in L1 across the next taps).
constexpr int kR = 8;
template <bool AXIS0, bool ADJ>
__global__ void __launch_bounds__(kThreads) k_ats_conv(const double* __restrict__ x, double* __restrict__ y, int NA, int W,
                                                      const double* __restrict__ taps, int t0, int t1) {
  const int N = AXIS0 ? NA : W;                      // length of the convolution axis
  const int nblk = (N + kR - 1) / kR;                // output blocks along it
  const int other = AXIS0 ? W : NA;                  // the other axis
  const long long idx = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= (long long)nblk * other) return;
  // axis 0: idx = blk * W + j (threads of a warp: consecutive j);  axis 1: idx = i * nblk + blk
  const int blk = AXIS0 ? (int)(idx / other) : (int)(idx % nblk);
  const int o = AXIS0 ? (int)(idx % other) : (int)(idx / nblk);
  const int n0 = blk * kR;
  const int c = (N - 1) / 2;
  const long long stride = AXIS0 ? W : 1, base = AXIS0 ? o : (long long)o * W;
  auto ld = [&](int m) { return (m >= 0 && m < N) ? x[base + (long long)m * stride] : 0.0; };
  // fwd: sample index of output n, tap t: n + c - t (decreasing in t);  adj: n - c + t (increasing)
  const int i0 = ADJ ? n0 - c + t0 : n0 + c - t0;
  double acc[kR], v[kR];
#pragma unroll
  for (int r = 0; r < kR; r++) { acc[r] = 0.0; v[r] = ld(i0 + r); }
  const int nt = t1 - t0 + 1;
  for (int ib = 0; ib < nt; ib += kR) {
#pragma unroll
    for (int u = 0; u < kR; u++) {
      const double gv = (ib + u < nt) ? taps[t0 + ib + u] : 0.0;
      if (ADJ) {
#pragma unroll
        for (int r = 0; r < kR; r++) acc[r] = fma(gv, v[(r + u) & (kR - 1)], acc[r]);
        v[u] = ld(i0 + ib + u + kR);                                   // enters as r = 7 of the next tap
      } else {
#pragma unroll
        for (int r = 0; r < kR; r++) acc[r] = fma(gv, v[(r - u) & (kR - 1)], acc[r]);
        v[(kR - 1 - u) & (kR - 1)] = ld(i0 - ib - u - 1);              // enters as r = 0 of the next tap
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kR; r++)
    if (n0 + r < N) y[base + (long long)(n0 + r) * stride] = acc[r];
}

// per row: max / first argmax of two arrays (jnp.amax semantics)
__device__ __forceinline__ void block_argmax(double& v, int& ix, double* sv, int* si) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double v2 = __shfl_down_sync(0xffffffffu, v, o);
    const int i2 = __shfl_down_sync(0xffffffffu, ix, o);
    if (v2 > v || (v2 == v && i2 < ix)) { v = v2; ix = i2; }
  }
  if (lane == 0) { sv[wid] = v; si[wid] = ix; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; w++)
      if (sv[w] > v || (sv[w] == v && si[w] < ix)) { v = sv[w]; ix = si[w]; }
    sv[0] = v; si[0] = ix;
  }
  __syncthreads();
  v = sv[0]; ix = si[0];
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) k_ats_rowstat(const double* __restrict__ X, const double* __restrict__ Y2, int W,
                                                         RowStat* st) {
  __shared__ double sv[kThreads / 32];
  __shared__ int si[kThreads / 32];
  const int i = blockIdx.x;
  double vx = -INFINITY, vy = -INFINITY;
  int jx = 0x7fffffff, jy = 0x7fffffff;
  for (int j = threadIdx.x; j < W; j += kThreads) {
    const double a = X[(long long)i * W + j], b = Y2[(long long)i * W + j];
    if (a > vx) { vx = a; jx = j; }
    if (b > vy) { vy = b; jy = j; }
  }
  block_argmax(vx, jx, sv, si);
  block_argmax(vy, jy, sv, si);
  if (threadIdx.x == 0) { st[i].mx = vx; st[i].my = vy; st[i].jx = jx; st[i].jy = jy; }
}

// R[u][q] = mean over the angle rows of unit (row0 + u) and the wavelength columns of unit q of  (mx/my)_row * Y2
__global__ void __launch_bounds__(kThreads) k_ats_reduce(const AtsGeom g, const double* __restrict__ Y2, const RowStat* st,
                                                        double* __restrict__ R) {
  const long long idx = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= (long long)g.nrows * g.nl) return;
  const int u = (int)(idx / g.nl), q = (int)(idx % g.nl);
  const int i0 = (g.row0 + u) * g.ang_step, i1 = min(i0 + g.ang_step, g.NA);
  const int j0 = q * g.lam_step, j1 = min(j0 + g.lam_step, g.W);
  double acc = 0.0;
  for (int i = i0; i < i1; i++) {
    double s = 0.0;
    for (int j = j0; j < j1; j++) s += Y2[(long long)i * g.W + j];
    acc += st[i].mx / st[i].my * (s / (double)(j1 - j0));
  }
  R[idx] = acc / (double)(i1 - i0);
}

struct AtsCall {
  const double* params;  // one row: uses lam, amp1, amp2
  const double* e_amps;  // [nrows]
  const double* noise;   // [nrows][nl] or null
  double* thry;          // [nrows][nl]
  const double* thry_bar;
  double* Rbar;          // [nrows][nl]
  double* amp_bar;       // [2]  (amp1, amp2) -- accumulated, zero it
};

// wavelength of unit q: mean of the lam-axis samples it averages (thomson_diagnostic.py:98-100)
__device__ __forceinline__ double unit_lam(const AtsGeom& g, int q) {
  const int j0 = q * g.lam_step, j1 = min(j0 + g.lam_step, g.W);
  return g.lam_min + g.dlam * 0.5 * (double)(j0 + j1 - 1);
}

__global__ void __launch_bounds__(kThreads) k_ats_finish(const AtsGeom g, const AtsCall c, const double* __restrict__ R,
                                                        UnitStat* us) {
  __shared__ double sv[kThreads / 32];
  __shared__ int si[kThreads / 32];
  const int u = blockIdx.x;
  double v = -INFINITY;
  int jr = 0x7fffffff;
  for (int q = threadIdx.x; q < g.nl; q += kThreads) {
    const double a = R[(long long)u * g.nl + q];
    if (a > v) { v = a; jr = q; }
  }
  block_argmax(v, jr, sv, si);
  if (threadIdx.x == 0) { us[u].mr = v; us[u].jr = jr; }
  const double lam = c.params[TSFF_P_LAM], a1 = c.params[TSFF_P_AMP1], a2 = c.params[TSFF_P_AMP2];
  const double sc = c.e_amps[u] / v;
  for (int q = threadIdx.x; q < g.nl; q += kThreads) {
    const double t = sc * R[(long long)u * g.nl + q] * (unit_lam(g, q) < lam ? a1 : a2);
    c.thry[(long long)u * g.nl + q] = t + (c.noise ? c.noise[(long long)u * g.nl + q] : 0.0);
  }
}

// ---- adjoint ------------------------------------------------------------------------------------------------------
// T = e R / m * amp(q)  ->  Rbar, amp cotangents; the max term lands on R[u][jr]
__global__ void __launch_bounds__(kThreads) k_ats_finish_bwd(const AtsGeom g, const AtsCall c, const double* __restrict__ R,
                                                            const UnitStat* us) {
  __shared__ double sred[3 * (kThreads / 32)];
  const int u = blockIdx.x;
  const double lam = c.params[TSFF_P_LAM], a1 = c.params[TSFF_P_AMP1], a2 = c.params[TSFF_P_AMP2];
  const double m = us[u].mr, e = c.e_amps[u];
  double s_m = 0.0, s_a1 = 0.0, s_a2 = 0.0;
  for (int q = threadIdx.x; q < g.nl; q += kThreads) {
    const double tb = c.thry_bar[(long long)u * g.nl + q], r = R[(long long)u * g.nl + q];
    const bool blue = unit_lam(g, q) < lam;
    const double amp = blue ? a1 : a2;
    c.Rbar[(long long)u * g.nl + q] = tb * e * amp / m;
    s_m += -tb * e * amp * r / (m * m);
    if (blue) s_a1 += tb * e * r / m; else s_a2 += tb * e * r / m;
  }
  double vals[3] = {s_m, s_a1, s_a2};
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < 3; k++) {
    const double s = warp_sum(vals[k]);
    if (lane == 0) sred[k * (kThreads / 32) + wid] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < 3; k++)
      for (int w = 0; w < kThreads / 32; w++) t[k] += sred[k * (kThreads / 32) + w];
    c.Rbar[(long long)u * g.nl + us[u].jr] += t[0];
    atomicAdd(&c.amp_bar[0], t[1]);
    atomicAdd(&c.amp_bar[1], t[2]);
  }
}

// Zbar[i][j] = Rbar[unit(i)][unit(j)] / (group sizes);  Y2bar = r_i Zbar;  rbar_i = sum_j Zbar Y2  (second pass below)
__global__ void __launch_bounds__(kThreads) k_ats_reduce_bwd(const AtsGeom g, const double* __restrict__ Rbar,
                                                            const RowStat* st, double* __restrict__ Zbar) {
  const long long idx = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= (long long)g.NA * g.W) return;
  const int i = (int)(idx / g.W), j = (int)(idx % g.W);
  const int U = i / g.ang_step, q = j / g.lam_step;
  double v = 0.0;
  if (U >= g.row0 && U < g.row0 + g.nrows) {
    const int i0 = U * g.ang_step, i1 = min(i0 + g.ang_step, g.NA), j0 = q * g.lam_step, j1 = min(j0 + g.lam_step, g.W);
    v = Rbar[(long long)(U - g.row0) * g.nl + q] / ((double)(i1 - i0) * (double)(j1 - j0));
  }
  Zbar[idx] = v;
}

// per row: rbar = sum_j Zbar Y2; Y2bar = r Zbar (in place) + the max(Y2) term; returns the max(X) cotangent in xmax[i]
__global__ void __launch_bounds__(kThreads) k_ats_rescale_bwd(const double* __restrict__ Y2, double* __restrict__ Zbar, int W,
                                                             const RowStat* st, double* __restrict__ xmaxbar) {
  __shared__ double sred[kThreads / 32];
  const int i = blockIdx.x;
  const double r = st[i].mx / st[i].my;
  double s = 0.0;
  for (int j = threadIdx.x; j < W; j += kThreads) {
    const double zb = Zbar[(long long)i * W + j];
    s += zb * Y2[(long long)i * W + j];
    Zbar[(long long)i * W + j] = r * zb;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double rbar = 0.0;
    for (int w = 0; w < kThreads / 32; w++) rbar += sred[w];
    Zbar[(long long)i * W + st[i].jy] += -rbar * st[i].mx / (st[i].my * st[i].my);
    xmaxbar[i] = rbar / st[i].my;
  }
}

__global__ void __launch_bounds__(kThreads) k_ats_add_xmax(double* __restrict__ Xbar, int W, const RowStat* st,
                                                          const double* __restrict__ xmaxbar, int NA) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i < NA) Xbar[(long long)i * W + st[i].jx] += xmaxbar[i];
}

int check_cfg(const tsff_ats_cfg* c) {
  if (!c || c->NA < 2 || c->W < 2 || c->lam_step < 1 || c->ang_step < 1 || !c->taps_ang || !c->taps_lam) {
    set_error("bad ATS configuration"); return TSFF_E_INVALID;
  }
  if (c->norm != 0) { set_error("PhysParams.norm > 0 is not supported (all reference decks use 0)"); return TSFF_E_INVALID; }
  const int na = (c->NA + c->ang_step - 1) / c->ang_step;
  if (c->row_start < 0 || c->row_end > na || c->row_end <= c->row_start) { set_error("bad lineout row range"); return TSFF_E_INVALID; }
  if (c->ang_t0 < 0 || c->ang_t1 >= c->NA || c->lam_t0 < 0 || c->lam_t1 >= c->W || c->ang_t0 > c->ang_t1 || c->lam_t0 > c->lam_t1) {
    set_error("bad tap support"); return TSFF_E_INVALID;
  }
  return TSFF_OK;
}
}  // namespace

extern "C" size_t tsff_ats_saved_bytes(const tsff_ats_cfg* c) { return check_cfg(c) ? 0 : ats_layout(ats_geom(c)).saved_bytes; }
extern "C" size_t tsff_ats_workspace_bytes(const tsff_ats_cfg* c) { return check_cfg(c) ? 0 : ats_layout(ats_geom(c)).ws_bytes; }

extern "C" int tsff_ats_fwd(const tsff_ats_cfg* cfg, const double* modl, const double* params, const double* e_amps,
                            const double* noise, double* thry, void* saved, void* ws, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!modl || !params || !e_amps || !thry || !saved || !ws) { set_error("null argument"); return TSFF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AtsGeom g = ats_geom(cfg);
  const AtsLayout L = ats_layout(g);
  char* sv = static_cast<char*>(saved);
  char* w = static_cast<char*>(ws);
  double* Y1 = (double*)(w + L.w_Y1);
  double* Y2 = (double*)(sv + L.s_Y2);
  double* R = (double*)(sv + L.s_R);
  RowStat* rs = (RowStat*)(sv + L.s_stats);
  UnitStat* us = (UnitStat*)(sv + L.s_stats + (size_t)g.NA * sizeof(RowStat));
  const unsigned nb = (unsigned)(((long long)g.NA * g.W + kThreads - 1) / kThreads);
  const unsigned nb0 = (unsigned)(((long long)((g.NA + kR - 1) / kR) * g.W + kThreads - 1) / kThreads);   // axis-0 conv: 8 rows per thread
  const unsigned nb1 = (unsigned)(((long long)((g.W + kR - 1) / kR) * g.NA + kThreads - 1) / kThreads);   // axis-1 conv: 8 columns per thread
  k_ats_conv<true, false><<<nb0, kThreads, 0, st>>>(modl, Y1, g.NA, g.W, cfg->taps_ang, g.ta0, g.ta1);
  k_ats_conv<false, false><<<nb1, kThreads, 0, st>>>(Y1, Y2, g.NA, g.W, cfg->taps_lam, g.tl0, g.tl1);
  k_ats_rowstat<<<(unsigned)g.NA, kThreads, 0, st>>>(modl, Y2, g.W, rs);
  k_ats_reduce<<<(unsigned)(((long long)g.nrows * g.nl + kThreads - 1) / kThreads), kThreads, 0, st>>>(g, Y2, rs, R);
  AtsCall c;
  memset(&c, 0, sizeof(c));
  c.params = params; c.e_amps = e_amps; c.noise = noise; c.thry = thry;
  k_ats_finish<<<(unsigned)g.nrows, kThreads, 0, st>>>(g, c, R, us);
  TSFF_LAUNCH_OK("tsff_ats_fwd");
  return TSFF_OK;
}

extern "C" int tsff_ats_bwd(const tsff_ats_cfg* cfg, const double* params, const double* e_amps, const void* saved,
                            const double* thry_bar, double* modl_bar, double* amp_bar, void* ws, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!params || !e_amps || !saved || !thry_bar || !modl_bar || !amp_bar || !ws) { set_error("null argument"); return TSFF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AtsGeom g = ats_geom(cfg);
  const AtsLayout L = ats_layout(g);
  const char* sv = static_cast<const char*>(saved);
  char* w = static_cast<char*>(ws);
  const double* Y2 = (const double*)(sv + L.s_Y2);
  const double* R = (const double*)(sv + L.s_R);
  const RowStat* rs = (const RowStat*)(sv + L.s_stats);
  const UnitStat* us = (const UnitStat*)(sv + L.s_stats + (size_t)g.NA * sizeof(RowStat));
  double* Zbar = (double*)(w + L.w_Zbar);
  double* Y1bar = (double*)(w + L.w_Y1);
  double* Rbar = (double*)(w + L.w_Rbar);
  TSFF_CUDA_OK(cudaMemsetAsync(amp_bar, 0, 2 * sizeof(double), st));
  AtsCall c;
  memset(&c, 0, sizeof(c));
  c.params = params; c.e_amps = e_amps; c.thry_bar = thry_bar; c.Rbar = Rbar; c.amp_bar = amp_bar;
  k_ats_finish_bwd<<<(unsigned)g.nrows, kThreads, 0, st>>>(g, c, R, us);
  const unsigned nb = (unsigned)(((long long)g.NA * g.W + kThreads - 1) / kThreads);
  k_ats_reduce_bwd<<<nb, kThreads, 0, st>>>(g, Rbar, rs, Zbar);
  double* xmaxbar = (double*)(w + L.w_xmax);
  k_ats_rescale_bwd<<<(unsigned)g.NA, kThreads, 0, st>>>(Y2, Zbar, g.W, rs, xmaxbar);
  const unsigned nb0 = (unsigned)(((long long)((g.NA + kR - 1) / kR) * g.W + kThreads - 1) / kThreads);
  const unsigned nb1 = (unsigned)(((long long)((g.W + kR - 1) / kR) * g.NA + kThreads - 1) / kThreads);
  k_ats_conv<false, true><<<nb1, kThreads, 0, st>>>(Zbar, Y1bar, g.NA, g.W, cfg->taps_lam, g.tl0, g.tl1);
  k_ats_conv<true, true><<<nb0, kThreads, 0, st>>>(Y1bar, modl_bar, g.NA, g.W, cfg->taps_ang, g.ta0, g.ta1);
  k_ats_add_xmax<<<(unsigned)((g.NA + kThreads - 1) / kThreads), kThreads, 0, st>>>(modl_bar, g.W, rs, xmaxbar, g.NA);
  TSFF_LAUNCH_OK("tsff_ats_bwd");
  return TSFF_OK;
}
