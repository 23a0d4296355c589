// tsff_irf.cu -- instrument response + CCD binning + amplitude scaling (irf.add_electron_IRF / add_ion_IRF,
// tsadar/core/physics/irf.py:50-132) and the masked loss (loss_function.py:190-267, 386-418), forward + adjoint.
//
// The reference convolves with a full-length Gaussian (jnp.convolve(model, g, "same"), irf.py:72,114).  The taps
// are negligible beyond a few sigma, so the kernels truncate them at cut_sigma (default 12 sigma: e^-72 relative,
// below the 1e-21 dynamic range of the golden vector).  All arithmetic FP64; HBM traffic is (W + nbins) doubles per
// spectrum, so these stages are bandwidth/latency-bound and tiny next to the form factor.
//
//   k_irf_conv       tiled 'same' convolution (shared-memory halo) + per-tile max of x and of conv(x)
//   k_irf_finish     per lineout: global maxima, rescale, bin to pixels, amplitude normalisation
//   k_irf_bwd_pre    per lineout: reverse of k_irf_finish -> cotangent of conv(x), amplitudes, max(x) term
//   k_irf_bwd_conv   tiled correlation with the same taps -> cotangent of the model spectrum
//   k_loss           fused weighted loss + its gradient wrt the theory spectrum
#include "tsff_common.cuh"

using namespace tsff;

namespace {

constexpr int kThreads = 256;
constexpr int kConvR = 8;                        // consecutive outputs per thread in the convolution kernels
constexpr int kConvThreads = 128;
constexpr int kTile = kConvThreads * kConvR;     // outputs per CTA in the convolution kernels

// Register-tiled sliding dot product: acc[r] += sum_i s[i0 + r + DIR * i] * sg[i], i = 0 .. ntp - 1 (ntp a multiple of 8;
// taps beyond the real ones are zero), every output summed in tap order.  The eight-sample window lives in registers and
// rotates by one per tap (fully unrolled: no moves): one shared-memory load of a sample and one broadcast load of a tap
// feed eight DFMAs, instead of two loads per DFMA.
// The halo is stored with one pad element per eight samples (physical index i + (i >> 3)): thread t reads sample 8 t + c, a
// 64-byte lane stride that would put the 32 lanes of a warp into two bank pairs (16-way conflicts: the walk was bound by the
// shared-memory pipe, 17 of the kernel's 20 us at a fit batch of two lineouts); with the pad the stride is 72 bytes, conflict-free.
__device__ __forceinline__ int hpad(int i) { return i + (i >> 3); }

template <int DIR>
__device__ __forceinline__ void conv_tile8(const double* __restrict__ sp, const double* __restrict__ sg, int ntp, int i0,
                                           double (&acc)[kConvR]) {
  auto ldh = [sp](int i) { return sp[hpad(i)]; };
  double v[kConvR];
#pragma unroll
  for (int r = 0; r < kConvR; r++) v[r] = ldh(i0 + r);
  for (int ib = 0; ib < ntp; ib += kConvR) {
#pragma unroll
    for (int u = 0; u < kConvR; u++) {
      const double gv = sg[ib + u];
      if (DIR > 0) {
#pragma unroll
        for (int r = 0; r < kConvR; r++) acc[r] = fma(v[(r + u) & (kConvR - 1)], gv, acc[r]);
        v[u] = ldh(i0 + ib + u + kConvR);                                 // enters as r = 7 of the next tap
      } else {
#pragma unroll
        for (int r = 0; r < kConvR; r++) acc[r] = fma(v[(r - u) & (kConvR - 1)], gv, acc[r]);
        v[(kConvR - 1 - u) & (kConvR - 1)] = ldh(i0 - ib - u - 1);        // enters as r = 0 of the next tap
      }
    }
  }
}

struct IrfGeom {
  int W, nbins, r, K;      // r = W / nbins samples per pixel; K = tap half-width in samples
  double dlam, half;       // sample spacing; tap centre offset: d = n - m - half  (half = 0.5 for even W, 0 for odd)
  double inv2s2, gnorm;    // Gaussian 1/(2 sigma^2), 1/(sigma sqrt(2 pi))
  int ntiles;
};

IrfGeom irf_geom(const tsff_irf_cfg* c) {
  IrfGeom g;
  g.W = c->W; g.nbins = c->nbins; g.r = c->W / c->nbins;
  g.dlam = (c->lam_max - c->lam_min) / (double)(c->W - 1);
  g.half = (c->W % 2 == 0) ? 0.5 : 0.0;
  const double cut = c->cut_sigma > 0 ? c->cut_sigma : 12.0;
  g.K = (int)ceil(cut * c->stddev / g.dlam) + 1;
  if (g.K > c->W) g.K = c->W;
  g.inv2s2 = 1.0 / (2.0 * c->stddev * c->stddev);
  g.gnorm = 1.0 / (c->stddev * sqrt(2.0 * kPi));
  g.ntiles = (c->W + kTile - 1) / kTile;
  return g;
}

struct IrfStats {  // per lineout, written by k_irf_finish, read by the backward kernels
  double mx, my, mb;
  int im, iy, qm, pad;
  double Mb, Mr;   // PhysParams.norm > 0: maxima of the rescaled convolution on the blue / red side of the probe wavelength
  int ib, ir;      // and their sample indices
};

struct IrfLayout {
  size_t w_yc, w_pmax, w_ycbar, bytes;
};
IrfLayout irf_layout(const IrfGeom& g, int64_t B) {
  IrfLayout L;
  size_t o = 0;
  L.w_yc = o; o += align_up((size_t)B * g.W * 8);
  L.w_pmax = o; o += align_up((size_t)B * g.ntiles * 4 * 8);
  L.w_ycbar = o; o += align_up((size_t)B * g.W * 8);
  L.bytes = o;
  return L;
}

// tap weight for output n, input m:  G((n - m - half) dlam)     (irf.py:110-114 with 'same' alignment)
__device__ __forceinline__ double tap(const IrfGeom& g, int n, int m) {
  const double d = ((double)(n - m) - g.half) * g.dlam;
  return g.gnorm * exp(-d * d * g.inv2s2);
}

// ---- forward ------------------------------------------------------------------------------------------------
// dynamic smem: 8 zeros | x halo [kTile + 2K] | 8 zeros | taps [ntp] (zero-padded to a multiple of 8)
__global__ void __launch_bounds__(kConvThreads) k_irf_conv(const IrfGeom g, const double* __restrict__ x, double* __restrict__ yc,
                                                          double* __restrict__ pmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_rv[2 * (kConvThreads / 32)];
  __shared__ int s_ri[2 * (kConvThreads / 32)];
  const int ntp = (2 * g.K + 1 + kConvR - 1) / kConvR * kConvR;
  double* s_x = reinterpret_cast<double*>(smem_raw) + 9;           // logical index -8 -> physical -9
  double* s_g = s_x + hpad(kTile + 2 * g.K + 8) + 1;
  const int tile = blockIdx.x % g.ntiles;
  const long long b = blockIdx.x / g.ntiles;
  const int n0 = tile * kTile;
  const double* xb = x + b * g.W;
  for (int i = threadIdx.x - 8; i < kTile + 2 * g.K + 8; i += kConvThreads) {
    const int m = n0 - g.K + i;
    s_x[hpad(i)] = (i >= 0 && i < kTile + 2 * g.K && m >= 0 && m < g.W) ? xb[m] : 0.0;
  }
  // taps for offsets o = n - m in [-K, K]: s_g[o + K]
  for (int i = threadIdx.x; i < ntp; i += kConvThreads) {
    const double d = ((double)(i - g.K) - g.half) * g.dlam;
    s_g[i] = i <= 2 * g.K ? g.gnorm * exp(-d * d * g.inv2s2) : 0.0;
  }
  __syncthreads();
  double vmax_y = -INFINITY, vmax_x = -INFINITY;
  int imax_y = 0, imax_x = 0;
  {
    // y[n] = sum_m x[m] G(n - m - half);  m = n - o, o in [-K, K];  s_x index of m: t - o + K = t + 2K - i, i = o + K
    const int t0 = threadIdx.x * kConvR;
    double acc[kConvR];
#pragma unroll
    for (int r = 0; r < kConvR; r++) acc[r] = 0.0;
    conv_tile8<-1>(s_x, s_g, ntp, t0 + 2 * g.K, acc);
#pragma unroll
    for (int r = 0; r < kConvR; r++) {
      const int n = n0 + t0 + r;
      if (n < g.W) {
        yc[b * g.W + n] = acc[r];
        if (acc[r] > vmax_y) { vmax_y = acc[r]; imax_y = n; }
        const double xv = s_x[hpad(t0 + r + g.K)];
        if (xv > vmax_x) { vmax_x = xv; imax_x = n; }
      }
    }
  }
  // block arg-max (first index wins on ties, like jnp.argmax)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double v = __shfl_down_sync(0xffffffffu, vmax_y, o); int i = __shfl_down_sync(0xffffffffu, imax_y, o);
    if (v > vmax_y || (v == vmax_y && i < imax_y)) { vmax_y = v; imax_y = i; }
    v = __shfl_down_sync(0xffffffffu, vmax_x, o); i = __shfl_down_sync(0xffffffffu, imax_x, o);
    if (v > vmax_x || (v == vmax_x && i < imax_x)) { vmax_x = v; imax_x = i; }
  }
  if (lane == 0) { s_rv[2 * wid] = vmax_y; s_ri[2 * wid] = imax_y; s_rv[2 * wid + 1] = vmax_x; s_ri[2 * wid + 1] = imax_x; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kConvThreads / 32; w++) {
      if (s_rv[2 * w] > vmax_y || (s_rv[2 * w] == vmax_y && s_ri[2 * w] < imax_y)) { vmax_y = s_rv[2 * w]; imax_y = s_ri[2 * w]; }
      if (s_rv[2 * w + 1] > vmax_x || (s_rv[2 * w + 1] == vmax_x && s_ri[2 * w + 1] < imax_x)) { vmax_x = s_rv[2 * w + 1]; imax_x = s_ri[2 * w + 1]; }
    }
    double* pm = pmax + (b * g.ntiles + tile) * 4;
    pm[0] = vmax_y; pm[1] = (double)imax_y; pm[2] = vmax_x; pm[3] = (double)imax_x;
  }
}

struct IrfCall {
  int kind, norm, NP;
  double lam_min;
  const double *params, *amps, *noise;
  double* thry;
  IrfStats* stats;
  // backward
  const double* thry_bar;
  double* ycbar;
  double* amp_bar;   // [B][3]
  double* xbar_max;  // [B] cotangent that lands on x[argmax x]
};

// dynamic smem: yb[nbins]
__global__ void __launch_bounds__(kThreads) k_irf_finish(const IrfGeom g, const IrfCall c, const double* __restrict__ yc,
                                                        const double* __restrict__ pmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_stat[4];
  __shared__ int s_idx[3];
  double* s_yb = reinterpret_cast<double*>(smem_raw);
  const long long b = blockIdx.x;
  if (threadIdx.x == 0) {
    double my = -INFINITY, mx = -INFINITY; int iy = 0, im = 0;
    for (int t = 0; t < g.ntiles; t++) {
      const double* pm = pmax + (b * g.ntiles + t) * 4;
      if (pm[0] > my) { my = pm[0]; iy = (int)pm[1]; }
      if (pm[2] > mx) { mx = pm[2]; im = (int)pm[3]; }
    }
    s_stat[0] = mx; s_stat[1] = my; s_idx[0] = im; s_idx[1] = iy;
  }
  __syncthreads();
  const double s = s_stat[0] / s_stat[1];  // irf.py:73,115  max(model)/max(conv)
  if (c.norm > 0) {
    // PhysParams.norm > 0 (irf.py:117-124; ion :74): the blue and the red side of the probe wavelength are normalised to their
    // own maxima at full resolution, THEN binned; no amps / max scaling afterwards.  (Boolean-mask indexing with a traced mask:
    // the reference can run this branch only un-jitted; no deck uses it.)
    __shared__ double s_mv[2 * (kThreads / 32)];
    __shared__ int s_mi[2 * (kThreads / 32)];
    __shared__ double s_M[2];
    __shared__ int s_I[2];
    const double* p = c.params + b * c.NP;
    if (c.kind == 0) {
      double vb = -INFINITY, vr = -INFINITY;
      int ib = 0, ir = 0;
      for (int n = threadIdx.x; n < g.W; n += kThreads) {
        const double lam = c.lam_min + (double)n * g.dlam, y = s * yc[b * g.W + n];
        if (lam < p[P_LAM]) { if (y > vb) { vb = y; ib = n; } }
        else if (lam > p[P_LAM]) { if (y > vr) { vr = y; ir = n; } }
      }
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double v = __shfl_down_sync(0xffffffffu, vb, o); int i = __shfl_down_sync(0xffffffffu, ib, o);
        if (v > vb || (v == vb && i < ib)) { vb = v; ib = i; }
        v = __shfl_down_sync(0xffffffffu, vr, o); i = __shfl_down_sync(0xffffffffu, ir, o);
        if (v > vr || (v == vr && i < ir)) { vr = v; ir = i; }
      }
      if (lane == 0) { s_mv[2 * wid] = vb; s_mi[2 * wid] = ib; s_mv[2 * wid + 1] = vr; s_mi[2 * wid + 1] = ir; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; w++) {
          if (s_mv[2 * w] > vb || (s_mv[2 * w] == vb && s_mi[2 * w] < ib)) { vb = s_mv[2 * w]; ib = s_mi[2 * w]; }
          if (s_mv[2 * w + 1] > vr || (s_mv[2 * w + 1] == vr && s_mi[2 * w + 1] < ir)) { vr = s_mv[2 * w + 1]; ir = s_mi[2 * w + 1]; }
        }
        s_M[0] = vb; s_M[1] = vr; s_I[0] = ib; s_I[1] = ir;
      }
      __syncthreads();
    }
    for (int q = threadIdx.x; q < g.nbins; q += kThreads) {
      double a = 0.0;
      for (int k = 0; k < g.r; k++) {
        const int n = q * g.r + k;
        const double y = s * yc[b * g.W + n];
        if (c.kind == 0) {
          const double lam = c.lam_min + (double)n * g.dlam;
          a += lam < p[P_LAM] ? p[P_AMP1] * (y / s_M[0]) : p[P_AMP2] * (y / s_M[1]);
        } else {
          a += y;
        }
      }
      double v = a / (double)g.r;
      if (c.noise) v += c.noise[b * g.nbins + q];
      c.thry[b * g.nbins + q] = v;
    }
    if (threadIdx.x == 0) {
      IrfStats st; st.mx = s_stat[0]; st.my = s_stat[1]; st.mb = 1.0; st.im = s_idx[0]; st.iy = s_idx[1]; st.qm = 0; st.pad = 0;
      st.Mb = c.kind == 0 ? s_M[0] : 1.0; st.Mr = c.kind == 0 ? s_M[1] : 1.0; st.ib = c.kind == 0 ? s_I[0] : 0; st.ir = c.kind == 0 ? s_I[1] : 0;
      c.stats[b] = st;
    }
    return;
  }
  for (int q = threadIdx.x; q < g.nbins; q += kThreads) {
    double a = 0.0;
    for (int k = 0; k < g.r; k++) a += yc[b * g.W + q * g.r + k];
    s_yb[q] = s * a / (double)g.r;             // irf.py:74,124  reshape(1024,-1).mean
  }
  __syncthreads();
  // arg-max over the bins, first index on ties (jnp.argmax); by the whole CTA: a serial walk of 1024 bins by one thread was half
  // of this kernel at the batch sizes fits run at
  __shared__ double s_wv[kThreads / 32];
  __shared__ int s_wi[kThreads / 32];
  {
    double v = -INFINITY; int ix = 0x7fffffff;
    for (int q = threadIdx.x; q < g.nbins; q += kThreads) if (s_yb[q] > v) { v = s_yb[q]; ix = q; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double v2 = __shfl_down_sync(0xffffffffu, v, o);
      const int i2 = __shfl_down_sync(0xffffffffu, ix, o);
      if (v2 > v || (v2 == v && i2 < ix)) { v = v2; ix = i2; }
    }
    if ((threadIdx.x & 31) == 0) { s_wv[threadIdx.x >> 5] = v; s_wi[threadIdx.x >> 5] = ix; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double mb = s_wv[0]; int qm = s_wi[0];
    for (int w = 1; w < kThreads / 32; w++) if (s_wv[w] > mb || (s_wv[w] == mb && s_wi[w] < qm)) { mb = s_wv[w]; qm = s_wi[w]; }
    if (qm == 0x7fffffff) qm = 0;   // every bin NaN or -inf: the serial walk's answer
    s_stat[2] = mb; s_idx[2] = qm;
    IrfStats st; st.mx = s_stat[0]; st.my = s_stat[1]; st.mb = mb; st.im = s_idx[0]; st.iy = s_idx[1]; st.qm = qm; st.pad = 0;
    st.Mb = st.Mr = 1.0; st.ib = st.ir = 0;
    c.stats[b] = st;
  }
  __syncthreads();
  const double mb = s_stat[2];
  const double* p = c.params + b * c.NP;
  const double amps = c.amps[b];
  for (int q = threadIdx.x; q < g.nbins; q += kThreads) {
    // binned wavelength axis: mean of r consecutive samples of the uniform axis (irf.py:76,126)
    const double lamq = c.lam_min + ((double)(q * g.r) + 0.5 * (double)(g.r - 1)) * g.dlam;
    double v;
    if (c.kind == 0) {  // electron: amps*y/max(y), then amp1 (lam < lamL) or amp2   irf.py:127-130
      const double a = (lamq < p[P_LAM]) ? p[P_AMP1] : p[P_AMP2];
      v = a * (amps * s_yb[q] / mb);
    } else {            // ion: amp3*amps*y/max(y)                                   irf.py:77
      v = p[P_AMP3] * amps * s_yb[q] / mb;
    }
    if (c.noise) v += c.noise[b * g.nbins + q];   // thomson_diagnostic.py:139-140
    c.thry[b * g.nbins + q] = v;
  }
}

// ---- backward -------------------------------------------------------------------------------------------------
// dynamic smem: ybbar[nbins]
__global__ void __launch_bounds__(kThreads) k_irf_bwd_pre(const IrfGeom g, const IrfCall c, const double* __restrict__ yc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sred[4 * (kThreads / 32)];
  __shared__ double s_tot[4];
  double* s_ybbar = reinterpret_cast<double*>(smem_raw);
  const long long b = blockIdx.x;
  const IrfStats st = c.stats[b];
  const double s = st.mx / st.my;
  const double* p = c.params + b * c.NP;
  const double amps = c.amps[b];
  if (c.norm > 0) {
    // reverse of the norm > 0 branch of k_irf_finish: out_q = mean_k z_n, z_n = a_n y_n / M_side(n), y_n = s yc_n
    double acc[4] = {0.0, 0.0, 0.0, 0.0};   // amp1_bar, amp2_bar, Mb_bar, Mr_bar
    double sbar = 0.0;
    for (int n = threadIdx.x; n < g.W; n += kThreads) {
      const double zb = c.thry_bar[b * g.nbins + n / g.r] / (double)g.r;
      const double ycn = yc[b * g.W + n], y = s * ycn;
      double ybar;
      if (c.kind == 0) {
        const double lam = c.lam_min + (double)n * g.dlam;
        const bool blue = lam < p[P_LAM];
        const double a = blue ? p[P_AMP1] : p[P_AMP2], M = blue ? st.Mb : st.Mr;
        acc[blue ? 0 : 1] += zb * y / M;
        acc[blue ? 2 : 3] += -zb * a * y / (M * M);
        ybar = zb * a / M;
      } else {
        ybar = zb;
      }
      c.ycbar[b * g.W + n] = s * ybar;
      sbar += ybar * ycn;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < 4; k++) {
      const double v = warp_sum(acc[k]);
      if (lane == 0) sred[k * (kThreads / 32) + wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      double v = 0.0;
      for (int w = 0; w < kThreads / 32; w++) v += sred[threadIdx.x * (kThreads / 32) + w];
      s_tot[threadIdx.x] = v;
    }
    __syncthreads();
    sbar = warp_sum(sbar);
    if (lane == 0) sred[wid] = sbar;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < kThreads / 32; w++) t += sred[w];
      double* ab = c.amp_bar + b * 3;
      ab[0] = c.kind == 0 ? s_tot[0] : 0.0; ab[1] = c.kind == 0 ? s_tot[1] : 0.0; ab[2] = 0.0;
      if (c.kind == 0) {   // the side maxima route their cotangents to their arg-max samples (y = s yc there)
        c.ycbar[b * g.W + st.ib] += s * s_tot[2];
        c.ycbar[b * g.W + st.ir] += s * s_tot[3];
        t += s_tot[2] * yc[b * g.W + st.ib] + s_tot[3] * yc[b * g.W + st.ir];
      }
      c.ycbar[b * g.W + st.iy] += -t * st.mx / (st.my * st.my);   // s = mx / my
      c.xbar_max[b] = t / st.my;
    }
    return;
  }
  // out_q = A_q * yb_q / mb with A_q = a_q*amps (electron) or amp3*amps (ion); yb_q = s * mean_k yc
  double part[4] = {0.0, 0.0, 0.0, 0.0};  // mb_bar, amp1_bar|amp3_bar, amp2_bar, unused
  for (int q = threadIdx.x; q < g.nbins; q += kThreads) {
    const double lamq = c.lam_min + ((double)(q * g.r) + 0.5 * (double)(g.r - 1)) * g.dlam;
    double ysum = 0.0;
    for (int k = 0; k < g.r; k++) ysum += yc[b * g.W + q * g.r + k];
    const double ybq = s * ysum / (double)g.r;
    const double ob = c.thry_bar[b * g.nbins + q];
    double A;
    if (c.kind == 0) {
      const bool blue = lamq < p[P_LAM];
      A = (blue ? p[P_AMP1] : p[P_AMP2]) * amps;
      part[blue ? 1 : 2] += ob * amps * ybq / st.mb;
    } else {
      A = p[P_AMP3] * amps;
      part[1] += ob * amps * ybq / st.mb;
    }
    s_ybbar[q] = ob * A / st.mb;
    part[0] += -ob * A * ybq / (st.mb * st.mb);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < 3; k++) {
    double v = warp_sum(part[k]);
    if (lane == 0) sred[k * (kThreads / 32) + wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < kThreads / 32; w++) v += sred[threadIdx.x * (kThreads / 32) + w];
    s_tot[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s_ybbar[st.qm] += s_tot[0];  // max(yb) routes its cotangent to the arg-max bin
    double* ab = c.amp_bar + b * 3;
    if (c.kind == 0) { ab[0] = s_tot[1]; ab[1] = s_tot[2]; ab[2] = 0.0; }
    else             { ab[0] = 0.0; ab[1] = 0.0; ab[2] = s_tot[1]; }
  }
  __syncthreads();
  // y_n = s * yc_n  ->  ycbar_n = s * ybbar_{n/r} / r ;  s_bar = sum_n ybbar_{n/r}/r * yc_n
  double sbar = 0.0;
  for (int n = threadIdx.x; n < g.W; n += kThreads) {
    const double yb = s_ybbar[n / g.r] / (double)g.r;
    const double v = yc[b * g.W + n];
    c.ycbar[b * g.W + n] = s * yb;
    sbar += yb * v;
  }
  sbar = warp_sum(sbar);
  __syncthreads();
  if (lane == 0) sred[wid] = sbar;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; w++) t += sred[w];
    // s = mx/my
    c.ycbar[b * g.W + st.iy] += -t * st.mx / (st.my * st.my);
    c.xbar_max[b] = t / st.my;
  }
}

// xbar_m = sum_n ycbar_n G(n - m - half)  (+ the max(x) term).  dynamic smem: ycbar halo [kTile + 2K] | taps
__global__ void __launch_bounds__(kConvThreads) k_irf_bwd_conv(const IrfGeom g, const IrfCall c, double* __restrict__ xbar) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ntp = (2 * g.K + 1 + kConvR - 1) / kConvR * kConvR;
  double* s_y = reinterpret_cast<double*>(smem_raw) + 9;
  double* s_g = s_y + hpad(kTile + 2 * g.K + 8) + 1;
  const int tile = blockIdx.x % g.ntiles;
  const long long b = blockIdx.x / g.ntiles;
  const int m0 = tile * kTile;
  for (int i = threadIdx.x - 8; i < kTile + 2 * g.K + 8; i += kConvThreads) {
    const int n = m0 - g.K + i;
    s_y[hpad(i)] = (i >= 0 && i < kTile + 2 * g.K && n >= 0 && n < g.W) ? c.ycbar[b * g.W + n] : 0.0;
  }
  for (int i = threadIdx.x; i < ntp; i += kConvThreads) {
    const double d = ((double)(i - g.K) - g.half) * g.dlam;
    s_g[i] = i <= 2 * g.K ? g.gnorm * exp(-d * d * g.inv2s2) : 0.0;
  }
  __syncthreads();
  const int im = c.stats[b].im;
  // n = m + o, o in [-K, K];  s_y index of n: t + o + K = t + i
  const int t0 = threadIdx.x * kConvR;
  double acc[kConvR];
#pragma unroll
  for (int r = 0; r < kConvR; r++) acc[r] = 0.0;
  conv_tile8<1>(s_y, s_g, ntp, t0, acc);
#pragma unroll
  for (int r = 0; r < kConvR; r++) {
    const int m = m0 + t0 + r;
    if (m < g.W) xbar[b * g.W + m] = acc[r] + (m == im ? c.xbar_max[b] : 0.0);
  }
}

// ---- loss ---------------------------------------------------------------------------------------------------
// loss = scale * sum_{b,q} weight[q] * err(d, t);  theory_bar = d loss / d theory      (loss_function.py:386-418)
__global__ void __launch_bounds__(kThreads) k_loss(long long total, int n, const double* __restrict__ t, const double* __restrict__ d,
                                                  const double* __restrict__ w, double uncert, double scale, int method,
                                                  double* __restrict__ loss, double* __restrict__ tbar) {
  __shared__ double sred[kThreads / 32];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const double wq = w[i % n];
    double e = 0.0, ge = 0.0;
    if (wq != 0.0) {
      const double tv = t[i], dv = d[i], diff = dv - tv;
      if (method == 0)      { e = diff * diff / uncert; ge = -2.0 * diff / uncert; }                // l2
      else if (method == 1) { e = fabs(diff) / uncert; ge = (diff > 0 ? -1.0 : (diff < 0 ? 1.0 : 0.0)) / uncert; }  // l1
      else if (method == 2) { e = log(cosh(diff)); ge = -tanh(diff); }                               // log-cosh
      else                  { e = tv - dv * log(tv); ge = 1.0 - dv / tv; }                           // poisson
    }
    acc += wq * e;
    if (tbar) tbar[i] = scale * wq * ge;
  }
  acc = warp_sum(acc);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sred[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < kThreads / 32; k++) s += sred[k];
    atomicAdd(loss, scale * s);
  }
}

int check_cfg(const tsff_irf_cfg* c) {
  if (!c || c->W < 2 || c->nbins < 1 || c->W % c->nbins != 0 || !(c->stddev > 0.0)) {
    set_error("bad irf cfg (W must be a multiple of nbins, stddev > 0)");
    return TSFF_E_INVALID;
  }
  if (c->norm < 0) { set_error("PhysParams.norm must be >= 0"); return TSFF_E_INVALID; }
  if (c->kind != 0 && c->kind != 1) { set_error("irf kind must be 0 (electron) or 1 (ion)"); return TSFF_E_INVALID; }
  return TSFF_OK;
}

}  // namespace

extern "C" size_t tsff_irf_workspace_bytes(const tsff_irf_cfg* c, int64_t B) {
  if (check_cfg(c) || B < 1) return 0;
  return irf_layout(irf_geom(c), B).bytes + align_up((size_t)B * 8);
}
extern "C" size_t tsff_irf_saved_bytes(const tsff_irf_cfg* c, int64_t B) {
  if (check_cfg(c) || B < 1) return 0;
  return align_up((size_t)B * sizeof(IrfStats)) + align_up((size_t)B * c->W * 8);
}

extern "C" int tsff_irf_fwd(const tsff_irf_cfg* c, int64_t B, const double* modl, const double* params, int32_t NP,
                            const double* amps, const double* noise, double* thry, void* saved, void* ws, void* stream) {
  int rc = check_cfg(c);
  if (rc) return rc;
  if (B == 0) return TSFF_OK;
  if (!modl || !params || !amps || !thry || !saved || !ws || B < 1) { set_error("null argument"); return TSFF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const IrfGeom g = irf_geom(c);
  const IrfLayout L = irf_layout(g, B);
  char* w = static_cast<char*>(ws);
  char* sv = static_cast<char*>(saved);
  // conv(x) is kept in `saved` for the backward pass
  double* yc = (double*)(sv + align_up((size_t)B * sizeof(IrfStats)));
  double* pmax = (double*)(w + L.w_pmax);
  const size_t smem = (size_t)((kTile + 2 * g.K + 16) * 9 / 8 + 12 + 2 * g.K + 1 + kConvR) * 8;
  if (smem > 200 * 1024) { set_error("IRF too wide for shared memory (K=%d)", g.K); return TSFF_E_INVALID; }
  TSFF_SMEM_OPTIN(k_irf_conv);
  k_irf_conv<<<(unsigned)(B * g.ntiles), kConvThreads, smem, st>>>(g, modl, yc, pmax);
  TSFF_LAUNCH_OK("k_irf_conv");
  IrfCall call;
  memset(&call, 0, sizeof(call));
  call.kind = c->kind; call.norm = c->norm; call.NP = NP; call.lam_min = c->lam_min;
  call.params = params; call.amps = amps; call.noise = noise; call.thry = thry; call.stats = (IrfStats*)sv;
  k_irf_finish<<<(unsigned)B, kThreads, (size_t)g.nbins * 8, st>>>(g, call, yc, pmax);
  TSFF_LAUNCH_OK("k_irf_finish");
  return TSFF_OK;
}

extern "C" int tsff_irf_bwd(const tsff_irf_cfg* c, int64_t B, const double* params, int32_t NP, const double* amps,
                            const void* saved, const double* thry_bar, double* modl_bar, double* amp_bar, void* ws,
                            void* stream) {
  int rc = check_cfg(c);
  if (rc) return rc;
  if (B == 0) return TSFF_OK;
  if (!params || !amps || !saved || !thry_bar || !modl_bar || !amp_bar || !ws || B < 1) { set_error("null argument"); return TSFF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const IrfGeom g = irf_geom(c);
  const IrfLayout L = irf_layout(g, B);
  char* w = static_cast<char*>(ws);
  const char* sv = static_cast<const char*>(saved);
  const double* yc = (const double*)(sv + align_up((size_t)B * sizeof(IrfStats)));
  IrfCall call;
  memset(&call, 0, sizeof(call));
  call.kind = c->kind; call.norm = c->norm; call.NP = NP; call.lam_min = c->lam_min;
  call.params = params; call.amps = amps; call.stats = (IrfStats*)sv;
  call.thry_bar = thry_bar; call.ycbar = (double*)(w + L.w_ycbar); call.amp_bar = amp_bar;
  call.xbar_max = (double*)(w + L.bytes);
  k_irf_bwd_pre<<<(unsigned)B, kThreads, (size_t)g.nbins * 8, st>>>(g, call, yc);
  TSFF_LAUNCH_OK("k_irf_bwd_pre");
  const size_t smem = (size_t)((kTile + 2 * g.K + 16) * 9 / 8 + 12 + 2 * g.K + 1 + kConvR) * 8;
  TSFF_SMEM_OPTIN(k_irf_bwd_conv);
  k_irf_bwd_conv<<<(unsigned)(B * g.ntiles), kConvThreads, smem, st>>>(g, call, modl_bar);
  TSFF_LAUNCH_OK("k_irf_bwd_conv");
  return TSFF_OK;
}

extern "C" int tsff_loss_fwd_bwd(int64_t B, int32_t n, const double* theory, const double* data, const double* weight,
                                 double uncert, double scale, int method, double* loss_out, double* theory_bar,
                                 void* stream) {
  if (B == 0 && loss_out) return TSFF_OK;   // nothing to accumulate
  if (!theory || !data || !weight || !loss_out || B < 1 || n < 1 || method < 0 || method > 3) { set_error("bad argument"); return TSFF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = (long long)B * n;
  long long blocks = (total + kThreads - 1) / kThreads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_loss<<<(unsigned)blocks, kThreads, 0, st>>>(total, n, theory, data, weight, uncert, scale, method, loss_out, theory_bar);
  TSFF_LAUNCH_OK("k_loss");
  return TSFF_OK;
}
