// tsff_tree.cuh -- block-multipole ("treecode") evaluation of the principal-value sums.
//
// Reference: ratintn / ratcen, tsadar/core/physics/ratintn.py:4-52 (called at form_factor.py:266-268, 385-386).
// With the summation-by-parts form of tsff_pv.cuh,
//
//     I(xi) = sum_{i=1..M-1} p_i W(g_i) + p_0 E_0(g_0) + p_M E_M(g_M),      g_i = z_i - xi,   x = h/g,
//     W(g)   = sum_{j>=0} x^(2j+1) / ((2j+1)(j+1))                          (second difference of g ln|g|, |x| < 1)
//     E_0(g) = sum_{k>=1} (-1)^(k+1) x^k / (k(k+1)),    E_M(g) = sum_{k>=1} x^k / (k(k+1)),
//
// the nodes are cut into blocks of kTS = 64.  For a pole whose nearest node n lies in block bn, the three blocks
// wb0..wb0+2 (wb0 = clamp(bn-1, 0, NB-3)) are its NEAR WINDOW: their nodes are summed one by one (FP32 series
// x(1 + x^2/6 + x^4/15) for |i-n| > kMidHalf = 8, six terms for kNearHalf = 3 < |i-n| <= 8, exact FP64 logs for the
// 2*kNearHalf+1 nodes around the pole and for an end node inside the window).  Every other block is FAR and enters through a Laurent expansion about its centre c_b:
//
//     sum_{i in b} p_i W(g_i)  (+ end-node term)  =  sum_{m<K} A_{b,m} t^(m+1),       t = s h / (z_{c_b} - xi),  s = 32,
//     A_{b,m} = (1/s) sum_j C(m,2j) / ((2j+1)(j+1)) s^(-2j) mu_{m-2j},               mu_k = sum_{i in b} p_i (-(i-c_b)/s)^k
//
// |t| <= 31.5/96: K = 12 terms truncate at 2e-8 of the sum (tools/tree_proto.py).  Cost per pole: NB far blocks x
// (1 MUFU.RCP + K packed FMAs) + 192 near nodes, instead of (nodes) x (1 MUFU + ~9 FMA): 9x fewer issue slots at
// 4096 nodes.  The adjoint is the transpose: per block the local coefficients L_{b,m} = sum_p Ibar_p t^(m+1) are
// gathered over the far poles, then spread to the nodes with the same static weights q_m(e) that build A from p.
#pragma once
#include "tsff_pv.cuh"

namespace tsff {

constexpr int kTS = 64;              // nodes per block
constexpr int kTK = 12;              // expansion order
constexpr double kTs = 32.0;         // scale of the block-local coordinate (half a block)
constexpr int kTWin = 3 * kTS;       // near-window nodes per pole

TSFF_HD int tree_npad(int nodes) {   // nodes = M + 1; at least one full window
  int n = (nodes + kTS - 1) / kTS * kTS;
  return n < kTWin ? kTWin : n;
}

TSFF_HD float2 f2(float x, float y) {
  float2 r;
  r.x = x;
  r.y = y;
  return r;
}

TSFF_HD double tree_binom(int n, int k) {
  if (k < 0 || k > n) return 0.0;
  double r = 1.0;
  for (int i = 1; i <= k; i++) r = r * (double)(n - k + i) / (double)i;
  return r;
}
// coefficient of mu_{m-2j} in s * A_m
TSFF_HD double tree_cm(int m, int j) {
  double s2j = 1.0;
  for (int i = 0; i < 2 * j; i++) s2j /= kTs;
  return tree_binom(m, 2 * j) / ((double)(2 * j + 1) * (double)(j + 1)) * s2j;
}

// q_m(e): d A_{b,m} / d p_i for an interior node at offset e = i - c_b  (also the spreading weight of the adjoint)
TSFF_HD double tree_q(int m, double e) {
  const double eh = -e / kTs;
  double acc = 0.0;
  for (int j = 0; 2 * j <= m; j++) {
    double pw = 1.0;
    for (int i = 0; i < m - 2 * j; i++) pw *= eh;
    acc += tree_cm(m, j) * pw;
  }
  return acc / kTs;
}
// the same for an end node: `last` = false -> node 0 (kernel E_0), true -> node M (kernel E_M)
TSFF_HD double tree_q_end(int m, double e, bool last) {
  const double eh = -e / kTs;
  double acc = 0.0;
  for (int r = 0; r <= m; r++) {
    double pw = 1.0;
    for (int i = 0; i < r; i++) pw *= eh;
    double sc = 1.0;
    for (int i = 0; i < m + 1 - r; i++) sc /= kTs;
    const double sg = last ? 1.0 : (((m - r) & 1) ? -1.0 : 1.0);
    acc += sg * tree_binom(m, r) / ((double)(m + 1 - r) * (double)(m + 2 - r)) * pw * sc;
  }
  return acc;
}

// Expansion coefficients of block b (FP64).  pget(i) = p_i for 0 <= i <= M.  cmtab: kTK x (kTK/2) table of tree_cm.
// qend: [2][kTK] rows tree_q_end(m, -c, false), tree_q_end(m, M % kTS - c, true) (the static table's last two rows).
template <typename PGet>
TSFF_HD void tree_block_coeffs(PGet pget, int M, int b, const double* cmtab, const double* qend, double* A /*[kTK]*/) {
  double mu[kTK];
  for (int k = 0; k < kTK; k++) mu[k] = 0.0;
  const double c = (double)(kTS * b) + 0.5 * (double)(kTS - 1);
  for (int k = 0; k < kTS; k++) {
    const int i = kTS * b + k;
    if (i < 1 || i > M - 1) continue;
    const double eh = -((double)i - c) / kTs;
    double pw = pget(i);
    for (int q = 0; q < kTK; q++) {
      mu[q] += pw;
      pw *= eh;
    }
  }
  for (int m = 0; m < kTK; m++) {
    double a = 0.0;
    for (int j = 0; 2 * j <= m; j++) a += cmtab[m * (kTK / 2) + j] * mu[m - 2 * j];
    A[m] = a / kTs;
  }
  if (b == 0)
    for (int m = 0; m < kTK; m++) A[m] += pget(0) * qend[m];
  if (M >= kTS * b && M < kTS * (b + 1))
    for (int m = 0; m < kTK; m++) A[m] += pget(M) * qend[kTK + m];
}

// A pole in block-local form.  n = nearest node (clamped to [0, M]), delta = xi - z_n:
//   un = -n (exact),  ndh = -delta/h (|.| <= 1/2 unless the pole lies outside the grid),  wb0 = first window block
struct TreePole {
  float un, ndh;
  int wb0;
};
TSFF_HD TreePole tree_pole(double xi, double z0, double h, int M, int NB) {
  double r = rint((xi - z0) / h);
  if (!(r >= 0.0)) r = 0.0;  // also NaN
  if (r > (double)M) r = (double)M;
  TreePole t;
  t.un = (float)(-r);
  t.ndh = (float)(-(xi - (z0 + r * h)) / h);
  int bn = (int)r / kTS;
  int w = bn - 1;
  if (w < 0) w = 0;
  if (w > NB - 3) w = NB - 3;
  t.wb0 = w;
  return t;
}

constexpr float kInvTs = (float)(1.0 / kTs);

// Far field for R poles of one thread.  sAB: [NB][kTK] pairs (A_m, (m+1) A_m), 16-byte aligned rows.
// accI += sum_far A t^(m+1);  accJ += t^2 sum_far (m+1) A t^m   (dI/dxi = accJ / (s h), applied by the caller).
// Four blocks are summed in FP32 before they enter the FP64 accumulator: left and right of the pole the block sums are
// O(1) with opposite signs, and a running FP32 sum over all of them costs 2e-7 (tools/tree_proto.py).
template <int R>
TSFF_HD void tree_far(const float4* sAB, int NB, const TreePole (&tp)[R], double (&accI)[R], double (&accJ)[R]) {
  constexpr int H = kTK / 2;
  for (int b0 = 0; b0 < NB; b0 += 4) {
    float aI[R], aJ[R];
#pragma unroll
    for (int r = 0; r < R; r++) aI[r] = aJ[r] = 0.f;
#pragma unroll
    for (int bb = 0; bb < 4; bb++) {
      const int b = b0 + bb;
      if (b < NB) {
        const float cb = (float)(2 * b) + (float)(0.5 * (kTS - 1) / kTs);  // c_b / s, exact
        float2 tt[R], acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float u = fmaf(tp[r].un, kInvTs, cb);                       // (c_b - n)/s, exact
          const float g = fmaf(tp[r].ndh, kInvTs, u);                       // (z_cb - xi)/(s h)
          const bool far = (unsigned)(b - tp[r].wb0) > 2u;
          const float t = far ? rcp_approx(g) : 0.f;
          tt[r] = f2(t, t);
        }
        float4 c = sAB[b * H + H - 1];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = ffma2(f2(c.z, c.w), tt[r], f2(c.x, c.y));
#pragma unroll
        for (int q = H - 2; q >= 0; q--) {
          c = sAB[b * H + q];
#pragma unroll
          for (int r = 0; r < R; r++) {
            acc[r] = ffma2(acc[r], tt[r], f2(c.z, c.w));
            acc[r] = ffma2(acc[r], tt[r], f2(c.x, c.y));
          }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
          aI[r] = fmaf(tt[r].x, acc[r].x, aI[r]);
          aJ[r] = fmaf(tt[r].x * tt[r].x, acc[r].y, aJ[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      accI[r] += (double)aI[r];
      accJ[r] += (double)aJ[r];
    }
  }
}

// Near window of ONE pole: the 192 nodes of blocks wb0..wb0+2 except those with |i - n| <= kNearHalf.
// sW: node weights p_i (FP32; zero at i = 0, i >= M and in the padding), 16-byte aligned.
// accI += sum p_i W(g_i);  accJ += sum p_i h dW/dxi(g_i)   (dI/dxi = accJ / h, applied by the caller).
// Two nodes share one packed instruction; groups of 32 nodes enter the FP64 accumulators.
TSFF_HD void tree_near(const float* sW, const TreePole tp, double& accI, double& accJ) {
  const float4* w4 = reinterpret_cast<const float4*>(sW + kTS * tp.wb0);
  const float ub = (float)(kTS * tp.wb0) + tp.un;  // i0 - n, exact
  const float2 one = f2(1.f, 1.f);
  const float lim = (float)kNearHalf + 0.5f, mid = (float)kMidHalf + 0.5f;
  for (int q0 = 0; q0 < kTWin / 4; q0 += 8) {
    // does this group of 32 nodes [ub + 4 q0, ub + 4 q0 + 31] come within kMidHalf nodes of the pole?
    const float ulo = ub + (float)(4 * q0);
    const bool touch = (ulo <= mid) && (ulo + 31.f >= -mid);
    float2 aI = f2(0.f, 0.f), aJ = f2(0.f, 0.f);
    if (!TSFF_WARP_ANY(touch)) {
      // |i - n| > kMidHalf: W = x (1 + x^2/6 + x^4/15),  h dW/dxi = x^2 (1 + x^2/2 + x^4/3)
      const float2 c2 = f2(1.f / 6.f, 1.f / 6.f), c4 = f2(1.f / 15.f, 1.f / 15.f), d2 = f2(0.5f, 0.5f), d4 = f2(1.f / 3.f, 1.f / 3.f);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const float4 w = w4[q0 + q];
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
          const float u0 = ulo + (float)(4 * q + 2 * hlf);
          const float2 x = f2(rcp_approx(u0 + tp.ndh), rcp_approx((u0 + 1.f) + tp.ndh));
          const float2 s2 = fmul2(x, x);
          const float2 wv = hlf ? f2(w.z, w.w) : f2(w.x, w.y);
          aI = ffma2(fmul2(wv, x), ffma2(ffma2(s2, c4, c2), s2, one), aI);
          aJ = ffma2(fmul2(wv, s2), ffma2(ffma2(s2, d4, d2), s2, one), aJ);
        }
      }
    } else {
      // next to the pole: six terms (x <= 1/3.5: the seventh is 3e-9), nodes with |i - n| <= kNearHalf left to FP64
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const float4 w = w4[q0 + q];
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
          const float u0 = ulo + (float)(4 * q + 2 * hlf), u1 = u0 + 1.f;
          const float x0 = fabsf(u0) > lim ? rcp_approx(u0 + tp.ndh) : 0.f;
          const float x1 = fabsf(u1) > lim ? rcp_approx(u1 + tp.ndh) : 0.f;
          const float2 x = f2(x0, x1);
          const float2 s2 = fmul2(x, x);
          const float2 wv = hlf ? f2(w.z, w.w) : f2(w.x, w.y);
          float2 pI = f2(1.f / 66.f, 1.f / 66.f), pJ = f2(1.f / 6.f, 1.f / 6.f);
          pI = ffma2(pI, s2, f2(1.f / 45.f, 1.f / 45.f));
          pJ = ffma2(pJ, s2, f2(0.2f, 0.2f));
          pI = ffma2(pI, s2, f2(1.f / 28.f, 1.f / 28.f));
          pJ = ffma2(pJ, s2, f2(0.25f, 0.25f));
          pI = ffma2(pI, s2, f2(1.f / 15.f, 1.f / 15.f));
          pJ = ffma2(pJ, s2, f2(1.f / 3.f, 1.f / 3.f));
          pI = ffma2(pI, s2, f2(1.f / 6.f, 1.f / 6.f));
          pJ = ffma2(pJ, s2, f2(0.5f, 0.5f));
          pI = ffma2(pI, s2, one);
          pJ = ffma2(pJ, s2, one);
          aI = ffma2(fmul2(wv, x), pI, aI);
          aJ = ffma2(fmul2(wv, s2), pJ, aJ);
        }
      }
    }
    accI += (double)(aI.x + aI.y);
    accJ += (double)(aJ.x + aJ.y);
  }
}

// Exact FP64 part of I and dI/dxi for one pole: the interior nodes with |i - n| <= kNearHalf, and an end node when its
// block lies inside the pole's near window (otherwise the end node is part of that block's far expansion).
template <typename PGet>
TSFF_HD void tree_near_exact(double xi, double z0, double h, int M, int wb0, PGet pget, double& I, double& dI) {
  double rn = rint((xi - z0) / h);
  if (!(rn >= 0.0)) rn = 0.0;
  if (rn > (double)M) rn = (double)M;
  const int n = (int)rn;
  int lo = n - kNearHalf, hi = n + kNearHalf;
  if (lo < 1) lo = 1;
  if (hi > M - 1) hi = M - 1;
  const double ih = 1.0 / h;
  double sI = 0.0, sJ = 0.0;
  if (lo <= hi) {
    double gm = z0 + (double)(lo - 1) * h - xi;
    double lm = log(fmax(fabs(gm), 1e-300)), lc = log(fmax(fabs(gm + h), 1e-300));
    for (int i = lo; i <= hi; i++) {
      const double gc = z0 + (double)i * h - xi;
      const double gp = z0 + (double)(i + 1) * h - xi;
      const double lp = log(fmax(fabs(gp), 1e-300));
      const double p = pget(i);
      sI += p * (gp * lp - 2.0 * gc * lc + gm * lm) * ih;   // W
      sJ += -p * (lp - 2.0 * lc + lm) * ih;                 // dW/dxi
      gm = gc;
      lm = lc;
      lc = lp;
    }
  }
  if (wb0 == 0) {
    const double g0 = z0 - xi;
    const double l0 = log(fmax(fabs(g0), 1e-300)), l1 = log(fmax(fabs(g0 + h), 1e-300));
    const double p0 = pget(0);
    sI += p0 * (((g0 + h) * l1 - g0 * l0) * ih - 1.0 - l0);
    sJ += p0 * (-(l1 - l0) * ih + 1.0 / g0);
  }
  if ((unsigned)(M / kTS - wb0) <= 2u) {
    const double gM = z0 + (double)M * h - xi;
    const double lM = log(fmax(fabs(gM), 1e-300)), lM1 = log(fmax(fabs(gM - h), 1e-300));
    const double pM = pget(M);
    sI += pM * (((gM - h) * lM1 - gM * lM) * ih + 1.0 + lM);
    sJ += pM * (-(lM1 - lM) * ih - 1.0 / gM);
  }
  I = sI;
  dI = sJ;
}

}  // namespace tsff
