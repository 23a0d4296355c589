// tsff_tree.cuh -- block-multipole ("treecode") evaluation of the principal-value sums.
//
// Reference: ratintn / ratcen, tsadar/core/physics/ratintn.py:4-52 (called at form_factor.py:266-268, 385-386).
// With the summation-by-parts form of tsff_pv.cuh,
//
//     I(xi) = sum_{i=1..M-1} p_i W(g_i) + p_0 E_0(g_0) + p_M E_M(g_M),      g_i = z_i - xi,   x = h/g,
//     W(g)   = sum_{j>=0} x^(2j+1) / ((2j+1)(j+1))                          (second difference of g ln|g|, |x| < 1)
//     E_0(g) = sum_{k>=1} (-1)^(k+1) x^k / (k(k+1)),    E_M(g) = sum_{k>=1} x^k / (k(k+1)),
//
// the nodes are cut into blocks of 64 (level 1) and 256 (level 2).  For a pole whose nearest node n lies in level-1
// block bn and level-2 block Bn:
//   NEAR WINDOW  level-1 blocks wb0..wb0+2 (wb0 = clamp(bn-1, 0, NB-3)) = twelve level-0 groups of 16 nodes.  The pole's
//                group and its two neighbours are summed node by node (FP32 six-term series, exact FP64 logs for the
//                2*kNearHalf+1 nodes around the pole and for an end node inside the window); the other nine groups lie
//                at least 17 nodes from the pole and enter through their own level-0 expansions (interior nodes only:
//                the end-node terms of the window stay with the exact part);
//   FAR, level 1 the other level-1 blocks whose parent lies in the level-2 window w2..w2+2 (w2 = clamp(Bn-1, ...));
//   FAR, level 2 every level-2 block outside that window.
// A far block of S nodes enters through its Laurent expansion about the block centre c:
//
//     sum_{i in blk} p_i W(g_i) (+ end-node term) = sum_{m<K} A_m t^(m+1),   t = s h / (z_c - xi),  s = S/2,
//     A_m = (1/s) sum_j C(m,2j) / ((2j+1)(j+1)) s^(-2j) mu_{m-2j},           mu_k = sum_{i in blk} p_i (-(i-c)/s)^k
//
// |t| <= 0.332 at both levels; K = 14 truncates at ~3e-9 of the sum.  The two leading terms A_0 t + A_1 t^2 carry the
// magnitude (for an EPW-resonance pole essentially all of I sits in them) and are evaluated in FP64 from the FP32
// reciprocal plus one Newton step; the tail t^3 (A_2 + A_3 t + ...) and the derivative series run as one packed-FP32
// Horner recurrence (fma.rn.f32x2: lanes = (tail of I, dI/dxi)).  Cost per pole at 4096 nodes: ~32 block evaluations
// x (1 MUFU.RCP + 13 packed FMAs + 6 FP64 ops) + 192 near nodes, instead of 4094 x (1 MUFU + ~9 FMA).
// The adjoint is the transpose (tsff_pv_kernels.cuh): per block the local coefficients L_m = sum_p Ibar_p t^(m+1) are
// gathered over the far poles, then spread to the nodes with the same static weights q_m(e) that build A from p.
#pragma once
#include "tsff_pv.cuh"

namespace tsff {

constexpr int kTS0 = 16;             // nodes per level-0 block (the 16-node groups of the near window)
constexpr double kTs0 = 8.0;
constexpr int kTS = 64;              // nodes per level-1 block
constexpr int kTS2 = 256;            // nodes per level-2 block
constexpr int kTK = 14;              // expansion order of the forward sweep
// expansion order of the adjoint sweep: gradients need 1e-4, not 1e-5.  Measured at the benchmark shape (W = 1024, V = 4096,
// f32 tables) against the oracle's autograd: K = 12 -> fe_bar 3.0e-7, 4.18 ms per 16384 lineouts for the whole adjoint;
// K = 10 -> 5.1e-7, 3.79 ms; K = 8 -> 2.2e-6, 3.68 ms.  (Even values only: two orders per packed instruction.)
#ifndef TSFF_KTKA
#define TSFF_KTKA 10
#endif
constexpr int kTKA = TSFF_KTKA;
constexpr double kTs = 32.0;         // scale of the level-1 block-local coordinate (half a block)
constexpr double kTs2 = 128.0;
constexpr int kTWin = 3 * kTS;       // near-window nodes per pole
constexpr float kInvTs = (float)(1.0 / kTs);

TSFF_HD int tree_npad(int nodes) {   // nodes = M + 1, padded to whole level-2 blocks (zero weights)
  return (nodes + kTS2 - 1) / kTS2 * kTS2;
}

// Per-lineout blob staged into shared memory by one bulk copy (all sizes multiples of 16 bytes):
//   W   [npad]        float     node weights p_i (interior nodes 1..M-1, zero elsewhere)
//   AB1 [NB ][kTK/2]  float4    level-1 packed Horner coefficients, step m: (A_{m+2} or 0, (m+1) A_m)
//   AB2 [NB2][kTK/2]  float4    level-2 ...
//   LD1 [NB ]         double2   level-1 leading coefficients (A_0, A_1)
//   LD2 [NB2]         double2
//   AB0 [NB0][kTK/2]  float4    level-0 (16-node groups, interior nodes only, no end-node rows)
//   LD0 [NB0]         double2
struct TreeBlob {
  int npad, NB, NB2, NB0;
  int oW, oAB1, oAB2, oLD1, oLD2, oAB0, oLD0, bytes;  // byte offsets
};
TSFF_HD TreeBlob tree_blob(int npad) {
  TreeBlob t;
  t.npad = npad;
  t.NB = npad / kTS;
  t.NB2 = npad / kTS2;
  t.oW = 0;
  t.oAB1 = npad * 4;
  t.oAB2 = t.oAB1 + t.NB * (kTK / 2) * 16;
  t.oLD1 = t.oAB2 + t.NB2 * (kTK / 2) * 16;
  t.oLD2 = t.oLD1 + t.NB * 16;
  t.NB0 = npad / kTS0;
  t.oAB0 = t.oLD2 + t.NB2 * 16;
  t.oLD0 = t.oAB0 + t.NB0 * (kTK / 2) * 16;
  t.bytes = t.oLD0 + t.NB0 * 16;
  return t;
}

TSFF_HD float2 f2(float x, float y) {
  float2 r;
  r.x = x;
  r.y = y;
  return r;
}

TSFF_HD constexpr double tree_binom(int n, int k) {
  if (k < 0 || k > n) return 0.0;
  double r = 1.0;
  for (int i = 1; i <= k; i++) r = r * (double)(n - k + i) / (double)i;
  return r;
}
// coefficient of mu_{m-2j} in s * A_m
TSFF_HD constexpr double tree_cm(int m, int j, double s) {
  double s2j = 1.0;
  for (int i = 0; i < 2 * j; i++) s2j /= s;
  return tree_binom(m, 2 * j) / ((double)(2 * j + 1) * (double)(j + 1)) * s2j;
}
// x_o^k for the level-0 in-block coordinate x_o = -(o - (kTS0 - 1)/2) / kTs0 of offset o (exact dyadic rationals up to rounding)
TSFF_HD constexpr double tree_xpow0(int o, int k) {
  const double x = -((double)o - 0.5 * (double)(kTS0 - 1)) * (1.0 / kTs0);
  double v = 1.0;
  for (int q = 0; q < k; q++) v *= x;
  return v;
}
// moment translation child c (of four) -> parent: x_parent = (x_child - D) / 4 with D = 2 c - 3, so
//   mu_parent_k += sum_{j <= k} T(c, k, j) mu_child_j,   T = C(k, j) (-D)^(k-j) / 4^k     (exact in double)
TSFF_HD constexpr double tree_T(int c, int k, int j) {
  if (j > k) return 0.0;
  const double D = 2.0 * (double)c - 3.0;
  double v = tree_binom(k, j);
  for (int q = 0; q < k - j; q++) v *= -D;
  for (int q = 0; q < k; q++) v *= 0.25;
  return v;
}
// The same translation child c -> parent for a RUN-TIME c, as a Taylor shift: m'_k = 4^-k sum_j C(k,j) d^(k-j) m_j with d = 3 - 2c,
// by the in-place Pascal recurrence (45 multiply-adds for K = 10, no branches).  The prep kernel's lanes hold the four children
// of a parent side by side: a switch over four compile-time matrices made every warp execute all four 105-term products.
TSFF_HD void tree_translate_shift(int c, const double* m0, double* m) {
  const double d = 3.0 - 2.0 * (double)c;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < kTK; k++) m[k] = m0[k];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 1; i < kTK; i++) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = kTK - 1; k >= i; k--) m[k] = fma(d, m[k - 1], m[k]);
  }
  double sc = 1.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 1; k < kTK; k++) {
    sc *= 0.25;
    m[k] *= sc;
  }
}
// q_m(e): d A_m / d p_i for an interior node at offset e = i - c  (also the spreading weight of the adjoint)
TSFF_HD double tree_q(int m, double e, double s) {
  const double eh = -e / s;
  double acc = 0.0;
  for (int j = 0; 2 * j <= m; j++) {
    double pw = 1.0;
    for (int i = 0; i < m - 2 * j; i++) pw *= eh;
    acc += tree_cm(m, j, s) * pw;
  }
  return acc / s;
}
// the same for an end node: `last` = false -> node 0 (kernel E_0), true -> node M (kernel E_M)
TSFF_HD double tree_q_end(int m, double e, bool last, double s) {
  const double eh = -e / s;
  double acc = 0.0;
  for (int r = 0; r <= m; r++) {
    double pw = 1.0;
    for (int i = 0; i < r; i++) pw *= eh;
    double sc = 1.0;
    for (int i = 0; i < m + 1 - r; i++) sc /= s;
    const double sg = last ? 1.0 : (((m - r) & 1) ? -1.0 : 1.0);
    acc += sg * tree_binom(m, r) / ((double)(m + 1 - r) * (double)(m + 2 - r)) * pw * sc;
  }
  return acc;
}

// Static tables (depend only on the node count M+1): built once per context / call.
//   QA  [(kTS+2)][kTKA]   adjoint spreading weights q_m(e) of the kTS in-block offsets (level 1), then the rows of
//                         node 0 and node M
//   CM1 [kTK][kTK/2], QE1 [2][kTK]    level-1 forward tables (cm matrix, end-node rows)
//   CM2 [kTK][kTK/2], QE2 [2][kTK]    level-2 forward tables
constexpr int kTsQA = 0;
constexpr int kTsCM1 = (kTS + 2) * kTKA;
constexpr int kTsQE1 = kTsCM1 + kTK * (kTK / 2);
constexpr int kTsCM2 = kTsQE1 + 2 * kTK;
constexpr int kTsQE2 = kTsCM2 + kTK * (kTK / 2);
//   E1  [kTS][kTK]        powers x^k of the level-1 in-block coordinate x = -(i - c)/s of the kTS offsets
//   T12 [4][kTK][kTK]     moment translation child -> parent: mu2_k += sum_j T12[c][k][j] mu1_j(child c)
constexpr int kTsE1 = kTsQE2 + 2 * kTK;
constexpr int kTsT12 = kTsE1 + kTS * kTK;
//   CM0 [kTK][kTK/2]      level-0 cm matrix
constexpr int kTsCM0 = kTsT12 + 4 * kTK * kTK;
//   QA0 [kTS0][kTKA]      adjoint spreading weights of the level-0 in-block offsets (interior nodes only)
constexpr int kTsQA0 = kTsCM0 + kTK * (kTK / 2);
//   QA2 [(kTS2+2)][kTKA]  adjoint spreading weights of the level-2 in-block offsets, then the rows of node 0 and node M
constexpr int kTsQA2 = kTsQA0 + kTS0 * kTKA;
constexpr int kTreeStaticDoubles = kTsQA2 + (kTS2 + 2) * kTKA;
TSFF_HD double tree_static_entry(int i, int M) {
  const double c1 = 0.5 * (double)(kTS - 1), c2 = 0.5 * (double)(kTS2 - 1);
  if (i < kTsCM1) {
    const int k = i / kTKA, m = i % kTKA;
    if (k < kTS) return tree_q(m, (double)k - c1, kTs);
    if (k == kTS) return tree_q_end(m, 0.0 - c1, false, kTs);
    return tree_q_end(m, (double)(M % kTS) - c1, true, kTs);
  }
  if (i < kTsQE1) return tree_cm((i - kTsCM1) / (kTK / 2), (i - kTsCM1) % (kTK / 2), kTs);
  if (i < kTsCM2) {
    const int k = i - kTsQE1;
    return k < kTK ? tree_q_end(k, 0.0 - c1, false, kTs) : tree_q_end(k - kTK, (double)(M % kTS) - c1, true, kTs);
  }
  if (i < kTsQE2) return tree_cm((i - kTsCM2) / (kTK / 2), (i - kTsCM2) % (kTK / 2), kTs2);
  if (i < kTsE1) {
    const int k = i - kTsQE2;
    return k < kTK ? tree_q_end(k, 0.0 - c2, false, kTs2) : tree_q_end(k - kTK, (double)(M % kTS2) - c2, true, kTs2);
  }
  if (i < kTsT12) {
    const int o = (i - kTsE1) / kTK, k = (i - kTsE1) % kTK;
    const double x = -((double)o - c1) / kTs;
    double pw = 1.0;
    for (int q = 0; q < k; q++) pw *= x;
    return pw;
  }
  if (i >= kTsQA2) {
    const int k = (i - kTsQA2) / kTKA, m = (i - kTsQA2) % kTKA;
    if (k < kTS2) return tree_q(m, (double)k - c2, kTs2);
    if (k == kTS2) return tree_q_end(m, 0.0 - c2, false, kTs2);
    return tree_q_end(m, (double)(M % kTS2) - c2, true, kTs2);
  }
  if (i >= kTsQA0) return tree_q((i - kTsQA0) % kTKA, (double)((i - kTsQA0) / kTKA) - 0.5 * (double)(kTS0 - 1), kTs0);
  if (i >= kTsCM0) return tree_cm((i - kTsCM0) / (kTK / 2), (i - kTsCM0) % (kTK / 2), kTs0);
  {
    // x2 = (x1 - D)/4 with D = (c_child - c_parent)/s1 = 2 c - 3 for child c = 0..3:  x2^k = 4^-k sum_j C(k,j) (-D)^(k-j) x1^j
    const int c = (i - kTsT12) / (kTK * kTK), k = ((i - kTsT12) / kTK) % kTK, j = (i - kTsT12) % kTK;
    return tree_T(c, k, j);
  }
}

// Raw moments of block b of S nodes over the node subset {first, first+stride, ...} (a warp splits a block over its
// lanes and adds the partial moments).  pget(i) = p_i for 0 <= i <= M; interior nodes only.
template <typename PGet>
TSFF_HD void tree_block_moments(PGet pget, int M, int b, int S, double s, int first, int stride, double* mu /*[kTK]*/) {
  for (int k = 0; k < kTK; k++) mu[k] = 0.0;
  const double c = (double)(S * b) + 0.5 * (double)(S - 1);
  for (int k = first; k < S; k += stride) {
    const int i = S * b + k;
    if (i < 1 || i > M - 1) continue;
    const double eh = -((double)i - c) / s;
    double pw = pget(i);
    for (int q = 0; q < kTK; q++) {
      mu[q] += pw;
      pw *= eh;
    }
  }
}
// Expansion coefficients from the (complete) moments, end-node terms included.
template <typename PGet>
TSFF_HD void tree_coeffs_from_moments(PGet pget, int M, int b, int S, double s, const double* mu, const double* cmtab,
                                      const double* qend, double* A /*[kTK]*/) {
  for (int m = 0; m < kTK; m++) {
    double a = 0.0;
    for (int j = 0; 2 * j <= m; j++) a += cmtab[m * (kTK / 2) + j] * mu[m - 2 * j];
    A[m] = a / s;
  }
  if (!qend) return;   // level 0: interior nodes only
  if (b == 0)
    for (int m = 0; m < kTK; m++) A[m] += pget(0) * qend[m];
  if (M >= S * b && M < S * (b + 1))
    for (int m = 0; m < kTK; m++) A[m] += pget(M) * qend[kTK + m];
}
// Pack the coefficients of one block: ab[kTK/2] float4 = steps (2q, 2q+1), each (A_{m+2} or 0, (m+1) A_m); ld = (A_0, A_1)
TSFF_HD void tree_pack(const double* A, float4* ab, double* ld) {
  for (int q = 0; q < kTK / 2; q++) {
    const int m0 = 2 * q, m1 = 2 * q + 1;
    float4 v;
    v.x = m0 + 2 < kTK ? (float)A[m0 + 2] : 0.f;
    v.y = (float)((double)(m0 + 1) * A[m0]);
    v.z = m1 + 2 < kTK ? (float)A[m1 + 2] : 0.f;
    v.w = (float)((double)(m1 + 1) * A[m1]);
    ab[q] = v;
  }
  ld[0] = A[0];
  ld[1] = A[1];
}

// A pole in block-local form.  n = nearest node (clamped to [0, M]), delta = xi - z_n:
//   un = -n (exact),  ndh = -delta/h (|.| <= 1/2 unless the pole lies outside the grid),
//   wb0 = first level-1 window block,  w2 = first level-2 window block,  gw = (un + ndh) in double (= (z_0 - xi)/h)
struct TreePole {
  float un, ndh;
  int wb0, w2;
  double gw;
};
TSFF_HD TreePole tree_pole(double xi, double z0, double h, int M, int npad) {
  const double ih = fast_rcp(h);
  double r = rint((xi - z0) * ih);
  if (!(r >= 0.0)) r = 0.0;  // also NaN
  if (r > (double)M) r = (double)M;
  const double dh = -(xi - (z0 + r * h)) * ih;
  TreePole t;
  t.un = (float)(-r);
  t.ndh = (float)dh;
  t.gw = -r + dh;
  const int NB = npad / kTS, NB2 = npad / kTS2;
  int w = (int)r / kTS - 1;
  if (w > NB - 3) w = NB - 3;
  if (w < 0) w = 0;
  t.wb0 = w;
  int w2 = (int)r / kTS2 - 1;
  if (w2 > NB2 - 3) w2 = NB2 - 3;
  if (w2 < 0) w2 = 0;
  t.w2 = w2;
  return t;
}

// One far block for R poles: packed FP32 Horner (tail of I, dI/dxi) + FP64 leading terms.
//   cbs = c_b / s (exact in FP32 and FP64), invs = 1/s, ab = the block's kTK/2 float4, ld = (A_0, A_1)
//   use[r]: the block is far for pole r.  aI/aJ: FP32 partial sums, acc64: FP64 sum of the leading terms.
template <int R>
TSFF_HD void tree_far_block(const float4* ab, const double* ld, float cbs, float invs, const TreePole (&tp)[R],
                            const bool (&use)[R], float (&aI)[R], float (&aJ)[R], double (&acc64)[R]) {
  constexpr int H = kTK / 2;
  float2 tt[R], acc[R];
  float t[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const float u = fmaf(tp[r].un, invs, cbs);    // (c_b - n)/s, exact
    const float g = fmaf(tp[r].ndh, invs, u);     // (z_cb - xi)/(s h)
    t[r] = use[r] ? rcp_approx(g) : 0.f;
    tt[r] = f2(t[r], t[r]);
  }
  float4 c = ab[H - 1];
#pragma unroll
  for (int r = 0; r < R; r++) acc[r] = ffma2(f2(c.z, c.w), tt[r], f2(c.x, c.y));
#pragma unroll
  for (int q = H - 2; q >= 0; q--) {
    c = ab[q];
#pragma unroll
    for (int r = 0; r < R; r++) {
      acc[r] = ffma2(acc[r], tt[r], f2(c.z, c.w));
      acc[r] = ffma2(acc[r], tt[r], f2(c.x, c.y));
    }
  }
  const double A0 = ld[0], A1 = ld[1];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const float t2 = t[r] * t[r];
    aI[r] = fmaf(t2 * t[r], acc[r].x, aI[r]);
    aJ[r] = fmaf(t2, acc[r].y, aJ[r]);
    // leading terms in FP64: one Newton step on the FP32 reciprocal (t = 0 stays 0 for a masked block)
    const double g64 = fma(tp[r].gw, (double)invs, (double)cbs);
    double t64 = (double)t[r];
    t64 = fma(t64, fma(-g64, t64, 1.0), t64);
    acc64[r] = fma(t64, fma(A1, t64, A0), acc64[r]);
  }
}

// Far field for R poles of one thread.  blob: the staged per-lineout blob (shared memory), tb its layout.
// accI += sum_far A t^(m+1);  accJ1 / accJ2: t^2 sum (m+1) A t^m per level (dI/dxi = accJ1/(s1 h) + accJ2/(s2 h)).
// Must be called by all 32 lanes (warp min/max of the level-2 windows).
template <int R>
TSFF_HD void tree_far(const unsigned char* blob, const TreeBlob tb, const TreePole (&tp)[R], double (&accI)[R], double (&accJ1)[R],
                      double (&accJ2)[R]) {
  constexpr int H = kTK / 2;
  const float4* ab1 = reinterpret_cast<const float4*>(blob + tb.oAB1);
  const float4* ab2 = reinterpret_cast<const float4*>(blob + tb.oAB2);
  const double* ld1 = reinterpret_cast<const double*>(blob + tb.oLD1);
  const double* ld2 = reinterpret_cast<const double*>(blob + tb.oLD2);
  float aI[R], aJ[R];
  double a64[R];
  bool use[R];
#pragma unroll
  for (int r = 0; r < R; r++) { aI[r] = aJ[r] = 0.f; a64[r] = 0.0; }
  // ---- level 2: every block outside the level-2 window
  for (int B = 0; B < tb.NB2; B++) {
#pragma unroll
    for (int r = 0; r < R; r++) use[r] = (unsigned)(B - tp[r].w2) > 2u;
    tree_far_block<R>(ab2 + B * H, ld2 + 2 * B, (float)(2 * B) + (float)(0.5 * (kTS2 - 1) / kTs2), (float)(1.0 / kTs2), tp, use, aI, aJ, a64);
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    accI[r] += a64[r] + (double)aI[r];
    accJ2[r] += (double)aJ[r];
    aI[r] = aJ[r] = 0.f;
    a64[r] = 0.0;
  }
  // ---- level 1: children of the level-2 windows of this warp's poles, minus each pole's near window
  int wlo = tp[0].w2, whi = tp[0].w2;
#pragma unroll
  for (int r = 1; r < R; r++) { wlo = tp[r].w2 < wlo ? tp[r].w2 : wlo; whi = tp[r].w2 > whi ? tp[r].w2 : whi; }
#if defined(__CUDA_ARCH__)
  wlo = __reduce_min_sync(0xffffffffu, wlo);
  whi = __reduce_max_sync(0xffffffffu, whi);
#endif
  int bend = 4 * (whi + 3);
  if (bend > tb.NB) bend = tb.NB;
  for (int b0 = 4 * wlo; b0 < bend; b0 += 4) {
#pragma unroll
    for (int bb = 0; bb < 4; bb++) {
      const int b = b0 + bb;
#pragma unroll
      for (int r = 0; r < R; r++) use[r] = ((unsigned)((b >> 2) - tp[r].w2) <= 2u) && ((unsigned)(b - tp[r].wb0) > 2u);
      tree_far_block<R>(ab1 + b * H, ld1 + 2 * b, (float)(2 * b) + (float)(0.5 * (kTS - 1) / kTs), (float)(1.0 / kTs), tp, use, aI, aJ, a64);
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    accI[r] += a64[r] + (double)aI[r];
    accJ1[r] += (double)aJ[r];
  }
}

// Near window of ONE pole: the 192 nodes of blocks wb0..wb0+2 except those with |i - n| <= kNearHalf.
// blob: the staged per-lineout blob; its node weights p_i are FP32, zero at i = 0, i >= M and in the padding.
// returns I = sum p_i W(g_i),  J = sum p_i h dW/dxi(g_i)   (dI/dxi = J / h, applied by the caller).
// The window is walked in 12 groups of 16 nodes, ROTATED so that every lane starts at the group that holds its own
// pole: the pole's group and its two neighbours are summed node by node (six-term series with the exact-zone mask, two
// nodes per packed instruction), the other nine through their level-0 expansions (tree_far_block, one reciprocal and 13
// packed FMAs per group instead of 16 reciprocals and ~120 FP32 instructions) -- the same instruction stream for all
// lanes whatever their pole positions (no vote, no divergence).
struct TreeAcc {
  double I, J;
};
TSFF_HD_NOINLINE TreeAcc tree_near(const unsigned char* blob, const TreeBlob tb, const TreePole tp) {
  const float* sW = reinterpret_cast<const float*>(blob + tb.oW);
  double accI = 0.0;
  const float4* w4 = reinterpret_cast<const float4*>(sW + kTS * tp.wb0);
  const float ub = (float)(kTS * tp.wb0) + tp.un;  // i0 - n, exact (<= 0)
  const int qn = ((int)(-ub)) >> 4;                // group of the pole's nearest node, 0..11
  const float2 one = f2(1.f, 1.f);
  const float lim = (float)kNearHalf + 0.5f;
  float2 aJ = f2(0.f, 0.f);
  // ---- the pole's group and its neighbours: six terms (x <= 1/3.5: the seventh is 3e-9); |i - n| <= kNearHalf -> FP64
#pragma unroll 1
  for (int j = -1; j <= 1; j++) {
    int q = qn + j;
    q = q < 0 ? q + 12 : (q >= 12 ? q - 12 : q);
    const float ulo = ub + (float)(16 * q);
    float2 aI = f2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const float4 w = w4[4 * q + c];
#pragma unroll
      for (int hlf = 0; hlf < 2; hlf++) {
        const float u0 = ulo + (float)(4 * c + 2 * hlf), u1 = u0 + 1.f;
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
    accI += (double)aI.x + (double)aI.y;
  }
  const double accJ = (double)aJ.x + (double)aJ.y;
  // ---- the other nine groups (|i - n| >= 17, |t| = 8 h / |z_c - xi| <= 0.34): level-0 expansions
  const float4* ab0 = reinterpret_cast<const float4*>(blob + tb.oAB0);
  const double* ld0 = reinterpret_cast<const double*>(blob + tb.oLD0);
  const TreePole tp1[1] = {tp};
  const bool use1[1] = {true};
  float fI[1] = {0.f}, fJ[1] = {0.f};
  double f64[1] = {0.0};
#pragma unroll 1
  for (int j = 2; j <= 10; j++) {
    int q = qn + j;
    q = q >= 12 ? q - 12 : q;
    const int b0 = (kTS / kTS0) * tp.wb0 + q;
    tree_far_block<1>(ab0 + b0 * (kTK / 2), ld0 + 2 * b0, (float)(2 * b0) + (float)(0.5 * (kTS0 - 1) / kTs0), (float)(1.0 / kTs0), tp1,
                      use1, fI, fJ, f64);
  }
  TreeAcc r;
  r.I = accI + f64[0] + (double)fI[0];
  r.J = accJ + (double)fJ[0] * (1.0 / kTs0);   // the block sums carry 1 / (s0 h), the node sums 1 / h
  return r;
}

// Exact FP64 part of I and dI/dxi for one pole: the interior nodes with |i - n| <= kNearHalf, and an end node when its
// block lies inside the pole's near window (otherwise the end node is part of that block's far expansion).
// n = the pole's nearest node as split by tree_pole (n = -un): the FP32 near window masks exactly these nodes.
template <typename PGet>
TSFF_HD void tree_near_exact(double xi, double z0, double h, int M, int n, int wb0, PGet pget, double& I, double& dI) {
  int lo = n - kNearHalf, hi = n + kNearHalf;
  if (lo < 1) lo = 1;
  if (hi > M - 1) hi = M - 1;
  const double ih = fast_rcp(h);
  double sI = 0.0, sJ = 0.0;
  if (lo <= hi) {
    double gm = z0 + (double)(lo - 1) * h - xi;
    double lm = log_abs(gm), lc = log_abs(gm + h);
    for (int i = lo; i <= hi; i++) {
      const double gc = z0 + (double)i * h - xi;
      const double gp = z0 + (double)(i + 1) * h - xi;
      const double lp = log_abs(gp);
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
    const double l0 = log_abs(g0), l1 = log_abs(g0 + h);
    const double p0 = pget(0);
    sI += p0 * (((g0 + h) * l1 - g0 * l0) * ih - 1.0 - l0);
    sJ += p0 * (-(l1 - l0) * ih + fast_rcp(g0));
  }
  if ((unsigned)(M / kTS - wb0) <= 2u) {
    const double gM = z0 + (double)M * h - xi;
    const double lM = log_abs(gM), lM1 = log_abs(gM - h);
    const double pM = pget(M);
    sI += pM * (((gM - h) * lM1 - gM * lM) * ih + 1.0 + lM);
    sJ += pM * (-(lM1 - lM) * ih - fast_rcp(gM));
  }
  I = sI;
  dI = sJ;
}

}  // namespace tsff
