// tsff_pv.cuh -- the O(poles x nodes) principal-value ("rational integration") sums in FP32.
//
// Reference: ratintn / ratcen, tsadar/core/physics/ratintn.py:4-52, called at form_factor.py:266-268
// (fixed pole grid xi2 against nodes xi1) and :385-386 (pole = phase velocity, nodes = the f-table grid).
//
// Restatement used here (exact algebra, checked against the oracle's literal ratcen):  with uniform nodes
// z_i = z_0 + i h (i = 0..M, M = N-2: the reference's slices drop the last interval), node values p_i, g_i = z_i - xi,
// phi(g) = g ln|g|,
//
//     I(xi) = sum_{i<M} [dp_i + (pav_i - gav_i s_i) ln|g_{i+1}/g_i|]                      (ratcen, ratintn.py:41-52)
//           = sum_{i=1..M-1} p_i W(g_i) + p_0 E_0 + p_M E_M                               (summation by parts, twice)
//     W(g)  = [phi(g+h) - 2 phi(g) + phi(g-h)] / h
//     E_0   = [phi(g_1) - phi(g_0)]/h - 1 - ln|g_0|,   E_M = [phi(g_{M-1}) - phi(g_M)]/h + 1 + ln|g_M|
//
// I is linear in p with pole-dependent weights W, so the forward sweep (sum over nodes for each pole) and the
// adjoint sweep (sum over poles for each node) are transposes of one Cauchy-type kernel.  Far from the pole
//     W(g)      =  x (1 + x^2/6 + x^4/15 + ...),        x = h/g
//     dW/dxi    = (x/g)(1 + x^2/2 + x^4/3 + ...)
// (one MUFU.RCP and a handful of FFMA per pair; terms are O(h/g), so FP32 rounding stays ~1e-8 of the sum);
// the 2*kNearHalf+1 nodes next to the pole and the two end nodes are evaluated exactly in FP64 once per pole.
// A first version summed D_i g_i lg2|g_i| (one MUFU.LG2 per pair): it is 15% faster but its terms are ~100x
// larger than the sum, which costs two digits -- measured 2e-5 instead of 1e-6 on S at sharp EPW resonances.
// Precision of g: the pole is split in FP64 into its nearest node n and the remainder delta = xi - z_n
// (|delta| <= h/2); g_i = (i - n) h - delta is formed by FFMAs from exact small integers.
#pragma once
#include "tsff_math.cuh"

namespace tsff {

TSFF_HD float lg2_approx(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return log2f(x);
#endif
}

constexpr float kTinyG = 1e-30f;  // |g| clamp: keeps lg2 finite so that g*lg2|g| -> 0 when a pole sits on a node
constexpr int kPvBlk = 32;        // nodes per FP32 partial sum

// Split a pole position into (nearest node index, remainder) in FP64 and hand both to FP32 exactly.
TSFF_HD void pole_split(double xi, double z0, double h, int nnodes, float& u0, float& ndelta) {
  double r = rint((xi - z0) / h);
  if (!(r >= 0.0)) r = 0.0;  // also catches NaN
  if (r > (double)(nnodes - 1)) r = (double)(nnodes - 1);
  double delta = xi - (z0 + r * h);
  u0 = (float)(-r);           // exact: |r| < 2^24
  ndelta = (float)(-delta);
}

constexpr int kNearHalf = 8;   // nodes with |i - n_p| <= kNearHalf are handled exactly in FP64
constexpr int kMidHalf = 48;   // beyond this distance the series is cut after x^2 (x^4/15 < 1.3e-8 relative)

#if defined(__CUDA_ARCH__)
#define TSFF_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define TSFF_WARP_ANY(p) (p)
#endif

TSFF_HD float rcp_approx(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.0f / x;
#endif
}

// Blackwell packed FP32 (fma.rn.f32x2 -> SASS FFMA2): two FMAs per issued instruction.  The sweeps are limited by
// instruction dispatch next to the MUFU pipe (measured: T ~ 1.1 N_fma + 3 clk per warp-pair), so packing the two poles
// (or nodes) a thread owns into one register pair halves the FMA-pipe instruction count.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1,%2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
#else
inline float2 make_f2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
inline float2 ffma2(float2 a, float2 b, float2 c) { return make_f2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float2 fmul2(float2 a, float2 b) { return make_f2(a.x * b.x, a.y * b.y); }
#endif

struct FarCoef {  // series coefficients with the powers of h folded in
  float h_hi, h_lo, c2, c4, d2, d4;
};
TSFF_HD FarCoef far_coef(double h) {
  FarCoef c;
  c.h_hi = (float)h;
  c.h_lo = (float)(h - (double)c.h_hi);
  c.c2 = (float)(h * h / 6.0);
  c.c4 = (float)(h * h * h * h / 15.0);
  c.d2 = (float)(h * h / 2.0);
  c.d4 = (float)(h * h * h * h / 3.0);
  return c;
}

// Thread-owns-pole far-field accumulation over all node blocks.  sPh: pole-independent node weights p_i*h for the
// interior nodes 1..M-1 (zero at i = 0, i >= M and in the padding).  For R poles per thread:
//     accI[r] = sum_far p_i W(g_i),     accJ[r] = sum_far p_i dW/dxi(g_i),
// "far" = |i - n_p| > kNearHalf.  GRP = nodes per FP32 partial sum before it is folded into the FP64 accumulator
// (the terms are O(h/g), so 32-term FP32 partial sums cost nothing in accuracy; F2F shares the XU pipe with MUFU.RCP).
// Blocks farther than kMidHalf nodes from every pole of the warp take the short series (8 FMA-pipe ops + 1 MUFU.RCP
// per pair, both pipes balanced); the few blocks around the poles take the long series with the near-node mask.
// Must be called by all 32 lanes of a warp (warp vote).
template <int R, bool WITH_J, int GRP = 32>
TSFF_HD void pv_accumulate(const float* sPh, int nblk, const FarCoef cf, const float (&u0)[R], const float (&ndelta)[R],
                           double (&accI)[R], double (&accJ)[R]) {
  float u[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    u[r] = u0[r];
    accI[r] = 0.0;
    accJ[r] = 0.0;
  }
  const float4* s4 = reinterpret_cast<const float4*>(sPh);
  const float mid_lo = -(float)(kMidHalf + kPvBlk - 1), mid_hi = (float)kMidHalf;
  for (int b = 0; b < nblk; b++) {
    float gb[R];
    bool mid_blk = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
      gb[r] = fmaf(u[r], cf.h_hi, fmaf(u[r], cf.h_lo, ndelta[r]));
      mid_blk = mid_blk || (u[r] >= mid_lo && u[r] <= mid_hi);  // block [u, u+31] meets [-kMidHalf, kMidHalf]
    }
    if (!TSFF_WARP_ANY(mid_blk)) {
#pragma unroll
      for (int q0 = 0; q0 < kPvBlk / 4; q0 += GRP / 4) {
        float aI[R], aJ[R];
#pragma unroll
        for (int r = 0; r < R; r++) aI[r] = aJ[r] = 0.f;
#pragma unroll
        for (int q = q0; q < q0 + GRP / 4; q++) {
          const float4 d = s4[b * (kPvBlk / 4) + q];
          const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int r = 0; r < R; r++) {
              const float g = fmaf((float)(4 * q + c), cf.h_hi, gb[r]);
              const float rg = rcp_approx(g);
              const float s2 = rg * rg;
              aI[r] = fmaf(dd[c] * rg, fmaf(s2, cf.c2, 1.f), aI[r]);
              if (WITH_J) aJ[r] = fmaf(dd[c] * s2, fmaf(s2, cf.d2, 1.f), aJ[r]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
          accI[r] += (double)aI[r];
          if (WITH_J) accJ[r] += (double)aJ[r];
        }
      }
    } else {
      // rare path (the blocks around the poles): long series, and the near nodes masked (summed exactly elsewhere)
      float aI[R], aJ[R];
#pragma unroll
      for (int r = 0; r < R; r++) aI[r] = aJ[r] = 0.f;
#pragma unroll 4
      for (int k = 0; k < kPvBlk; k++) {
        const float w = sPh[b * kPvBlk + k];
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float g = fmaf((float)k, cf.h_hi, gb[r]);
          const bool far = fabsf(u[r] + (float)k) > (float)kNearHalf + 0.5f;
          const float rg = far ? rcp_approx(g) : 0.f;
          const float s2 = rg * rg;
          aI[r] = fmaf(w, rg * fmaf(fmaf(s2, cf.c4, cf.c2), s2, 1.f), aI[r]);
          if (WITH_J) aJ[r] = fmaf(w, s2 * fmaf(fmaf(s2, cf.d4, cf.d2), s2, 1.f), aJ[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        accI[r] += (double)aI[r];
        if (WITH_J) accJ[r] += (double)aJ[r];
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) u[r] += (float)kPvBlk;
  }
}

TSFF_HD double pv_phi(double g) { return g * log(fmax(fabs(g), 1e-300)); }

// Exact FP64 part of I and dI/dxi for one pole: the near nodes (interior ones among n-kNearHalf..n+kNearHalf) and
// the two end nodes.  `pget(i)` returns p_i as double.
template <typename PGet>
TSFF_HD void pv_near_exact(double xi, double z0, double h, int nodes, PGet pget, double& I, double& dI) {
  const int M = nodes - 1;
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
    double gm = z0 + (double)(lo - 1) * h - xi, gc = gm + h;
    double lm = log(fmax(fabs(gm), 1e-300)), lc = log(fmax(fabs(gc), 1e-300));
    for (int i = lo; i <= hi; i++) {
      const double gc_i = z0 + (double)i * h - xi;
      const double gp = z0 + (double)(i + 1) * h - xi;
      const double lp = log(fmax(fabs(gp), 1e-300));
      const double p = pget(i);
      sI += p * (gp * lp - 2.0 * gc_i * lc + gm * lm) * ih;   // W
      sJ += -p * (lp - 2.0 * lc + lm) * ih;                   // dW/dxi = -[phi'(g+h) - 2 phi'(g) + phi'(g-h)]/h
      gm = gc_i;
      lm = lc;
      lc = lp;
    }
    (void)gc;
  }
  const double g0 = z0 - xi, gM = z0 + (double)M * h - xi;
  const double l0 = log(fmax(fabs(g0), 1e-300)), l1 = log(fmax(fabs(g0 + h), 1e-300));
  const double lM = log(fmax(fabs(gM), 1e-300)), lM1 = log(fmax(fabs(gM - h), 1e-300));
  const double p0 = pget(0), pM = pget(M);
  sI += p0 * (((g0 + h) * l1 - g0 * l0) * ih - 1.0 - l0);
  sJ += p0 * (-(l1 - l0) * ih + 1.0 / g0);
  sI += pM * (((gM - h) * lM1 - gM * lM) * ih + 1.0 + lM);
  sJ += pM * (-(lM1 - lM) * ih - 1.0 / gM);
  I = sI;
  dI = sJ;
}

// FP64 twin of pv_accumulate (validation / "exact" mode): same algebra, log2 in double.
template <int R, bool WITH_J>
TSFF_HD void pv_accumulate_f64(const double* D, int nnodes, double h, const double (&g0)[R], double (&accI)[R],
                               double (&accJ)[R]) {
#pragma unroll
  for (int r = 0; r < R; r++) accI[r] = accJ[r] = 0.0;
  for (int i = 0; i < nnodes; i++) {
    const double d = D[i];
#pragma unroll
    for (int r = 0; r < R; r++) {
      double g = g0[r] + (double)i * h;
      double l = log2(fmax(fabs(g), 1e-300));
      accI[r] += d * g * l;
      if (WITH_J) accJ[r] += d * l;
    }
  }
}

// Endpoint terms and final values (FP64, once per pole).  g0 = z_0 - xi, gM = z_M - xi.
TSFF_HD void pv_finish(double accI, double accJ, double p0, double pM, double g0, double gM, double& I,
                       double& dIdxi) {
  double l0 = log(fmax(fabs(g0), 1e-300)), lM = log(fmax(fabs(gM), 1e-300));
  I = (pM - p0) + pM * lM - p0 * l0 + kLn2 * accI;
  dIdxi = -pM / gM + p0 / g0 - kLn2 * accJ;
}

// Node weights D_i from node values p (FP64 in, FP32 out); i in [0, M]; zero beyond.
TSFF_HD double pv_weight(const double* p, int M, double h, int i) {
  if (i > M) return 0.0;
  double sR = (i < M) ? (p[i + 1] - p[i]) / h : 0.0;
  double sL = (i > 0) ? (p[i] - p[i - 1]) / h : 0.0;
  return sR - sL;
}

}  // namespace tsff
