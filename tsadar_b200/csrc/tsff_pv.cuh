// tsff_pv.cuh -- shared pieces of the principal-value ("rational integration") sums.
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
// adjoint sweep (sum over poles for each node) are transposes of one Cauchy-type kernel.  The FP32 production path
// (block-multipole far field, near-window series, exact FP64 logs next to the pole) is in tsff_tree.cuh; this header
// holds the packed-FP32 helpers and the plain FP64 validation path (log form, O(poles x nodes)).
// History: a first version summed D_i g_i lg2|g_i| pairwise in FP32 (one MUFU.LG2 per pair): its terms are ~100x
// larger than the sum, which costs two digits -- measured 2e-5 instead of 1e-6 on S at sharp EPW resonances; the
// second summed the W-series pairwise (one MUFU.RCP per pair, profiles/r01a_*): accurate but 9x the instructions.
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

constexpr int kNearHalf = 3;   // nodes with |i - n_p| <= kNearHalf are handled exactly in FP64

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

TSFF_HD double pv_phi(double g) { return g * log_abs(g); }

// FP64 validation path ("exact" mode): log form  I = sum_i D_i g_i ln|g_i| + endpoint terms, D = second difference
// of p / h (pv_weight), log2 in double.
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
  double l0 = log_abs(g0), lM = log_abs(gM);
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
