// tsff_pv.cuh -- the O(poles x nodes) principal-value ("rational integration") sums in FP32.
//
// Reference: ratintn / ratcen, tsadar/core/physics/ratintn.py:4-52, called at form_factor.py:266-268
// (fixed pole grid xi2 against nodes xi1) and :385-386 (pole = phase velocity, nodes = the f-table grid).
//
// Restatement used here (exact algebra, checked against the oracle's literal ratcen):  with uniform nodes
// z_i = z_0 + i h (i = 0..M, M = N-2: the reference's slices drop the last interval), node values p_i,
// s_i = (p_{i+1}-p_i)/h, g_i = z_i - xi,
//
//     I(xi) = sum_{i<M} [dp_i + (pav_i - gav_i s_i) ln|g_{i+1}/g_i|]
//           = (p_M - p_0) + p_M ln|g_M| - p_0 ln|g_0| + sum_{i=0..M} D_i g_i ln|g_i|,
//     D_0 = s_0,  D_i = s_i - s_{i-1},  D_M = -s_{M-1}                      (summation by parts)
//     dI/dxi = -p_M/g_M + p_0/g_0 - sum_i D_i ln|g_i|                        (sum_i D_i = 0)
//
// so one pass needs ONE MUFU.LG2 per (pole,node) pair and yields both I and dI/dxi; the weights D_i do not
// depend on the pole and are staged once in shared memory.  Precision: the pole is
// split in FP64 into its nearest node n and the remainder delta = xi - z_n (|delta| <= h/2), and
// g_i = (i - n) h - delta is formed by one FFMA from exact small integers, so g keeps full FP32 relative
// accuracy next to the pole.  FP32 partial sums cover 32 nodes and are folded into FP64 accumulators.
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

// Thread-owns-pole accumulation over all node blocks.  sD: pole-independent weights D_i (zero padded to a
// multiple of 32).  For R poles per thread:  accI[r] = sum_i D_i g_i lg2|g_i|,  accJ[r] = sum_i D_i lg2|g_i|.
template <int R, bool WITH_J>
TSFF_HD void pv_accumulate(const float* sD, int nblk, float h, const float (&u0)[R], const float (&ndelta)[R],
                           double (&accI)[R], double (&accJ)[R]) {
  float u[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    u[r] = u0[r];
    accI[r] = 0.0;
    accJ[r] = 0.0;
  }
  const float4* sD4 = reinterpret_cast<const float4*>(sD);
  for (int b = 0; b < nblk; b++) {
    float gb[R], aI[R], aJ[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      gb[r] = fmaf(u[r], h, ndelta[r]);
      aI[r] = 0.f;
      aJ[r] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < kPvBlk / 4; q++) {
      const float4 d = sD4[b * (kPvBlk / 4) + q];
      const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float g = fmaf((float)(4 * q + c), h, gb[r]);
          const float l = lg2_approx(fmaxf(fabsf(g), kTinyG));
          aI[r] = fmaf(dd[c], g * l, aI[r]);
          if (WITH_J) aJ[r] = fmaf(dd[c], l, aJ[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      accI[r] += (double)aI[r];
      if (WITH_J) accJ[r] += (double)aJ[r];
      u[r] += (float)kPvBlk;
    }
  }
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
