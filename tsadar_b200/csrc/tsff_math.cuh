// tsff_math.cuh -- per-lineout and per-(omega,angle) arithmetic of the Thomson-scattering form factor,
// forward and hand-written adjoint, shared by all kernels.  FP64 throughout (the per-omega
// assembly must stay in double; only the O(P*V) principal-value sums run in FP32, see tsff_pv.cuh).
//
// Reference being restated (ergodicio/tsadar): tsadar/core/physics/form_factor.py:182-296 (1V),
// :349-388 (calc_chi_vals), ratintn.py:4-52.  Every function names the lines it follows.
//
// The functions are __host__ __device__ so that tests/hostsim can execute exactly this arithmetic on
// the CPU against the oracle's autograd; the product only ever runs them inside CUDA kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TSFF_HD __host__ __device__ __forceinline__
#define TSFF_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define TSFF_HD inline
#define TSFF_HD_NOINLINE inline
struct float4 { float x, y, z, w; };  // host-only stand-ins (tests/hostsim)
struct float2 { float x, y; };
#endif

#define TSFF_MAX_IONS 4

namespace tsff {

// form_factor.py:123-125, 207-209
constexpr double kC = 2.99792458e10;
constexpr double kMe = 510.9896 / (kC * kC);
constexpr double kMp = kMe * 1836.1;
constexpr double kRe = 2.8179e-13;
constexpr double kPi = 3.14159265358979323846;
constexpr double kCst2 = 4.0 * kPi * kC * kC * kRe;  // constants^2 = 4 pi Esq / Me, Esq = Me C^2 re
constexpr double kOmgLNum = 2.0 * kPi * 1e7 * kC;
constexpr double kLn2 = 0.693147180559945309417;
constexpr double kInvSqrt2Pi = 0.398942280401432677940;

// parameter block layout, one row per lineout (include/tsff.h documents the same indices)
enum ParamIdx { P_TE = 0, P_NE, P_LAM, P_VA, P_UD, P_NE_GRAD, P_TE_GRAD, P_AMP1, P_AMP2, P_AMP3, P_ION0 };
enum IonIdx { ION_A = 0, ION_Z, ION_TI, ION_FRACT, ION_STRIDE };

// per (lineout, gradient point) scalars
struct LG {
  double ne_g, omgL, omgpe2, kL, vTe, Va6, ud6;
  double c_kldi[TSFF_MAX_IONS], inv_s2vTi[TSFF_MAX_IONS], ioncf[TSFF_MAX_IONS];
};

TSFF_HD void lg_zero(LG& b) {
  b.ne_g = b.omgL = b.omgpe2 = b.kL = b.vTe = b.Va6 = b.ud6 = 0.0;
  for (int i = 0; i < TSFF_MAX_IONS; i++) b.c_kldi[i] = b.inv_s2vTi[i] = b.ioncf[i] = 0.0;
}
constexpr int kLGDoubles = 7 + 3 * TSFF_MAX_IONS;

TSFF_HD double grad_factor(double grad, int g, int G) {
  // jnp.linspace(1 - grad/200, 1 + grad/200, G)[g]  (form_factor.py:182-195); G == 1 -> lower end
  double lo = 1.0 - grad / 200.0;
  if (G <= 1) return lo;
  return lo + (double)g * ((grad / 100.0) / (double)(G - 1));
}
TSFF_HD double grad_factor_d(int g, int G) {  // d factor / d grad
  if (G <= 1) return -1.0 / 200.0;
  return -1.0 / 200.0 + (double)g / (100.0 * (double)(G - 1));
}

// form_factor.py:182-243 (lineout-level part)
TSFF_HD void lg_forward(const double* p, int nI, int g, int G, double lam_shift, LG& o) {
  double ne_g = 1.0e20 * p[P_NE] * grad_factor(p[P_NE_GRAD], g, G);
  double Te_g = p[P_TE] * grad_factor(p[P_TE_GRAD], g, G);
  o.ne_g = ne_g;
  o.omgL = kOmgLNum / (p[P_LAM] + lam_shift);
  o.omgpe2 = kCst2 * ne_g;
  o.kL = sqrt(o.omgL * o.omgL - o.omgpe2) / kC;
  o.vTe = sqrt(Te_g / kMe);
  o.Va6 = p[P_VA] * 1e6;
  o.ud6 = p[P_UD] * 1e6;
  double Zbar = 0.0;
  for (int i = 0; i < nI; i++) Zbar += p[P_ION0 + i * ION_STRIDE + ION_Z] * p[P_ION0 + i * ION_STRIDE + ION_FRACT];
  for (int i = 0; i < nI; i++) {
    const double* q = p + P_ION0 + i * ION_STRIDE;
    double Mi = q[ION_A] * kMp;
    double ni = q[ION_FRACT] * ne_g / Zbar;
    double omgpi = sqrt(kCst2) * q[ION_Z] * sqrt(ni * kMe / Mi);
    double vTi = sqrt(q[ION_TI] / Mi);
    o.c_kldi[i] = vTi / omgpi;
    o.inv_s2vTi[i] = 1.0 / (1.4142135623730951 * vTi);
    o.ioncf[i] = q[ION_FRACT] * q[ION_Z] * q[ION_Z] / (Zbar * vTi);
  }
}

// adjoint of lg_forward: accumulates into pbar[NP]
TSFF_HD void lg_backward(const double* p, int nI, int g, int G, double lam_shift, const LG& b, double* pbar) {
  double fne = grad_factor(p[P_NE_GRAD], g, G), fTe = grad_factor(p[P_TE_GRAD], g, G);
  double ne_g = 1.0e20 * p[P_NE] * fne;
  double Te_g = p[P_TE] * fTe;
  double lamL = p[P_LAM] + lam_shift;
  double omgL = kOmgLNum / lamL;
  double omgpe2 = kCst2 * ne_g;
  double kLC = sqrt(omgL * omgL - omgpe2);  // kL * C
  double vTe = sqrt(Te_g / kMe);
  double ne_g_bar = b.ne_g, omgL_bar = b.omgL, omgpe2_bar = b.omgpe2, Te_g_bar = 0.0;
  // kL = sqrt(omgL^2 - omgpe2)/C
  omgL_bar += b.kL * omgL / (kLC * kC);
  omgpe2_bar += -b.kL / (2.0 * kLC * kC);
  // vTe = sqrt(Te_g/Me)
  Te_g_bar += b.vTe / (2.0 * vTe * kMe);
  pbar[P_VA] += b.Va6 * 1e6;
  pbar[P_UD] += b.ud6 * 1e6;
  // ions
  double Zbar = 0.0;
  for (int i = 0; i < nI; i++) Zbar += p[P_ION0 + i * ION_STRIDE + ION_Z] * p[P_ION0 + i * ION_STRIDE + ION_FRACT];
  double Zbar_bar = 0.0;
  for (int i = 0; i < nI; i++) {
    const double* q = p + P_ION0 + i * ION_STRIDE;
    double* qb = pbar + P_ION0 + i * ION_STRIDE;
    double Mi = q[ION_A] * kMp;
    double ni = q[ION_FRACT] * ne_g / Zbar;
    double sq = sqrt(ni * kMe / Mi);
    double omgpi = sqrt(kCst2) * q[ION_Z] * sq;
    double vTi = sqrt(q[ION_TI] / Mi);
    double vTi_bar = 0.0, omgpi_bar = 0.0;
    // c_kldi = vTi/omgpi
    vTi_bar += b.c_kldi[i] / omgpi;
    omgpi_bar += -b.c_kldi[i] * vTi / (omgpi * omgpi);
    // inv_s2vTi = 1/(sqrt2 vTi)
    vTi_bar += -b.inv_s2vTi[i] / (1.4142135623730951 * vTi * vTi);
    // ioncf = fract Z^2 / (Zbar vTi)
    double ioncf = q[ION_FRACT] * q[ION_Z] * q[ION_Z] / (Zbar * vTi);
    qb[ION_FRACT] += b.ioncf[i] * ioncf / q[ION_FRACT];
    qb[ION_Z] += b.ioncf[i] * 2.0 * ioncf / q[ION_Z];
    Zbar_bar += -b.ioncf[i] * ioncf / Zbar;
    vTi_bar += -b.ioncf[i] * ioncf / vTi;
    // vTi = sqrt(Ti/Mi)
    qb[ION_TI] += vTi_bar / (2.0 * vTi * Mi);
    // omgpi = cst Z sqrt(ni Me/Mi)
    qb[ION_Z] += omgpi_bar * omgpi / q[ION_Z];
    double ni_bar = omgpi_bar * omgpi / (2.0 * ni);
    // ni = fract ne_g / Zbar
    qb[ION_FRACT] += ni_bar * ne_g / Zbar;
    ne_g_bar += ni_bar * q[ION_FRACT] / Zbar;
    Zbar_bar += -ni_bar * ni / Zbar;
  }
  for (int i = 0; i < nI; i++) {
    const double* q = p + P_ION0 + i * ION_STRIDE;
    double* qb = pbar + P_ION0 + i * ION_STRIDE;
    qb[ION_Z] += Zbar_bar * q[ION_FRACT];
    qb[ION_FRACT] += Zbar_bar * q[ION_Z];
  }
  // omgpe2 = cst2 ne_g ; omgL = num/lamL
  ne_g_bar += omgpe2_bar * kCst2;
  pbar[P_LAM] += -omgL_bar * omgL / lamL;
  // ne_g = 1e20 ne fne ; Te_g = Te fTe
  pbar[P_NE] += ne_g_bar * 1.0e20 * fne;
  pbar[P_NE_GRAD] += ne_g_bar * 1.0e20 * p[P_NE] * grad_factor_d(g, G);
  pbar[P_TE] += Te_g_bar * fTe;
  pbar[P_TE_GRAD] += Te_g_bar * p[P_TE] * grad_factor_d(g, G);
}

// ------------------------------------------------------------------------------------------------
// FP64 reciprocal, square root and exp(y <= 0) at ~1-2 ulp from the MUFU seeds (rcp/rsqrt.approx.ftz.f64) and Newton
// steps: 5-20 instructions instead of the 30-60 of the IEEE-exact library paths.  The per-pole FP64 stage runs ~15
// divisions, 2 square roots and 1 exponential per pole; the reference is float64 and parity is asked to 1e-5.
// Arguments are positive normal numbers (the kernels' NaN policy is "propagate": NaN in -> NaN out still holds).
// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
// Polynomial coefficients as constant-bank operands: an FP64 immediate costs two UMOV per use (the compiler re-materialises the
// 64-bit constants of a Horner chain through one uniform-register pair: 12 % of k_table_fwd's issue slots), a c[bank][offset]
// operand costs nothing.
static __constant__ double kExpC[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
                                        1.0 / 5040.0,      1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,     1.0 / 6.0,
                                        0.5,               1.0,              1.0};
static __constant__ double kLogC[10] = {1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0,
                                        1.0 / 9.0,  1.0 / 7.0,  1.0 / 5.0,  1.0 / 3.0,  1.0};
static __constant__ double kLn2Split[3] = {1.4426950408889634, 0.693147180369123816490, 1.90821492927058770002e-10};
#endif
TSFF_HD double fast_rcp(double d) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  return r;
#else
  return 1.0 / d;
#endif
}
TSFF_HD double fast_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double h = 0.5 * r, s = x * r;             // Goldschmidt: s -> sqrt(x), h -> 1/(2 sqrt(x))
  double e = fma(-s, h, 0.5);
  s = fma(s, e, s);
  h = fma(h, e, h);
  e = fma(-s, h, 0.5);
  s = fma(s, e, s);
  h = fma(h, e, h);
  return fma(fma(-s, s, x), h, s);
#else
  return sqrt(x);
#endif
}
TSFF_HD double fast_exp_neg(double y) {       // exp(y), y <= 0
#if defined(__CUDA_ARCH__)
  if (!(y > -700.0)) return y == y ? 0.0 : y;
  const double n = rint(y * kLn2Split[0]);
  double f = fma(-n, kLn2Split[1], y);   // ln2 split hi/lo
  f = fma(-n, kLn2Split[2], f);
  double p = kExpC[0];
#pragma unroll
  for (int i = 1; i < 13; i++) p = fma(p, f, kExpC[i]);
  return __hiloint2double(__double2hiint(p) + ((int)n << 20), __double2loint(p));   // p * 2^n, n >= -1010
#else
  return exp(y);
#endif
}

// ------------------------------------------------------------------------------------------------
// ln(x) for a positive, normal double to ~2e-15 relative-to-1 accuracy in ~30 instructions (the CUDA library log costs
// ~100 SASS instructions inline, and the exact near-pole terms of the PV sums need 11 logs per pole).
// x = 2^e m, m in [sqrt(1/2), sqrt(2)):  ln x = e ln2 + 2 atanh(s),  s = (m-1)/(m+1),  |s| <= 0.1716.
// The quotient comes from rcp.approx.ftz.f64 (MUFU.RCP64H) and two Newton steps.
// ------------------------------------------------------------------------------------------------
TSFF_HD double fast_log_pos(double x) {
#if defined(__CUDA_ARCH__)
  int hi = __double2hiint(x), lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;       // m in [1, 2)
  if (hi >= 0x3ff6a09e) {                     // m >= sqrt(2) (to 20 bits): halve
    hi -= 0x00100000;
    e += 1;
  }
  const double m = __hiloint2double(hi, lo);
  const double d = m + 1.0;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  const double s = (m - 1.0) * r;
#else
  int e;
  double m = frexp(x, &e);                    // m in [0.5, 1)
  if (m < 0.70710678118654752) { m *= 2.0; e -= 1; }
  const double s = (m - 1.0) / (m + 1.0);
#endif
  const double s2 = s * s;
#if defined(__CUDA_ARCH__)
  double p = kLogC[0];
#pragma unroll
  for (int i = 1; i < 10; i++) p = fma(p, s2, kLogC[i]);
#else
  double p = 1.0 / 19.0;
  p = fma(p, s2, 1.0 / 17.0);
  p = fma(p, s2, 1.0 / 15.0);
  p = fma(p, s2, 1.0 / 13.0);
  p = fma(p, s2, 1.0 / 11.0);
  p = fma(p, s2, 1.0 / 9.0);
  p = fma(p, s2, 1.0 / 7.0);
  p = fma(p, s2, 1.0 / 5.0);
  p = fma(p, s2, 1.0 / 3.0);
  p = fma(p, s2, 1.0);
#endif
  return fma((double)e, kLn2, 2.0 * s * p);
}
// ln|g| with the clamp the PV code uses (a pole exactly on a node: g ln|g| -> 0)
TSFF_HD double log_abs(double g) { return fast_log_pos(fmax(fabs(g), 1e-300)); }

// ------------------------------------------------------------------------------------------------
// Z' table lookup: jnp.interp(xii, xi2, Zpi[0], left=xii**-2, right=xii**-2), imag with 0 fills
// (form_factor.py:247-248).  xi2 = arange(-8.2, 8.2, 0.01) is uniform; zr/zi hold Zpi rows.
// Returns values and d/dxii.
// ------------------------------------------------------------------------------------------------
struct ZTab {
  const double* zr;
  const double* zi;
  int n;        // 1640
  double x0;    // -8.2
  double h;     // 0.01
  double xlast; // xi2[n-1] as computed by arange
};

TSFF_HD void zprime_lerp(const ZTab& z, double x, double& zr, double& zi, double& dzr, double& dzi) {
  if (x < z.x0 || x > z.xlast) {
    double x2 = x * x;
    zr = fast_rcp(x2);
    dzr = -2.0 * zr * fast_rcp(x);
    zi = 0.0;
    dzi = 0.0;
    return;
  }
  const double ih = fast_rcp(z.h);
  double u = (x - z.x0) * ih;
  int i = (int)u;
  if (i > z.n - 2) i = z.n - 2;
  if (i < 0) i = 0;
  double t = u - (double)i;
  double r0 = z.zr[i], r1 = z.zr[i + 1], i0 = z.zi[i], i1 = z.zi[i + 1];
  zr = r0 + t * (r1 - r0);
  zi = i0 + t * (i1 - i0);
  dzr = (r1 - r0) * ih;
  dzi = (i1 - i0) * ih;
}

// Frozen-cell variants (second-order path).  jax.hessian of a jnp.interp treats the cell index as a
// constant: inside a cell the interpolant is linear in x, so its second derivative is zero and the kinks at the nodes do
// not contribute.  A finite difference of the adjoint gradient across a node DOES see the kink (on average it reproduces the
// curvature of the underlying smooth function), so to reproduce the reference's Hessian the perturbed evaluations must
// extend the cell of the unperturbed point linearly.  mode 0: normal; 1: normal + record the cell used; 2: use the
// recorded cell (t may leave [0, 1]).  Cell codes: >= 0 the cell; -1 / -2 the left / right out-of-table branch.
constexpr int kCellStride = 1 + TSFF_MAX_IONS;   // per (omega, angle) point: T-table cell, then one Z' cell per ion

TSFF_HD void zprime_lerp_cell(const ZTab& z, double x, double& zr, double& zi, double& dzr, double& dzi, int mode, int* cell) {
  if (mode != 2) {
    zprime_lerp(z, x, zr, zi, dzr, dzi);
    if (mode == 1) {
      int c = -1;
      if (!(x < z.x0 || x > z.xlast)) {
        c = (int)((x - z.x0) * fast_rcp(z.h));
        c = c > z.n - 2 ? z.n - 2 : (c < 0 ? 0 : c);
      }
      *cell = c;
    }
    return;
  }
  const int i = *cell;
  if (i < 0) {
    double x2 = x * x;
    zr = fast_rcp(x2);
    dzr = -2.0 * zr * fast_rcp(x);
    zi = 0.0;
    dzi = 0.0;
    return;
  }
  const double ih = fast_rcp(z.h);
  const double t = (x - z.x0) * ih - (double)i;
  double r0 = z.zr[i], r1 = z.zr[i + 1], i0 = z.zi[i], i1 = z.zi[i + 1];
  zr = r0 + t * (r1 - r0);
  zi = i0 + t * (i1 - i0);
  dzr = (r1 - r0) * ih;
  dzi = (i1 - i0) * ih;
}

// edge-clamped linear interpolation on a uniform grid (jnp.interp default; form_factor.py:270,376-377)
// returns value; idx/t/slope for the adjoint (slope = 0 and weights collapse on the edge when clamped)
template <typename T>
TSFF_HD double lerp_uniform(const T* f, int n, double x0, double h, double x, int& i, double& t, double& slope) {
  const double ih = fast_rcp(h);
  double u = (x - x0) * ih;
  if (!(u > 0.0)) {  // left clamp (also NaN)
    i = 0; t = 0.0; slope = 0.0;
    return (double)f[0];
  }
  if (u >= (double)(n - 1)) {
    i = n - 2; t = 1.0; slope = 0.0;
    return (double)f[n - 1];
  }
  i = (int)u;
  t = u - (double)i;
  double a = (double)f[i], b = (double)f[i + 1];
  slope = (b - a) * ih;
  return a + t * (b - a);
}

template <typename T>
TSFF_HD double lerp_uniform_cell(const T* f, int n, double x0, double h, double x, int& i, double& t, double& slope, int mode,
                                 int* cell) {
  if (mode != 2) {
    const double v = lerp_uniform(f, n, x0, h, x, i, t, slope);
    if (mode == 1) {
      const double u = (x - x0) * fast_rcp(h);
      *cell = !(u > 0.0) ? -1 : (u >= (double)(n - 1) ? -2 : i);
    }
    return v;
  }
  const int c = *cell;
  if (c == -1) { i = 0; t = 0.0; slope = 0.0; return (double)f[0]; }
  if (c == -2) { i = n - 2; t = 1.0; slope = 0.0; return (double)f[n - 1]; }
  const double ih = fast_rcp(h);
  i = c;
  t = (x - x0) * ih - (double)c;
  const double a = (double)f[c], b = (double)f[c + 1];
  slope = (b - a) * ih;
  return a + t * (b - a);
}

// ------------------------------------------------------------------------------------------------
// point kinematics (form_factor.py:215-228, 253)
// ------------------------------------------------------------------------------------------------
struct Kin {
  double ks, k2, k, omgdop, w, xie, ikl2;
};

TSFF_HD void kin_forward(const LG& L, double omgs, double cth, Kin& q) {
  q.ks = fast_sqrt(omgs * omgs - L.omgpe2) * (1.0 / kC);
  q.k2 = q.ks * q.ks + L.kL * L.kL - 2.0 * q.ks * L.kL * cth;
  q.k = fast_sqrt(q.k2);
  q.omgdop = omgs - L.omgL - q.k * L.Va6;
  q.w = q.omgdop * fast_rcp(q.k);
  const double ivTe = fast_rcp(L.vTe);
  q.xie = (q.w - L.ud6) * ivTe;
  q.ikl2 = L.omgpe2 * ivTe * ivTe * fast_rcp(q.k2);
}

// ion susceptibility (form_factor.py:231-249) -> chiI, plus the ion-feature sum  sum_i ioncf_i exp(-xii^2)
struct IonOut {
  double chiIr, chiIi, sion;
};

TSFF_HD void ion_forward(const LG& L, int nI, const ZTab& zt, const Kin& q, IonOut& o, int cell_mode = 0, int* cells = nullptr) {
  o.chiIr = o.chiIi = o.sion = 0.0;
#pragma unroll
  for (int i = 0; i < TSFF_MAX_IONS; i++) {   // fixed trip count + predicate: the per-ion arrays stay in registers
    if (i >= nI) break;
    double xii = L.inv_s2vTi[i] * q.w;
    double ikldi2 = fast_rcp(L.c_kldi[i] * L.c_kldi[i] * q.k2);
    double zr, zi, dzr, dzi;
    zprime_lerp_cell(zt, xii, zr, zi, dzr, dzi, cell_mode, cells + i);
    o.chiIr += -0.5 * ikldi2 * zr;
    o.chiIi += -0.5 * ikldi2 * zi;
    o.sion += L.ioncf[i] * fast_exp_neg(-xii * xii);
  }
}

// spectral density assembly (form_factor.py:273-296): returns PsLam
struct Asm {
  double er, ei, eps2, ce2, a1, Sion, Sele, dop, cP, P;
};

TSFF_HD double assemble_forward(const LG& L, const Kin& q, const IonOut& io, double chiEr, double chiEi, double fphi,
                                double omgs, Asm& s) {
  s.er = 1.0 + chiEr + io.chiIr;
  s.ei = chiEi + io.chiIi;
  s.eps2 = s.er * s.er + s.ei * s.ei;
  s.ce2 = chiEr * chiEr + chiEi * chiEi;
  s.a1 = (1.0 + io.chiIr) * (1.0 + io.chiIr) + io.chiIi * io.chiIi;
  const double ike = fast_rcp(q.k * s.eps2);
  s.Sion = io.sion * s.ce2 * kInvSqrt2Pi * ike;
  s.Sele = s.a1 * fphi * ike * fast_rcp(L.vTe);
  s.dop = 1.0 + 2.0 * q.omgdop * fast_rcp(L.omgL);
  s.cP = kRe * kRe * omgs * omgs / (2.0 * kPi * kC);  // re^2 * 2 pi C / lams^2, lams = 2 pi C / omgs
  s.P = (s.Sion + s.Sele) * s.dop * L.ne_g * s.cP;
  return s.P;
}

// adjoints produced by the assembly + ion + kinematics reverse sweep for one point
struct PointBar {
  double chiEr, chiEi, fphi;  // cotangents handed to the chi_e provider
};

// Reverse of assemble_forward and ion_forward down to (chiE, fphi, kinematic scalars).  `Pbar` is the
// cotangent of PsLam.  Kinematic cotangents are accumulated in kb (k, k2, omgdop, w) and LG cotangents in Lb.
struct KinBar {
  double k, k2, omgdop, w, xie, ikl2;
};

TSFF_HD void assemble_backward(const LG& L, int nI, const ZTab& zt, const Kin& q, const IonOut& io, double chiEr,
                               double chiEi, double fphi, const Asm& s, double Pbar, PointBar& pb, KinBar& kb,
                               LG& Lb, int cell_mode = 0, int* cells = nullptr) {
  double Ssum = s.Sion + s.Sele;
  double Ssum_bar = Pbar * s.dop * L.ne_g * s.cP;
  double dop_bar = Pbar * Ssum * L.ne_g * s.cP;
  Lb.ne_g += Pbar * Ssum * s.dop * s.cP;
  const double iomgL = fast_rcp(L.omgL), ivTe = fast_rcp(L.vTe), ik = fast_rcp(q.k), ik2 = fast_rcp(q.k2);
  kb.omgdop += dop_bar * 2.0 * iomgL;
  Lb.omgL += -dop_bar * 2.0 * q.omgdop * iomgL * iomgL;
  // Sion = sion ce2 c / (k eps2)
  double inv_keps = fast_rcp(q.k * s.eps2);
  double sion_bar = Ssum_bar * s.ce2 * kInvSqrt2Pi * inv_keps;
  double ce2_bar = Ssum_bar * io.sion * kInvSqrt2Pi * inv_keps;
  double eps2_bar = -Ssum_bar * Ssum * (inv_keps * q.k);
  kb.k += -Ssum_bar * Ssum * ik;
  // Sele = a1 fphi / (k vTe eps2)
  double a1_bar = Ssum_bar * fphi * inv_keps * ivTe;
  pb.fphi = Ssum_bar * s.a1 * inv_keps * ivTe;
  Lb.vTe += -Ssum_bar * s.Sele * ivTe;
  double er_bar = 2.0 * s.er * eps2_bar, ei_bar = 2.0 * s.ei * eps2_bar;
  pb.chiEr = er_bar + 2.0 * chiEr * ce2_bar;
  pb.chiEi = ei_bar + 2.0 * chiEi * ce2_bar;
  double chiIr_bar = er_bar + 2.0 * (1.0 + io.chiIr) * a1_bar;
  double chiIi_bar = ei_bar + 2.0 * io.chiIi * a1_bar;
#pragma unroll
  for (int i = 0; i < TSFF_MAX_IONS; i++) {
    if (i >= nI) break;
    double xii = L.inv_s2vTi[i] * q.w;
    double ikldi2 = fast_rcp(L.c_kldi[i] * L.c_kldi[i] * q.k2);
    double zr, zi, dzr, dzi;
    zprime_lerp_cell(zt, xii, zr, zi, dzr, dzi, cell_mode == 2 ? 2 : 0, cells + i);
    double E = fast_exp_neg(-xii * xii);
    Lb.ioncf[i] += sion_bar * E;
    double xii_bar = sion_bar * L.ioncf[i] * E * (-2.0 * xii);
    double ikldi2_bar = -0.5 * (zr * chiIr_bar + zi * chiIi_bar);
    xii_bar += -0.5 * ikldi2 * (dzr * chiIr_bar + dzi * chiIi_bar);
    Lb.c_kldi[i] += ikldi2_bar * (-2.0 * ikldi2 * fast_rcp(L.c_kldi[i]));
    kb.k2 += -ikldi2_bar * ikldi2 * ik2;
    Lb.inv_s2vTi[i] += xii_bar * q.w;
    kb.w += xii_bar * L.inv_s2vTi[i];
  }
}

// reverse of kin_forward: consumes kb (incl. xie, ikl2 cotangents), accumulates LG cotangents
TSFF_HD void kin_backward(const LG& L, double omgs, double cth, const Kin& q, KinBar kb, LG& Lb) {
  const double ivTe = fast_rcp(L.vTe), ik = fast_rcp(q.k), ik2 = ik * ik;
  // ikl2 = omgpe2 / (vTe^2 k2)
  Lb.omgpe2 += kb.ikl2 * q.ikl2 * fast_rcp(L.omgpe2);
  Lb.vTe += -2.0 * kb.ikl2 * q.ikl2 * ivTe;
  kb.k2 += -kb.ikl2 * q.ikl2 * ik2;
  // xie = (w - ud6)/vTe
  kb.w += kb.xie * ivTe;
  Lb.ud6 += -kb.xie * ivTe;
  Lb.vTe += -kb.xie * q.xie * ivTe;
  // w = omgdop / k
  kb.omgdop += kb.w * ik;
  kb.k += -kb.w * q.w * ik;
  // omgdop = omgs - omgL - k Va6
  Lb.omgL += -kb.omgdop;
  kb.k += -kb.omgdop * L.Va6;
  Lb.Va6 += -kb.omgdop * q.k;
  // k = sqrt(k2)
  kb.k2 += kb.k * (0.5 * ik);
  // k2 = ks^2 + kL^2 - 2 ks kL cth
  double ks_bar = kb.k2 * (2.0 * q.ks - 2.0 * L.kL * cth);
  Lb.kL += kb.k2 * (2.0 * L.kL - 2.0 * q.ks * cth);
  // ks = sqrt(omgs^2 - omgpe2)/C
  Lb.omgpe2 += -ks_bar * fast_rcp(2.0 * kC * kC * q.ks);
}

// ------------------------------------------------------------------------------------------------
// The same point chain with the loop-invariant reciprocals taken out ("_x" variants, used by the 1V kernels).
// A (lineout, gradient point) has ~8 reciprocals that do not depend on (omega, angle): 1/vTe, 1/omgL, 1/omgpe2, 1/c_kldi,
// 1/c_kldi^2, ...; 1/k and 1/ks come for free with the square roots (Goldschmidt carries 1/(2 sqrt x)); 1/(k eps^2) is
// formed once and handed from the forward assembly to its reverse, as are the ions' Z' cell values and exp(-xii^2).
// Per point this leaves ONE FP64 reciprocal (1/eps^2) and two square roots, against ~26 MUFU seeds + Newton steps before.
// Results differ from the plain functions by rounding only (tests/hostsim checks both against each other).
// ------------------------------------------------------------------------------------------------
struct LGX {
  double ivTe, iomgL, iomgpe2, opv, kL2, zih;          // opv = omgpe2 / vTe^2 ; zih = 1 / (Z' table spacing)
  double ickl[TSFF_MAX_IONS], ickl2[TSFF_MAX_IONS];    // 1 / c_kldi, 1 / c_kldi^2
};
constexpr int kLGXDoubles = 6 + 2 * TSFF_MAX_IONS;
// field f of the LGX of L, in declaration order (so that kLGXDoubles threads can fill a shared LGX, one division each)
TSFF_HD double lgx_field(const LG& L, int nI, double zh, int f) {
  switch (f) {
    case 0: return 1.0 / L.vTe;
    case 1: return 1.0 / L.omgL;
    case 2: return 1.0 / L.omgpe2;
    case 3: return L.omgpe2 / (L.vTe * L.vTe);
    case 4: return L.kL * L.kL;
    case 5: return 1.0 / zh;
    default: break;
  }
  const int i = (f - 6) % TSFF_MAX_IONS;
  if (i >= nI) return 0.0;
  return f - 6 < TSFF_MAX_IONS ? 1.0 / L.c_kldi[i] : 1.0 / (L.c_kldi[i] * L.c_kldi[i]);
}
TSFF_HD void lgx_make(const LG& L, int nI, double zh, LGX& X) {
  X.ivTe = lgx_field(L, nI, zh, 0);
  X.iomgL = lgx_field(L, nI, zh, 1);
  X.iomgpe2 = lgx_field(L, nI, zh, 2);
  X.opv = lgx_field(L, nI, zh, 3);
  X.kL2 = lgx_field(L, nI, zh, 4);
  X.zih = lgx_field(L, nI, zh, 5);
  for (int i = 0; i < TSFF_MAX_IONS; i++) {
    X.ickl[i] = lgx_field(L, nI, zh, 6 + i);
    X.ickl2[i] = lgx_field(L, nI, zh, 6 + TSFF_MAX_IONS + i);
  }
}

// sqrt(x) and 1/sqrt(x) together (x a positive normal number)
TSFF_HD void fast_sqrt_rsqrt(double x, double& s, double& r) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double h = 0.5 * y, g = x * y;
  double e = fma(-g, h, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-g, h, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  s = fma(fma(-g, g, x), h, g);
  r = h + h;
#else
  s = sqrt(x);
  r = 1.0 / s;
#endif
}

struct KinX {
  double ks, k2, k, omgdop, w, xie, ikl2;   // as Kin
  double ik, ik2, iks;                      // 1/k, 1/k^2, 1/ks
};

TSFF_HD void kin_forward_x(const LG& L, const LGX& X, double omgs, double cth, KinX& q) {
  double rs;
  fast_sqrt_rsqrt(omgs * omgs - L.omgpe2, q.ks, rs);
  q.ks *= (1.0 / kC);
  q.iks = rs * kC;
  q.k2 = q.ks * q.ks + X.kL2 - 2.0 * q.ks * L.kL * cth;
  fast_sqrt_rsqrt(q.k2, q.k, q.ik);
  q.ik2 = q.ik * q.ik;
  q.omgdop = omgs - L.omgL - q.k * L.Va6;
  q.w = q.omgdop * q.ik;
  q.xie = (q.w - L.ud6) * X.ivTe;
  q.ikl2 = X.opv * q.ik2;
}

// Z' lookup with the inverse spacing passed in; mode / cell as zprime_lerp_cell
TSFF_HD void zprime_lerp_x(const ZTab& z, double ih, double x, double& zr, double& zi, double& dzr, double& dzi, int mode,
                           int* cell) {
  int i;
  if (mode == 2) {
    i = *cell;
  } else {
    i = -1;
    if (!(x < z.x0 || x > z.xlast)) {
      i = (int)((x - z.x0) * ih);
      i = i > z.n - 2 ? z.n - 2 : (i < 0 ? 0 : i);
    }
    if (mode == 1) *cell = i;
  }
  if (i < 0) {
    const double ix = fast_rcp(x);
    zr = ix * ix;
    dzr = -2.0 * zr * ix;
    zi = 0.0;
    dzi = 0.0;
    return;
  }
  const double t = (x - z.x0) * ih - (double)i;
  const double r0 = z.zr[i], r1 = z.zr[i + 1], i0 = z.zi[i], i1 = z.zi[i + 1];
  zr = r0 + t * (r1 - r0);
  zi = i0 + t * (i1 - i0);
  dzr = (r1 - r0) * ih;
  dzi = (i1 - i0) * ih;
}

// ion susceptibility; keeps what its reverse needs (per ion: xii, 1/(k lambda_Di)^2, the Z' values / slopes, exp(-xii^2))
struct IonX {
  double chiIr, chiIi, sion;
  double xii[TSFF_MAX_IONS], ikldi2[TSFF_MAX_IONS], zr[TSFF_MAX_IONS], zi[TSFF_MAX_IONS], dzr[TSFF_MAX_IONS],
      dzi[TSFF_MAX_IONS], E[TSFF_MAX_IONS];
};

// NI > 0: the ion count at compile time (the per-ion arrays of the unused slots then cost no registers); NI == 0: nI at run time
template <int NI>
TSFF_HD void ion_forward_x(const LG& L, const LGX& X, int nI, const ZTab& zt, const KinX& q, IonX& o, int cell_mode = 0,
                           int* cells = nullptr) {
  o.chiIr = o.chiIi = o.sion = 0.0;
#pragma unroll
  for (int i = 0; i < TSFF_MAX_IONS; i++) {
    if (NI > 0 ? i >= NI : i >= nI) break;
    const double xii = L.inv_s2vTi[i] * q.w;
    const double ikldi2 = X.ickl2[i] * q.ik2;
    zprime_lerp_x(zt, X.zih, xii, o.zr[i], o.zi[i], o.dzr[i], o.dzi[i], cell_mode, cells + i);
    o.xii[i] = xii;
    o.ikldi2[i] = ikldi2;
    o.E[i] = fast_exp_neg(-xii * xii);
    o.chiIr += -0.5 * ikldi2 * o.zr[i];
    o.chiIi += -0.5 * ikldi2 * o.zi[i];
    o.sion += L.ioncf[i] * o.E[i];
  }
}

constexpr double kCP = kRe * kRe / (2.0 * kPi * kC);   // re^2 * 2 pi C / lams^2 = kCP * omgs^2

struct AsmX {
  double er, ei, eps2, ce2, a1, Sion, Sele, dop, cP, P, ike;   // ike = 1 / (k eps2)
};

TSFF_HD double assemble_forward_x(const LG& L, const LGX& X, const KinX& q, const IonX& io, double chiEr, double chiEi,
                                  double fphi, double omgs, AsmX& s) {
  const double ci = 1.0 + io.chiIr;
  s.er = ci + chiEr;
  s.ei = chiEi + io.chiIi;
  s.eps2 = s.er * s.er + s.ei * s.ei;
  s.ce2 = chiEr * chiEr + chiEi * chiEi;
  s.a1 = ci * ci + io.chiIi * io.chiIi;
  s.ike = q.ik * fast_rcp(s.eps2);
  s.Sion = io.sion * s.ce2 * kInvSqrt2Pi * s.ike;
  s.Sele = s.a1 * fphi * s.ike * X.ivTe;
  s.dop = 1.0 + 2.0 * q.omgdop * X.iomgL;
  s.cP = omgs * omgs * kCP;
  s.P = (s.Sion + s.Sele) * s.dop * L.ne_g * s.cP;
  return s.P;
}

template <int NI>
TSFF_HD void assemble_backward_x(const LG& L, const LGX& X, int nI, const KinX& q, const IonX& io, double chiEr, double chiEi,
                                 double fphi, const AsmX& s, double Pbar, PointBar& pb, KinBar& kb, LG& Lb) {
  const double Ssum = s.Sion + s.Sele;
  const double pc = Pbar * s.cP;
  const double Ssum_bar = pc * s.dop * L.ne_g;
  const double dop_bar = pc * Ssum * L.ne_g;
  Lb.ne_g += pc * Ssum * s.dop;
  const double t2 = dop_bar * 2.0 * X.iomgL;
  kb.omgdop += t2;
  Lb.omgL += -t2 * q.omgdop * X.iomgL;
  const double sk = Ssum_bar * s.ike;                 // Ssum_bar / (k eps2)
  const double sion_bar = sk * s.ce2 * kInvSqrt2Pi;
  const double ce2_bar = sk * io.sion * kInvSqrt2Pi;
  const double eps2_bar = -sk * Ssum * q.k;
  kb.k += -Ssum_bar * Ssum * q.ik;
  const double skv = sk * X.ivTe;
  const double a1_bar = skv * fphi;
  pb.fphi = skv * s.a1;
  Lb.vTe += -Ssum_bar * s.Sele * X.ivTe;
  const double er_bar = 2.0 * s.er * eps2_bar, ei_bar = 2.0 * s.ei * eps2_bar;
  pb.chiEr = er_bar + 2.0 * chiEr * ce2_bar;
  pb.chiEi = ei_bar + 2.0 * chiEi * ce2_bar;
  const double chiIr_bar = er_bar + 2.0 * (1.0 + io.chiIr) * a1_bar;
  const double chiIi_bar = ei_bar + 2.0 * io.chiIi * a1_bar;
#pragma unroll
  for (int i = 0; i < TSFF_MAX_IONS; i++) {
    if (NI > 0 ? i >= NI : i >= nI) break;
    const double xii = io.xii[i], ikldi2 = io.ikldi2[i], E = io.E[i];
    Lb.ioncf[i] += sion_bar * E;
    double xii_bar = sion_bar * L.ioncf[i] * E * (-2.0 * xii);
    const double ikldi2_bar = -0.5 * (io.zr[i] * chiIr_bar + io.zi[i] * chiIi_bar);
    xii_bar += -0.5 * ikldi2 * (io.dzr[i] * chiIr_bar + io.dzi[i] * chiIi_bar);
    const double ib = ikldi2_bar * ikldi2;
    Lb.c_kldi[i] += -2.0 * ib * X.ickl[i];
    kb.k2 += -ib * q.ik2;
    Lb.inv_s2vTi[i] += xii_bar * q.w;
    kb.w += xii_bar * L.inv_s2vTi[i];
  }
}

TSFF_HD void kin_backward_x(const LG& L, const LGX& X, double cth, const KinX& q, KinBar kb, LG& Lb) {
  const double ib = kb.ikl2 * q.ikl2;                 // ikl2 = omgpe2 / (vTe^2 k2)
  Lb.omgpe2 += ib * X.iomgpe2;
  kb.k2 += -ib * q.ik2;
  const double xv = kb.xie * X.ivTe;                  // xie = (w - ud6) / vTe
  kb.w += xv;
  Lb.ud6 += -xv;
  Lb.vTe += -(2.0 * ib + kb.xie * q.xie) * X.ivTe;
  const double wk = kb.w * q.ik;                      // w = omgdop / k
  kb.omgdop += wk;
  kb.k += -wk * q.w;
  Lb.omgL += -kb.omgdop;                              // omgdop = omgs - omgL - k Va6
  kb.k += -kb.omgdop * L.Va6;
  Lb.Va6 += -kb.omgdop * q.k;
  kb.k2 += kb.k * (0.5 * q.ik);                       // k = sqrt(k2)
  const double ks_bar = kb.k2 * (2.0 * q.ks - 2.0 * L.kL * cth);   // k2 = ks^2 + kL^2 - 2 ks kL cth
  Lb.kL += kb.k2 * (2.0 * L.kL - 2.0 * q.ks * cth);
  Lb.omgpe2 += -ks_bar * (0.5 / (kC * kC)) * q.iks;  // ks = sqrt(omgs^2 - omgpe2) / C
}

// ------------------------------------------------------------------------------------------------
// interpax cubic Hermite on a uniform grid with pre-computed node slopes (form_factor.py:256,263;
// SURVEY.md A-note 1).  lnf[V], slope[V]; returns H(x) (or `fill` outside [x0, x_last]) and the pieces
// for the adjoint.
// ------------------------------------------------------------------------------------------------
struct Herm {
  int i;         // right node of the cell (1..V-1); 0 when outside
  double t;      // (x - x[i-1]) / h
  double dHdx;   // derivative wrt x (0 outside)
  bool inside;
};

TSFF_HD double hermite_uniform(const double* lnf, const double* slope, int V, double x0, double h, double x,
                               double fill, Herm& o) {
  double xlast = x0 + (double)(V - 1) * h;
  if (x < x0 || x > xlast || !(x == x)) {
    o.i = 0; o.t = 0.0; o.dHdx = 0.0; o.inside = false;
    return fill;
  }
  // searchsorted(x, xq, 'right') clipped to [1, V-1]
  int i = (int)floor((x - x0) / h) + 1;
  if (i < 1) i = 1;
  if (i > V - 1) i = V - 1;
  double t = (x - (x0 + (double)(i - 1) * h)) / h;
  double f0 = lnf[i - 1], f1 = lnf[i], m0 = slope[i - 1] * h, m1 = slope[i] * h;
  double c2 = -3.0 * f0 + 3.0 * f1 - 2.0 * m0 - m1;
  double c3 = 2.0 * f0 - 2.0 * f1 + m0 + m1;
  o.i = i; o.t = t; o.inside = true;
  o.dHdx = (m0 + t * (2.0 * c2 + 3.0 * c3 * t)) / h;
  return f0 + t * (m0 + t * (c2 + c3 * t));
}

// the same with the inverse spacing passed in (no FP64 divisions: the table kernels evaluate this once per (omega, angle))
TSFF_HD double hermite_uniform_ih(const double* lnf, const double* slope, int V, double x0, double h, double ih, double x,
                                  double fill, Herm& o) {
  const double xlast = x0 + (double)(V - 1) * h;
  if (x < x0 || x > xlast || !(x == x)) {
    o.i = 0; o.t = 0.0; o.dHdx = 0.0; o.inside = false;
    return fill;
  }
  const double u = (x - x0) * ih;
  int i = (int)u + 1;                       // u >= 0 here
  if (i > V - 1) i = V - 1;
  const double t = u - (double)(i - 1);
  const double f0 = lnf[i - 1], f1 = lnf[i], m0 = slope[i - 1] * h, m1 = slope[i] * h;
  const double d = f1 - f0;
  const double c2 = 3.0 * d - 2.0 * m0 - m1;
  const double c3 = m0 + m1 - 2.0 * d;
  o.i = i; o.t = t; o.inside = true;
  o.dHdx = (m0 + t * (2.0 * c2 + 3.0 * c3 * t)) * ih;
  return f0 + t * (m0 + t * (c2 + c3 * t));
}

// lerp_uniform with the inverse spacing passed in
template <typename T>
TSFF_HD double lerp_uniform_ih(const T* f, int n, double x0, double ih, double x, int& i, double& t, double& slope) {
  const double u = (x - x0) * ih;
  if (!(u > 0.0)) {  // left clamp (also NaN)
    i = 0; t = 0.0; slope = 0.0;
    return (double)f[0];
  }
  if (u >= (double)(n - 1)) {
    i = n - 2; t = 1.0; slope = 0.0;
    return (double)f[n - 1];
  }
  i = (int)u;
  t = u - (double)i;
  const double a = (double)f[i], b = (double)f[i + 1];
  slope = (b - a) * ih;
  return a + t * (b - a);
}

// ---- branch-free forms (the table adjoint keeps one point chain per thread in flight, so its speed is the length of that
// chain; straight-line code lets the compiler overlap the look-ahead point with the reverse sweep of the current one) ----
// exp(y) for y in [-700, 700] (clamped above), 0 below, NaN -> NaN; even / odd split of the degree-12 polynomial
TSFF_HD double fast_exp_bf(double y) {
#if defined(__CUDA_ARCH__)
  const double yc = fmin(fmax(y, -700.0), 700.0);
  const double n = rint(yc * kLn2Split[0]);
  double f = fma(-n, kLn2Split[1], yc);
  f = fma(-n, kLn2Split[2], f);
  const double f2 = f * f;
  double pe = kExpC[0];                      // even powers: 1/12!, 1/10!, ..., 1
#pragma unroll
  for (int i = 2; i < 13; i += 2) pe = fma(pe, f2, kExpC[i]);
  double po = kExpC[1];                      // odd powers: 1/11!, 1/9!, ..., 1
#pragma unroll
  for (int i = 3; i < 13; i += 2) po = fma(po, f2, kExpC[i]);
  const double p = fma(po, f, pe);
  const double r = __hiloint2double(__double2hiint(p) + ((int)n << 20), __double2loint(p));
  return y > -700.0 ? r : (y == y ? 0.0 : y);
#else
  return y > -700.0 ? exp(fmin(y, 700.0)) : (y == y ? 0.0 : y);
#endif
}

TSFF_HD double hermite_uniform_bf(const double* lnf, const double* slope, int V, double x0, double h, double ih, double x,
                                  double fill, Herm& o) {
  const double xlast = x0 + (double)(V - 1) * h;
  const bool inside = (x >= x0) && (x <= xlast);
  const double u = fmin(fmax((x - x0) * ih, 0.0), (double)(V - 1));
  int i = (int)u + 1;
  i = i > V - 1 ? V - 1 : i;
  const double t = u - (double)(i - 1);
  const double f0 = lnf[i - 1], f1 = lnf[i], m0 = slope[i - 1] * h, m1 = slope[i] * h;
  const double d = f1 - f0;
  const double c2 = 3.0 * d - 2.0 * m0 - m1;
  const double c3 = m0 + m1 - 2.0 * d;
  o.i = inside ? i : 0;
  o.t = inside ? t : 0.0;
  o.inside = inside;
  o.dHdx = inside ? (m0 + t * (2.0 * c2 + 3.0 * c3 * t)) * ih : 0.0;
  return inside ? f0 + t * (m0 + t * (c2 + c3 * t)) : fill;
}

template <typename T>
TSFF_HD double lerp_uniform_bf(const T* f, int n, double x0, double ih, double x, int& i, double& t, double& slope) {
  const double u = (x - x0) * ih;
  const bool edge = !(u > 0.0) || (u >= (double)(n - 1));
  const double uc = fmin(fmax(u, 0.0), (double)(n - 1));
  int ii = (int)uc;
  ii = ii > n - 2 ? n - 2 : ii;
  const double tt = uc - (double)ii;
  const double a = (double)f[ii], b = (double)f[ii + 1];
  i = ii;
  t = tt;
  slope = edge ? 0.0 : (b - a) * ih;
  return a + tt * (b - a);
}

// Z' lookup from an interleaved table zz[i] = (Zr_i, Zi_i) (shared memory), asymptotic form outside; no branches
struct ZZ { double r, i; };
TSFF_HD void zprime_lerp_bf(const ZZ* zz, int n, double x0, double xlast, double ih, double x, double& zr, double& zi, double& dzr,
                            double& dzi) {
  const bool out = (x < x0) || (x > xlast);
  const double u = fmin(fmax((x - x0) * ih, 0.0), (double)(n - 2));
  const int i = (int)u;
  const double t = (x - x0) * ih - (double)i;
  const ZZ a = zz[i], b = zz[i + 1];
  const double ix = fast_rcp(out ? x : 1.0);
  const double ar = ix * ix;
  zr = out ? ar : a.r + t * (b.r - a.r);
  zi = out ? 0.0 : a.i + t * (b.i - a.i);
  dzr = out ? -2.0 * ar * ix : (b.r - a.r) * ih;
  dzi = out ? 0.0 : (b.i - a.i) * ih;
}

// the same from the two separate rows (global memory)
struct ZRows { const double* zr; const double* zi; };
TSFF_HD void zprime_lerp_bf(ZRows z, int n, double x0, double xlast, double ih, double x, double& zr, double& zi, double& dzr,
                            double& dzi) {
  const bool out = (x < x0) || (x > xlast);
  const double u = fmin(fmax((x - x0) * ih, 0.0), (double)(n - 2));
  const int i = (int)u;
  const double t = (x - x0) * ih - (double)i;
  const double ar0 = z.zr[i], ar1 = z.zr[i + 1], ai0 = z.zi[i], ai1 = z.zi[i + 1];
  const double ix = fast_rcp(out ? x : 1.0);
  const double ar = ix * ix;
  zr = out ? ar : ar0 + t * (ar1 - ar0);
  zi = out ? 0.0 : ai0 + t * (ai1 - ai0);
  dzr = out ? -2.0 * ar * ix : (ar1 - ar0) * ih;
  dzi = out ? 0.0 : (ai1 - ai0) * ih;
}

template <int NI, typename ZT>
TSFF_HD void ion_forward_bf(const LG& L, const LGX& X, int nI, ZT zz, const ZTab& zt, const KinX& q, IonX& o) {
  o.chiIr = o.chiIi = o.sion = 0.0;
#pragma unroll
  for (int i = 0; i < TSFF_MAX_IONS; i++) {
    if (NI > 0 ? i >= NI : i >= nI) break;
    const double xii = L.inv_s2vTi[i] * q.w;
    const double ikldi2 = X.ickl2[i] * q.ik2;
    zprime_lerp_bf(zz, zt.n, zt.x0, zt.xlast, X.zih, xii, o.zr[i], o.zi[i], o.dzr[i], o.dzi[i]);
    o.xii[i] = xii;
    o.ikldi2[i] = ikldi2;
    o.E[i] = fast_exp_bf(-xii * xii);
    o.chiIr += -0.5 * ikldi2 * o.zr[i];
    o.chiIi += -0.5 * ikldi2 * o.zi[i];
    o.sion += L.ioncf[i] * o.E[i];
  }
}

// adjoint weights of H wrt (lnf[i-1], lnf[i], slope[i-1], slope[i])
TSFF_HD void hermite_weights(double t, double h, double& wf0, double& wf1, double& wm0, double& wm1) {
  double t2 = t * t, t3 = t2 * t;
  wf0 = 1.0 - 3.0 * t2 + 2.0 * t3;
  wf1 = 3.0 * t2 - 2.0 * t3;
  wm0 = (t - 2.0 * t2 + t3) * h;
  wm1 = (-t2 + t3) * h;
}

}  // namespace tsff
