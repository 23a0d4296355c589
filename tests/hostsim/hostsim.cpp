// Host-side self-consistency checks of tsadar_b200/csrc/tsff_math.cuh and tsff_pv.cuh (the exact headers the CUDA
// kernels compile), built with g++ by tests/test_hostsim.py.  TEST ONLY -- the product has no CPU path.
//  1. lg_backward      vs central differences of lg_forward
//  2. assemble/ion/kin backward vs central differences of the forward point chain
//  3. pv_accumulate_f64 + pv_finish (FP64 validation path) vs the literal ratintn formula (ratintn.py:4-52)
//  4. hermite_uniform derivative / weights vs differences
//  5. the block-multipole evaluation (tsff_tree.cuh: far expansion + near-window series + exact near nodes) vs ratintn
#include <cstdio>
#include <cmath>
#include <vector>
#include <complex>
#include <algorithm>
#include "../../tsadar_b200/csrc/tsff_math.cuh"
#include "../../tsadar_b200/csrc/tsff_pv.cuh"
#include "../../tsadar_b200/csrc/tsff_tree.cuh"
using namespace tsff;

static double relerr(double a, double b) { return fabs(a - b) / std::max(1e-300, std::max(fabs(a), fabs(b))); }

// literal ratintn (ratintn.py:4-52) in double
static double ratintn_literal(const std::vector<double>& f, const std::vector<double>& z, double xi) {
  int N = (int)f.size();
  double out = 0.0;
  for (int i = 0; i + 2 < N; i++) {
    double g0 = z[i] - xi, g1 = z[i + 1] - xi;
    double fdif = f[i + 1] - f[i], gdif = g1 - g0, fav = 0.5 * (f[i + 1] + f[i]), gav = 0.5 * (g1 + g0);
    double tmp = fav * gdif - gav * fdif;
    double rfn = fdif / gdif + tmp * std::log(std::complex<double>((gav + 0.5 * gdif) / (gav - 0.5 * gdif), 0.0)).real() / (gdif * gdif);
    out += rfn * (z[i + 1] - z[i]);
  }
  return out;
}

struct Tables { std::vector<double> zr, zi; ZTab zt; };

static double point_P(const double* params, int nI, double lam_shift, double omgs, double cth, const ZTab& zt, double chiEr_in,
                      double chiEi_in, double fphi, bool use_xie_dep) {
  LG L; lg_zero(L); lg_forward(params, nI, 0, 1, lam_shift, L);
  Kin q; kin_forward(L, omgs, cth, q);
  IonOut io; ion_forward(L, nI, zt, q, io);
  // make chiE depend on kinematics the way the kernels do: chiEr = -ikl2 * I(xie), chiEi = pi ikl2 dfe(xie), fphi(xie)
  double I = use_xie_dep ? sin(q.xie) * 0.7 : chiEr_in;
  double dfe = use_xie_dep ? cos(0.5 * q.xie) * 0.3 : chiEi_in;
  double fp = use_xie_dep ? exp(-0.5 * q.xie * q.xie) * 0.4 : fphi;
  Asm s;
  return assemble_forward(L, q, io, -q.ikl2 * I, kPi * q.ikl2 * dfe, fp, omgs, s);
}

int main() {
  int fails = 0;
  // Z' like table (smooth stand-in; the real table is checked on the GPU against the oracle)
  Tables T; T.zr.resize(1640); T.zi.resize(1640);
  for (int i = 0; i < 1640; i++) { double x = -8.2 + 0.01 * i; T.zr[i] = -2.0 * (1 - 2 * x * x) * exp(-x * x) + 0.1 / (1 + x * x); T.zi[i] = -3.5 * x * exp(-x * x); }
  T.zt = {T.zr.data(), T.zi.data(), 1640, -8.2, 0.01, -8.2 + 1639 * 0.01};
  const int nI = 2, NP = 10 + 4 * nI;
  double params[NP] = {0.8, 0.35, 526.5, 1.5, -0.7, 3.0, 2.0, 1, 1, 1, /*ion1*/ 40.0, 8.0, 0.2, 0.6, /*ion2*/ 1.0, 1.0, 0.3, 0.4};
  const double lam_shift = 0.3;
  // ---- 1. lg_backward
  {
    const int G = 3;
    for (int g = 0; g < G; g++) {
      // scalar functional: random linear combination of LG outputs
      auto fun = [&](const double* p) {
        LG L; lg_zero(L); lg_forward(p, nI, g, G, lam_shift, L);
        double s = 1e-20 * L.ne_g + 1e-15 * L.omgL + 1e-30 * L.omgpe2 + 1e-5 * L.kL + 1e-9 * L.vTe + 1e-6 * L.Va6 + 2e-6 * L.ud6;
        for (int i = 0; i < nI; i++) s += (i + 1) * (1e9 * L.c_kldi[i] + 1e6 * L.inv_s2vTi[i] + 1e5 * L.ioncf[i]);
        return s;
      };
      LG b; lg_zero(b);
      b.ne_g = 1e-20; b.omgL = 1e-15; b.omgpe2 = 1e-30; b.kL = 1e-5; b.vTe = 1e-9; b.Va6 = 1e-6; b.ud6 = 2e-6;
      for (int i = 0; i < nI; i++) { b.c_kldi[i] = (i + 1) * 1e9; b.inv_s2vTi[i] = (i + 1) * 1e6; b.ioncf[i] = (i + 1) * 1e5; }
      double pbar[NP] = {0};
      lg_backward(params, nI, g, G, lam_shift, b, pbar);
      for (int k = 0; k < NP; k++) {
        if (k == P_AMP1 || k == P_AMP2 || k == P_AMP3 || (k >= P_ION0 && (k - P_ION0) % 4 == ION_A)) continue;
        double p2[NP]; std::copy(params, params + NP, p2);
        double h = 1e-4 * std::max(1.0, fabs(params[k]));
        p2[k] = params[k] + h; double fp = fun(p2);
        p2[k] = params[k] - h; double fm = fun(p2);
        double fd = (fp - fm) / (2 * h);
        double e = fabs(fd - pbar[k]) / std::max(1e-3, std::max(fabs(fd), fabs(pbar[k])));
        if (e > 2e-6) { printf("FAIL lg_backward g=%d k=%d fd=%.10e ad=%.10e\n", g, k, fd, pbar[k]); fails++; }
      }
    }
    printf("lg_backward checked\n");
  }
  // ---- 2. point chain: d P / d params through kin/ion/assemble (+ synthetic xie-dependent chi_e)
  {
    const double lam_nm[3] = {450.0, 526.2, 610.0};
    for (int c = 0; c < 3; c++) {
      const double omgs = 2e7 * kPi * kC / lam_nm[c], cth = cos(60.0 * kPi / 180);
      LG L; lg_zero(L); lg_forward(params, nI, 0, 1, lam_shift, L);
      Kin q; kin_forward(L, omgs, cth, q);
      IonOut io; ion_forward(L, nI, T.zt, q, io);
      double I = sin(q.xie) * 0.7, dI = cos(q.xie) * 0.7, dfe = cos(0.5 * q.xie) * 0.3, ddfe = -0.15 * sin(0.5 * q.xie);
      double fp = exp(-0.5 * q.xie * q.xie) * 0.4, dfp = -q.xie * fp;
      double chiEr = -q.ikl2 * I, chiEi = kPi * q.ikl2 * dfe;
      Asm s; double P = assemble_forward(L, q, io, chiEr, chiEi, fp, omgs, s);
      PointBar pb; KinBar kb = {0, 0, 0, 0, 0, 0}; LG Lb; lg_zero(Lb);
      assemble_backward(L, nI, T.zt, q, io, chiEr, chiEi, fp, s, 1.0, pb, kb, Lb);
      kb.ikl2 += -I * pb.chiEr + kPi * dfe * pb.chiEi;
      kb.xie += (-q.ikl2 * pb.chiEr) * dI + (kPi * q.ikl2 * pb.chiEi) * ddfe + pb.fphi * dfp;
      kin_backward(L, omgs, cth, q, kb, Lb);
      double pbar[NP] = {0};
      lg_backward(params, nI, 0, 1, lam_shift, Lb, pbar);
      for (int k = 0; k < NP; k++) {
        if (k == P_AMP1 || k == P_AMP2 || k == P_AMP3 || (k >= P_ION0 && (k - P_ION0) % 4 == ION_A)) continue;
        if (k == P_NE_GRAD || k == P_TE_GRAD) { /* G=1 path */ }
        double p2[NP]; std::copy(params, params + NP, p2);
        double h = 1e-6 * std::max(1.0, fabs(params[k]));
        p2[k] = params[k] + h; double f1 = point_P(p2, nI, lam_shift, omgs, cth, T.zt, 0, 0, 0, true);
        p2[k] = params[k] - h; double f0 = point_P(p2, nI, lam_shift, omgs, cth, T.zt, 0, 0, 0, true);
        double fd = (f1 - f0) / (2 * h);
        double e = fabs(fd - pbar[k]) / std::max(1e-12 * fabs(P), std::max(fabs(fd), fabs(pbar[k])));
        if (e > 2e-5) { printf("FAIL point chain lam=%g k=%d fd=%.10e ad=%.10e (P=%.4e)\n", lam_nm[c], k, fd, pbar[k], P); fails++; }
      }
    }
    printf("point chain checked\n");
  }
  // ---- 2b. the "_x" point chain (invariant reciprocals hoisted) against the plain one: values and every cotangent
  {
    const double lam_nm[4] = {450.0, 526.2, 526.6, 610.0};
    double worst = 0;
    for (int c = 0; c < 4; c++) {
      const double omgs = 2e7 * kPi * kC / lam_nm[c], cth = cos((50.0 + 7 * c) * kPi / 180);
      LG L; lg_zero(L); lg_forward(params, nI, 0, 1, lam_shift, L);
      LGX X; lgx_make(L, nI, T.zt.h, X);
      Kin q; kin_forward(L, omgs, cth, q);
      KinX qx; kin_forward_x(L, X, omgs, cth, qx);
      IonOut io; ion_forward(L, nI, T.zt, q, io);
      IonX iox; ion_forward_x<0>(L, X, nI, T.zt, qx, iox);
      IonX iox2; ion_forward_x<2>(L, X, 0, T.zt, qx, iox2);
      const double chiEr = -q.ikl2 * 0.3, chiEi = kPi * q.ikl2 * 0.2, fp = 0.37;
      Asm s; const double P = assemble_forward(L, q, io, chiEr, chiEi, fp, omgs, s);
      AsmX sx; const double Px = assemble_forward_x(L, X, qx, iox, chiEr, chiEi, fp, omgs, sx);
      PointBar pb, pbx; KinBar kb = {0, 0, 0, 0, 0, 0}, kbx = {0, 0, 0, 0, 0, 0}; LG Lb, Lbx; lg_zero(Lb); lg_zero(Lbx);
      assemble_backward(L, nI, T.zt, q, io, chiEr, chiEi, fp, s, 1.3, pb, kb, Lb);
      assemble_backward_x<0>(L, X, nI, qx, iox, chiEr, chiEi, fp, sx, 1.3, pbx, kbx, Lbx);
      kb.ikl2 += 0.11 * pb.chiEr; kb.xie += 0.7 * pb.fphi;
      kbx.ikl2 += 0.11 * pbx.chiEr; kbx.xie += 0.7 * pbx.fphi;
      kin_backward(L, omgs, cth, q, kb, Lb);
      kin_backward_x(L, X, cth, qx, kbx, Lbx);
      double a[kLGDoubles + 8], b[kLGDoubles + 8];
      a[0] = Lb.ne_g; a[1] = Lb.omgL; a[2] = Lb.omgpe2; a[3] = Lb.kL; a[4] = Lb.vTe; a[5] = Lb.Va6; a[6] = Lb.ud6;
      b[0] = Lbx.ne_g; b[1] = Lbx.omgL; b[2] = Lbx.omgpe2; b[3] = Lbx.kL; b[4] = Lbx.vTe; b[5] = Lbx.Va6; b[6] = Lbx.ud6;
      for (int i = 0; i < TSFF_MAX_IONS; i++) {
        a[7 + i] = Lb.c_kldi[i]; a[11 + i] = Lb.inv_s2vTi[i]; a[15 + i] = Lb.ioncf[i];
        b[7 + i] = Lbx.c_kldi[i]; b[11 + i] = Lbx.inv_s2vTi[i]; b[15 + i] = Lbx.ioncf[i];
      }
      a[19] = P; b[19] = Px; a[20] = q.xie; b[20] = qx.xie; a[21] = q.ikl2; b[21] = qx.ikl2;
      a[22] = pb.chiEr; b[22] = pbx.chiEr; a[23] = pb.chiEi; b[23] = pbx.chiEi; a[24] = pb.fphi; b[24] = pbx.fphi;
      a[25] = io.chiIr; b[25] = iox2.chiIr; a[26] = io.sion; b[26] = iox2.sion;
      for (int k = 0; k < 27; k++) {
        const double e = relerr(a[k], b[k]);
        if (a[k] != 0.0 || b[k] != 0.0) worst = std::max(worst, e);
        if (e > 1e-11) { printf("FAIL x-chain lam=%g k=%d plain=%.15e x=%.15e\n", lam_nm[c], k, a[k], b[k]); fails++; }
      }
    }
    // Hermite / lerp with the inverse spacing passed in
    const int V = 64; const double x0 = -6 + 6.0 / V, h = 12.0 / V;
    std::vector<double> lnf(V), sl(V);
    for (int i = 0; i < V; i++) { double x = x0 + i * h; lnf[i] = -0.5 * x * x; sl[i] = -x; }
    for (double x : {-5.9, -5.5, -1.234, 0.0, 0.777, 4.9, 5.9, 7.0}) {
      Herm h1, h2;
      const double H1 = hermite_uniform(lnf.data(), sl.data(), V, x0, h, x, -50.0, h1);
      const double H2 = hermite_uniform_ih(lnf.data(), sl.data(), V, x0, h, 1.0 / h, x, -50.0, h2);
      int i1, i2; double t1, t2, s1, s2;
      const double l1 = lerp_uniform(lnf.data(), V, x0, h, x, i1, t1, s1), l2 = lerp_uniform_ih(lnf.data(), V, x0, 1.0 / h, x, i2, t2, s2);
      if (relerr(H1, H2) > 1e-13 || fabs(h1.dHdx - h2.dHdx) > 1e-12 * std::max(1.0, fabs(h1.dHdx)) || h1.inside != h2.inside ||
          (h1.i != h2.i && fabs(h1.t + (h1.i - h2.i) - h2.t) > 1e-9) || relerr(l1, l2) > 1e-13 || fabs(s1 - s2) > 1e-12 * std::max(1.0, fabs(s1))) {
        printf("FAIL hermite/lerp ih x=%g H %.15e %.15e dH %.12e %.12e lerp %.15e %.15e\n", x, H1, H2, h1.dHdx, h2.dHdx, l1, l2); fails++;
      }
    }
    {  // branch-free forms against the branchy ones
      std::vector<ZZ> zz(1640);
      for (int i = 0; i < 1640; i++) { zz[i].r = T.zr[i]; zz[i].i = T.zi[i]; }
      for (double x : {-9.0, -8.2, -8.19999, -3.3, -0.004, 0.0, 0.5, 2.22, 8.1899, 8.195, 12.0, 300.0}) {
        double a[4], b[4];
        zprime_lerp_x(T.zt, 1.0 / T.zt.h, x, a[0], a[1], a[2], a[3], 0, nullptr);
        zprime_lerp_bf(zz.data(), T.zt.n, T.zt.x0, T.zt.xlast, 1.0 / T.zt.h, x, b[0], b[1], b[2], b[3]);
        for (int k = 0; k < 4; k++)
          if (relerr(a[k], b[k]) > 1e-12 && fabs(a[k] - b[k]) > 1e-15) { printf("FAIL zprime bf x=%g k=%d %.15e %.15e\n", x, k, a[k], b[k]); fails++; }
      }
      for (double x : {-7.0, -5.9, -5.5, -1.234, 0.0, 0.777, 4.9, 5.9, 5.95, 7.0}) {
        Herm h1, h2;
        const double H1 = hermite_uniform_ih(lnf.data(), sl.data(), V, x0, h, 1.0 / h, x, -50.0, h1);
        const double H2 = hermite_uniform_bf(lnf.data(), sl.data(), V, x0, h, 1.0 / h, x, -50.0, h2);
        int i1, i2; double t1, t2, s1, s2;
        const double l1 = lerp_uniform_ih(lnf.data(), V, x0, 1.0 / h, x, i1, t1, s1), l2 = lerp_uniform_bf(lnf.data(), V, x0, 1.0 / h, x, i2, t2, s2);
        if (H1 != H2 || h1.dHdx != h2.dHdx || h1.inside != h2.inside || h1.i != h2.i || h1.t != h2.t || l1 != l2 || s1 != s2 || i1 != i2 || t1 != t2) {
          printf("FAIL hermite/lerp bf x=%g H %.15e %.15e lerp %.15e %.15e i %d %d t %g %g\n", x, H1, H2, l1, l2, i1, i2, t1, t2); fails++;
        }
      }
      for (double y : {-800.0, -700.0, -699.0, -50.0, -1.0, -1e-9, 0.0, 0.3, 5.0}) {
        if (relerr(fast_exp_bf(y), y > -700.0 ? exp(y) : 0.0) > 1e-14) { printf("FAIL exp bf y=%g\n", y); fails++; }
      }
    }
    printf("x-chain vs plain chain: worst rel diff %.2e\n", worst);
  }
  // ---- 3. PV sums
  {
    const int N = 512; const double z0 = -6 + 6.0 / N, h = 12.0 / N;
    std::vector<double> f(N), z(N);
    for (int i = 0; i < N; i++) { z[i] = z0 + i * h; f[i] = -z[i] * exp(-0.5 * z[i] * z[i]) * 0.4 + 0.01 * sin(3 * z[i]); }
    const int M = N - 2, nodes = M + 1, npad = (nodes + 31) / 32 * 32;
    std::vector<double> D64(npad, 0.0);
    for (int i = 0; i < npad; i++) D64[i] = pv_weight(f.data(), M, h, i);
    double maxe64 = 0, maxed = 0;
    const double xis[] = {-7.3, -5.99, -2.345678, -0.0117, 0.0, 0.4321, 1.0 + 1e-9, 3.3333, 5.97, 6.8};
    for (double xi : xis) {
      double ref = ratintn_literal(f, z, xi);
      double g0[1] = {z0 - xi}, bI[1], bJ[1];
      pv_accumulate_f64<1, true>(D64.data(), nodes, h, g0, bI, bJ);
      double I64, dI64; pv_finish(bI[0], bJ[0], f[0], f[M], z0 - xi, z0 + M * h - xi, I64, dI64);
      double e = 1e-6;
      double fd = (ratintn_literal(f, z, xi + e) - ratintn_literal(f, z, xi - e)) / (2 * e);
      maxe64 = std::max(maxe64, fabs(I64 - ref));
      maxed = std::max(maxed, fabs(dI64 - fd) / std::max(1.0, fabs(fd)));
      if (fabs(I64 - ref) > 1e-11 || fabs(dI64 - fd) > 2e-4 * std::max(1.0, fabs(fd))) {
        printf("FAIL pv xi=%g ref=%.12e I64=%.12e dI64=%.8e fd=%.8e\n", xi, ref, I64, dI64, fd); fails++;
      }
    }
    printf("pv (FP64 log form): max |I64-ref| = %.3e, max rel dI err = %.3e\n", maxe64, maxed);
  }
  // ---- 4. Hermite
  {
    const int V = 64; const double x0 = -6 + 6.0 / V, h = 12.0 / V;
    std::vector<double> lnf(V), sl(V);
    for (int i = 0; i < V; i++) { double x = x0 + i * h; lnf[i] = -0.5 * x * x - 0.1 * x * x * x * x / 10; }
    for (int i = 0; i < V; i++) sl[i] = (i == 0) ? (lnf[1] - lnf[0]) / h : (i == V - 1) ? (lnf[V - 1] - lnf[V - 2]) / h : (lnf[i + 1] - lnf[i - 1]) / (2 * h);
    for (double x : {-5.5, -1.234, 0.0, 0.777, 4.9}) {
      Herm hm; double H = hermite_uniform(lnf.data(), sl.data(), V, x0, h, x, -50.0, hm);
      Herm h2; double e = 1e-6;
      double fd = (hermite_uniform(lnf.data(), sl.data(), V, x0, h, x + e, -50.0, h2) - hermite_uniform(lnf.data(), sl.data(), V, x0, h, x - e, -50.0, h2)) / (2 * e);
      double wf0, wf1, wm0, wm1; hermite_weights(hm.t, h, wf0, wf1, wm0, wm1);
      double Hw = wf0 * lnf[hm.i - 1] + wf1 * lnf[hm.i] + wm0 * sl[hm.i - 1] + wm1 * sl[hm.i];
      if (fabs(fd - hm.dHdx) > 1e-6 * std::max(1.0, fabs(fd)) || fabs(Hw - H) > 1e-12 * std::max(1.0, fabs(H))) {
        printf("FAIL hermite x=%g H=%.10e Hw=%.10e dH=%.8e fd=%.8e\n", x, H, Hw, hm.dHdx, fd); fails++;
      }
    }
    printf("hermite checked\n");
  }
  // ---- 4b. child -> parent moment translation: Pascal recurrence (run-time child) vs the explicit matrix tree_T
  {
    double worst = 0;
    for (int c = 0; c < 4; c++) {
      double m0[kTK], m[kTK];
      for (int k = 0; k < kTK; k++) m0[k] = sin(1.0 + 0.7 * k + c) / (1.0 + 0.3 * k);
      tree_translate_shift(c, m0, m);
      for (int k = 0; k < kTK; k++) {
        double ref = 0.0, mag = 0.0;
        for (int j = 0; j <= k; j++) { ref += tree_T(c, k, j) * m0[j]; mag += fabs(tree_T(c, k, j) * m0[j]); }
        const double e = fabs(m[k] - ref) / std::max(1e-300, mag);
        worst = std::max(worst, e);
        if (e > 1e-14) { printf("FAIL translate c=%d k=%d %.16e %.16e\n", c, k, m[k], ref); fails++; }
      }
    }
    printf("moment translation (recurrence vs matrix): worst %.2e of the term magnitudes\n", worst);
  }
  // ---- 5. block-multipole PV evaluation (tsff_tree.cuh) vs the literal ratintn, several grid sizes
  for (int N : {4096, 1024, 512, 130, 37}) {
    const double z0 = -6 + 6.0 / N, h = 12.0 / N;
    std::vector<double> f(N), z(N);
    for (int i = 0; i < N; i++) { z[i] = z0 + i * h; f[i] = -z[i] * exp(-0.5 * z[i] * z[i]) * 0.4 + 0.01 * sin(3 * z[i]); }
    const int M = N - 2, npad = tree_npad(M + 1);
    const TreeBlob tb = tree_blob(npad);
    std::vector<double> ts(kTreeStaticDoubles);
    for (int i = 0; i < kTreeStaticDoubles; i++) ts[i] = tree_static_entry(i, M);
    std::vector<double> blob8(tb.bytes / 8 + 2, 0.0);   // 8-byte aligned storage
    unsigned char* blob = reinterpret_cast<unsigned char*>(blob8.data());
    float* W = reinterpret_cast<float*>(blob + tb.oW);
    for (int i = 1; i <= M - 1; i++) W[i] = (float)f[i];
    auto pget = [&](int i) { return f[i]; };
    for (int b = 0; b < tb.NB0; b++) {   // level 0: interior nodes only, no end-node rows
      double mu[kTK], A[kTK];
      tree_block_moments(pget, M, b, kTS0, kTs0, 0, 1, mu);
      tree_coeffs_from_moments(pget, M, b, kTS0, kTs0, mu, ts.data() + kTsCM0, (const double*)nullptr, A);
      tree_pack(A, reinterpret_cast<float4*>(blob + tb.oAB0) + b * (kTK / 2), reinterpret_cast<double*>(blob + tb.oLD0) + 2 * b);
    }
    for (int lvl = 0; lvl < 2; lvl++) {
      const int S = lvl ? kTS2 : kTS, nb = lvl ? tb.NB2 : tb.NB;
      const double s = lvl ? kTs2 : kTs;
      for (int b = 0; b < nb; b++) {
        double mu[kTK], A[kTK];
        tree_block_moments(pget, M, b, S, s, 0, 1, mu);
        tree_coeffs_from_moments(pget, M, b, S, s, mu, ts.data() + (lvl ? kTsCM2 : kTsCM1), ts.data() + (lvl ? kTsQE2 : kTsQE1), A);
        tree_pack(A, reinterpret_cast<float4*>(blob + (lvl ? tb.oAB2 : tb.oAB1)) + b * (kTK / 2),
                  reinterpret_cast<double*>(blob + (lvl ? tb.oLD2 : tb.oLD1)) + 2 * b);
      }
    }
    {  // the prep kernel's route to the level-2 moments (static tables E1, T12) vs the direct sums
      double worst = 0;
      for (int B2 = 0; B2 < tb.NB2; B2++) {
        double direct[kTK]; tree_block_moments(pget, M, B2, kTS2, kTs2, 0, 1, direct);
        double viaT[kTK] = {0};
        for (int c = 0; c < 4; c++) {
          double mu1[kTK] = {0};
          for (int o = 0; o < kTS; o++) { const int i = kTS * (4 * B2 + c) + o; if (i >= 1 && i <= M - 1) for (int k = 0; k < kTK; k++) mu1[k] += f[i] * ts[kTsE1 + o * kTK + k]; }
          for (int k = 0; k < kTK; k++) for (int j = 0; j <= k; j++) viaT[k] += ts[kTsT12 + (c * kTK + k) * kTK + j] * mu1[j];
        }
        double nrm = 0; for (int k = 0; k < kTK; k++) nrm = std::max(nrm, fabs(direct[k]));
        for (int k = 0; k < kTK; k++) worst = std::max(worst, fabs(viaT[k] - direct[k]) / std::max(nrm, 1e-300));
      }
      if (worst > 1e-12) { printf("FAIL moment translation N=%d err=%.3e\n", N, worst); fails++; }
    }
    double maxe = 0, maxd = 0, scale = 0;
    std::vector<double> xis = {-7.3, -5.99, -5.2, -2.345678, -0.0117, 0.0, 0.4321, 1.0 + 1e-9, 3.3333, 5.2, 5.97, 6.8};
    for (int k = 0; k < 60; k++) xis.push_back(-7.0 + 14.0 * (k + 0.37) / 60.0);
    for (double xi : xis) scale = std::max(scale, fabs(ratintn_literal(f, z, xi)));
    for (double xi : xis) {
      const double ref = ratintn_literal(f, z, xi);
      TreePole tp[1] = {tree_pole(xi, z0, h, M, npad)};
      double fI[1] = {0}, fJ1[1] = {0}, fJ2[1] = {0}, eI, eJ;
      tree_far<1>(blob, tb, tp, fI, fJ1, fJ2);
      const TreeAcc na = tree_near(blob, tb, tp[0]);
      tree_near_exact(xi, z0, h, M, (int)(-tp[0].un), tp[0].wb0, pget, eI, eJ);
      const double I = eI + na.I + fI[0], dI = eJ + na.J / h + fJ1[0] / (kTs * h) + fJ2[0] / (kTs2 * h);
      const double e = 1e-6;
      const double fd = (ratintn_literal(f, z, xi + e) - ratintn_literal(f, z, xi - e)) / (2 * e);
      maxe = std::max(maxe, fabs(I - ref) / scale); maxd = std::max(maxd, fabs(dI - fd) / std::max(1.0, fabs(fd)));
      if (fabs(I - ref) > 1.5e-7 * scale || fabs(dI - fd) > 2e-4 * std::max(1.0, fabs(fd))) {
        printf("FAIL tree N=%d xi=%g ref=%.12e I=%.12e dI=%.8e fd=%.8e\n", N, xi, ref, I, dI, fd); fails++;
      }
    }
    printf("tree N=%d: max |I-ref|/scale = %.3e, max rel dI err = %.3e\n", N, maxe, maxd);
  }
  printf(fails ? "HOSTSIM FAILED (%d)\n" : "HOSTSIM OK\n", fails);
  return fails ? 1 : 0;
}
