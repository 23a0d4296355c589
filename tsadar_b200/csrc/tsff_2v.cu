// tsff_2v.cu -- TSFF_MODE_2V: FormFactor.calc_in_2D (form_factor.py:449-587) for a 2-D electron distribution f(vx, vy)
// on a V x V grid, with the per-pole susceptibility of calc_chi_vals (:349-388):
//
//   per pole (gradient point, wavelength, angle):  k = ks - kL as a 2-vector (:512-516), xi_e vector (:552), |xi_e|,
//   beta = atan(xi_y / xi_x) + pi (1 - H(xi_x)) (:558);
//   F[a][b] = f(cos(beta) v_a - sin(beta) v_b, sin(beta) v_a + cos(beta) v_b)   bicubic (interpax interp2d "cubic",
//             extrap=True: tensor-product cubic Hermite, node slopes = mean of adjacent secants / one-sided at the ends,
//             the edge cell's polynomial outside the grid), rotate() :300-324 incl. its Fortran-order reshape;
//   f1[b] = dv sum_a F[a][b] (:371);  df = gradient(f1) (:372);  f(|xi|), f'(|xi|) by lerp (:376-377);
//   chi_e'' = pi/(k lambda_D)^2 f'(|xi|) (:381);  chi_e' = -1/(k lambda_D)^2 ratintn(df, vx - |xi|, vx) (:385-386);
//   then the same S(k, omega) assembly as the 1V path (:562-584).
//
// Work: V^2 bicubic interpolations (16 taps each) per pole -- 4.0e9 per forward at arts-2d (246 784 poles, V = 128).
// All FP64 (a 1e-7 error in f1' is amplified ~250x at an EPW resonance).  The f table sits in shared memory (padded
// rows, 129 KB at V = 128); a CTA of 1024 threads works on 1024/V poles at a time, thread b of a group owning column b
// of its pole's rotated table (no reductions in the hot loop).  Bound: FP64 pipe + shared-memory gathers.
// The adjoint (k_ff2v_bwd) scatters f1bar through the same 16 taps into a CTA-private fbar in shared memory, gathers
// d f1 / d beta from the patch derivatives, and reverses the 2-vector kinematics.
#include "tsff_common.cuh"

using namespace tsff;

namespace {
constexpr int kThreads2V = 1024;   // adjoint
constexpr int kThreadsFwd2V = 1024; // forward

struct Args2V {
  int W, A, G, nI, V, NP, P;       // P = G*W*A poles per parameter set
  double lam_shift, v0, dv, cos_va, sin_va, cos_ud, sin_ud;
  const double *omgs, *costh, *sinth;
  ZTab zt;
  const double* params;   // [B][NP]
  const double* fe;       // [B][V][V]
  double* ff;             // [B][G][W][A]
  double* f1save;         // [B][P][V] projected tables per pole (forward -> backward), or null
  // calc_all_chi_vals entry (form_factor.py:390-447): poles given by the caller instead of the kinematics
  const double *beta_in, *xie_in, *klde_in;   // [P] each, or null
  double* chi_out;                            // [3][P]: fe_vphi, chiEI, chiERrat
  // backward
  const double* ff_bar;   // [B][G][W][A]
  double* fe_bar;         // [B][V][V]
  double* fe_part;        // [grid][V][V] per-CTA partial tables of the adjoint (workspace), summed in order by k_ff2v_reduce
  double* lgbar;          // [B][G][kLGDoubles] (atomically accumulated: zero it)
};

// 1-D cubic-Hermite node weights on a uniform grid (interpax "cubic"): query coordinate q -> first node index i0 (clamped
// to >= 0; a clamped node carries weight 0) and weights on nodes i0 .. i0+3.
__device__ __forceinline__ void hermite4(double q, double v0, double idv, int V, int& i0, double (&w)[4]) {
  const double u = (q - v0) * idv;
  int i = (int)floor(u) + 1;           // searchsorted(x, q, "right")
  i = i < 1 ? 1 : (i > V - 1 ? V - 1 : i);
  const double t = u - (double)(i - 1);
  const double t2 = t * t;
  const double h00 = (2.0 * t - 3.0) * t2 + 1.0, h01 = 1.0 - h00;
  const double h10 = ((t - 2.0) * t + 1.0) * t, h11 = (t - 1.0) * t2;
  if (i == 1) {                        // slope at node 0 is one-sided
    w[0] = 0.0; w[1] = h00 - h10 - 0.5 * h11; w[2] = h01 + h10; w[3] = 0.5 * h11;
    i0 = 0;                            // nodes (0 [unused], 0, 1, 2): shift so that w[1] sits on node 0
    // layout below expects nodes i0 + {0,1,2,3} = {i-2, i-1, i, i+1}; with i = 1 node i-2 = -1 does not exist
    w[0] = w[1]; w[1] = w[2]; w[2] = w[3]; w[3] = 0.0;   // now weights on nodes 0, 1, 2, (3: zero)
    return;
  }
  if (i == V - 1) {                    // slope at node V-1 is one-sided; node i+1 = V does not exist
    w[0] = -0.5 * h10; w[1] = h00 - h11; w[2] = h01 + 0.5 * h10 + h11; w[3] = 0.0;
    i0 = i - 2;
    if (i0 + 3 > V - 1) {              // keep the 4-node window inside the table: shift left by one, weights follow
      i0 -= 1;
      w[3] = w[2]; w[2] = w[1]; w[1] = w[0]; w[0] = 0.0;
    }
    return;
  }
  w[0] = -0.5 * h10; w[1] = h00 - 0.5 * h11; w[2] = h01 + 0.5 * h10; w[3] = 0.5 * h11;
  i0 = i - 2;
}

// dynamic smem: f [V][V+1] | f1 [NG][V] | df [NG][V] | red [NG][8] | pole scalars [NG][8]
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_ff2v_fwd(const Args2V a, long long b_lineout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int V = a.V, VP = V + 1;
  const int GS = (V + 31) / 32 * 32;           // threads per pole group
  const int NG = NT / GS;              // poles in flight per CTA
  double* sf = reinterpret_cast<double*>(smem_raw);
  double* sf1 = sf + (size_t)V * VP;
  double* sdf = sf1 + (size_t)NG * V;
  double* sred = sdf + (size_t)NG * V;         // [NG][8]
  __shared__ LG sL[8];                          // G <= 8 gradient points
  const double* fe = a.fe + b_lineout * (long long)V * V;
  for (int i = threadIdx.x; i < V * V; i += NT) sf[(i / V) * VP + (i % V)] = fe[i];
  if (threadIdx.x < a.G) {
    LG L;
    lg_zero(L);
    lg_forward(a.params + b_lineout * a.NP, a.nI, threadIdx.x, a.G, a.lam_shift, L);
    sL[threadIdx.x] = L;
  }
  __syncthreads();
  const int grp = threadIdx.x / GS, tb = threadIdx.x % GS;
  const bool active_grp = grp < NG;
  const double idv = fast_rcp(a.dv);
  const int WA = a.W * a.A;
  const int M = V - 2;                          // ratintn uses nodes 0..M (N-2 intervals)

  for (long long p0 = (long long)blockIdx.x * NG; p0 < a.P; p0 += (long long)gridDim.x * NG) {
    const long long p = p0 + grp;
    const bool valid = active_grp && p < a.P;
    // ---- kinematics of this group's pole (every thread of the group computes them: ~200 FP64 ops vs ~9000 below)
    Kin q;
    double cb = 1.0, sb = 0.0, xmag = 0.0, omgs = 0.0;
    int g = 0, j = 0, ia = 0;
    if (valid && a.beta_in) {                    // calc_all_chi_vals: (beta, |xi|, k lambda_De) straight from the caller
      sincos(a.beta_in[p], &sb, &cb);
      xmag = a.xie_in[p];
      const double kl = a.klde_in[p];
      q.ikl2 = 1.0 / (kl * kl);
    } else if (valid) {
      g = (int)(p / WA);
      const int r = (int)(p % WA);
      j = r / a.A; ia = r % a.A;
      const LG& L = sL[g];
      omgs = a.omgs[j];
      const double ks = fast_sqrt(omgs * omgs - L.omgpe2) * (1.0 / kC);
      const double kx = a.costh[ia] * ks - L.kL, ky = a.sinth[ia] * ks;     // :513-515
      q.ks = ks;
      q.k2 = kx * kx + ky * ky;
      q.k = fast_sqrt(q.k2);
      q.omgdop = omgs - L.omgL - (kx * L.Va6 * a.cos_va + ky * L.Va6 * a.sin_va);   // :519
      q.w = q.omgdop * fast_rcp(q.k);
      const double ivTe = fast_rcp(L.vTe), ok2 = q.omgdop * fast_rcp(q.k2);
      const double xx = (ok2 * kx - L.ud6 * a.cos_ud) * ivTe, xy = (ok2 * ky - L.ud6 * a.sin_ud) * ivTe;   // :552
      xmag = sqrt(xx * xx + xy * xy);
      q.xie = xmag;
      q.ikl2 = L.omgpe2 * ivTe * ivTe * fast_rcp(q.k2);
      const double beta = atan(xy / xx) + (xx < 0.0 ? kPi : 0.0);           // :558 (heaviside(0) = 1)
      sincos(beta, &sb, &cb);
    }
    // ---- rotate + project: f1[b] = dv sum_a f(cb v_a - sb v_b, sb v_a + cb v_b)
    if (valid && tb < V) {
      const double vb = a.v0 + (double)tb * a.dv;
      double acc = 0.0;
      // Neighbouring lanes (b, b+1) read table cells (-sb, cb) apart: with VP = 1 mod 16 the shared-memory bank moves by
      // cb - sb per lane, and near beta = 45 / 225 deg a whole half-warp sits on one bank (measured 2.2x on the kernel).
      // There the threads start their walk over a at a lane-dependent offset (a = it +- b mod V), which adds +-(cb + sb) to
      // the bank stride; every thread still sums all V points, only the order of the sum changes.
      int aoff = 0;
      if (fabs(cb - sb) < 0.75) aoff = fabs(2.0 * cb) >= fabs(2.0 * sb) ? tb : (tb == 0 ? 0 : V - tb);
      // (a sliding 4x4 register window along the walk -- load only the entering row / column, ~5.7 instead of 16 LDS.64 per
      // point -- was built and measured in round 2: at the 512 threads it needs for 128 registers it ran 39.1 ms against the
      // 27.4 ms of this plain form: divergent shifts, 48 register moves per point and half the warps to hide latency)
      for (int it = 0; it < V; it++) {
        int aa = it + aoff;
        if (aa >= V) aa -= V;
        const double va = a.v0 + (double)aa * a.dv;
        const double xq = cb * va - sb * vb, yq = sb * va + cb * vb;
        int ix, iy;
        double wx[4], wy[4];
        hermite4(xq, a.v0, idv, V, ix, wx);
        hermite4(yq, a.v0, idv, V, iy, wy);
        const double* base = sf + ix * VP + iy;
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const double* row = base + m * VP;
          s = fma(wx[m], fma(wy[0], row[0], fma(wy[1], row[1], fma(wy[2], row[2], wy[3] * row[3]))), s);
        }
        acc += s;
      }
      sf1[grp * V + tb] = acc * a.dv;
      if (a.f1save) a.f1save[(b_lineout * a.P + p) * V + tb] = acc * a.dv;
    }
    __syncthreads();
    if (valid && tb < V) {   // np.gradient(f1, dv) (:372)
      const double* f1 = sf1 + grp * V;
      sdf[grp * V + tb] = tb == 0 ? (f1[1] - f1[0]) * idv : (tb == V - 1 ? (f1[V - 1] - f1[V - 2]) * idv : (f1[tb + 1] - f1[tb - 1]) * (0.5 * idv));
    }
    __syncthreads();
    // ---- PV integral of df against 1/(v - |xi|): exact second-difference form, one node per thread (tsff_pv.cuh)
    double term = 0.0;
    if (valid && tb <= M) {
      const double* df = sdf + grp * V;
      const double gi = a.v0 + (double)tb * a.dv - xmag;
      const double h = a.dv;
      if (tb == 0) term = df[0] * ((pv_phi(gi + h) - pv_phi(gi)) * idv - 1.0 - log_abs(gi));
      else if (tb == M) term = df[M] * ((pv_phi(gi - h) - pv_phi(gi)) * idv + 1.0 + log_abs(gi));
      else term = df[tb] * (pv_phi(gi + h) - 2.0 * pv_phi(gi) + pv_phi(gi - h)) * idv;
    }
    term = warp_sum(term);
    if (active_grp && (tb & 31) == 0) sred[grp * 8 + (tb >> 5)] = term;
    __syncthreads();
    if (valid && tb == 0) {
      double I = 0.0;
      for (int w = 0; w < GS / 32; w++) I += sred[grp * 8 + w];
      const LG& L = sL[g];
      int i_f; double t_f, sl_f;
      const double fphi = lerp_uniform(sf1 + grp * V, V, a.v0, a.dv, xmag, i_f, t_f, sl_f);   // :376
      const double d0 = sdf[grp * V + i_f], d1 = sdf[grp * V + i_f + 1];
      const double dfe = d0 + t_f * (d1 - d0);                                                // :377
      if (a.chi_out) {
        a.chi_out[p] = fphi;                                 // fe_vphi   :376
        a.chi_out[a.P + p] = kPi * q.ikl2 * dfe;             // chiEI     :381
        a.chi_out[2 * (long long)a.P + p] = -q.ikl2 * I;     // chiERrat  :385-386
      } else {
        IonOut io;
        ion_forward(L, a.nI, a.zt, q, io);
        Asm s;
        const double Pl = assemble_forward(L, q, io, -q.ikl2 * I, kPi * q.ikl2 * dfe, fphi, omgs, s);
        a.ff[((b_lineout * a.G + g) * (long long)a.W + j) * a.A + ia] = Pl;
      }
    }
    __syncthreads();
  }
}

// ---- adjoint ------------------------------------------------------------------------------------------------------
// hermite4 plus the derivative weights d w / d t (for d f / d beta through the query coordinates)
__device__ __forceinline__ void hermite4d(double q, double v0, double idv, int V, int& i0, double (&w)[4], double (&dw)[4]) {
  const double u = (q - v0) * idv;
  int i = (int)floor(u) + 1;
  i = i < 1 ? 1 : (i > V - 1 ? V - 1 : i);
  const double t = u - (double)(i - 1);
  const double t2 = t * t;
  const double h00 = (2.0 * t - 3.0) * t2 + 1.0, h01 = 1.0 - h00;
  const double h10 = ((t - 2.0) * t + 1.0) * t, h11 = (t - 1.0) * t2;
  const double g00 = 6.0 * (t2 - t), g01 = -g00;
  const double g10 = (3.0 * t - 4.0) * t + 1.0, g11 = (3.0 * t - 2.0) * t;
  if (i == 1) {
    i0 = 0;
    w[0] = h00 - h10 - 0.5 * h11; w[1] = h01 + h10; w[2] = 0.5 * h11; w[3] = 0.0;
    dw[0] = g00 - g10 - 0.5 * g11; dw[1] = g01 + g10; dw[2] = 0.5 * g11; dw[3] = 0.0;
    return;
  }
  if (i == V - 1) {
    i0 = i - 2;
    w[0] = -0.5 * h10; w[1] = h00 - h11; w[2] = h01 + 0.5 * h10 + h11; w[3] = 0.0;
    dw[0] = -0.5 * g10; dw[1] = g00 - g11; dw[2] = g01 + 0.5 * g10 + g11; dw[3] = 0.0;
    if (i0 + 3 > V - 1) {
      i0 -= 1;
      w[3] = w[2]; w[2] = w[1]; w[1] = w[0]; w[0] = 0.0;
      dw[3] = dw[2]; dw[2] = dw[1]; dw[1] = dw[0]; dw[0] = 0.0;
    }
    return;
  }
  i0 = i - 2;
  w[0] = -0.5 * h10; w[1] = h00 - 0.5 * h11; w[2] = h01 + 0.5 * h10; w[3] = 0.5 * h11;
  dw[0] = -0.5 * g10; dw[1] = g00 - 0.5 * g11; dw[2] = g01 + 0.5 * g10; dw[3] = 0.5 * g11;
}

__device__ __forceinline__ void atomic_add_shared(double* addr, double v) { atomicAdd(addr, v); }

// ---- FIXED-POINT accumulation in shared memory ------------------------------------------------------------------------
// sm_100 has ONE native shared-memory atomic add, the 32-bit integer one (ATOMS.ADD); FP64, FP32 and 64-bit integer adds
// on shared memory compile to compare-and-swap spin loops (ATOMS.CAST.SPIN, ~3x the shared-memory wavefronts).  A table
// cell is therefore a fixed-point number in CARRY-SAVE form, two 32-bit words (L, H) with value = H * 2^24 + L: an addend
// q (a signed integer, |q| < 2^51) is split into its low 24 bits (added to L) and the signed rest (added to H), both with
// the non-returning form of the native atomic -- fire and forget, no dependent second add waiting on the first one's
// return value.  L collects at most ~25 addends per pole and is normalised (carry moved into H) by a sweep of the CTA after
// every batch of poles, long before its 8 bits of headroom (255 addends) run out; H holds |sum| / 2^24 < 2^31, i.e. sums up
// to 2^55 units.  Integer addition is exact and associative, so the cell sums do not depend on the order in which the warps
// arrive (deterministic).  The addend arrives as t = value * 2^E + 1.5 * 2^52 (round to nearest in the FMA / add that forms
// it): for |value * 2^E| < 2^51 the low 52 mantissa bits of t hold 2^51 + q, so q = bits(t) - 0x4338000000000000 -- no
// F2I conversion.
constexpr double kFxMagic = 6755399441055744.0;   // 1.5 * 2^52
constexpr int kFxDigit = 24;
__device__ __forceinline__ long long fx_load(const unsigned int* cell) {
  return (long long)(int)cell[1] * (1LL << kFxDigit) + (long long)cell[0];
}
__device__ __forceinline__ void fx_store(unsigned int* cell, long long v) {
  cell[0] = (unsigned int)(v & ((1LL << kFxDigit) - 1));
  cell[1] = (unsigned int)(int)(v >> kFxDigit);
}
// Out-of-grid points (a query outside the table in x or y) extrapolate the edge cell's cubic: their weights reach ~1e9 and
// would eat the fixed-point range, and all of their taps land in the BAND of cells within three rows / columns of the table
// edge (hermite4's clamped windows).  They are accumulated in FP64 (CAS adds, ~15 % of the points) in a compact copy of
// that band: rows 0..2 and V-3..V-1 in full (6 V cells), then columns 0..2 and V-3..V-1 of the other rows.
__device__ __forceinline__ int band_cells(int V) { return 12 * V - 36; }
__device__ __forceinline__ int band_index(int r, int c, int V) {
  if (r < 3) return r * V + c;
  if (r >= V - 3) return (3 + r - (V - 3)) * V + c;
  const int cc = c < 3 ? c : (c >= V - 3 ? 3 + c - (V - 3) : -1);
  return cc < 0 ? -1 : 6 * V + (r - 3) * 6 + cc;
}

// Four fixed-point adds (carry-save form, see above): eight independent non-returning atomics.  t = value * 2^E + kFxMagic.
__device__ __forceinline__ void fx_add4(unsigned int* c0, unsigned int* c1, unsigned int* c2, unsigned int* c3, double t0, double t1,
                                        double t2, double t3) {
  unsigned int* cell[4] = {c0, c1, c2, c3};
  const double t[4] = {t0, t1, t2, t3};
#pragma unroll
  for (int n = 0; n < 4; n++) {
    const long long q = __double_as_longlong(t[n]) - 0x4338000000000000LL;
    atomicAdd(cell[n], (unsigned int)(q & ((1LL << kFxDigit) - 1)));
    atomicAdd(cell[n] + 1, (unsigned int)(int)(q >> kFxDigit));
  }
}

// One CTA works on kThreads2V / GS poles at a time (as the forward).  Per pole:
//   f1 (saved by the forward) -> df; per-node PV weights Wt_i and dWt_i/dxi -> I, dI/dxi; thread 0 reverses the
//   assembly -> Ibar, dfe_bar, fphi_bar, kinematic cotangents; df_bar -> f1_bar; then the rotate/project adjoint: every
//   (a, b) point scatters f1_bar[b] dv wx wy into the CTA's private fbar in shared memory -- in-grid points into the
//   64-bit fixed-point table (two native 32-bit atomics per tap, fx_add), out-of-grid points into the FP64 edge band (a warp
//   owns a line of the mesh and its lanes take points four nodes apart so that their 4x4 stencils barely collide) -- and
//   gathers d f1[b] / d beta from f; finally the kinematics reverse (|xi|, beta -> parameters).
// Fixed-point scale: the table holds sums scaled by 2^E.  Per pole a cell gains at most ~1.7 max_b|f1_bar_b| dv from the
// in-grid points (the |wx wy| of the ~16 mesh points whose stencils cover it sum to (int |w|)^2 = 1.64); kFxGrow = 26 is the
// bound used.  E is chosen at the first batch of poles for kFxGrow * (poles of this CTA) * (first batch's max) with a
// factor 4 to spare (table sums up to 2^54 units), which leaves ~2^-36 of a typical addend as the rounding unit; a running bound is kept, and should a
// later batch exceed it (or bring addends >= 2^49 units) the whole table is shifted right first (rare, exact to one unit).
// dynamic smem: f [V][V+1] f32 | fbar [V][V+1] int64 | 2 x [NG][V] f64 | band [12 V - 36] f64 | red [NG][16] | scal [NG][16]
// WANT_PARAMS = false: no kinematic parameter is trainable (only the table is, as in the reference's arts-2d deck):
// d f1 / d beta, the 2-vector kinematics reverse and the lineout-scalar cotangents are skipped.
constexpr double kFxGrow = 26.0;
template <bool WANT_PARAMS>
__global__ void __launch_bounds__(kThreads2V, 1) k_ff2v_bwd(const Args2V a, long long b_lineout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int V = a.V, VP = V + 1;
  const int GS = (V + 31) / 32 * 32, NG = kThreads2V / GS, NWG = GS / 32;
  float* sf = reinterpret_cast<float*>(smem_raw);                 // f as FP32 (only d/dbeta reads it; 1e-4 is asked)
  unsigned int* sfx = reinterpret_cast<unsigned int*>(smem_raw + (((size_t)V * VP * 4 + 15) / 16) * 16);   // [V][VP] x (lo, hi)
  double* sA = reinterpret_cast<double*>(sfx) + (size_t)V * VP;   // [NG][V]: f1, later df_bar
  double* sB = sA + (size_t)NG * V;            // [NG][V]: df, later f1_bar
  double* sband = sB + (size_t)NG * V;         // [12 V - 36] out-of-grid contributions (FP64)
  double* sred = sband + band_cells(V);        // [NG][16]
  double* ssc = sred + (size_t)NG * 16;        // [NG][16]
  __shared__ LG sL[8];
  __shared__ double sLb[8][kLGDoubles];
  __shared__ double s_scale, s_bound;          // 2^E; running bound on |cell| in table units / 2^E
  __shared__ int s_E, s_shift, s_have;
  const double* fe = a.fe + b_lineout * (long long)V * V;
  if (WANT_PARAMS)
    for (int i = threadIdx.x; i < V * V; i += kThreads2V) sf[(i / V) * VP + (i % V)] = (float)fe[i];
  for (int i = threadIdx.x; i < 2 * V * VP; i += kThreads2V) sfx[i] = 0u;
  for (int i = threadIdx.x; i < band_cells(V); i += kThreads2V) sband[i] = 0.0;
  for (int i = threadIdx.x; i < 8 * kLGDoubles; i += kThreads2V) (&sLb[0][0])[i] = 0.0;
  if (threadIdx.x < a.G) {
    LG L;
    lg_zero(L);
    lg_forward(a.params + b_lineout * a.NP, a.nI, threadIdx.x, a.G, a.lam_shift, L);
    sL[threadIdx.x] = L;
  }
  if (threadIdx.x == 0) { s_scale = 0.0; s_bound = 0.0; s_E = 0; s_shift = 0; s_have = 0; }
  __syncthreads();
  const int grp = threadIdx.x / GS, tb = threadIdx.x % GS, wg = tb >> 5, lane = tb & 31;
  const bool active_grp = grp < NG;
  const double idv = fast_rcp(a.dv), h = a.dv;
  const int WA = a.W * a.A, M = V - 2;
  const double vlast = a.v0 + (double)(V - 1) * a.dv;
  const long long nbatch_cta = (((long long)a.P + NG - 1) / NG - blockIdx.x + gridDim.x - 1) / gridDim.x;   // batches this CTA will see

  for (long long p0 = (long long)blockIdx.x * NG; p0 < a.P; p0 += (long long)gridDim.x * NG) {
    const long long p = p0 + grp;
    const bool valid = active_grp && p < a.P;
    Kin q;
    double cb = 1.0, sb = 0.0, xmag = 0.0, omgs = 0.0, kx = 0.0, ky = 0.0, xx = 1.0, xy = 0.0;
    int g = 0, j = 0, ia = 0;
    if (valid) {
      g = (int)(p / WA);
      const int r = (int)(p % WA);
      j = r / a.A; ia = r % a.A;
      const LG& L = sL[g];
      omgs = a.omgs[j];
      const double ks = fast_sqrt(omgs * omgs - L.omgpe2) * (1.0 / kC);
      kx = a.costh[ia] * ks - L.kL; ky = a.sinth[ia] * ks;
      q.ks = ks;
      q.k2 = kx * kx + ky * ky;
      q.k = fast_sqrt(q.k2);
      q.omgdop = omgs - L.omgL - (kx * L.Va6 * a.cos_va + ky * L.Va6 * a.sin_va);
      q.w = q.omgdop * fast_rcp(q.k);
      const double ivTe = fast_rcp(L.vTe), ok2 = q.omgdop * fast_rcp(q.k2);
      xx = (ok2 * kx - L.ud6 * a.cos_ud) * ivTe; xy = (ok2 * ky - L.ud6 * a.sin_ud) * ivTe;
      xmag = sqrt(xx * xx + xy * xy);
      q.xie = xmag;
      q.ikl2 = L.omgpe2 * ivTe * ivTe * fast_rcp(q.k2);
      const double beta = atan(xy / xx) + (xx < 0.0 ? kPi : 0.0);
      sincos(beta, &sb, &cb);
    }
    // ---- f1 (saved), df, PV node weights
    if (valid && tb < V) sA[grp * V + tb] = a.f1save[(b_lineout * a.P + p) * V + tb];
    __syncthreads();
    if (valid && tb < V) {
      const double* f1 = sA + grp * V;
      sB[grp * V + tb] = tb == 0 ? (f1[1] - f1[0]) * idv : (tb == V - 1 ? (f1[V - 1] - f1[V - 2]) * idv : (f1[tb + 1] - f1[tb - 1]) * (0.5 * idv));
    }
    __syncthreads();
    double tI = 0.0, tJ = 0.0, wt = 0.0;        // wt: this thread's PV node weight (kept in a register until df_bar)
    if (valid && tb < V) {
      double dwt = 0.0;
      if (tb <= M) {
        const double gi = a.v0 + (double)tb * h - xmag;
        const double lc = log_abs(gi);
        if (tb == 0) {
          const double lp = log_abs(gi + h);
          wt = ((gi + h) * lp - gi * lc) * idv - 1.0 - lc;
          dwt = -(lp - lc) * idv + fast_rcp(gi);
        } else if (tb == M) {
          const double lm = log_abs(gi - h);
          wt = ((gi - h) * lm - gi * lc) * idv + 1.0 + lc;
          dwt = -(lm - lc) * idv - fast_rcp(gi);
        } else {
          const double lp = log_abs(gi + h), lm = log_abs(gi - h);
          wt = ((gi + h) * lp - 2.0 * gi * lc + (gi - h) * lm) * idv;
          dwt = -(lp - 2.0 * lc + lm) * idv;
        }
      }
      tI = sB[grp * V + tb] * wt;
      tJ = sB[grp * V + tb] * dwt;
    }
    tI = warp_sum(tI);
    tJ = warp_sum(tJ);
    if (active_grp && lane == 0) { sred[grp * 16 + wg] = tI; sred[grp * 16 + 8 + wg] = tJ; }
    __syncthreads();
    // ---- thread 0 of the group: reverse of the assembly
    LG Lb;
    lg_zero(Lb);
    KinBar kb = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (valid && tb == 0) {
      double I = 0.0, J = 0.0;
      for (int w = 0; w < NWG; w++) { I += sred[grp * 16 + w]; J += sred[grp * 16 + 8 + w]; }
      const LG& L = sL[g];
      int i_f; double t_f, sl_f;
      const double fphi = lerp_uniform(sA + grp * V, V, a.v0, a.dv, xmag, i_f, t_f, sl_f);
      const double d0 = sB[grp * V + i_f], d1 = sB[grp * V + i_f + 1];
      const bool clamped = (xmag <= a.v0) || (xmag >= a.v0 + (V - 1) * a.dv) || !(xmag == xmag);
      const double dfe = d0 + t_f * (d1 - d0);
      const double sl_d = clamped ? 0.0 : (d1 - d0) * idv;
      IonOut io;
      ion_forward(L, a.nI, a.zt, q, io);
      Asm s;
      const double chiEr = -q.ikl2 * I, chiEi = kPi * q.ikl2 * dfe;
      assemble_forward(L, q, io, chiEr, chiEi, fphi, omgs, s);
      PointBar pb;
      const double Pbar = a.ff_bar[((b_lineout * a.G + g) * (long long)a.W + j) * a.A + ia];
      assemble_backward(L, a.nI, a.zt, q, io, chiEr, chiEi, fphi, s, Pbar, pb, kb, Lb);
      kb.ikl2 += -I * pb.chiEr + kPi * dfe * pb.chiEi;
      const double Ibar = -q.ikl2 * pb.chiEr, dfe_bar = kPi * q.ikl2 * pb.chiEi;
      double* sc = ssc + grp * 16;
      sc[0] = Ibar; sc[1] = dfe_bar; sc[2] = pb.fphi; sc[3] = (double)i_f; sc[4] = t_f;
      sc[5] = Ibar * J + dfe_bar * sl_d + pb.fphi * sl_f;   // cotangent of |xi|
    }
    __syncthreads();
    // ---- df_bar (into sA), then f1_bar (into sB)
    double Ibar = 0.0, dfe_bar = 0.0, fphi_bar = 0.0, t_f = 0.0;
    int i_f = 0;
    if (valid) {
      const double* sc = ssc + grp * 16;
      Ibar = sc[0]; dfe_bar = sc[1]; fphi_bar = sc[2]; i_f = (int)sc[3]; t_f = sc[4];
    }
    if (valid && tb < V) {
      double v = Ibar * wt;
      if (tb == i_f) v += (1.0 - t_f) * dfe_bar;
      if (tb == i_f + 1) v += t_f * dfe_bar;
      sA[grp * V + tb] = v;
    }
    __syncthreads();
    double amax = 0.0;                           // max_b |f1_bar_b| dv of this group's pole (for the fixed-point bound)
    if (valid && tb < V) {
      const double* dfb = sA + grp * V;
      const int k = tb;
      double fb = 0.0;
      if (k >= 1) fb += dfb[k - 1] * ((k - 1 == 0) ? idv : 0.5 * idv);
      if (k <= V - 2) fb -= dfb[k + 1] * ((k + 1 == V - 1) ? idv : 0.5 * idv);
      if (k == 0) fb -= dfb[0] * idv;
      if (k == V - 1) fb += dfb[V - 1] * idv;
      if (k == i_f) fb += (1.0 - t_f) * fphi_bar;
      if (k == i_f + 1) fb += t_f * fphi_bar;
      sB[grp * V + k] = fb;
      amax = fabs(fb) * a.dv;
      if (!(amax < 1e300)) amax = 0.0;           // NaN / inf cotangents propagate through the band-free FP64 flush below
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) sred[(threadIdx.x >> 5)] = amax;          // one slot per warp of the CTA (sred holds NG * 16 >= 32 doubles)
    __syncthreads();
    // ---- fixed-point scale for this batch (thread 0), table shift if the running bound would overflow
    if (threadIdx.x == 0) {
      double m = 0.0;
      for (int w = 0; w < kThreads2V / 32; w++) m = fmax(m, sred[w]);
      int shift = 0;
      if (m > 0.0) {
        int em;
        frexp(m, &em);                                       // m < 2^em
        if (!s_have) {
          int eb;
          frexp(kFxGrow * (double)NG * m * (double)nbatch_cta * 4.0, &eb);
          int E = 53 - eb;
          if (E > 49 - em) E = 49 - em;
          s_E = E; s_scale = ldexp(1.0, E); s_bound = 0.0; s_have = 1;
        }
        const double need = s_bound + kFxGrow * (double)NG * m;
        int en;
        frexp(need, &en);
        int E = s_E;
        if (en + E > 53) E = 53 - en;
        if (em + E > 49) E = 49 - em;
        shift = s_E - E;
        if (shift > 0) { s_E = E; s_scale = ldexp(1.0, E); }
        s_bound = need;
      }
      s_shift = shift;
    }
    __syncthreads();
    if (s_shift > 0) {
      const int sh = s_shift;
      for (int i = threadIdx.x; i < V * VP; i += kThreads2V) {
        long long v = fx_load(sfx + 2 * i);
        v = sh >= 63 ? (v < 0 ? -1 : 0) : (v >> sh);
        fx_store(sfx + 2 * i, v);
      }
      __syncthreads();
    }
    const double scale = s_scale;
    // ---- rotate/project adjoint: warp wg of the group owns lines o = wg, wg + NWG, ... of the mesh (columns b, or rows a),
    //      lanes take the points 4 lane + c along the line.
    //      (Measured alternative, round 2: one thread per line walking it with a 4x4 accumulator window in registers and
    //      flushing only the row / column that leaves the window -- ~11 instead of 32 atomics per point -- needs 128 registers,
    //      i.e. 512 threads, and executes ~470 instructions per warp-step with the divergent shifts: 122 ms against the 104 ms
    //      of this form at arts-2d, table-only.)
    double bbar = 0.0;
    if (valid) {
      // lanes step 4 nodes along a (address step 4 (cb VP + sb), i.e. 4 (cb + sb) banks since VP = 1 mod 16) or along b
      // (4 (cb - sb)): take the direction with the larger bank step, the other one piles the lanes onto a few banks
      const bool along_b = fabs(cb - sb) > fabs(cb + sb);
      for (int o = wg; o < V; o += NWG) {
        for (int c = 0; c < 4; c++) {
          const int in = 4 * lane + c;
          if (in >= V) break;
          const int aa = along_b ? o : in, b = along_b ? in : o;
          const double wcol = sB[grp * V + b] * a.dv;
          if (wcol == 0.0) continue;
          const double vb = a.v0 + (double)b * a.dv;
          const double va = a.v0 + (double)aa * a.dv;
          const double xq = cb * va - sb * vb, yq = sb * va + cb * vb;
          int ix, iy;
          double wx[4], wy[4], dwx[4], dwy[4];
          if (WANT_PARAMS) {
            hermite4d(xq, a.v0, idv, V, ix, wx, dwx);
            hermite4d(yq, a.v0, idv, V, iy, wy, dwy);
          } else {
            hermite4(xq, a.v0, idv, V, ix, wx);
            hermite4(yq, a.v0, idv, V, iy, wy);
          }
          const bool ingrid = xq >= a.v0 && xq <= vlast && yq >= a.v0 && yq <= vlast && wcol == wcol;
          if (WANT_PARAMS) {
            const float* fb = sf + ix * VP + iy;
            double sx = 0.0, sy = 0.0;
#pragma unroll
            for (int m = 0; m < 4; m++) {
              double r0 = 0.0, r1 = 0.0;
#pragma unroll
              for (int n = 0; n < 4; n++) {
                const double fv = (double)fb[m * VP + n];
                r0 = fma(wy[n], fv, r0);
                r1 = fma(dwy[n], fv, r1);
              }
              sx = fma(dwx[m], r0, sx);
              sy = fma(wx[m], r1, sy);
            }
            // d xq / d beta = -yq, d yq / d beta = xq;  d/dxq = idv d/dt
            bbar += wcol * (-yq * sx + xq * sy) * idv;
          }
          if (ingrid) {
            // one DFMA per tap forms value * 2^E + magic; eight independent native atomics per stencil row
            unsigned int* ob = sfx + 2 * (ix * VP + iy);
            const double ws = wcol * scale;
#pragma unroll
            for (int m = 0; m < 4; m++) {
              const double wm = ws * wx[m];
              unsigned int* r = ob + 2 * (m * VP);
              fx_add4(r, r + 2, r + 4, r + 6, fma(wm, wy[0], kFxMagic), fma(wm, wy[1], kFxMagic), fma(wm, wy[2], kFxMagic),
                      fma(wm, wy[3], kFxMagic));
            }
          } else {
#pragma unroll
            for (int m = 0; m < 4; m++) {
#pragma unroll
              for (int n = 0; n < 4; n++) {
                const double wv = wcol * wx[m] * wy[n];
                if (wv != 0.0) atomic_add_shared(sband + band_index(ix + m, iy + n, V), wv);
              }
            }
          }
        }
      }
    }
    bbar = warp_sum(bbar);
    __syncthreads();                                       // all adds of this batch are in; sred is reused below
    for (int i = threadIdx.x; i < V * VP; i += kThreads2V) {   // carry-save normalisation: L's carry into H
      const unsigned int L = sfx[2 * i];
      if (L >> kFxDigit) {
        sfx[2 * i] = L & ((1u << kFxDigit) - 1u);
        sfx[2 * i + 1] += L >> kFxDigit;
      }
    }
    if (active_grp && lane == 0) sred[grp * 16 + wg] = bbar;
    __syncthreads();
    // ---- kinematics reverse for (|xi|, beta) and the rest of kb
    if (WANT_PARAMS && valid && tb == 0) {
      double beta_bar = 0.0;
      for (int w = 0; w < NWG; w++) beta_bar += sred[grp * 16 + w];
      const LG& L = sL[g];
      const double xmag_bar = ssc[grp * 16 + 5];
      const double ixm = fast_rcp(xmag);
      // xmag = sqrt(xx^2 + xy^2); beta = atan(xy/xx) (+ const)
      const double xx_bar = xmag_bar * xx * ixm - beta_bar * xy * ixm * ixm;
      const double xy_bar = xmag_bar * xy * ixm + beta_bar * xx * ixm * ixm;
      // xx = (ok2 kx - udx) / vTe,  ok2 = omgdop / k2
      const double ivTe = fast_rcp(L.vTe), ik2 = fast_rcp(q.k2), ok2 = q.omgdop * ik2;
      Lb.vTe += -(xx_bar * xx + xy_bar * xy) * ivTe;
      Lb.ud6 += -(xx_bar * a.cos_ud + xy_bar * a.sin_ud) * ivTe;
      const double ok2_bar = (xx_bar * kx + xy_bar * ky) * ivTe;
      double kx_bar = xx_bar * ok2 * ivTe, ky_bar = xy_bar * ok2 * ivTe;
      kb.omgdop += ok2_bar * ik2;
      kb.k2 += -ok2_bar * ok2 * ik2;
      // ikl2 = omgpe2 / (vTe^2 k2)
      Lb.omgpe2 += kb.ikl2 * q.ikl2 * fast_rcp(L.omgpe2);
      Lb.vTe += -2.0 * kb.ikl2 * q.ikl2 * ivTe;
      kb.k2 += -kb.ikl2 * q.ikl2 * ik2;
      // w = omgdop / k
      const double ik = fast_rcp(q.k);
      kb.omgdop += kb.w * ik;
      kb.k += -kb.w * q.w * ik;
      // omgdop = omgs - omgL - (kx Vax + ky Vay)
      Lb.omgL += -kb.omgdop;
      kx_bar += -kb.omgdop * L.Va6 * a.cos_va;
      ky_bar += -kb.omgdop * L.Va6 * a.sin_va;
      Lb.Va6 += -kb.omgdop * (kx * a.cos_va + ky * a.sin_va);
      // k = sqrt(k2); k2 = kx^2 + ky^2
      kb.k2 += kb.k * (0.5 * ik);
      kx_bar += kb.k2 * 2.0 * kx;
      ky_bar += kb.k2 * 2.0 * ky;
      // kx = cos(sa) ks - kL, ky = sin(sa) ks;  ks = sqrt(omgs^2 - omgpe2)/C
      const double ks_bar = kx_bar * a.costh[ia] + ky_bar * a.sinth[ia];
      Lb.kL += -kx_bar;
      Lb.omgpe2 += -ks_bar * fast_rcp(2.0 * kC * kC * q.ks);
      double vals[kLGDoubles];
      vals[0] = Lb.ne_g; vals[1] = Lb.omgL; vals[2] = Lb.omgpe2; vals[3] = Lb.kL; vals[4] = Lb.vTe; vals[5] = Lb.Va6; vals[6] = Lb.ud6;
      for (int i = 0; i < TSFF_MAX_IONS; i++) { vals[7 + i] = Lb.c_kldi[i]; vals[7 + TSFF_MAX_IONS + i] = Lb.inv_s2vTi[i]; vals[7 + 2 * TSFF_MAX_IONS + i] = Lb.ioncf[i]; }
      for (int i = 0; i < kLGDoubles; i++) if (vals[i] != 0.0) atomicAdd(&sLb[g][i], vals[i]);
    }
    __syncthreads();
  }
  // ---- flush: this CTA's table (fixed point -> FP64, plus the edge band) goes to ITS OWN slab of the workspace; the slabs
  //      are summed in a fixed order by k_ff2v_reduce (no floating-point atomics across CTAs: deterministic)
  double* part = a.fe_part + (long long)blockIdx.x * V * V;
  const double inv_scale = s_have ? ldexp(1.0, -s_E) : 0.0;
  for (int i = threadIdx.x; i < V * V; i += kThreads2V) {
    const int r = i / V, c = i % V;
    double v = (double)fx_load(sfx + 2 * (r * VP + c)) * inv_scale;
    const int bi = band_index(r, c, V);
    if (bi >= 0) v += sband[bi];
    part[i] = v;
  }
  for (int i = threadIdx.x; i < a.G * kLGDoubles; i += kThreads2V) {
    const double v = sLb[i / kLGDoubles][i % kLGDoubles];
    if (v != 0.0) atomicAdd(&a.lgbar[(b_lineout * a.G) * kLGDoubles + i], v);
  }
}

// fe_bar[b][i] = sum over the CTAs' slabs, in CTA order
__global__ void __launch_bounds__(256) k_ff2v_reduce(const double* part, int nparts, int n, double* out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < nparts; k++) s += part[(long long)k * n + i];
  out[i] = s;
}

__global__ void k_ff2v_params_bar(const Args2V a, double* params_bar, int64_t B) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= B) return;
  double* pbar = params_bar + b * a.NP;
  for (int k = 0; k < a.NP; k++) pbar[k] = 0.0;
  for (int g = 0; g < a.G; g++) {
    const double* src = a.lgbar + (b * a.G + g) * kLGDoubles;
    LG Lb;
    Lb.ne_g = src[0]; Lb.omgL = src[1]; Lb.omgpe2 = src[2]; Lb.kL = src[3]; Lb.vTe = src[4]; Lb.Va6 = src[5]; Lb.ud6 = src[6];
    for (int i = 0; i < TSFF_MAX_IONS; i++) { Lb.c_kldi[i] = src[7 + i]; Lb.inv_s2vTi[i] = src[7 + TSFF_MAX_IONS + i]; Lb.ioncf[i] = src[7 + 2 * TSFF_MAX_IONS + i]; }
    lg_backward(a.params + b * a.NP, a.nI, g, a.G, a.lam_shift, Lb, pbar);
  }
}
}  // namespace

namespace tsff {
size_t ff2v_saved_bytes(const tsff_ctx* c, int64_t B) { return align_up((size_t)B * c->G * c->W * c->A * c->V * 8); }
size_t ff2v_ws_bytes(const tsff_ctx* c, int64_t B) {
  return align_up((size_t)B * c->G * kLGDoubles * 8) + align_up((size_t)c->sm_count * c->V * c->V * 8);   // lgbar | per-CTA partial tables
}

int ff2v_fwd(tsff_ctx* c, int64_t B, const double* params, const double* fe, double* ff_out, void* saved, cudaStream_t st) {
  const int V = c->V;
  if (V > 128 || V < 8) { set_error("2V path: V must be in [8, 128] (got %d)", V); return TSFF_E_INVALID; }
  if (c->G > 8) { set_error("2V path: at most 8 gradient points"); return TSFF_E_INVALID; }
  Args2V a;
  memset(&a, 0, sizeof(a));
  a.W = c->W; a.A = c->A; a.G = c->G; a.nI = c->I; a.V = V; a.NP = c->NP; a.P = c->G * c->W * c->A;
  a.lam_shift = c->lam_shift; a.v0 = c->v0; a.dv = c->dv;
  a.cos_va = cos(c->va_angle_deg * kPi / 180.0); a.sin_va = sin(c->va_angle_deg * kPi / 180.0);
  a.cos_ud = cos(c->ud_angle_deg * kPi / 180.0); a.sin_ud = sin(c->ud_angle_deg * kPi / 180.0);
  a.omgs = c->omgs; a.costh = c->costh; a.sinth = c->sinth; a.zt = c->zt;
  a.params = params; a.fe = fe; a.ff = ff_out; a.f1save = static_cast<double*>(saved);
  const int GS = (V + 31) / 32 * 32, NG = kThreadsFwd2V / GS;
  const size_t smem = ((size_t)V * (V + 1) + (size_t)NG * V * 2 + (size_t)NG * 16) * 8;
  TSFF_SMEM_OPTIN(k_ff2v_fwd<kThreadsFwd2V>);
  const long long nbatch = ((long long)a.P + NG - 1) / NG;
  const unsigned grid = (unsigned)(nbatch < c->sm_count ? nbatch : c->sm_count);
  for (int64_t b = 0; b < B; b++) {
    k_ff2v_fwd<kThreadsFwd2V><<<grid, kThreadsFwd2V, smem, st>>>(a, (long long)b);
    TSFF_LAUNCH_OK("k_ff2v_fwd");
  }
  return TSFF_OK;
}

// FormFactor.calc_all_chi_vals (form_factor.py:390-447): chi_out = [fe_vphi | chiEI | chiERrat], each [P]
int chi2v_fwd(tsff_ctx* c, const double* fe, const double* beta, const double* xie_mag, const double* klde_mag, int64_t P,
              double* chi_out, cudaStream_t st) {
  const int V = c->V;
  if (V > 128 || V < 8) { set_error("2V path: V must be in [8, 128] (got %d)", V); return TSFF_E_INVALID; }
  if (P > 0x7fffffffLL) { set_error("2V path: too many poles"); return TSFF_E_INVALID; }
  Args2V a;
  memset(&a, 0, sizeof(a));
  a.W = 1; a.A = 1; a.G = 0; a.nI = c->I; a.V = V; a.NP = c->NP; a.P = (int)P;
  a.v0 = c->v0; a.dv = c->dv; a.zt = c->zt;
  a.fe = fe; a.beta_in = beta; a.xie_in = xie_mag; a.klde_in = klde_mag; a.chi_out = chi_out;
  const int GS = (V + 31) / 32 * 32, NG = kThreadsFwd2V / GS;
  const size_t smem = ((size_t)V * (V + 1) + (size_t)NG * V * 2 + (size_t)NG * 16) * 8;
  TSFF_SMEM_OPTIN(k_ff2v_fwd<kThreadsFwd2V>);
  const long long nbatch = ((long long)a.P + NG - 1) / NG;
  const unsigned grid = (unsigned)(nbatch < c->sm_count ? nbatch : c->sm_count);
  k_ff2v_fwd<kThreadsFwd2V><<<grid, kThreadsFwd2V, smem, st>>>(a, 0LL);
  TSFF_LAUNCH_OK("k_ff2v_fwd (chi)");
  return TSFF_OK;
}

int ff2v_bwd(tsff_ctx* c, int64_t B, const double* params, const double* fe, const void* saved, const double* ff_bar,
             double* params_bar, double* fe_bar, void* ws, cudaStream_t st) {
  const int V = c->V;
  if (V > 128 || V < 8) { set_error("2V path: V must be in [8, 128] (got %d)", V); return TSFF_E_INVALID; }
  Args2V a;
  memset(&a, 0, sizeof(a));
  a.W = c->W; a.A = c->A; a.G = c->G; a.nI = c->I; a.V = V; a.NP = c->NP; a.P = c->G * c->W * c->A;
  a.lam_shift = c->lam_shift; a.v0 = c->v0; a.dv = c->dv;
  a.cos_va = cos(c->va_angle_deg * kPi / 180.0); a.sin_va = sin(c->va_angle_deg * kPi / 180.0);
  a.cos_ud = cos(c->ud_angle_deg * kPi / 180.0); a.sin_ud = sin(c->ud_angle_deg * kPi / 180.0);
  a.omgs = c->omgs; a.costh = c->costh; a.sinth = c->sinth; a.zt = c->zt;
  a.params = params; a.fe = fe; a.f1save = const_cast<double*>(static_cast<const double*>(saved));
  a.ff_bar = ff_bar; a.fe_bar = fe_bar; a.lgbar = static_cast<double*>(ws);
  a.fe_part = reinterpret_cast<double*>(static_cast<char*>(ws) + align_up((size_t)B * c->G * kLGDoubles * 8));
  TSFF_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)B * c->G * kLGDoubles * 8, st));
  const int GS = (V + 31) / 32 * 32, NG = kThreads2V / GS;
  const size_t smem = (((size_t)V * (V + 1) * 4 + 15) / 16) * 16 + ((size_t)V * (V + 1) + (size_t)NG * V * 2 + (size_t)(12 * V - 36) + (size_t)NG * 32) * 8;
  TSFF_CUDA_OK(cudaFuncSetAttribute(k_ff2v_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 228 * 1024 - 4096));
  TSFF_CUDA_OK(cudaFuncSetAttribute(k_ff2v_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 228 * 1024 - 4096));  // 224 KB at V = 128
  const long long nbatch = ((long long)a.P + NG - 1) / NG;
  const unsigned grid = (unsigned)(nbatch < c->sm_count ? nbatch : c->sm_count);
  for (int64_t b = 0; b < B; b++) {
    if (params_bar) k_ff2v_bwd<true><<<grid, kThreads2V, smem, st>>>(a, (long long)b);
    else k_ff2v_bwd<false><<<grid, kThreads2V, smem, st>>>(a, (long long)b);
    TSFF_LAUNCH_OK("k_ff2v_bwd");
    k_ff2v_reduce<<<(unsigned)((V * V + 255) / 256), 256, 0, st>>>(a.fe_part, (int)grid, V * V, fe_bar + b * (long long)V * V);
    TSFF_LAUNCH_OK("k_ff2v_reduce");
  }
  if (params_bar) {
    k_ff2v_params_bar<<<(unsigned)((B + 63) / 64), 64, 0, st>>>(a, params_bar, B);
    TSFF_LAUNCH_OK("k_ff2v_params_bar");
  }
  return TSFF_OK;
}
}  // namespace tsff
