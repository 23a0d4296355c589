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
// The adjoint (scatter of f1bar through the 16 taps into fbar, d/dbeta through the patch derivatives) is the next row.
#include "tsff_common.cuh"

using namespace tsff;

namespace {
constexpr int kThreads2V = 1024;

struct Args2V {
  int W, A, G, nI, V, NP, P;       // P = G*W*A poles per parameter set
  double lam_shift, v0, dv, cos_va, sin_va, cos_ud, sin_ud;
  const double *omgs, *costh, *sinth;
  ZTab zt;
  const double* params;   // [B][NP]
  const double* fe;       // [B][V][V]
  double* ff;             // [B][G][W][A]
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
template <int NGMAX>
__global__ void __launch_bounds__(kThreads2V, 1) k_ff2v_fwd(const Args2V a, long long b_lineout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int V = a.V, VP = V + 1;
  const int GS = (V + 31) / 32 * 32;           // threads per pole group
  const int NG = kThreads2V / GS;              // poles in flight per CTA
  double* sf = reinterpret_cast<double*>(smem_raw);
  double* sf1 = sf + (size_t)V * VP;
  double* sdf = sf1 + (size_t)NG * V;
  double* sred = sdf + (size_t)NG * V;         // [NG][8]
  double* spol = sred + (size_t)NG * 8;        // [NG][8]: beta-cos, beta-sin, |xi|, valid
  __shared__ LG sL[8];                          // G <= 8 gradient points
  const double* fe = a.fe + b_lineout * (long long)V * V;
  for (int i = threadIdx.x; i < V * V; i += kThreads2V) sf[(i / V) * VP + (i % V)] = fe[i];
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
    if (valid) {
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
      for (int aa = 0; aa < V; aa++) {
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
      IonOut io;
      ion_forward(L, a.nI, a.zt, q, io);
      Asm s;
      const double Pl = assemble_forward(L, q, io, -q.ikl2 * I, kPi * q.ikl2 * dfe, fphi, omgs, s);
      a.ff[((b_lineout * a.G + g) * (long long)a.W + j) * a.A + ia] = Pl;
    }
    __syncthreads();
  }
}
}  // namespace

namespace tsff {
int ff2v_fwd(tsff_ctx* c, int64_t B, const double* params, const double* fe, double* ff_out, cudaStream_t st) {
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
  a.params = params; a.fe = fe; a.ff = ff_out;
  const int GS = (V + 31) / 32 * 32, NG = kThreads2V / GS;
  const size_t smem = ((size_t)V * (V + 1) + (size_t)NG * V * 2 + (size_t)NG * 16) * 8;
  TSFF_SMEM_OPTIN(k_ff2v_fwd<32>);
  const long long nbatch = ((long long)a.P + NG - 1) / NG;
  const unsigned grid = (unsigned)(nbatch < c->sm_count ? nbatch : c->sm_count);
  for (int64_t b = 0; b < B; b++) {
    k_ff2v_fwd<32><<<grid, kThreads2V, smem, st>>>(a, (long long)b);
    TSFF_LAUNCH_OK("k_ff2v_fwd");
  }
  return TSFF_OK;
}
}  // namespace tsff
