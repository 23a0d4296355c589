// tsff_table.cu -- TSFF_MODE_TABLE: the reference's 1V FormFactor.__call__ (form_factor.py:163-298).
//
//   f(v) --log, cubic Hermite--> ratmod on xi1 (1024) --gradient--> ratdf --ratintn--> PV table T on xi2 (1640)
//   per (omega, angle): kinematics, chi_i from Z', fphi = exp(cubic(log f))(xi_e), Im chi_e from the forward
//   difference of fphi along omega, Re chi_e = -lerp(T)(xi_e)/(k lambda_De)^2, S(k, omega), angle sum.
//
// Kernels
//   k_table_prep        per lineout (FP64): LG scalars, log f, node slopes, ratmod, ratdf, PV weights D
//   k_pv_poles          (tsff_pv_kernels.cuh) PV table on the fixed pole grid xi2           [MUFU-bound]
//   k_table_fwd         per-(omega, angle) assembly in FP64; a warp owns 31 consecutive wavelengths so the forward
//                       difference along omega is a register shuffle; angle sum in-thread        [FP64-pipe-bound]
//   k_table_bwd         hand-written reverse of k_table_fwd (30 outputs + 2 halo lanes per warp)
//   k_table_tbar        per lineout: Tbar -> pole descriptors + endpoint cotangents
//   k_pv_nodes          adjoint PV sweep -> Dbar                                                [MUFU-bound]
//   k_table_bwd_finish  per lineout: Dbar -> ratdf_bar -> ratmod_bar -> (log f, slope)_bar -> fe_bar; params_bar
#include "tsff_pv_kernels.cuh"

using namespace tsff;

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct TableLayout {
  size_t s_lg, s_lnf, s_slope, s_ratmod, s_T, saved_bytes;
  size_t w_mpart, w_D, w_D64, w_pend, w_ratdf, w_desc, w_Dbar, w_zero_begin, w_Tbar, w_lnfbar, w_slopebar, w_pnear, w_lgbar, w_zero_end,
      ws_bytes;
};

// forward with the fused angle sum: few lineouts leave most SMs idle at one CTA per (lineout, wavelength tile); the angles are then
// split over CTAs too, each writing its partial sum, and a second small kernel adds the partials in chunk order (deterministic)
constexpr int kFwdJ = 31;  // outputs per warp (lane 31 is the right halo)
inline bool table_fwd_split_angles(const tsff_ctx* c, int64_t B) {
  const long long tiles = (c->W + (256 / 32) * kFwdJ - 1) / ((256 / 32) * kFwdJ);
  return c->A > 1 && B * tiles < 2LL * c->sm_count;
}

TableLayout table_layout(const tsff_ctx* c, int64_t B) {
  TableLayout L;
  size_t o = 0;
  L.s_lg = o; o += align_up((size_t)B * c->G * kLGDoubles * 8);
  L.s_lnf = o; o += align_up((size_t)B * c->V * 8);
  L.s_slope = o; o += align_up((size_t)B * c->V * 8);
  L.s_ratmod = o; o += align_up((size_t)B * kXi1N * 8);
  L.s_T = o; o += align_up((size_t)B * kXi2N * 8);
  L.saved_bytes = o;
  o = 0;
  L.w_D = o; o += align_up((size_t)B * tree_blob(c->pv_npad).bytes);
  L.w_D64 = o; o += align_up(c->pv_precision == TSFF_PV_FP64 ? (size_t)B * c->pv_npad * 8 : 0);
  L.w_pend = o; o += align_up((size_t)B * 2 * 8);
  L.w_ratdf = o; o += align_up((size_t)B * kXi1N * 8);
  L.w_desc = o; o += align_up((size_t)B * kXi2N * 16);
  L.w_zero_begin = o;
  L.w_Dbar = o; o += align_up((size_t)B * c->pv_npad * 8);
  L.w_Tbar = o; o += align_up((size_t)B * kXi2N * 8);
  L.w_lnfbar = o; o += align_up((size_t)B * c->V * 8);
  L.w_slopebar = o; o += align_up((size_t)B * c->V * 8);
  L.w_pnear = o; o += align_up((size_t)B * kXi1N * 8);
  L.w_lgbar = o; o += align_up((size_t)B * c->G * kLGDoubles * 8);
  L.w_zero_end = o;
  // last: the only region whose size depends on (A, W) -- the pair path addresses one context's workspace with the other's layout
  L.w_mpart = o; o += align_up(table_fwd_split_angles(c, B) ? (size_t)c->A * B * c->W * 8 : 0);
  L.ws_bytes = o;
  return L;
}

struct TableArgs {
  int W, A, G, nI, V, NP, nodes, npad, ntiles;
  int jrep;     // forward: groups of 31 wavelengths a warp takes in turn
  int kper;     // backward: consecutive wavelengths per thread
  int stage_z;  // backward: Z' table staged in shared memory (pays once a CTA evaluates a few thousand points)
  int asplit;   // angle chunks per wavelength tile (ARTS: one lineout, 241 angles -- the grid would not fill the device otherwise)
  double lam_shift, v0, dv, xi1_0, xi1_h, xi2_0, xi2_h;
  const double *omgs, *costh, *wts, *jmul, *xi2;
  ZTab zt;
  const double* params;
  const void* fe;
  double *lg, *lnf, *slope, *ratmod, *T;
  unsigned char* D;  // per-lineout tree blobs
  const double* tstat;
  double* D64;
  double* pend;
  double* ratdf;
  double* modl;
  double* mpart;   // forward, angles split over CTAs: partial angle sums [asplit][B][W] (then k_table_modl_reduce), else null
  long long Bn;    // lineouts of the call
  double* ff;
  // backward
  const double* modl_bar;
  const double* ff_bar;
  float4* desc;
  double *Tbar, *lnfbar, *slopebar, *pnear, *Dbar, *lgbar;
  const double* lgbar2;   // pair path: the second window's LG cotangents (its params_bar is added by this window's finish), else null
  double lam_shift2;
  double* params_bar;
  void* fe_bar;
  int* cells;      // frozen lerp cells [B][G][W][A][kCellStride] (second-order path) or null
  int cell_mode;   // 0 off, 1 record, 2 replay
};

__device__ __forceinline__ void load_lg(const double* src, LG& L) {
  L.ne_g = src[0]; L.omgL = src[1]; L.omgpe2 = src[2]; L.kL = src[3]; L.vTe = src[4]; L.Va6 = src[5]; L.ud6 = src[6];
#pragma unroll
  for (int i = 0; i < TSFF_MAX_IONS; i++) {
    L.c_kldi[i] = src[7 + i]; L.inv_s2vTi[i] = src[7 + TSFF_MAX_IONS + i]; L.ioncf[i] = src[7 + 2 * TSFF_MAX_IONS + i];
  }
}
__device__ __forceinline__ void store_lg(double* dst, const LG& L) {
  dst[0] = L.ne_g; dst[1] = L.omgL; dst[2] = L.omgpe2; dst[3] = L.kL; dst[4] = L.vTe; dst[5] = L.Va6; dst[6] = L.ud6;
#pragma unroll
  for (int i = 0; i < TSFF_MAX_IONS; i++) {
    dst[7 + i] = L.c_kldi[i]; dst[7 + TSFF_MAX_IONS + i] = L.inv_s2vTi[i]; dst[7 + 2 * TSFF_MAX_IONS + i] = L.ioncf[i];
  }
}

// np.gradient with uniform spacing at node i of an n-array held in shared memory
__device__ __forceinline__ double grad_sm(const double* f, int n, double ih, int i) {
  if (i <= 0) return (f[1] - f[0]) * ih;
  if (i >= n - 1) return (f[n - 1] - f[n - 2]) * ih;
  return (f[i + 1] - f[i - 1]) * (0.5 * ih);
}

// exp of an interpolated log f: log f <= 0 for every normalised table the reference produces (fast path, ~1 ulp, 20
// instructions against ~45 for the library exp); anything else goes through exp()
__device__ __forceinline__ double exp_logf(double H) { return H <= 0.0 ? fast_exp_neg(H) : exp(H); }

// ---- prep -------------------------------------------------------------------------------------------------
// dynamic smem: lnf[V] | slope[V] | ratmod[1024] | ratdf[1024] | tree_prep scratch
template <typename T>
__global__ void __launch_bounds__(kThreads) k_table_prep(const TableArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_lnf = reinterpret_cast<double*>(smem_raw);
  double* s_slope = s_lnf + a.V;
  double* s_rat = s_slope + a.V;
  double* s_p = s_rat + kXi1N;
  const long long b = blockIdx.x;
  const T* fe = static_cast<const T*>(a.fe) + b * a.V;
  if (threadIdx.x < a.G) {
    LG L;
    lg_zero(L);
    lg_forward(a.params + b * a.NP, a.nI, threadIdx.x, a.G, a.lam_shift, L);
    store_lg(a.lg + (b * a.G + threadIdx.x) * kLGDoubles, L);
  }
  for (int i = threadIdx.x; i < a.V; i += kThreads) s_lnf[i] = log((double)fe[i]);   // form_factor.py:256,263
  __syncthreads();
  const double idv = 1.0 / a.dv;
  for (int i = threadIdx.x; i < a.V; i += kThreads) {
    // interpax cubic node slopes: mean of adjacent secants, one-sided at the ends (= np.gradient on a uniform grid)
    double s = grad_sm(s_lnf, a.V, idv, i);
    s_slope[i] = s;
    a.lnf[b * a.V + i] = s_lnf[i];
    a.slope[b * a.V + i] = s;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < kXi1N; n += kThreads) {
    Herm hm;
    double H = hermite_uniform(s_lnf, s_slope, a.V, a.v0, a.dv, a.xi1_0 + (double)n * a.xi1_h, kFillLog, hm);
    double r = exp(H);                                                                // form_factor.py:263
    s_rat[n] = r;
    a.ratmod[b * kXi1N + n] = r;
  }
  __syncthreads();
  const double ih = 1.0 / a.xi1_h;
  for (int n = threadIdx.x; n < kXi1N; n += kThreads) s_p[n] = grad_sm(s_rat, kXi1N, ih, n);  // form_factor.py:264
  __syncthreads();
  const int M = a.nodes - 1;
  tree_prep_cta(s_p, M, a.npad, a.D + b * tree_blob(a.npad).bytes, a.tstat, s_p + kXi1N);
  for (int i = threadIdx.x; i < kXi1N; i += kThreads) {
    if (a.D64) a.D64[b * a.npad + i] = pv_weight(s_p, M, a.xi1_h, i);                  // FP64 validation path
    a.ratdf[b * kXi1N + i] = s_p[i];
  }
  if (threadIdx.x == 0) {
    a.pend[2 * b] = s_p[0];
    a.pend[2 * b + 1] = s_p[M];
  }
}

// ---- forward assembly -----------------------------------------------------------------------------------------
// dynamic smem: lnf[V] | slope[V] | T[1640]
// FROZEN: the second-order path's cell record / replay (tsff_ctx_set_frozen_cells); a template so that the normal path carries
// none of it
#ifndef TSFF_TFWD_MINB
#define TSFF_TFWD_MINB 4
#endif
template <bool WRITE_FF, bool FROZEN = false>
__global__ void __launch_bounds__(kThreads, TSFF_TFWD_MINB) k_table_fwd(const TableArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LG sL;   // CTA-uniform scalars of the (lineout, gradient point)
  __shared__ LGX sX;  // ... and their reciprocals
  double* s_lnf = reinterpret_cast<double*>(smem_raw);
  double* s_slope = s_lnf + a.V;
  double* s_T = s_slope + a.V;
  const int chunk = blockIdx.x % a.asplit;
  const int tile = (blockIdx.x / a.asplit) % a.ntiles;
  const long long b = blockIdx.x / ((long long)a.asplit * a.ntiles);
  const int aper = (a.A + a.asplit - 1) / a.asplit, ia0 = chunk * aper, ia1 = min(a.A, ia0 + aper);
  for (int i = threadIdx.x; i < a.V; i += kThreads) {
    s_lnf[i] = a.lnf[b * a.V + i];
    s_slope[i] = a.slope[b * a.V + i];
  }
  for (int i = threadIdx.x; i < kXi2N; i += kThreads) s_T[i] = a.T[b * kXi2N + i];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double idv = 1.0 / a.dv, ih2 = 1.0 / a.xi2_h;
  // a warp takes a.jrep groups of 31 wavelengths in turn: the staged tables (18 KB per CTA) are then paid once per jrep x 248
  // wavelengths (at one group per warp the staging and its barrier held ~20 % of the kernel's warp-time)
  for (int rep = 0; rep < a.jrep; rep++) {
    const int j = ((tile * a.jrep + rep) * kWarps + wid) * kFwdJ + lane;
    const int jc = min(j, a.W - 1);
    const bool out = (lane < kFwdJ) && (j < a.W);
    const double omgs = a.omgs[jc];
    double acc = 0.0;
    for (int g = 0; g < a.G; g++) {
      if (rep == 0 || a.G > 1) {
        __syncthreads();
        if (threadIdx.x < kLGDoubles) reinterpret_cast<double*>(&sL)[threadIdx.x] = a.lg[(b * a.G + g) * kLGDoubles + threadIdx.x];
        __syncthreads();
        if (threadIdx.x < kLGXDoubles) reinterpret_cast<double*>(&sX)[threadIdx.x] = lgx_field(sL, a.nI, a.zt.h, threadIdx.x);
        __syncthreads();
      }
      const LG& L = sL;
      const LGX& X = sX;
      for (int ia = ia0; ia < ia1; ia++) {
        KinX q;
        kin_forward_x(L, X, omgs, a.costh[ia], q);
        Herm hm;
        const double fphi = exp_logf(hermite_uniform_ih(s_lnf, s_slope, a.V, a.v0, a.dv, idv, q.xie, kFillLog, hm));  // :256
        const double xi_n = __shfl_down_sync(0xffffffffu, q.xie, 1);
        const double fphi_n = __shfl_down_sync(0xffffffffu, fphi, 1);
        const double df = (j + 1 < a.W && lane < 31) ? (fphi_n - fphi) * fast_rcp(xi_n - q.xie) : 0.0;   // :258-259
        if (out) {
          int ip; double tp, slp;
          int* cp = FROZEN ? a.cells + ((((b * a.G + g) * (long long)a.W + j) * a.A + ia) * kCellStride) : nullptr;
          const int cm = FROZEN ? a.cell_mode : 0;
          const double Tl = FROZEN ? lerp_uniform_cell(s_T, kXi2N, a.xi2_0, a.xi2_h, q.xie, ip, tp, slp, cm, cp)
                                   : lerp_uniform_ih(s_T, kXi2N, a.xi2_0, ih2, q.xie, ip, tp, slp);       // :270
          const double chiEr = -q.ikl2 * Tl;                                                                // :271
          const double chiEi = kPi * q.ikl2 * df;                                                           // :261
          IonX io;
          ion_forward_x<0>(L, X, a.nI, a.zt, q, io, cm, cp + 1);
          AsmX s;
          const double P = assemble_forward_x(L, X, q, io, chiEr, chiEi, fphi, omgs, s);
          if (WRITE_FF) a.ff[((b * a.G + g) * (long long)a.W + j) * a.A + ia] = P;
          acc += a.wts[ia] * P;
        }
      }
    }
    if (out && a.mpart) a.mpart[((long long)chunk * a.Bn + b) * a.W + j] = acc;
    else if (out && a.modl) a.modl[b * a.W + j] = a.jmul[j] * acc / (double)a.G;
  }
}

__global__ void __launch_bounds__(kThreads) k_table_modl_reduce(const TableArgs a, long long total) {
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= total) return;
  double s = 0.0;
  for (int ch = 0; ch < a.asplit; ch++) s += a.mpart[(long long)ch * total + t];
  a.modl[t] = a.jmul[t % a.W] * s / (double)a.G;
}

// ---- backward assembly ----------------------------------------------------------------------------------------
// A thread owns `kper` CONSECUTIVE wavelengths of one (lineout, angle chunk) and walks them in order.  The forward difference
// along omega (form_factor.py:258-259) couples point j to j + 1: the thread looks one point ahead (kinematics + f(xi) only) and
// carries d loss / d df_j to the next point in two registers; the share of its last forward difference that belongs to the first
// point of the NEXT thread is applied by this thread itself (the reverse is linear in the cotangents, and the look-ahead already
// holds that point's kinematics), so threads exchange nothing.  The scatter targets -- the T-table cell and the log-f Hermite cell
// -- change slowly along omega: each is accumulated in registers over the run of equal cell and flushed with atomics when the
// cell changes (a move to the adjacent cell keeps the shared node's partial sum in registers); when a whole warp ends on the
// same cell (the ion-acoustic window: thousands of wavelengths inside a few cells) the final flush is reduced across the warp.
#ifndef TSFF_TBWD_MINB
#define TSFF_TBWD_MINB 2
#endif
#ifndef TSFF_TBWD_K
#define TSFF_TBWD_K 4          // preferred wavelengths per thread (the launch picks 1 .. 2x this, see table_bwd_t)
#endif
#ifndef TSFF_TBWD_THREADS
#define TSFF_TBWD_THREADS 256
#endif
#ifndef TSFF_TBWD_RECOMP_KIN
#define TSFF_TBWD_RECOMP_KIN 1  // 1: carry only (xie, f(xie), Hermite cell) of the look-ahead point and redo its kinematics
#endif
constexpr int kBwdThreads = TSFF_TBWD_THREADS;

__device__ __forceinline__ void run_flush2(double* dst, int k, double v0, double v1) {
  if (k < 0) return;
  if (v0 != 0.0) atomicAdd(&dst[k], v0);
  if (v1 != 0.0) atomicAdd(&dst[k + 1], v1);
}
// accumulate (w0, w1) onto cells (k, k + 1) of dst, run-length compressed in (rk, r0, r1)
__device__ __forceinline__ void run_add2(double* dst, int& rk, double& r0, double& r1, int k, double w0, double w1) {
  if (k != rk) {
    if (k == rk + 1) {            // moved one cell up: node rk + 1 stays in registers
      if (r0 != 0.0) atomicAdd(&dst[rk], r0);
      r0 = r1; r1 = 0.0;
    } else if (k == rk - 1) {     // one cell down: node rk stays
      if (r1 != 0.0) atomicAdd(&dst[rk + 1], r1);
      r1 = r0; r0 = 0.0;
    } else {
      run_flush2(dst, rk, r0, r1);
      r0 = r1 = 0.0;
    }
    rk = k;
  }
  r0 += w0;
  r1 += w1;
}
// final flush: one lane issues the atomics when the whole warp ended on the same cell
__device__ __forceinline__ void run_flush2_warp(double* dst, int k, double v0, double v1) {
  const int k0 = __shfl_sync(0xffffffffu, k, 0);
  if (__all_sync(0xffffffffu, k == k0)) {
    v0 = warp_sum(v0);
    v1 = warp_sum(v1);
    if ((threadIdx.x & 31) == 0) run_flush2(dst, k, v0, v1);
  } else {
    run_flush2(dst, k, v0, v1);
  }
}

// dynamic smem: lnf[V] | slope[V] | T[1640] | Z'[1640] as (re, im) pairs | omgs of the CTA's wavelengths (+1) | modl_bar * jmul of them
__host__ __device__ inline size_t table_bwd_smem(int V, int kper) {
  return (size_t)(2 * V + kXi2N + 2 * kXi2N + 2 * (kBwdThreads * kper + 1)) * 8;
}

template <bool FROZEN, int NI>
__global__ void __launch_bounds__(kBwdThreads, TSFF_TBWD_MINB) k_table_bwd(const TableArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sred[kLGDoubles * (kBwdThreads / 32)];
  __shared__ LG sL;   // the (lineout, gradient point) scalars are CTA-uniform: shared, not registers
  __shared__ LGX sX;
  __shared__ int s_next;
  double* s_lnf = reinterpret_cast<double*>(smem_raw);
  double* s_slope = s_lnf + a.V;
  double* s_T = s_slope + a.V;
  ZZ* s_zz = reinterpret_cast<ZZ*>(s_T + kXi2N);
  double* s_omg = reinterpret_cast<double*>(s_zz + kXi2N);
  const int span = kBwdThreads * a.kper + 1;   // wavelengths of the CTA plus the look-ahead of its last one
  double* s_pm = s_omg + span;
  const int chunk = blockIdx.x % a.asplit;
  const int tile = (blockIdx.x / a.asplit) % a.ntiles;
  const long long b = blockIdx.x / ((long long)a.asplit * a.ntiles);
  const int aper = (a.A + a.asplit - 1) / a.asplit, ia0 = chunk * aper, ia1 = min(a.A, ia0 + aper);
  const int jbase = tile * kBwdThreads * a.kper;
  const double iG = 1.0 / (double)a.G;
  for (int i = threadIdx.x; i < a.V; i += kBwdThreads) {
    s_lnf[i] = a.lnf[b * a.V + i];
    s_slope[i] = a.slope[b * a.V + i];
  }
  for (int i = threadIdx.x; i < kXi2N; i += kBwdThreads) {
    s_T[i] = a.T[b * kXi2N + i];
    if (a.stage_z) {
      s_zz[i].r = a.zt.zr[i];
      s_zz[i].i = a.zt.zi[i];
    }
  }
  for (int i = threadIdx.x; i < span; i += kBwdThreads) {
    const int j = min(jbase + i, a.W - 1);
    s_omg[i] = a.omgs[j];
    s_pm[i] = a.modl_bar ? a.modl_bar[b * a.W + j] * a.jmul[j] * iG : 0.0;
  }
  const int lane = threadIdx.x & 31;
  constexpr int kNW = kBwdThreads / 32;
  const double idv = 1.0 / a.dv, ih2 = 1.0 / a.xi2_h;
  double* Tbar = a.Tbar + b * kXi2N;
  double* lnfbar = a.lnfbar + b * a.V;
  double* slopebar = a.slopebar + b * a.V;
  // work items of the CTA: (warp-sized chunk of wavelength runs, angle), drawn by the warps from a shared counter
  const int nitems = kNW * (ia1 - ia0);
  for (int g = 0; g < a.G; g++) {
    __syncthreads();   // first pass: the staged tables; later passes: everyone is done with the previous sL
    if (threadIdx.x < kLGDoubles) reinterpret_cast<double*>(&sL)[threadIdx.x] = a.lg[(b * a.G + g) * kLGDoubles + threadIdx.x];
    if (threadIdx.x == 32) s_next = 0;
    __syncthreads();
    if (threadIdx.x < kLGXDoubles) reinterpret_cast<double*>(&sX)[threadIdx.x] = lgx_field(sL, a.nI, a.zt.h, threadIdx.x);
    __syncthreads();
    const LG& L = sL;
    const LGX& X = sX;
    LG Lb;
    lg_zero(Lb);
    for (;;) {
      int item = 0;
      if (lane == 0) item = atomicAdd(&s_next, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= nitems) break;
      const int ia = ia0 + item / kNW;
      const int j0 = jbase + ((item % kNW) * 32 + lane) * a.kper;
      const int jend = min(j0 + a.kper, a.W);
      const double cth = a.costh[ia], wt = a.wts[ia];
      int kT = -1, kH = -1;                                   // open runs: T cell (nodes kT, kT + 1); Hermite cell (nodes kH - 1, kH)
      double tb0 = 0.0, tb1 = 0.0, cf0 = 0.0, cf1 = 0.0, cm0 = 0.0, cm1 = 0.0;
      if (j0 < a.W) {
        KinX q;
        kin_forward_x(L, X, s_omg[j0 - jbase], cth, q);
        Herm hm;
        double fphi = fast_exp_bf(hermite_uniform_bf(s_lnf, s_slope, a.V, a.v0, a.dv, idv, q.xie, kFillLog, hm));
        double carry_f = 0.0, carry_x = 0.0;                  // d loss / d (fphi_j, xie_j) through df_{j-1}
        for (int j = j0; j < jend; j++) {
          const double omgs = s_omg[j - jbase];
          // look one point ahead for the forward difference (form_factor.py:258-259); the last wavelength has none: the staged
          // omgs repeats it, the difference quotient is discarded
          const bool has_df = j + 1 < a.W;
          KinX qn;
          kin_forward_x(L, X, s_omg[j + 1 - jbase], cth, qn);
          Herm hn;
          const double fphi_n = fast_exp_bf(hermite_uniform_bf(s_lnf, s_slope, a.V, a.v0, a.dv, idv, qn.xie, kFillLog, hn));
          const double idelta = has_df ? fast_rcp(qn.xie - q.xie) : 0.0;
          const double df = has_df ? (fphi_n - fphi) * idelta : 0.0;
          // reverse of the assembly at point j
          double Pbar = s_pm[j - jbase] * wt;
          if (a.ff_bar) Pbar += a.ff_bar[((b * a.G + g) * (long long)a.W + j) * a.A + ia];
          int ip; double tp, slp, Tl;
          IonX io;
          if (FROZEN) {
            int* cp = a.cells + ((((b * a.G + g) * (long long)a.W + j) * a.A + ia) * kCellStride);
            Tl = lerp_uniform_cell(s_T, kXi2N, a.xi2_0, a.xi2_h, q.xie, ip, tp, slp, 2, cp);   // the adjoint re-uses the recorded cells
            ion_forward_x<NI>(L, X, a.nI, a.zt, q, io, 2, cp + 1);
          } else {
            Tl = lerp_uniform_bf(s_T, kXi2N, a.xi2_0, ih2, q.xie, ip, tp, slp);
            if (a.stage_z) ion_forward_bf<NI>(L, X, a.nI, (const ZZ*)s_zz, a.zt, q, io);
            else ion_forward_bf<NI>(L, X, a.nI, ZRows{a.zt.zr, a.zt.zi}, a.zt, q, io);
          }
          const double chiEr = -q.ikl2 * Tl, chiEi = kPi * q.ikl2 * df;
          AsmX s;
          assemble_forward_x(L, X, q, io, chiEr, chiEi, fphi, omgs, s);
          PointBar pb;
          KinBar kb = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
          assemble_backward_x<NI>(L, X, a.nI, q, io, chiEr, chiEi, fphi, s, Pbar, pb, kb, Lb);
          const double dfbar = has_df ? kPi * q.ikl2 * pb.chiEi : 0.0;
          kb.ikl2 += -Tl * pb.chiEr + kPi * df * pb.chiEi;
          const double gd = dfbar * idelta;
          const double fphibar = pb.fphi - gd + carry_f;
          double xiebar = kb.xie + gd * df + carry_x;
          carry_f = gd;
          carry_x = -gd * df;
          const double Tlbar = -q.ikl2 * pb.chiEr;
          xiebar += Tlbar * slp;
          const double Hbar = hm.inside ? fphibar * fphi : 0.0;
          xiebar += Hbar * hm.dHdx;
          double wf0, wf1, wm0, wm1;
          hermite_weights(hm.t, a.dv, wf0, wf1, wm0, wm1);
          kb.xie = xiebar;
          kin_backward_x(L, X, cth, q, kb, Lb);
          run_add2(Tbar, kT, tb0, tb1, ip, (1.0 - tp) * Tlbar, tp * Tlbar);
          if (hm.inside) {
            int kk = kH;
            run_add2(lnfbar - 1, kk, cf0, cf1, hm.i, Hbar * wf0, Hbar * wf1);
            run_add2(slopebar - 1, kH, cm0, cm1, hm.i, Hbar * wm0, Hbar * wm1);
          }
#if TSFF_TBWD_RECOMP_KIN
          kin_forward_x(L, X, s_omg[j + 1 - jbase], cth, q);
#else
          q = qn;
#endif
          hm = hn; fphi = fphi_n;
        }
        // the point after this thread's last one (q holds it) takes its share of the last forward difference
        if (jend < a.W) {
          double xiebar = carry_x;
          if (hm.inside) {
            const double Hbar = carry_f * fphi;
            xiebar += Hbar * hm.dHdx;
            double wf0, wf1, wm0, wm1;
            hermite_weights(hm.t, a.dv, wf0, wf1, wm0, wm1);
            int kk = kH;
            run_add2(lnfbar - 1, kk, cf0, cf1, hm.i, Hbar * wf0, Hbar * wf1);
            run_add2(slopebar - 1, kH, cm0, cm1, hm.i, Hbar * wm0, Hbar * wm1);
          }
          KinBar kb = {0.0, 0.0, 0.0, 0.0, xiebar, 0.0};
          kin_backward_x(L, X, cth, q, kb, Lb);
        }
      }
      // all 32 lanes (idle ones with empty runs) take part in the final flush
      run_flush2_warp(Tbar, kT, tb0, tb1);
      run_flush2_warp(lnfbar - 1, kH, cf0, cf1);
      run_flush2_warp(slopebar - 1, kH, cm0, cm1);
    }
    double vals[kLGDoubles];
    store_lg(vals, Lb);
    block_accumulate<kBwdThreads / 32>(vals, kLGDoubles, sred, a.lgbar + (b * a.G + g) * kLGDoubles);
  }
}

// ---- Tbar -> descriptors for the far-field sweep + exact near / endpoint contributions ---------------------------
__global__ void __launch_bounds__(kThreads) k_table_tbar(const TableArgs a, int nsplit) {
  const long long b = blockIdx.x / nsplit;
  const int per = (kXi2N + nsplit - 1) / nsplit, p0 = (blockIdx.x % nsplit) * per, p1 = min(kXi2N, p0 + per);
  for (int p = p0 + threadIdx.x; p < p1; p += kThreads) {
    const double xi = a.xi2[p];
    const double tb = a.Tbar[b * kXi2N + p];
    int np, wb0;
    a.desc[b * kXi2N + p] = pv_desc(xi, tb, a.xi1_0, a.xi1_h, a.nodes, a.npad, np, wb0);
    pv_bwd_pole_exact(xi, tb, a.xi1_0, a.xi1_h, a.nodes, np, wb0, a.pnear + b * kXi1N);
  }
}

// ---- finish ---------------------------------------------------------------------------------------------------
// dynamic smem: pb[1024] | hb[1024] | lnfb[V] | slb[V]
template <typename T>
__global__ void __launch_bounds__(kThreads) k_table_bwd_finish(const TableArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_pb = reinterpret_cast<double*>(smem_raw);
  double* s_hb = s_pb + kXi1N;
  double* s_lnfb = s_hb + kXi1N;
  double* s_slb = s_lnfb + a.V;
  const long long b = blockIdx.x;
  const int V = a.V, M = a.nodes - 1;
  const double* pfar = a.Dbar + b * a.npad;
  const double ih = 1.0 / a.xi1_h;
  // 1. ratdf_bar = far-field sweep + exact near/endpoint terms
  for (int i = threadIdx.x; i < kXi1N; i += kThreads) s_pb[i] = (i <= M ? pfar[i] : 0.0) + a.pnear[b * kXi1N + i];
  __syncthreads();
  // 2. ratmod_bar (adjoint of np.gradient) and Hbar_n = ratmod_bar_n * ratmod_n inside the f grid
  const double xlast = a.v0 + (double)(V - 1) * a.dv;
  for (int k = threadIdx.x; k < kXi1N; k += kThreads) {
    double rb = 0.0;
    if (k >= 1) rb += s_pb[k - 1] * ((k - 1 == 0) ? ih : 0.5 * ih);
    if (k <= kXi1N - 2) rb -= s_pb[k + 1] * ((k + 1 == kXi1N - 1) ? ih : 0.5 * ih);
    if (k == 0) rb -= s_pb[0] * ih;
    if (k == kXi1N - 1) rb += s_pb[kXi1N - 1] * ih;
    const double x = a.xi1_0 + (double)k * a.xi1_h;
    const bool inside = !(x < a.v0 || x > xlast);
    s_hb[k] = inside ? rb * a.ratmod[b * kXi1N + k] : 0.0;
  }
  __syncthreads();
  // 3. gather the Hermite adjoint at the xi1 points onto the f nodes (deterministic, no atomics)
  for (int k = threadIdx.x; k < V; k += kThreads) {
    double lf = a.lnfbar[b * V + k], sl = a.slopebar[b * V + k];
    // xi1 points whose cell [i-1, i] touches node k lie in [v_{k-1}, v_{k+1}]
    const double xlo = a.v0 + (double)(k - 1) * a.dv, xhi = a.v0 + (double)(k + 1) * a.dv;
    int nlo = (int)floor((xlo - a.xi1_0) / a.xi1_h) - 1, nhi = (int)ceil((xhi - a.xi1_0) / a.xi1_h) + 1;
    nlo = max(nlo, 0); nhi = min(nhi, kXi1N - 1);
    for (int n = nlo; n <= nhi; n++) {
      const double hb = s_hb[n];
      if (hb == 0.0) continue;
      const double x = a.xi1_0 + (double)n * a.xi1_h;
      int i = (int)floor((x - a.v0) / a.dv) + 1;
      i = min(max(i, 1), V - 1);
      if (k != i - 1 && k != i) continue;
      const double t = (x - (a.v0 + (double)(i - 1) * a.dv)) / a.dv;
      double wf0, wf1, wm0, wm1;
      hermite_weights(t, a.dv, wf0, wf1, wm0, wm1);
      if (k == i - 1) { lf += hb * wf0; sl += hb * wm0; }
      else            { lf += hb * wf1; sl += hb * wm1; }
    }
    s_lnfb[k] = lf;
    s_slb[k] = sl;
  }
  __syncthreads();
  // 4. slope adjoint (np.gradient stencil on log f), then fe_bar = lnf_bar / fe
  const double idv = 1.0 / a.dv;
  const T* fe = static_cast<const T*>(a.fe) + b * V;
  T* fe_bar = static_cast<T*>(a.fe_bar) + b * V;
  for (int k = threadIdx.x; k < V; k += kThreads) {
    double lb = s_lnfb[k];
    if (k >= 1) lb += s_slb[k - 1] * ((k - 1 == 0) ? idv : 0.5 * idv);
    if (k <= V - 2) lb -= s_slb[k + 1] * ((k + 1 == V - 1) ? idv : 0.5 * idv);
    if (k == 0) lb -= s_slb[0] * idv;
    if (k == V - 1) lb += s_slb[V - 1] * idv;
    fe_bar[k] = (T)(lb / (double)fe[k]);
  }
  // params_bar: this window's LG cotangents (thread 0) and, on the pair path, the second window's (thread 32), each reversed into its
  // own accumulator, then added -- one launch, two short serial chains side by side
  __shared__ double spb[2][10 + 4 * TSFF_MAX_IONS];
  if (threadIdx.x == 0 || (threadIdx.x == 32 && a.lgbar2)) {
    const int w = threadIdx.x == 0 ? 0 : 1;
    double pb[10 + 4 * TSFF_MAX_IONS];
    for (int k = 0; k < a.NP; k++) pb[k] = 0.0;
    for (int g = 0; g < a.G; g++) {
      LG Lb;
      load_lg((w ? a.lgbar2 : a.lgbar) + (b * a.G + g) * kLGDoubles, Lb);
      lg_backward(a.params + b * a.NP, a.nI, g, a.G, w ? a.lam_shift2 : a.lam_shift, Lb, pb);
    }
    for (int k = 0; k < a.NP; k++) spb[w][k] = pb[k];
  }
  __syncthreads();
  if (threadIdx.x < a.NP) a.params_bar[b * a.NP + threadIdx.x] = spb[0][threadIdx.x] + (a.lgbar2 ? spb[1][threadIdx.x] : 0.0);
}

// ---- two windows of one plasma (pair path) ----------------------------------------------------------------------
// LG scalars alone (the second window re-uses the first one's f-dependent tables, so k_table_prep does not run for it)
__global__ void __launch_bounds__(128) k_table_lg(const TableArgs a, long long BG) {
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  if (t >= BG) return;
  LG L;
  lg_zero(L);
  lg_forward(a.params + (t / a.G) * a.NP, a.nI, (int)(t % a.G), a.G, a.lam_shift, L);
  store_lg(a.lg + t * kLGDoubles, L);
}
// angle chunks per wavelength tile: 1 when the (lineout, tile) grid alone gives two CTAs per SM, else enough to get there
int table_angle_split(long long ctas, int A, int sm_count) {
  if (ctas >= 2LL * sm_count) return 1;
  long long want = (2LL * sm_count + ctas - 1) / ctas;
  if (want > A) want = A;
  return want < 1 ? 1 : (int)want;
}

void fill_static(const tsff_ctx* c, TableArgs& a) {
  a.cells = c->cells; a.cell_mode = c->cell_mode;
  a.W = c->W; a.A = c->A; a.G = c->G; a.nI = c->I; a.V = c->V; a.NP = c->NP;
  a.nodes = c->pv_nodes; a.npad = c->pv_npad;
  a.lam_shift = c->lam_shift; a.v0 = c->v0; a.dv = c->dv;
  a.xi1_0 = c->xi1_0; a.xi1_h = c->xi1_h; a.xi2_0 = c->zt.x0; a.xi2_h = c->zt.h;
  a.omgs = c->omgs; a.costh = c->costh; a.wts = c->wts; a.jmul = c->jmul; a.xi2 = c->xi2; a.zt = c->zt;
  a.tstat = c->tstat;
}

void bind_saved(const TableLayout& L, char* sv, TableArgs& a) {
  a.lg = (double*)(sv + L.s_lg); a.lnf = (double*)(sv + L.s_lnf); a.slope = (double*)(sv + L.s_slope);
  a.ratmod = (double*)(sv + L.s_ratmod); a.T = (double*)(sv + L.s_T);
}

template <typename T>
int table_fwd_t(tsff_ctx* c, int64_t B, const double* params, const void* fe, double* modl_out, double* ff_out, void* saved,
                void* ws, cudaStream_t st, const void* tables_from = nullptr) {
  const TableLayout L = table_layout(c, B);
  char* w = static_cast<char*>(ws);
  TableArgs a;
  memset(&a, 0, sizeof(a));
  fill_static(c, a);
  bind_saved(L, static_cast<char*>(saved), a);
  a.params = params; a.fe = fe;
  a.D = (unsigned char*)(w + L.w_D);
  a.D64 = c->pv_precision == TSFF_PV_FP64 ? (double*)(w + L.w_D64) : nullptr;
  a.pend = (double*)(w + L.w_pend); a.ratdf = (double*)(w + L.w_ratdf);
  a.modl = modl_out; a.ff = ff_out;
  if (tables_from) {
    // pair path, second window: its own LG scalars; log f, slopes, ratmod and the PV table are the first window's
    const long long BG = (long long)B * c->G;
    k_table_lg<<<(unsigned)((BG + 127) / 128), 128, 0, st>>>(a, BG);
    TSFF_LAUNCH_OK("k_table_lg");
    double* own_lg = a.lg;
    bind_saved(L, const_cast<char*>(static_cast<const char*>(tables_from)), a);
    a.lg = own_lg;
  } else {
  {
    const size_t smem = (size_t)(2 * c->V + 2 * kXi1N) * 8 + tree_prep_scratch_bytes(c->pv_npad);
    TSFF_SMEM_OPTIN(k_table_prep<T>);
    k_table_prep<T><<<(unsigned)B, kThreads, smem, st>>>(a);
    TSFF_LAUNCH_OK("k_table_prep");
  }
  {
    PvPolesArgs p;
    p.blob = a.D; p.D64 = a.D64; p.pend = a.pend; p.poles = c->xi2; p.pole_bstride = 0;
    p.pnodes = a.ratdf; p.pnode_stride = kXi1N;
    p.z0 = c->xi1_0; p.h = c->xi1_h; p.nodes = c->pv_nodes; p.npad = c->pv_npad; p.P = kXi2N;
    p.outI = a.T; p.outdI = nullptr;
    const size_t smem = (size_t)tree_blob(c->pv_npad).bytes;
    TSFF_SMEM_OPTIN((k_pv_poles<1, TSFF_PV_FP32>));
    TSFF_SMEM_OPTIN((k_pv_poles<2, TSFF_PV_FP32>));
    if (c->pv_precision == TSFF_PV_FP64) {
      p.ntiles = (kXi2N + kPvThreads - 1) / kPvThreads;
      k_pv_poles<1, TSFF_PV_FP64><<<(unsigned)(B * p.ntiles), kPvThreads, 0, st>>>(p);
    } else if ((long long)B * ((kXi2N + 2 * kPvThreads - 1) / (2 * kPvThreads)) >= 2LL * c->sm_count) {
      p.ntiles = (kXi2N + 2 * kPvThreads - 1) / (2 * kPvThreads);
      k_pv_poles<2, TSFF_PV_FP32><<<(unsigned)(B * p.ntiles), kPvThreads, smem, st>>>(p);
    } else {
      p.ntiles = (kXi2N + kPvThreads - 1) / kPvThreads;
      k_pv_poles<1, TSFF_PV_FP32><<<(unsigned)(B * p.ntiles), kPvThreads, smem, st>>>(p);
    }
    TSFF_LAUNCH_OK("k_pv_poles");
  }
  }
  {
    const size_t smem = (size_t)(2 * c->V + kXi2N) * 8;
    // groups per warp: 2 while the grid keeps >= 8 CTAs per SM, else 1 (A/B: 1 -> 6.12, 2 -> 6.03, 4 -> 6.06, 8 -> 6.07 ms on the 1d deck)
    a.jrep = 2;
    while (a.jrep > 1 && (long long)B * ((c->W + kWarps * kFwdJ * a.jrep - 1) / (kWarps * kFwdJ * a.jrep)) < 8LL * c->sm_count) a.jrep /= 2;
#ifdef TSFF_TFWD_JREP
    a.jrep = TSFF_TFWD_JREP;
#endif
    a.ntiles = (c->W + kWarps * kFwdJ * a.jrep - 1) / (kWarps * kFwdJ * a.jrep);
    // the fused angle sum (modl) needs all angles in one CTA; the plain formfactor output can split them
    const bool split_modl = modl_out && !ff_out && table_fwd_split_angles(c, B);
    if (split_modl) a.jrep = 1, a.ntiles = (c->W + kWarps * kFwdJ - 1) / (kWarps * kFwdJ);
    a.asplit = (modl_out && !split_modl) ? 1 : table_angle_split(B * a.ntiles, c->A, c->sm_count);
    a.mpart = (split_modl && a.asplit > 1) ? (double*)(w + L.w_mpart) : nullptr;
    a.Bn = B;
    if (c->ev[0] && c->ev[1]) TSFF_CUDA_OK(cudaEventRecord(c->ev[0], st));
    const unsigned grid = (unsigned)(B * a.ntiles * a.asplit);
    const bool frozen = a.cells && a.cell_mode;
    if (ff_out && frozen) { TSFF_SMEM_OPTIN((k_table_fwd<true, true>)); k_table_fwd<true, true><<<grid, kThreads, smem, st>>>(a); }
    else if (ff_out) { TSFF_SMEM_OPTIN((k_table_fwd<true, false>)); k_table_fwd<true, false><<<grid, kThreads, smem, st>>>(a); }
    else if (frozen) { TSFF_SMEM_OPTIN((k_table_fwd<false, true>)); k_table_fwd<false, true><<<grid, kThreads, smem, st>>>(a); }
    else { TSFF_SMEM_OPTIN((k_table_fwd<false, false>)); k_table_fwd<false, false><<<grid, kThreads, smem, st>>>(a); }
    TSFF_LAUNCH_OK("k_table_fwd");
    if (a.mpart) {
      const long long total = (long long)B * c->W;
      k_table_modl_reduce<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, st>>>(a, total);
      TSFF_LAUNCH_OK("k_table_modl_reduce");
    }
    if (c->ev[0] && c->ev[1]) TSFF_CUDA_OK(cudaEventRecord(c->ev[1], st));
  }
  return TSFF_OK;
}

// pair path: `tables_from` = the first window's saved buffer (this window's f-dependent tables live there); `acc_ws` non-null =
// accumulate-only: this window's table cotangents are added into that workspace (the first window's, already zeroed) and nothing
// else runs -- the first window's call (skip_zero) then finishes both; params_bar of this window is added by table_pair_bwd.
template <typename T>
int table_bwd_t(tsff_ctx* c, int64_t B, const double* params, const void* fe, const void* saved, const double* modl_bar,
                const double* ff_bar, double* params_bar, void* fe_bar, void* ws, cudaStream_t st, const void* tables_from = nullptr,
                void* acc_ws = nullptr, bool skip_zero = false, const double* lgbar2 = nullptr, double lam_shift2 = 0.0) {
  const TableLayout L = table_layout(c, B);
  char* w = static_cast<char*>(ws);
  TableArgs a;
  memset(&a, 0, sizeof(a));
  fill_static(c, a);
  bind_saved(L, const_cast<char*>(static_cast<const char*>(saved)), a);
  a.params = params; a.fe = fe;
  a.modl_bar = modl_bar; a.ff_bar = ff_bar;
  a.desc = (float4*)(w + L.w_desc); a.Dbar = (double*)(w + L.w_Dbar); a.Tbar = (double*)(w + L.w_Tbar);
  a.lnfbar = (double*)(w + L.w_lnfbar); a.slopebar = (double*)(w + L.w_slopebar); a.pnear = (double*)(w + L.w_pnear);
  a.lgbar = (double*)(w + L.w_lgbar);
  a.params_bar = params_bar; a.fe_bar = fe_bar;
  a.lgbar2 = lgbar2; a.lam_shift2 = lam_shift2;
  if (tables_from) {
    double* own_lg = a.lg;
    bind_saved(L, const_cast<char*>(static_cast<const char*>(tables_from)), a);
    a.lg = own_lg;
  }
  if (acc_ws) {
    char* wa = static_cast<char*>(acc_ws);
    a.Tbar = (double*)(wa + L.w_Tbar); a.lnfbar = (double*)(wa + L.w_lnfbar); a.slopebar = (double*)(wa + L.w_slopebar);
    TSFF_CUDA_OK(cudaMemsetAsync(w + L.w_lgbar, 0, L.w_zero_end - L.w_lgbar, st));
  } else if (!skip_zero) {
    TSFF_CUDA_OK(cudaMemsetAsync(w + L.w_zero_begin, 0, L.w_zero_end - L.w_zero_begin, st));
  }
  {
    // wavelengths per thread (kper) and angle chunks per tile (asplit): the pair with the smallest modelled time
    //   waves x (prologue + kper x angles-per-chunk x (1 + look-ahead share)),  waves = CTAs / (2 per SM), whole while few
    // (a long chain amortises the look-ahead point and the run flushes; a short one fills the device when lineouts are few)
    {
      const double slots = 2.0 * c->sm_count;
      double best = 1e300;
      for (int kt = 1; kt <= 2 * TSFF_TBWD_K; kt++) {
        const int nt = (c->W + kBwdThreads * kt - 1) / (kBwdThreads * kt);
        const int kb = (c->W + kBwdThreads * nt - 1) / (kBwdThreads * nt);   // balanced over the tiles
        if (kb != kt) continue;
        for (int as = 1; as <= c->A; as++) {
          const int aper = (c->A + as - 1) / as;
          if (as > 1 && (c->A + as - 2) / (as - 1) == aper) continue;        // same chunk length as the previous split
          const double ctas = (double)B * nt * as;
          double waves = ctas / slots;
          if (waves < 8.0) waves = ceil(waves - 1e-9);
          const double over = kb > TSFF_TBWD_K ? 1.0 + 0.05 * (kb - TSFF_TBWD_K) : 1.0;   // long chains: measured slower (tail effects)
          const double t = waves * (1.5 + kb * aper * (1.0 + 0.6 / kb) * over);
          if (t < best) { best = t; a.kper = kb; a.ntiles = nt; a.asplit = as; }
        }
      }
    }
    const size_t smem = table_bwd_smem(c->V, a.kper);
    a.stage_z = (long long)a.kper * ((c->A + a.asplit - 1) / a.asplit) >= 8 ? 1 : 0;   // >= 2048 points per CTA
    if (c->ev[2] && c->ev[3]) TSFF_CUDA_OK(cudaEventRecord(c->ev[2], st));
    const unsigned grid = (unsigned)(B * a.ntiles * a.asplit);
    const bool frozen = a.cells && a.cell_mode == 2;
#define TSFF_TBWD_LAUNCH(FZ, NI_)                                    \
  do {                                                               \
    TSFF_SMEM_OPTIN((k_table_bwd<FZ, NI_>));                         \
    k_table_bwd<FZ, NI_><<<grid, kBwdThreads, smem, st>>>(a);          \
  } while (0)
    if (frozen) TSFF_TBWD_LAUNCH(true, 0);
    else if (c->I == 1) TSFF_TBWD_LAUNCH(false, 1);
    else if (c->I == 2) TSFF_TBWD_LAUNCH(false, 2);
    else TSFF_TBWD_LAUNCH(false, 0);
#undef TSFF_TBWD_LAUNCH
    TSFF_LAUNCH_OK("k_table_bwd");
    if (c->ev[2] && c->ev[3]) TSFF_CUDA_OK(cudaEventRecord(c->ev[3], st));
  }
  if (acc_ws) return TSFF_OK;
  {
    // per lineout 1640 poles x (descriptor + exact near zone in FP64): one CTA per lineout when lineouts fill the device, else split
    const int ts = B >= c->sm_count ? 1 : (kXi2N + kThreads - 1) / kThreads;
    k_table_tbar<<<(unsigned)(B * ts), kThreads, 0, st>>>(a, ts);
  }
  TSFF_LAUNCH_OK("k_table_tbar");
  {
    PvNodesArgs n;
    n.desc = a.desc; n.tstat = c->tstat; n.P = kXi2N; n.nodes = c->pv_nodes; n.npad = c->pv_npad; n.pbar = a.Dbar;
    n.nsplit = pv_nodes_split(B, n.P, c->sm_count);
    TSFF_SMEM_OPTIN(k_pv_nodes);
    k_pv_nodes<<<(unsigned)(B * n.nsplit), kPvThreads, pv_nodes_smem(n.npad), st>>>(n);
    TSFF_LAUNCH_OK("k_pv_nodes");
  }
  {
    const size_t smem = (size_t)(2 * kXi1N + 2 * c->V) * 8;
    TSFF_SMEM_OPTIN(k_table_bwd_finish<T>);
    k_table_bwd_finish<T><<<(unsigned)B, kThreads, smem, st>>>(a);
    TSFF_LAUNCH_OK("k_table_bwd_finish");
  }
  return TSFF_OK;
}

}  // namespace

namespace {
template <typename T>
int table_pair_bwd_t(tsff_ctx* ca, tsff_ctx* cb, int64_t B, const double* params, const void* fe, const void* saved_a,
                     const void* saved_b, const double* modl_bar_a, const double* modl_bar_b, double* params_bar, void* fe_bar,
                     void* ws_a, void* ws_b, cudaStream_t st) {
  const TableLayout L = table_layout(ca, B);
  TSFF_CUDA_OK(cudaMemsetAsync(static_cast<char*>(ws_a) + L.w_zero_begin, 0, L.w_zero_end - L.w_zero_begin, st));
  int rc = table_bwd_t<T>(cb, B, params, fe, saved_b, modl_bar_b, nullptr, params_bar, fe_bar, ws_b, st, saved_a, ws_a, false);
  if (rc) return rc;
  const double* lgbar_b = (const double*)(static_cast<char*>(ws_b) + table_layout(cb, B).w_lgbar);
  return table_bwd_t<T>(ca, B, params, fe, saved_a, modl_bar_a, nullptr, params_bar, fe_bar, ws_a, st, nullptr, nullptr, true, lgbar_b,
                        cb->lam_shift);
}
}  // namespace

namespace tsff {
int table_fwd(tsff_ctx* c, int64_t B, const double* params, const void* fe, int fe_dtype, double* modl_out, double* ff_out,
              void* saved, void* ws, cudaStream_t st);
// the two contexts must describe the same plasma tables: same V / velocity grid, gradient points, ions, PV precision
bool table_pair_compatible(const tsff_ctx* a, const tsff_ctx* b) {
  return a->mode == TSFF_MODE_TABLE && b->mode == TSFF_MODE_TABLE && a->device == b->device && a->V == b->V && a->G == b->G &&
         a->I == b->I && a->NP == b->NP && a->v0 == b->v0 && a->dv == b->dv && a->pv_precision == b->pv_precision &&
         a->pv_npad == b->pv_npad && !a->cell_mode && !b->cell_mode;
}
int table_pair_fwd(tsff_ctx* ca, tsff_ctx* cb, int64_t B, const double* params, const void* fe, int fe_dtype, double* modl_a,
                   double* modl_b, void* saved_a, void* saved_b, void* ws_a, cudaStream_t st) {
  int rc = table_fwd(ca, B, params, fe, fe_dtype, modl_a, nullptr, saved_a, ws_a, st);
  if (rc) return rc;
  return fe_dtype == TSFF_F32 ? table_fwd_t<float>(cb, B, params, fe, modl_b, nullptr, saved_b, ws_a, st, saved_a)
                              : table_fwd_t<double>(cb, B, params, fe, modl_b, nullptr, saved_b, ws_a, st, saved_a);
}
int table_pair_bwd(tsff_ctx* ca, tsff_ctx* cb, int64_t B, const double* params, const void* fe, int fe_dtype, const void* saved_a,
                   const void* saved_b, const double* modl_bar_a, const double* modl_bar_b, double* params_bar, void* fe_bar,
                   void* ws_a, void* ws_b, cudaStream_t st) {
  if (table_bwd_smem(ca->V, 2 * TSFF_TBWD_K) > 200 * 1024) { set_error("V=%d too large for the table-mode adjoint", ca->V); return TSFF_E_INVALID; }
  return fe_dtype == TSFF_F32
             ? table_pair_bwd_t<float>(ca, cb, B, params, fe, saved_a, saved_b, modl_bar_a, modl_bar_b, params_bar, fe_bar, ws_a, ws_b, st)
             : table_pair_bwd_t<double>(ca, cb, B, params, fe, saved_a, saved_b, modl_bar_a, modl_bar_b, params_bar, fe_bar, ws_a, ws_b, st);
}
size_t table_saved_bytes(const tsff_ctx* c, int64_t B) { return table_layout(c, B).saved_bytes; }
size_t table_ws_bytes(const tsff_ctx* c, int64_t B) { return table_layout(c, B).ws_bytes; }

int table_fwd(tsff_ctx* c, int64_t B, const double* params, const void* fe, int fe_dtype, double* modl_out, double* ff_out,
              void* saved, void* ws, cudaStream_t st) {
  if ((size_t)(2 * c->V + 2 * kXi1N) * 8 > 200 * 1024) { set_error("V=%d too large for table mode", c->V); return TSFF_E_INVALID; }
  return fe_dtype == TSFF_F32 ? table_fwd_t<float>(c, B, params, fe, modl_out, ff_out, saved, ws, st)
                              : table_fwd_t<double>(c, B, params, fe, modl_out, ff_out, saved, ws, st);
}
int table_bwd(tsff_ctx* c, int64_t B, const double* params, const void* fe, int fe_dtype, const void* saved,
              const double* modl_bar, const double* ff_bar, double* params_bar, void* fe_bar, void* ws, cudaStream_t st) {
  if (table_bwd_smem(c->V, 2 * TSFF_TBWD_K) > 200 * 1024) { set_error("V=%d too large for the table-mode adjoint", c->V); return TSFF_E_INVALID; }
  return fe_dtype == TSFF_F32 ? table_bwd_t<float>(c, B, params, fe, saved, modl_bar, ff_bar, params_bar, fe_bar, ws, st)
                              : table_bwd_t<double>(c, B, params, fe, saved, modl_bar, ff_bar, params_bar, fe_bar, ws, st);
}
}  // namespace tsff
