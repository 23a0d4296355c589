// tsff_direct.cu -- TSFF_MODE_DIRECT: 1V kinematics (form_factor.py:182-253, 273-296) with the
// calc_chi_vals electron susceptibility (form_factor.py:369-388) evaluated at every pole xi_e(omega, angle)
// against the lineout's own f-table.  This is the synthetic-sweep workload and the per-pole
// stage of the 2V path.
//
// Kernels
//   k_direct_lg         per (lineout, gradient point): LG scalars
//   k_direct_prep       per lineout: df = gradient(f), tree blob (FP32 node weights + block coefficients), endpoints
//   k_direct_fwd        thread-owns-pole: FP64 kinematics -> FP32 PV sweep (MUFU-bound) -> FP64 assembly
//   k_reduce_modl       mean over G, weighted angle sum, static per-wavelength multiplier
//   k_direct_bwd_poles  FP64 reverse of the assembly per pole; emits Ibar descriptors, lerp scatter, LG cotangents
//   k_pv_nodes          (tsff_pv_kernels.cuh) thread-owns-node adjoint PV sweep (MUFU-bound)
//   k_direct_bwd_finish per lineout: Dbar -> df_bar -> fe_bar; LG cotangents -> params_bar
#include <stdlib.h>

#include "tsff_pv_kernels.cuh"

using namespace tsff;

namespace {

constexpr int kThreads = 256;
constexpr int kLGSDoubles = kLGDoubles + kLGXDoubles;   // a saved (lineout, gradient point) row: the LG scalars, then their LGX

// np.gradient(f, dv) at node i (form_factor.py:372): central inside, first-order one-sided at the ends
// (division by the uniform spacing as a multiplication by its reciprocal: every consumer of df goes through this one
// function, so the node values seen by the tree weights, the exact zone and the adjoint agree bit for bit)
template <typename T>
__device__ __forceinline__ double grad_at(const T* f, int V, double dv, int i) {
  const double idv = fast_rcp(dv);
  if (i <= 0) return ((double)f[1] - (double)f[0]) * idv;
  if (i >= V - 1) return ((double)f[V - 1] - (double)f[V - 2]) * idv;
  return ((double)f[i + 1] - (double)f[i - 1]) * (0.5 * idv);
}

struct LGS { LG L; LGX X; };

struct DirectLayout {  // byte offsets inside `saved` and `ws` for a batch of B lineouts
  size_t s_lg, s_I, s_dI, saved_bytes;
  size_t w_D, w_D64, w_pend, w_ff, w_desc, w_accfe, w_accdf, w_pendbar, w_Dbar, w_lgbar, w_zero_begin, w_zero_end,
      ws_bytes;
};

DirectLayout direct_layout(const tsff_ctx* c, int64_t B) {
  DirectLayout L;
  const size_t P = (size_t)c->G * c->W * c->A;
  size_t o = 0;
  L.s_lg = o; o += align_up((size_t)B * c->G * kLGSDoubles * 8);
  L.s_I = o; o += align_up((size_t)B * P * 8);
  L.s_dI = o; o += align_up((size_t)B * P * 8);
  L.saved_bytes = o;
  o = 0;
  L.w_D = o; o += align_up((size_t)B * tree_blob(c->pv_npad).bytes);
  L.w_D64 = o; o += align_up(c->pv_precision == TSFF_PV_FP64 ? (size_t)B * c->pv_npad * 8 : 0);
  L.w_pend = o; o += align_up((size_t)B * 2 * 8);
  L.w_ff = o; o += align_up((size_t)B * P * 8);
  L.w_desc = o; o += align_up((size_t)B * P * 16);
  L.w_zero_begin = o;
  L.w_Dbar = o; o += align_up((size_t)B * c->pv_npad * 8);
  L.w_accfe = o; o += align_up((size_t)B * c->V * 8);
  L.w_accdf = o; o += align_up((size_t)B * c->V * 8);
  L.w_pendbar = o; o += align_up((size_t)B * 2 * 8);
  L.w_lgbar = o; o += align_up((size_t)B * c->G * kLGDoubles * 8);
  L.w_zero_end = o;
  L.ws_bytes = o;
  return L;
}

struct DirectArgs {
  // static
  int W, A, G, nI, V, NP, nodes, npad, ntiles;
  int stage_fe;   // forward: the f table is staged into shared memory behind the blob
  double lam_shift, v0, dv;
  const double *omgs, *costh, *wts, *jmul;
  ZTab zt;
  // per call
  const double* params;
  const void* fe;
  double* lg;        // [B][G][kLGDoubles]
  double* sI;        // [B][P]
  double* sdI;       // [B][P]
  unsigned char* D;  // [B][blob bytes]  per-lineout tree blob: FP32 node weights + block expansion coefficients
  const double* tstat;
  double* D64;       // [B][npad] or null
  double* pend;      // [B][2]
  double* ff;        // [B][G][W][A]  (null: not wanted and the angle sum is fused, see modl1)
  double* modl1;     // [B][W] or null: A == 1 and G == 1 -> the pole kernel writes modl itself (no k_reduce_modl pass)
  // backward
  const double* modl_bar;
  const double* ff_bar;
  float4* desc;
  double *accfe, *accdf, *pendbar, *Dbar, *lgbar;
  double* params_bar;
  void* fe_bar;
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

// ---- prep -------------------------------------------------------------------------------------------------
// LG scalars of every (lineout, gradient point): one thread each (FP64 divisions and square roots; kept out of the
// per-lineout CTA of k_direct_prep, where a single thread's serial chain would hold back the whole CTA)
__global__ void __launch_bounds__(128) k_direct_lg(const DirectArgs a, long long BG) {
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  if (t >= BG) return;
  const long long b = t / a.G;
  LG L;
  lg_zero(L);
  lg_forward(a.params + b * a.NP, a.nI, (int)(t % a.G), a.G, a.lam_shift, L);
  store_lg(a.lg + t * kLGSDoubles, L);
  LGX X;
  lgx_make(L, a.nI, a.zt.h, X);
  for (int i = 0; i < kLGXDoubles; i++) a.lg[t * kLGSDoubles + kLGDoubles + i] = reinterpret_cast<const double*>(&X)[i];
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 3) k_direct_prep(const DirectArgs a) {
  const long long b = blockIdx.x;
  const T* fe = static_cast<const T*>(a.fe) + b * a.V;
  const int M = a.nodes - 1;
  {
    // node values p_i = gradient(f)_i (16 consecutive nodes per thread), then the tree blob.  The lineout's f row is staged in
    // shared memory first (coalesced): a thread's 16-node run read straight from global memory touches 32 sectors per request
    extern __shared__ __align__(16) unsigned char prep_smem[];
    // (one pad element per 16: a warp's lanes read elements 16 apart, which would otherwise all fall into two banks)
    T* sfe = reinterpret_cast<T*>(prep_smem + tree_prep_scratch_bytes(a.npad));
    for (int i = threadIdx.x; i < a.V; i += kThreads) sfe[i + (i >> 4)] = fe[i];
    const int V = a.V;
    const double idv = fast_rcp(a.dv);
    auto at = [sfe](int i) { return (double)sfe[i + (i >> 4)]; };
    // np.gradient at node i, the arithmetic of grad_at (tree_prep_cta_f synchronises the CTA before its first pget)
    auto pget = [at, V, idv](int i) {
      if (i <= 0) return (at(1) - at(0)) * idv;
      if (i >= V - 1) return (at(V - 1) - at(V - 2)) * idv;
      return (at(i + 1) - at(i - 1)) * (0.5 * idv);
    };
    tree_prep_cta_f(pget, M, a.npad, a.D + b * tree_blob(a.npad).bytes, a.tstat, reinterpret_cast<double*>(prep_smem));
  }
  if (a.D64) {  // log-form weights for the FP64 validation path
    for (int i = threadIdx.x; i < a.npad; i += kThreads) {
      double d = 0.0;
      if (i <= M) {
        double pc = grad_at(fe, a.V, a.dv, i);
        double sR = (i < M) ? (grad_at(fe, a.V, a.dv, i + 1) - pc) / a.dv : 0.0;
        double sL = (i > 0) ? (pc - grad_at(fe, a.V, a.dv, i - 1)) / a.dv : 0.0;
        d = sR - sL;
      }
      a.D64[b * a.npad + i] = d;
    }
  }
  if (threadIdx.x == 0) {
    a.pend[2 * b] = grad_at(fe, a.V, a.dv, 0);
    a.pend[2 * b + 1] = grad_at(fe, a.V, a.dv, M);
  }
}

// FP64 tail of one pole: exact near nodes, ion susceptibility, lerp of f and f', assembly; stores P, I, dI/dxi.
// Not inlined: shared by the R poles of a thread (the kernel would otherwise carry R copies of the FP64 log code).
template <typename T, int PREC>
__device__ __noinline__ void direct_point(const DirectArgs& a, const LG& sL, const LGX& sX, const T* fe, long long b, int g, int idx,
                                          int n, int wb0, double farI, double farJ, double g0d) {
  const int j = idx / a.A, ia = idx % a.A;
  const int M = a.nodes - 1;
  const double omgs = a.omgs[j];
  KinX q;
  kin_forward_x(sL, sX, omgs, a.costh[ia], q);
  IonX io;
  ion_forward_x<0>(sL, sX, a.nI, a.zt, q, io);
  double I, dI;
  if (PREC == TSFF_PV_FP32) {
    const int V = a.V;
    const double dv = a.dv;
    tree_near_exact(q.xie, a.v0, dv, M, n, wb0, [fe, V, dv](int i) { return grad_at(fe, V, dv, i); }, I, dI);
    I += farI;
    dI += farJ;
  } else {
    const double p0 = a.pend[2 * b], pM = a.pend[2 * b + 1];
    pv_finish(farI, farJ, p0, pM, g0d, g0d + (double)(a.nodes - 1) * a.dv, I, dI);
  }
  int i_f; double t_f, sl_f;
  const double fphi = lerp_uniform(fe, a.V, a.v0, a.dv, q.xie, i_f, t_f, sl_f);   // form_factor.py:376
  const double d0 = grad_at(fe, a.V, a.dv, i_f), d1 = grad_at(fe, a.V, a.dv, i_f + 1);
  const double dfe = d0 + t_f * (d1 - d0);  // form_factor.py:377 (clamped: t_f in {0,1} picks the edge value)
  const double chiEr = -q.ikl2 * I;          // form_factor.py:385-386
  const double chiEi = kPi * q.ikl2 * dfe;   // form_factor.py:381
  AsmX s;
  const double P = assemble_forward_x(sL, sX, q, io, chiEr, chiEi, fphi, omgs, s);
  const long long pidx = ((b * a.G + g) * (long long)a.W + j) * a.A + ia;
  if (a.ff) a.ff[pidx] = P;
  if (a.modl1) a.modl1[pidx] = a.jmul[j] * (0.0 + P * a.wts[0]) / 1.0;   // k_reduce_modl's arithmetic for A = G = 1
  a.sI[pidx] = I;
  a.sdI[pidx] = dI;
}

// ---- forward ------------------------------------------------------------------------------------------------
// grid.x = B * G * ntiles; a tile covers kThreads*R consecutive (j,a) pairs of one (lineout, gradient point).
template <int R, typename T, int PREC, int MINB = 3>
__global__ void __launch_bounds__(kThreads, MINB) k_direct_fwd(const __grid_constant__ DirectArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ LGS sLS;   // the (lineout, gradient point) scalars and their reciprocals
  const LG& sL = sLS.L;
  const LGX& sX = sLS.X;
  const int tile = blockIdx.x % a.ntiles;
  const long long bg = blockIdx.x / a.ntiles;
  const int g = (int)(bg % a.G);
  const long long b = bg / a.G;
  const int M = a.nodes - 1;
  const TreeBlob tb = tree_blob(a.npad);
  if (threadIdx.x < kLGSDoubles) reinterpret_cast<double*>(&sLS)[threadIdx.x] = a.lg[bg * kLGSDoubles + threadIdx.x];
  // the lineout's f table rides along with the blob (same bulk copy, same barrier) when it is a float table that keeps three
  // CTAs per SM: the exact zone (7 nodes around the pole) and the lerps of f and f' gather from it per lane -- from shared
  // memory instead of 32 scattered global loads per warp (the region held 20 % of the kernel's stall samples)
  const T* fe = static_cast<const T*>(a.fe) + b * a.V;
  if (PREC == TSFF_PV_FP32) {
    const uint32_t febytes = a.stage_fe ? (uint32_t)(a.V * sizeof(T)) : 0u;
    stage_blob2(smem_raw, a.D + b * tb.bytes, (uint32_t)tb.bytes, smem_raw + tb.bytes, fe, febytes, &bar);
    if (a.stage_fe) fe = reinterpret_cast<const T*>(smem_raw + tb.bytes);
  } else {
    __syncthreads();
  }
  const int WA = a.W * a.A;

  TreePole tp[R];
  double g0d[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    int idx = (tile * kThreads + threadIdx.x) * R + r;
    if (idx >= WA) idx = WA - 1;
    KinX q;
    kin_forward_x(sL, sX, a.omgs[idx / a.A], a.costh[idx % a.A], q);
    tp[r] = tree_pole(q.xie, a.v0, a.dv, M, a.npad);
    g0d[r] = a.v0 - q.xie;
  }
  double accI[R], accJ[R], accJ2[R], nrI[R], nrJ[R];
  if (PREC == TSFF_PV_FP32) {
#pragma unroll
    for (int r = 0; r < R; r++) accI[r] = accJ[r] = accJ2[r] = nrI[r] = nrJ[r] = 0.0;
    tree_far<R>(smem_raw, tb, tp, accI, accJ, accJ2);
#pragma unroll
    for (int r = 0; r < R; r++) {
      const TreeAcc na = tree_near(smem_raw, tb, tp[r]);
      nrI[r] = na.I;
      nrJ[r] = na.J;
    }
  } else {
    pv_accumulate_f64<R, true>(a.D64 + b * a.npad, a.nodes, a.dv, g0d, accI, accJ);
  }

#pragma unroll
  for (int r = 0; r < R; r++) {
    const int idx = (tile * kThreads + threadIdx.x) * R + r;
    if (idx >= WA) continue;
    double farI = accI[r], farJ = accJ[r];
    if (PREC == TSFF_PV_FP32) {
      farI += nrI[r];
      farJ = accJ[r] / (kTs * a.dv) + accJ2[r] / (kTs2 * a.dv) + nrJ[r] / a.dv;
    }
    direct_point<T, PREC>(a, sL, sX, fe, b, g, idx, (int)(-tp[r].un), tp[r].wb0, farI, farJ, g0d[r]);
  }
}

// modl[b][j] = jmul[j] * sum_a w_a * mean_g ff[b][g][j][a]   (generate_spectra.py:164-165, 193, 197, 210-216)
__global__ void __launch_bounds__(kThreads) k_reduce_modl(const double* ff, const double* wts, const double* jmul, int G,
                                                          int W, int A, long long total, double* modl) {
  long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= total) return;
  const long long b = t / W;
  const int j = (int)(t % W);
  double s = 0.0;
  for (int g = 0; g < G; g++) {
    const double* row = ff + ((b * G + g) * (long long)W + j) * A;
    for (int ia = 0; ia < A; ia++) s += row[ia] * wts[ia];
  }
  modl[t] = jmul[j] * s / (double)G;
}

// ---- backward: poles --------------------------------------------------------------------------------------
#ifndef TSFF_BWDP_MINB
#define TSFF_BWDP_MINB 4
#endif
#ifndef TSFF_BWDP_RB
#define TSFF_BWDP_RB 4
#endif
#ifndef TSFF_BWDP_THREADS
#define TSFF_BWDP_THREADS 128
#endif
#ifndef TSFF_BWDP_STAGE_Z
#define TSFF_BWDP_STAGE_Z 0    // 1: Z' table staged per CTA (measured slower: the CTA is short, the table is L1-resident)
#endif
#ifndef TSFF_BWDP_WARPFIN
#define TSFF_BWDP_WARPFIN 0
#endif
constexpr int kBwdpThreads = TSFF_BWDP_THREADS;

// Exact-zone adjoint of one pole, the kNearHalf nodes either side of its nearest node n:  d I / d p_i = W_i, the second difference
// of u ln|u| at u_i = (z_i - xi)/h (the h ln h parts cancel in the difference).  The FORWARD needs these weights in FP64 (they
// multiply p_i in a sum with cancellation against the far field); the adjoint only scales them by Ibar, so FP32 logarithms
// (u_i formed in FP64, then rounded) give the cotangent to ~1e-6 of its scale against a 1e-4 bar -- 9 MUFU.LG2 instead of
// 9 FP64 log polynomials (half of this kernel's FP64 instructions before).  End nodes keep the FP64 form (rare).
__device__ __forceinline__ void pv_bwd_pole_exact_f32(double gw, double xi, double Ibar, double z0, double h, int nodes, int n,
                                                      int wb0, double* pnear, int ei0, double ev0, int ei1, double ev1) {
  const int M = nodes - 1;
  const int lo = max(1, n - kNearHalf), hi = min(M - 1, n + kNearHalf);
  if (lo <= hi && Ibar != 0.0) {
    auto phi = [gw](int i) {
      const float u = (float)(gw + (double)i);
      return u * (__log2f(fmaxf(fabsf(u), 1e-30f)) * 0.69314718f);
    };
    float pm = phi(lo - 1), pc = phi(lo);
    for (int i = lo; i <= hi; i++) {
      const float pp = phi(i + 1);
      double v = Ibar * (double)(pp - 2.f * pc + pm);
      if (i == ei0) { v += ev0; ev0 = 0.0; }
      if (i == ei1) { v += ev1; ev1 = 0.0; }
      atomicAdd(&pnear[i], v);
      pm = pc;
      pc = pp;
    }
  }
  if (ev0 != 0.0) atomicAdd(&pnear[ei0], ev0);
  if (ev1 != 0.0) atomicAdd(&pnear[ei1], ev1);
  if (Ibar == 0.0) return;
  const double ih = fast_rcp(h);
  if (wb0 == 0) {
    const double g0 = z0 - xi;
    atomicAdd(&pnear[0], Ibar * ((pv_phi(g0 + h) - pv_phi(g0)) * ih - 1.0 - log_abs(g0)));
  }
  if ((unsigned)(M / kTS - wb0) <= 2u) {
    const double gM = z0 + (double)M * h - xi;
    atomicAdd(&pnear[M], Ibar * ((pv_phi(gM - h) - pv_phi(gM)) * ih + 1.0 + log_abs(gM)));
  }
}

template <int R, typename T, int NI>
__global__ void __launch_bounds__(kBwdpThreads, TSFF_BWDP_MINB) k_direct_bwd_poles(const DirectArgs a) {
  __shared__ LGS sLS;
  __shared__ double sred[kLGDoubles * (kBwdpThreads / 32)];
#if TSFF_BWDP_STAGE_Z
  __shared__ ZZ s_zz[kXi2N];   // the Z' table as (re, im) pairs
#endif
  const int tile = blockIdx.x % a.ntiles;
  const long long bg = blockIdx.x / a.ntiles;
  const int g = (int)(bg % a.G);
  const long long b = bg / a.G;
  if (threadIdx.x < kLGSDoubles) reinterpret_cast<double*>(&sLS)[threadIdx.x] = a.lg[bg * kLGSDoubles + threadIdx.x];
#if TSFF_BWDP_STAGE_Z
  for (int i = threadIdx.x; i < kXi2N; i += kBwdpThreads) {
    s_zz[i].r = a.zt.zr[i];
    s_zz[i].i = a.zt.zi[i];
  }
#endif
  __syncthreads();
  const LG& L = sLS.L;
  const LGX& X = sLS.X;
  const T* fe = static_cast<const T*>(a.fe) + b * a.V;
  const int WA = a.W * a.A;
  const double idv = fast_rcp(a.dv), iG = 1.0 / (double)a.G;
  const double xlast = a.v0 + (a.V - 1) * a.dv;
  LG Lb;
  lg_zero(Lb);
  for (int r = 0; r < R; r++) {
    const int idx = (tile * kBwdpThreads + threadIdx.x) * R + r;
    if (idx >= WA) continue;
    const int j = idx / a.A, ia = idx % a.A;
    const long long pidx = ((b * a.G + g) * (long long)a.W + j) * a.A + ia;
    double Pbar = 0.0;
    if (a.modl_bar) Pbar += a.modl_bar[b * a.W + j] * a.jmul[j] * a.wts[ia] * iG;
    if (a.ff_bar) Pbar += a.ff_bar[pidx];
    const double omgs = a.omgs[j], cth = a.costh[ia];
    KinX q;
    kin_forward_x(L, X, omgs, cth, q);
    IonX io;
#if TSFF_BWDP_STAGE_Z
    ion_forward_bf<NI>(L, X, a.nI, (const ZZ*)s_zz, a.zt, q, io);
#else
    ion_forward_bf<NI>(L, X, a.nI, ZRows{a.zt.zr, a.zt.zi}, a.zt, q, io);
#endif
    const double I = a.sI[pidx], dI = a.sdI[pidx];
    int i_f; double t_f, sl_f;
    const double fphi = lerp_uniform_bf(fe, a.V, a.v0, idv, q.xie, i_f, t_f, sl_f);
    const double d0 = grad_at(fe, a.V, a.dv, i_f), d1 = grad_at(fe, a.V, a.dv, i_f + 1);
    const bool clamped = (q.xie <= a.v0) || (q.xie >= xlast) || !(q.xie == q.xie);
    const double dfe = d0 + t_f * (d1 - d0);
    const double sl_d = clamped ? 0.0 : (d1 - d0) * idv;
    const double chiEr = -q.ikl2 * I, chiEi = kPi * q.ikl2 * dfe;
    AsmX s;
    assemble_forward_x(L, X, q, io, chiEr, chiEi, fphi, omgs, s);
    PointBar pb;
    KinBar kb = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    assemble_backward_x<NI>(L, X, a.nI, q, io, chiEr, chiEi, fphi, s, Pbar, pb, kb, Lb);
    // chi_e provider reverse (form_factor.py:376-386)
    kb.ikl2 += -I * pb.chiEr + kPi * dfe * pb.chiEi;
    const double Ibar = -q.ikl2 * pb.chiEr;
    const double dfe_bar = kPi * q.ikl2 * pb.chiEi;
    kb.xie += Ibar * dI + dfe_bar * sl_d + pb.fphi * sl_f;
    kin_backward_x(L, X, cth, q, kb, Lb);
    if (pb.fphi != 0.0) {
      atomicAdd(&a.accfe[b * a.V + i_f], (1.0 - t_f) * pb.fphi);
      atomicAdd(&a.accfe[b * a.V + i_f + 1], t_f * pb.fphi);
    }
    // d I / d p_i for the nodes next to the pole (and an end node inside its window); the rest is k_pv_nodes'
    // (far blocks + near-window series)
    const TreePole tp = tree_pole(q.xie, a.v0, a.dv, a.nodes - 1, a.npad);
    const int np = (int)(-tp.un);
    a.desc[b * ((long long)a.G * WA) + (long long)g * WA + idx] =
        make_float4(tp.un, tp.ndh, (float)Ibar, __int_as_float((np >> 4) | (tp.wb0 << 10) | (tp.w2 << 18)));   // = pv_desc
    pv_bwd_pole_exact_f32(tp.gw, q.xie, Ibar, a.v0, a.dv, a.nodes, np, tp.wb0, a.accdf + b * a.V, i_f, (1.0 - t_f) * dfe_bar,
                          i_f + 1, t_f * dfe_bar);
  }
  double vals[kLGDoubles];
  store_lg(vals, Lb);
#if TSFF_BWDP_WARPFIN
  warp_accumulate(vals, kLGDoubles, a.lgbar + bg * kLGDoubles);
#else
  block_accumulate<kBwdpThreads / 32>(vals, kLGDoubles, sred, a.lgbar + bg * kLGDoubles);
#endif
}

// ---- backward: finish ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) k_direct_bwd_finish(const DirectArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* spb = reinterpret_cast<double*>(smem_raw);  // pbar[V]
  const long long b = blockIdx.x;
  const int V = a.V, M = a.nodes - 1;
  const double* pfar = a.Dbar + b * a.npad;  // far-field part of df_bar from k_pv_nodes
  const double ih = 1.0 / a.dv;
  for (int i = threadIdx.x; i < V; i += kThreads) spb[i] = a.accdf[b * V + i] + (i <= M ? pfar[i] : 0.0);
  __syncthreads();
  T* fe_bar = static_cast<T*>(a.fe_bar) + b * V;
  for (int k = threadIdx.x; k < V; k += kThreads) {
    double fb = a.accfe[b * V + k];
    // f_k enters p_{k-1} (+), p_{k+1} (-), and p_k at the two ends
    if (k >= 1) fb += spb[k - 1] * ((k - 1 == 0) ? ih : 0.5 * ih);
    if (k <= V - 2) fb -= spb[k + 1] * ((k + 1 == V - 1) ? ih : 0.5 * ih);
    if (k == 0) fb -= spb[0] * ih;
    if (k == V - 1) fb += spb[V - 1] * ih;
    fe_bar[k] = (T)fb;
  }
  if (threadIdx.x == 0) {
    double* pbar = a.params_bar + b * a.NP;
    for (int k = 0; k < a.NP; k++) pbar[k] = 0.0;
    for (int g = 0; g < a.G; g++) {
      LG Lb;
      load_lg(a.lgbar + (b * a.G + g) * kLGDoubles, Lb);
      lg_backward(a.params + b * a.NP, a.nI, g, a.G, a.lam_shift, Lb, pbar);
    }
  }
}

// params_bar alone (the table cotangent is finished inside k_pv_nodes when one CTA owns the whole lineout)
__global__ void __launch_bounds__(64) k_direct_params_bar(const DirectArgs a, long long B) {
  const long long b = (long long)blockIdx.x * 64 + threadIdx.x;
  if (b >= B) return;
  double* pbar = a.params_bar + b * a.NP;
  for (int k = 0; k < a.NP; k++) pbar[k] = 0.0;
  for (int g = 0; g < a.G; g++) {
    LG Lb;
    load_lg(a.lgbar + (b * a.G + g) * kLGDoubles, Lb);
    lg_backward(a.params + b * a.NP, a.nI, g, a.G, a.lam_shift, Lb, pbar);
  }
}

void fill_static(const tsff_ctx* c, DirectArgs& a) {
  a.W = c->W; a.A = c->A; a.G = c->G; a.nI = c->I; a.V = c->V; a.NP = c->NP;
  a.nodes = c->pv_nodes; a.npad = c->pv_npad;
  a.lam_shift = c->lam_shift; a.v0 = c->v0; a.dv = c->dv;
  a.omgs = c->omgs; a.costh = c->costh; a.wts = c->wts; a.jmul = c->jmul; a.zt = c->zt;
  a.tstat = c->tstat;
}

template <typename T>
int direct_fwd_t(tsff_ctx* c, int64_t B, const double* params, const void* fe, double* modl_out, double* ff_out,
                 void* saved, void* ws, cudaStream_t st) {
  const DirectLayout L = direct_layout(c, B);
  char* sv = static_cast<char*>(saved);
  char* w = static_cast<char*>(ws);
  DirectArgs a;
  memset(&a, 0, sizeof(a));
  fill_static(c, a);
  a.params = params; a.fe = fe;
  a.lg = (double*)(sv + L.s_lg); a.sI = (double*)(sv + L.s_I); a.sdI = (double*)(sv + L.s_dI);
  a.D = (unsigned char*)(w + L.w_D);
  a.D64 = c->pv_precision == TSFF_PV_FP64 ? (double*)(w + L.w_D64) : nullptr;
  a.pend = (double*)(w + L.w_pend);
  const bool fuse_modl = modl_out && c->A == 1 && c->G == 1;
  a.ff = ff_out ? ff_out : (fuse_modl ? nullptr : (double*)(w + L.w_ff));
  a.modl1 = fuse_modl ? modl_out : nullptr;
  {
    const long long BG = (long long)B * c->G;
    k_direct_lg<<<(unsigned)((BG + 127) / 128), 128, 0, st>>>(a, BG);
    TSFF_LAUNCH_OK("k_direct_lg");
    const size_t psm = tree_prep_scratch_bytes(c->pv_npad) + (size_t)(c->V + c->V / 16 + 1) * sizeof(T);
    TSFF_SMEM_OPTIN(k_direct_prep<T>);
    k_direct_prep<T><<<(unsigned)B, kThreads, psm, st>>>(a);
  }
  TSFF_LAUNCH_OK("k_direct_prep");
  const int WA = c->W * c->A;
  size_t smem = (size_t)tree_blob(c->pv_npad).bytes;
  // stage the f table too when it is 16-byte copyable and three CTAs per SM still fit (float tables up to ~4096 nodes)
  {
    const size_t febytes = (size_t)c->V * sizeof(T);
    const bool aligned = febytes % 16 == 0 && (reinterpret_cast<uintptr_t>(fe) % 16 == 0);
    a.stage_fe = (c->pv_precision != TSFF_PV_FP64 && aligned && 3 * (smem + febytes + 1024) <= 228 * 1024) ? 1 : 0;
    if (a.stage_fe) smem += febytes;
  }
  // poles per thread: 2 while the grid still fills the device, else 1 (4 poles/thread at 128 registers measured 7% slower)
  const long long tiles4 = (WA + 4 * kThreads - 1) / (4 * kThreads), tiles2 = (WA + 2 * kThreads - 1) / (2 * kThreads);
  if (c->ev[0] && c->ev[1]) TSFF_CUDA_OK(cudaEventRecord(c->ev[0], st));
  if (c->pv_precision == TSFF_PV_FP64) {
    a.ntiles = (WA + kThreads - 1) / kThreads;
    k_direct_fwd<1, T, TSFF_PV_FP64><<<(unsigned)(B * c->G * a.ntiles), kThreads, 0, st>>>(a);
  } else if ((long long)B * c->G * tiles4 >= 2LL * c->sm_count && c->tune_fwd_r4) {  // tuning switch: measured slower
    a.ntiles = (int)tiles4;
    TSFF_SMEM_OPTIN((k_direct_fwd<4, T, TSFF_PV_FP32, 2>));
    k_direct_fwd<4, T, TSFF_PV_FP32, 2><<<(unsigned)(B * c->G * a.ntiles), kThreads, smem, st>>>(a);
  } else if ((long long)B * c->G * tiles2 >= 2LL * c->sm_count) {
    a.ntiles = (int)tiles2;
    TSFF_SMEM_OPTIN((k_direct_fwd<2, T, TSFF_PV_FP32, 3>));
    k_direct_fwd<2, T, TSFF_PV_FP32, 3><<<(unsigned)(B * c->G * a.ntiles), kThreads, smem, st>>>(a);
  } else {
    a.ntiles = (WA + kThreads - 1) / kThreads;
    TSFF_SMEM_OPTIN((k_direct_fwd<1, T, TSFF_PV_FP32, 4>));
    k_direct_fwd<1, T, TSFF_PV_FP32, 4><<<(unsigned)(B * c->G * a.ntiles), kThreads, smem, st>>>(a);
  }
  TSFF_LAUNCH_OK("k_direct_fwd");
  if (c->ev[0] && c->ev[1]) TSFF_CUDA_OK(cudaEventRecord(c->ev[1], st));
  if (modl_out && !fuse_modl) {
    const long long total = (long long)B * c->W;
    k_reduce_modl<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, st>>>(a.ff, c->wts, c->jmul, c->G, c->W, c->A,
                                                                                     total, modl_out);
    TSFF_LAUNCH_OK("k_reduce_modl");
  }
  return TSFF_OK;
}

template <typename T>
int direct_bwd_t(tsff_ctx* c, int64_t B, const double* params, const void* fe, const void* saved, const double* modl_bar,
                 const double* ff_bar, double* params_bar, void* fe_bar, void* ws, cudaStream_t st) {
  const DirectLayout L = direct_layout(c, B);
  const char* sv = static_cast<const char*>(saved);
  char* w = static_cast<char*>(ws);
  DirectArgs a;
  memset(&a, 0, sizeof(a));
  fill_static(c, a);
  a.params = params; a.fe = fe;
  a.lg = (double*)(sv + L.s_lg); a.sI = (double*)(sv + L.s_I); a.sdI = (double*)(sv + L.s_dI);
  a.modl_bar = modl_bar; a.ff_bar = ff_bar;
  a.desc = (float4*)(w + L.w_desc); a.accfe = (double*)(w + L.w_accfe); a.accdf = (double*)(w + L.w_accdf);
  a.pendbar = (double*)(w + L.w_pendbar); a.Dbar = (double*)(w + L.w_Dbar); a.lgbar = (double*)(w + L.w_lgbar);
  a.params_bar = params_bar; a.fe_bar = fe_bar;
  const int WA = c->W * c->A;
  PvNodesArgs n;
  n.desc = a.desc; n.tstat = c->tstat; n.P = c->G * WA; n.nodes = c->pv_nodes; n.npad = c->pv_npad; n.pbar = a.Dbar;
  n.nsplit = pv_nodes_split(B, n.P, c->sm_count);
  // one CTA per lineout in the node sweep: it finishes fe_bar itself and Dbar (first in the zeroed range) is never touched
  const bool fused = n.nsplit == 1;
  const size_t z0 = fused ? L.w_accfe : L.w_zero_begin;
  TSFF_CUDA_OK(cudaMemsetAsync(w + z0, 0, L.w_zero_end - z0, st));
  constexpr int RB = TSFF_BWDP_RB;
  a.ntiles = (WA + RB * kBwdpThreads - 1) / (RB * kBwdpThreads);
  if (c->I == 1) k_direct_bwd_poles<RB, T, 1><<<(unsigned)(B * c->G * a.ntiles), kBwdpThreads, 0, st>>>(a);
  else k_direct_bwd_poles<RB, T, 0><<<(unsigned)(B * c->G * a.ntiles), kBwdpThreads, 0, st>>>(a);
  TSFF_LAUNCH_OK("k_direct_bwd_poles");
  if (fused) {
    n.accdf = a.accdf; n.accfe = a.accfe; n.fe_bar = fe_bar; n.fe_f32 = sizeof(T) == 4; n.ih = 1.0 / c->dv;
  }
  if (c->ev[2] && c->ev[3]) TSFF_CUDA_OK(cudaEventRecord(c->ev[2], st));
  TSFF_SMEM_OPTIN(k_pv_nodes);
  k_pv_nodes<<<(unsigned)(B * n.nsplit), kPvThreads, pv_nodes_smem(n.npad), st>>>(n);
  TSFF_LAUNCH_OK("k_pv_nodes");
  if (c->ev[2] && c->ev[3]) TSFF_CUDA_OK(cudaEventRecord(c->ev[3], st));
  if (fused) {
    k_direct_params_bar<<<(unsigned)((B + 63) / 64), 64, 0, st>>>(a, (long long)B);
    TSFF_LAUNCH_OK("k_direct_params_bar");
    return TSFF_OK;
  }
  const size_t smem = (size_t)c->V * 8;
  TSFF_SMEM_OPTIN(k_direct_bwd_finish<T>);
  k_direct_bwd_finish<T><<<(unsigned)B, kThreads, smem, st>>>(a);
  TSFF_LAUNCH_OK("k_direct_bwd_finish");
  return TSFF_OK;
}

}  // namespace

namespace tsff {
size_t direct_saved_bytes(const tsff_ctx* c, int64_t B) { return direct_layout(c, B).saved_bytes; }
size_t direct_ws_bytes(const tsff_ctx* c, int64_t B) { return direct_layout(c, B).ws_bytes; }

int direct_fwd(tsff_ctx* c, int64_t B, const double* params, const void* fe, int fe_dtype, double* modl_out, double* ff_out,
               void* saved, void* ws, cudaStream_t st) {
  if ((size_t)tree_blob(c->pv_npad).bytes > 200 * 1024 || tree_prep_scratch_bytes(c->pv_npad) + (size_t)(c->V + c->V / 16 + 1) * 8 > 200 * 1024) { set_error("V=%d too large for shared-memory staging", c->V); return TSFF_E_INVALID; }
  return fe_dtype == TSFF_F32 ? direct_fwd_t<float>(c, B, params, fe, modl_out, ff_out, saved, ws, st)
                              : direct_fwd_t<double>(c, B, params, fe, modl_out, ff_out, saved, ws, st);
}
int direct_bwd(tsff_ctx* c, int64_t B, const double* params, const void* fe, int fe_dtype, const void* saved,
               const double* modl_bar, const double* ff_bar, double* params_bar, void* fe_bar, void* ws, cudaStream_t st) {
  if ((size_t)c->V * 8 > 200 * 1024) { set_error("V=%d too large", c->V); return TSFF_E_INVALID; }
  return fe_dtype == TSFF_F32 ? direct_bwd_t<float>(c, B, params, fe, saved, modl_bar, ff_bar, params_bar, fe_bar, ws, st)
                              : direct_bwd_t<double>(c, B, params, fe, saved, modl_bar, ff_bar, params_bar, fe_bar, ws, st);
}
}  // namespace tsff
