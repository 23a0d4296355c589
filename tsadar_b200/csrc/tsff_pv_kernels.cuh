// tsff_pv_kernels.cuh -- the principal-value sweeps shared by every mode (block-multipole form, tsff_tree.cuh).
//
//   tree_prep_cta : per lineout: FP32 node weights + the Laurent coefficients of every 64- and 256-node block
//   k_pv_poles    : thread-owns-pole forward sweep   I(xi_p), dI/dxi_p
//   k_pv_nodes    : adjoint sweep  pbar_i = sum_p Ibar_p dI_p/dp_i  (far: block-owner gathers local coefficients over
//                   the poles, near: node-owner loops over the poles whose window covers its block)
//
// Bound: instruction issue (FP32 FMA pipe + XU side by side); no HBM traffic to speak of.
// Shared memory: the per-lineout blob (node weights 16 KB + block coefficients 10 KB at 4096 nodes) staged by TMA
// bulk copies.
#pragma once
#include "tsff_common.cuh"
#include "tsff_tree.cuh"

namespace tsff {

constexpr int kPvThreads = 256;

constexpr int kPrepStatic = (kTsE1 - kTsCM1) + 4 * kTK * kTK;   // staged static tables: CM1 QE1 CM2 QE2 | T12
TSFF_HD size_t tree_prep_scratch_bytes(int npad) { return ((size_t)(npad / kTS + npad / kTS2) * kTK * 2 + kPrepStatic) * 8; }   // moments + coefficients + tables

#if defined(__CUDACC__)
// One bulk copy completed on an mbarrier, in pieces of at most 32 KB.  Called by all threads; one use per kernel.
__device__ __forceinline__ void stage_blob(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u)
      bulk_g2s(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, min(32768u, bytes - o), bar);
  }
  mbar_wait(bar, 0);
}

// The same with a second region (the lineout's f table) on the same mbarrier: dst2 / src2 / bytes2 (bytes2 = 0: none).
__device__ __forceinline__ void stage_blob2(void* dst, const void* src, uint32_t bytes, void* dst2, const void* src2, uint32_t bytes2,
                                            uint64_t* bar) {
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes + bytes2);
    for (uint32_t o = 0; o < bytes; o += 32768u)
      bulk_g2s(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, min(32768u, bytes - o), bar);
    if (bytes2) bulk_g2s(dst2, src2, bytes2, bar);
  }
  mbar_wait(bar, 0);
}

// Static tables of the expansion (tree_static_entry): built once per context / call.
static __global__ void __launch_bounds__(256) k_tree_static(int M, double* out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kTreeStaticDoubles; i += gridDim.x * blockDim.x) out[i] = tree_static_entry(i, M);
}
constexpr int kTreeStaticGrid = (kTreeStaticDoubles + 255) / 256;   // one entry per thread (the entries are small serial loops)

template <int C>
__device__ __forceinline__ void tree_translate_fixed(const double (&m0)[kTK], double (&m)[kTK]) {
#pragma unroll
  for (int k = 0; k < kTK; k++) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j <= k; j++) acc = fma(tree_T(C, k, j), m0[j], acc);   // folded: C, k, j are compile-time here
    m[k] = acc;
  }
}
__device__ __forceinline__ void tree_translate_child(int c, const double (&m0)[kTK], double (&m)[kTK]) {
  tree_translate_shift(c, m0, m);   // run-time child index without divergence (tsff_tree.cuh)
}

// Per-lineout preparation, executed by one CTA.  pget(i), 0 <= i <= M: node values p_i (FP64).
// Writes the lineout's blob (tree_blob layout) to global memory: weights, packed coefficients of both levels, leading
// coefficients.  All phases are spread over the whole CTA:
//   1. level-1 moments  mu1[b][k] = sum_i p_i x_i^k           thread (b, quarter): 16 consecutive nodes held in registers,
//                                                             powers by recurrence, the four quarters added by shuffles
//                                                             (also writes the FP32 node weights)
//   2. level-2 moments from the four children by translation  thread (B, k), static matrices T12
//   3. coefficients A_m from the moments (+ end-node rows)    thread (block, m)
//   4. packing into the Horner layout                          thread (block, q)
// scratch: shared, tree_prep_scratch_bytes(npad).  Ends with the CTA synchronised.  blockDim.x must be a multiple of 32.
template <typename PGet>
__device__ __forceinline__ void tree_prep_cta_f(PGet pget, int M, int npad, unsigned char* blob, const double* tstat,
                                                double* scratch) {
  const TreeBlob tb = tree_blob(npad);
  const int NB = tb.NB, NB2 = tb.NB2, NT = NB + NB2;
  double* mu = scratch;              // [NT][kTK]  (level 1 first)
  double* A = scratch + NT * kTK;    // [NT][kTK]
  double* sCM = A + NT * kTK;        // static tables CM1 QE1 CM2 QE2 (offsets relative to kTsCM1), then T12
  double* sT12 = sCM + (kTsE1 - kTsCM1);
  float* Wt = reinterpret_cast<float*>(blob + tb.oW);
  // the static tables are fetched while phase 1 runs (their first use is behind its barrier)
  for (int i = threadIdx.x; i < kPrepStatic; i += blockDim.x)
    sCM[i] = i < kTsE1 - kTsCM1 ? tstat[kTsCM1 + i] : tstat[kTsT12 + (i - (kTsE1 - kTsCM1))];
  constexpr int kQ = 16;             // nodes per thread
  static_assert(kQ == kTS0, "one thread per level-0 block");
  __syncthreads();                   // sT12 is used inside phase 1
  for (int base = 0; base < NB * (kTS / kQ); base += blockDim.x) {   // warp-uniform trip count (shuffles inside)
    const int item = base + threadIdx.x;
    const bool valid = item < NB * (kTS / kQ);
    const int b = item / (kTS / kQ), qd = item % (kTS / kQ);
    double m[kTK];
#pragma unroll
    for (int k = 0; k < kTK; k++) m[k] = 0.0;
    if (valid) {
      // the thread's 16 nodes ARE one level-0 block: moments about its own centre (scale 8) by power recurrence
      const int b0 = (kTS / kTS0) * b + qd;
      double m0[kTK];
#pragma unroll
      for (int k = 0; k < kTK; k++) m0[k] = 0.0;
      float wv[kQ];
#pragma unroll
      for (int o = 0; o < kQ; o++) {
        const int i = kTS0 * b0 + o;
        const double pv = (i >= 1 && i <= M - 1) ? pget(i) : 0.0;
        wv[o] = (float)pv;
        // m0[k] += p x_o^k with x_o = -(o - 7.5) / 8: o and k are unrolled, so the powers fold to immediate operands (one DFMA
        // per (node, order) instead of a DADD and a DMUL)
#pragma unroll
        for (int k = 0; k < kTK; k++) m0[k] = fma(pv, tree_xpow0(o, k), m0[k]);
      }
      float4* w4 = reinterpret_cast<float4*>(Wt + kTS0 * b0);
#pragma unroll
      for (int o = 0; o < kQ; o += 4) w4[o / 4] = make_float4(wv[o], wv[o + 1], wv[o + 2], wv[o + 3]);
      // this block's share of the level-1 moments: translation child qd -> parent; the 105 matrix entries of each child
      // are compile-time constants (immediate operands: no shared-memory traffic in this hot loop)
      tree_translate_child(qd, m0, m);
      // level-0 coefficients (interior nodes only, no end-node rows), packed, straight to the blob
      double A0[kTK];
#pragma unroll
      for (int mm = 0; mm < kTK; mm++) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < kTK / 2; j++)
          if (2 * j <= mm) {
            constexpr double s0 = kTs0;
            const double cmv = tree_cm(mm, j, s0);   // folded at compile time (mm, j are unrolled)
            acc = fma(cmv, m0[mm - 2 * j], acc);
          }
        A0[mm] = acc * (1.0 / kTs0);
      }
      float4* ab = reinterpret_cast<float4*>(blob + tb.oAB0) + b0 * (kTK / 2);
#pragma unroll
      for (int q = 0; q < kTK / 2; q++) {
        const int e0 = 2 * q, e1 = 2 * q + 1;
        ab[q] = make_float4(e0 + 2 < kTK ? (float)A0[e0 + 2] : 0.f, (float)((double)(e0 + 1) * A0[e0]),
                            e1 + 2 < kTK ? (float)A0[e1 + 2] : 0.f, (float)((double)(e1 + 1) * A0[e1]));
      }
      double* dl = reinterpret_cast<double*>(blob + tb.oLD0) + 2 * b0;
      dl[0] = A0[0];
      dl[1] = A0[1];
    }
#pragma unroll
    for (int k = 0; k < kTK; k++) {
      m[k] += __shfl_xor_sync(0xffffffffu, m[k], 1);
      m[k] += __shfl_xor_sync(0xffffffffu, m[k], 2);
    }
    if (valid && qd == 0) {
#pragma unroll
      for (int k = 0; k < kTK; k++) mu[b * kTK + k] = m[k];
    }
  }
  __syncthreads();
  const double* T12 = sT12;
  for (int it = threadIdx.x; it < NB2 * kTK; it += blockDim.x) {
    const int B = it / kTK, k = it % kTK;
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const double* m1 = mu + (4 * B + c) * kTK;
      const double* t = T12 + (c * kTK + k) * kTK;
#pragma unroll
      for (int j = 0; j < kTK; j++) acc = fma(t[j], m1[j], acc);   // T12[c][k][j] = 0 for j > k
    }
    mu[NB * kTK + it] = acc;
  }
  __syncthreads();
  for (int it = threadIdx.x; it < NT * kTK; it += blockDim.x) {
    const int blk = it / kTK, m = it % kTK;
    const bool lvl2 = blk >= NB;
    const int b = lvl2 ? blk - NB : blk, S = lvl2 ? kTS2 : kTS;
    const double s = lvl2 ? kTs2 : kTs;
    const double* cm = sCM + ((lvl2 ? kTsCM2 : kTsCM1) - kTsCM1);
    const double* qe = sCM + ((lvl2 ? kTsQE2 : kTsQE1) - kTsCM1);
    const double* mb = mu + blk * kTK;
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < kTK / 2; j++) a = fma(cm[m * (kTK / 2) + j], 2 * j <= m ? mb[m - 2 * j] : 0.0, a);   // cm = 0 there too
    a *= 1.0 / s;   // s is a power of two
    if (b == 0) a = fma(pget(0), qe[m], a);
    if (M >= S * b && M < S * (b + 1)) a = fma(pget(M), qe[kTK + m], a);
    A[it] = a;
  }
  __syncthreads();
  for (int it = threadIdx.x; it < NT * (kTK / 2); it += blockDim.x) {
    const int blk = it / (kTK / 2), q = it % (kTK / 2);
    const bool lvl2 = blk >= NB;
    const int b = lvl2 ? blk - NB : blk;
    const double* Ab = A + blk * kTK;
    const int m0 = 2 * q, m1 = 2 * q + 1;
    float4 v;
    v.x = m0 + 2 < kTK ? (float)Ab[m0 + 2] : 0.f;
    v.y = (float)((double)(m0 + 1) * Ab[m0]);
    v.z = m1 + 2 < kTK ? (float)Ab[m1 + 2] : 0.f;
    v.w = (float)((double)(m1 + 1) * Ab[m1]);
    reinterpret_cast<float4*>(blob + (lvl2 ? tb.oAB2 : tb.oAB1))[b * (kTK / 2) + q] = v;
    if (q == 0) {
      double* dl = reinterpret_cast<double*>(blob + (lvl2 ? tb.oLD2 : tb.oLD1)) + 2 * b;
      dl[0] = Ab[0];
      dl[1] = Ab[1];
    }
  }
  __syncthreads();
}
// p[0..M] given as an array (shared or global memory)
__device__ __forceinline__ void tree_prep_cta(const double* p, int M, int npad, unsigned char* blob, const double* tstat,
                                              double* scratch) {
  tree_prep_cta_f([p](int i) { return p[i]; }, M, npad, blob, tstat, scratch);
}
#endif

struct PvPolesArgs {
  const unsigned char* blob;  // [B][blob_bytes]  per-lineout tree blob (tree_prep_cta)
  const double* D64;     // [B][npad] log-form weights D_i (FP64 validation path) or nullptr
  const double* pend;    // [B][2]   (p_0, p_M)   (FP64 validation path)
  const double* pnodes;  // node values p_i (FP64), row b at pnodes + b*pnode_stride
  long long pnode_stride;
  const double* poles;   // pole positions, row b at poles + b*pole_bstride
  long long pole_bstride;  // 0: all lineouts share one pole list (table mode)
  double z0, h;
  int nodes, npad, P, ntiles;
  double* outI;          // [B][P]
  double* outdI;         // [B][P] or nullptr
};

#if defined(__CUDACC__)
template <int R, int PREC>
__global__ void __launch_bounds__(kPvThreads) k_pv_poles(const PvPolesArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const long long b = blockIdx.x / a.ntiles;
  const int tile = blockIdx.x % a.ntiles;
  const int M = a.nodes - 1;
  const TreeBlob tb = tree_blob(a.npad);
  if (PREC == TSFF_PV_FP32) stage_blob(smem_raw, a.blob + b * tb.bytes, (uint32_t)tb.bytes, &bar);

  const double* poles = a.poles + b * a.pole_bstride;
  double xi[R];
  TreePole tp[R];
  double g0d[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    int p = (tile * kPvThreads + threadIdx.x) * R + r;
    xi[r] = poles[p < a.P ? p : a.P - 1];
    tp[r] = tree_pole(xi[r], a.z0, a.h, M, a.npad);
    g0d[r] = a.z0 - xi[r];
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
    pv_accumulate_f64<R, true>(a.D64 + b * a.npad, a.nodes, a.h, g0d, accI, accJ);
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    int p = (tile * kPvThreads + threadIdx.x) * R + r;
    if (p < a.P) {
      double I, dI;
      if (PREC == TSFF_PV_FP32) {
        const double* pn = a.pnodes + b * a.pnode_stride;
        tree_near_exact(xi[r], a.z0, a.h, M, (int)(-tp[r].un), tp[r].wb0, [pn](int i) { return pn[i]; }, I, dI);
        I += accI[r] + nrI[r];
        dI += accJ[r] / (kTs * a.h) + accJ2[r] / (kTs2 * a.h) + nrJ[r] / a.h;
      } else {
        const double p0 = a.pend[2 * b], pM = a.pend[2 * b + 1];
        pv_finish(accI[r], accJ[r], p0, pM, g0d[r], g0d[r] + (double)(a.nodes - 1) * a.h, I, dI);
      }
      a.outI[b * a.P + p] = I;
      if (a.outdI) a.outdI[b * a.P + p] = dI;
    }
  }
}
#endif

// ---- adjoint ----------------------------------------------------------------------------------------------------
struct PvNodesArgs {
  const float4* desc;   // [B][P]  (un = -n_p, ndh = -delta_p/h, Ibar_p, packed keys gn | wb0 << 10 | w2 << 18: pv_desc)
  const double* tstat;  // k_tree_static table
  int P, nodes, npad, nsplit;
  double* pbar;         // [B][npad]  d loss / d p_i without the exact near part (nsplit > 1: atomically accumulated, zero it)
  // Optional fused epilogue (direct mode, nsplit == 1, fe_bar != null): p = gradient(f) on V = nodes + 1 table entries,
  // so the lineout's table cotangent is finished here instead of in a second pass over pbar:
  //   fe_bar[k] = accfe[k] + sum_i d p_i / d f_k (pbar_i + accdf_i)         (np.gradient adjoint, form_factor.py:372)
  // accdf / accfe [B][V]: the exact near-zone and lerp contributions scattered by the pole kernel.  pbar is not written.
  const double* accdf = nullptr;
  const double* accfe = nullptr;
  void* fe_bar = nullptr;
  int fe_f32 = 0;
  double ih = 0.0;      // 1 / dv
};

constexpr int kNodeChunk = 1024;   // poles staged per shared-memory chunk (2 x 16 KB)
constexpr int kTreeMaxNpad = 256 * kTS;  // 1024 level-0 groups (10-bit key), 256 level-1 and 64 level-2 blocks

inline size_t pv_nodes_smem(int npad) {
  const int NB = npad / kTS, NB0 = npad / kTS0;
  return (size_t)kNodeChunk * 16 + (size_t)(npad + 2) * 8 + (size_t)(NB + npad / kTS2) * kTKA * 8 + (size_t)NB0 * kTKA * 4 + (size_t)(NB0 + 2) * 4 * 2 + 64;
}

// pole splits per lineout: 1 when the batch alone fills the device, else enough CTAs for two per SM (each split
// re-does the O(nodes) spreading, so never more than one split per kNodeChunk poles)
inline int pv_nodes_split(int64_t B, int P, int sm_count) {
  if (B >= 2LL * sm_count) return 1;
  long long want = (2LL * sm_count + B - 1) / B;
  // a few lineouts (a fit batch of 2-8): latency is what counts, and the pole loops shrink with the split while the O(nodes)
  // spreading each split repeats is short -- allow splits down to 256 poles; otherwise one split per chunk of poles
  const int grain = B * 8 <= sm_count ? 256 : kNodeChunk;
  long long most = (P + grain - 1) / grain;
  if (want > most) want = most;
  return want < 1 ? 1 : (int)want;
}

#if defined(__CUDACC__)
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// Adjoint partition = the transpose of the forward sweep's.  For a pole with nearest node n, level-0 group gn = n / 16,
// level-1 window wb0..wb0+2 (twelve level-0 groups) and level-2 window w2..w2+2:
//   FAR 2    level-2 blocks outside the level-2 window: L2_{B,m} += Ibar t^(m+1), t = 128 h / (z_c - xi)
//   FAR 1    level-1 blocks inside the level-2 window but outside the level-1 window: L1_{b,m}, t = 32 h / (z_c - xi)
//   FAR 0    level-0 groups of the window at least two groups from gn: L0_{b,m} += Ibar t^(m+1), t = 8 h / (z_c - xi)
//   NEAR     the nodes of groups gn-1 .. gn+1, one by one (six-term series); |i - n| <= kNearHalf masked -- those nodes, and
//            an end node inside the window, get their exact FP64 terms from the caller (pv_bwd_pole_exact)
// One CTA per (lineout, pole split).  Per chunk of kNodeChunk poles:
//   1. stage the descriptors in shared memory
//   2. counting-sort the descriptors by gn: every set of poles used below is one contiguous range of the sorted list
//   3. FAR 2: thread (block, pole subset) over all poles of the chunk, two poles per packed instruction
//   4. FAR 1 / FAR 0: a warp takes one parent block at a time, lanes = 4 children x 8 pole subsets, over the poles whose
//      window holds the parent; the subsets are added by shuffles
//   5. NEAR: thread <-> node, over the poles with gn within one group of the node's group
// FP32 inside a chunk (at most 64 poles per partial sum in the near loop), FP64 across.  Then L1 and L0 are spread to
// the nodes with the static weights q_m(e) and everything is written (or atomically added) to pbar.
static __global__ void __launch_bounds__(kPvThreads, 3) k_pv_nodes(const PvNodesArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NB = a.npad / kTS, NB0 = a.npad / kTS0, NB2 = a.npad / kTS2, M = a.nodes - 1;
  float4* ssort = reinterpret_cast<float4*>(smem_raw);                 // [kNodeChunk] descriptors sorted by gn
  double* spbar = reinterpret_cast<double*>(ssort + kNodeChunk);       // [npad + 2]  (the fused epilogue needs nodes + 1 <= npad + 1)
  double* sL = spbar + a.npad + 2;                                     // [NB][kTKA]
  double* sL2 = sL + NB * kTKA;                                        // [NB2][kTKA]
  float* sL0 = reinterpret_cast<float*>(sL2 + NB2 * kTKA);             // [NB0][kTKA]  (FP32: three CTAs per SM fit with it)
  int* shist = reinterpret_cast<int*>(sL0 + NB0 * kTKA);               // [NB0 + 2]  first sorted slot of key gn
  int* scur = shist + NB0 + 2;                                         // [NB0 + 2]
  const long long b = blockIdx.x / a.nsplit;
  const int split = blockIdx.x % a.nsplit;
  const int per = (a.P + a.nsplit - 1) / a.nsplit;
  const int p_begin = split * per, p_end = min(a.P, p_begin + per);
  const float4* desc = a.desc + b * a.P;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  for (int i = threadIdx.x; i < a.npad + 2; i += kPvThreads) spbar[i] = 0.0;
  for (int i = threadIdx.x; i < (NB + NB2) * kTKA; i += kPvThreads) sL[i] = 0.0;
  for (int i = threadIdx.x; i < NB0 * kTKA; i += kPvThreads) sL0[i] = 0.f;
  if (a.fe_bar) {   // fused epilogue: pull this lineout's accumulator rows towards L2 now, they are read at the very end
    const int Vv = a.nodes + 1;
    for (int i = threadIdx.x * 16; i < Vv; i += kPvThreads * 16) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.accdf + b * Vv + i));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.accfe + b * Vv + i));
    }
  }
  // far-2 thread layout: NBP (power of two >= NB2, <= 64) blocks x Q pole subsets
  int NBP = 1;
  while (NBP < NB2) NBP <<= 1;
  const int Q = kPvThreads / NBP;
  const int fb = threadIdx.x % NBP, fq = threadIdx.x / NBP;
  const float cb = (float)(2 * fb) + (float)(0.5 * (kTS2 - 1) / kTs2);
  const float invs2 = (float)(1.0 / kTs2);
  double L64[kTKA];
#pragma unroll
  for (int m = 0; m < kTKA; m++) L64[m] = 0.0;
  const float lim = (float)kNearHalf + 0.5f;
  // first / last level-0 key of the poles whose window starts at level-1 block w
  auto glo = [NB](int w) { return w <= 0 ? 0 : (kTS / kTS0) * (w + 1); };
  auto ghi = [NB, NB0](int w) { return w >= NB - 3 ? NB0 - 1 : (kTS / kTS0) * (w + 1) + (kTS / kTS0) - 1; };
  // the same for the level-2 window start w2
  auto g2lo = [](int w) { return w <= 0 ? 0 : (kTS2 / kTS0) * (w + 1); };
  auto g2hi = [NB2, NB0](int w) { return w >= NB2 - 3 ? NB0 - 1 : (kTS2 / kTS0) * (w + 1) + (kTS2 / kTS0) - 1; };

  for (int c0 = p_begin; c0 < p_end; c0 += kNodeChunk) {
    const int nc = min(kNodeChunk, p_end - c0);
    __syncthreads();
    for (int k = threadIdx.x; k < NB0 + 2; k += kPvThreads) shist[k] = 0;
    __syncthreads();
    // ---- counting sort by gn = n / 16 (n = -un), straight from global memory (the second read hits L1 / L2)
    for (int k = threadIdx.x; k < nc; k += kPvThreads) atomicAdd(&shist[(__float_as_int(desc[c0 + k].w) & 1023) + 1], 1);
    __syncthreads();
    if (wid == 0) {  // inclusive scan -> shist[k] = first sorted slot of key k, shist[NB0] = nc
      int carry = 0;
      for (int k0 = 0; k0 <= NB0; k0 += 32) {
        const int k = k0 + lane;
        int v = (k <= NB0) ? shist[k] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v += n;
        }
        v += carry;
        if (k <= NB0) { shist[k] = v; scur[k] = v; }
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nc; k += kPvThreads) {
      const float4 d = desc[c0 + k];
      ssort[atomicAdd(&scur[__float_as_int(d.w) & 1023], 1)] = d;
    }
    __syncthreads();
    // ---- far 2 (the order of the poles does not matter): level-2 window start w2 = clamp(n / 256 - 1, 0, NB2 - 3)
    if (fb < NB2) {
      float2 Lp[kTKA];
#pragma unroll
      for (int m = 0; m < kTKA; m++) Lp[m] = make_float2(0.f, 0.f);
      for (int k = fq; k < nc; k += 2 * Q) {
        const float4 d0 = ssort[k];
        const bool has1 = (k + Q) < nc;
        const float4 d1 = has1 ? ssort[k + Q] : d0;
        const int w20 = __float_as_int(d0.w) >> 18, w21 = __float_as_int(d1.w) >> 18;
        const float g0 = fmaf(d0.y, invs2, fmaf(d0.x, invs2, cb));
        const float g1 = fmaf(d1.y, invs2, fmaf(d1.x, invs2, cb));
        const bool far0 = (unsigned)(fb - w20) > 2u;
        const bool far1 = has1 && ((unsigned)(fb - w21) > 2u);
        const float2 t = make_float2(far0 ? rcp_approx(g0) : 0.f, far1 ? rcp_approx(g1) : 0.f);
        float2 pw = fmul2(make_float2(d0.z, d1.z), t);
#pragma unroll
        for (int m = 0; m < kTKA; m++) {
          Lp[m] = fadd2(Lp[m], pw);
          pw = fmul2(pw, t);
        }
      }
#pragma unroll
      for (int m = 0; m < kTKA; m++) L64[m] += (double)Lp[m].x + (double)Lp[m].y;
    }
    // ---- far 1: a warp takes one level-2 block at a time; lane = (child level-1 block lane / 8, pole subset lane % 8) over
    // the poles whose level-2 window holds it; a child inside the pole's level-1 window is masked
    for (int B2 = wid; B2 < NB2; B2 += kPvThreads / 32) {
      const int b1 = (kTS2 / kTS) * B2 + (lane >> 3), sub = lane & 7;
      const int klo = shist[g2lo(max(B2 - 2, 0))], khi = shist[g2hi(min(B2, NB2 - 3)) + 1];
      const float cb1 = (float)(2 * b1) + (float)(0.5 * (kTS - 1) / kTs);
      float2 Lp[kTKA];
#pragma unroll
      for (int m = 0; m < kTKA; m++) Lp[m] = make_float2(0.f, 0.f);
      for (int k = klo + sub; k < khi; k += 16) {
        const float4 d0 = ssort[k];
        const bool has1 = (k + 8) < khi;
        const float4 d1 = has1 ? ssort[k + 8] : d0;
        const bool far0 = (unsigned)(b1 - ((__float_as_int(d0.w) >> 10) & 255)) > 2u;
        const bool far1 = has1 && ((unsigned)(b1 - ((__float_as_int(d1.w) >> 10) & 255)) > 2u);
        const float g0 = fmaf(d0.y, kInvTs, fmaf(d0.x, kInvTs, cb1));
        const float g1 = fmaf(d1.y, kInvTs, fmaf(d1.x, kInvTs, cb1));
        const float2 t = make_float2(far0 ? rcp_approx(g0) : 0.f, far1 ? rcp_approx(g1) : 0.f);
        float2 pw = fmul2(make_float2(d0.z, d1.z), t);
#pragma unroll
        for (int m = 0; m < kTKA; m++) {
          Lp[m] = fadd2(Lp[m], pw);
          pw = fmul2(pw, t);
        }
      }
#pragma unroll
      for (int m = 0; m < kTKA; m++) {
        float v = Lp[m].x + Lp[m].y;
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (sub == 0) sL[b1 * kTKA + m] += (double)v;   // one writer per block
      }
    }
    // ---- far 0: a warp takes one level-1 block at a time (strided over the warps: every warp samples the whole grid, so
    // bunched poles do not leave warps idle); lane = (child group c = lane / 8, pole subset lane % 8) over the poles whose
    // window [wb0, wb0 + 2] holds the block; the eight subsets are added by shuffles
    for (int b1 = wid; b1 < NB; b1 += kPvThreads / 32) {
      const int b0 = (kTS / kTS0) * b1 + (lane >> 3), sub = lane & 7;
      const int klo = shist[glo(max(b1 - 2, 0))], khi = shist[ghi(min(b1, NB - 3)) + 1];
      const float cb0 = (float)(2 * b0) + (float)(0.5 * (kTS0 - 1) / kTs0);
      const float invs0 = (float)(1.0 / kTs0);
      float2 Lp[kTKA];
#pragma unroll
      for (int m = 0; m < kTKA; m++) Lp[m] = make_float2(0.f, 0.f);
      for (int k = klo + sub; k < khi; k += 16) {
        const float4 d0 = ssort[k];
        const bool has1 = (k + 8) < khi;
        const float4 d1 = has1 ? ssort[k + 8] : d0;
        const int g0n = __float_as_int(d0.w) & 1023, g1n = __float_as_int(d1.w) & 1023;
        const bool far0 = abs(b0 - g0n) >= 2;
        const bool far1 = has1 && abs(b0 - g1n) >= 2;
        const float g0 = fmaf(d0.y, invs0, fmaf(d0.x, invs0, cb0));
        const float g1 = fmaf(d1.y, invs0, fmaf(d1.x, invs0, cb0));
        const float2 t = make_float2(far0 ? rcp_approx(g0) : 0.f, far1 ? rcp_approx(g1) : 0.f);
        float2 pw = fmul2(make_float2(d0.z, d1.z), t);
#pragma unroll
        for (int m = 0; m < kTKA; m++) {
          Lp[m] = fadd2(Lp[m], pw);
          pw = fmul2(pw, t);
        }
      }
#pragma unroll
      for (int m = 0; m < kTKA; m++) {
        float v = Lp[m].x + Lp[m].y;             // FP32 like the sums themselves; FP64 across chunks
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (sub == 0) sL0[b0 * kTKA + m] += v;   // one writer per group
      }
    }
    // ---- near: thread <-> node i (its slot of spbar is private), poles with gn in [gi - 1, gi + 1]
    for (int i = threadIdx.x; i < a.npad; i += kPvThreads) {
      if (i < 1 || i > M - 1) continue;
      const int gi = i >> 4;
      const int klo = shist[max(gi - 1, 0)], khi = shist[min(gi + 1, NB0 - 1) + 1];
      const float fi = (float)i;
      double acc64 = 0.0;
      for (int kc = klo; kc < khi; kc += 64) {
        const int ke = min(kc + 64, khi);
        float acc = 0.f;
        for (int k = kc; k < ke; k++) {
          const float4 d = ssort[k];
          const float u = fi + d.x;   // i - n_p, exact
          const float x = fabsf(u) > lim ? rcp_approx(u + d.y) : 0.f;
          const float s2 = x * x;
          float pI = fmaf(1.f / 66.f, s2, 1.f / 45.f);
          pI = fmaf(pI, s2, 1.f / 28.f);
          pI = fmaf(pI, s2, 1.f / 15.f);
          pI = fmaf(pI, s2, 1.f / 6.f);
          pI = fmaf(pI, s2, 1.f);
          acc = fmaf(d.z * x, pI, acc);
        }
        acc64 += (double)acc;
      }
      spbar[i] += acc64;
    }
  }
  if (fb < NB2) {
#pragma unroll
    for (int m = 0; m < kTKA; m++) atomicAdd(&sL2[fb * kTKA + m], L64[m]);
  }
  __syncthreads();
  // ---- spread the local coefficients of both levels to the nodes and write out
  double* out = a.pbar + b * a.npad;
  const double* q = a.tstat + kTsQA;
  const bool fused = a.fe_bar != nullptr;
  const int V = a.nodes + 1;
  {
    // node i = threadIdx.x + k kPvThreads: its in-block offsets i % 16, i % 64, i % 256 do not depend on k, so a thread needs
    // one row of spreading weights per level; one pass per level keeps only twelve of them in registers at a time (the
    // slots of spbar a thread touches are its own)
    static_assert(kPvThreads % kTS2 == 0, "the level-2 offset of a thread's nodes must not depend on k");
    const double* q2 = a.tstat + kTsQA2;
    double wq[kTKA];
#pragma unroll
    for (int m = 0; m < kTKA; m++) wq[m] = a.tstat[kTsQA0 + (threadIdx.x % kTS0) * kTKA + m];
    for (int i = threadIdx.x; i < a.npad; i += kPvThreads) {
      if (i < 1 || i > M - 1) continue;
      const float* l0 = sL0 + (i / kTS0) * kTKA;
      double v = spbar[i];
#pragma unroll
      for (int m = 0; m < kTKA; m++) v += (double)l0[m] * wq[m];
      spbar[i] = v;
    }
#pragma unroll
    for (int m = 0; m < kTKA; m++) wq[m] = q[(threadIdx.x % kTS) * kTKA + m];
    for (int i = threadIdx.x; i < a.npad; i += kPvThreads) {
      if (i < 1 || i > M - 1) continue;
      const double* l1 = sL + (i / kTS) * kTKA;
      double v = spbar[i];
#pragma unroll
      for (int m = 0; m < kTKA; m++) v += l1[m] * wq[m];
      spbar[i] = v;
    }
#pragma unroll
    for (int m = 0; m < kTKA; m++) wq[m] = q2[(threadIdx.x % kTS2) * kTKA + m];
    for (int i = threadIdx.x; i < a.npad; i += kPvThreads) {
      double v = 0.0;
      const int nb = i / kTS;
      if (i >= 1 && i <= M - 1) {
        const double* l2 = sL2 + (i / kTS2) * kTKA;
        v = spbar[i];
#pragma unroll
        for (int m = 0; m < kTKA; m++) v += l2[m] * wq[m];
      } else if (i == 0) {
#pragma unroll
        for (int m = 0; m < kTKA; m++) v += sL[m] * q[kTS * kTKA + m] + sL2[m] * q2[kTS2 * kTKA + m];
      } else if (i == M) {
#pragma unroll
        for (int m = 0; m < kTKA; m++) v += sL[nb * kTKA + m] * q[(kTS + 1) * kTKA + m] + sL2[(i / kTS2) * kTKA + m] * q2[(kTS2 + 1) * kTKA + m];
      }
      if (fused) spbar[i] = v;   // each thread rewrites only the slot it has just read
      else if (a.nsplit == 1) out[i] = v;
      else if (v != 0.0) atomicAdd(&out[i], v);
    }
  }
  if (!fused) return;
  // + the pole kernel's exact near-zone / lerp contributions (eight independent loads in flight per thread; the rows
  // were prefetched into L2 when the CTA started)
  const double* adf = a.accdf + b * V;
  const double* afe = a.accfe + b * V;
  constexpr int U = 8;
  for (int i0 = threadIdx.x; i0 < V; i0 += U * kPvThreads) {
    double t[U];
#pragma unroll
    for (int u = 0; u < U; u++) { const int i = i0 + u * kPvThreads; t[u] = i < V ? __ldg(adf + i) : 0.0; }
#pragma unroll
    for (int u = 0; u < U; u++) { const int i = i0 + u * kPvThreads; if (i < V) spbar[i] += t[u]; }
  }
  __syncthreads();
  // f_k enters p_{k-1} (+), p_{k+1} (-), and p_k at the two ends (central differences inside, one-sided at the ends)
  const double ih = a.ih;
  for (int k0 = threadIdx.x; k0 < V; k0 += U * kPvThreads) {
    double t[U];
#pragma unroll
    for (int u = 0; u < U; u++) { const int k = k0 + u * kPvThreads; t[u] = k < V ? __ldg(afe + k) : 0.0; }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int k = k0 + u * kPvThreads;
      if (k >= V) break;
      double fb = t[u];
      if (k >= 1) fb += spbar[k - 1] * ((k - 1 == 0) ? ih : 0.5 * ih);
      if (k <= V - 2) fb -= spbar[k + 1] * ((k + 1 == V - 1) ? ih : 0.5 * ih);
      if (k == 0) fb -= spbar[0] * ih;
      if (k == V - 1) fb += spbar[V - 1] * ih;
      if (a.fe_f32) static_cast<float*>(a.fe_bar)[b * V + k] = (float)fb;
      else static_cast<double*>(a.fe_bar)[b * V + k] = fb;
    }
  }
}

// Exact (FP64) adjoint contributions of one pole (nearest node n as split by tree_pole / pv_desc): the kNearHalf nodes
// either side of it, and an end node when its
// block lies in the pole's near window.  Atomically added to pnear[0..M].  Two extra contributions (ei0, ev0),
// (ei1, ev1) to the same array (the lerp adjoint of the direct mode, which lands on nodes next to the pole) ride on the
// same atomics when they fall on a node of the exact zone; pass ev = 0 to skip.
__device__ __forceinline__ void pv_bwd_pole_exact(double xi, double Ibar, double z0, double h, int nodes, int n, int wb0,
                                                  double* pnear, int ei0 = 0, double ev0 = 0.0, int ei1 = 0, double ev1 = 0.0) {
  const int M = nodes - 1;
  const int lo = max(1, n - kNearHalf), hi = min(M - 1, n + kNearHalf);
  const double ih = fast_rcp(h);
  if (lo <= hi && Ibar != 0.0) {
    double pm = pv_phi(z0 + (double)(lo - 1) * h - xi), pc = pv_phi(z0 + (double)lo * h - xi);
    for (int i = lo; i <= hi; i++) {
      const double pp = pv_phi(z0 + (double)(i + 1) * h - xi);
      double v = Ibar * (pp - 2.0 * pc + pm) * ih;
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
  if (wb0 == 0) {
    const double g0 = z0 - xi;
    atomicAdd(&pnear[0], Ibar * ((pv_phi(g0 + h) - pv_phi(g0)) * ih - 1.0 - log_abs(g0)));
  }
  if ((unsigned)(M / kTS - wb0) <= 2u) {
    const double gM = z0 + (double)M * h - xi;
    atomicAdd(&pnear[M], Ibar * ((pv_phi(gM - h) - pv_phi(gM)) * ih + 1.0 + log_abs(gM)));
  }
}

// descriptor of one pole for k_pv_nodes
__device__ __forceinline__ float4 pv_desc(double xi, double Ibar, double z0, double h, int nodes, int npad, int& n, int& wb0) {
  const TreePole t = tree_pole(xi, z0, h, nodes - 1, npad);
  n = (int)(-t.un);
  wb0 = t.wb0;
  // window keys packed for the node sweep: level-0 group gn (10 bits) | level-1 window start (8 bits) | level-2 window start
  return make_float4(t.un, t.ndh, (float)Ibar, __int_as_float((n >> 4) | (t.wb0 << 10) | (t.w2 << 18)));
}
#endif

}  // namespace tsff
