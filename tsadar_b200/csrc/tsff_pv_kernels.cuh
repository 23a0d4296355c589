// tsff_pv_kernels.cuh -- the two O(poles x nodes) kernels shared by every mode.
//
//   k_pv_poles : thread-owns-pole forward sweep   I(xi_p), dI/dxi_p          (1 MUFU.RCP per pair)
//   k_pv_nodes : thread-owns-node adjoint sweep   pbar_i = sum_p Ibar_p W(g_{p,i})  (1 MUFU.RCP per pair)
//
// Bound: MUFU (XU) pipe, 16 lanes/clk/SM; the FP32 FMA pipe carries 3-4 ops per pair beside it.
// Shared memory: the pole-independent weights D (<= 16 KB for 4096 nodes) staged by one TMA bulk copy;
// pole descriptors for the adjoint sweep staged in 8 KB chunks.
#pragma once
#include "tsff_common.cuh"

namespace tsff {

constexpr int kPvThreads = 256;

struct PvPolesArgs {
  const float* D;        // [B][npad] far-field node weights p_i*h (FP32; zero at i = 0, i >= M)
  const double* D64;     // [B][npad] log-form weights D_i (FP64 validation path) or nullptr
  const double* pend;    // [B][2]   (p_0, p_M)
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
  float* sD = reinterpret_cast<float*>(smem_raw);
  const long long b = blockIdx.x / a.ntiles;
  const int tile = blockIdx.x % a.ntiles;
  if (PREC == TSFF_PV_FP32) stage_bulk(sD, a.D + b * a.npad, (uint32_t)a.npad * 4u, &bar);

  const double* poles = a.poles + b * a.pole_bstride;
  double xi[R];
  float u0[R], nd[R];
  double g0d[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    int p = tile * (kPvThreads * R) + r * kPvThreads + threadIdx.x;
    xi[r] = poles[p < a.P ? p : a.P - 1];
    pole_split(xi[r], a.z0, a.h, a.nodes, u0[r], nd[r]);
    g0d[r] = a.z0 - xi[r];
  }
  double accI[R], accJ[R];
  if (PREC == TSFF_PV_FP32) {
    pv_accumulate<R, true>(sD, a.npad / kPvBlk, far_coef(a.h), u0, nd, accI, accJ);
  } else {
    pv_accumulate_f64<R, true>(a.D64 + b * a.npad, a.nodes, a.h, g0d, accI, accJ);
  }
  const double p0 = a.pend[2 * b], pM = a.pend[2 * b + 1];
#pragma unroll
  for (int r = 0; r < R; r++) {
    int p = tile * (kPvThreads * R) + r * kPvThreads + threadIdx.x;
    if (p < a.P) {
      double I, dI;
      if (PREC == TSFF_PV_FP32) {
        const double* pn = a.pnodes + b * a.pnode_stride;
        pv_near_exact(xi[r], a.z0, a.h, a.nodes, [pn](int i) { return pn[i]; }, I, dI);
        I += accI[r];
        dI += accJ[r];
      } else {
        pv_finish(accI[r], accJ[r], p0, pM, g0d[r], g0d[r] + (double)(a.nodes - 1) * a.h, I, dI);
      }
      a.outI[b * a.P + p] = I;
      if (a.outdI) a.outdI[b * a.P + p] = dI;
    }
  }
}

struct PvNodesArgs {
  const float4* desc;  // [B][P]  (u0 = -n_p, nd = -delta_p, Ibar_p * h, unused)
  int P, nodes, npad, ntiles;
  float h;
  double* pbar;        // [B][npad]  far-field part of d loss / d p_i for interior nodes 1..M-1 (0 elsewhere)
};

constexpr int kNodeChunk = 512;  // poles staged per shared-memory chunk (8 KB)

// pbar_i (far part) = sum_p Ibar_p W(g_{p,i}) over poles with |i - n_p| > kNearHalf.
// Poles are staged in shared memory in chunks; for every sub-chunk of 64 poles the range of nearest-node indices is
// recorded, so that a warp whose nodes are farther than kMidHalf from all of them runs the short, unmasked series
// (5.25 FMA-pipe ops + 1 MUFU.RCP per pair: MUFU-bound), and only the sub-chunks next to the warp's nodes pay for the
// long series and the near-node mask.
template <int R>
__global__ void __launch_bounds__(kPvThreads) k_pv_nodes(const PvNodesArgs a) {
  __shared__ float4 sdesc[kNodeChunk];
  __shared__ float srange[kNodeChunk / 64][2];  // (max u0, min u0) = (-n_min, -n_max) per 64-pole sub-chunk
  const long long b = blockIdx.x / a.ntiles;
  const int tile = blockIdx.x % a.ntiles;
  const int i0 = (tile * kPvThreads + threadIdx.x) * R;
  const float fi0 = (float)i0;
  const float c2 = a.h * a.h * (1.f / 6.f), c4 = a.h * a.h * a.h * a.h * (1.f / 15.f);
  const float4* desc = a.desc + b * a.P;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) acc[r] = 0.0;
  for (int c0 = 0; c0 < a.P; c0 += kNodeChunk) {
    const int nc = min(kNodeChunk, a.P - c0);
    __syncthreads();
    for (int k = threadIdx.x; k < kNodeChunk; k += kPvThreads)
      sdesc[k] = (k < nc) ? desc[c0 + k] : make_float4(1e6f, 0.f, 0.f, 0.f);
    __syncthreads();
    // one warp per sub-chunk: range of u0 over its real poles
    for (int sc = wid; sc < kNodeChunk / 64; sc += kPvThreads / 32) {
      float hi = -3.0e38f, lo = 3.0e38f;
      for (int k = lane; k < 64; k += 32) {
        const int kk = sc * 64 + k;
        if (kk < nc) { const float v = sdesc[kk].x; hi = fmaxf(hi, v); lo = fminf(lo, v); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      }
      if (lane == 0) { srange[sc][0] = hi; srange[sc][1] = lo; }
    }
    __syncthreads();
    for (int s0 = 0; s0 < kNodeChunk; s0 += 64) {
      if (s0 >= nc) break;
      // offsets i - n_p of this thread's nodes against the sub-chunk's poles lie in [fi0 + lo, fi0 + R-1 + hi]
      const float omin = fi0 + srange[s0 / 64][1], omax = fi0 + (float)(R - 1) + srange[s0 / 64][0];
      const bool mid = (omin <= (float)kMidHalf) && (omax >= -(float)kMidHalf);
      float part[R];
#pragma unroll
      for (int r = 0; r < R; r++) part[r] = 0.f;
      if (!__any_sync(0xffffffffu, mid)) {
#pragma unroll 8
        for (int k = 0; k < 64; k++) {
          const float4 d = sdesc[s0 + k];
          const float gbase = fmaf(fi0 + d.x, a.h, d.y);  // g at node i0
#pragma unroll
          for (int r = 0; r < R; r++) {
            const float g = fmaf((float)r, a.h, gbase);
            const float rg = rcp_approx(g);
            part[r] = fmaf(d.z * rg, fmaf(rg * rg, c2, 1.f), part[r]);   // Ibar h * W / h
          }
        }
      } else {
#pragma unroll 4
        for (int k = 0; k < 64; k++) {
          const float4 d = sdesc[s0 + k];
          const float u = fi0 + d.x;              // i0 - n_p, exact
          const float gbase = fmaf(u, a.h, d.y);
#pragma unroll
          for (int r = 0; r < R; r++) {
            const float g = fmaf((float)r, a.h, gbase);
            const bool far = fabsf(u + (float)r) > (float)kNearHalf + 0.5f;
            const float rg = far ? rcp_approx(g) : 0.f;
            const float s2 = rg * rg;
            part[r] = fmaf(d.z, rg * fmaf(fmaf(s2, c4, c2), s2, 1.f), part[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) acc[r] += (double)part[r];
    }
  }
  const int M = a.nodes - 1;
#pragma unroll
  for (int r = 0; r < R; r++)
    if (i0 + r < a.npad) a.pbar[b * a.npad + i0 + r] = (i0 + r >= 1 && i0 + r <= M - 1) ? acc[r] : 0.0;
}

// Exact (FP64) adjoint contributions of one pole: the kNearHalf nodes either side of it and the two end nodes
// (whose weights are first differences plus the explicit endpoint terms of I).  Atomically added to pnear[0..M].
__device__ __forceinline__ void pv_bwd_pole_exact(double xi, double Ibar, double z0, double h, int nodes, double* pnear) {
  if (Ibar == 0.0) return;
  const int M = nodes - 1;
  double rn = rint((xi - z0) / h);
  if (!(rn >= 0.0)) rn = 0.0;
  if (rn > (double)M) rn = (double)M;
  const int n = (int)rn;
  const int lo = max(1, n - kNearHalf), hi = min(M - 1, n + kNearHalf);
  const double ih = 1.0 / h;
  if (lo <= hi) {
    double pm = pv_phi(z0 + (double)(lo - 1) * h - xi), pc = pv_phi(z0 + (double)lo * h - xi);
    for (int i = lo; i <= hi; i++) {
      const double pp = pv_phi(z0 + (double)(i + 1) * h - xi);
      atomicAdd(&pnear[i], Ibar * (pp - 2.0 * pc + pm) * ih);
      pm = pc;
      pc = pp;
    }
  }
  const double g0 = z0 - xi, gM = z0 + (double)M * h - xi;
  atomicAdd(&pnear[0], Ibar * ((pv_phi(g0 + h) - pv_phi(g0)) * ih - 1.0 - log(fmax(fabs(g0), 1e-300))));
  atomicAdd(&pnear[M], Ibar * ((pv_phi(gM - h) - pv_phi(gM)) * ih + 1.0 + log(fmax(fabs(gM), 1e-300))));
}
#endif

}  // namespace tsff
