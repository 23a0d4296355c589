// tsff_pv_kernels.cuh -- the two O(poles x nodes) kernels shared by every mode.
//
//   k_pv_poles : thread-owns-pole forward sweep   I(xi_p), dI/dxi_p          (1 MUFU.LG2 per pair)
//   k_pv_nodes : thread-owns-node adjoint sweep   Dbar_i = sum_p Ibar_p g ln|g|  (1 MUFU.LG2 per pair)
//
// Bound: MUFU (XU) pipe, 16 lanes/clk/SM; the FP32 FMA pipe carries 3-4 ops per pair beside it.
// Shared memory: the pole-independent weights D (<= 16 KB for 4096 nodes) staged by one TMA bulk copy;
// pole descriptors for the adjoint sweep staged in 8 KB chunks.
#pragma once
#include "tsff_common.cuh"

namespace tsff {

constexpr int kPvThreads = 256;

struct PvPolesArgs {
  const float* D;        // [B][npad] weights (FP32)
  const double* D64;     // [B][npad] weights (FP64 validation path) or nullptr
  const double* pend;    // [B][2]   (p_0, p_M)
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
    pv_accumulate<R, true>(sD, a.npad / kPvBlk, (float)a.h, u0, nd, accI, accJ);
  } else {
    pv_accumulate_f64<R, true>(a.D64 + b * a.npad, a.nodes, a.h, g0d, accI, accJ);
  }
  const double p0 = a.pend[2 * b], pM = a.pend[2 * b + 1];
#pragma unroll
  for (int r = 0; r < R; r++) {
    int p = tile * (kPvThreads * R) + r * kPvThreads + threadIdx.x;
    if (p < a.P) {
      double I, dI;
      pv_finish(accI[r], accJ[r], p0, pM, g0d[r], g0d[r] + (double)(a.nodes - 1) * a.h, I, dI);
      a.outI[b * a.P + p] = I;
      if (a.outdI) a.outdI[b * a.P + p] = dI;
    }
  }
}

struct PvNodesArgs {
  const float4* desc;  // [B][P]  (u0 = -n_p, nd = -delta_p, Ibar_p, unused)
  int P, npad, ntiles;
  float h;
  double* Dbar;        // [B][npad]   Dbar_i = sum_p Ibar_p g_{p,i} ln|g_{p,i}|
};

constexpr int kNodeChunk = 512;  // poles staged per shared-memory chunk (8 KB)

template <int R>
__global__ void __launch_bounds__(kPvThreads) k_pv_nodes(const PvNodesArgs a) {
  __shared__ float4 sdesc[kNodeChunk];
  const long long b = blockIdx.x / a.ntiles;
  const int tile = blockIdx.x % a.ntiles;
  const int i0 = (tile * kPvThreads + threadIdx.x) * R;
  const float fi0 = (float)i0;
  const float4* desc = a.desc + b * a.P;
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) acc[r] = 0.0;
  for (int c0 = 0; c0 < a.P; c0 += kNodeChunk) {
    const int nc = min(kNodeChunk, a.P - c0);
    __syncthreads();
    for (int k = threadIdx.x; k < kNodeChunk; k += kPvThreads)
      sdesc[k] = (k < nc) ? desc[c0 + k] : make_float4(0.f, 1.f, 0.f, 0.f);
    __syncthreads();
    for (int s0 = 0; s0 < kNodeChunk; s0 += 64) {
      if (s0 >= nc) break;
      float part[R];
#pragma unroll
      for (int r = 0; r < R; r++) part[r] = 0.f;
#pragma unroll 8
      for (int k = 0; k < 64; k++) {
        const float4 d = sdesc[s0 + k];
        const float gbase = fmaf(fi0 + d.x, a.h, d.y);
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float g = fmaf((float)r, a.h, gbase);
          const float l = lg2_approx(fmaxf(fabsf(g), kTinyG));
          part[r] = fmaf(d.z, g * l, part[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) acc[r] += (double)part[r];
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++)
    if (i0 + r < a.npad) a.Dbar[b * a.npad + i0 + r] = kLn2 * acc[r];
}
#endif

}  // namespace tsff
