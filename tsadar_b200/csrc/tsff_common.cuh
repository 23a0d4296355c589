// tsff_common.cuh -- context layout, error handling, small device helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tsff.h"
#include "tsff_math.cuh"
#include "tsff_pv.cuh"

namespace tsff {

constexpr int kXi1N = 1024;      // form_factor.py:130,137
constexpr int kXi2N = 1640;      // form_factor.py:138  arange(-8.2, 8.2, 0.01)
constexpr double kXiMinMax = 8.2;
constexpr double kFillLog = -50.0;  // extrap=[-50,-50]  form_factor.py:256,263

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

void set_error(const char* fmt, ...);

#define TSFF_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      tsff::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return TSFF_E_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

// RAII: make `device` current for the duration of a C-ABI call and restore the caller's device afterwards (a process may
// hold contexts on several GPUs; the library must neither depend on nor change the caller's current device).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { ok = false; return; }
    if (cur != device) {
      ok = cudaSetDevice(device) == cudaSuccess;
      if (ok) prev = cur;
    }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define TSFF_ON_DEVICE(ctx)                                                                  \
  tsff::DeviceGuard _dev_guard((ctx)->device);                                               \
  if (!_dev_guard.ok) { tsff::set_error("cannot make device %d current", (ctx)->device); return TSFF_E_CUDA; }

// Opt a kernel in to large dynamic shared memory.  Always raise the cap to the device maximum: the attribute is a
// per-function CAP, so setting it to "what this launch needs" would make a later, larger launch fail.
#define TSFF_SMEM_OPTIN(kernel)                                                                          \
  TSFF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))

#define TSFF_LAUNCH_OK(name)                                                                 \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      tsff::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));              \
      return TSFF_E_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

// ---- 1-D TMA bulk copy global -> shared (cp.async.bulk, SASS UBLKCP) completed on an mbarrier ----------
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// Stage `bytes` (multiple of 16, 16-byte aligned both sides) from global into shared with one bulk copy.
// Must be called by all threads of the CTA; `bar` is a shared 8-byte slot.  One use per kernel (parity 0).
__device__ __forceinline__ void stage_bulk(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes);
    bulk_g2s(dst_smem, src_gmem, bytes, bar);
  }
  mbar_wait(bar, 0);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Segmented warp reduction over contiguous runs of lanes with equal `key` (scatter-adds whose target index changes slowly
// along the warp: one atomic per run instead of one per lane).  run_head: lane index of the first lane of this lane's run;
// run_sum: suffix sums within runs -- the head lane ends up with the run's total.  All 32 lanes must call both.
__device__ __forceinline__ int run_head(int key) {
  const int lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync(0xffffffffu, key, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != key);
  return 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
}
__device__ __forceinline__ double run_sum(double v, int head) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double vo = __shfl_down_sync(0xffffffffu, v, o);
    const int ho = __shfl_down_sync(0xffffffffu, head, o);
    if (lane + o < 32 && ho == head) v += vo;
  }
  return v;
}

// Sum `n` per-thread values across the CTA and atomically add the totals to dst[0..n).  sred: >= n*nwarps doubles.
template <int NW>
__device__ __forceinline__ void block_accumulate(const double* vals, int n, double* sred, double* dst) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < n; k++) {
    double s = warp_sum(vals[k]);
    if (lane == 0) sred[k * NW + wid] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; w++) s += sred[k * NW + w];
    if (s != 0.0) atomicAdd(&dst[k], s);
  }
  __syncthreads();
}
// The same per warp, no CTA barrier: each warp adds its own totals (warps of a CTA then finish independently).
__device__ __forceinline__ void warp_accumulate(const double* vals, int n, double* dst) {
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < n; k++) {
    const double s = warp_sum(vals[k]);
    if (lane == 0 && s != 0.0) atomicAdd(&dst[k], s);
  }
}
#endif  // __CUDACC__

}  // namespace tsff

// ---- the context -----------------------------------------------------------------------------------------
struct tsff_ctx {
  int device;
  int sm_count;
  int mode, W, A, G, I, V, NP, pv_precision;
  double lam_min, lam_max, lam_shift, v0, dv;
  // device-resident static tables (one allocation)
  void* dev_blob;
  double* omgs;   // [W]   2e7 pi C / linspace(lam_min, lam_max, W)       form_factor.py:132-135
  double* lam_nm; // [W]   wavelength axis in nm (= lams*1e7)
  double* costh;  // [A]   cos(sa)
  double* sinth;  // [A]   sin(sa)   (2V mode: ks as a vector, form_factor.py:514)
  double ud_angle_deg, va_angle_deg;
  double* wts;    // [A]
  double* jmul;   // [W]
  double* zr;     // [1640] Zpi[0]                                        form_factor.py:139
  double* zi;     // [1640] Zpi[1]
  double* xi2;    // [1640]
  tsff::ZTab zt;
  // table mode: nodes xi1 (uniform), M = 1022
  double xi1_0, xi1_h;
  // PV geometry for the active mode
  int pv_nodes;  // M+1 nodes used by ratintn (N-1)
  int pv_npad;   // padded to whole 64-node blocks, at least three (tsff_tree.cuh)
  double pv_z0, pv_h;
  double* tstat; // static expansion tables (k_tree_static) for pv_nodes
  cudaEvent_t ev[4];  // optional profile events (fwd start/stop, bwd start/stop)
  int* cells;         // frozen lerp cells of the table-mode assembly (tsff_ctx_set_frozen_cells), or null
  int cell_mode;      // 0 off, 1 record, 2 replay
  long long cells_B;  // lineouts the cells buffer holds
  int tune_fwd_r4;    // TSFF_FWD_R4 tuning switch, read once at creation (four poles per thread in k_direct_fwd: measured slower)
};
