// tsff_arts.cu -- the ARTS angular contraction of FitModel.electron_spectrum, spectype "angular_full"
// (tsadar/core/physics/generate_spectra.py:193-197, 210-216), forward and adjoint, hand-written (no cuBLAS):
//
//     modlE[r][j] = jmul[j] * sum_a weights[r][a] * (1/G) sum_g formfactor[g][j][a]          r < NA (1024), j < W, a < A (241)
//     ff_bar[g][j][a] = (jmul[j] / G) * sum_r modl_bar[r][j] * weights[r][a]
//
// 5e8 FP64 multiply-adds at the arts-1d shape (1024 x 2048 x 241): a DFMA-pipe problem (28 us at the B200's 64 DFMA per
// clock per SM); the operands (4 MB of formfactor, 2 MB of weights) sit in L2.  One kernel serves both directions: a
// shared-memory tiled C[M][N] = rowscale[m] * colscale[n] * sum_k A(m, k) B(k, n) with element strides for A and B, 128 x 64
// (or 64 x 32) output tiles, 16-deep k slices, register prefetch of the next slice.  The mean over the G gradient
// points is folded into the B (forward) load and into the C (adjoint) store.  FP64 throughout (the oracle comparison is
// 1e-12); the k sum runs in index order per output, so results are deterministic.
#include "tsff_common.cuh"

using namespace tsff;

namespace {
constexpr int kTK = 16, kThreads = 256;

struct GemmArgs {
  int M, N, K;
  const double* A; long long a_m, a_k;        // A(m, k) = A[m * a_m + k * a_k]
  const double* B; long long b_k, b_n;        // B(k, n) = sum_{g < b_G} B[g * b_g + k * b_k + n * b_n] * b_scale
  int b_G; long long b_g; double b_scale;
  const double* rowscale;                      // [M] or null
  const double* colscale;                      // [N] or null
  double* C; long long c_m, c_n;              // C[g * c_g + m * c_m + n * c_n] for g < c_G (the same value to every g)
  int c_G; long long c_g; double c_scale;
};

// BM x BN output tile per CTA, (BM/16) x (BN/16) CONSECUTIVE outputs per thread (operands come out of shared memory as 128-bit
// loads: 6 LDS.128 per 32 DFMA at 128 x 64), 16-deep k slices.  The next slice is fetched from global memory into registers
// while the current one is multiplied (one shared buffer, two barriers per slice): the first version of this kernel waited for
// L2 every slice and ran at 7 % of the FP64 pipe.
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src, bool ok) {
  // 8-byte asynchronous copy global -> shared; src-size 0 writes zeros (out-of-range elements)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(ok ? 8 : 0) : "memory");
}

// KFAST: both operands have k as their unit-stride index (forward) or neither has (adjoint); G1: one gradient point (no mean /
// replication loops).  Compile-time, so that each instantiation carries one loader and stays small.
// G1: operand slices go global -> shared with cp.async into a two-stage ring (no registers held across the multiply loop: with
// register prefetch ptxas sank the loads to just before their shared-memory stores and every slice waited for L2 again);
// the operand scale moves to the epilogue.  !G1 (mean over gradient points folded into the B load): register prefetch.
template <int BM, int BN, bool KFAST, bool G1>
__global__ void __launch_bounds__(kThreads, 2) k_arts_gemm(const GemmArgs p) {
  constexpr int RM = BM / 16, RN = BN / 16;            // outputs per thread
  constexpr int LA = BM * kTK / kThreads, LB = BN * kTK / kThreads;   // elements each thread stages per slice
  constexpr int PA = BM + 2, PB = BN + 2;              // padded rows (16-byte aligned, staggered banks)
  constexpr int STAGE = kTK * (PA + PB);               // doubles per ring stage
  extern __shared__ __align__(16) double smem[];      // [G1 ? 2 : 1][STAGE]
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;       // thread -> outputs (m0 + ty * RM + i, n0 + tx * RN + j)
  double acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; i++)
#pragma unroll
    for (int j = 0; j < RN; j++) acc[i][j] = 0.0;
  // gridDim.z = 1 or 2 halves of the k range (whole slices each); with two, both add atomically onto a zeroed C: a sum of two
  // terms does not depend on their order, so the result stays deterministic
  const int kslices = (p.K + kTK - 1) / kTK, zper = (kslices + gridDim.z - 1) / gridDim.z;
  const int kbeg = blockIdx.z * zper * kTK, kend = min(p.K, (int)(blockIdx.z + 1) * zper * kTK);
  // loaders walk the unit-stride dimension of each operand with consecutive threads
  auto issue = [&](int k0, int buf) {                 // G1: cp.async of slice k0 into ring stage buf
    double* sa = smem + buf * STAGE;
    double* sb = sa + kTK * PA;
#pragma unroll
    for (int q = 0; q < LA; q++) {
      const int e = threadIdx.x + q * kThreads;
      const int kk = KFAST ? e % kTK : e / BM, mm = KFAST ? e / kTK : e % BM;
      const int m = m0 + mm, k = k0 + kk;
      const bool ok = m < p.M && k < kend;
      cp_async8(&sa[kk * PA + mm], ok ? p.A + m * p.a_m + k * p.a_k : p.A, ok);
    }
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int e = threadIdx.x + q * kThreads;
      const int kk = KFAST ? e % kTK : e / BN, nn = KFAST ? e / kTK : e % BN;
      const int n = n0 + nn, k = k0 + kk;
      const bool ok = n < p.N && k < kend;
      cp_async8(&sb[kk * PB + nn], ok ? p.B + k * p.b_k + n * p.b_n : p.B, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double ra[LA], rb[LB];
  auto fetch = [&](int k0) {                          // !G1: slice k0 into registers
#pragma unroll
    for (int q = 0; q < LA; q++) {
      const int e = threadIdx.x + q * kThreads;
      const int kk = KFAST ? e % kTK : e / BM, mm = KFAST ? e / kTK : e % BM;
      const int m = m0 + mm, k = k0 + kk;
      ra[q] = (m < p.M && k < kend) ? p.A[m * p.a_m + k * p.a_k] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int e = threadIdx.x + q * kThreads;
      const int kk = KFAST ? e % kTK : e / BN, nn = KFAST ? e / kTK : e % BN;
      const int n = n0 + nn, k = k0 + kk;
      double v = 0.0;
      if (n < p.N && k < kend) {
        const double* src = p.B + k * p.b_k + n * p.b_n;
        for (int g = 0; g < p.b_G; g++) v += src[g * p.b_g];
      }
      rb[q] = v;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int q = 0; q < LA; q++) {
      const int e = threadIdx.x + q * kThreads;
      smem[(KFAST ? e % kTK : e / BM) * PA + (KFAST ? e / kTK : e % BM)] = ra[q];
    }
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int e = threadIdx.x + q * kThreads;
      smem[kTK * PA + (KFAST ? e % kTK : e / BN) * PB + (KFAST ? e / kTK : e % BN)] = rb[q];
    }
  };
  int buf = 0;
  if (G1) issue(kbeg, 0);
  else fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += kTK) {
    const bool more = k0 + kTK < kend;
    if (G1) {
      if (more) {
        issue(k0 + kTK, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    } else {
      stage();
    }
    __syncthreads();
    if (!G1 && more) fetch(k0 + kTK);
    const double* sa = smem + (G1 ? buf * STAGE : 0);
    const double* sb = sa + kTK * PA;
#pragma unroll
    for (int kk = 0; kk < kTK; kk++) {
      double av[RM], bv[RN];
#pragma unroll
      for (int i = 0; i < RM; i += 2) {
        const double2 t = *reinterpret_cast<const double2*>(&sa[kk * PA + ty * RM + i]);
        av[i] = t.x; av[i + 1] = t.y;
      }
#pragma unroll
      for (int j = 0; j < RN; j += 2) {
        const double2 t = *reinterpret_cast<const double2*>(&sb[kk * PB + tx * RN + j]);
        bv[j] = t.x; bv[j + 1] = t.y;
      }
#pragma unroll
      for (int i = 0; i < RM; i++)
#pragma unroll
        for (int j = 0; j < RN; j++) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < RM; i++) {
    const int m = m0 + ty * RM + i;
    if (m >= p.M) continue;
    const double rs = (p.rowscale ? p.rowscale[m] : 1.0) * p.c_scale * p.b_scale;
#pragma unroll
    for (int j = 0; j < RN; j++) {
      const int n = n0 + tx * RN + j;
      if (n >= p.N) continue;
      const double v = acc[i][j] * rs * (p.colscale ? p.colscale[n] : 1.0);
      const int cG = G1 ? 1 : p.c_G;
      if (gridDim.z == 1) {
        for (int g = 0; g < cG; g++) p.C[g * p.c_g + m * p.c_m + n * p.c_n] = v;
      } else {
        for (int g = 0; g < cG; g++) atomicAdd(&p.C[g * p.c_g + m * p.c_m + n * p.c_n], v);
      }
    }
  }
}

template <int BM, int BN>
constexpr size_t gemm_smem(bool g1) { return (size_t)(g1 ? 2 : 1) * kTK * (BM + 2 + BN + 2) * sizeof(double); }

int launch(const GemmArgs& p, cudaStream_t st) {
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // the large tile when it still gives every SM a CTA; else 64 x 64 tiles, the k range in two halves when that is what fills
  // the device (the adjoint at the arts-1d shape: 2048 x 241 outputs, k = 1024)
  const long long big = (long long)((p.N + 63) / 64) * ((p.M + 127) / 128);
  const bool kfast = p.a_k == 1 && p.b_k == 1, g1 = p.b_G == 1 && p.c_G == 1;
  if (!kfast && (p.a_k == 1 || p.b_k == 1)) { set_error("k_arts_gemm: mixed operand layouts"); return TSFF_E_INVALID; }
#define TSFF_GEMM_ONE(BM_, BN_, KF_, G1_, grid)                                                          \
  do {                                                                                                   \
    TSFF_SMEM_OPTIN((k_arts_gemm<BM_, BN_, KF_, G1_>));                                                  \
    k_arts_gemm<BM_, BN_, KF_, G1_><<<grid, kThreads, gemm_smem<BM_, BN_>(G1_), st>>>(p);                \
  } while (0)
#define TSFF_GEMM(BM_, BN_, grid)                                                               \
  do {                                                                                          \
    if (kfast && g1) TSFF_GEMM_ONE(BM_, BN_, true, true, grid);                                 \
    else if (kfast) TSFF_GEMM_ONE(BM_, BN_, true, false, grid);                                 \
    else if (g1) TSFF_GEMM_ONE(BM_, BN_, false, true, grid);                                    \
    else TSFF_GEMM_ONE(BM_, BN_, false, false, grid);                                           \
  } while (0)
  if (big >= sms) {
    dim3 grid((unsigned)((p.N + 63) / 64), (unsigned)((p.M + 127) / 128));
    TSFF_GEMM(128, 64, grid);
  } else {
    const long long tiles = (long long)((p.N + 63) / 64) * ((p.M + 63) / 64);
    const unsigned split = (tiles < 2LL * sms && p.K >= 8 * kTK) ? 2u : 1u;
    if (split == 2) TSFF_CUDA_OK(cudaMemsetAsync(p.C, 0, (size_t)p.M * p.N * p.c_G * sizeof(double), st));
    dim3 grid((unsigned)((p.N + 63) / 64), (unsigned)((p.M + 63) / 64), split);
    TSFF_GEMM(64, 64, grid);
  }
#undef TSFF_GEMM
#undef TSFF_GEMM_ONE
  TSFF_LAUNCH_OK("k_arts_gemm");
  return TSFF_OK;
}
}  // namespace

extern "C" int tsff_arts_weights_fwd(const double* ff, int32_t G, int32_t W, int32_t A, const double* weights, int32_t NA,
                                     const double* jmul, double* modl, void* stream) {
  if (!ff || !weights || !modl || G < 1 || W < 1 || A < 1 || NA < 1) { set_error("tsff_arts_weights_fwd: bad argument"); return TSFF_E_INVALID; }
  GemmArgs p;
  memset(&p, 0, sizeof(p));
  p.M = NA; p.N = W; p.K = A;                                    // m = image row r, n = wavelength j, k = angle a
  p.A = weights; p.a_m = A; p.a_k = 1;
  p.B = ff; p.b_k = 1; p.b_n = A; p.b_G = G; p.b_g = (long long)W * A; p.b_scale = 1.0 / (double)G;   // mean over gradient points (:193)
  p.colscale = jmul;                                             // IAW filter (:210-216)
  p.C = modl; p.c_m = W; p.c_n = 1; p.c_G = 1; p.c_g = 0; p.c_scale = 1.0;
  return launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int tsff_arts_weights_bwd(const double* modl_bar, int32_t G, int32_t W, int32_t A, const double* weights, int32_t NA,
                                     const double* jmul, double* ff_bar, void* stream) {
  if (!modl_bar || !weights || !ff_bar || G < 1 || W < 1 || A < 1 || NA < 1) { set_error("tsff_arts_weights_bwd: bad argument"); return TSFF_E_INVALID; }
  GemmArgs p;
  memset(&p, 0, sizeof(p));
  p.M = W; p.N = A; p.K = NA;                                    // m = wavelength j, n = angle a, k = image row r
  p.A = modl_bar; p.a_m = 1; p.a_k = W;
  p.B = weights; p.b_k = A; p.b_n = 1; p.b_G = 1; p.b_g = 0; p.b_scale = 1.0;
  p.rowscale = jmul;
  p.C = ff_bar; p.c_m = A; p.c_n = 1; p.c_G = G; p.c_g = (long long)W * A; p.c_scale = 1.0 / (double)G;
  return launch(p, static_cast<cudaStream_t>(stream));
}
