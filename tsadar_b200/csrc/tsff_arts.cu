// tsff_arts.cu -- the ARTS angular contraction of FitModel.electron_spectrum, spectype "angular_full"
// (tsadar/core/physics/generate_spectra.py:193-197, 210-216), forward and adjoint, hand-written (no cuBLAS):
//
//     modlE[r][j] = jmul[j] * sum_a weights[r][a] * (1/G) sum_g formfactor[g][j][a]          r < NA (1024), j < W, a < A (241)
//     ff_bar[g][j][a] = (jmul[j] / G) * sum_r modl_bar[r][j] * weights[r][a]
//
// 5e8 FP64 multiply-adds at the arts-1d shape (1024 x 2048 x 241): a DFMA-pipe problem (28 us at the B200's 64 DFMA per
// clock per SM); the operands (4 MB of formfactor, 2 MB of weights) sit in L2.  One kernel serves both directions: a
// shared-memory tiled C[M][N] = rowscale[m] * colscale[n] * sum_k A(m, k) B(k, n) with element strides for A and B, 64 x 64
// output tiles, 16-deep k slices, 4 x 4 outputs per thread (16 DFMA per 8 shared loads).  The mean over the G gradient
// points is folded into the B (forward) load and into the C (adjoint) store.  FP64 throughout (the oracle comparison is
// 1e-12); the k sum runs in index order per output, so results are deterministic.
#include "tsff_common.cuh"

using namespace tsff;

namespace {
constexpr int kTM = 64, kTN = 64, kTK = 16, kThreads = 256;

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

__global__ void __launch_bounds__(kThreads) k_arts_gemm(const GemmArgs p) {
  __shared__ double sA[kTK][kTM + 1];
  __shared__ double sB[kTK][kTN + 1];
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;       // thread -> outputs (m0 + ty + 16 i, n0 + tx + 16 j)
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
  // loaders walk the unit-stride dimension of each operand with consecutive threads
  const bool a_kfast = p.a_k == 1, b_kfast = p.b_k == 1;
  for (int k0 = 0; k0 < p.K; k0 += kTK) {
    for (int e = threadIdx.x; e < kTM * kTK; e += kThreads) {
      const int kk = a_kfast ? e % kTK : e / kTM, mm = a_kfast ? e / kTK : e % kTM;
      const int m = m0 + mm, k = k0 + kk;
      sA[kk][mm] = (m < p.M && k < p.K) ? p.A[m * p.a_m + k * p.a_k] : 0.0;
    }
    for (int e = threadIdx.x; e < kTN * kTK; e += kThreads) {
      const int kk = b_kfast ? e % kTK : e / kTN, nn = b_kfast ? e / kTK : e % kTN;
      const int n = n0 + nn, k = k0 + kk;
      double v = 0.0;
      if (n < p.N && k < p.K) {
        const double* src = p.B + k * p.b_k + n * p.b_n;
        for (int g = 0; g < p.b_G; g++) v += src[g * p.b_g];
        v *= p.b_scale;
      }
      sB[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; kk++) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) av[i] = sA[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; j++) bv[j] = sB[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int m = m0 + ty + 16 * i;
    if (m >= p.M) continue;
    const double rs = (p.rowscale ? p.rowscale[m] : 1.0) * p.c_scale;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int n = n0 + tx + 16 * j;
      if (n >= p.N) continue;
      const double v = acc[i][j] * rs * (p.colscale ? p.colscale[n] : 1.0);
      for (int g = 0; g < p.c_G; g++) p.C[g * p.c_g + m * p.c_m + n * p.c_n] = v;
    }
  }
}

int launch(const GemmArgs& p, cudaStream_t st) {
  dim3 grid((unsigned)((p.N + kTN - 1) / kTN), (unsigned)((p.M + kTM - 1) / kTM));
  k_arts_gemm<<<grid, kThreads, 0, st>>>(p);
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
