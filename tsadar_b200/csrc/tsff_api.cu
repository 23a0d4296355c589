// tsff_api.cu -- C ABI entry points (include/tsff.h): mode dispatch and the stand-alone principal-value
// integral (boundary B1: vmap(ratintn), ratintn.py:4-23).
#include "tsff_pv_kernels.cuh"

using namespace tsff;

namespace tsff {
size_t direct_saved_bytes(const tsff_ctx*, int64_t);
size_t direct_ws_bytes(const tsff_ctx*, int64_t);
int direct_fwd(tsff_ctx*, int64_t, const double*, const void*, int, double*, double*, void*, void*, cudaStream_t);
int direct_bwd(tsff_ctx*, int64_t, const double*, const void*, int, const void*, const double*, const double*, double*, void*,
               void*, cudaStream_t);
size_t table_saved_bytes(const tsff_ctx*, int64_t);
size_t table_ws_bytes(const tsff_ctx*, int64_t);
int table_fwd(tsff_ctx*, int64_t, const double*, const void*, int, double*, double*, void*, void*, cudaStream_t);
bool table_pair_compatible(const tsff_ctx*, const tsff_ctx*);
int table_pair_fwd(tsff_ctx*, tsff_ctx*, int64_t, const double*, const void*, int, double*, double*, void*, void*, void*, cudaStream_t);
int table_pair_bwd(tsff_ctx*, tsff_ctx*, int64_t, const double*, const void*, int, const void*, const void*, const double*, const double*,
                   double*, void*, void*, void*, cudaStream_t);
int table_bwd(tsff_ctx*, int64_t, const double*, const void*, int, const void*, const double*, const double*, double*, void*,
              void*, cudaStream_t);
int ff2v_fwd(tsff_ctx*, int64_t, const double*, const double*, double*, void*, cudaStream_t);
int ff2v_bwd(tsff_ctx*, int64_t, const double*, const double*, const void*, const double*, double*, double*, void*, cudaStream_t);
int chi2v_fwd(tsff_ctx*, const double*, const double*, const double*, const double*, int64_t, double*, cudaStream_t);
size_t ff2v_saved_bytes(const tsff_ctx*, int64_t);
size_t ff2v_ws_bytes(const tsff_ctx*, int64_t);
}  // namespace tsff

extern "C" size_t tsff_ff_saved_bytes(const tsff_ctx* c, int64_t B) {
  if (!c || B < 1) return 0;
  if (c->mode == TSFF_MODE_2V) return ff2v_saved_bytes(c, B);
  return c->mode == TSFF_MODE_TABLE ? table_saved_bytes(c, B) : direct_saved_bytes(c, B);
}
extern "C" size_t tsff_ff_workspace_bytes(const tsff_ctx* c, int64_t B) {
  if (!c || B < 1) return 0;
  if (c->mode == TSFF_MODE_2V) return ff2v_ws_bytes(c, B);
  return c->mode == TSFF_MODE_TABLE ? table_ws_bytes(c, B) : direct_ws_bytes(c, B);
}

// ---- second-order path: frozen lerp cells (table mode) -----------------------------------------------------------------------
extern "C" size_t tsff_ff_cells_bytes(const tsff_ctx* c, int64_t B) {
  if (!c || B < 1 || c->mode != TSFF_MODE_TABLE) return 0;
  return (size_t)B * c->G * c->W * c->A * kCellStride * sizeof(int);
}
extern "C" int tsff_ctx_set_frozen_cells(tsff_ctx* c, int mode, int32_t* cells, int64_t B) {
  if (!c) { set_error("null ctx"); return TSFF_E_INVALID; }
  if (mode == 0) { c->cells = nullptr; c->cell_mode = 0; c->cells_B = 0; return TSFF_OK; }
  if (c->mode != TSFF_MODE_TABLE) { set_error("frozen cells: table mode only"); return TSFF_E_INVALID; }
  if ((mode != 1 && mode != 2) || !cells || B < 1) { set_error("frozen cells: bad argument"); return TSFF_E_INVALID; }
  c->cells = cells; c->cell_mode = mode; c->cells_B = B;
  return TSFF_OK;
}

static int check_common(const tsff_ctx* c, int64_t B, const void* params, const void* fe, int fe_dtype, const void* saved,
                        const void* ws) {
  if (!c || !params || !fe || !saved || !ws) { set_error("null argument"); return TSFF_E_INVALID; }
  if (B < 1) { set_error("B must be >= 1"); return TSFF_E_INVALID; }
  if (fe_dtype != TSFF_F32 && fe_dtype != TSFF_F64) { set_error("fe_dtype must be TSFF_F32 or TSFF_F64"); return TSFF_E_INVALID; }
  if (c->cell_mode && B > c->cells_B) { set_error("frozen-cell buffer holds %lld lineouts, call has %lld", c->cells_B, (long long)B); return TSFF_E_INVALID; }
  const long long blocks = (long long)B * c->G * (((long long)c->W * c->A + 255) / 256);
  if (blocks > 0x7fffffffLL) { set_error("batch too large for one launch: split the call"); return TSFF_E_INVALID; }
  return TSFF_OK;
}

extern "C" int tsff_ff_fwd(tsff_ctx* c, int64_t B, const double* params, const void* fe, int fe_dtype, double* modl_out,
                           double* ff_out, void* saved, void* ws, void* stream) {
  if (c && B == 0) return TSFF_OK;   // an empty batch (vmap over zero lineouts) is a no-op, as in the reference
  int rc = check_common(c, B, params, fe, fe_dtype, saved, ws);
  if (rc) return rc;
  if (!modl_out && !ff_out) { set_error("no output requested"); return TSFF_E_INVALID; }
  TSFF_ON_DEVICE(c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->mode == TSFF_MODE_2V) {
    if (fe_dtype != TSFF_F64 || !ff_out || modl_out) { set_error("2V mode: fe must be float64 [B][V][V]; only ff_out is produced"); return TSFF_E_INVALID; }
    return ff2v_fwd(c, B, params, static_cast<const double*>(fe), ff_out, saved, st);
  }
  return c->mode == TSFF_MODE_TABLE ? table_fwd(c, B, params, fe, fe_dtype, modl_out, ff_out, saved, ws, st)
                                    : direct_fwd(c, B, params, fe, fe_dtype, modl_out, ff_out, saved, ws, st);
}

extern "C" int tsff_chi2v_fwd(tsff_ctx* c, const double* fe, const double* beta, const double* xie_mag, const double* klde_mag,
                              int64_t P, double* chi_out, void* stream) {
  if (!c) { set_error("null context"); return TSFF_E_INVALID; }
  if (c->mode != TSFF_MODE_2V) { set_error("tsff_chi2v_fwd needs a TSFF_MODE_2V context"); return TSFF_E_INVALID; }
  if (P == 0) return TSFF_OK;
  if (P < 0 || !fe || !beta || !xie_mag || !klde_mag || !chi_out) { set_error("null argument"); return TSFF_E_INVALID; }
  TSFF_ON_DEVICE(c);
  return chi2v_fwd(c, fe, beta, xie_mag, klde_mag, P, chi_out, static_cast<cudaStream_t>(stream));
}

extern "C" int tsff_ff_bwd(tsff_ctx* c, int64_t B, const double* params, const void* fe, int fe_dtype, const void* saved,
                           const double* modl_bar, const double* ff_bar, double* params_bar, void* fe_bar, void* ws,
                           void* stream) {
  if (c && B == 0) return TSFF_OK;
  int rc = check_common(c, B, params, fe, fe_dtype, saved, ws);
  if (rc) return rc;
  if (!modl_bar && !ff_bar) { set_error("no cotangent given"); return TSFF_E_INVALID; }
  if (!fe_bar || (!params_bar && c->mode != TSFF_MODE_2V)) { set_error("null output"); return TSFF_E_INVALID; }
  TSFF_ON_DEVICE(c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->mode == TSFF_MODE_2V) {
    if (fe_dtype != TSFF_F64 || !ff_bar || modl_bar) { set_error("2V mode: fe float64, cotangent of the formfactor only"); return TSFF_E_INVALID; }
    return ff2v_bwd(c, B, params, static_cast<const double*>(fe), saved, ff_bar, params_bar, static_cast<double*>(fe_bar), ws, st);
  }
  return c->mode == TSFF_MODE_TABLE
             ? table_bwd(c, B, params, fe, fe_dtype, saved, modl_bar, ff_bar, params_bar, fe_bar, ws, st)
             : direct_bwd(c, B, params, fe, fe_dtype, saved, modl_bar, ff_bar, params_bar, fe_bar, ws, st);
}

// ---- B3': two windows of one plasma ---------------------------------------------------------------------------------
static int check_pair(const tsff_ctx* a, const tsff_ctx* b) {
  if (!a || !b) { set_error("null context"); return TSFF_E_INVALID; }
  if (!table_pair_compatible(a, b)) {
    set_error("pair calls need two TSFF_MODE_TABLE contexts on one device with the same V / velocity grid, G, I, PV precision, frozen cells off");
    return TSFF_E_INVALID;
  }
  return TSFF_OK;
}
extern "C" int tsff_ff_pair_fwd(tsff_ctx* a, tsff_ctx* b, int64_t B, const double* params, const void* fe, int fe_dtype,
                                double* modl_a, double* modl_b, void* saved_a, void* saved_b, void* ws_a, void* stream) {
  int rc = check_pair(a, b);
  if (rc) return rc;
  if (B == 0) return TSFF_OK;
  if ((rc = check_common(a, B, params, fe, fe_dtype, saved_a, ws_a))) return rc;
  if ((rc = check_common(b, B, params, fe, fe_dtype, saved_b, ws_a))) return rc;
  if (!modl_a || !modl_b) { set_error("null output"); return TSFF_E_INVALID; }
  TSFF_ON_DEVICE(a);
  return table_pair_fwd(a, b, B, params, fe, fe_dtype, modl_a, modl_b, saved_a, saved_b, ws_a, static_cast<cudaStream_t>(stream));
}
extern "C" int tsff_ff_pair_bwd(tsff_ctx* a, tsff_ctx* b, int64_t B, const double* params, const void* fe, int fe_dtype,
                                const void* saved_a, const void* saved_b, const double* modl_bar_a, const double* modl_bar_b,
                                double* params_bar, void* fe_bar, void* ws_a, void* ws_b, void* stream) {
  int rc = check_pair(a, b);
  if (rc) return rc;
  if (B == 0) return TSFF_OK;
  if ((rc = check_common(a, B, params, fe, fe_dtype, saved_a, ws_a))) return rc;
  if ((rc = check_common(b, B, params, fe, fe_dtype, saved_b, ws_b))) return rc;
  if (!modl_bar_a || !modl_bar_b) { set_error("no cotangent given"); return TSFF_E_INVALID; }
  if (!fe_bar || !params_bar) { set_error("null output"); return TSFF_E_INVALID; }
  TSFF_ON_DEVICE(a);
  return table_pair_bwd(a, b, B, params, fe, fe_dtype, saved_a, saved_b, modl_bar_a, modl_bar_b, params_bar, fe_bar, ws_a, ws_b,
                        static_cast<cudaStream_t>(stream));
}

// ---- B1: stand-alone PV integral -----------------------------------------------------------------------------
namespace {
constexpr int kThreads = 256;

struct PvLayout {
  size_t D, tstat, D64, pend, desc, Dbar, pendbar, tI, tdI, bytes;
  int nodes, npad;
};
PvLayout pv_layout(int64_t B, int64_t N, int64_t P) {
  PvLayout L;
  L.nodes = (int)N - 1;
  L.npad = tree_npad(L.nodes);
  size_t o = 0;
  L.D = o; o += align_up((size_t)B * tree_blob(L.npad).bytes);
  L.tstat = o; o += align_up((size_t)kTreeStaticDoubles * 8);
  L.D64 = o; o += align_up((size_t)B * L.npad * 8);
  L.pend = o; o += align_up((size_t)B * 2 * 8);
  L.desc = o; o += align_up((size_t)B * P * 16);
  L.Dbar = o; o += align_up((size_t)B * L.npad * 8);
  L.pendbar = o; o += align_up((size_t)B * L.npad * 8);
  L.tI = o; o += align_up((size_t)B * P * 8);
  L.tdI = o; o += align_up((size_t)B * P * 8);
  L.bytes = o;
  return L;
}

__global__ void __launch_bounds__(kThreads) k_pv_prep(const double* f, int N, double h, int npad, unsigned char* blob,
                                                      const double* tstat, double* D64, double* pend) {
  const long long b = blockIdx.x;
  const double* fb = f + b * N;
  const int M = N - 2;
  extern __shared__ __align__(16) unsigned char prep_smem[];
  tree_prep_cta(fb, M, npad, blob + b * tree_blob(npad).bytes, tstat, reinterpret_cast<double*>(prep_smem));
  for (int i = threadIdx.x; i < npad; i += kThreads) D64[b * npad + i] = pv_weight(fb, M, h, i);  // FP64 validation path
  if (threadIdx.x == 0) {
    pend[2 * b] = fb[0];
    pend[2 * b + 1] = fb[M];
  }
}

__global__ void __launch_bounds__(kThreads) k_pv_desc(const double* pole, const double* out_bar, int P, double z0, double h,
                                                      int nodes, int npad, float4* desc, double* pnear) {
  const long long b = blockIdx.x;
  for (int p = threadIdx.x; p < P; p += kThreads) {
    const double xi = pole[b * P + p], ob = out_bar[b * P + p];
    int np, wb0;
    desc[b * P + p] = pv_desc(xi, ob, z0, h, nodes, npad, np, wb0);
    pv_bwd_pole_exact(xi, ob, z0, h, nodes, np, wb0, pnear + b * npad);
  }
}

__global__ void __launch_bounds__(kThreads) k_pv_bwd_finish(const double* pfar, const double* pnear, int N, int npad,
                                                            double* f_bar) {
  const long long b = blockIdx.x;
  for (int i = threadIdx.x; i < N; i += kThreads) f_bar[b * N + i] = (i < npad) ? pfar[b * npad + i] + pnear[b * npad + i] : 0.0;
}

__global__ void k_mul(const double* x, const double* y, long long n, double* out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = x[t] * y[t];
}

int launch_poles(const PvLayout& L, int64_t B, int64_t N, int64_t P, const double* f, char* w, double z0, double h, const double* pole, double* out,
                 double* dout, int prec, cudaStream_t st) {
  PvPolesArgs p;
  p.blob = (unsigned char*)(w + L.D); p.D64 = (double*)(w + L.D64); p.pend = (double*)(w + L.pend);
  p.pnodes = f; p.pnode_stride = N;
  p.poles = pole; p.pole_bstride = P; p.z0 = z0; p.h = h; p.nodes = L.nodes; p.npad = L.npad; p.P = (int)P;
  p.outI = out; p.outdI = dout;
  p.ntiles = (int)((P + kPvThreads - 1) / kPvThreads);
  const size_t smem = (size_t)tree_blob(L.npad).bytes;
  if (prec == TSFF_PV_FP64) {
    k_pv_poles<1, TSFF_PV_FP64><<<(unsigned)(B * p.ntiles), kPvThreads, 0, st>>>(p);
  } else {
    TSFF_SMEM_OPTIN((k_pv_poles<1, TSFF_PV_FP32>));
    k_pv_poles<1, TSFF_PV_FP32><<<(unsigned)(B * p.ntiles), kPvThreads, smem, st>>>(p);
  }
  TSFF_LAUNCH_OK("k_pv_poles");
  return TSFF_OK;
}
}  // namespace

extern "C" size_t tsff_pv_workspace_bytes(int64_t B, int64_t N, int64_t P) {
  if (B < 1 || N < 4 || P < 1) return 0;
  return pv_layout(B, N, P).bytes;
}

extern "C" int tsff_pv_fwd(int64_t B, int64_t N, int64_t P, const double* f, double z0, double h, const double* pole,
                           double* out, double* dout_dpole, int pv_precision, void* ws, void* stream) {
  if (B == 0 || P == 0) return TSFF_OK;
  if (!f || !pole || !out || !ws || B < 1 || N < 4 || P < 1) { set_error("bad argument"); return TSFF_E_INVALID; }
  if (tree_npad((int)N - 1) > kTreeMaxNpad) { set_error("N too large (limit %d nodes)", kTreeMaxNpad); return TSFF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PvLayout L = pv_layout(B, N, P);
  char* w = static_cast<char*>(ws);
  k_tree_static<<<kTreeStaticGrid, 256, 0, st>>>((int)N - 2, (double*)(w + L.tstat));
  k_pv_prep<<<(unsigned)B, kThreads, tree_prep_scratch_bytes(L.npad), st>>>(f, (int)N, h, L.npad, (unsigned char*)(w + L.D), (double*)(w + L.tstat),
                                            (double*)(w + L.D64), (double*)(w + L.pend));
  TSFF_LAUNCH_OK("k_pv_prep");
  return launch_poles(L, B, N, P, f, w, z0, h, pole, out, dout_dpole, pv_precision, st);
}

extern "C" int tsff_pv_bwd(int64_t B, int64_t N, int64_t P, const double* f, double z0, double h, const double* pole,
                           const double* out_bar, double* f_bar, double* pole_bar, void* ws, void* stream) {
  if (B == 0) return TSFF_OK;
  if (!f || !pole || !out_bar || !f_bar || !ws || B < 1 || N < 4 || P < 1) { set_error("bad argument"); return TSFF_E_INVALID; }
  if (tree_npad((int)N - 1) > kTreeMaxNpad) { set_error("N too large (limit %d nodes)", kTreeMaxNpad); return TSFF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PvLayout L = pv_layout(B, N, P);
  char* w = static_cast<char*>(ws);
  TSFF_CUDA_OK(cudaMemsetAsync(w + L.pendbar, 0, (size_t)B * L.npad * 8, st));
  k_pv_desc<<<(unsigned)B, kThreads, 0, st>>>(pole, out_bar, (int)P, z0, h, L.nodes, L.npad, (float4*)(w + L.desc), (double*)(w + L.pendbar));
  TSFF_LAUNCH_OK("k_pv_desc");
  k_tree_static<<<kTreeStaticGrid, 256, 0, st>>>((int)N - 2, (double*)(w + L.tstat));
  PvNodesArgs n;
  n.desc = (float4*)(w + L.desc); n.tstat = (double*)(w + L.tstat); n.P = (int)P; n.nodes = L.nodes; n.npad = L.npad;
  n.pbar = (double*)(w + L.Dbar); n.nsplit = 1;
  TSFF_SMEM_OPTIN(k_pv_nodes);
  k_pv_nodes<<<(unsigned)B, kPvThreads, pv_nodes_smem(n.npad), st>>>(n);
  TSFF_LAUNCH_OK("k_pv_nodes");
  k_pv_bwd_finish<<<(unsigned)B, kThreads, 0, st>>>((double*)(w + L.Dbar), (double*)(w + L.pendbar), (int)N, L.npad, f_bar);
  TSFF_LAUNCH_OK("k_pv_bwd_finish");
  if (pole_bar) {
    k_pv_prep<<<(unsigned)B, kThreads, tree_prep_scratch_bytes(L.npad), st>>>(f, (int)N, h, L.npad, (unsigned char*)(w + L.D), (double*)(w + L.tstat),
                                              (double*)(w + L.D64), (double*)(w + L.pend));
    TSFF_LAUNCH_OK("k_pv_prep");
    int rc = launch_poles(L, B, N, P, f, w, z0, h, pole, (double*)(w + L.tI), (double*)(w + L.tdI), TSFF_PV_FP32, st);
    if (rc) return rc;
    const long long n2 = (long long)B * P;
    k_mul<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(out_bar, (double*)(w + L.tdI), n2, pole_bar);
    TSFF_LAUNCH_OK("k_mul");
  }
  return TSFF_OK;
}
