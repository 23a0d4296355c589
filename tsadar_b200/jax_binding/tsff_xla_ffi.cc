// tsff_xla_ffi.cc -- XLA FFI handlers over the C ABI of libtsff (include/tsff.h): the reference-side binding of
// INTEGRATION.md section 3.  NOT built by tsadar_b200/build.py: it needs the XLA FFI headers that ship inside the
// jaxlib wheel (`python -c "import jax; print(jax.ffi.include_dir())"`), which this image does not have.  Build on a
// machine with JAX:
//
//   nvcc -std=c++17 -shared -Xcompiler -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude \
//        tsadar_b200/jax_binding/tsff_xla_ffi.cc -Ltsadar_b200/_lib -ltsff -o tsadar_b200/_lib/libtsff_xla.so
//
// A tsff_ctx is created from Python (ctypes, tsadar_b200/_ffi.py) once per (lambda range, npts, angles, f grid) and handed
// to the handlers as an int64 attribute; it is immutable, so XLA may call from several host threads / streams.
#include <cuda_runtime_api.h>

#include <cstdint>

#include "tsff.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

inline ffi::Error status(int rc) { return rc ? ffi::Error::Internal(tsff_last_error()) : ffi::Error::Success(); }
inline int fe_dtype(const ffi::AnyBuffer& fe) { return fe.element_type() == ffi::F32 ? TSFF_F32 : TSFF_F64; }
// Lineouts in the call.  The Python side registers every call with vmap_method="broadcast_all": under the reference's
// vmap over lineouts (thomson_diagnostic.py:35-36) XLA hands the handler operands with extra LEADING batch dimensions
// (params [N, B, NP], fe [N, B, V], ...).  All leading dimensions are folded into the kernel's own batch axis -- one launch
// for the whole vmapped batch, never one call per lineout.  saved / ws arrive as [N, bytes(B)] >= bytes(N * B) (every
// segment of the layout is 256-byte aligned, so N * align(x) >= align(N * x)) and are used as one flat buffer.
inline int64_t lineouts(const ffi::Buffer<ffi::F64>& params) {
  const auto d = params.dimensions();
  return static_cast<int64_t>(params.element_count()) / d[d.size() - 1];
}

// FormFactor.__call__ + FitModel angle sum (form_factor.py:163-298, generate_spectra.py:164-165,193-197):
// params [B, NP] f64, fe [B, V] f32|f64 -> modl [B, W] f64; saved / ws are result buffers sized by tsff_ff_*_bytes.
ffi::Error FfFwd(cudaStream_t stream, int64_t ctx, ffi::Buffer<ffi::F64> params, ffi::AnyBuffer fe,
                 ffi::ResultBuffer<ffi::F64> modl, ffi::ResultBuffer<ffi::U8> saved, ffi::ResultBuffer<ffi::U8> ws) {
  const int64_t B = lineouts(params);
  return status(tsff_ff_fwd(reinterpret_cast<tsff_ctx*>(ctx), B, params.typed_data(), fe.untyped_data(), fe_dtype(fe),
                            modl->typed_data(), nullptr, saved->typed_data(), ws->typed_data(), stream));
}

// the same returning the full formfactor [B, G, W, A] (ARTS: the weight matrix product follows in JAX)
ffi::Error FfFullFwd(cudaStream_t stream, int64_t ctx, ffi::Buffer<ffi::F64> params, ffi::AnyBuffer fe,
                     ffi::ResultBuffer<ffi::F64> ff, ffi::ResultBuffer<ffi::U8> saved, ffi::ResultBuffer<ffi::U8> ws) {
  const int64_t B = lineouts(params);
  return status(tsff_ff_fwd(reinterpret_cast<tsff_ctx*>(ctx), B, params.typed_data(), fe.untyped_data(), fe_dtype(fe),
                            nullptr, ff->typed_data(), saved->typed_data(), ws->typed_data(), stream));
}

// VJP of FfFwd (replaces XLA's reverse mode of the same graph, loss_function.py:107-108)
ffi::Error FfBwd(cudaStream_t stream, int64_t ctx, ffi::Buffer<ffi::F64> params, ffi::AnyBuffer fe, ffi::Buffer<ffi::U8> saved,
                 ffi::Buffer<ffi::F64> modl_bar, ffi::ResultBuffer<ffi::F64> params_bar, ffi::Result<ffi::AnyBuffer> fe_bar,
                 ffi::ResultBuffer<ffi::U8> ws) {
  const int64_t B = lineouts(params);
  return status(tsff_ff_bwd(reinterpret_cast<tsff_ctx*>(ctx), B, params.typed_data(), fe.untyped_data(), fe_dtype(fe),
                            saved.typed_data(), modl_bar.typed_data(), nullptr, params_bar->typed_data(),
                            fe_bar->untyped_data(), ws->typed_data(), stream));
}

ffi::Error FfFullBwd(cudaStream_t stream, int64_t ctx, ffi::Buffer<ffi::F64> params, ffi::AnyBuffer fe,
                     ffi::Buffer<ffi::U8> saved, ffi::Buffer<ffi::F64> ff_bar, ffi::ResultBuffer<ffi::F64> params_bar,
                     ffi::Result<ffi::AnyBuffer> fe_bar, ffi::ResultBuffer<ffi::U8> ws) {
  const int64_t B = lineouts(params);
  return status(tsff_ff_bwd(reinterpret_cast<tsff_ctx*>(ctx), B, params.typed_data(), fe.untyped_data(), fe_dtype(fe),
                            saved.typed_data(), nullptr, ff_bar.typed_data(), params_bar->typed_data(),
                            fe_bar->untyped_data(), ws->typed_data(), stream));
}

// Two windows of one plasma (the electron and ion FormFactor instances of FitModel, generate_spectra.py:136-165): ctx = window a
// (its f-dependent tables serve both), ctx_b = window b.  One call forward, one call for the VJP of both outputs.
ffi::Error FfPairFwd(cudaStream_t stream, int64_t ctx, int64_t ctx_b, ffi::Buffer<ffi::F64> params, ffi::AnyBuffer fe,
                     ffi::ResultBuffer<ffi::F64> modl_a, ffi::ResultBuffer<ffi::F64> modl_b, ffi::ResultBuffer<ffi::U8> saved_a,
                     ffi::ResultBuffer<ffi::U8> saved_b, ffi::ResultBuffer<ffi::U8> ws) {
  const int64_t B = lineouts(params);
  return status(tsff_ff_pair_fwd(reinterpret_cast<tsff_ctx*>(ctx), reinterpret_cast<tsff_ctx*>(ctx_b), B, params.typed_data(),
                                 fe.untyped_data(), fe_dtype(fe), modl_a->typed_data(), modl_b->typed_data(), saved_a->typed_data(),
                                 saved_b->typed_data(), ws->typed_data(), stream));
}

ffi::Error FfPairBwd(cudaStream_t stream, int64_t ctx, int64_t ctx_b, ffi::Buffer<ffi::F64> params, ffi::AnyBuffer fe,
                     ffi::Buffer<ffi::U8> saved_a, ffi::Buffer<ffi::U8> saved_b, ffi::Buffer<ffi::F64> bar_a,
                     ffi::Buffer<ffi::F64> bar_b, ffi::ResultBuffer<ffi::F64> params_bar, ffi::Result<ffi::AnyBuffer> fe_bar,
                     ffi::ResultBuffer<ffi::U8> ws_a, ffi::ResultBuffer<ffi::U8> ws_b) {
  const int64_t B = lineouts(params);
  return status(tsff_ff_pair_bwd(reinterpret_cast<tsff_ctx*>(ctx), reinterpret_cast<tsff_ctx*>(ctx_b), B, params.typed_data(),
                                 fe.untyped_data(), fe_dtype(fe), saved_a.typed_data(), saved_b.typed_data(), bar_a.typed_data(),
                                 bar_b.typed_data(), params_bar->typed_data(), fe_bar->untyped_data(), ws_a->typed_data(),
                                 ws_b->typed_data(), stream));
}

// vmap(ratintn) on uniform nodes (ratintn.py:4-23): f [B, N], pole [B, P] -> out [B, P], dout_dpole [B, P]
ffi::Error PvFwd(cudaStream_t stream, double z0, double h, ffi::Buffer<ffi::F64> f, ffi::Buffer<ffi::F64> pole,
                 ffi::ResultBuffer<ffi::F64> out, ffi::ResultBuffer<ffi::F64> dout, ffi::ResultBuffer<ffi::U8> ws) {
  const auto fd = f.dimensions(), pd = pole.dimensions();
  const int64_t N = fd[fd.size() - 1], P = pd[pd.size() - 1], B = static_cast<int64_t>(f.element_count()) / N;   // leading dims folded
  return status(tsff_pv_fwd(B, N, P, f.typed_data(), z0, h, pole.typed_data(), out->typed_data(), dout->typed_data(),
                            TSFF_PV_FP32, ws->typed_data(), stream));
}

// FormFactor.calc_all_chi_vals (form_factor.py:390-447): DF [V, V]; beta, xie_mag, klde_mag flattened [P] -> chi [3, P]
ffi::Error Chi2vFwd(cudaStream_t stream, int64_t ctx, ffi::Buffer<ffi::F64> fe, ffi::Buffer<ffi::F64> beta,
                    ffi::Buffer<ffi::F64> xie_mag, ffi::Buffer<ffi::F64> klde_mag, ffi::ResultBuffer<ffi::F64> chi) {
  const int64_t P = static_cast<int64_t>(beta.element_count());
  return status(tsff_chi2v_fwd(reinterpret_cast<tsff_ctx*>(ctx), fe.typed_data(), beta.typed_data(), xie_mag.typed_data(),
                               klde_mag.typed_data(), P, chi->typed_data(), stream));
}

// masked loss + seed cotangent (loss_function.py:190-267, 386-418)
ffi::Error LossFwdBwd(cudaStream_t stream, double uncert, double scale, int64_t method, ffi::Buffer<ffi::F64> theory,
                      ffi::Buffer<ffi::F64> data, ffi::Buffer<ffi::F64> weight, ffi::ResultBuffer<ffi::F64> loss,
                      ffi::ResultBuffer<ffi::F64> theory_bar) {
  const auto td = theory.dimensions();
  const int32_t n = static_cast<int32_t>(td[td.size() - 1]);
  const int64_t B = static_cast<int64_t>(theory.element_count()) / n;
  cudaError_t e = cudaMemsetAsync(loss->typed_data(), 0, sizeof(double), stream);   // the entry point accumulates
  if (e != cudaSuccess) return ffi::Error::Internal(cudaGetErrorString(e));
  return status(tsff_loss_fwd_bwd(B, n, theory.typed_data(), data.typed_data(), weight.typed_data(), uncert, scale,
                                  static_cast<int>(method), loss->typed_data(), theory_bar->typed_data(), stream));
}

}  // namespace

#define TSFF_FF_BINDING()                                                                         \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("ctx").Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::AnyBuffer>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(TsffFfFwd, FfFwd,
                              TSFF_FF_BINDING().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TsffFfFullFwd, FfFullFwd,
                              TSFF_FF_BINDING().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TsffFfBwd, FfBwd,
                              TSFF_FF_BINDING().Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::AnyBuffer>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TsffFfFullBwd, FfFullBwd,
                              TSFF_FF_BINDING().Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::AnyBuffer>().Ret<ffi::Buffer<ffi::U8>>());
#define TSFF_PAIR_BINDING()                                                                                              \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("ctx").Attr<int64_t>("ctx_b")                       \
      .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::AnyBuffer>()
XLA_FFI_DEFINE_HANDLER_SYMBOL(TsffFfPairFwd, FfPairFwd,
                              TSFF_PAIR_BINDING().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>().Ret<ffi::Buffer<ffi::U8>>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TsffFfPairBwd, FfPairBwd,
                              TSFF_PAIR_BINDING().Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::AnyBuffer>().Ret<ffi::Buffer<ffi::U8>>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TsffPvFwd, PvFwd,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<double>("z0").Attr<double>("h")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TsffChi2vFwd, Chi2vFwd,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("ctx")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TsffLossFwdBwd, LossFwdBwd,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<double>("uncert").Attr<double>("scale")
                                  .Attr<int64_t>("method").Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());
