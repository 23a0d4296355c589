"""jax.ffi + jax.custom_vjp wrappers over libtsff: the module a TSADAR maintainer drops into
`tsadar/core/physics/` so that `FormFactor`, `ThomsonScatteringDiagnostic`, `LossFunction`, the yaml decks and the
optax / SciPy optimisers stay untouched.

JAX is not installable in the image this repository is developed in, so this module is NOT imported by the package or by
any test that runs here; importing it without JAX raises a clear ImportError.  It is a thin translation of the C ABI
(include/tsff.h): contexts are created through ctypes (tsadar_b200/_ffi.py, same struct as the torch harness), the
handlers of tsff_xla_ffi.cc are registered as FFI targets, and each forward/backward pair is tied by jax.custom_vjp."""
from __future__ import annotations

import ctypes
import os

try:
    import jax
    import jax.numpy as jnp
    import numpy as np
except ImportError as e:  # pragma: no cover - JAX is absent where this repository's tests run
    raise ImportError("tsadar_b200.jax_binding.tsff_jax needs jax/jaxlib (and libtsff_xla.so built from "
                      "tsff_xla_ffi.cc against jax.ffi.include_dir())") from e

# Batching rule of every custom call: "broadcast_all" -- under jax.vmap (the reference vmaps its model over lineouts,
# thomson_diagnostic.py:35-36) XLA passes operands with the mapped axis as an extra leading dimension (unmapped operands
# broadcast to it) and the handlers of tsff_xla_ffi.cc fold all leading dimensions into the kernels' own batch axis B:
# ONE launch for the whole vmapped batch.  ("sequential" would issue one FFI call per lineout and defeat the batched
# kernels; "expand_dims" leaves unmapped operands with a leading 1, which the flat [B, ...] kernels cannot index.)
_VMAP = "broadcast_all"

_HERE = os.path.dirname(os.path.abspath(__file__))
_XLA_LIB = ctypes.CDLL(os.environ.get("TSFF_XLA_LIB", os.path.join(_HERE, "..", "_lib", "libtsff_xla.so")))
for _name in ("TsffFfFwd", "TsffFfFullFwd", "TsffFfBwd", "TsffFfFullBwd", "TsffFfPairFwd", "TsffFfPairBwd", "TsffPvFwd", "TsffChi2vFwd",
              "TsffLossFwdBwd"):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_XLA_LIB, _name)), platform="CUDA")


def make_form_factor(ctx: int, B: int, W: int, V: int, NP: int, saved_bytes: int, ws_bytes: int):
    """-> f(params [B, NP] f64, fe [B, V]) = modl [B, W] f64, differentiable in both arguments.
    `ctx` = the integer value of a tsff_ctx* made with tsff_ctx_create (ctypes); saved_bytes / ws_bytes from
    tsff_ff_saved_bytes / tsff_ff_workspace_bytes for this batch size.  The lineout axis is the kernel's own batch axis, so
    the reference's vmap over lineouts (thomson_diagnostic.py:35) can be dropped -- or kept: with B = 1 per call the
    vmapped call arrives as ONE handler invocation with params [N, 1, NP] and runs as a batch of N (see _VMAP)."""
    cid = np.int64(ctx)

    @jax.custom_vjp
    def ff(params, fe):
        return _fwd(params, fe)[0]

    def _fwd(params, fe):
        modl, saved, _ = jax.ffi.ffi_call(
            "TsffFfFwd",
            (jax.ShapeDtypeStruct((B, W), jnp.float64), jax.ShapeDtypeStruct((saved_bytes,), jnp.uint8),
             jax.ShapeDtypeStruct((ws_bytes,), jnp.uint8)), vmap_method=_VMAP)(params, fe, ctx=cid)
        return modl, (params, fe, saved)

    def _bwd(res, modl_bar):
        params, fe, saved = res
        pbar, fbar, _ = jax.ffi.ffi_call(
            "TsffFfBwd",
            (jax.ShapeDtypeStruct((B, NP), jnp.float64), jax.ShapeDtypeStruct((B, V), fe.dtype),
             jax.ShapeDtypeStruct((ws_bytes,), jnp.uint8)), vmap_method=_VMAP)(params, fe, saved, modl_bar, ctx=cid)
        return pbar, fbar

    ff.defvjp(_fwd, _bwd)
    return ff


def make_form_factor_pair(ctx_a: int, ctx_b: int, B: int, W_a: int, W_b: int, V: int, NP: int, saved_a: int, saved_b: int,
                          ws_a: int, ws_b: int):
    """-> f(params [B, NP], fe [B, V]) = (modl_a [B, W_a], modl_b [B, W_b]): the electron and ion windows of FitModel.__call__
    (generate_spectra.py:332-336) in one custom call each way; the f-dependent tables are built once (tsff_ff_pair_fwd / _bwd).
    ws_a must be max(workspace bytes of a, of b)."""
    ca, cb = np.int64(ctx_a), np.int64(ctx_b)

    @jax.custom_vjp
    def ff(params, fe):
        return _fwd(params, fe)[0]

    def _fwd(params, fe):
        ma, mb, sa, sb, _ = jax.ffi.ffi_call(
            "TsffFfPairFwd",
            (jax.ShapeDtypeStruct((B, W_a), jnp.float64), jax.ShapeDtypeStruct((B, W_b), jnp.float64),
             jax.ShapeDtypeStruct((saved_a,), jnp.uint8), jax.ShapeDtypeStruct((saved_b,), jnp.uint8),
             jax.ShapeDtypeStruct((ws_a,), jnp.uint8)), vmap_method=_VMAP)(params, fe, ctx=ca, ctx_b=cb)
        return (ma, mb), (params, fe, sa, sb)

    def _bwd(res, bars):
        params, fe, sa, sb = res
        pbar, fbar, _, _ = jax.ffi.ffi_call(
            "TsffFfPairBwd",
            (jax.ShapeDtypeStruct((B, NP), jnp.float64), jax.ShapeDtypeStruct((B, V), fe.dtype),
             jax.ShapeDtypeStruct((ws_a,), jnp.uint8), jax.ShapeDtypeStruct((ws_b,), jnp.uint8)),
            vmap_method=_VMAP)(params, fe, sa, sb, bars[0], bars[1], ctx=ca, ctx_b=cb)
        return pbar, fbar

    ff.defvjp(_fwd, _bwd)
    return ff


def make_form_factor_full(ctx: int, B: int, G: int, W: int, A: int, fe_shape, NP: int, saved_bytes: int, ws_bytes: int):
    """The same returning the full formfactor [B, G, W, A] (ARTS; TSFF_MODE_2V takes fe [B, V, V] float64)."""
    cid = np.int64(ctx)

    @jax.custom_vjp
    def ff(params, fe):
        return _fwd(params, fe)[0]

    def _fwd(params, fe):
        out, saved, _ = jax.ffi.ffi_call(
            "TsffFfFullFwd",
            (jax.ShapeDtypeStruct((B, G, W, A), jnp.float64), jax.ShapeDtypeStruct((saved_bytes,), jnp.uint8),
             jax.ShapeDtypeStruct((ws_bytes,), jnp.uint8)), vmap_method=_VMAP)(params, fe, ctx=cid)
        return out, (params, fe, saved)

    def _bwd(res, ff_bar):
        params, fe, saved = res
        pbar, fbar, _ = jax.ffi.ffi_call(
            "TsffFfFullBwd",
            (jax.ShapeDtypeStruct((B, NP), jnp.float64), jax.ShapeDtypeStruct(tuple(fe_shape), fe.dtype),
             jax.ShapeDtypeStruct((ws_bytes,), jnp.uint8)), vmap_method=_VMAP)(params, fe, saved, ff_bar, ctx=cid)
        return pbar, fbar

    ff.defvjp(_fwd, _bwd)
    return ff


def pack_params(params):
    """ThomsonParams.__call__ dict (ts_params.py:599-603) -> block [B, NP] in the column order of include/tsff.h."""
    ions = sorted([k for k in params if k.startswith("ion-")], key=lambda s: int(s.split("-")[1]))
    e, g = params["electron"], params["general"]
    cols = [e["Te"], e["ne"], g["lam"], g["Va"], g["ud"], g["ne_gradient"], g["Te_gradient"], g["amp1"], g["amp2"], g["amp3"]]
    for k in ions:
        cols += [params[k]["A"], params[k]["Z"], params[k]["Ti"], params[k]["fract"]]
    cols = [jnp.atleast_1d(jnp.asarray(c, dtype=jnp.float64)).reshape(-1) for c in cols]
    B = max(c.shape[0] for c in cols)
    return jnp.stack([jnp.broadcast_to(c, (B,)) for c in cols], axis=1)


def calc_all_chi_vals(ctx: int, DF, beta, xie_mag, klde_mag):
    """FormFactor.calc_all_chi_vals (form_factor.py:390-447) on a TSFF_MODE_2V context -> (fe_vphi, chiEI, chiERrat), each
    shaped like beta.  Forward only: inside calc_in_2D the whole stage is one custom call with its own VJP."""
    P = int(np.prod(beta.shape))
    chi = jax.ffi.ffi_call("TsffChi2vFwd", jax.ShapeDtypeStruct((3, P), jnp.float64), vmap_method=_VMAP)(
        DF, beta.reshape(-1), xie_mag.reshape(-1), klde_mag.reshape(-1), ctx=np.int64(ctx))
    return chi[0].reshape(beta.shape), chi[1].reshape(beta.shape), chi[2].reshape(beta.shape)
