"""Thin torch <-> libtsff plumbing: context lifetime, workspace tensors, and the autograd Function that plays
the role `jax.custom_vjp` plays in the JAX binding.  torch is used for device memory and
streams only; all arithmetic happens in the CUDA kernels behind the C ABI."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _ffi

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _zprime_table():
    t = np.load(os.path.join(_DATA, "zprime_table.npz"))
    return (np.ascontiguousarray(t["x"], dtype=np.float64), np.ascontiguousarray(t["re"], dtype=np.float64),
            np.ascontiguousarray(t["im"], dtype=np.float64))


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _require_cuda(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{name} must be a CUDA tensor: tsadar_b200 has no CPU path")
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")


class FormFactorEngine:
    """One immutable libtsff context (= the static state of the reference's FormFactor.__init__,
    form_factor.py:120-161) bound to one GPU."""

    def __init__(self, lambda_range, npts, lam_shift, sa_deg, weights, num_grad_points, n_ions, vx, mode="table",
                 jmul=None, pv_precision="fp32", device=None, ud_ang=0.0, va_ang=0.0, w_slice=None):
        """w_slice = (j0, j1): this engine covers the wavelength points [j0, j1) of linspace(lambda_range, npts) (the
        W-axis shard of one rank, tsadar_b200/parallel.py); jmul is then the local slice."""
        if not torch.cuda.is_available():
            raise RuntimeError("tsadar_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        vx = np.asarray(vx, dtype=np.float64)
        dv = float(vx[1] - vx[0])
        if not np.allclose(np.diff(vx), dv, rtol=1e-9, atol=0):
            raise ValueError("the f-table grid must be uniform (DistributionFunction1V, base.py:149-151)")
        sa_deg = np.ascontiguousarray(np.asarray(sa_deg, dtype=np.float64).reshape(-1))
        weights = np.asarray(weights, dtype=np.float64)
        if weights.ndim == 0:
            weights = np.full(sa_deg.shape, float(weights))
        weights = np.ascontiguousarray(weights.reshape(-1))
        if weights.shape != sa_deg.shape:
            raise ValueError("weights must have one entry per scattering angle")
        j0, j1 = (0, int(npts)) if w_slice is None else (int(w_slice[0]), int(w_slice[1]))
        if not (0 <= j0 < j1 <= int(npts)):
            raise ValueError(f"w_slice {w_slice} outside [0, {npts})")
        self.W, self.A, self.G, self.I, self.V = j1 - j0, int(sa_deg.size), int(num_grad_points), int(n_ions), int(vx.size)
        self.NP = _ffi.P_ION0 + _ffi.ION_STRIDE * self.I
        self.mode = mode
        zx, zr, zi = _zprime_table()
        cfg = _ffi.StaticCfg()
        cfg.abi_version = _ffi.TSFF_ABI_VERSION
        cfg.mode = {"table": _ffi.TSFF_MODE_TABLE, "direct": _ffi.TSFF_MODE_DIRECT, "2v": _ffi.TSFF_MODE_2V}[mode]
        cfg.ud_angle_deg, cfg.va_angle_deg = float(ud_ang or 0.0), float(va_ang or 0.0)
        cfg.W, cfg.A, cfg.G, cfg.I, cfg.V = self.W, self.A, self.G, self.I, self.V
        cfg.W_total, cfg.w_offset = int(npts), j0
        cfg.pv_precision = _ffi.TSFF_PV_FP64 if pv_precision == "fp64" else _ffi.TSFF_PV_FP32
        cfg.lam_min, cfg.lam_max, cfg.lam_shift = float(lambda_range[0]), float(lambda_range[1]), float(lam_shift)
        cfg.v0, cfg.dv = float(vx[0]), dv
        cfg.sa_deg, cfg.weights = _dptr(sa_deg), _dptr(weights)
        if jmul is not None:
            jmul = np.ascontiguousarray(np.asarray(jmul, dtype=np.float64).reshape(-1))
            assert jmul.size == self.W
            cfg.jmul = _dptr(jmul)
        cfg.zp_x, cfg.zp_re, cfg.zp_im, cfg.zp_n = _dptr(zx), _dptr(zr), _dptr(zi), int(zx.size)
        self._ctx = C.c_void_p()
        _ffi.check(_ffi.lib().tsff_ctx_create(self.device.index, C.byref(cfg), C.byref(self._ctx)))
        # wavelength axis in nm as FitModel returns it (lams * 1e7, generate_spectra.py:163,191)
        lam = np.linspace(cfg.lam_min, cfg.lam_max, int(npts))[j0:j1]
        omgs = 2e7 * np.pi * 2.99792458e10 / lam
        self.lam_cm = 2 * np.pi * 2.99792458e10 / omgs
        self._buf = {}
        self.launches = 0

    def __del__(self):
        try:
            if getattr(self, "_ctx", None) and self._ctx.value:
                _ffi.lib().tsff_ctx_destroy(self._ctx)
                self._ctx = C.c_void_p()
        except Exception:
            pass

    # ---- buffers ------------------------------------------------------------------------------------------
    def _scratch(self, key, nbytes):
        t = self._buf.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._buf[key] = t
        return t

    def saved_bytes(self, B):
        return int(_ffi.lib().tsff_ff_saved_bytes(self._ctx, B))

    def workspace_bytes(self, B):
        return int(_ffi.lib().tsff_ff_workspace_bytes(self._ctx, B))

    # ---- raw calls ----------------------------------------------------------------------------------------
    def forward(self, params, fe, want_ff=False, want_modl=True, saved=None):
        """params [B,NP] f64 cuda, fe [B,V] f32|f64 cuda -> (modl [B,W] | None, ff [B,G,W,A] | None, saved)."""
        _require_cuda(params, torch.float64, "params")
        if fe.dtype not in (torch.float32, torch.float64):
            raise RuntimeError("fe must be float32 or float64")
        _require_cuda(fe, fe.dtype, "fe")
        B = params.shape[0]
        if self.mode == "2v":
            assert params.shape == (B, self.NP) and fe.shape == (B, self.V, self.V) and fe.dtype == torch.float64, (params.shape, fe.shape)
            want_ff, want_modl = True, False
        else:
            assert params.shape == (B, self.NP) and fe.shape == (B, self.V), (params.shape, fe.shape)
        modl = torch.empty((B, self.W), dtype=torch.float64, device=self.device) if want_modl else None
        ff = torch.empty((B, self.G, self.W, self.A), dtype=torch.float64, device=self.device) if want_ff else None
        if saved is None:
            saved = torch.empty(self.saved_bytes(B), dtype=torch.uint8, device=self.device)
        ws = self._scratch("ws", self.workspace_bytes(B))
        st = torch.cuda.current_stream(self.device).cuda_stream
        _ffi.check(_ffi.lib().tsff_ff_fwd(
            self._ctx, B, params.data_ptr(), fe.data_ptr(), _ffi.TSFF_F32 if fe.dtype == torch.float32 else _ffi.TSFF_F64,
            modl.data_ptr() if modl is not None else None, ff.data_ptr() if ff is not None else None,
            saved.data_ptr(), ws.data_ptr(), st))
        return modl, ff, saved

    def chi_vals_2v(self, fe, beta, xie_mag, klde_mag):
        """2V mode: FormFactor.calc_all_chi_vals -- fe [V,V], beta / xie_mag / klde_mag of any common shape (float64 cuda)
        -> (fe_vphi, chiEI, chiERrat) of that shape."""
        assert self.mode == "2v", "chi_vals_2v needs a 2V engine"
        for name, t in (("fe", fe), ("beta", beta), ("xie_mag", xie_mag), ("klde_mag", klde_mag)):
            _require_cuda(t, torch.float64, name)
        assert fe.shape == (self.V, self.V), fe.shape
        shape = beta.shape
        P = beta.numel()
        assert xie_mag.numel() == P and klde_mag.numel() == P
        out = torch.empty((3, P), dtype=torch.float64, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _ffi.check(_ffi.lib().tsff_chi2v_fwd(self._ctx, fe.contiguous().data_ptr(), beta.contiguous().data_ptr(),
                                             xie_mag.contiguous().data_ptr(), klde_mag.contiguous().data_ptr(), P, out.data_ptr(), st))
        return out[0].reshape(shape), out[1].reshape(shape), out[2].reshape(shape)

    def backward(self, params, fe, saved, modl_bar=None, ff_bar=None, params_bar=None, fe_bar=None, want_params=True):
        """want_params=False (2V mode only): skip params_bar (returned as None) when no kinematic parameter is trainable."""
        B = params.shape[0]
        if modl_bar is not None:
            _require_cuda(modl_bar, torch.float64, "modl_bar")
        if ff_bar is not None:
            _require_cuda(ff_bar, torch.float64, "ff_bar")
        skip_params = (not want_params) and self.mode == "2v"
        if params_bar is None and not skip_params:
            params_bar = torch.empty((B, self.NP), dtype=torch.float64, device=self.device)
        if fe_bar is None:
            fe_bar = torch.empty_like(fe)
        ws = self._scratch("ws", self.workspace_bytes(B))
        st = torch.cuda.current_stream(self.device).cuda_stream
        _ffi.check(_ffi.lib().tsff_ff_bwd(
            self._ctx, B, params.data_ptr(), fe.data_ptr(), _ffi.TSFF_F32 if fe.dtype == torch.float32 else _ffi.TSFF_F64,
            saved.data_ptr(), modl_bar.data_ptr() if modl_bar is not None else None,
            ff_bar.data_ptr() if ff_bar is not None else None, None if skip_params else params_bar.data_ptr(), fe_bar.data_ptr(),
            ws.data_ptr(), st))
        return params_bar, fe_bar

    def set_frozen_cells(self, mode, B=None):
        """Second-order path (table mode): mode "record" -- the next forward stores the cells of its linear interpolations;
        "replay" -- forward and backward extend the recorded cells linearly (what jax.hessian differentiates); "off"."""
        m = {"off": 0, "record": 1, "replay": 2}[mode]
        self._cell_mode = m
        if m == 0:
            _ffi.check(_ffi.lib().tsff_ctx_set_frozen_cells(self._ctx, 0, None, 0))
            return
        if m == 1:
            self._cells_B = int(B)
            self._cells = torch.zeros(int(_ffi.lib().tsff_ff_cells_bytes(self._ctx, self._cells_B)), dtype=torch.uint8, device=self.device)
        _ffi.check(_ffi.lib().tsff_ctx_set_frozen_cells(self._ctx, m, self._cells.data_ptr(), self._cells_B))

    def set_profile_events(self, fwd=None, bwd=None):
        """fwd/bwd: (start, stop) torch.cuda.Event pairs (enable_timing=True) recorded around the dominant kernels."""
        self._prof = (fwd, bwd)  # keep the events alive
        h = [None, None, None, None]
        for k, pair in enumerate((fwd, bwd)):
            if pair is not None:
                for e in pair:
                    e.record()  # torch creates the cudaEvent lazily on first record
                h[2 * k], h[2 * k + 1] = pair[0].cuda_event, pair[1].cuda_event
        _ffi.check(_ffi.lib().tsff_ctx_set_profile_events(self._ctx, h[0], h[1], h[2], h[3]))

    # number of kernel launches one forward / backward call makes (for bench.py's gpu_launches claim)
    def launches_fwd(self, want_modl=True):
        if self.mode == "table":
            return 3
        if self.mode == "2v":
            return 1
        return 3 + (1 if want_modl and not (self.A == 1 and self.G == 1) else 0)   # A = G = 1: the angle sum is fused

    def launches_bwd(self):
        return 4 if self.mode == "table" else 3  # (+1 memset node, not a kernel of ours)


class _FFPairFunction(torch.autograd.Function):
    """Two spectral windows of one plasma (the reference's electron and ion FormFactor instances on the same parameters and f):
    tsff_ff_pair_fwd / _bwd build the f-dependent tables once and run one principal-value adjoint sweep for both."""

    @staticmethod
    def forward(ctx, eng_a, eng_b, params, fe):
        _require_cuda(params, torch.float64, "params")
        _require_cuda(fe, fe.dtype, "fe")
        B = params.shape[0]
        assert params.shape == (B, eng_a.NP) and fe.shape == (B, eng_a.V) and eng_b.V == eng_a.V and eng_b.NP == eng_a.NP
        dev = eng_a.device
        modl_a = torch.empty((B, eng_a.W), dtype=torch.float64, device=dev)
        modl_b = torch.empty((B, eng_b.W), dtype=torch.float64, device=dev)
        saved_a = torch.empty(eng_a.saved_bytes(B), dtype=torch.uint8, device=dev)
        saved_b = torch.empty(eng_b.saved_bytes(B), dtype=torch.uint8, device=dev)
        ws_a = eng_a._scratch("ws", max(eng_a.workspace_bytes(B), eng_b.workspace_bytes(B)))
        st = torch.cuda.current_stream(dev).cuda_stream
        _ffi.check(_ffi.lib().tsff_ff_pair_fwd(
            eng_a._ctx, eng_b._ctx, B, params.data_ptr(), fe.data_ptr(), _ffi.TSFF_F32 if fe.dtype == torch.float32 else _ffi.TSFF_F64,
            modl_a.data_ptr(), modl_b.data_ptr(), saved_a.data_ptr(), saved_b.data_ptr(), ws_a.data_ptr(), st))
        ctx.eng_a, ctx.eng_b = eng_a, eng_b
        ctx.save_for_backward(params, fe, saved_a, saved_b)
        return modl_a, modl_b

    @staticmethod
    def backward(ctx, bar_a, bar_b):
        params, fe, saved_a, saved_b = ctx.saved_tensors
        eng_a, eng_b = ctx.eng_a, ctx.eng_b
        B = params.shape[0]
        bar_a = (torch.zeros((B, eng_a.W), dtype=torch.float64, device=params.device) if bar_a is None else bar_a).contiguous()
        bar_b = (torch.zeros((B, eng_b.W), dtype=torch.float64, device=params.device) if bar_b is None else bar_b).contiguous()
        params_bar = torch.empty((B, eng_a.NP), dtype=torch.float64, device=params.device)
        fe_bar = torch.empty_like(fe)
        ws_a = eng_a._scratch("ws", max(eng_a.workspace_bytes(B), eng_b.workspace_bytes(B)))
        ws_b = eng_b._scratch("ws", eng_b.workspace_bytes(B))
        st = torch.cuda.current_stream(params.device).cuda_stream
        _ffi.check(_ffi.lib().tsff_ff_pair_bwd(
            eng_a._ctx, eng_b._ctx, B, params.data_ptr(), fe.data_ptr(), _ffi.TSFF_F32 if fe.dtype == torch.float32 else _ffi.TSFF_F64,
            saved_a.data_ptr(), saved_b.data_ptr(), bar_a.data_ptr(), bar_b.data_ptr(), params_bar.data_ptr(), fe_bar.data_ptr(),
            ws_a.data_ptr(), ws_b.data_ptr(), st))
        return None, None, params_bar, fe_bar


def form_factor_modl_pair(eng_a, eng_b, params, fe):
    """Differentiable (modl_a [B,W_a], modl_b [B,W_b]) of two table-mode engines on the same (params, fe)."""
    return _FFPairFunction.apply(eng_a, eng_b, params, fe)


def pair_compatible(eng_a, eng_b):
    return (eng_a.mode == "table" and eng_b.mode == "table" and eng_a.V == eng_b.V and eng_a.G == eng_b.G and eng_a.NP == eng_b.NP
            and eng_a.device == eng_b.device and not getattr(eng_a, "_cell_mode", 0) and not getattr(eng_b, "_cell_mode", 0))


class _FFFunction(torch.autograd.Function):
    """custom-VJP wrapper: forward saves the kernel's residual buffer, backward calls tsff_ff_bwd."""

    @staticmethod
    def forward(ctx, engine, params, fe, want_ff):
        modl, ff, saved = engine.forward(params, fe, want_ff=want_ff, want_modl=not want_ff)
        ctx.engine, ctx.want_ff = engine, want_ff
        ctx.save_for_backward(params, fe, saved)
        return ff if want_ff else modl

    @staticmethod
    def backward(ctx, out_bar):
        params, fe, saved = ctx.saved_tensors
        out_bar = out_bar.contiguous()
        if ctx.want_ff:
            pb, fb = ctx.engine.backward(params, fe, saved, ff_bar=out_bar, want_params=ctx.needs_input_grad[1])
        else:
            pb, fb = ctx.engine.backward(params, fe, saved, modl_bar=out_bar)
        return None, pb, fb, None


def form_factor_modl(engine, params, fe):
    """Differentiable angle-integrated spectrum modl[B,W]."""
    return _FFFunction.apply(engine, params, fe, False)


def form_factor_full(engine, params, fe):
    """Differentiable formfactor[B,G,W,A]."""
    return _FFFunction.apply(engine, params, fe, True)


# ---- B1: stand-alone PV integral ------------------------------------------------------------------------------
def pv_integral(f, z0, h, pole, precision="fp32", want_grad=True):
    """vmap(ratintn)(f, z[None]-pole[:,None], z) for uniform nodes (ratintn.py:4-23).  f [B,N], pole [B,P] f64 cuda."""
    _require_cuda(f, torch.float64, "f")
    _require_cuda(pole, torch.float64, "pole")
    B, N = f.shape
    P = pole.shape[1]
    out = torch.empty((B, P), dtype=torch.float64, device=f.device)
    dout = torch.empty((B, P), dtype=torch.float64, device=f.device) if want_grad else None
    ws = torch.empty(int(_ffi.lib().tsff_pv_workspace_bytes(B, N, P)), dtype=torch.uint8, device=f.device)
    st = torch.cuda.current_stream(f.device).cuda_stream
    _ffi.check(_ffi.lib().tsff_pv_fwd(B, N, P, f.data_ptr(), float(z0), float(h), pole.data_ptr(), out.data_ptr(),
                                      dout.data_ptr() if dout is not None else None,
                                      _ffi.TSFF_PV_FP64 if precision == "fp64" else _ffi.TSFF_PV_FP32, ws.data_ptr(), st))
    return out, dout


def pv_integral_vjp(f, z0, h, pole, out_bar):
    _require_cuda(out_bar, torch.float64, "out_bar")
    B, N = f.shape
    P = pole.shape[1]
    f_bar = torch.empty_like(f)
    pole_bar = torch.empty_like(pole)
    ws = torch.empty(int(_ffi.lib().tsff_pv_workspace_bytes(B, N, P)), dtype=torch.uint8, device=f.device)
    st = torch.cuda.current_stream(f.device).cuda_stream
    _ffi.check(_ffi.lib().tsff_pv_bwd(B, N, P, f.data_ptr(), float(z0), float(h), pole.data_ptr(), out_bar.data_ptr(),
                                      f_bar.data_ptr(), pole_bar.data_ptr(), ws.data_ptr(), st))
    return f_bar, pole_bar


def loss_fwd_bwd(theory, data, weight, uncert=1.0, scale=1.0, method="l2", loss_out=None, want_grad=True):
    """Fused masked loss + its gradient wrt theory (tsff_loss_fwd_bwd; loss_function.py:190-267, 386-418).
    theory, data [B,n]; weight [n] (static window weights); returns (loss device scalar, theory_bar)."""
    _require_cuda(theory, torch.float64, "theory")
    _require_cuda(data, torch.float64, "data")
    _require_cuda(weight, torch.float64, "weight")
    B, n = theory.shape
    if loss_out is None:
        loss_out = torch.zeros(1, dtype=torch.float64, device=theory.device)
    else:
        loss_out.zero_()
    tbar = torch.empty_like(theory) if want_grad else None
    st = torch.cuda.current_stream(theory.device).cuda_stream
    m = {"l2": 0, "l1": 1, "log-cosh": 2, "poisson": 3}[method]
    _ffi.check(_ffi.lib().tsff_loss_fwd_bwd(B, n, theory.data_ptr(), data.data_ptr(), weight.data_ptr(), float(uncert),
                                            float(scale), m, loss_out.data_ptr(), tbar.data_ptr() if want_grad else None, st))
    return loss_out, tbar


def microbench(kind, iters=4096):
    """Measured FFMA (kind 0) / MUFU.LG2 (kind 1) issue peak on the current device, ops per second."""
    sink = torch.zeros(4, dtype=torch.float32, device="cuda")
    ops = C.c_double()
    st = torch.cuda.current_stream().cuda_stream
    L = _ffi.lib()
    _ffi.check(L.tsff_microbench(kind, 64, C.byref(ops), sink.data_ptr(), st))  # warm
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0.0
    for _ in range(3):
        e0.record()
        _ffi.check(L.tsff_microbench(kind, iters, C.byref(ops), sink.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, ops.value / (e0.elapsed_time(e1) * 1e-3))
    return best
