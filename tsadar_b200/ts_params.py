"""ThomsonParams -- torch mirror of tsadar.core.modules.ts_params.ThomsonParams (ts_params.py:498-645) :
normalised trainable leaves, sigmoid/affine de-normalisation, ion-fraction renormalisation, and the f(v) producers
DLM1V / 'mx' (here) and Arbitrary1V / Arbitrary2V / SphericalHarmonics (tsadar_b200/distribution_functions.py).

The reference keeps this stage in JAX (out of scope for the kernels, "next" row N3); it is
mirrored here only so that the drop-in tests read like the reference's.  It is cheap elementwise host-side glue in torch
(autograd carries the kernels' cotangents back to the normalised leaves)."""
from __future__ import annotations

import numpy as np
import torch
from scipy.special import gamma, gammaincc

from .distribution_functions import Arbitrary1V, Arbitrary2V, SphericalHarmonics

DT = torch.float64


def _inv_act(x):
    return np.log(1e-2 + x / (1 - x + 1e-2))  # ts_params.py:344 (not the logit: values shift on the first call)


def vgrid(nvx):
    vmax = 6.0
    dv = 2 * vmax / nvx
    return np.linspace(-vmax + dv / 2, vmax - dv / 2, nvx)


def dlm_table(vx):
    """Stand-in for the missing blob DLM_x_-3_-10_10_m_-1_2_5.mat (base.py:266-272): projected super-Gaussians on
    vx_ax = linspace(-10,10,20001) x m_ax = linspace(2,5,31), lerped in v onto vx."""
    vx_ax = np.linspace(-10, 10, 20001)
    m_ax = np.linspace(2, 5, 31)
    cols = []
    for m in m_ax:
        alpha = np.sqrt(3.0 * gamma(3.0 / m) / 2.0 / gamma(5.0 / m))
        it = gamma(2.0 / m) * gammaincc(2.0 / m, (np.abs(vx_ax) / (alpha * np.sqrt(2.0))) ** m)
        cols.append(np.interp(vx, vx_ax, it))
    return m_ax, np.stack(cols, axis=1)


class _Scalar:
    """One (possibly batched) scalar parameter: stored normalised, trainable when active."""

    def __init__(self, cfg, batch_size, activate, device, raw=False):
        self.active = bool(cfg.get("active", False)) and activate
        self.raw = raw
        self.scale = 1.0 if raw else cfg["ub"] - cfg["lb"]
        self.shift = 0.0 if raw else cfg["lb"]
        x = np.full(batch_size, (cfg["val"] - self.shift) / self.scale, dtype=np.float64)
        if self.active:
            x = _inv_act(x)
        self.value = torch.tensor(x, dtype=DT, device=device, requires_grad=self.active)

    def physical(self):
        v = torch.sigmoid(self.value) if self.active else self.value
        return v * self.scale + self.shift


class ThomsonParams:
    def __init__(self, param_cfg, num_params, batch=True, activate=False, device=None, dlm_m_offset=0.0):
        self.param_cfg = param_cfg
        self.B = int(num_params) if batch else 1
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        B, dev = self.B, self.device
        e = param_cfg["electron"]
        self.leaves = {("electron", k): _Scalar(e[k], B, activate, dev) for k in ["Te", "ne"]}
        for k in ["lam", "amp1", "amp2", "amp3", "ne_gradient", "Te_gradient", "ud", "Va"]:
            self.leaves[("general", k)] = _Scalar(param_cfg["general"][k], B, activate, dev)
        self.ions = sorted([k for k in param_cfg if k.startswith("ion-")], key=lambda s: int(s.split("-")[1]))
        assert self.ions, "No ion species found in input deck"
        for ion in self.ions:
            c = param_cfg[ion]
            self.leaves[(ion, "Ti")] = _Scalar(c["Ti"], B, activate, dev)
            self.leaves[(ion, "Z")] = _Scalar(c["Z"], B, activate, dev)
            self.leaves[(ion, "fract")] = _Scalar(c["fract"], B, activate, dev, raw=True)
        fe = e["fe"]
        self.fe_type = fe["type"].casefold()
        self.fe_dim = int(fe["dim"])
        self.vx = vgrid(fe["nvx"])
        self.dv = self.vx[1] - self.vx[0]
        self.dist = None
        fe_trainable = bool(fe.get("active", False))      # get_filter_spec: fe leaves follow cfg["fe"]["active"] alone
        if self.fe_dim == 2:
            # ElectronParams.init_dists (ts_params.py:155-166): one table per image, no batch mode
            if batch:
                raise NotImplementedError("Batch mode not implemented for 2D distributions as a precautionary measure against memory issues")
            if "sph" in self.fe_type:
                self.dist = SphericalHarmonics(fe, dev, fe_trainable)
            elif self.fe_type == "arbitrary":
                self.dist = Arbitrary2V(fe, dev, fe_trainable)
            else:
                raise NotImplementedError(f"Unknown 2D distribution type: {fe['type']}")
        elif self.fe_dim != 1:
            raise NotImplementedError(f"Not implemented distribution dimension: {fe['dim']}")
        elif self.fe_type == "dlm":
            mcfg = dict(val=fe["params"]["m"]["val"], lb=2.0, ub=5.0, active=fe.get("active", False))  # scale 3, shift 2
            self.leaves[("electron", "m")] = _Scalar(mcfg, B, activate, dev)
            self.m_offset = float(dlm_m_offset)
            m_ax, tab = dlm_table(self.vx)
            self.m_ax = torch.tensor(m_ax, dtype=DT, device=dev)
            self.f_vx_m = torch.tensor(tab, dtype=DT, device=dev)  # [V, 31]
        elif self.fe_type == "mx":
            f = np.exp(-(self.vx**2 / 2))
            self.f_fixed = torch.tensor(f / f.sum() / self.dv, dtype=DT, device=dev)
        elif self.fe_type == "arbitrary":
            self.dist = Arbitrary1V(fe, B, dev, fe_trainable)
        else:
            raise NotImplementedError(f"Unknown 1D distribution type: {fe['type']}")

    def parameters(self):
        """The trainable leaves (what eqx.partition(ts_params, get_filter_spec(...)) selects, ts_params.py:648-685)."""
        out = [s.value for s in self.leaves.values() if s.active]
        if self.dist is not None:
            out += [t for t in self.dist.leaves().values() if t.requires_grad]
        return out

    def _fe(self):
        if self.dist is not None:
            return self.dist()
        if self.fe_type == "mx":
            return self.f_fixed.reshape(1, -1).expand(self.B, -1)
        m = self.leaves[("electron", "m")].physical() + self.m_offset          # base.py:286
        i = torch.clamp(torch.searchsorted(self.m_ax, m.detach().contiguous(), right=True), 1, self.m_ax.numel() - 1)
        w = (m - self.m_ax[i - 1]) / (self.m_ax[i] - self.m_ax[i - 1])
        f = self.f_vx_m[:, i - 1].T * (1 - w)[:, None] + self.f_vx_m[:, i].T * w[:, None]   # jnp.interp in m (base.py:292)
        return f / f.sum(dim=1, keepdim=True) / self.dv                                        # base.py:294

    def __call__(self):
        out = {"electron": {"Te": self.leaves[("electron", "Te")].physical(), "ne": self.leaves[("electron", "ne")].physical(),
                            "fe": self._fe(), "v": self.vx if self.fe_dim == 2 else np.broadcast_to(self.vx, (self.B, self.vx.size))},
               "general": {k: self.leaves[("general", k)].physical() for k in
                           ["lam", "amp1", "amp2", "amp3", "ne_gradient", "Te_gradient", "ud", "Va"]}}
        fsum = 0
        for n, ion in enumerate(self.ions):
            c = self.param_cfg[ion]
            out[ion] = {"A": torch.full((self.B,), float(c["A"]["val"]), dtype=DT, device=self.device),
                        "fract": self.leaves[(ion, "fract")].physical(), "Ti": self.leaves[(ion, "Ti")].physical(),
                        "Z": self.leaves[(ion, "Z")].physical()}
            if n > 0 and c["Ti"].get("same", False):
                out[ion]["Ti"] = out["ion-1"]["Ti"]                           # ts_params.py:555-557
            fsum = fsum + out[ion]["fract"]
        for ion in self.ions:
            out[ion]["fract"] = out[ion]["fract"] / fsum                      # ts_params.py:559-561
        return out

    def get_unnormed_params(self):
        """Physical values of the parameters WITHOUT the f table and its grid (ts_params.py:168-200, 565-581): Te, ne and
        what the distribution itself reports -- m (DLM), f (arbitrary tables), flm (spherical harmonics); a Maxwellian
        reports nothing."""
        p = self()
        fe = p["electron"].pop("fe")
        p["electron"].pop("v")
        if self.fe_dim == 1 and self.fe_type == "dlm":
            p["electron"]["m"] = self.leaves[("electron", "m")].physical()
        elif self.fe_type == "arbitrary":
            p["electron"]["f"] = fe
        elif isinstance(self.dist, SphericalHarmonics):
            p["electron"].update(self.dist.get_unnormed_params())
        return p

    def get_fitted_params(self, param_cfg):
        """(fitted_params, num_params): the active entries of get_unnormed_params (ts_params.py:605-645); m counts when the
        distribution is active, f / flm are always reported (flm with the assembled table and its grid)."""
        param_dict = self.get_unnormed_params()
        num_params = 0
        fitted = {}
        for k in param_dict:
            fitted[k] = {}
            for k2 in param_dict[k]:
                if k2 == "m":
                    if param_cfg[k]["fe"]["active"]:
                        fitted[k][k2] = param_dict[k][k2]
                        num_params += 1
                elif k2 in ("f", "fe", "flm"):
                    fitted[k][k2] = param_dict[k][k2]
                    if k2 == "flm":
                        out = self()
                        fitted[k][k2]["fvxvy"] = out["electron"]["fe"]
                        fitted[k][k2]["v"] = out["electron"]["v"]
                elif param_cfg[k][k2]["active"]:
                    fitted[k][k2] = param_dict[k][k2]
                    num_params += 1
        return fitted, num_params


# ---- fused path: the transforms + the DLM1V producer as ONE kernel each way ----------------------
class _ParamsFn(torch.autograd.Function):
    """(x_active [B, NLA]) -> (params block [B, NP], fe [B, V]) through tsff_params_fwd; VJP tsff_params_bwd."""

    @staticmethod
    def forward(ctx, xa, owner):
        from . import _ffi
        B = xa.shape[0]
        block = torch.empty((B, owner.NP), dtype=DT, device=xa.device)
        fe = torch.empty((B, owner.V), dtype=owner.fe_dtype, device=xa.device)
        st = torch.cuda.current_stream(xa.device).cuda_stream
        _ffi.check(_ffi.lib().tsff_params_fwd(owner._cfg_ref(), B, xa.data_ptr(), owner.x_static.data_ptr(), block.data_ptr(),
                                              fe.data_ptr(), st))
        ctx.owner = owner
        ctx.save_for_backward(xa)
        return block, fe

    @staticmethod
    def backward(ctx, block_bar, fe_bar):
        from . import _ffi
        (xa,) = ctx.saved_tensors
        owner = ctx.owner
        B = xa.shape[0]
        if block_bar is None:
            block_bar = torch.zeros((B, owner.NP), dtype=DT, device=xa.device)
        out = torch.empty_like(xa)
        st = torch.cuda.current_stream(xa.device).cuda_stream
        _ffi.check(_ffi.lib().tsff_params_bwd(owner._cfg_ref(), B, xa.data_ptr(), owner.x_static.data_ptr(),
                                              block_bar.contiguous().data_ptr(),
                                              fe_bar.contiguous().data_ptr() if fe_bar is not None else None, out.data_ptr(), st))
        return out, None


class FusedThomsonParams:
    """ThomsonParams for the 1-D DLM / Maxwellian decks with every per-step stage on the device in two launches:
    `__call__` = tsff_params_fwd (de-normalisation of all leaves, ion-fraction renormalisation, DLM table lerp in m and
    normalisation), its reverse = tsff_params_bwd.  All trainable leaves of all lineouts live in ONE tensor `x` [B, NLA]
    (column order `active_names`), which is what the fused optimiser update (tsff_adam_step, fit.fused_adam_fit) advances.
    Values are identical to ThomsonParams (tests/test_gpu_params_kernels.py).  The returned dict carries the standard entries
    (views of the block) and the pre-packed operands of the form-factor call, so nothing is re-stacked downstream."""

    def __init__(self, param_cfg, num_params, batch=True, activate=False, device=None, dlm_m_offset=0.0, fe_dtype=torch.float64):
        import ctypes as C
        from . import _ffi
        ref = ThomsonParams(param_cfg, num_params, batch=batch, activate=activate, device=device, dlm_m_offset=dlm_m_offset)
        if ref.dist is not None or ref.fe_dim != 1:
            raise NotImplementedError("FusedThomsonParams covers the DLM and Maxwellian 1-D distributions; use ThomsonParams for the others")
        self.param_cfg, self.B, self.device, self.ions = param_cfg, ref.B, ref.device, ref.ions
        self.vx, self.dv, self.fe_type, self.fe_dim, self.fe_dtype = ref.vx, ref.dv, ref.fe_type, 1, fe_dtype
        I = len(self.ions)
        self.NP, self.V = _ffi.P_ION0 + _ffi.ION_STRIDE * I, int(self.vx.size)
        keys = [("electron", "Te"), ("electron", "ne")] + [("general", k) for k in ["lam", "Va", "ud", "ne_gradient", "Te_gradient", "amp1", "amp2", "amp3"]]
        for ion in self.ions:
            keys += [(ion, "Z"), (ion, "Ti"), (ion, "fract")]
        has_m = self.fe_type == "dlm"
        keys.append(("electron", "m"))
        self.NL = len(keys)
        cfg = _ffi.ParamsCfg()
        cfg.I, cfg.V, cfg.fe_dtype = I, self.V, (_ffi.TSFF_F32 if fe_dtype == torch.float32 else _ffi.TSFF_F64)
        cfg.dv, cfg.m_offset = float(self.dv), float(dlm_m_offset)
        xs = torch.zeros((self.B, self.NL), dtype=DT, device=self.device)
        cols, self.active_names = [], []
        for k, key in enumerate(keys):
            s = ref.leaves.get(key)
            cfg.active_slot[k], cfg.scale[k], cfg.shift[k] = -1, 1.0, 0.0
            if s is None:
                continue
            cfg.scale[k], cfg.shift[k] = float(s.scale), float(s.shift)
            if s.active:
                cfg.active_slot[k] = len(cols)
                cols.append(s.value.detach())
                self.active_names.append(key)
            else:
                xs[:, k] = s.value.detach()
        cfg.NLA = len(cols)
        for i, ion in enumerate(self.ions):
            cfg.ionA[i] = float(param_cfg[ion]["A"]["val"])
            cfg.ti_same[i] = int(bool(i > 0 and param_cfg[ion]["Ti"].get("same", False)))
        if has_m:
            self._tab = ref.f_vx_m.t().contiguous()                   # [31][V]
            cfg.nm, cfg.m0, cfg.dm = int(ref.m_ax.numel()), float(ref.m_ax[0]), float(ref.m_ax[1] - ref.m_ax[0])
        else:
            self._tab = (ref.f_fixed * 1.0).reshape(1, -1).contiguous()
            cfg.nm, cfg.m0, cfg.dm = 1, 0.0, 1.0
        cfg.f_vx_m = self._tab.data_ptr()
        self._cfg, self._C = cfg, C
        self.x_static = xs
        self.x = (torch.stack(cols, dim=1).contiguous() if cols else torch.zeros((self.B, 0), dtype=DT, device=self.device)).requires_grad_(bool(cols))
        self.NLA = cfg.NLA
        self._keys = keys

    def _cfg_ref(self):
        return self._C.byref(self._cfg)

    def parameters(self):
        return [self.x] if self.NLA else []

    def physical(self):
        """-> (block [B, NP], fe [B, V]), differentiable with respect to `x`."""
        return _ParamsFn.apply(self.x, self)

    def __call__(self):
        from . import _ffi
        block, fe = self.physical()
        out = {"electron": {"Te": block[:, _ffi.P_TE], "ne": block[:, _ffi.P_NE], "fe": fe, "v": np.broadcast_to(self.vx, (self.B, self.vx.size))},
               "general": {"lam": block[:, _ffi.P_LAM], "Va": block[:, _ffi.P_VA], "ud": block[:, _ffi.P_UD], "ne_gradient": block[:, _ffi.P_NE_GRAD],
                           "Te_gradient": block[:, _ffi.P_TE_GRAD], "amp1": block[:, _ffi.P_AMP1], "amp2": block[:, _ffi.P_AMP2],
                           "amp3": block[:, _ffi.P_AMP3]}}
        for i, ion in enumerate(self.ions):
            o = _ffi.P_ION0 + _ffi.ION_STRIDE * i
            out[ion] = {"A": block[:, o + _ffi.ION_A], "Z": block[:, o + _ffi.ION_Z], "Ti": block[:, o + _ffi.ION_TI], "fract": block[:, o + _ffi.ION_FRACT]}
        out["_packed"] = (block, fe, self.vx, True, len(self.ions))
        return out
