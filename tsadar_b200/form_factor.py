"""FormFactor -- host-side mirror of tsadar.core.physics.form_factor.FormFactor (form_factor.py:48-587) for the 1V path.

Same constructor arguments and the same `__call__(params) -> (formfactor, lams)` contract; `params` is the nested dict
the reference's ThomsonParams.__call__ returns (ts_params.py:599-603) with torch CUDA tensors as leaves.  Scalars
(shape ()) mean one lineout, as seen by the reference inside `vmap`; leaves of shape [B] are a batch of lineouts (what
`vmap(FitModel)` maps over, thomson_diagnostic.py:35) and add a leading batch axis to the outputs.
All arithmetic is done by the CUDA kernels (tsff_ff_fwd / tsff_ff_bwd); gradients flow through torch autograd via the
custom Function in engine.py."""
from __future__ import annotations

import numpy as np
import torch

from . import _ffi
from .engine import FormFactorEngine, form_factor_full, form_factor_modl

C_CM = 2.99792458e10


def pack_params(params, device):
    """nested dict -> (block [B, NP] float64, fe [B, V], vx numpy [V], batched?)"""
    if "_packed" in params:            # FusedThomsonParams: the block and the table come out of tsff_params_fwd already packed
        return params["_packed"]
    ions = sorted([k for k in params if k.startswith("ion-")], key=lambda s: int(s.split("-")[1]))
    ele, gen = params["electron"], params["general"]

    def T(x):
        return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float64)

    cols = [T(ele["Te"]), T(ele["ne"]), T(gen["lam"]), T(gen["Va"]), T(gen["ud"]), T(gen["ne_gradient"]),
            T(gen["Te_gradient"]), T(gen["amp1"]), T(gen["amp2"]), T(gen["amp3"])]
    for k in ions:
        cols += [T(params[k]["A"]), T(params[k]["Z"]), T(params[k]["Ti"]), T(params[k]["fract"])]
    batched = any(c.dim() > 0 and c.numel() > 1 for c in cols) or T(ele["fe"]).dim() == 2
    B = max([c.numel() for c in cols] + [T(ele["fe"]).shape[0] if T(ele["fe"]).dim() == 2 else 1])
    cols = [c.to(device=device, dtype=torch.float64).reshape(-1).expand(B) for c in cols]
    block = torch.stack(cols, dim=1).contiguous()
    fe = T(ele["fe"]).to(device)
    fe = fe.reshape(1, -1).expand(B, -1) if fe.dim() == 1 else fe
    if fe.dtype not in (torch.float32, torch.float64):
        fe = fe.double()
    v = ele["v"]
    v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    vx = np.asarray(v[0] if v.ndim == 2 else v, dtype=np.float64)
    return block, fe.contiguous(), vx, batched, len(ions)


class FormFactor:
    def __init__(self, lambda_range, npts, lam_shift, scattering_angles, num_grad_points, ud_ang, va_ang, mode="table",
                 pv_precision="fp32", w_shard=None):
        # form_factor.py:120-161 -- the static grids live in the libtsff context, created lazily once the f-grid is seen
        self.lambda_range = [float(lambda_range[0]), float(lambda_range[1])]
        self.npts = int(npts)
        self.lam_shift = float(lam_shift)
        self.scattering_angles = scattering_angles
        self.num_grad_points = int(num_grad_points)
        self.ud_angle, self.va_angle = ud_ang, va_ang
        self.mode, self.pv_precision = mode, pv_precision
        self._engines = {}
        # w_shard (parallel.WShard): this instance evaluates only its rank's wavelengths [j0, j1e) of the axis; operands pass
        # through copy_to_shards so that their cotangents are summed over the ranks in the backward pass
        self.w_shard = w_shard
        self._w_slice = None if w_shard is None else (w_shard.j0, w_shard.j1e)
        lam = np.linspace(self.lambda_range[0], self.lambda_range[1], self.npts)
        if self._w_slice is not None:
            lam = lam[self._w_slice[0]:self._w_slice[1]]
        omgs = 2e7 * np.pi * C_CM / lam
        self._lams = (2 * np.pi * C_CM / omgs)[None, :, None]

    def _lams_on(self, dev):
        """the wavelength axis on `dev`, uploaded once (a per-call host-to-device copy would also break CUDA-graph capture)"""
        c = getattr(self, "_lams_dev", None)
        if c is None or c.device != dev:
            c = self._lams_dev = torch.as_tensor(self._lams, device=dev)
        return c

    def _enter(self, t):
        if self.w_shard is None:
            return t
        from .parallel import copy_to_shards
        return copy_to_shards(t, self.w_shard.group)

    def engine(self, vx, n_ions, weights=None, jmul=None):
        key = (vx.size, float(vx[0]), float(vx[1] - vx[0]), n_ions, None if weights is None else tuple(np.ravel(weights)),
               None if jmul is None else hash(np.asarray(jmul).tobytes()))
        if key not in self._engines:
            sa = np.asarray(self.scattering_angles["sa"], dtype=np.float64).reshape(-1)
            w = np.ones_like(sa) if weights is None else weights
            self._engines[key] = FormFactorEngine(self.lambda_range, self.npts, self.lam_shift, sa, w, self.num_grad_points,
                                                  n_ions, vx, mode=self.mode, jmul=jmul, pv_precision=self.pv_precision,
                                                  w_slice=self._w_slice)
        return self._engines[key]

    def __call__(self, params):
        """-> formfactor [G,W,A] (or [B,G,W,A]), lams [1,W,1]   (form_factor.py:163-298)"""
        dev = torch.device("cuda", torch.cuda.current_device())
        block, fe, vx, batched, nI = pack_params(params, dev)
        eng = self.engine(vx, nI)
        ff = form_factor_full(eng, self._enter(block), self._enter(fe))
        return (ff if batched else ff[0]), self._lams_on(dev)

    def modl(self, params, weights, jmul=None):
        """Fused FormFactor + FitModel angle integration -> modl [B,W] (generate_spectra.py:164-165,193,197)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        block, fe, vx, batched, nI = pack_params(params, dev)
        eng = self.engine(vx, nI, weights=weights, jmul=jmul)
        return form_factor_modl(eng, block, fe), block

    def _engine_2v(self, vx, nI):
        key = ("2v", vx.size, float(vx[0]), float(vx[1] - vx[0]), nI)
        if key not in self._engines:
            sa = np.asarray(self.scattering_angles["sa"], dtype=np.float64).reshape(-1)
            self._engines[key] = FormFactorEngine(self.lambda_range, self.npts, self.lam_shift, sa, np.ones_like(sa),
                                                  self.num_grad_points, nI, vx, mode="2v", ud_ang=self.ud_angle, va_ang=self.va_angle,
                                                  w_slice=self._w_slice)
        return self._engines[key]

    def calc_all_chi_vals(self, x, DF, beta, xie_mag, klde_mag):
        """-> fe_vphi, chiEI, chiERrat, each shaped like beta   (form_factor.py:390-447; per pole :349-388).
        x: the velocity grid vx[V]; DF: the table [V,V]; beta, xie_mag [G,W,A]; klde_mag [G,W,A] or [G,W,A,1].
        Forward only (the reference differentiates this stage through calc_in_2D)."""
        dev = torch.device("cuda", torch.cuda.current_device())

        def T(t):
            t = t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t), dtype=torch.float64)
            return t.detach().to(device=dev, dtype=torch.float64).contiguous()

        vx = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        vx = np.asarray(vx, dtype=np.float64).reshape(-1)
        beta = T(beta)
        klde = T(klde_mag).reshape(beta.shape)             # the reference carries a trailing unit axis (:565)
        eng = self._engine_2v(vx, 1)
        return eng.chi_vals_2v(T(DF), beta, T(xie_mag).reshape(beta.shape), klde)

    def calc_in_2D(self, params):
        """-> formfactor [G,W,A], lams [1,W,1]   (form_factor.py:449-587).  params["electron"]["fe"] is the 2-D table
        DF[V,V] on vx x vx (one parameter set, as in the reference's angular spectypes).  Differentiable: the custom
        Function calls tsff_ff_bwd (rotate/project adjoint, d/dbeta, kinematics reverse)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        ele = params["electron"]
        fe = ele["fe"] if isinstance(ele["fe"], torch.Tensor) else torch.as_tensor(np.asarray(ele["fe"]), dtype=torch.float64)
        fe = fe.to(device=dev, dtype=torch.float64)
        fe = fe.reshape((-1,) + tuple(fe.shape[-2:]))
        assert fe.shape[0] == 1 and fe.shape[1] == fe.shape[2], "calc_in_2D takes one 2-D table [V, V]"
        p1 = {k: (dict(v) if isinstance(v, dict) else v) for k, v in params.items()}
        p1["electron"] = dict(ele)
        p1["electron"]["fe"] = torch.zeros(fe.shape[1], dtype=torch.float64)     # placeholder row for pack_params
        v = ele["v"]
        v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        p1["electron"]["v"] = np.asarray(v, dtype=np.float64).reshape(-1)[: fe.shape[1]] if np.ndim(v) == 1 else np.asarray(v[0], dtype=np.float64)
        block, _, vx, _, nI = pack_params(p1, dev)
        eng = self._engine_2v(vx, nI)
        ff = form_factor_full(eng, self._enter(block[:1].contiguous()), self._enter(fe.contiguous()))
        return ff[0], self._lams_on(dev)
