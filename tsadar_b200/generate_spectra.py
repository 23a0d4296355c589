"""FitModel -- mirror of tsadar.core.physics.generate_spectra.FitModel (generate_spectra.py:8-220) for 1V distributions.
Temporal / imaging / 1d spectypes: the mean over gradient points, the weighted angle sum and the IAW filter are fused
into the form-factor kernel (`modl` output of tsff_ff_fwd).  "angular_full" (ARTS): the kernel returns the full
formfactor [G, W, A]; the mean over gradient points, the angular weight matrix product (generate_spectra.py:193-197) and
the IAW filter (:210-216) are one hand-written FP64 tiled contraction, tsff_arts_weights_fwd / _bwd (csrc/tsff_arts.cu)."""
from __future__ import annotations

import numpy as np
import torch

from . import _ffi
from .form_factor import FormFactor, pack_params


class _ArtsWeights(torch.autograd.Function):
    """modlE [NA, W] = jmul * (weights [NA, A] @ mean_g(ff [G, W, A]).T)  -- tsff_arts_weights_fwd, VJP tsff_arts_weights_bwd."""

    @staticmethod
    def forward(ctx, ff, wmat, jmul):
        if not ff.is_cuda:
            raise RuntimeError("tsadar_b200 has no CPU path")
        ff = ff.contiguous()
        G, W, A = ff.shape
        NA = wmat.shape[0]
        out = torch.empty((NA, W), dtype=torch.float64, device=ff.device)
        st = torch.cuda.current_stream(ff.device).cuda_stream
        _ffi.check(_ffi.lib().tsff_arts_weights_fwd(ff.data_ptr(), G, W, A, wmat.data_ptr(), NA,
                                                    jmul.data_ptr() if jmul is not None else None, out.data_ptr(), st))
        ctx.shape = (G, W, A)
        ctx.wmat, ctx.jmul = wmat, jmul
        return out

    @staticmethod
    def backward(ctx, g):
        G, W, A = ctx.shape
        g = g.contiguous()
        ffb = torch.empty((G, W, A), dtype=torch.float64, device=g.device)
        st = torch.cuda.current_stream(g.device).cuda_stream
        _ffi.check(_ffi.lib().tsff_arts_weights_bwd(g.data_ptr(), G, W, A, ctx.wmat.data_ptr(), ctx.wmat.shape[0],
                                                    ctx.jmul.data_ptr() if ctx.jmul is not None else None, ffb.data_ptr(), st))
        return ffb, None, None


def arts_weights(ff, wmat, jmul=None):
    """ff [G, W, A] float64 cuda, wmat [NA, A], jmul [W] or None (both float64 cuda, contiguous) -> modlE [NA, W]."""
    return _ArtsWeights.apply(ff, wmat, jmul)


class FitModel:
    # wavelength-axis sharding pays only when one image is many milliseconds of work: the 2V path (V^2 bicubic points per pole,
    # 130 ms per arts-2d image) always qualifies; the 1V table path (arts-1d: 0.16 ms of kernels per image) would be slowed
    # down by the all-gather / all-reduce latency (measured 2.41 -> 2.86 ms on two GPUs), so it needs this many poles
    SHARD_MIN_POLES_1V = 8_000_000

    def __init__(self, config, scattering_angles, mode="table", pv_precision="fp32", shard_group=False, force_shard=False):
        """shard_group: False = single GPU; None or a process group = "angular_full" spectra are evaluated W-sharded over
        that group's ranks (tsadar_b200/parallel.py), every rank returning the full modlE -- unless the image is too small
        for the collectives to pay (SHARD_MIN_POLES_1V; force_shard=True overrides)."""
        self.config = config
        self.scattering_angles = scattering_angles
        gen = config["parameters"]["general"]
        assert gen["Te_gradient"]["num_grad_points"] == gen["ne_gradient"]["num_grad_points"], \
            "Number of gradient points for Te and ne must be the same"
        G = gen["Te_gradient"]["num_grad_points"]
        self.dim = int(config["parameters"]["electron"]["fe"]["dim"])
        oth = config["other"]
        self.w_shard = None
        if shard_group is not False and config["other"]["extraoptions"]["spectype"] == "angular_full":
            from .parallel import WShard, _active
            nA_ = np.asarray(scattering_angles["sa"]).size
            worth = self.dim == 2 or force_shard or oth["npts"] * nA_ * G >= self.SHARD_MIN_POLES_1V
            if _active(shard_group) and worth:
                self.w_shard = WShard(oth["npts"], group=shard_group, halo=(self.dim == 1 and mode == "table"))
        self.electron_form_factor = FormFactor(oth["lamrangE"], npts=oth["npts"], lam_shift=config["data"]["ele_lam_shift"],
                                               scattering_angles=scattering_angles, num_grad_points=G,
                                               va_ang=config["parameters"]["general"].get("Va", {}).get("angle", 0.0) if self.dim == 2 else None,
                                               ud_ang=config["parameters"]["general"].get("ud", {}).get("angle", 0.0) if self.dim == 2 else None,
                                               mode=mode, pv_precision=pv_precision, w_shard=self.w_shard)
        self.ion_form_factor = FormFactor(oth["lamrangI"], npts=oth["npts"], lam_shift=0, scattering_angles=scattering_angles,
                                          num_grad_points=G, va_ang=None, ud_ang=None, mode=mode, pv_precision=pv_precision)
        # `weights[0]`: a scalar when `sa` comes straight from get_scattering_angles (tests, forward mode), the per-angle
        # vector after lineouts.py:103.  Both are "one weight per angle" for the kernel.
        self.angular_full = config["other"]["extraoptions"]["spectype"] == "angular_full"
        nA = np.asarray(scattering_angles["sa"]).size
        if self.angular_full:
            self._wmat = np.asarray(scattering_angles["weights"], dtype=np.float64)   # [1024, A]
            assert self._wmat.ndim == 2 and self._wmat.shape[1] == nA
            self._w = np.ones(nA)
            self._w_ion = np.ascontiguousarray(self._wmat[0])      # ion_spectrum always sums with weights[0] (:165): row 0 of the matrix
        else:
            w0 = np.asarray(scattering_angles["weights"])[0]
            self._w = np.full(nA, float(w0)) if np.ndim(w0) == 0 else np.asarray(w0, dtype=np.float64)
            self._w_ion = self._w
        lamE = np.linspace(oth["lamrangE"][0], oth["lamrangE"][1], oth["npts"])
        self._jmulE = None
        # iawoff (generate_spectra.py:199-208, "set the ion feature to 0"): as written the branch is ill-formed (its two indices
        # are swapped, so jnp.zeros gets a negative size) and no deck switches it on; its evident meaning -- the model zeroed
        # between the samples nearest lam - 3 nm and lam + 3 nm -- is applied after the kernel (_apply_iawoff)
        self._iawoff = bool(oth.get("iawoff", 0))
        f = oth.get("iawfilter", [0])
        if f[0]:
            fb, fr = f[3] - f[2] / 2, f[3] + f[2] / 2
            if oth["lamrangE"][0] < fr and oth["lamrangE"][1] > fb:
                self._jmulE = np.where((fb < lamE) & (fr > lamE), 10.0 ** (-f[1]), 1.0)   # generate_spectra.py:210-216

    def _apply_iawoff(self, modlE, block):
        """modlE [..., W] with the samples in [nearest(lam - 3), nearest(lam + 3)) zeroed, per parameter set (row of block)."""
        oth = self.config["other"]
        lamE = torch.linspace(oth["lamrangE"][0], oth["lamrangE"][1], oth["npts"], dtype=torch.float64, device=modlE.device)
        lam = block[:, _ffi.P_LAM].detach()
        i_lo = torch.argmin((lamE[None, :] - (lam[:, None] - 3.0)).abs(), dim=1)
        i_hi = torch.argmin((lamE[None, :] - (lam[:, None] + 3.0)).abs(), dim=1)
        j = torch.arange(oth["npts"], device=modlE.device)[None, :]
        inside = (j >= i_lo[:, None]) & (j < i_hi[:, None]) & ((lam > oth["lamrangE"][0]) & (lam < oth["lamrangE"][1]))[:, None]
        if modlE.dim() == 2 and modlE.shape[0] != block.shape[0]:       # angular_full: one parameter set, [NA, W] image
            inside = inside[:1]
        return torch.where(inside, torch.zeros((), dtype=modlE.dtype, device=modlE.device), modlE)

    def _spectrum_2v(self, ff_obj, all_params, jmul, weights):
        """dim == 2 with a non-ARTS spectype (generate_spectra.py:159-160, 187-188 + :164-165, 193, 197): calc_in_2D, mean over
        the gradient points, weighted angle sum -- the contraction kernel with a one-row weight matrix."""
        ff, _ = ff_obj.calc_in_2D(all_params)                                          # [G, W, A]
        dev = ff.device
        w = torch.tensor(np.ascontiguousarray(weights).reshape(1, -1), dtype=torch.float64, device=dev)
        jm = None if jmul is None else torch.tensor(np.ascontiguousarray(jmul), dtype=torch.float64, device=dev)
        return arts_weights(ff, w, jm)                                                 # [1, W]

    @staticmethod
    def _block_2v(all_params, dev):
        """the one-row parameter block of a 2V parameter set (the instrument stage reads lam and the amplitudes from it)"""
        p1 = {k: (dict(v) if isinstance(v, dict) else v) for k, v in all_params.items()}
        p1["electron"]["fe"] = torch.zeros(4, dtype=torch.float64)
        p1["electron"]["v"] = np.zeros(4)
        return pack_params(p1, dev)[0][:1].contiguous()

    def ion_spectrum(self, all_params):
        if self.config["other"]["extraoptions"]["load_ion_spec"] and self.dim == 2:
            modlI = self._spectrum_2v(self.ion_form_factor, all_params, None, self._w_ion)
            lamI = np.linspace(*self.config["other"]["lamrangI"], self.config["other"]["npts"])
            return lamI, modlI, self._block_2v(all_params, modlI.device)
        if self.config["other"]["extraoptions"]["load_ion_spec"]:
            modlI, block = self.ion_form_factor.modl(all_params, self._w_ion)
            lamI = np.linspace(*self.config["other"]["lamrangI"], self.config["other"]["npts"])
            return lamI, modlI, block
        return np.zeros(1), 0, None

    def electron_spectrum(self, all_params):
        if self.config["other"]["extraoptions"]["load_ele_spec"] and self.angular_full:
            # generate_spectra.py:185-188: 1V table or 2V table (calc_in_2D); [G, W, A], one parameter set per image
            ff, _ = self.electron_form_factor(all_params) if self.dim == 1 else self.electron_form_factor.calc_in_2D(all_params)
            if ff.dim() == 4:
                assert ff.shape[0] == 1, "angular_full takes a single parameter set (thomson_diagnostic.py:37-38)"
                ff = ff[0]
            dev = ff.device
            wm = getattr(self, "_wmat_dev", None)
            if wm is None or wm.device != dev:
                wm = self._wmat_dev = torch.tensor(self._wmat, dtype=torch.float64, device=dev).contiguous()
            jm = self._jmulE
            if self.w_shard is not None:                                         # this rank's wavelengths, halo dropped
                ff = ff[:, : self.w_shard.keep]
                jm = None if jm is None else jm[self.w_shard.j0:self.w_shard.j1]
            jkey = None if jm is None else (self.w_shard.j0 if self.w_shard is not None else 0)
            if jm is not None and (getattr(self, "_jm_dev", None) is None or self._jm_key != (jkey, dev)):
                self._jm_dev, self._jm_key = torch.tensor(np.ascontiguousarray(jm), dtype=torch.float64, device=dev), (jkey, dev)
            # mean over gradient points (:193), weights @ ThryE.T (:194-195), IAW filter (:210-216): one kernel
            modlE = arts_weights(ff, wm, self._jm_dev if jm is not None else None)
            if self.w_shard is not None:
                from .parallel import gather_columns
                modlE = gather_columns(modlE.contiguous(), self.w_shard.npts, self.w_shard.group)
            block, _, _, _, _ = pack_params(all_params, dev)
            if self._iawoff:
                modlE = self._apply_iawoff(modlE, block)
            lamE = np.linspace(*self.config["other"]["lamrangE"], self.config["other"]["npts"])
            return lamE, modlE, block
        if self.config["other"]["extraoptions"]["load_ele_spec"] and self.dim == 2:
            modlE = self._spectrum_2v(self.electron_form_factor, all_params, self._jmulE, self._w)
            block = self._block_2v(all_params, modlE.device)
            if self._iawoff:
                modlE = self._apply_iawoff(modlE, block)
            lamE = np.linspace(*self.config["other"]["lamrangE"], self.config["other"]["npts"])
            return lamE, modlE, block
        if self.config["other"]["extraoptions"]["load_ele_spec"]:
            modlE, block = self.electron_form_factor.modl(all_params, self._w, jmul=self._jmulE)
            if self._iawoff:
                modlE = self._apply_iawoff(modlE, block)
            lamE = np.linspace(*self.config["other"]["lamrangE"], self.config["other"]["npts"])
            return lamE, modlE, block
        return [], 0, None

    # ---- breakdown used by the reference's postprocessing plots (generate_spectra.py:222-330): the angle-integrated spectra
    # together with the raw formfactor [.., G, W, A] they were integrated from
    def ion_spectrum_detailed(self, all_params):
        if not self.config["other"]["extraoptions"]["load_ion_spec"]:
            return np.zeros(1), 0, 0
        ff = self.ion_form_factor
        ThryI, _ = ff(all_params) if self.dim == 1 else ff.calc_in_2D(all_params)
        lamI = np.linspace(*self.config["other"]["lamrangI"], self.config["other"]["npts"])
        w = torch.tensor(self._w, dtype=torch.float64, device=ThryI.device)
        modlI = (ThryI.mean(dim=-3) * w).sum(dim=-1)                                # :259-260
        return lamI, modlI, ThryI

    def electron_spectrum_detailed(self, all_params):
        if not self.config["other"]["extraoptions"]["load_ele_spec"]:
            return [], 0, 0
        ff = self.electron_form_factor
        ThryE, _ = ff(all_params) if self.dim == 1 else ff.calc_in_2D(all_params)
        dev = ThryE.device
        lamE = np.linspace(*self.config["other"]["lamrangE"], self.config["other"]["npts"])
        mean = ThryE.mean(dim=-3)                                                   # over gradient points  :296
        if self.angular_full:
            wm = torch.tensor(self._wmat, dtype=torch.float64, device=dev)
            modlE = torch.matmul(wm, mean.transpose(-1, -2))                        # :297-298
        else:
            modlE = (mean * torch.tensor(self._w, dtype=torch.float64, device=dev)).sum(dim=-1)   # :300
        if self._jmulE is not None:                                                 # :311-318
            jm = torch.tensor(self._jmulE, dtype=torch.float64, device=dev)
            modlE = modlE * jm
            inside = torch.tensor(self._jmulE != 1.0, device=dev)
            ThryE = torch.where(inside[:, None], ThryE * 1e-9, ThryE)               # the reference scales the raw one by 1e-9
        return lamE, modlE, ThryE

    def detailed_spectrum(self, all_params):
        lamAxisI, modlI, ThryI = self.ion_spectrum_detailed(all_params)
        lamAxisE, modlE, ThryE = self.electron_spectrum_detailed(all_params)
        return modlE, modlI, ThryE, ThryI, lamAxisE, lamAxisI

    def __call__(self, all_params):
        ex, oth = self.config["other"]["extraoptions"], self.config["other"]
        if (ex["load_ion_spec"] and ex["load_ele_spec"] and self.dim == 1 and not self.angular_full and self.w_shard is None
                and self.electron_form_factor.mode == "table" and self.ion_form_factor.mode == "table"):
            # both windows of one plasma: the f-dependent tables (form_factor.py:256-270) are built once, the principal-value
            # adjoint runs once (tsff_ff_pair_fwd / _bwd) -- the reference derives them in each FormFactor instance
            from .engine import form_factor_modl_pair, pair_compatible
            dev = torch.device("cuda", torch.cuda.current_device())
            block, fe, vx, _, nI = pack_params(all_params, dev)
            engE = self.electron_form_factor.engine(vx, nI, weights=self._w, jmul=self._jmulE)
            engI = self.ion_form_factor.engine(vx, nI, weights=self._w_ion)
            if pair_compatible(engE, engI):
                modlE, modlI = form_factor_modl_pair(engE, engI, block, fe)
                if self._iawoff:
                    modlE = self._apply_iawoff(modlE, block)
                return modlE, modlI, np.linspace(*oth["lamrangE"], oth["npts"]), np.linspace(*oth["lamrangI"], oth["npts"])
        lamAxisI, modlI, _ = self.ion_spectrum(all_params)
        lamAxisE, modlE, _ = self.electron_spectrum(all_params)
        return modlE, modlI, lamAxisE, lamAxisI
