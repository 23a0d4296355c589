"""ThomsonScatteringDiagnostic -- mirror of tsadar.core.thomson_diagnostic (thomson_diagnostic.py:10-142) for the
temporal / imaging / 1d spectypes: FitModel + instrument response + noise, batched over lineouts."""
from __future__ import annotations

import torch

from . import irf
from .generate_spectra import FitModel


def _dev_vec(x, B, dev):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float64)
    return t.to(device=dev, dtype=torch.float64).reshape(-1).expand(B).contiguous()


class ThomsonScatteringDiagnostic:
    def __init__(self, cfg, scattering_angles, mode="table", pv_precision="fp32"):
        self.cfg = cfg
        self.scattering_angles = scattering_angles
        st = cfg["other"]["extraoptions"]["spectype"]
        if not ("temporal" in st or "imaging" in st or "1d" in st):
            raise NotImplementedError(f"spectype {st}: only the vmapped 1V spectypes are built (DESIGN.md: scope)")
        self.model = FitModel(cfg, scattering_angles, mode=mode, pv_precision=pv_precision)

    def __call__(self, ts_params, batch):
        """-> ThryE [B,1024], ThryI [B,1024], lamAxisE, lamAxisI   (thomson_diagnostic.py:109-142)"""
        physical_params = ts_params() if callable(ts_params) else ts_params
        oth = self.cfg["other"]
        ex = oth["extraoptions"]
        dev = torch.device("cuda", torch.cuda.current_device())
        ThryE = ThryI = 0
        lamAxisE, lamAxisI = [], []
        if ex["load_ion_spec"]:
            lamI, modlI, block = self.model.ion_spectrum(physical_params)
            B = modlI.shape[0]
            noise = self._noise(batch["noise_i"], B, dev)
            lamAxisI, ThryI = irf.add_ion_IRF(self.cfg, oth["lamrangI"], oth["npts"], modlI, _dev_vec(batch["i_amps"], B, dev), block, noise)
        if ex["load_ele_spec"]:
            lamE, modlE, block = self.model.electron_spectrum(physical_params)
            B = modlE.shape[0]
            noise = self._noise(batch["noise_e"], B, dev)
            lamAxisE, ThryE = irf.add_electron_IRF(self.cfg, oth["lamrangE"], oth["npts"], modlE, _dev_vec(batch["e_amps"], B, dev), block, noise)
        return ThryE, ThryI, lamAxisE, lamAxisI

    @staticmethod
    def _noise(n, B, dev, nbins=1024):
        t = n if isinstance(n, torch.Tensor) else torch.as_tensor(n, dtype=torch.float64)
        t = t.to(device=dev, dtype=torch.float64)
        if t.numel() == 1 and float(t.reshape(-1)[0]) == 0.0:
            return None
        return t.expand(B, nbins).contiguous() if t.dim() < 2 or t.shape != (B, nbins) else t.contiguous()
