"""Angular (ARTS) instrument stage -- mirror of irf.add_ATS_IRF (irf.py:5-47) followed by
ThomsonScatteringDiagnostic.reduce_ATS_to_resunit and the noise add (thomson_diagnostic.py:78-107, 139) on the CUDA
kernels tsff_ats_fwd / tsff_ats_bwd."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ffi


def gaussian_taps(axis, fwhm):
    """The reference's full-length tap vector (irf.py:22-33) and the index range outside which it is < 1e-30 of its peak."""
    axis = np.asarray(axis, dtype=np.float64).reshape(-1)
    stddev = fwhm / 2.3548
    origin = (axis.max() + axis.min()) / 2.0
    taps = (1.0 / (stddev * np.sqrt(2.0 * np.pi))) * np.exp(-((axis - origin) ** 2.0) / (2.0 * stddev**2.0))
    nz = np.nonzero(taps > 1e-30 * taps.max())[0]
    return taps, int(nz[0]), int(nz[-1])


class AtsStage:
    """Static state of the ATS stage for one deck: tap vectors on the device and the geometry of the reduction."""

    def __init__(self, config, scattering_angles, lam_range, W, n_lam_data, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        oth = config["other"]
        wid = oth["PhysParams"]["widIRF"]
        ang_axis = np.asarray(scattering_angles["angAxis"], dtype=np.float64).reshape(-1)
        lam_axis = np.linspace(lam_range[0], lam_range[1], W)
        ta, a0, a1 = gaussian_taps(ang_axis, wid["ang_FWHM_ele"])
        tl, l0, l1 = gaussian_taps(lam_axis, wid["spect_FWHM_ele"])
        self._ta = torch.tensor(ta, dtype=torch.float64, device=self.device)
        self._tl = torch.tensor(tl, dtype=torch.float64, device=self.device)
        c = _ffi.AtsCfg()
        c.NA, c.W = int(ang_axis.size), int(W)
        c.lam_step = int(round(W / n_lam_data))                      # thomson_diagnostic.py:93
        c.ang_step = int(round(ang_axis.size / oth["CCDsize"][0]))   # thomson_diagnostic.py:94
        c.row_start, c.row_end = int(config["data"]["lineouts"]["start"]), int(config["data"]["lineouts"]["end"])
        c.norm = int(oth["PhysParams"]["norm"])
        c.ang_t0, c.ang_t1, c.lam_t0, c.lam_t1 = a0, a1, l0, l1
        c.lam_min, c.lam_max = float(lam_range[0]), float(lam_range[1])
        c.taps_ang, c.taps_lam = self._ta.data_ptr(), self._tl.data_ptr()
        self.cfg = c
        na = -(-c.NA // c.ang_step)
        c.row_end = min(c.row_end, na)
        self.nrows = c.row_end - c.row_start
        self.nl = -(-c.W // c.lam_step)
        self.lam_units = np.array([lam_axis[i:i + c.lam_step].mean() for i in range(0, W, c.lam_step)])
        L = _ffi.lib()
        self.saved_bytes = int(L.tsff_ats_saved_bytes(C.byref(c)))
        self.ws_bytes = int(L.tsff_ats_workspace_bytes(C.byref(c)))
        if self.saved_bytes == 0:
            raise RuntimeError(f"libtsff: {L.tsff_last_error().decode()}")
        self._ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)

    def __call__(self, modlE, block, e_amps, noise=None):
        """modlE [NA, W], block [1, NP] (physical parameter row), e_amps [nrows] -> ThryE [nrows, n_lam_units]."""
        return _AtsFunction.apply(modlE.contiguous(), block, e_amps.contiguous(), noise, self)


class _AtsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, modl, block, e_amps, noise, stage):
        L = _ffi.lib()
        dev = modl.device
        assert modl.shape == (stage.cfg.NA, stage.cfg.W) and modl.dtype == torch.float64 and modl.is_cuda
        assert e_amps.numel() == stage.nrows and e_amps.dtype == torch.float64
        thry = torch.empty((stage.nrows, stage.nl), dtype=torch.float64, device=dev)
        saved = torch.empty(stage.saved_bytes, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _ffi.check(L.tsff_ats_fwd(C.byref(stage.cfg), modl.data_ptr(), block.data_ptr(), e_amps.data_ptr(),
                                  noise.data_ptr() if noise is not None else None, thry.data_ptr(), saved.data_ptr(),
                                  stage._ws.data_ptr(), st))
        ctx.stage = stage
        ctx.save_for_backward(block, e_amps, saved)
        return thry

    @staticmethod
    def backward(ctx, thry_bar):
        block, e_amps, saved = ctx.saved_tensors
        stage = ctx.stage
        L = _ffi.lib()
        dev = block.device
        modl_bar = torch.empty((stage.cfg.NA, stage.cfg.W), dtype=torch.float64, device=dev)
        amp_bar = torch.empty(2, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _ffi.check(L.tsff_ats_bwd(C.byref(stage.cfg), block.data_ptr(), e_amps.data_ptr(), saved.data_ptr(),
                                  thry_bar.contiguous().data_ptr(), modl_bar.data_ptr(), amp_bar.data_ptr(),
                                  stage._ws.data_ptr(), st))
        block_bar = torch.zeros_like(block)
        block_bar[0, _ffi.P_AMP1] = amp_bar[0]
        block_bar[0, _ffi.P_AMP2] = amp_bar[1]
        return modl_bar, block_bar, None, None, None
