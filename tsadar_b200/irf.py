"""Instrument response -- mirror of tsadar.core.physics.irf (irf.py:50-132) on the CUDA kernels tsff_irf_fwd/bwd.
Batched over lineouts (the reference vmaps postprocess_theory, thomson_diagnostic.py:36)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ffi


def _cfg(W, nbins, lam_min, lam_max, stddev, kind, norm=0, cut_sigma=0.0):
    c = _ffi.IrfCfg()
    c.W, c.nbins, c.norm, c.kind = int(W), int(nbins), int(norm), int(kind)
    c.lam_min, c.lam_max, c.stddev, c.cut_sigma = float(lam_min), float(lam_max), float(stddev), float(cut_sigma)
    return c


class _IrfFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, modl, block, amps, noise, cfg):
        L = _ffi.lib()
        B, W = modl.shape
        dev = modl.device
        thry = torch.empty((B, cfg.nbins), dtype=torch.float64, device=dev)
        saved = torch.empty(int(L.tsff_irf_saved_bytes(C.byref(cfg), B)), dtype=torch.uint8, device=dev)
        ws = torch.empty(int(L.tsff_irf_workspace_bytes(C.byref(cfg), B)), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _ffi.check(L.tsff_irf_fwd(C.byref(cfg), B, modl.data_ptr(), block.data_ptr(), block.shape[1], amps.data_ptr(),
                                  noise.data_ptr() if noise is not None else None, thry.data_ptr(), saved.data_ptr(),
                                  ws.data_ptr(), st))
        ctx.cfg = cfg
        ctx.save_for_backward(block, amps, saved)
        return thry

    @staticmethod
    def backward(ctx, thry_bar):
        block, amps, saved = ctx.saved_tensors
        cfg = ctx.cfg
        L = _ffi.lib()
        B = block.shape[0]
        dev = block.device
        modl_bar = torch.empty((B, cfg.W), dtype=torch.float64, device=dev)
        amp_bar = torch.empty((B, 3), dtype=torch.float64, device=dev)
        ws = torch.empty(int(L.tsff_irf_workspace_bytes(C.byref(cfg), B)), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _ffi.check(L.tsff_irf_bwd(C.byref(cfg), B, block.data_ptr(), block.shape[1], amps.data_ptr(), saved.data_ptr(),
                                  thry_bar.contiguous().data_ptr(), modl_bar.data_ptr(), amp_bar.data_ptr(), ws.data_ptr(), st))
        block_bar = torch.zeros_like(block)
        block_bar[:, _ffi.P_AMP1:_ffi.P_AMP3 + 1] = amp_bar
        return modl_bar, block_bar, None, None, None


def binned_axis(lam_min, lam_max, W, nbins):
    """lamAxis.reshape(1024,-1).mean(1)  (irf.py:76,126)"""
    return np.linspace(lam_min, lam_max, W).reshape(nbins, -1).mean(axis=1)


def add_electron_IRF(config, lam_range, W, modlE, amps, block, noise=None, nbins=1024):
    """irf.add_electron_IRF (irf.py:90-132): modlE [B,W] -> (lamAxisE [nbins], ThryE [B,nbins])."""
    pp = config["other"]["PhysParams"]
    cfg = _cfg(W, nbins, lam_range[0], lam_range[1], pp["widIRF"]["spect_stddev_ele"], 0, pp["norm"])
    thry = _IrfFunction.apply(modlE.contiguous(), block, amps.contiguous(), noise, cfg)
    if pp["norm"] > 0:     # the reference bins the axis only in its norm == 0 branch (irf.py:125-126)
        return np.linspace(lam_range[0], lam_range[1], W), thry
    return binned_axis(lam_range[0], lam_range[1], W, nbins), thry


def add_ion_IRF(config, lam_range, W, modlI, amps, block, noise=None, nbins=1024):
    """irf.add_ion_IRF (irf.py:50-87)."""
    pp = config["other"]["PhysParams"]
    std = pp["widIRF"]["spect_stddev_ion"]
    if not std:
        return np.linspace(lam_range[0], lam_range[1], W), modlI if noise is None else modlI + noise
    cfg = _cfg(W, nbins, lam_range[0], lam_range[1], std, 1, pp["norm"])
    thry = _IrfFunction.apply(modlI.contiguous(), block, amps.contiguous(), noise, cfg)
    if pp["norm"] > 0:     # irf.py:76-78: axis binned and amp3 * amps / max applied only when norm == 0
        return np.linspace(lam_range[0], lam_range[1], W), thry
    return binned_axis(lam_range[0], lam_range[1], W, nbins), thry
