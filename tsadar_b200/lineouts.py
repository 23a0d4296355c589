"""Lineout extraction -- the data-side stage in front of the fit, mirror of the extraction part of
tsadar.utils.process.lineouts.get_lineouts (lineouts.py:85-165) on tsff_lineouts_fwd: the CCD image stays on the device and
the batches the fit consumes (`e_data` / `i_data`, `e_amps` / `i_amps`) are produced there."""
from __future__ import annotations

import numpy as np
import torch

from . import _ffi


def extract_lineouts(image, pixels, dpixel, gain, axis_y=None, windows=None):
    """image [NY, NX] float64 CUDA tensor (wavelength x time/space); pixels: lineout centres (ints); windows: list of (lo, hi)
    ranges on axis_y over which the amplitude is taken (the fit windows, lineouts.py:128-137, 145-152) or None = everywhere.
    -> (data [L, NY], amps [L]) CUDA tensors."""
    if not (isinstance(image, torch.Tensor) and image.is_cuda and image.dtype == torch.float64 and image.is_contiguous()):
        raise RuntimeError("image must be a contiguous float64 CUDA tensor: tsadar_b200 has no CPU path")
    NY, NX = image.shape
    px_h = np.ascontiguousarray(np.asarray(pixels, dtype=np.int32).reshape(-1))
    L = int(px_h.size)
    dev = image.device
    px_d = torch.as_tensor(px_h, device=dev)
    win = None
    if windows is not None:
        ay = np.asarray(axis_y, dtype=np.float64)
        m = np.zeros(NY, dtype=bool)
        for lo, hi in windows:
            m |= (lo < ay) & (ay < hi)
        win = torch.as_tensor(m.astype(np.uint8), device=dev)
    data = torch.empty((L, NY), dtype=torch.float64, device=dev)
    amps = torch.empty(L, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _ffi.check(_ffi.lib().tsff_lineouts_fwd(image.data_ptr(), NY, NX, px_d.data_ptr(), px_h.ctypes.data, L, int(dpixel), float(gain),
                                            win.data_ptr() if win is not None else None, data.data_ptr(), amps.data_ptr(), st))
    return data, amps
