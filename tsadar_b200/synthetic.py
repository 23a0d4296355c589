"""Synthetic lineouts of the benchmark shape: host-side input generation only.

Per lineout: W = 1024 wavelengths on [400, 700] nm, one scattering angle (60 deg), one ion species
(A=40, Z=8, Ti=0.2 keV), V = 4096 velocity nodes; parameters drawn with default_rng(seed) from the ranges of the
reference's tests/test_inverse/test_1d_random.py:33-39; f(v) = projected super-Gaussian of order m."""
from __future__ import annotations

import numpy as np
from scipy.special import gamma, gammaincc

from . import _ffi

LAM_RANGE = (400.0, 700.0)
W_SYN, V_SYN = 1024, 4096
SA_SYN = np.array([60.0])


def vgrid(nvx):
    vmax = 6.0
    dv = 2 * vmax / nvx
    return np.linspace(-vmax + dv / 2, vmax - dv / 2, nvx)


def super_gaussian_projected(vx, m):
    alpha = np.sqrt(3.0 * gamma(3.0 / m) / 2.0 / gamma(5.0 / m))
    f = gamma(2.0 / m) * gammaincc(2.0 / m, (np.abs(vx) / (alpha * np.sqrt(2.0))) ** m)
    return f / np.sum(f) / (vx[1] - vx[0])


def make_lineouts(n, seed=42, nvx=V_SYN, n_unique=None, dtype=np.float32):
    """Returns params [n, NP] float64 and fe [n, nvx] (dtype).  To keep host-side generation cheap for very
    large n, only `n_unique` distinct f-tables are evaluated (default min(n, 256)) and tiled; the parameter rows
    are all distinct."""
    rng = np.random.default_rng(seed)
    NP = _ffi.P_ION0 + _ffi.ION_STRIDE
    p = np.zeros((n, NP))
    p[:, _ffi.P_TE] = rng.uniform(0.5, 1.5, n)
    p[:, _ffi.P_NE] = rng.uniform(0.1, 0.7, n)
    p[:, _ffi.P_LAM] = rng.uniform(523.0, 527.0, n)
    p[:, _ffi.P_AMP1] = rng.uniform(0.5, 2.5, n)
    p[:, _ffi.P_AMP2] = rng.uniform(0.5, 2.5, n)
    p[:, _ffi.P_AMP3] = 1.0
    p[:, _ffi.P_ION0 + _ffi.ION_A] = 40.0
    p[:, _ffi.P_ION0 + _ffi.ION_Z] = 8.0
    p[:, _ffi.P_ION0 + _ffi.ION_TI] = 0.2
    p[:, _ffi.P_ION0 + _ffi.ION_FRACT] = 1.0
    m = rng.uniform(2.0, 3.5, n)
    vx = vgrid(nvx)
    nu = min(n, 256) if n_unique is None else min(n, n_unique)
    tabs = np.stack([super_gaussian_projected(vx, m[i]) for i in range(nu)]).astype(dtype)
    fe = np.empty((n, nvx), dtype=dtype)
    for s in range(0, n, nu):
        e = min(n, s + nu)
        fe[s:e] = tabs[: e - s]
    return p, fe, vx, m
