"""The four loss functionals of LossFunction.loss_functionals (loss_function.py:386-418: l1, l2, log-cosh, poisson) through
tsff_loss_fwd_bwd: value against the NumPy oracle's calc_ei_error (window masks + nanmean, loss_function.py:190-267), seed
cotangent against float64 autograd of the same formula; then end to end through LossFunction.vg_loss for every functional
(value vs oracle.loss_1d, gradient vs central differences of the same CUDA loss)."""
import copy

import numpy as np
import pytest
import torch

from oracle import np_oracle as O, params_oracle as P
from tests.common import SA_P9, load_cfg

pytestmark = pytest.mark.gpu

METHODS = ["l2", "l1", "log-cosh", "poisson"]


def _cfg(method):
    cfg = load_cfg("cfg_1d")
    cfg["optimizer"]["loss_method"] = method
    return cfg


def _window_weights(cfg, lam):
    fr = cfg["data"]["fit_rng"]
    mb = (lam > fr["blue_min"]) & (lam < fr["blue_max"])
    mr = (lam > fr["red_min"]) & (lam < fr["red_max"])
    return 0.5 * (mb / mb.sum() + mr / mr.sum())


@pytest.mark.parametrize("method", METHODS)
def test_loss_kernel_value_and_seed_vs_oracle(method):
    from tsadar_b200.engine import loss_fwd_bwd
    cfg = _cfg(method)
    rng = np.random.default_rng(3)
    B, n = 5, 1024
    lam = np.linspace(400.0, 700.0, n)
    t = 0.6 * np.exp(-0.5 * ((lam - 470) / 12.0) ** 2) + 0.5 * np.exp(-0.5 * ((lam - 590) / 15.0) ** 2) + 0.02
    t = t[None, :] * rng.uniform(0.7, 1.4, (B, 1))
    d = t * (1 + 0.2 * rng.normal(size=(B, n))) + 0.01
    t[:, :50] = -1.0        # outside every fit window: a poisson log of a negative theory must not leak into the loss
    uncert = 1.7
    w = _window_weights(cfg, lam)
    tt, dt, wt = (torch.tensor(a, device="cuda") for a in (t, d, w))
    loss, tbar = loss_fwd_bwd(tt, dt, wt, uncert, 1.0 / B, method)
    # oracle value: the reference's per-batch nanmean over the window of the whole [B, n] block (loss_function.py:371)
    batch = dict(e_data=d, i_data=d)
    ex = dict(cfg["other"]["extraoptions"])
    assert ex["fit_EPWb"] and ex["fit_EPWr"] and not ex["fit_IAW"]
    with np.errstate(invalid="ignore"):
        _, e_err = O.calc_ei_error(cfg, batch, t, lam, t, lam, [uncert, uncert])
    assert np.isfinite(e_err)
    assert abs(float(loss) - e_err) <= 1e-12 * abs(e_err), (float(loss), e_err)
    # seed cotangent: autograd of the functional in float64
    tq = torch.tensor(t, requires_grad=True)
    dq, wq = torch.tensor(d), torch.tensor(w)
    m = wq > 0
    diff = dq - tq
    if method == "l2":
        e = diff**2 / uncert
    elif method == "l1":
        e = diff.abs() / uncert
    elif method == "log-cosh":
        e = torch.log(torch.cosh(diff))
    else:
        e = tq - dq * torch.log(torch.where(m, tq, torch.ones_like(tq)))
    (torch.where(m, e, torch.zeros_like(e)) * wq).sum().div(B).backward()
    np.testing.assert_allclose(tbar.cpu().numpy(), tq.grad.numpy(), rtol=1e-10, atol=1e-300)   # the kernel divides, autograd multiplies by a reciprocal
    assert float(tbar[:, :50].abs().max()) == 0.0


@pytest.mark.parametrize("method", ["l1", "log-cosh", "poisson"])
def test_loss_function_end_to_end_other_functionals(method):
    """LossFunction.vg_loss with loss_method = l1 / log-cosh / poisson on a 2-lineout batch: the value against
    oracle.loss_1d (whole NumPy chain) and the gradient against central differences of the CUDA loss."""
    from tsadar_b200.loss_function import LossFunction
    from tsadar_b200.ts_params import ThomsonParams
    from tsadar_b200.fit import ravel_leaves, unravel_into, value_and_grad
    cfg = _cfg(method)
    cfg["other"]["points_per_pixel"] = 1
    cfg["other"]["npts"] = 1024
    cfg["parameters"]["electron"]["fe"]["nvx"] = 64
    B = 2
    lamb = np.linspace(400, 700, 1024)
    e_data = 0.6 * np.exp(-0.5 * ((lamb - 470) / 12.0) ** 2) + 0.5 * np.exp(-0.5 * ((lamb - 590) / 15.0) ** 2) + 0.01
    batch = dict(e_data=np.stack([e_data, 1.2 * e_data]), i_data=np.ones((B, 1024)), e_amps=np.array([1.0, 1.2]),
                 i_amps=np.ones(B), noise_e=np.zeros((B, 1024)), noise_i=np.zeros(1))
    loss_fn = LossFunction(cfg, SA_P9, batch)
    tp = ThomsonParams(cfg["parameters"], num_params=B, batch=True, activate=True)
    with torch.no_grad():
        tp.leaves[("electron", "Te")].value[1] += 0.3
        tp.leaves[("electron", "ne")].value[1] -= 0.2
    (loss, _), grads = loss_fn.vg_loss(tp, batch)
    # value: NumPy oracle of the whole chain on the physical parameters the mirror produced
    phys = tp()
    plist = []
    for b in range(B):
        q = {"electron": dict(Te=float(phys["electron"]["Te"][b]), ne=float(phys["electron"]["ne"][b]),
                              fe=phys["electron"]["fe"][b].detach().cpu().numpy(), v=tp.vx),
             "general": {k: float(v[b]) for k, v in phys["general"].items()}}
        for ion in tp.ions:
            q[ion] = {k: float(v[b]) for k, v in phys[ion].items()}
        plist.append(q)
    ref, _, _ = O.loss_1d(plist, cfg, SA_P9, batch, i_norm=loss_fn.i_norm, e_norm=loss_fn.e_norm)
    assert abs(float(loss) - ref) <= 1e-6 * abs(ref), (method, float(loss), ref)
    # gradient: the FP32-sweep path against the FP64 validation path of the same kernels (1e-4), and the latter against central
    # differences of its own loss (differences of the FP32 path would drown in its 1e-7 rounding noise)
    leaves = tp.parameters()
    x0 = ravel_leaves(leaves)
    g32 = np.concatenate([t.detach().cpu().numpy().ravel() for t in grads])
    loss_fn64 = LossFunction(cfg, SA_P9, batch, pv_precision="fp64")
    closure = lambda q: loss_fn64.calc_loss(q, batch)[0]
    _, g = value_and_grad(closure, tp)
    assert np.all(np.isfinite(g)) and np.abs(g).max() > 0
    assert np.all(np.abs(g32 - g) <= 1e-4 * np.maximum(np.abs(g), 1e-4 * np.abs(g).max())), (method, g32, g)
    for k in range(x0.size):
        h = 1e-5
        vals = []
        for dx in (-h, h):
            x = x0.copy()
            x[k] += dx
            unravel_into(leaves, x)
            with torch.no_grad():
                vals.append(float(closure(tp)))
        unravel_into(leaves, x0)
        fd = (vals[1] - vals[0]) / (2 * h)
        assert abs(g[k] - fd) <= 2e-4 * max(abs(fd), 1e-3 * np.abs(g).max()), (method, k, g[k], fd)
