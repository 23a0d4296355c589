"""Replay of the reference's only gradient-dependent assertion that runs without a GPU in the reference,
tests/test_inverse/test_1d_random.py:57-174: a synthetic spectrum is generated from parameters drawn with
default_rng(42), the parameters are re-drawn, and L-BFGS-B (jac=True) must recover Te, ne, m, amp1, amp2, lam to
rtol 0.1.  Here the loss and its gradient come from the CUDA forward + hand-written adjoint kernels."""
import copy

import numpy as np
import pytest
import torch

from tests.common import SA_P9, load_cfg, dummy_batch_1d

pytestmark = pytest.mark.gpu


def _perturb_params_(rng, params):            # test_1d_random.py:22-47, same draw order
    params["electron"]["fe"]["params"]["m"]["val"] = float(rng.uniform(2.0, 3.5))
    params["electron"]["Te"]["val"] = float(rng.uniform(0.5, 1.5))
    params["electron"]["ne"]["val"] = float(rng.uniform(0.1, 0.7))
    params["general"]["amp1"]["val"] = float(rng.uniform(0.5, 2.5))
    params["general"]["amp2"]["val"] = float(rng.uniform(0.5, 2.5))
    params["general"]["lam"]["val"] = float(rng.uniform(523, 527))
    return params


def _flat(p):
    out = {}
    for k, v in p.items():
        for kk, vv in v.items():
            if isinstance(vv, torch.Tensor) and vv.numel() == 1:
                out[(k, kk)] = float(vv.detach().reshape(-1)[0])
    return out


def _setup():
    from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
    from tsadar_b200.ts_params import ThomsonParams
    cfg = load_cfg("cfg_1d")
    rng = np.random.default_rng(42)
    ts_diag = ThomsonScatteringDiagnostic(cfg, scattering_angles=SA_P9)
    cfg["parameters"] = _perturb_params_(rng, cfg["parameters"])
    gt = ThomsonParams(copy.deepcopy(cfg["parameters"]), num_params=1, batch=True, activate=True)
    with torch.no_grad():
        ThryE_gt, _, _, _ = ts_diag(gt, dummy_batch_1d())
    ThryE_gt = ThryE_gt.detach()
    cfg["parameters"] = _perturb_params_(rng, cfg["parameters"])
    fit = ThomsonParams(copy.deepcopy(cfg["parameters"]), num_params=1, batch=True, activate=True)

    def loss_fn(tp):
        ThryE, _, _, _ = ts_diag(tp, dummy_batch_1d())
        return torch.mean(torch.square(ThryE - ThryE_gt))

    return gt, fit, loss_fn


def test_1d_inverse_lbfgsb_recovers_parameters():
    from tsadar_b200.fit import scipy_fit
    gt, fit, loss_fn = _setup()
    l0 = float(loss_fn(fit).detach())
    res = scipy_fit(loss_fn, fit, method="L-BFGS-B")
    assert res["fun"] < 1e-4 * l0, (res["fun"], l0)
    g, l = _flat(gt.get_unnormed_params()), _flat(fit.get_unnormed_params())
    assert ("electron", "m") in g and ("electron", "Te") in g
    for key in g:
        np.testing.assert_allclose(l[key], g[key], atol=0, rtol=0.1, err_msg=str(key))   # the reference's assertion


def test_1d_inverse_adam_reduces_loss():
    """The optax branch of the same test (test_1d_random.py:121-130): adam(0.004), 100 steps."""
    from tsadar_b200.fit import adam_fit
    gt, fit, loss_fn = _setup()
    hist = adam_fit(loss_fn, fit, 0.004, 100)
    assert np.all(np.isfinite(hist))
    assert hist[-1] < 0.5 * hist[0], (hist[0], hist[-1])
