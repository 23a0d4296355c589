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
    from tsadar_b200.fit import batch_to_device
    ts_diag = ThomsonScatteringDiagnostic(cfg, scattering_angles=SA_P9)
    cfg["parameters"] = _perturb_params_(rng, cfg["parameters"])
    gt = ThomsonParams(copy.deepcopy(cfg["parameters"]), num_params=1, batch=True, activate=True)
    dummy = batch_to_device(dummy_batch_1d())          # device-resident, so that the step can also be graph-captured
    with torch.no_grad():
        ThryE_gt, _, _, _ = ts_diag(gt, dummy)
    ThryE_gt = ThryE_gt.detach()
    cfg["parameters"] = _perturb_params_(rng, cfg["parameters"])
    fit = ThomsonParams(copy.deepcopy(cfg["parameters"]), num_params=1, batch=True, activate=True)

    def loss_fn(tp):
        ThryE, _, _, _ = ts_diag(tp, dummy)
        return torch.mean(torch.square(ThryE - ThryE_gt))

    return gt, fit, loss_fn


@pytest.mark.parametrize("graphed", [False, True])
def test_1d_inverse_lbfgsb_recovers_parameters(graphed):
    from tsadar_b200.fit import scipy_fit
    gt, fit, loss_fn = _setup()
    l0 = float(loss_fn(fit).detach())
    res = scipy_fit(loss_fn, fit, method="L-BFGS-B", cuda_graph=graphed)
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


def test_hessian_of_the_fit_loss_and_sigmas():
    """Second-order path (loss_function.py:110,170-188; postprocess.get_sigmas): the finite-difference-of-adjoint Hessian is
    symmetric, block-diagonal over lineouts, matches second differences of the loss itself, and is positive definite at
    the optimum of a noiseless synthetic fit (so every sigma is real and positive)."""
    from tsadar_b200.loss_function import LossFunction, get_sigmas
    from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
    from tsadar_b200.ts_params import ThomsonParams
    from tsadar_b200.fit import ravel_leaves, unravel_into
    cfg = load_cfg("cfg_1d")
    cfg["other"]["points_per_pixel"] = 1
    cfg["other"]["npts"] = 1024
    B = 2
    tp = ThomsonParams(copy.deepcopy(cfg["parameters"]), num_params=B, batch=True, activate=True)
    with torch.no_grad():
        tp.leaves[("electron", "Te")].value[1] += 0.2
    diag = ThomsonScatteringDiagnostic(cfg, scattering_angles=SA_P9)
    dummy = dict(i_data=np.ones((B, 1024)), e_data=np.ones((B, 1024)), noise_e=np.zeros((B, 1024)), noise_i=np.zeros((B, 1024)),
                 e_amps=np.ones(B), i_amps=np.ones(B))
    with torch.no_grad():
        data, _, _, _ = diag(tp, dummy)
    batch = dict(dummy, e_data=data.detach().cpu().numpy())           # noiseless data generated at the current parameters
    lf = LossFunction(cfg, SA_P9, batch)
    H, rows = lf.h_loss_wrt_params(tp, batch)
    n = H.shape[0]
    assert n == len(tp.parameters()) * B
    assert np.allclose(H, H.T)
    # different lineouts do not talk to each other
    for a, (la, ea) in enumerate(rows):
        for b, (lb, eb) in enumerate(rows):
            if ea != eb:
                assert abs(H[a, b]) <= 1e-6 * np.sqrt(abs(H[a, a] * H[b, b])), (a, b, H[a, b])
    # one diagonal entry against the second difference of the loss itself
    leaves = tp.parameters()
    x0 = ravel_leaves(leaves)
    k, h = 0, 2e-3
    vals = []
    for dx in (-h, 0.0, h):
        x = x0.copy(); x[k] += dx
        unravel_into(leaves, x)
        vals.append(float(lf.loss_for_hess(tp, batch).detach()))
    unravel_into(leaves, x0)
    d2 = (vals[0] - 2 * vals[1] + vals[2]) / h**2
    assert abs(d2 - H[k, k]) <= 2e-3 * abs(H[k, k]), (d2, H[k, k])
    assert np.linalg.eigvalsh(H).min() > 0
    sig = get_sigmas(H, rows, B)
    assert sig.shape == (B, len(leaves)) and np.all(sig > 0) and np.all(np.isfinite(sig))
