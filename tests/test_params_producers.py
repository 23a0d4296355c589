"""f(v) producers: the torch mirrors in tsadar_b200/distribution_functions.py against the NumPy
restatement of the reference (oracle/params_oracle.py), values and gradients, on the CPU (they are host-side glue in
torch; the same code runs on the GPU in front of the kernels).  Analytic pins of the restatement itself: the l=1
harmonics are dipoles along vx and vy, tables are normalised, the folded filters equal the reference's recurrences."""
import copy

import numpy as np
import pytest
import torch

from oracle import params_oracle as P
from tests.common import load_cfg
from tsadar_b200 import distribution_functions as D
from tsadar_b200.ts_params import ThomsonParams


def test_butterworth_matrix_equals_the_recurrence():
    rng = np.random.default_rng(0)
    for n in (16, 64, 257):
        x = rng.normal(size=n)
        ref = P.second_order_butterworth(x, 100, 6, "forward_backward")
        np.testing.assert_allclose(D.butterworth_matrix(n) @ x, ref, rtol=0, atol=1e-12 * np.abs(ref).max())


def test_hann_matrix_equals_convolve_same():
    rng = np.random.default_rng(1)
    x = rng.normal(size=64)
    np.testing.assert_allclose(D.hann_same_matrix(64, 16) @ x, P.smooth1d(x, 16), rtol=0, atol=1e-14)


def _cfg_1v_arbitrary(nvx=96):
    cfg = load_cfg("cfg_1d")["parameters"]
    cfg["electron"]["fe"] = {"active": True, "dim": 1, "nvx": nvx, "type": "arbitrary", "params": {"init_m": 2.6}}
    return cfg


def test_arbitrary1v_matches_oracle_and_gradient():
    cfg = _cfg_1v_arbitrary()
    tp = ThomsonParams(cfg, num_params=3, batch=True, activate=True, device="cpu")
    vx, fval = P.arbitrary1v_init(96, 2.6)
    ref = P.arbitrary1v_call(vx, fval)
    got = tp()["electron"]["fe"]
    assert got.shape == (3, 96)
    np.testing.assert_allclose(got.detach().numpy()[1], ref, rtol=1e-11)
    assert abs(ref.sum() * (vx[1] - vx[0]) - 1) < 1e-13
    assert any(t is tp.dist.fval for t in tp.parameters())          # the table is a trainable leaf
    # gradient of a linear functional vs central differences of the oracle
    c = np.random.default_rng(2).normal(size=96)
    (got[1] * torch.tensor(c)).sum().backward()
    g = tp.dist.fval.grad[1].numpy()
    for k in (3, 40, 90):
        e = np.zeros(96); e[k] = 1e-6
        fd = (c @ P.arbitrary1v_call(vx, fval + e) - c @ P.arbitrary1v_call(vx, fval - e)) / 2e-6
        assert abs(g[k] - fd) <= 1e-6 * max(abs(fd), np.abs(g).max() * 1e-3), (k, g[k], fd)


def _cfg_2v(flm_type="mora-yahi", nvx=48, nvr=32):
    cfg = load_cfg("cfg_arts2v")["parameters"]
    cfg["electron"]["fe"]["nvx"] = nvx
    cfg["electron"]["fe"]["params"]["nvr"] = nvr
    cfg["electron"]["fe"]["params"]["flm_type"] = flm_type
    return cfg


@pytest.mark.parametrize("flm_type", ["mora-yahi", "arbitrary"])
def test_spherical_harmonics_matches_oracle(flm_type):
    cfg = _cfg_2v(flm_type)
    tp = ThomsonParams(cfg, num_params=1, batch=False, device="cpu")
    out = tp()
    got = out["electron"]["fe"].detach().numpy()
    vx, ref = P.spherical_harmonics_fe(cfg["electron"]["fe"])
    assert got.shape == ref.shape == (48, 48)
    np.testing.assert_allclose(out["electron"]["v"], vx, rtol=0, atol=0)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-300)
    assert abs(ref.sum() * (vx[1] - vx[0]) ** 2 - 1) < 1e-12
    with pytest.raises(NotImplementedError):
        ThomsonParams(cfg, num_params=1, batch=True, device="cpu")      # ts_params.py:156-159


def test_spherical_harmonics_l1_are_dipoles_along_vx_and_vy():
    """Re Y_1^0 ∝ vx/|v| and Re Y_1^1 ∝ -vy/|v| on the reference's (th, phi) convention: a positive f_10 shifts the
    first moment along +x (axis 1 of the "xy" mesh), f_11 along -y."""
    cfg = _cfg_2v("arbitrary")
    tp = ThomsonParams(cfg, num_params=1, batch=False, device="cpu")
    sh = tp.dist
    vx = sh.vx
    X, Y = np.meshgrid(vx, vx)
    r = np.sqrt(X**2 + Y**2)
    np.testing.assert_allclose(sh._ylm[(1, 0)].numpy().reshape(48, 48), np.sqrt(3 / (4 * np.pi)) * X / r, atol=1e-14)
    np.testing.assert_allclose(sh._ylm[(1, 1)].numpy().reshape(48, 48), -np.sqrt(3 / (8 * np.pi)) * Y / r, atol=1e-14)


def test_spherical_harmonics_leaf_gradients_vs_oracle_fd():
    cfg = _cfg_2v("mora-yahi")
    cfg["electron"]["fe"]["params"]["LTx"] = 40.0     # strong enough anisotropy for a well-conditioned derivative
    cfg["electron"]["fe"]["params"]["LTy"] = 70.0
    tp = ThomsonParams(cfg, num_params=1, batch=False, device="cpu")
    names = list(tp.dist.leaves().keys())
    assert names == ["normed_m", "flm[1][0].log_10_LT", "flm[1][1].log_10_LT"]
    assert len(tp.parameters()) == 3        # Te, ne inactive in the deck; fe active
    c = np.random.default_rng(3).normal(size=(48, 48))
    (tp()["electron"]["fe"] * torch.tensor(c)).sum().backward()
    fe_cfg = cfg["electron"]["fe"]
    nm0 = float(tp.dist.normed_m.detach())

    def fun(nm, ltx, lty):
        return float((P.spherical_harmonics_fe(fe_cfg, nm, {(1, 0): {"log_10_LT": ltx}, (1, 1): {"log_10_LT": lty}})[1] * c).sum())

    x0 = np.array([nm0, np.log10(40.0), np.log10(70.0)])
    for k, leaf in enumerate(tp.dist.leaves().values()):
        e = np.zeros(3); e[k] = 1e-5
        fd = (fun(*(x0 + e)) - fun(*(x0 - e))) / 2e-5
        assert abs(float(leaf.grad) - fd) <= 1e-6 * abs(fd) + 1e-12, (k, float(leaf.grad), fd)


@pytest.mark.parametrize("learn_log", [True, False])
def test_arbitrary2v_matches_oracle(learn_log):
    cfg = _cfg_2v()
    cfg["electron"]["fe"] = {"active": True, "dim": 2, "nvx": 40, "type": "arbitrary", "params": {"init_m": 2.3, "learn_log": learn_log}}
    tp = ThomsonParams(cfg, num_params=1, batch=False, device="cpu")
    got = tp()["electron"]["fe"]
    vx, fval = P.arbitrary2v_init(40, 2.3, learn_log)
    np.testing.assert_allclose(tp.dist.fval.detach().numpy(), fval, rtol=1e-14)
    np.testing.assert_allclose(got.detach().numpy(), P.arbitrary2v_call(vx, fval, learn_log), rtol=1e-12)
    assert tp.parameters()[-1] is tp.dist.fval


def test_thomson_params_oracle_2v_dispatch():
    cfg = _cfg_2v()
    p = P.thomson_params(copy.deepcopy(cfg), activate=False)
    assert p["electron"]["fe"].shape == (48, 48) and p["electron"]["v"].shape == (48,)


def test_get_unnormed_and_fitted_params():
    """ts_params.py:565-581, 605-645: the reported parameters exclude the f table; fitted = the active ones (+ m when the
    distribution is active, + flm with the assembled table for spherical harmonics)."""
    cfg = load_cfg("cfg_1d")["parameters"]
    tp = ThomsonParams(cfg, num_params=2, batch=True, activate=True, device="cpu")
    un = tp.get_unnormed_params()
    assert set(un["electron"]) == {"Te", "ne", "m"} and "fe" not in un["electron"]
    fitted, n = tp.get_fitted_params(cfg)
    active = [(k, k2) for k in cfg for k2 in cfg[k] if k2 != "fe" and cfg[k][k2].get("active")]
    assert n == len(active) + 1                                   # + m
    assert set(fitted["electron"]) == {"Te", "ne", "m"} and set(fitted["general"]) == {"amp1", "amp2", "lam"}
    assert fitted["ion-1"] == {}
    cfg2 = _cfg_2v("mora-yahi")
    tp2 = ThomsonParams(cfg2, num_params=1, batch=False, device="cpu")
    fitted2, n2 = tp2.get_fitted_params(cfg2)
    assert n2 == 0 and set(fitted2["electron"]) == {"flm"}
    assert fitted2["electron"]["flm"]["fvxvy"].shape == (48, 48) and set(fitted2["electron"]["flm"]) >= {0, 1, "fvxvy", "v"}


def test_flm_nn_radial_model_matches_oracle():
    """SphericalHarmonics with flm_type 'nn' (FLM_NN, spherical_harmonics.py:14-49: two eqx.nn.MLP(1, 1, 32, 3)) on the same
    weights as the NumPy restatement; and the table is differentiable with respect to every MLP weight."""
    import copy
    import torch
    from oracle import params_oracle as P
    from tests.common import load_cfg
    from tsadar_b200.distribution_functions import SphericalHarmonics
    fe = copy.deepcopy(load_cfg("cfg_arts2v")["parameters"]["electron"]["fe"])
    fe["nvx"], fe["params"]["nvr"], fe["params"]["flm_type"], fe["params"]["Nl"] = 24, 16, "nn", 1
    sh = SphericalHarmonics(fe, torch.device("cpu"), True)
    f = sh()
    leaves = {}
    for key, net in sh.flm.items():
        w = {n: [(W.detach().numpy(), b.detach().numpy()) for W, b in ls] for n, ls in net.nets.items()}
        leaves[key] = {"mag_weights": w["mag"], "sign_weights": w["sign"]}
    vx, ref = P.spherical_harmonics_fe(fe, flm_leaves=leaves)
    assert np.abs(f.detach().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.abs(ref - ref.T).max() > 1e-6 * np.abs(ref).max()        # the l = 1 terms make the table anisotropic
    (f * torch.linspace(0, 1, f.numel(), dtype=torch.float64).reshape(f.shape)).sum().backward()
    g = [t.grad for t in sh.leaves().values()]
    assert len(g) == 1 + 2 * 2 * 8 and all(x is not None and torch.isfinite(x).all() for x in g)
    # weights exported from elsewhere (e.g. a JAX run) are taken as they are
    fe2 = copy.deepcopy(fe)
    fe2["params"]["nn_weights"] = {key: {n: [(W.detach().numpy(), b.detach().numpy()) for W, b in ls] for n, ls in net.nets.items()}
                                   for key, net in sh.flm.items()}
    assert torch.equal(SphericalHarmonics(fe2, torch.device("cpu"), False)(), f.detach())
