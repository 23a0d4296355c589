"""Multiplexed shots (tsadar/inverse/loss_function.py:101, 287-317; configs/arts-2d/inputs.yaml lists two
shots): the second shot sees the electron distribution rotated by data.shot_rot (vector_tools.rotate, vector_tools.py:94-138),
the loss is the sum of both shots' errors.  Single GPU here (both shots on one rank); the split over ranks is covered by the
gloo tests (tests/test_parallel_gloo.py::test_shot_sharding_recipe)."""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O, params_oracle as P
from tests.common import load_cfg

pytestmark = pytest.mark.gpu


def _deck(npts=32):
    cfg = load_cfg("cfg_arts2v")
    cfg["other"]["lamrangE"] = [cfg["data"]["fit_rng"]["forward_epw_start"], cfg["data"]["fit_rng"]["forward_epw_end"]]
    cfg["other"]["lamrangI"] = [cfg["data"]["fit_rng"]["forward_iaw_start"], cfg["data"]["fit_rng"]["forward_iaw_end"]]
    cfg["other"]["npts"] = npts
    cfg["other"]["extraoptions"]["spectype"] = "angular_full"
    cfg["parameters"]["electron"]["fe"]["nvx"] = 24
    cfg["parameters"]["electron"]["fe"]["params"]["nvr"] = 16
    cfg["parameters"]["electron"]["fe"]["params"]["LTx"] = 60.0
    cfg["parameters"]["electron"]["fe"]["params"]["LTy"] = 90.0
    cfg["parameters"]["electron"]["fe"]["active"] = True
    cfg["data"]["shotnum"] = [101675, 101676]            # a list = multiplexed (loss_function.py:101)
    cfg["data"]["shot_rot"] = 33.0
    tab = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tsadar_b200", "data", "arts_angles.npz"))
    sa = dict(sa=np.arange(19, 139.5, 0.5), weights=tab["weightMatrix"], angAxis=tab["angsFRED"])
    return cfg, sa


def test_multiplexed_loss_matches_oracle_sum_of_two_shots():
    from tsadar_b200.loss_function import LossFunction
    from tsadar_b200.ts_params import ThomsonParams
    cfg, sa = _deck()
    n_lam = cfg["other"]["npts"] // 2
    rows = cfg["data"]["lineouts"]["end"] - cfg["data"]["lineouts"]["start"]
    rng = np.random.default_rng(5)
    mk = lambda s: dict(i_data=np.ones((1024, n_lam)), e_data=rng.uniform(0.2, 1.0, (rows, n_lam)) * s, noise_e=np.array([0.0]),
                        noise_i=np.array([0.0]), e_amps=np.array([1.0]), i_amps=np.array([1.0]))
    batch = {"b1": mk(1.0), "b2": mk(0.8)}
    lf = LossFunction(cfg, sa, batch["b1"])
    assert lf.multiplex_ang and lf.shot is None
    tp = ThomsonParams(cfg["parameters"], num_params=1, batch=False, activate=True)
    (loss, _), grads = lf.vg_loss(tp, batch)
    assert all(g is not None and torch.isfinite(g).all() for g in grads) and any(float(g.abs().sum()) > 0 for g in grads)
    # oracle: the same two diagnostics, the second with the rotated table
    p = P.thomson_params(cfg["parameters"], activate=True)
    unc = [lf.i_norm**2, lf.e_norm**2]
    tot = 0.0
    for key, rot in (("b1", None), ("b2", cfg["data"]["shot_rot"] * np.pi / 180.0)):
        q = {k: (dict(v) if isinstance(v, dict) else v) for k, v in p.items()}
        if rot is not None:
            q["electron"]["fe"] = O.rotate_pixels(np.squeeze(p["electron"]["fe"]), rot)
        ThryE, lamb, _ = O.diagnostic_arts(q, cfg, sa, batch[key])
        _, e_err = O.calc_ei_error(cfg, batch[key], 0.0, np.zeros(1), ThryE, np.asarray(lamb), unc)
        tot += e_err
    assert abs(float(loss) - tot) <= 1e-7 * abs(tot), (float(loss), tot)
    # and it is not the un-rotated sum (the rotation matters for this anisotropic table)
    q = {k: (dict(v) if isinstance(v, dict) else v) for k, v in p.items()}
    ThryE0, lamb, _ = O.diagnostic_arts(q, cfg, sa, batch["b2"])
    _, e0 = O.calc_ei_error(cfg, batch["b2"], 0.0, np.zeros(1), ThryE0, np.asarray(lamb), unc)
    ThryE1, _, _ = O.diagnostic_arts(q, cfg, sa, batch["b1"])
    _, e1 = O.calc_ei_error(cfg, batch["b1"], 0.0, np.zeros(1), ThryE1, np.asarray(lamb), unc)
    assert abs((e0 + e1) - tot) > 1e-6 * abs(tot)
