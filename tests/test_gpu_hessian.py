"""Second-order path (loss_function.py:110, 170-188): LossFunction.h_loss_wrt_params -- central
differences of the gradient the adjoint kernels return -- against the ORACLE's Hessian: torch autograd differentiated
twice (double backward) through the float64 restatement of the same chain
(ThomsonParams transforms -> DLM table lerp -> form factor (table mode) -> angle sum -> IAW filter -> IRF -> loss_for_hess)."""
import copy

import numpy as np
import pytest
import torch

from oracle import np_oracle as O, torch_oracle as TO
from tests.common import SA_P9, load_cfg

pytestmark = pytest.mark.gpu


def _oracle_loss_for_hess(x, tp, cfg, names, batch, grids, jmul):
    """loss_for_hess of ONE lineout as a function of the flat vector x of its active normalised leaves (float64 torch)."""
    w0 = float(SA_P9["weights"][0])
    val = {k: tp.leaves[k].value.detach().cpu()[0] for k in tp.leaves}
    for i, k in enumerate(names):
        val[k] = x[i]

    def phys(key):
        s = tp.leaves[key]
        return (torch.sigmoid(val[key]) if s.active else val[key]) * s.scale + s.shift

    m_ax, tab = tp.m_ax.cpu(), tp.f_vx_m.cpu()
    m = phys(("electron", "m"))
    i = int(torch.clamp(torch.searchsorted(m_ax, m.detach().reshape(1), right=True), 1, m_ax.numel() - 1))
    w = (m - m_ax[i - 1]) / (m_ax[i] - m_ax[i - 1])
    f = tab[:, i - 1] * (1 - w) + tab[:, i] * w
    fe = f / f.sum() / tp.dv
    p = dict(Te=phys(("electron", "Te")), ne=phys(("electron", "ne")), lam=phys(("general", "lam")), Va=phys(("general", "Va")),
             ud=phys(("general", "ud")), ne_gradient=phys(("general", "ne_gradient")), Te_gradient=phys(("general", "Te_gradient")),
             ions=[dict(A=torch.tensor(40.0, dtype=torch.float64), Z=phys(("ion-1", "Z")), Ti=phys(("ion-1", "Ti")),
                        fract=phys(("ion-1", "fract")) / phys(("ion-1", "fract")))])
    ff = TO.form_factor_1v(p, fe, tp.vx, grids, SA_P9["sa"], 1, 0.0)
    modl = TO.modl_from_ff(ff, np.full(10, w0), jmul)
    lb, thry = TO.add_electron_irf(grids.lam_axis, modl, batch["e_amps"][0], phys(("general", "lam")), phys(("general", "amp1")),
                                   phys(("general", "amp2")), cfg["other"]["PhysParams"]["widIRF"]["spect_stddev_ele"])
    fr = cfg["data"]["fit_rng"]
    d = torch.tensor(batch["e_data"][0])
    mask = ((lb > fr["blue_min"]) & (lb < fr["blue_max"])) | ((lb > fr["red_min"]) & (lb < fr["red_max"]))
    e = (d - thry) ** 2 / (d.abs() + 1e-10)
    return 0.5 * e[torch.as_tensor(mask)].sum()           # both EPW windows fitted: the halves are averaged (loss_function.py:262-264)


@pytest.mark.parametrize("pv,tol", [("fp64", 1e-3), ("fp32", 2e-2)])
def test_hessian_matches_the_oracles_double_backward(pv, tol):
    from tsadar_b200.loss_function import LossFunction
    from tsadar_b200.ts_params import ThomsonParams
    cfg = load_cfg("cfg_1d")
    cfg["other"]["points_per_pixel"] = 1
    cfg["other"]["npts"] = 1024
    cfg["parameters"]["electron"]["fe"]["nvx"] = 64
    B = 1
    lamb = np.linspace(400, 700, 1024)
    e_data = 0.6 * np.exp(-0.5 * ((lamb - 470) / 12.0) ** 2) + 0.5 * np.exp(-0.5 * ((lamb - 590) / 15.0) ** 2) + 0.01
    batch = dict(e_data=e_data[None], i_data=np.ones((B, 1024)), e_amps=np.array([1.0]), i_amps=np.ones(B),
                 noise_e=np.zeros((B, 1024)), noise_i=np.zeros((B, 1024)))
    lf = LossFunction(cfg, SA_P9, batch, pv_precision=pv)
    tp = ThomsonParams(copy.deepcopy(cfg["parameters"]), num_params=B, batch=True, activate=True)
    H, rows = lf.h_loss_wrt_params(tp, batch)
    names = [k for k, s in tp.leaves.items() if s.active]
    assert H.shape == (len(names), len(names))

    grids = O.Grids(cfg["other"]["lamrangE"], 1024)
    jmul = np.where((528 - 12 < grids.lam_axis) & (528 + 12 > grids.lam_axis), 1e-4, 1.0)
    x0 = torch.tensor([float(tp.leaves[k].value.detach().cpu()[0]) for k in names], dtype=torch.float64, requires_grad=True)
    L = _oracle_loss_for_hess(x0, tp, cfg, names, batch, grids, jmul)
    with torch.no_grad():
        got_L = float(lf.loss_for_hess(tp, batch))
    assert abs(got_L - float(L)) <= 1e-5 * abs(float(L)), (got_L, float(L))
    (g,) = torch.autograd.grad(L, x0, create_graph=True)
    Href = np.stack([torch.autograd.grad(g[i], x0, retain_graph=True)[0].numpy() for i in range(len(names))])
    assert np.allclose(Href, Href.T, rtol=1e-8, atol=1e-10 * np.abs(Href).max())
    scale = np.sqrt(np.abs(np.outer(np.diag(Href), np.diag(Href))))      # entrywise natural size
    err = np.abs(H - Href) / np.maximum(scale, 1e-300)
    assert err.max() <= tol, (pv, err.max(), names)
    # without the frozen cells the differences see the kinks of the linear interpolations: a different (smoother-function)
    # Hessian, which is NOT what the reference's jax.hessian returns
    if pv == "fp64":
        H2, _ = lf.h_loss_wrt_params(tp, batch, frozen_cells=False)
        assert (np.abs(H2 - Href) / np.maximum(scale, 1e-300)).max() > 10 * tol
    # and the sigmas derived from it (postprocess.get_sigmas)
    from tsadar_b200.loss_function import get_sigmas
    sig, sig_ref = get_sigmas(H, rows, B), get_sigmas(Href, rows, B)
    np.testing.assert_allclose(sig, sig_ref, rtol=50 * tol)
