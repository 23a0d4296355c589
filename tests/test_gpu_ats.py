"""GPU parity, ARTS rows of SURVEY.md 8 (a7 weight-matrix product, a8 add_ATS_IRF, a9 reduce_ATS_to_resunit):
tsff_ats_fwd / tsff_ats_bwd vs the NumPy / torch-f64 oracle, and the angular_full diagnostic end to end.
The reference's goldens for this path (ThryE-arts1v.npy) are missing blobs, so parity here is oracle-vs-kernel only
("parity unpinned" for the ATS stage itself; the form factor underneath is pinned by the 1-D golden)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O, torch_oracle as TO, params_oracle as P
from tests.common import load_cfg

pytestmark = pytest.mark.gpu


def _stage_cfg(NA, W, n_lam_data, ccd0, start, end, spect_fwhm, ang_fwhm):
    return {"other": {"PhysParams": {"norm": 0, "widIRF": {"spect_FWHM_ele": spect_fwhm, "ang_FWHM_ele": ang_fwhm}},
                      "CCDsize": [ccd0, n_lam_data]},
            "data": {"lineouts": {"start": start, "end": end}}}


@pytest.mark.parametrize("NA,W,n_lam_data,ccd0,start,end", [(128, 192, 64, 64, 5, 50), (96, 130, 65, 96, 0, 96), (64, 101, 50, 21, 1, 20)])
def test_ats_kernel_forward_and_vjp(NA, W, n_lam_data, ccd0, start, end):
    from tsadar_b200.ats import AtsStage
    from tsadar_b200 import _ffi
    rng = np.random.default_rng(NA + W)
    ang = np.sort(rng.uniform(20, 140, NA))          # non-uniform angle axis, like angsFRED
    lam = np.linspace(400.0, 700.0, W)
    a, l = np.meshgrid(ang, lam, indexing="ij")
    modl = (1.0 + 0.5 * np.sin(a / 9.0)) * (np.exp(-0.5 * ((l - 470 - 0.3 * a) / 9.0) ** 2) + 0.7 * np.exp(-0.5 * ((l - 600 + 0.2 * a) / 12.0) ** 2)) \
        + 0.01 * rng.uniform(size=(NA, W))
    cfg = _stage_cfg(NA, W, n_lam_data, ccd0, start, end, 4.0, 3.0)
    st = AtsStage(cfg, {"angAxis": ang}, (400.0, 700.0), W, n_lam_data)
    lamL, amp1, amp2 = 526.3, 0.8, 1.3
    block = torch.zeros((1, _ffi.P_ION0 + 4), dtype=torch.float64, device="cuda")
    block[0, _ffi.P_LAM], block[0, _ffi.P_AMP1], block[0, _ffi.P_AMP2] = lamL, amp1, amp2
    block.requires_grad_(True)
    e_amps = rng.uniform(0.5, 2.0, st.nrows)
    noise = 0.01 * rng.normal(size=(st.nrows, st.nl))
    mt = torch.tensor(modl, device="cuda", requires_grad=True)
    thry = st(mt, block, torch.tensor(e_amps, device="cuda"), torch.tensor(noise, device="cuda"))
    # oracle values
    _, y = O.add_ats_irf(lam, ang, modl, 4.0, 3.0)
    ref, lamb = O.reduce_ats_to_resunit(y, lam, lamL, amp1, amp2, e_amps[:, None], n_lam_data, ccd0, start, end)
    ref = ref + noise
    got = thry.detach().cpu().numpy()
    assert got.shape == ref.shape
    np.testing.assert_allclose(st.lam_units, lamb, rtol=1e-14)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12
    # VJP vs torch autograd of the oracle chain
    cot = rng.normal(size=ref.shape)
    (thry * torch.tensor(cot, device="cuda")).sum().backward()
    mo = torch.tensor(modl, requires_grad=True)
    a1, a2 = torch.tensor(amp1, dtype=torch.float64, requires_grad=True), torch.tensor(amp2, dtype=torch.float64, requires_grad=True)
    yo, _ = TO.ats_chain(mo, lam, ang, 4.0, 3.0, lamL, a1, a2, e_amps, n_lam_data, ccd0, start, end)
    (yo * torch.tensor(cot)).sum().backward()
    gm = mo.grad.numpy()
    assert np.abs(mt.grad.cpu().numpy() - gm).max() / np.abs(gm).max() < 1e-10
    gb = block.grad.cpu().numpy()[0]
    assert abs(gb[_ffi.P_AMP1] - a1.grad.item()) <= 1e-10 * abs(a1.grad.item())
    assert abs(gb[_ffi.P_AMP2] - a2.grad.item()) <= 1e-10 * abs(a2.grad.item())


def _arts_setup(npts=256):
    import os
    cfg = load_cfg("cfg_arts1v")
    cfg["other"]["lamrangE"] = [cfg["data"]["fit_rng"]["forward_epw_start"], cfg["data"]["fit_rng"]["forward_epw_end"]]
    cfg["other"]["lamrangI"] = [cfg["data"]["fit_rng"]["forward_iaw_start"], cfg["data"]["fit_rng"]["forward_iaw_end"]]
    cfg["other"]["npts"] = npts
    cfg["other"]["extraoptions"]["spectype"] = "angular_full"       # tests/test_forward/test_angular_1v.py:55
    tab = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tsadar_b200", "data", "arts_angles.npz"))
    sa = dict(sa=np.arange(19, 139.5, 0.5), weights=tab["weightMatrix"], angAxis=tab["angsFRED"])   # calibration.py:457-458, 487-491
    n_lam = npts // 2
    batch = dict(i_data=np.ones((1024, n_lam)), e_data=np.ones((1024, n_lam)), noise_e=np.array([0.0]), noise_i=np.array([0.0]),
                 e_amps=np.array([1.0]), i_amps=np.array([1.0]))
    return cfg, sa, batch


def test_arts1v_diagnostic_forward_matches_oracle():
    """test_arts1d_forward_pass (tests/test_forward/test_angular_1v.py:17-75) at a reduced npts: A = 241 angles,
    [1024, 241] weight matrix, ATS IRF, reduction to resolution units."""
    from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
    from tsadar_b200.ts_params import ThomsonParams
    cfg, sa, batch = _arts_setup(256)
    p = P.thomson_params(cfg["parameters"], activate=True)
    ref, lamb, modl_ref = O.diagnostic_arts(p, cfg, sa, batch)
    ts_diag = ThomsonScatteringDiagnostic(cfg, scattering_angles=sa)
    ts_params = ThomsonParams(cfg["parameters"], num_params=1, batch=False, activate=True)
    ThryE, ThryI, lamE, _ = ts_diag(ts_params, batch)
    got = ThryE.detach().cpu().numpy()
    assert got.shape == ref.shape == (cfg["data"]["lineouts"]["end"] - cfg["data"]["lineouts"]["start"], 128)
    np.testing.assert_allclose(lamE, lamb, rtol=1e-13)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5


def test_arts1v_diagnostic_full_deck_shape_matches_oracle():
    """The arts-1d deck at its OWN shape (BASELINE.json configs[2]; tests/test_forward/test_angular_1v.py): npts = 2048
    wavelengths x 241 angles = 493 568 (omega, angle) points, [1024, 241] weight matrix, ATS IRF on the [1024, 2048] image,
    reduction to resolution units -- the whole diagnostic against the NumPy oracle."""
    from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
    from tsadar_b200.ts_params import ThomsonParams
    cfg, sa, batch = _arts_setup(2048)
    p = P.thomson_params(cfg["parameters"], activate=True)
    ref, lamb, modl_ref = O.diagnostic_arts(p, cfg, sa, batch)
    ts_diag = ThomsonScatteringDiagnostic(cfg, scattering_angles=sa)
    ts_params = ThomsonParams(cfg["parameters"], num_params=1, batch=False, activate=True)
    ThryE, ThryI, lamE, _ = ts_diag(ts_params, batch)
    got = ThryE.detach().cpu().numpy()
    assert got.shape == ref.shape == (cfg["data"]["lineouts"]["end"] - cfg["data"]["lineouts"]["start"], 1024)
    np.testing.assert_allclose(lamE, lamb, rtol=1e-13)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"arts-1d full shape: max|diff|/max {err:.2e}")
    assert err < 1e-5


def test_arts1v_loss_gradient_matches_oracle():
    """d loss / d (Te, ne, amp1, amp2, lam) through weights-GEMM -> ATS IRF -> reduction, vs torch autograd of the oracle."""
    from tsadar_b200.generate_spectra import FitModel
    from tsadar_b200.ats import AtsStage
    from tsadar_b200 import _ffi
    cfg, sa, batch = _arts_setup(128)
    p = P.thomson_params(cfg["parameters"], activate=True)
    oth = cfg["other"]
    rng = np.random.default_rng(4)
    # GPU chain with the physical parameter block as the differentiable leaf
    fm = FitModel(cfg, sa)
    vx, fe = np.asarray(p["electron"]["v"]), np.asarray(p["electron"]["fe"])
    row = np.array([p["electron"]["Te"], p["electron"]["ne"], p["general"]["lam"], p["general"]["Va"], p["general"]["ud"],
                    p["general"]["ne_gradient"], p["general"]["Te_gradient"], p["general"]["amp1"], p["general"]["amp2"], p["general"]["amp3"],
                    p["ion-1"]["A"], p["ion-1"]["Z"], p["ion-1"]["Ti"], p["ion-1"]["fract"]], dtype=np.float64)
    from tsadar_b200.engine import form_factor_full
    block = torch.tensor(row[None], device="cuda", requires_grad=True)
    fet = torch.tensor(fe[None], device="cuda", requires_grad=True)
    eng = fm.electron_form_factor.engine(vx, 1)
    ff = form_factor_full(eng, block, fet)[0]
    from tsadar_b200.generate_spectra import arts_weights
    wm = torch.tensor(sa["weights"], device="cuda").contiguous()
    modlE = arts_weights(ff, wm, torch.tensor(fm._jmulE, device="cuda"))          # tsff_arts_weights_fwd / _bwd
    st = AtsStage(cfg, sa, oth["lamrangE"], oth["npts"], 64)
    thry = st(modlE, block, torch.ones(st.nrows, dtype=torch.float64, device="cuda"))
    cot = rng.normal(size=tuple(thry.shape))
    (thry * torch.tensor(cot, device="cuda")).sum().backward()
    # oracle chain
    grids = O.Grids(oth["lamrangE"], oth["npts"])
    leaves, pt = TO.params_from_block(row, 1)
    feo = torch.tensor(fe, requires_grad=True)
    ffo = TO.form_factor_1v(pt, feo, vx, grids, sa["sa"], 1, cfg["data"]["ele_lam_shift"])
    mo = torch.matmul(torch.tensor(sa["weights"]), ffo.mean(0).t()) * torch.tensor(fm._jmulE)
    yo, _ = TO.ats_chain(mo, grids.lam_axis, sa["angAxis"], oth["PhysParams"]["widIRF"]["spect_FWHM_ele"], oth["PhysParams"]["widIRF"]["ang_FWHM_ele"],
                         leaves[2], leaves[7], leaves[8], np.ones(st.nrows), 64, oth["CCDsize"][0], cfg["data"]["lineouts"]["start"], cfg["data"]["lineouts"]["end"])
    (yo * torch.tensor(cot)).sum().backward()
    gp = leaves.grad.numpy()
    gb = block.grad.cpu().numpy()[0]
    for k in (0, 1, 7, 8):
        assert abs(gb[k] - gp[k]) <= 1e-4 * abs(gp[k]), (k, gb[k], gp[k])
    gf = feo.grad.numpy()
    assert np.abs(fet.grad.cpu().numpy()[0] - gf).max() / np.abs(gf).max() < 1e-4


@pytest.mark.parametrize("G,W,A,NA", [(1, 2048, 241, 1024), (3, 301, 241, 1024), (2, 65, 17, 130)])
def test_arts_weight_contraction_kernel_vs_numpy(G, W, A, NA):
    """tsff_arts_weights_fwd / _bwd (generate_spectra.py:193-197, 210-216: mean over gradient points, weights @ ThryE.T, IAW
    filter) against the same arithmetic in NumPy float64 at 1e-12, forward and VJP, full arts-1d shape and ragged shapes."""
    from tsadar_b200.generate_spectra import arts_weights
    rng = np.random.default_rng(G + W)
    ff = rng.uniform(0.1, 2.0, (G, W, A)) * np.exp(rng.normal(size=(1, W, 1)))
    wm = rng.uniform(0.0, 1.0, (NA, A)) * (rng.uniform(size=(NA, A)) < 0.3)            # sparse-ish, like the FRED weight matrix
    jm = np.where(rng.uniform(size=W) < 0.1, 1e-4, 1.0)
    cot = rng.normal(size=(NA, W))
    fft = torch.tensor(ff, device="cuda", requires_grad=True)
    out = arts_weights(fft, torch.tensor(wm, device="cuda"), torch.tensor(jm, device="cuda"))
    ref = (wm @ ff.mean(0).T) * jm
    assert np.abs(out.detach().cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
    (out * torch.tensor(cot, device="cuda")).sum().backward()
    gref = np.broadcast_to(((cot * jm).T @ wm) / G, (G, W, A))
    assert np.abs(fft.grad.cpu().numpy() - gref).max() <= 1e-12 * np.abs(gref).max()
    out2 = arts_weights(fft.detach(), torch.tensor(wm, device="cuda"), None)               # no filter
    ref2 = wm @ ff.mean(0).T
    assert np.abs(out2.cpu().numpy() - ref2).max() <= 1e-12 * np.abs(ref2).max()
    assert torch.equal(out2, arts_weights(fft.detach(), torch.tensor(wm, device="cuda"), None))   # deterministic
    # the adjoint at this shape may run its k range in two halves added atomically: a two-term sum is order-independent
    f2 = torch.tensor(ff, device="cuda", requires_grad=True)
    (arts_weights(f2, torch.tensor(wm, device="cuda"), torch.tensor(jm, device="cuda")) * torch.tensor(cot, device="cuda")).sum().backward()
    assert torch.equal(f2.grad, fft.grad)


def test_angular_full_with_an_ion_spectrum():
    """postprocess_theory applies add_ion_IRF whatever the spectype (thomson_diagnostic.py:61-62): with load_ion_spec on, the
    ARTS diagnostic also returns the (un-batched) ion spectrum of its single parameter set."""
    from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
    from tsadar_b200.ts_params import ThomsonParams
    cfg, sa, batch = _arts_setup(1024)
    cfg["other"]["extraoptions"]["load_ion_spec"] = True
    p = P.thomson_params(cfg["parameters"], activate=True)
    ts_diag = ThomsonScatteringDiagnostic(cfg, scattering_angles=sa)
    tp = ThomsonParams(cfg["parameters"], num_params=1, batch=False, activate=True)
    ThryE, ThryI, lamE, lamI = ts_diag(tp, batch)
    gI = O.Grids(cfg["other"]["lamrangI"], 1024)
    lI, mI = O.fit_model_ion(p, gI, sa, cfg["parameters"]["general"]["Te_gradient"]["num_grad_points"])
    lI, tI = O.add_ion_irf(lI, mI, 1.0, float(p["general"]["amp3"]), cfg["other"]["PhysParams"]["widIRF"]["spect_stddev_ion"],
                           cfg["other"]["PhysParams"]["norm"])
    assert ThryI.shape == (1024,)
    assert np.abs(ThryI.detach().cpu().numpy() - tI).max() / np.abs(tI).max() < 1e-5
    np.testing.assert_allclose(np.asarray(lamI), lI, rtol=1e-12)
    assert isinstance(ThryE, torch.Tensor) and ThryE.shape[0] == cfg["data"]["lineouts"]["end"] - cfg["data"]["lineouts"]["start"]
