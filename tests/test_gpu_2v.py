"""GPU parity, 2V path: FormFactor.calc_in_2D forward vs the NumPy oracle.
PARITY UNPINNED: the reference golden ThryE-arts2v.npy is a missing blob and interpax's bicubic interp2d is restated
from its published algorithm; substitute pins: (i) an isotropic f through the 2V path equals the 1V direct
path fed with the projected f up to discretisation, (ii) oracle-vs-kernel agreement."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O, torch_oracle as TO

pytestmark = pytest.mark.gpu


def _params(V, DF, vx, Va=0.0, ud=0.0, ne_grad=0.0):
    return {"electron": dict(Te=0.8, ne=0.3, fe=DF, v=vx),
            "general": dict(lam=526.5, amp1=1.0, amp2=1.0, amp3=1.0, ne_gradient=ne_grad, Te_gradient=0.0, ud=ud, Va=Va),
            "ion-1": dict(A=40.0, Z=8.0, Ti=0.2, fract=1.0)}


def _grid(V):
    dv = 12.0 / V
    return np.linspace(-6 + dv / 2, 6 - dv / 2, V)


@pytest.mark.parametrize("V,W,sa,ud,Va,ud_ang,va_ang,G", [(32, 24, [40.0, 75.0, 120.0], 0.0, 0.0, 0.0, 0.0, 1),
                                                           (48, 16, [60.0, 100.0], 0.4, -0.8, 30.0, 110.0, 2),
                                                           (128, 5, [60.0, 130.0], 0.2, 0.3, 200.0, 45.0, 1),    # arts-2d table size
                                                           # mirrored detector: beta near 225 deg, where the forward kernel
                                                           # staggers the threads' walk over a (bank conflicts) -- both signs
                                                           (64, 12, [-35.0, -90.0, -150.0], 0.0, 0.0, 0.0, 0.0, 1),
                                                           (64, 12, [-20.0, 170.0, 210.0, 300.0], 0.3, 0.5, 75.0, 290.0, 1)])
def test_calc_in_2D_matches_oracle(V, W, sa, ud, Va, ud_ang, va_ang, G):
    from tsadar_b200.form_factor import FormFactor
    vx = _grid(V)
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    # anisotropic, drifting, non-Maxwellian table (normalised like Arbitrary2V: f / sum / dv^2)
    DF = np.exp(-0.5 * ((X - 0.3) ** 2 / 1.2 + (Y + 0.2) ** 2 / 0.8) ** 1.2) * (1 + 0.2 * np.tanh(X))
    DF = DF / DF.sum() / (vx[1] - vx[0]) ** 2
    sa = np.array(sa)
    p = _params(V, DF, vx, Va=Va, ud=ud, ne_grad=4.0 if G > 1 else 0.0)
    grids = O.Grids([400.0, 700.0], W)
    ref, lams = O.form_factor_2d(p, grids, sa, G, 0.0, ud_ang, va_ang)
    ff = FormFactor([400.0, 700.0], W, 0.0, {"sa": sa}, G, ud_ang, va_ang)
    pt = {k: dict(v) for k, v in p.items()}
    got, lam_t = ff.calc_in_2D(pt)
    got = got.cpu().numpy()
    assert got.shape == ref.shape == (G, W, len(sa))
    err = np.abs(got - ref) / np.abs(ref).max()
    assert err.max() < 1e-9, err.max()                       # all-FP64 kernel vs all-FP64 oracle
    m = np.abs(ref) > 1e-6 * np.abs(ref).max()
    assert (np.abs(got - ref)[m] / np.abs(ref)[m]).max() < 1e-7
    np.testing.assert_allclose(lam_t.cpu().numpy(), lams, rtol=1e-14)


def test_isotropic_2v_equals_1v_direct():
    """Substitute pin (i): for an isotropic Maxwellian the rotate/project stage returns the 1-D Maxwellian for every beta,
    so calc_in_2D must agree with the direct 1V path (same calc_chi_vals arithmetic) up to the bicubic discretisation."""
    from tsadar_b200.form_factor import FormFactor
    from tsadar_b200.engine import FormFactorEngine
    V, W = 64, 64
    vx = _grid(V)
    dv = vx[1] - vx[0]
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    DF = np.exp(-0.5 * (X**2 + Y**2)) / (2 * np.pi)
    sa = np.array([60.0])
    p = _params(V, DF, vx)
    ff2 = FormFactor([400.0, 700.0], W, 0.0, {"sa": sa}, 1, 0.0, 0.0)
    got2, _ = ff2.calc_in_2D({k: dict(v) for k, v in p.items()})
    f1 = DF.sum(axis=0) * dv
    eng = FormFactorEngine((400.0, 700.0), W, 0.0, sa, np.ones(1), 1, 1, vx, mode="direct", pv_precision="fp64")
    row = np.array([[0.8, 0.3, 526.5, 0, 0, 0, 0, 1, 1, 1, 40.0, 8.0, 0.2, 1.0]])
    _, ff1, _ = eng.forward(torch.tensor(row, device="cuda"), torch.tensor(f1[None], device="cuda"), want_ff=True)
    a, b = got2.cpu().numpy()[0, :, 0], ff1.cpu().numpy()[0, 0, :, 0]
    assert np.abs(a - b).max() / np.abs(b).max() < 2e-3      # bicubic rotation of a V = 64 Maxwellian: ~1e-4 in f1


def test_calc_in_2D_vjp_matches_autograd():
    """Adjoint of the 2V path (fe_bar [V,V] through the bicubic rotate/project scatter, params_bar through |xi|, beta and the
    kinematics) vs torch autograd of the oracle, with drift, flow and two gradient points."""
    from oracle import torch_oracle as TO
    from tsadar_b200.engine import FormFactorEngine
    V, W, G = 16, 6, 2
    vx = _grid(V)
    dv = vx[1] - vx[0]
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    DF = np.exp(-0.5 * ((X - 0.3) ** 2 / 1.2 + (Y + 0.2) ** 2 / 0.8) ** 1.2) * (1 + 0.2 * np.tanh(X))
    DF = DF / DF.sum() / dv**2
    sa = np.array([60.0, 100.0])
    row = np.array([0.8, 0.3, 526.5, -0.8, 0.4, 3.0, 1.0, 1, 1, 1, 40.0, 8.0, 0.2, 1.0])
    grids = O.Grids([400.0, 700.0], W)
    leaves, pt = TO.params_from_block(row, 1)
    fo = torch.tensor(DF, requires_grad=True)
    ffo = TO.form_factor_2d(pt, fo, vx, grids, sa, G, 0.0, 30.0, 110.0)
    rng = np.random.default_rng(2)
    cot = rng.normal(size=tuple(ffo.shape)) / np.abs(ffo.detach().numpy()).max()
    (ffo * torch.tensor(cot)).sum().backward()
    eng = FormFactorEngine((400.0, 700.0), W, 0.0, sa, np.ones(2), G, 1, vx, mode="2v", ud_ang=30.0, va_ang=110.0)
    ptg = torch.tensor(row[None], device="cuda")
    feg = torch.tensor(DF[None], device="cuda")
    _, ff, saved = eng.forward(ptg, feg, want_ff=True)
    assert np.abs(ff.cpu().numpy()[0] - ffo.detach().numpy()).max() / np.abs(ffo.detach().numpy()).max() < 1e-10
    pb, fb = eng.backward(ptg, feg, saved, ff_bar=torch.tensor(cot[None], device="cuda"))
    gp, gf = leaves.grad.numpy(), fo.grad.numpy()
    pb, fb = pb.cpu().numpy()[0], fb.cpu().numpy()[0]
    assert np.abs(fb - gf).max() / np.abs(gf).max() < 1e-9, np.abs(fb - gf).max() / np.abs(gf).max()
    for k in [0, 1, 2, 3, 4, 5, 6, 11, 12, 13]:
        assert abs(pb[k] - gp[k]) <= 1e-4 * max(abs(gp[k]), 1e-8 * np.abs(gp).max()), (k, pb[k], gp[k])


def test_calc_in_2D_vjp_full_size_table_fd():
    """V = 128 (the arts-2d table, 225 KB of shared memory in the adjoint kernel): the VJP against central differences of the
    kernel's own forward for Te, ne and one table entry."""
    from tsadar_b200.engine import FormFactorEngine
    V, W = 128, 3
    vx = _grid(V)
    dv = vx[1] - vx[0]
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    DF = np.exp(-0.5 * (X**2 + 1.3 * Y**2)) * (1 + 0.1 * X)
    DF = np.clip(DF, 1e-30, None)
    DF = DF / DF.sum() / dv**2
    sa = np.array([60.0, 120.0])
    row = np.array([[0.8, 0.3, 526.5, 0.3, 0.2, 0.0, 0.0, 1, 1, 1, 40.0, 8.0, 0.2, 1.0]])
    eng = FormFactorEngine((450.0, 600.0), W, 0.0, sa, np.ones(2), 1, 1, vx, mode="2v", ud_ang=20.0, va_ang=70.0)
    pt, ft = torch.tensor(row, device="cuda"), torch.tensor(DF[None], device="cuda")
    _, ff, saved = eng.forward(pt, ft, want_ff=True)
    rng = np.random.default_rng(0)
    cot = torch.tensor(rng.normal(size=tuple(ff.shape)) / float(ff.abs().max()), device="cuda")
    pb, fb = eng.backward(pt, ft, saved, ff_bar=cot)
    assert torch.isfinite(pb).all() and torch.isfinite(fb).all()

    def L(rw, tab):
        _, f2, _ = eng.forward(torch.tensor(rw, device="cuda"), torch.tensor(tab[None], device="cuda"), want_ff=True)
        return float((f2 * cot).sum())
    for k, h in ((0, 1e-6), (1, 1e-6)):
        rp, rm = row.copy(), row.copy()
        rp[0, k] += h; rm[0, k] -= h
        fd = (L(rp, DF) - L(rm, DF)) / (2 * h)
        assert abs(pb[0, k].item() - fd) <= 2e-4 * abs(fd), (k, pb[0, k].item(), fd)
    i, j = 70, 61
    h = 1e-4 * DF[i, j]
    Dp, Dm = DF.copy(), DF.copy()
    Dp[i, j] += h; Dm[i, j] -= h
    fd = (L(row, Dp) - L(row, Dm)) / (2 * h)
    assert abs(fb[0, i, j].item() - fd) <= 1e-4 * abs(fd), (fb[0, i, j].item(), fd)


def test_arts2v_diagnostic_with_spherical_harmonics_matches_oracle():
    """test_arts2d_forward_pass (tests/test_forward/test_angular_2v.py:18-94) at a reduced size: SphericalHarmonics
    (Mora-Yahi l=1) producer -> calc_in_2D over the 241 ARTS angles -> [1024, 241] weight matrix -> ATS IRF ->
    resolution units, against the NumPy oracle of the same chain.  PARITY UNPINNED (ThryE-arts2v.npy is a missing blob)."""
    import os
    from oracle import params_oracle as P
    from tests.common import load_cfg
    from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
    from tsadar_b200.ts_params import ThomsonParams
    cfg = load_cfg("cfg_arts2v")
    npts = 32
    cfg["other"]["lamrangE"] = [cfg["data"]["fit_rng"]["forward_epw_start"], cfg["data"]["fit_rng"]["forward_epw_end"]]
    cfg["other"]["lamrangI"] = [cfg["data"]["fit_rng"]["forward_iaw_start"], cfg["data"]["fit_rng"]["forward_iaw_end"]]
    cfg["other"]["npts"] = npts
    cfg["other"]["extraoptions"]["spectype"] = "angular_full"
    cfg["parameters"]["electron"]["fe"]["nvx"] = 24
    cfg["parameters"]["electron"]["fe"]["params"]["nvr"] = 16
    cfg["parameters"]["electron"]["fe"]["params"]["LTx"] = 60.0      # visible anisotropy at this resolution
    cfg["parameters"]["electron"]["fe"]["params"]["LTy"] = 90.0
    tab = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tsadar_b200", "data", "arts_angles.npz"))
    sa = dict(sa=np.arange(19, 139.5, 0.5), weights=tab["weightMatrix"], angAxis=tab["angsFRED"])
    n_lam = npts // 2
    batch = dict(i_data=np.ones((1024, n_lam)), e_data=np.ones((1024, n_lam)), noise_e=np.array([0.0]), noise_i=np.array([0.0]),
                 e_amps=np.array([1.0]), i_amps=np.array([1.0]))
    p = P.thomson_params(cfg["parameters"], activate=False)
    assert np.ndim(p["electron"]["fe"]) == 2
    ref, lamb, _ = O.diagnostic_arts(p, cfg, sa, batch)
    ts_diag = ThomsonScatteringDiagnostic(cfg, scattering_angles=sa)
    ts_params = ThomsonParams(cfg["parameters"], num_params=1, batch=False)        # the reference's call (:81)
    ThryE, _, lamE, _ = ts_diag(ts_params, batch)
    got = ThryE.detach().cpu().numpy()
    assert got.shape == ref.shape
    np.testing.assert_allclose(lamE, lamb, rtol=1e-13)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-8
    # and the chain is differentiable down to the producer's leaves (m of f00, log10 LTx, log10 LTy)
    (ThryE * ThryE).sum().backward()
    g = [t.grad for t in ts_params.parameters()]
    assert len(g) == 3 and all(x is not None and torch.isfinite(x).all() for x in g) and all(float(x.abs()) > 0 for x in g)


def test_calc_in_2D_vjp_table_only():
    """Only the table trainable (the reference's arts-2d deck: Te, ne inactive): tsff_ff_bwd with params_bar = NULL skips
    the d/dbeta gather and the kinematics reverse; fe_bar must be bit-identical to the full adjoint's."""
    from tsadar_b200.engine import FormFactorEngine, form_factor_full
    V, W = 48, 12
    vx = _grid(V)
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    DF = np.exp(-0.5 * ((X - 0.2) ** 2 / 1.1 + Y**2 / 0.9) ** 1.1)
    DF = DF / DF.sum() / (vx[1] - vx[0]) ** 2
    sa = np.array([50.0, 95.0, 130.0])
    row = np.array([[0.8, 0.3, 526.5, 0.3, 0.2, 0.0, 0.0, 1, 1, 1, 40.0, 8.0, 0.2, 1.0]])
    eng = FormFactorEngine((450.0, 600.0), W, 0.0, sa, np.ones(3), 1, 1, vx, mode="2v", ud_ang=20.0, va_ang=70.0)
    pt, ft = torch.tensor(row, device="cuda"), torch.tensor(DF[None], device="cuda")
    _, ff, saved = eng.forward(pt, ft, want_ff=True)
    cot = torch.tensor(np.random.default_rng(1).normal(size=tuple(ff.shape)) / float(ff.abs().max()), device="cuda")
    pb, fb = eng.backward(pt, ft, saved, ff_bar=cot)
    fb = fb.clone()
    pb2, fb2 = eng.backward(pt, ft, saved, ff_bar=cot, want_params=False)
    assert pb2 is None and pb is not None
    assert float((fb2 - fb).abs().max()) <= 1e-13 * float(fb.abs().max())      # same taps, atomics in another order
    # through autograd: a params tensor that does not require grad selects the short path
    ftg = ft.clone().requires_grad_(True)
    (form_factor_full(eng, pt, ftg) * cot).sum().backward()
    assert float((ftg.grad - fb).abs().max()) <= 1e-13 * float(fb.abs().max())


@pytest.mark.gpu
def test_calc_all_chi_vals_matches_oracle_per_pole():
    """Boundary B2: FormFactor.calc_all_chi_vals(vx, DF, beta, xie_mag, klde_mag) -> (fe_vphi, chiEI, chiERrat),
    form_factor.py:390-447, against the oracle's per-pole calc_chi_vals (:349-388), all quadrants of beta, poles inside
    and outside the grid."""
    from tsadar_b200.form_factor import FormFactor
    V = 64
    vx = _grid(V)
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    DF = np.exp(-0.5 * ((X - 0.3) ** 2 / 1.2 + (Y + 0.2) ** 2 / 0.8) ** 1.2) * (1 + 0.2 * np.tanh(X))
    DF = DF / DF.sum() / (vx[1] - vx[0]) ** 2
    rng = np.random.default_rng(4)
    G, W, A = 1, 7, 5
    beta = rng.uniform(-np.pi, 2 * np.pi, (G, W, A))
    beta[0, 0, :] = [0.0, np.pi / 4, 3 * np.pi / 4, 5 * np.pi / 4, np.pi / 2]      # the axis / diagonal cases of the bank logic
    xie = rng.uniform(0.0, 5.0, (G, W, A))
    xie[0, 1, 0] = 7.5                                                             # beyond the table: clamped lerps
    klde = rng.uniform(0.2, 3.0, (G, W, A, 1))                                     # trailing unit axis as in the reference (:565)
    ff = FormFactor([400.0, 700.0], 16, 0.0, {"sa": np.array([60.0])}, 1, 0.0, 0.0)
    fphi, chiEI, chiER = ff.calc_all_chi_vals(vx, DF, beta, xie, klde)
    assert fphi.shape == chiEI.shape == chiER.shape == (G, W, A)
    ref = np.array([O.calc_chi_vals_2d(vx, DF, b, x, k)[:3] for b, x, k in zip(beta.ravel(), xie.ravel(), klde.ravel())])
    for got, want, name in ((fphi, ref[:, 0], "fe_vphi"), (chiEI, ref[:, 1], "chiEI"), (chiER, ref[:, 2], "chiERrat")):
        got = got.cpu().numpy().ravel()
        assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want)), (name, np.max(np.abs(got - want)), np.max(np.abs(want)))


@pytest.mark.gpu
def test_chi2v_needs_2v_context():
    import torch
    from tsadar_b200 import _ffi
    from tsadar_b200.engine import FormFactorEngine
    vx = _grid(64)
    eng = FormFactorEngine((400.0, 700.0), 16, 0.0, np.array([60.0]), np.ones(1), 1, 1, vx, mode="direct")
    z = torch.zeros(4, dtype=torch.float64, device="cuda")
    rc = _ffi.lib().tsff_chi2v_fwd(eng._ctx, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), 1, z.data_ptr(), None)
    assert rc != 0
    with pytest.raises(RuntimeError, match="2V"):
        _ffi.check(rc)


def _arts2d_deck(npts):
    """The arts-2d deck (tests/configs/arts2v_test_defaults + arts2d_test_inputs): SphericalHarmonics table on 128 x 128."""
    import os
    from oracle import params_oracle as P
    from tests.common import load_cfg
    cfg = load_cfg("cfg_arts2v")
    cfg["other"]["lamrangE"] = [cfg["data"]["fit_rng"]["forward_epw_start"], cfg["data"]["fit_rng"]["forward_epw_end"]]
    cfg["other"]["npts"] = npts
    p = P.thomson_params(cfg["parameters"], activate=False)
    return cfg, p


@pytest.mark.gpu
def test_arts2d_deck_full_shape_pole_sample_vs_oracle(monkeypatch):
    """BASELINE.json configs[3] at its OWN shape: V = 128 table (the deck's SphericalHarmonics f), W = 1024 wavelengths x 241
    angles = 246 784 poles.  The oracle (4.0e9 bicubic interpolations in NumPy) cannot do the whole image, so a random
    sample of 256 of the deck's own (beta, |xi|, k lambda_De) poles is checked: (i) tsff_chi2v_fwd (boundary B2, the kernel
    the full forward runs) against the oracle's per-pole calc_chi_vals (form_factor.py:349-388) at 1e-9; (ii) the full
    tsff_ff_fwd image at those poles against the oracle's kinematics + assembly fed with the oracle's chi at 1e-9."""
    from tsadar_b200.engine import FormFactorEngine
    cfg, p = _arts2d_deck(1024)
    vx, DF = np.asarray(p["electron"]["v"], dtype=np.float64), np.asarray(p["electron"]["fe"], dtype=np.float64)
    assert DF.shape == (128, 128)
    sa = np.arange(19, 139.5, 0.5)
    gen = cfg["parameters"]["general"]
    ud_ang, va_ang = gen.get("ud", {}).get("angle", 0.0), gen.get("Va", {}).get("angle", 0.0)
    grids = O.Grids(cfg["other"]["lamrangE"], 1024)
    lam_shift = cfg["data"]["ele_lam_shift"]
    # the oracle's pole set for the whole image: form_factor_2d with the per-pole chi stubbed out (kinematics only)
    real_chi = O.calc_chi_vals_2d
    monkeypatch.setattr(O, "calc_chi_vals_2d", lambda vx_, DF_, b, x, k: (0.0, 0.0, 0.0, None))
    _, _, parts = O.form_factor_2d(p, grids, sa, 1, lam_shift, ud_ang, va_ang, return_parts=True)
    monkeypatch.setattr(O, "calc_chi_vals_2d", real_chi)
    beta, xmag, klde = parts["beta"], parts["xie_mag"], parts["kin"]["klde"] * np.ones_like(parts["beta"])
    assert beta.shape == (1, 1024, 241)
    rng = np.random.default_rng(11)
    flat = rng.choice(beta.size, 256, replace=False)
    eng = FormFactorEngine(cfg["other"]["lamrangE"], 1024, lam_shift, sa, np.ones(241), 1, 1, vx, mode="2v", ud_ang=ud_ang, va_ang=va_ang)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
    # (i) the chi kernel on the sampled poles
    fphi, chiEI, chiER = eng.chi_vals_2v(t(DF), t(beta.ravel()[flat]), t(xmag.ravel()[flat]), t(klde.ravel()[flat]))
    ref = np.array([real_chi(vx, DF, beta.ravel()[i], xmag.ravel()[i], klde.ravel()[i])[:3] for i in flat])
    for got, want, name in ((fphi, ref[:, 0], "fe_vphi"), (chiEI, ref[:, 1], "chiEI"), (chiER, ref[:, 2], "chiERrat")):
        e = np.max(np.abs(got.cpu().numpy() - want)) / np.max(np.abs(want))
        assert e <= 1e-9, (name, e)
    # (ii) the full image: the oracle's assembly with the oracle's chi at the sampled poles (the kernel's own chi elsewhere,
    # so that the vectorised oracle assembly can run on the whole grid)
    fa, ca, cr = eng.chi_vals_2v(t(DF), t(beta), t(xmag), t(klde))
    fe_vphi, chiE = fa.cpu().numpy().copy(), (cr.cpu().numpy() + 1j * ca.cpu().numpy())
    fe_vphi.ravel()[flat] = ref[:, 0]
    chiE.ravel()[flat] = ref[:, 2] + 1j * ref[:, 1]
    want_ff, _ = O._assemble(parts["kin"], chiE, parts["chiI"], fe_vphi, grids)
    row = np.array([[p["electron"]["Te"], p["electron"]["ne"], p["general"]["lam"], p["general"]["Va"], p["general"]["ud"],
                     p["general"]["ne_gradient"], p["general"]["Te_gradient"], p["general"]["amp1"], p["general"]["amp2"], p["general"]["amp3"],
                     p["ion-1"]["A"], p["ion-1"]["Z"], p["ion-1"]["Ti"], p["ion-1"]["fract"]]], dtype=np.float64)
    _, ff, _ = eng.forward(t(row), t(DF[None]), want_ff=True)
    got_ff = ff.cpu().numpy()[0]
    e = np.max(np.abs(got_ff.ravel()[flat] - want_ff.ravel()[flat])) / np.max(np.abs(want_ff.ravel()[flat]))
    print(f"arts-2d full shape, 256 sampled poles: formfactor max|diff|/max {e:.2e}")
    assert e <= 1e-9, e


@pytest.mark.gpu
def test_calc_in_2D_vjp_full_size_table_vs_oracle_autograd():
    """V = 128 (the arts-2d table: the adjoint's 224 KB shared-memory layout, fixed-point table + edge band) against the
    ORACLE's autograd (torch float64 restatement of calc_in_2D, bicubic taps and all): fe_bar at 1e-9 of its max norm,
    parameters at 1e-4.  The run is repeated: the interior of fe_bar must be bit-identical (integer accumulation)."""
    from tsadar_b200.engine import FormFactorEngine
    V, W = 128, 3
    vx = _grid(V)
    dv = vx[1] - vx[0]
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    DF = np.exp(-0.5 * (X**2 + 1.3 * Y**2)) * (1 + 0.1 * X)
    DF = np.clip(DF, 1e-30, None)
    DF = DF / DF.sum() / dv**2
    sa = np.array([60.0, 120.0])
    row = np.array([0.8, 0.3, 526.5, 0.3, 0.2, 0.0, 0.0, 1, 1, 1, 40.0, 8.0, 0.2, 1.0])
    eng = FormFactorEngine((450.0, 600.0), W, 0.0, sa, np.ones(2), 1, 1, vx, mode="2v", ud_ang=20.0, va_ang=70.0)
    pt, ft = torch.tensor(row[None], device="cuda"), torch.tensor(DF[None], device="cuda")
    _, ff, saved = eng.forward(pt, ft, want_ff=True)
    grids = O.Grids([450.0, 600.0], W)
    leaves, po = TO.params_from_block(row, 1)
    fo = torch.tensor(DF, requires_grad=True)
    ffo = TO.form_factor_2d(po, fo, vx, grids, sa, 1, 0.0, ud_ang=20.0, va_ang=70.0)
    assert np.abs(ff.cpu().numpy()[0] - ffo.detach().numpy()).max() / np.abs(ffo.detach().numpy()).max() < 1e-10
    cot = np.random.default_rng(0).normal(size=tuple(ffo.shape)) / float(ffo.detach().abs().max())
    (ffo * torch.tensor(cot)).sum().backward()
    gp, gf = leaves.grad.numpy(), fo.grad.numpy()
    pb, fb = eng.backward(pt, ft, saved, ff_bar=torch.tensor(cot[None], device="cuda"))
    pb, fb1 = pb.cpu().numpy()[0], fb.clone()
    e = np.abs(fb1.cpu().numpy()[0] - gf).max() / np.abs(gf).max()
    print(f"V = 128 fe_bar vs oracle autograd: {e:.2e}")
    assert e < 1e-9, e
    inner = np.abs(gf[3:-3, 3:-3]).max()
    assert np.abs(fb1.cpu().numpy()[0][3:-3, 3:-3] - gf[3:-3, 3:-3]).max() / inner < 1e-9       # and on the interior's own scale
    for k in [0, 1, 2, 3, 4, 11, 12, 13]:
        assert abs(pb[k] - gp[k]) <= 1e-4 * max(abs(gp[k]), 1e-8 * np.abs(gp).max()), (k, pb[k], gp[k])
    _, fb2 = eng.backward(pt, ft, saved, ff_bar=torch.tensor(cot[None], device="cuda"))
    assert torch.equal(fb2[0, 3:-3, 3:-3], fb1[0, 3:-3, 3:-3])                                 # deterministic interior
    _, fb3 = eng.backward(pt, ft, saved, ff_bar=torch.tensor(cot[None], device="cuda"), want_params=False)
    assert torch.equal(fb3[0, 3:-3, 3:-3], fb1[0, 3:-3, 3:-3])


@pytest.mark.gpu
def test_2v_table_with_the_decks_own_angular_spectype_matches_oracle():
    """The arts-2d deck's own spectype is "angular" (tests/configs/arts2v_test_defaults.yaml): a 2-D table through calc_in_2D,
    the un-batched angle sum with the first row of the weight matrix (generate_spectra.py:187-197) and the electron IRF
    (thomson_diagnostic.py:67-73) -- against the NumPy oracle of the same chain."""
    import os
    from oracle import params_oracle as P
    from tests.common import load_cfg
    from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
    from tsadar_b200.ts_params import ThomsonParams
    cfg = load_cfg("cfg_arts2v")
    assert cfg["other"]["extraoptions"]["spectype"] == "angular"
    cfg["other"]["lamrangE"] = [cfg["data"]["fit_rng"]["forward_epw_start"], cfg["data"]["fit_rng"]["forward_epw_end"]]
    cfg["other"]["lamrangI"] = [cfg["data"]["fit_rng"]["forward_iaw_start"], cfg["data"]["fit_rng"]["forward_iaw_end"]]
    cfg["other"]["npts"] = 1024
    cfg["parameters"]["electron"]["fe"]["nvx"] = 24
    cfg["parameters"]["electron"]["fe"]["params"]["nvr"] = 16
    cfg["parameters"]["electron"]["fe"]["params"]["LTx"] = 60.0
    cfg["parameters"]["electron"]["fe"]["active"] = True
    tab = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tsadar_b200", "data", "arts_angles.npz"))
    sa = dict(sa=np.arange(19, 139.5, 0.5)[::12], weights=tab["weightMatrix"][:, ::12], angAxis=tab["angsFRED"])     # 21 angles: the oracle's 2V path is slow
    batch = dict(i_data=np.ones(1024), e_data=np.ones(1024), noise_e=np.array([0.0]), noise_i=np.array([0.0]),
                 e_amps=np.array([1.0]), i_amps=np.array([1.0]))
    cfg["other"]["npts"] = 64           # 64 x 21 poles for the oracle; pixel binning needs W % 1024 == 0 -> 64 bins here
    from tsadar_b200 import irf
    p = P.thomson_params(cfg["parameters"], activate=True)
    grids = O.Grids(cfg["other"]["lamrangE"], 64)
    lamE, mE = O.fit_model_electron(p, grids, sa, cfg["other"], 1, cfg["data"]["ele_lam_shift"])
    diag = ThomsonScatteringDiagnostic(cfg, scattering_angles=sa)
    tp = ThomsonParams(cfg["parameters"], num_params=1, batch=False, activate=True)
    lam_got, modlE, block = diag.model.electron_spectrum(tp())
    assert modlE.shape == (1, 64)
    np.testing.assert_allclose(lam_got, lamE, rtol=1e-12)
    assert np.abs(modlE.detach().cpu().numpy()[0] - mE).max() / np.abs(mE).max() < 1e-8
    modlE.sum().backward()
    g = [t.grad for t in tp.parameters()]
    assert g and all(x is not None and torch.isfinite(x).all() for x in g)
