"""GPU parity, TSFF_MODE_TABLE = the reference's 1V FormFactor.__call__ (form_factor.py:163-298) + FitModel angle sum
(generate_spectra.py:171-220), through the C ABI.  Forward vs the NumPy oracle on the golden-vector deck
(tests/configs/1d-*.yaml, W=5120, A=10, V=128), VJP vs torch-f64 autograd."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O, torch_oracle as TO, params_oracle as P
from tests.common import SA_P9, DLM_M_OFFSET, load_cfg, params_to_row, row_to_params, rel_err_report

pytestmark = pytest.mark.gpu


def _engine(W, vx, sa, weights, G=1, nI=1, pv="fp32", lam=(400.0, 700.0), lam_shift=0.0, jmul=None):
    from tsadar_b200.engine import FormFactorEngine
    return FormFactorEngine(lam, W, lam_shift, sa, weights, G, nI, vx, mode="table", jmul=jmul, pv_precision=pv)


def _jmul(lam_axis, iawfilter):
    fb, fr = iawfilter[3] - iawfilter[2] / 2, iawfilter[3] + iawfilter[2] / 2
    return np.where((fb < lam_axis) & (fr > lam_axis), 10.0 ** (-iawfilter[1]), 1.0)


@pytest.mark.parametrize("pv", ["fp64", "fp32"])
def test_table_forward_golden_deck(pv):
    cfg = load_cfg("cfg_1d")
    p = P.thomson_params(cfg["parameters"], activate=True, dlm_m_offset=DLM_M_OFFSET)
    row = params_to_row(p)[None, :]
    fe, vx = p["electron"]["fe"][None, :], p["electron"]["v"]
    W = cfg["other"]["npts"]
    grids = O.Grids(cfg["other"]["lamrangE"], W)
    lamE, ref = O.fit_model_electron(p, grids, SA_P9, cfg["other"], 1, 0.0)
    w0 = float(SA_P9["weights"][0])  # the scalar-weight quirk of un-prepared `sa`
    eng = _engine(W, vx, SA_P9["sa"], w0, lam=cfg["other"]["lamrangE"], jmul=_jmul(grids.lam_axis, cfg["other"]["iawfilter"]), pv=pv)
    modl, ff, _ = eng.forward(torch.tensor(row, device="cuda"), torch.tensor(fe, device="cuda"), want_ff=True)
    got = modl.cpu().numpy()[0]
    pw, mx = rel_err_report(got, ref)
    ref_ff, _ = O.form_factor_1v(p, grids, SA_P9["sa"])
    pwf, mxf = rel_err_report(ff.cpu().numpy()[0], ref_ff)
    if pv == "fp64":
        assert pw < 1e-9 and mx < 1e-10 and pwf < 1e-9, (pw, mx, pwf, mxf)
    else:
        assert pw < 1e-5 and mx < 1e-5 and pwf < 1e-5, (pw, mx, pwf, mxf)


def test_table_forward_batch_two_ions_gradients():
    W, V, B, G, nI = 640, 96, 3, 2, 2
    vx = P.vgrid(V)
    rows, fes = [], []
    for b, (m, Te, ne) in enumerate([(2.0, 0.4, 0.15), (3.1, 1.2, 0.6), (4.5, 0.7, 0.3)]):
        rows.append([Te, ne, 526.5 + 0.3 * b, 0.5 * b, -0.2 * b, 4.0, 3.0, 1, 1, 1, 40.0, 8.0 + b, 0.2, 0.6, 1.0, 1.0, 0.1 + 0.1 * b, 0.4])
        fes.append(P.super_gaussian_projected(vx, m))
    rows, fes = np.array(rows), np.array(fes)
    grids = O.Grids([523.0, 530.0], W)
    eng = _engine(W, vx, SA_P9["sa"], SA_P9["weights"], G=G, nI=nI, lam=(523.0, 530.0), lam_shift=0.0)
    modl, _, _ = eng.forward(torch.tensor(rows, device="cuda"), torch.tensor(fes, device="cuda"))
    for b in range(B):
        ff, _ = O.form_factor_1v(row_to_params(rows[b], fes[b], vx, nI), grids, SA_P9["sa"], G, 0.0)
        ref = np.sum(np.mean(ff, 0) * SA_P9["weights"], 1)
        pw, mx = rel_err_report(modl.cpu().numpy()[b], ref)
        assert mx < 1e-5 and pw < 1e-5, (b, pw, mx)


@pytest.mark.parametrize("W,V,A", [(310, 64, 3), (96, 128, 10)])
def test_table_vjp(W, V, A):
    B, G, nI = 2, 2, 2
    vx = P.vgrid(V)
    sa = np.linspace(55.0, 65.0, A)
    wts = np.linspace(0.5, 1.5, A) / A
    rows = np.array([[0.6, 0.25, 526.0, 0.4, -0.3, 4.0, 3.0, 1, 1, 1, 40.0, 8.0, 0.2, 0.6, 1.0, 1.0, 0.3, 0.4],
                     [1.1, 0.5, 524.5, -0.6, 0.2, 1.0, 6.0, 1, 1, 1, 12.0, 6.0, 0.5, 0.3, 1.0, 1.0, 0.2, 0.7]])
    fes = np.array([P.super_gaussian_projected(vx, 2.3), P.super_gaussian_projected(vx, 3.6)])
    lam = (430.0, 640.0)
    grids = O.Grids(lam, W)
    rng = np.random.default_rng(4)
    eng = _engine(W, vx, sa, wts, G=G, nI=nI, lam=lam, lam_shift=0.05)
    pt, ft = torch.tensor(rows, device="cuda"), torch.tensor(fes, device="cuda")
    modl, _, saved = eng.forward(pt, ft)
    cot = rng.normal(size=(B, W)) / np.abs(modl.cpu().numpy()).max(axis=1, keepdims=True)
    pb, fb = eng.backward(pt, ft, saved, modl_bar=torch.tensor(cot, device="cuda"))
    pb, fb = pb.cpu().numpy(), fb.cpu().numpy()
    for b in range(B):
        leaves, p = TO.params_from_block(rows[b], nI)
        fet = torch.tensor(fes[b], requires_grad=True)
        ff = TO.form_factor_1v(p, fet, vx, grids, sa, G, 0.05)
        (TO.modl_from_ff(ff, wts) * torch.tensor(cot[b])).sum().backward()
        gp, gf = leaves.grad.numpy(), fet.grad.numpy()
        for k in [0, 1, 2, 3, 4, 5, 6, 11, 12, 13, 15, 16, 17]:
            assert abs(pb[b, k] - gp[k]) <= 1e-4 * max(abs(gp[k]), 1e-8 * np.abs(gp).max()), (b, k, pb[b, k], gp[k])
        assert np.abs(fb[b] - gf).max() / np.abs(gf).max() < 1e-4, (b, np.abs(fb[b] - gf).max() / np.abs(gf).max())


@pytest.mark.parametrize("fdt", [torch.float64, torch.float32])
def test_pair_of_windows_equals_the_two_separate_calls(fdt):
    """tsff_ff_pair_fwd / _bwd (electron + ion windows of one plasma: f-dependent tables built once, one PV adjoint sweep) against
    two tsff_ff_fwd / _bwd calls: same spectra bit for bit, cotangents = the sums (atomics reorder: 1e-10)."""
    from tsadar_b200.engine import form_factor_modl_pair, form_factor_modl
    from tsadar_b200.synthetic import vgrid, super_gaussian_projected
    B, W, nI = 3, 700, 2
    sa = np.linspace(53.6, 66.1, 5)
    vx = vgrid(192)
    rng = np.random.default_rng(5)
    fe = np.stack([super_gaussian_projected(vx, m) for m in (2.0, 2.6, 3.4)])
    p = np.zeros((B, 10 + 4 * nI))
    p[:, 0], p[:, 1], p[:, 2] = [0.5, 0.7, 0.9], [0.2, 0.3, 0.25], 526.5
    p[:, 3], p[:, 4] = [0.0, 0.4, -0.3], [0.0, 0.2, 0.5]
    p[:, 7:10] = 1.0
    p[:, 10:14] = [40.0, 8.0, 0.2, 0.6]
    p[:, 14:18] = [1.0, 1.0, 0.35, 0.4]
    engE = _engine(W, vx, sa, np.full(5, 0.2), nI=nI, lam=(380.0, 690.0), lam_shift=0.3, jmul=np.linspace(0.5, 1.5, W))
    # (a different wavelength count AND a different angle count: the two contexts' workspaces are laid out independently)
    engI = _engine(W + 37, vx, sa[:3], np.linspace(0.1, 0.3, 3), nI=nI, lam=(524.0, 529.0), lam_shift=0.0)
    cE = torch.tensor(rng.normal(size=(B, W)), device="cuda")
    cI = torch.tensor(rng.normal(size=(B, W + 37)), device="cuda")

    def run(pair):
        pt = torch.tensor(p, device="cuda", requires_grad=True)
        ft = torch.tensor(fe, device="cuda", dtype=fdt, requires_grad=True)
        if pair:
            mE, mI = form_factor_modl_pair(engE, engI, pt, ft)
        else:
            mE, mI = form_factor_modl(engE, pt, ft), form_factor_modl(engI, pt, ft)
        ((mE * cE).sum() + (mI * cI).sum()).backward()
        return mE.detach(), mI.detach(), pt.grad, ft.grad

    mE, mI, pb, fb = run(True)
    mE0, mI0, pb0, fb0 = run(False)
    assert torch.equal(mE, mE0) and torch.equal(mI, mI0)
    # params_bar: FP64 sums in a different order.  fe_bar: the PV adjoint sweep runs its far field in FP32, so one sweep over the
    # summed table cotangent and the sum of two sweeps agree to FP32 rounding of the sweep, not to FP64
    tol, tolf = 1e-10, 5e-6
    assert float((pb - pb0).abs().max()) <= tol * float(pb0.abs().max())
    assert float((fb.double() - fb0.double()).abs().max()) <= tolf * float(fb0.double().abs().max())
    # only one window's cotangent given
    pt = torch.tensor(p, device="cuda", requires_grad=True)
    ft = torch.tensor(fe, device="cuda", dtype=fdt, requires_grad=True)
    mE, mI = form_factor_modl_pair(engE, engI, pt, ft)
    (mI * cI).sum().backward()
    pt0 = torch.tensor(p, device="cuda", requires_grad=True)
    ft0 = torch.tensor(fe, device="cuda", dtype=fdt, requires_grad=True)
    (form_factor_modl(engI, pt0, ft0) * cI).sum().backward()
    assert float((pt.grad - pt0.grad).abs().max()) <= tol * float(pt0.grad.abs().max())
    assert float((ft.grad.double() - ft0.grad.double()).abs().max()) <= tolf * float(ft0.grad.double().abs().max())


def test_forward_with_angles_split_over_ctas_equals_the_unsplit_one():
    """A fit batch of a few lineouts runs the fused-angle-sum forward with the angles split over CTAs (partial sums reduced in chunk
    order by a second kernel); a batch that fills the device runs one CTA per (lineout, tile).  Same lineouts, same spectra (the
    angle sum is re-associated: 1e-14), and the split path repeats bit for bit."""
    from tsadar_b200.synthetic import vgrid, super_gaussian_projected
    W, A, nI = 900, 10, 1
    sa = np.linspace(53.6, 66.1, A)
    vx = vgrid(160)
    eng = _engine(W, vx, sa, np.linspace(0.05, 0.15, A), nI=nI, lam=(400.0, 700.0), jmul=np.linspace(0.8, 1.2, W))
    Bbig = 400
    rng = np.random.default_rng(11)
    p = np.zeros((Bbig, 14))
    p[:, 0], p[:, 1], p[:, 2] = rng.uniform(0.3, 1.2, Bbig), rng.uniform(0.1, 0.5, Bbig), 526.5
    p[:, 7:10] = 1.0
    p[:, 10:14] = [40.0, 8.0, 0.2, 1.0]
    fe = np.stack([super_gaussian_projected(vx, m) for m in rng.uniform(2.0, 4.0, Bbig)])
    pt, ft = torch.tensor(p, device="cuda"), torch.tensor(fe, device="cuda")
    big, _, _ = eng.forward(pt, ft)
    small, _, _ = eng.forward(pt[:3].contiguous(), ft[:3].contiguous())
    again, _, _ = eng.forward(pt[:3].contiguous(), ft[:3].contiguous())
    assert torch.equal(small, again)
    assert float((small - big[:3]).abs().max()) <= 1e-13 * float(big[:3].abs().max())
