"""Edge cases and size-independent properties of the CUDA path through the C ABI (task statement, parity section):
empty batches, ragged shapes (nothing a multiple of a tile), poles outside the f grid, NaN propagation confined to its
lineout, batch-position invariance, determinism, and -- at the FULL benchmark size (W = 1024, V = 4096), beside the oracle
parity of tests/test_gpu_headline.py at that same size -- linearity of the principal-value map in f, linearity of the VJP in
its cotangent and the adjoint identity <Ibar, PV f> = <PV^T Ibar, f>."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.common import row_to_params
from tsadar_b200 import engine as E
from tsadar_b200.engine import FormFactorEngine
from tsadar_b200.synthetic import make_lineouts, SA_SYN, LAM_RANGE, W_SYN, V_SYN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["direct", "table"])
def test_empty_batch_is_a_noop(mode):
    params, fe, vx, _ = make_lineouts(1, seed=0, nvx=128, dtype=np.float64)
    eng = FormFactorEngine(LAM_RANGE, 64, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode=mode)
    pt = torch.empty((0, params.shape[1]), dtype=torch.float64, device="cuda")
    ft = torch.empty((0, 128), dtype=torch.float64, device="cuda")
    modl, _, saved = eng.forward(pt, ft)
    assert modl.shape == (0, 64)
    pb, fb = eng.backward(pt, ft, saved, modl_bar=torch.empty((0, 64), dtype=torch.float64, device="cuda"))
    assert pb.shape == (0, params.shape[1]) and fb.shape == (0, 128)
    loss, tbar = E.loss_fwd_bwd(modl, modl, torch.ones(64, dtype=torch.float64, device="cuda"))
    assert float(loss) == 0.0 and tbar.shape == (0, 64)


@pytest.mark.parametrize("mode,tol", [("direct", 1e-5), ("table", 1e-5)])
def test_ragged_shapes(mode, tol):
    """W = 301 wavelengths, A = 3 angles, G = 3 gradient points, two ion species, V = 130 nodes: no axis is a multiple of a
    warp, a tile or a tree block."""
    W, V, G, nI = 301, 130, 3, 2
    sa = np.array([47.0, 63.5, 121.0])
    wts = np.array([0.2, 0.5, 0.3])
    params, fe, vx, _ = make_lineouts(3, seed=5, nvx=V, dtype=np.float64)
    rows = np.zeros((3, 10 + 4 * nI))
    rows[:, :10] = params[:, :10]
    rows[:, 5] = [0.0, 6.0, 3.0]          # ne gradient [%]
    rows[:, 6] = [4.0, 0.0, 2.0]          # Te gradient
    rows[:, 3] = [0.0, 1.5, -2.0]         # Va
    rows[:, 4] = [0.0, -0.7, 0.4]         # ud
    rows[:, 10:14] = [40.0, 8.0, 0.2, 0.7]
    rows[:, 14:18] = [1.0, 1.0, 0.4, 0.3]
    eng = FormFactorEngine(LAM_RANGE, W, 0.0, sa, wts, G, nI, vx, mode=mode)
    modl, _, _ = eng.forward(torch.tensor(rows, device="cuda"), torch.tensor(fe, device="cuda"))
    got = modl.cpu().numpy()
    grids = O.Grids(list(LAM_RANGE), W)
    fn = O.form_factor_direct if mode == "direct" else O.form_factor_1v
    for b in range(3):
        ff, _ = fn(row_to_params(rows[b], fe[b], vx, nI), grids, sa, G, 0.0)
        ref = (ff.mean(axis=0) * wts).sum(axis=1)
        assert np.abs(got[b] - ref).max() / np.abs(ref).max() < tol, (b, np.abs(got[b] - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("pv,tol", [("fp64", 1e-9), ("fp32", 5e-5)])
def test_poles_outside_the_f_grid(pv, tol):
    """A cold, dense plasma pushes xi_e = omega/(k vTe) far beyond the table's +-6: the lerps clamp (jnp.interp), the
    nearest node saturates at the ends, the PV integral sees every node as 'far'.  The semantics are pinned by the FP64
    path (1e-9).  The FP32 sweeps are held to 5e-5 here, not 1e-5: at Te = 20 eV the plasma wave is undamped (f' = 0 at
    the pole), the resonance is ~1e3 times sharper than at the benchmark's Te >= 0.5 keV and amplifies the 3e-7 relative
    error of the FP32 sums accordingly (measured 1.1e-5 of the peak)."""
    W, V = 96, 256
    params, fe, vx, _ = make_lineouts(2, seed=2, nvx=V, dtype=np.float64)
    params[:, 0] = [0.004, 0.02]          # Te [keV]
    params[:, 1] = [0.9, 0.5]
    eng = FormFactorEngine(LAM_RANGE, W, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct", pv_precision=pv)
    modl, ff, _ = eng.forward(torch.tensor(params, device="cuda"), torch.tensor(fe, device="cuda"), want_ff=True)
    got = modl.cpu().numpy()
    grids = O.Grids(list(LAM_RANGE), W)
    for b in range(2):
        ref, _ = O.form_factor_direct(row_to_params(params[b], fe[b], vx, 1), grids, SA_SYN, 1, 0.0, return_parts=False)
        ref = ref[0, :, 0]
        assert np.isfinite(got[b]).all()
        assert np.abs(got[b] - ref).max() / np.abs(ref).max() < tol


def test_nan_stays_in_its_lineout_and_batch_position_does_not_matter():
    params, fe, vx, _ = make_lineouts(5, seed=3, nvx=512, dtype=np.float64)
    eng = FormFactorEngine(LAM_RANGE, 128, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct")
    pt, ft = torch.tensor(params, device="cuda"), torch.tensor(fe, device="cuda")
    clean, _, _ = eng.forward(pt, ft)
    clean = clean.clone()
    bad = pt.clone()
    bad[2, 0] = float("nan")
    out, _, saved = eng.forward(bad, ft)
    assert torch.isnan(out[2]).all()
    keep = [0, 1, 3, 4]
    assert torch.equal(out[keep], clean[keep])                       # NaN policy: propagate, per lineout (vmap semantics)
    cot = torch.ones_like(out)
    pb, fb = eng.backward(bad, ft, saved, modl_bar=cot)
    assert torch.isfinite(pb[keep]).all() and torch.isfinite(fb[keep]).all()
    # the same lineout at another batch position and in another batch size: bit-identical spectrum
    perm = torch.tensor([4, 0, 3, 1, 2], device="cuda")
    out_p, _, _ = eng.forward(pt[perm].contiguous(), ft[perm].contiguous())
    assert torch.equal(out_p, clean[perm])
    one, _, _ = eng.forward(pt[3:4].contiguous(), ft[3:4].contiguous())
    assert torch.equal(one[0], clean[3])
    again, _, _ = eng.forward(pt, ft)
    assert torch.equal(again, clean)                                  # deterministic forward


def test_full_size_linearity_and_adjoint_identity_of_the_pv_map():
    """BASELINE.json's synthetic shape: N = 4096 nodes, P = 1024 poles per lineout, 64 lineouts.  The map f -> I is linear
    (ratintn.py:4-52): I(a f1 + b f2) = a I(f1) + b I(f2), and its hand-written adjoint satisfies the dot-product identity."""
    B, N, P = 64, V_SYN, W_SYN
    rng = np.random.default_rng(11)
    h = 12.0 / N
    z0 = -6 + h / 2
    z = z0 + h * np.arange(N)
    f1 = -z * np.exp(-0.5 * z**2) * rng.uniform(0.2, 0.6, (B, 1)) + 0.01 * np.sin(rng.uniform(1, 4, (B, 1)) * z)
    f2 = np.exp(-0.5 * ((z - rng.uniform(-2, 2, (B, 1))) / 0.7) ** 2) * rng.uniform(-0.3, 0.3, (B, 1))
    pole = np.sort(rng.uniform(-7.0, 7.0, (B, P)), axis=1)
    t = lambda a: torch.tensor(a, device="cuda")
    a_, b_ = 0.7, -1.9
    I1, _ = E.pv_integral(t(f1), z0, h, t(pole))
    I2, _ = E.pv_integral(t(f2), z0, h, t(pole))
    I12, _ = E.pv_integral(t(a_ * f1 + b_ * f2), z0, h, t(pole))
    sc = float(I12.abs().max())
    assert float((I12 - (a_ * I1 + b_ * I2)).abs().max()) / sc < 5e-7       # FP32 sweeps: 1e-5 is the spectrum bar
    Ibar = rng.normal(size=(B, P))
    fbar, _ = E.pv_integral_vjp(t(f1), z0, h, t(pole), t(Ibar))
    lhs = (t(Ibar) * I1).sum(dim=1)
    rhs = (fbar * t(f1)).sum(dim=1)
    scale = (t(Ibar).abs() * I1.abs()).sum(dim=1)
    assert float(((lhs - rhs).abs() / scale).max()) < 1e-5
    # and against the FP64 validation path of the same kernels
    I1d, _ = E.pv_integral(t(f1), z0, h, t(pole), precision="fp64")
    assert float((I1 - I1d).abs().max()) / float(I1d.abs().max()) < 1e-6


def test_full_size_vjp_is_linear_in_the_cotangent():
    """Synthetic sweep at full size (W = 1024, V = 4096, 32 lineouts): bwd(c1 + 2 c2) = bwd(c1) + 2 bwd(c2)."""
    B = 32
    params, fe, vx, _ = make_lineouts(B, seed=42)
    eng = FormFactorEngine(LAM_RANGE, W_SYN, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct")
    pt, ft = torch.tensor(params, device="cuda"), torch.tensor(fe, device="cuda")
    modl, _, saved = eng.forward(pt, ft)
    assert torch.isfinite(modl).all() and float(modl.min()) > 0.0          # a power spectrum
    g = torch.Generator(device="cuda").manual_seed(0)
    c1 = torch.randn(modl.shape, dtype=torch.float64, device="cuda", generator=g) / modl.abs().amax(dim=1, keepdim=True)
    c2 = torch.randn(modl.shape, dtype=torch.float64, device="cuda", generator=g) / modl.abs().amax(dim=1, keepdim=True)
    p1, f1 = [x.clone() for x in eng.backward(pt, ft, saved, modl_bar=c1)]
    p2, f2 = [x.clone() for x in eng.backward(pt, ft, saved, modl_bar=c2)]
    p12, f12 = eng.backward(pt, ft, saved, modl_bar=(c1 + 2 * c2).contiguous())
    ps = (p1.abs() + 2 * p2.abs()).amax(dim=0, keepdim=True) + 1e-300
    assert float(((p12 - (p1 + 2 * p2)).abs() / ps).max()) < 1e-4
    fs = (f1.abs() + 2 * f2.abs()).amax(dim=1, keepdim=True)
    assert float(((f12 - (f1 + 2 * f2)).abs() / fs).max()) < 1e-4


@pytest.mark.parametrize("V", [8192, 9216])   # (beyond ~11000 nodes the oracle would need ratcen's Taylor branch, ratintn.py:47-51)
def test_large_tables(V):
    """Maximum sizes: f tables of 8192 and 9216 nodes (the per-lineout tree blob is staged in shared memory: 14.5 B per
    node, 200 KB limit) against the oracle, forward and table cotangent; beyond the limit the library must refuse cleanly."""
    from oracle import torch_oracle as TO
    W = 48
    params, fe, vx, _ = make_lineouts(2, seed=9, nvx=V, dtype=np.float64)
    eng = FormFactorEngine(LAM_RANGE, W, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct")
    pt, ft = torch.tensor(params, device="cuda"), torch.tensor(fe, device="cuda")
    modl, _, saved = eng.forward(pt, ft)
    got = modl.cpu().numpy()
    grids = O.Grids(list(LAM_RANGE), W)
    for b in range(2):
        ref, _ = O.form_factor_direct(row_to_params(params[b], fe[b], vx, 1), grids, SA_SYN, 1, 0.0)
        ref = ref[0, :, 0]
        assert np.abs(got[b] - ref).max() / np.abs(ref).max() < 1e-5
    rng = np.random.default_rng(3)
    cot = rng.normal(size=(2, W)) / np.abs(got).max(axis=1, keepdims=True)
    pb, fb = eng.backward(pt, ft, saved, modl_bar=torch.tensor(cot, device="cuda"))
    leaves, p = TO.params_from_block(params[0], 1)
    fet = torch.tensor(fe[0], requires_grad=True)
    ffo = TO.form_factor_direct(p, fet, vx, grids, SA_SYN, 1, 0.0)
    (TO.modl_from_ff(ffo, np.array([1.0])) * torch.tensor(cot[0])).sum().backward()
    gf = fet.grad.numpy()
    assert np.abs(fb[0].cpu().numpy() - gf).max() / np.abs(gf).max() < 1e-4
    assert abs(pb[0, 0].item() - leaves.grad.numpy()[0]) <= 1e-4 * abs(leaves.grad.numpy()[0])


@pytest.mark.parametrize("V", [16384, 32768])
def test_table_too_large_is_refused(V):
    """16384 nodes pass the context check but their tree blob (14.5 B per node) exceeds the shared-memory staging limit;
    32768 nodes are refused when the context is created.  Either way: an error code and a message, no crash."""
    params, fe, vx, _ = make_lineouts(1, seed=9, nvx=V, dtype=np.float64)
    with pytest.raises(RuntimeError, match="too l"):
        eng = FormFactorEngine(LAM_RANGE, 16, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct")
        eng.forward(torch.tensor(params, device="cuda"), torch.tensor(fe, device="cuda"))


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_random_shapes_match_oracle(seed):
    """Seeded fuzz over the shape space: wavelengths, table length, angles, gradient points, ion species and mode drawn at
    random (odd sizes on purpose), every plasma parameter drawn from the range of the reference's random-fit test; spectrum
    vs the NumPy oracle at the north-star tolerance."""
    rng = np.random.default_rng(1000 + seed)
    W, V = int(rng.integers(17, 260)), int(rng.integers(40, 700))
    A, G, nI = int(rng.integers(1, 6)), int(rng.integers(1, 4)), int(rng.integers(1, 4))
    mode = ["direct", "table"][seed % 2]
    lam = (float(rng.uniform(380, 480)), float(rng.uniform(600, 720)))
    sa = np.sort(rng.uniform(35.0, 140.0, A))
    wts = rng.uniform(0.05, 1.0, A)
    B = 2
    params, fe, vx, _ = make_lineouts(B, seed=seed, nvx=V, dtype=np.float64)
    rows = np.zeros((B, 10 + 4 * nI))
    rows[:, :10] = params[:, :10]
    rows[:, 3] = rng.uniform(-3, 3, B)            # Va
    rows[:, 4] = rng.uniform(-1, 1, B)            # ud
    rows[:, 5] = rng.uniform(0, 8, B) * (G > 1)   # ne gradient [%]
    rows[:, 6] = rng.uniform(0, 8, B) * (G > 1)
    fr = rng.uniform(0.2, 1.0, nI)
    for i in range(nI):
        rows[:, 10 + 4 * i: 14 + 4 * i] = [float(rng.choice([1.0, 12.0, 40.0])), float(rng.uniform(1, 10)), float(rng.uniform(0.05, 0.6)), fr[i] / fr.sum()]
    eng = FormFactorEngine(lam, W, 0.0, sa, wts, G, nI, vx, mode=mode)
    modl, _, _ = eng.forward(torch.tensor(rows, device="cuda"), torch.tensor(fe, device="cuda"))
    got = modl.cpu().numpy()
    grids = O.Grids(list(lam), W)
    fn = O.form_factor_direct if mode == "direct" else O.form_factor_1v
    for b in range(B):
        ff, _ = fn(row_to_params(rows[b], fe[b], vx, nI), grids, sa, G, 0.0)
        ref = (ff.mean(axis=0) * wts).sum(axis=1)
        err = np.abs(got[b] - ref).max() / np.abs(ref).max()
        assert err < 1e-5, (seed, mode, W, V, A, G, nI, b, err)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_shapes_vjp_matches_autograd(seed):
    """The same fuzz for the hand-written adjoint (direct mode): params_bar and fe_bar vs torch autograd of the oracle."""
    from tests.test_gpu_direct import _torch_grads
    rng = np.random.default_rng(2000 + seed)
    W, V = int(rng.integers(17, 200)), int(rng.integers(40, 900))
    A, G, nI = int(rng.integers(1, 4)), int(rng.integers(1, 3)), int(rng.integers(1, 3))
    lam = (float(rng.uniform(380, 480)), float(rng.uniform(600, 720)))
    sa = np.sort(rng.uniform(35.0, 140.0, A))
    wts = rng.uniform(0.05, 1.0, A)
    B = 2
    params, fe, vx, _ = make_lineouts(B, seed=seed, nvx=V, dtype=np.float64)
    rows = np.zeros((B, 10 + 4 * nI))
    rows[:, :10] = params[:, :10]
    rows[:, 3], rows[:, 4] = rng.uniform(-3, 3, B), rng.uniform(-1, 1, B)
    rows[:, 5], rows[:, 6] = rng.uniform(0, 8, B) * (G > 1), rng.uniform(0, 8, B) * (G > 1)
    fr = rng.uniform(0.2, 1.0, nI)
    for i in range(nI):
        rows[:, 10 + 4 * i: 14 + 4 * i] = [float(rng.choice([1.0, 12.0, 40.0])), float(rng.uniform(1, 10)), float(rng.uniform(0.05, 0.6)), fr[i] / fr.sum()]
    eng = FormFactorEngine(lam, W, 0.0, sa, wts, G, nI, vx, mode="direct")
    pt, ft = torch.tensor(rows, device="cuda"), torch.tensor(fe, device="cuda")
    modl, _, saved = eng.forward(pt, ft)
    cot = rng.normal(size=(B, W)) / np.abs(modl.cpu().numpy()).max(axis=1, keepdims=True)
    pb, fb = eng.backward(pt, ft, saved, modl_bar=torch.tensor(cot, device="cuda"))
    gp, gf = _torch_grads(rows, fe, vx, O.Grids(list(lam), W), sa, wts, G, nI, cot)
    pb, fb = pb.cpu().numpy(), fb.cpu().numpy()
    cols = [0, 1, 2, 3, 4] + ([5, 6] if G > 1 else []) + [c for i in range(nI) for c in (11 + 4 * i, 12 + 4 * i, 13 + 4 * i)]
    for b in range(B):
        assert np.abs(fb[b] - gf[b]).max() / np.abs(gf[b]).max() < 1e-4, (seed, b)
        for k in cols:
            assert abs(pb[b, k] - gp[b, k]) <= 1e-4 * max(abs(gp[b, k]), 1e-6 * np.abs(gp[b]).max()), (seed, b, k, pb[b, k], gp[b, k])
