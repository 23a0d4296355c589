"""Oracle parity AT THE CONFIGURATION THE HEADLINE NUMBER IS QUOTED ON (BASELINE.json configs[4], bench.py's workload):
W = 1024 wavelengths, V = 4096 nodes, FLOAT32 f tables, direct-pole mode, FP32 block-multipole PV sweeps, and a batch large
enough (B >= 148) that tsff_ff_fwd launches the very template bench.py times (k_direct_fwd<2, float, FP32, 3>: two poles
per thread, three tree levels).  The batch is the bench's own (make_lineouts(seed=42)).

Bars (BASELINE.json north_star): spectrum <= 1e-5 relative, gradients <= 1e-4 relative, against the float64 oracle
(oracle/np_oracle.py: the reference's complex-log ratintn, all 1024 x 4094 pairs; oracle/torch_oracle.py: its autograd)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O, torch_oracle as TO
from tests.common import row_to_params, rel_err_report
from tsadar_b200.engine import FormFactorEngine, loss_fwd_bwd
from tsadar_b200.synthetic import make_lineouts, SA_SYN, LAM_RANGE, W_SYN, V_SYN

pytestmark = pytest.mark.gpu

B_BENCH_TEMPLATE = 160            # >= 148 lineouts: the two-poles-per-thread template of the benchmark
CHECKED = [0, 1, 37, 80, 121, 159]  # lineouts compared with the oracle (each ~1 s of oracle time)
ACTIVE = [0, 1, 2, 3, 4, 5, 6, 11, 12, 13]   # Te ne lam Va ud ne_grad Te_grad | Z Ti fract  (amps / A have zero gradient here)


@pytest.fixture(scope="module")
def headline():
    params, fe, vx, _ = make_lineouts(B_BENCH_TEMPLATE, seed=42)           # float32 tables, as bench.py
    assert fe.dtype == np.float32 and fe.shape == (B_BENCH_TEMPLATE, V_SYN)
    eng = FormFactorEngine(LAM_RANGE, W_SYN, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct")
    pt, ft = torch.tensor(params, device="cuda"), torch.tensor(fe, device="cuda")
    modl, _, saved = eng.forward(pt, ft)
    grids = O.Grids(list(LAM_RANGE), W_SYN)
    ref = {}
    for b in CHECKED:
        ff, _ = O.form_factor_direct(row_to_params(params[b], fe[b], vx, 1), grids, SA_SYN, 1, 0.0)
        ref[b] = ff[0, :, 0]
    return dict(params=params, fe=fe, vx=vx, eng=eng, pt=pt, ft=ft, modl=modl, saved=saved, grids=grids, ref=ref)


def test_headline_spectrum_vs_oracle(headline):
    got = headline["modl"].cpu().numpy()
    assert np.isfinite(got).all()
    worst_pw = worst_mx = 0.0
    for b in CHECKED:
        pw, mx = rel_err_report(got[b], headline["ref"][b])
        worst_pw, worst_mx = max(worst_pw, pw), max(worst_mx, mx)
        assert mx <= 1e-5, (b, pw, mx)       # max|diff| / max|S|
        assert pw <= 1e-5, (b, pw, mx)       # pointwise wherever |S| >= 1e-6 max|S|
    print(f"headline spectrum parity: pointwise {worst_pw:.2e}, max-norm {worst_mx:.2e}")


def test_headline_vjp_vs_oracle_autograd(headline):
    h = headline
    rng = np.random.default_rng(5)
    got = h["modl"].cpu().numpy()
    cot = rng.normal(size=got.shape) / np.abs(got).max(axis=1, keepdims=True)
    pb, fb = h["eng"].backward(h["pt"], h["ft"], h["saved"], modl_bar=torch.tensor(cot, device="cuda"))
    assert fb.dtype == torch.float32
    pb, fb = pb.cpu().numpy(), fb.cpu().numpy().astype(np.float64)
    worst_p = worst_f = 0.0
    for b in CHECKED:
        leaves, p = TO.params_from_block(h["params"][b], 1)
        fet = torch.tensor(h["fe"][b].astype(np.float64), requires_grad=True)
        ff = TO.form_factor_direct(p, fet, h["vx"], h["grids"], SA_SYN, 1, 0.0)
        (TO.modl_from_ff(ff, np.array([1.0])) * torch.tensor(cot[b])).sum().backward()
        gp, gf = leaves.grad.numpy(), fet.grad.numpy()
        for k in ACTIVE:
            sc = max(abs(gp[k]), 1e-8 * np.abs(gp).max())
            worst_p = max(worst_p, abs(pb[b, k] - gp[k]) / sc)
            assert abs(pb[b, k] - gp[k]) <= 1e-4 * sc, (b, k, pb[b, k], gp[k])
        ef = np.abs(fb[b] - gf).max() / np.abs(gf).max()
        worst_f = max(worst_f, ef)
        assert ef <= 1e-4, (b, ef)
        cos = np.dot(fb[b], gf) / np.linalg.norm(fb[b]) / np.linalg.norm(gf)
        assert cos > 1 - 1e-8, (b, cos)
    print(f"headline gradient parity: params {worst_p:.2e}, fe_bar max-norm {worst_f:.2e}")


def test_headline_step_through_the_fused_loss(headline):
    """The exact step bench.py times: tsff_ff_fwd -> tsff_loss_fwd_bwd (l2 against a target) -> tsff_ff_bwd, compared with
    the oracle's loss value and the autograd gradient of that loss."""
    h = headline
    B = B_BENCH_TEMPLATE
    pert = h["pt"].clone()
    pert[:, 0] *= 1.05
    pert[:, 1] *= 0.95
    target, _, _ = h["eng"].forward(pert, h["ft"])
    wq = torch.full((W_SYN,), 1.0 / W_SYN, dtype=torch.float64, device="cuda")
    modl, _, saved = h["eng"].forward(h["pt"], h["ft"])
    loss, tbar = loss_fwd_bwd(modl, target, wq, 1.0, 1.0 / B, "l2")
    pb, fb = h["eng"].backward(h["pt"], h["ft"], saved, modl_bar=tbar)
    pb, fb, tg = pb.cpu().numpy(), fb.cpu().numpy().astype(np.float64), target.cpu().numpy()
    for b in CHECKED[:3]:
        leaves, p = TO.params_from_block(h["params"][b], 1)
        fet = torch.tensor(h["fe"][b].astype(np.float64), requires_grad=True)
        ff = TO.form_factor_direct(p, fet, h["vx"], h["grids"], SA_SYN, 1, 0.0)
        lb = torch.sum((torch.tensor(tg[b]) - TO.modl_from_ff(ff, np.array([1.0]))) ** 2) / W_SYN / B
        lb.backward()
        gp, gf = leaves.grad.numpy(), fet.grad.numpy()
        for k in (0, 1, 2):
            assert abs(pb[b, k] - gp[k]) <= 1e-4 * abs(gp[k]), (b, k, pb[b, k], gp[k])
        assert np.abs(fb[b] - gf).max() / np.abs(gf).max() <= 1e-4
    # the loss itself: sum over the batch of per-lineout mean squared differences / B
    ref_loss = float(((modl - target) ** 2).mean(dim=1).sum() / B)
    assert abs(float(loss) - ref_loss) <= 1e-12 * abs(ref_loss)
