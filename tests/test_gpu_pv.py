"""GPU parity: stand-alone principal-value integral (boundary B1) vs the oracle's literal ratintn
(ratintn.py:4-52) and vs torch autograd of the same formula.  Through the C ABI (tsff_pv_fwd / tsff_pv_bwd)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O, torch_oracle as TO

pytestmark = pytest.mark.gpu


def _case(N, P, B, seed):
    rng = np.random.default_rng(seed)
    h = 12.0 / N
    z0 = -6 + h / 2
    z = z0 + h * np.arange(N)
    f = np.stack([-z * np.exp(-0.5 * z**2 / s**2) / s**3 * 0.4 + 0.01 * np.sin(3 * z) for s in rng.uniform(0.7, 1.3, B)])
    pole = rng.uniform(-7.5, 7.5, (B, P))
    pole[:, 0] = z[N // 3] + 1e-7  # a pole almost on a node
    if P > 1:
        pole[:, 1] = z[5]          # a pole exactly on a node (reference would return inf/nan there; we return the limit)
    return z, z0, h, f, pole


@pytest.mark.parametrize("N,P,B", [(1024, 1640, 2), (4096, 1024, 3), (128, 37, 5), (130, 1, 1)])
def test_pv_forward_matches_ratintn(N, P, B):
    from tsadar_b200 import engine as E
    z, z0, h, f, pole = _case(N, P, B, 1)
    ref = np.stack([O.ratintn(f[b][None, :], z[None, :] - pole[b][:, None], z) for b in range(B)])
    fd = torch.tensor(f, device="cuda")
    pd = torch.tensor(pole, device="cuda")
    out64, _ = E.pv_integral(fd, z0, h, pd, precision="fp64")
    out32, dout = E.pv_integral(fd, z0, h, pd, precision="fp32")
    ok = np.ones_like(ref, dtype=bool)
    if P > 1:
        ok[:, 1] = False  # pole exactly on a node
    scale = np.abs(ref[ok]).max()
    assert np.abs(out64.cpu().numpy() - ref)[ok].max() / scale < 1e-12
    # FP32 MUFU path: absolute error relative to the O(1) scale of the integral
    assert np.abs(out32.cpu().numpy() - ref)[ok].max() / scale < 1e-7
    assert np.isfinite(out32.cpu().numpy()).all()
    # derivative wrt the pole vs central differences of the oracle
    e = 1e-6
    refp = np.stack([O.ratintn(f[b][None, :], z[None, :] - (pole[b] + e)[:, None], z) for b in range(B)])
    refm = np.stack([O.ratintn(f[b][None, :], z[None, :] - (pole[b] - e)[:, None], z) for b in range(B)])
    fdv = (refp - refm) / (2 * e)
    okd = ok.copy()
    okd[:, 0] = False
    d = dout.cpu().numpy()
    if okd.any():
        assert (np.abs(d - fdv)[okd] / np.maximum(1.0, np.abs(fdv[okd]))).max() < 1e-4


def test_pv_adjoint_matches_autograd():
    from tsadar_b200 import engine as E
    N, P, B = 512, 300, 2
    z, z0, h, f, pole = _case(N, P, B, 2)
    pole = pole[:, 2:]
    P = pole.shape[1]
    rng = np.random.default_rng(3)
    obar = rng.normal(size=(B, P))
    ft = torch.tensor(f, dtype=torch.float64, requires_grad=True)
    pt = torch.tensor(pole, dtype=torch.float64, requires_grad=True)
    zt = torch.tensor(z)
    tot = 0.0
    for b in range(B):
        tot = tot + torch.sum(TO.t_ratintn(ft[b][None, :], zt[None, :] - pt[b][:, None], zt) * torch.tensor(obar[b]))
    tot.backward()
    fbar, pbar = E.pv_integral_vjp(torch.tensor(f, device="cuda"), z0, h, torch.tensor(pole, device="cuda"),
                                   torch.tensor(obar, device="cuda"))
    gf, gp = ft.grad.numpy(), pt.grad.numpy()
    assert np.abs(fbar.cpu().numpy() - gf).max() / np.abs(gf).max() < 1e-4
    assert np.abs(pbar.cpu().numpy() - gp).max() / np.abs(gp).max() < 1e-4
