"""Boundary B1: tsadar_b200.ratintn.ratintn(f, g, z) -- same name and arguments as the reference's
ratintn.ratintn (ratintn.py:4-23) -- against the oracle's complex-log restatement, values and gradients."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import torch_oracle as TO

pytestmark = pytest.mark.gpu


def _case(N=1024, P=37, seed=0):
    rng = np.random.default_rng(seed)
    z = np.linspace(-8.2 - np.sqrt(2) / 1024, 8.2 + np.sqrt(2) / 1024, N)          # xi1 of the reference (form_factor.py:137)
    f = np.gradient(np.exp(-z**2 / 2) * (1 + 0.2 * np.tanh(z)), z[1] - z[0])
    pole = np.sort(rng.uniform(-8.0, 8.0, P))
    return z, f, pole


def test_single_pole_and_vmapped_shapes_match_oracle():
    from tsadar_b200.ratintn import ratintn
    z, f, pole = _case()
    ft = torch.tensor(f, device="cuda")
    one = ratintn(ft, torch.tensor(z - pole[3], device="cuda"), z)                  # form_factor.py:385-386 (per pole)
    assert one.shape == (1,)
    ref1 = O.ratintn(f[None, :], (z - pole[3])[None, :], z)
    assert abs(float(one) - ref1[0]) <= 1e-12 * max(1.0, abs(ref1[0]))
    g = torch.tensor(z[None, :] - pole[:, None], device="cuda")                     # form_factor.py:266-268 (vmap over poles)
    many = ratintn(ft[None, :], g, z)
    assert many.shape == (len(pole), 1)
    ref = O.pv_table(f, z, pole)
    assert np.max(np.abs(many[:, 0].cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))
    fast = ratintn(ft, g, z, precision="fp32")
    assert np.max(np.abs(fast[:, 0].cpu().numpy() - ref)) <= 2e-6 * np.max(np.abs(ref))


def test_gradients_match_oracle_autograd():
    from tsadar_b200.ratintn import ratintn
    z, f, pole = _case(N=256, P=9, seed=1)
    cot = np.random.default_rng(2).normal(size=len(pole))
    ft = torch.tensor(f, device="cuda", requires_grad=True)
    pt = torch.tensor(pole, device="cuda", requires_grad=True)
    zt = torch.tensor(z, device="cuda")
    out = ratintn(ft, zt[None, :] - pt[:, None], z)
    (out[:, 0] * torch.tensor(cot, device="cuda")).sum().backward()
    fo = torch.tensor(f, requires_grad=True)
    po = torch.tensor(pole, requires_grad=True)
    zo = torch.tensor(z)
    ref = torch.stack([TO.t_ratintn(fo[None, :], (zo - po[k])[None, :], zo).reshape(()) for k in range(len(pole))])
    (ref * torch.tensor(cot)).sum().backward()
    gf, gp = ft.grad.cpu().numpy(), pt.grad.cpu().numpy()
    assert np.max(np.abs(gf - fo.grad.numpy())) <= 1e-6 * np.max(np.abs(fo.grad.numpy()))
    assert np.max(np.abs(gp - po.grad.numpy())) <= 1e-4 * np.max(np.abs(po.grad.numpy()))


def test_refuses_what_the_kernel_does_not_cover():
    from tsadar_b200.ratintn import ratintn
    z, f, pole = _case(N=64, P=3)
    ft = torch.tensor(f, device="cuda")
    zbad = z.copy(); zbad[10] += 0.3 * (z[1] - z[0])
    with pytest.raises(ValueError, match="uniform"):
        ratintn(ft, torch.tensor(zbad - 0.1, device="cuda"), zbad)
    with pytest.raises(ValueError, match="z - pole"):
        ratintn(ft, torch.tensor((z - 0.1) * (1 + 0.01 * z), device="cuda"), z)
