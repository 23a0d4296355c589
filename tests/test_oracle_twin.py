"""The torch-f64 twin (gradient oracle) must agree with the NumPy oracle forward, and its autograd gradients with
central finite differences of the NumPy oracle."""
import numpy as np
import torch

from oracle import np_oracle as O, torch_oracle as TO, params_oracle as P
from tests.common import row_to_params, SA_P9


def _row(nI=1):
    row = [0.8, 0.35, 526.0, 0.7, -0.4, 3.0, 2.0, 1.2, 0.9, 1.0, 40.0, 8.0, 0.2, 1.0]
    if nI == 2:
        row[13] = 0.6
        row += [1.0, 1.0, 0.3, 0.4]
    return np.array(row)


def _fe(V, m=2.7):
    vx = P.vgrid(V)
    return vx, P.super_gaussian_projected(vx, m)


def test_twin_forward_table_and_direct():
    V, W = 128, 96
    vx, fe = _fe(V)
    row = _row(2)
    g = O.Grids([480, 580], W)
    for name in ["form_factor_1v", "form_factor_direct"]:
        ref, _ = getattr(O, name)(row_to_params(row, fe, vx, 2), g, SA_P9["sa"], 2, 0.2)
        _, p = TO.params_from_block(row, 2, requires_grad=False)
        got = getattr(TO, name)(p, torch.tensor(fe), vx, g, SA_P9["sa"], 2, 0.2).numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-300)


def test_twin_gradients_vs_finite_differences():
    V, W = 64, 48
    vx, fe = _fe(V)
    row = _row(1)
    g = O.Grids([450, 620], W)
    sa = np.array([60.0])
    rng = np.random.default_rng(0)
    for name in ["form_factor_1v", "form_factor_direct"]:
        ref, _ = getattr(O, name)(row_to_params(row, fe, vx, 1), g, sa, 1, 0.0)
        cot = rng.normal(size=ref.shape) / np.abs(ref).max()
        leaves, p = TO.params_from_block(row, 1)
        fet = torch.tensor(fe, requires_grad=True)
        (getattr(TO, name)(p, fet, vx, g, sa, 1, 0.0) * torch.tensor(cot)).sum().backward()

        def fun(r, f):
            out, _ = getattr(O, name)(row_to_params(r, f, vx, 1), g, sa, 1, 0.0)
            return float(np.sum(out * cot))

        for k in [0, 1, 2, 3, 4, 12]:
            h = 1e-6 if k == 2 else 1e-6 * max(1.0, abs(row[k]))  # lam: the ion feature is razor sharp in lam
            rp, rm = row.copy(), row.copy()
            rp[k] += h
            rm[k] -= h
            fd = (fun(rp, fe) - fun(rm, fe)) / (2 * h)
            assert abs(fd - leaves.grad[k].item()) <= 2e-5 * max(abs(fd), 1e-6), (name, k, fd, leaves.grad[k].item())
        for i in [12, 20, 31, 32, 40, 50]:
            h = 1e-4 * fe[i]
            fp, fm = fe.copy(), fe.copy()
            fp[i] += h
            fm[i] -= h
            fd = (fun(row, fp) - fun(row, fm)) / (2 * h)
            assert abs(fd - fet.grad[i].item()) <= 1e-4 * max(abs(fd), 1e-3 * float(fet.grad.abs().max())), (name, i, fd, fet.grad[i].item())
