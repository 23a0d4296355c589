#!/usr/bin/env python
"""Parity report on a B200: CUDA path vs the NumPy/torch float64 oracle at the benchmark shape and the reference's
own shapes.  Prints the error figures SURVEY.md 8(d) asks for (max pointwise relative error where |S| >= 1e-6 max|S|,
max|diff|/max|S|, gradient errors); run under gpurun, paste the output into profiles/."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from oracle import np_oracle as O, torch_oracle as TO
from tests.common import row_to_params, rel_err_report
from tsadar_b200 import engine as E
from tsadar_b200.synthetic import make_lineouts, SA_SYN, vgrid


def pv_level(N, P, seed=1):
    rng = np.random.default_rng(seed)
    h = 12.0 / N
    z0 = -6 + h / 2
    z = z0 + h * np.arange(N)
    f = -z * np.exp(-0.5 * z**2) * 0.4 + 0.01 * np.sin(3 * z)
    pole = rng.uniform(-7.5, 7.5, (1, P))
    ref = O.ratintn(f[None, :], z[None, :] - pole[0][:, None], z).ravel()
    fd, pd = torch.tensor(f[None], device="cuda"), torch.tensor(pole, device="cuda")
    o32, d32 = E.pv_integral(fd, z0, h, pd, precision="fp32")
    o64, d64 = E.pv_integral(fd, z0, h, pd, precision="fp64")
    sc = np.abs(ref).max()
    e32 = np.abs(o32.cpu().numpy()[0] - ref) / sc
    dsc = np.abs(d64.cpu().numpy()).max()
    print(f"PV N={N} P={P}: fp32 max {e32.max():.2e} rms {np.sqrt((e32**2).mean()):.2e} | fp64 max "
          f"{np.abs(o64.cpu().numpy()[0] - ref).max() / sc:.2e} | dI/dxi fp32 vs fp64 max {np.abs(d32.cpu().numpy() - d64.cpu().numpy()).max() / dsc:.2e}")


def direct_shape(W, V, B, seed=42):
    params, fe, vx, _ = make_lineouts(B, seed=seed, nvx=V, dtype=np.float64)
    grids = O.Grids([400, 700], W)
    res = {}
    for pv in ("fp32", "fp64"):
        eng = E.FormFactorEngine((400.0, 700.0), W, 0.0, SA_SYN, np.array([1.0]), 1, 1, vgrid(V), mode="direct", pv_precision=pv)
        pt, ft = torch.tensor(params, device="cuda"), torch.tensor(fe, device="cuda")
        modl, _, saved = eng.forward(pt, ft)
        res[pv] = (eng, pt, ft, modl, saved)
    pws, mxs, pw64 = [], [], []
    refs = []
    for b in range(B):
        ff, _ = O.form_factor_direct(row_to_params(params[b], fe[b], vx, 1), grids, SA_SYN, 1, 0.0)
        ref = ff[0, :, 0]
        refs.append(ref)
        pw, mx = rel_err_report(res["fp32"][3][b].cpu().numpy(), ref)
        pws.append(pw); mxs.append(mx)
        pw64.append(rel_err_report(res["fp64"][3][b].cpu().numpy(), ref)[0])
    print(f"direct W={W} V={V} B={B}: spectrum fp32 pointwise max {max(pws):.2e} (median over lineouts {np.median(pws):.2e}), "
          f"max|d|/max|S| {max(mxs):.2e}; fp64 PV path pointwise {max(pw64):.2e}")
    # gradients vs torch autograd of the oracle (2 lineouts)
    rng = np.random.default_rng(5)
    nb = min(B, 2)
    cot = np.stack([rng.normal(size=W) / np.abs(refs[b]).max() for b in range(B)])
    eng, pt, ft, modl, saved = res["fp32"]
    pb, fb = eng.backward(pt, ft, saved, modl_bar=torch.tensor(cot, device="cuda"))
    pb, fb = pb.cpu().numpy(), fb.cpu().numpy()
    for b in range(nb):
        leaves, p = TO.params_from_block(params[b], 1)
        fet = torch.tensor(fe[b], requires_grad=True)
        ffo = TO.form_factor_direct(p, fet, vx, grids, SA_SYN, 1, 0.0)
        (TO.modl_from_ff(ffo, np.array([1.0])) * torch.tensor(cot[b])).sum().backward()
        gp, gf = leaves.grad.numpy(), fet.grad.numpy()
        act = [0, 1, 2, 3, 4, 11, 12, 13]
        ep = max(abs(pb[b, k] - gp[k]) / max(abs(gp[k]), 1e-8 * np.abs(gp).max()) for k in act)
        ef = np.abs(fb[b] - gf).max() / np.abs(gf).max()
        cs = np.dot(fb[b], gf) / np.linalg.norm(fb[b]) / np.linalg.norm(gf)
        print(f"   lineout {b}: params_bar max rel {ep:.2e}; fe_bar max|d|/max {ef:.2e}, 1-cos {1 - cs:.1e}")


if __name__ == "__main__":
    torch.cuda.set_device(0)
    for N, P in ((4096, 4000), (1024, 1640), (128, 500)):
        pv_level(N, P)
    direct_shape(1024, 4096, 4)
    direct_shape(1024, 512, 6)
