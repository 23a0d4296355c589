#!/usr/bin/env python
"""Device timings of the reference's own shapes -- not the bench.py contract, a side table:
  1d       table mode, W=5120, A=10, V=320->1024-node xi1 table, B lineouts, EPW+IAW windows, forward + VJP
  arts-1d  table mode, W=2048, A=241, one image: formfactor + weights GEMM + ATS stage, forward + VJP
  arts-2d  2V mode,   W=1024, A=241, V=128 (246 784 poles x 16 384 bicubic points), forward and VJP"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from tsadar_b200.engine import FormFactorEngine
from tsadar_b200.synthetic import vgrid, super_gaussian_projected

dev = torch.device("cuda")


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def row(B, nI=1):
    p = np.zeros((B, 10 + 4 * nI))
    p[:, 0], p[:, 1], p[:, 2] = 0.6, 0.25, 526.5
    p[:, 7:10] = 1.0
    p[:, 10:14] = [40.0, 8.0, 0.2, 1.0]
    return torch.tensor(p, device=dev)


P9 = np.linspace(53.637560, 66.1191, 10)
out = []
# ---- 1d
for B in (2, 80, 1024):
    vx = vgrid(320)
    fe = torch.tensor(np.tile(super_gaussian_projected(vx, 2.5), (B, 1)), device=dev)
    engE = FormFactorEngine((319.7, 739.6), 5120, 0.0, P9, np.full(10, 0.1), 1, 1, vx, mode="table")
    engI = FormFactorEngine((523.1, 530.0), 5120, 0.0, P9, np.full(10, 0.1), 1, 1, vx, mode="table")
    pr = row(B)
    cot = torch.randn(B, 5120, dtype=torch.float64, device=dev)
    def step():
        for e in (engE, engI):
            modl, _, saved = e.forward(pr, fe)
            e.backward(pr, fe, saved, modl_bar=cot)
    ms = timeit(step)
    out.append(f"1d       B={B:5d} (EPW+IAW windows, W=5120, A=10) fwd+VJP: {ms:8.3f} ms  -> {B / ms * 1e3:10.0f} lineouts/s")
# ---- arts-1d
vx = vgrid(256)
fe = torch.tensor(super_gaussian_projected(vx, 2.5)[None], device=dev)
sa = np.arange(19, 139.5, 0.5)
eng = FormFactorEngine((400.0, 700.0), 2048, 0.0, sa, np.ones(241), 1, 1, vx, mode="table")
pr = row(1)
cot = torch.randn(1, 1, 2048, 241, dtype=torch.float64, device=dev)
def step_a():
    _, ff, saved = eng.forward(pr, fe, want_ff=True, want_modl=False)
    eng.backward(pr, fe, saved, ff_bar=cot)
ms = timeit(step_a)
out.append(f"arts-1d  formfactor [1,2048,241] (493 568 points) fwd+VJP: {ms:8.3f} ms")
# ---- arts-2d forward
V = 128
vx = vgrid(V)
X, Y = np.meshgrid(vx, vx, indexing="ij")
DF = np.exp(-0.5 * (X**2 + Y**2)) / (2 * np.pi)
eng2 = FormFactorEngine((400.0, 700.0), 1024, 0.0, sa, np.ones(241), 1, 1, vx, mode="2v")
fe2 = torch.tensor(DF[None], device=dev)
ms = timeit(lambda: eng2.forward(pr, fe2, want_ff=True), n=2, warm=1)
npole = 1024 * 241
out.append(f"arts-2d  calc_in_2D forward, {npole} poles x {V*V} bicubic points: {ms:8.1f} ms  ({npole * V * V / ms / 1e6:.1f} G interpolations/s, "
           f"{npole * V * V * 70 / ms / 1e9:.1f} TFLOP/s-equivalent FP64 at ~70 ops/point)")
cot2 = torch.randn(1, 1, 1024, 241, dtype=torch.float64, device=dev)
_, ff2, saved2 = eng2.forward(pr, fe2, want_ff=True)
ms = timeit(lambda: eng2.backward(pr, fe2, saved2, ff_bar=cot2), n=2, warm=1)
out.append(f"arts-2d  calc_in_2D VJP (rotate/project scatter + d/dbeta): {ms:8.1f} ms")
ms = timeit(lambda: eng2.backward(pr, fe2, saved2, ff_bar=cot2, want_params=False), n=2, warm=1)
out.append(f"arts-2d  calc_in_2D VJP, table cotangent only (the reference's arts-2d deck: only f is trainable): {ms:8.1f} ms")
print("\n".join(out))
