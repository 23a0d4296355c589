#!/usr/bin/env python
"""Device timing of the reference's `1d` deck shape (table mode: W=5120, A=10, V=320, EPW + IAW windows), fwd + VJP."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from tsadar_b200.engine import FormFactorEngine
from tsadar_b200.synthetic import vgrid, super_gaussian_projected
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
P9 = np.linspace(53.637560, 66.1191, 10)
vx = vgrid(320)
fe = torch.tensor(np.tile(super_gaussian_projected(vx, 2.5), (B, 1)), device=dev)
engE = FormFactorEngine((319.7, 739.6), 5120, 0.0, P9, np.full(10, 0.1), 1, 1, vx, mode="table")
engI = FormFactorEngine((523.1, 530.0), 5120, 0.0, P9, np.full(10, 0.1), 1, 1, vx, mode="table")
p = np.zeros((B, 14)); p[:, 0], p[:, 1], p[:, 2] = 0.6, 0.25, 526.5; p[:, 7:10] = 1.0; p[:, 10:14] = [40.0, 8.0, 0.2, 1.0]
pr = torch.tensor(p, device=dev)
cot = torch.randn(B, 5120, dtype=torch.float64, device=dev)
from tsadar_b200.engine import _FFPairFunction
PAIR = "--separate" not in sys.argv
class _Ctx:   # the autograd Function's forward / backward driven by hand (no graph bookkeeping in the timed loop)
    def save_for_backward(self, *t): self.saved_tensors = t
def step():
    if PAIR:   # the product path (FitModel.__call__): both windows in one tsff_ff_pair_fwd / _bwd
        c = _Ctx()
        _FFPairFunction.forward(c, engE, engI, pr, fe)
        _FFPairFunction.backward(c, cot, cot)
        return
    for e in (engE, engI):
        modl, _, saved = e.forward(pr, fe)
        e.backward(pr, fe, saved, modl_bar=cot)
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"1d B={B} (EPW+IAW windows, W=5120, A=10{', paired' if PAIR else ', separate calls'}) fwd+VJP: {ms:.3f} ms -> {B / ms * 1e3:.0f} lineouts/s")
