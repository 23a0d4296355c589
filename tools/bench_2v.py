import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from tsadar_b200.engine import FormFactorEngine
from tsadar_b200.synthetic import vgrid
dev = torch.device("cuda")
def timeit(fn, n=2, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
p = np.zeros((1, 14)); p[:, 0], p[:, 1], p[:, 2] = 0.6, 0.25, 526.5; p[:, 7:10] = 1.0; p[:, 10:14] = [40.0, 8.0, 0.2, 1.0]
pr = torch.tensor(p, device=dev)
sa = np.arange(19, 139.5, 0.5)
if len(sys.argv) > 2 and sys.argv[2] == 'mirror':
    sa = -sa            # beta around 225 deg instead of 135 deg: the other diagonal of the table
V = 128; vx = vgrid(V)
X, Y = np.meshgrid(vx, vx, indexing="ij")
DF = np.exp(-0.5 * (X**2 + Y**2)) / (2 * np.pi)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng2 = FormFactorEngine((400.0, 700.0), W, 0.0, sa, np.ones(241), 1, 1, vx, mode="2v")
fe2 = torch.tensor(DF[None], device=dev)
ms = timeit(lambda: eng2.forward(pr, fe2, want_ff=True))
print(f"fwd {ms:.1f} ms")
cot2 = torch.randn(1, 1, W, 241, dtype=torch.float64, device=dev)
_, ff2, saved2 = eng2.forward(pr, fe2, want_ff=True)
ms = timeit(lambda: eng2.backward(pr, fe2, saved2, ff_bar=cot2))
print(f"bwd {ms:.1f} ms")
ms = timeit(lambda: eng2.backward(pr, fe2, saved2, ff_bar=cot2, want_params=False))
print(f"bwd table-only {ms:.1f} ms")
