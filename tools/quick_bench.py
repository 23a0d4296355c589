"""Quick device timing of the synthetic sweep (scratch tool used while developing; bench.py is the contract)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from tsadar_b200.engine import FormFactorEngine, microbench
from tsadar_b200.synthetic import make_lineouts, SA_SYN, LAM_RANGE, W_SYN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
params, fe, vx, _ = make_lineouts(B)
eng = FormFactorEngine(LAM_RANGE, W_SYN, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct")
pt, ft = torch.tensor(params, device="cuda"), torch.tensor(fe, device="cuda")
cot = torch.randn(B, W_SYN, dtype=torch.float64, device="cuda")
saved = torch.empty(eng.saved_bytes(B), dtype=torch.uint8, device="cuda")
pb = torch.empty_like(pt); fb = torch.empty_like(ft)
for _ in range(2):
    modl, _, _ = eng.forward(pt, ft, saved=saved)
    eng.backward(pt, ft, saved, modl_bar=cot, params_bar=pb, fe_bar=fb)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
n = 5
tf = tb = 0.0
for _ in range(n):
    ev[0].record(); eng.forward(pt, ft, saved=saved); ev[1].record(); eng.backward(pt, ft, saved, modl_bar=cot, params_bar=pb, fe_bar=fb); ev[2].record()
    torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
tf /= n; tb /= n
pairs = B * 1024 * 4094
print(f"B={B} fwd {tf:.3f} ms bwd {tb:.3f} ms  -> {B/((tf+tb)*1e-3):.0f} lineouts/s; fwd {pairs/(tf*1e-3)/1e12:.3f} Tpair/s bwd {pairs/(tb*1e-3)/1e12:.3f} Tpair/s")
print("ffma peak (op/s)", microbench(0), " lg2 peak (op/s)", microbench(1))
print("modl finite:", bool(torch.isfinite(modl).all()), "grad finite:", bool(torch.isfinite(pb).all() and torch.isfinite(fb).all()))
