import os, sys, copy
sys.path.insert(0, '.')
import numpy as np, torch
from tests.common import SA_P9, load_cfg
from tsadar_b200.loss_function import LossFunction
from tsadar_b200.ts_params import ThomsonParams
B=int(sys.argv[1]) if len(sys.argv)>1 else 2
cfg = load_cfg("cfg_1d")
lamb = np.linspace(400, 700, 1024)
e_data = 0.6 * np.exp(-0.5 * ((lamb - 470) / 12.0) ** 2) + 0.01
batch = dict(e_data=np.tile(e_data, (B, 1)), i_data=np.ones((B, 1024)), e_amps=np.ones(B), i_amps=np.ones(B), noise_e=np.zeros((B, 1024)), noise_i=np.zeros((B, 1024)))
batch_t = {k: torch.as_tensor(v, dtype=torch.float64, device="cuda") for k, v in batch.items()}
loss_fn = LossFunction(cfg, SA_P9, batch)
tp = ThomsonParams(copy.deepcopy(cfg["parameters"]), num_params=B, batch=True, activate=True)
for _ in range(3):
    (loss, aux), g = loss_fn.vg_loss(tp, batch_t)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    (loss, aux), g = loss_fn.vg_loss(tp, batch_t)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
import collections
c = collections.Counter(); t = collections.Counter()
for e in ev:
    c[e.name[:70]] += 1; t[e.name[:70]] += e.device_time
print("cuda kernels:", len(ev), "total device us:", sum(e.device_time for e in ev))
for k, v in t.most_common(25):
    print(f"{c[k]:4d} {v:8.1f} us  {k}")
