#!/usr/bin/env python
"""Per-kernel device time of the fused fit step (1d deck, B = 2, CUDA-graph replay) from the torch profiler: python tools/fit_profile.py"""
import os, sys, copy
sys.path.insert(0, os.getcwd())
import numpy as np, torch, collections
from tests.common import SA_P9, load_cfg
from tsadar_b200.loss_function import LossFunction
from tsadar_b200.ts_params import FusedThomsonParams
from tsadar_b200.fit import fused_adam_fit
B=2
cfg = load_cfg("cfg_1d")
lamb = np.linspace(400, 700, 1024)
e_data = 0.6 * np.exp(-0.5 * ((lamb - 470) / 12.0) ** 2) + 0.5 * np.exp(-0.5 * ((lamb - 590) / 15.0) ** 2) + 0.01
batch = dict(e_data=np.tile(e_data, (B, 1)), i_data=np.ones((B, 1024)), e_amps=np.ones(B), i_amps=np.ones(B), noise_e=np.zeros((B, 1024)), noise_i=np.zeros((B, 1024)))
batch_t = {k: torch.as_tensor(v, dtype=torch.float64, device="cuda") for k, v in batch.items()}
loss_fn = LossFunction(cfg, SA_P9, batch)
closure = lambda p: loss_fn.calc_loss(p, batch_t)[0]
fused_adam_fit(closure, FusedThomsonParams(copy.deepcopy(cfg["parameters"]), num_params=B, batch=True, activate=True), 0.01, 5)
from torch.profiler import profile, ProfilerActivity
fz = FusedThomsonParams(copy.deepcopy(cfg["parameters"]), num_params=B, batch=True, activate=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fused_adam_fit(closure, fz, 0.01, 50)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:70]][0] += 1; agg[e.name[:70]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"total device time {tot/50:.1f} us per step (50 steps + capture/warmup)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{v[0]:5d} {v[1]/50:8.1f} us/step  {k}")
