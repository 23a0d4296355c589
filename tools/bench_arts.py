#!/usr/bin/env python
"""Kernel-time breakdown of one ARTS diagnostic step (forward + backward) from the torch profiler.
    python tools/bench_arts.py cfg_arts1v 2048        |       python tools/bench_arts.py cfg_arts2v 256 64"""
import os, sys, collections
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from tests.common import load_cfg
from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
from tsadar_b200.ts_params import ThomsonParams
name, npts, nvx = sys.argv[1], int(sys.argv[2]), (int(sys.argv[3]) if len(sys.argv) > 3 else None)
tab = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tsadar_b200", "data", "arts_angles.npz"))
sa = dict(sa=np.arange(19, 139.5, 0.5), weights=tab["weightMatrix"], angAxis=tab["angsFRED"])
cfg = load_cfg(name)
cfg["other"]["lamrangE"] = [cfg["data"]["fit_rng"]["forward_epw_start"], cfg["data"]["fit_rng"]["forward_epw_end"]]
cfg["other"]["lamrangI"] = [cfg["data"]["fit_rng"]["forward_iaw_start"], cfg["data"]["fit_rng"]["forward_iaw_end"]]
cfg["other"]["npts"] = npts
cfg["other"]["extraoptions"]["spectype"] = "angular_full"
if nvx:
    cfg["parameters"]["electron"]["fe"]["nvx"] = nvx
for k in ("Te", "ne"):
    cfg["parameters"]["electron"][k]["active"] = True
n_lam = npts // 2
batch = dict(i_data=np.ones((1024, n_lam)), e_data=np.ones((1024, n_lam)), noise_e=np.array([0.0]), noise_i=np.array([0.0]),
             e_amps=np.array([1.0]), i_amps=np.array([1.0]))
diag = ThomsonScatteringDiagnostic(cfg, sa)
tp = ThomsonParams(cfg["parameters"], num_params=1, batch=False, activate=True)


def step():
    for t in tp.parameters():
        t.grad = None
    ThryE, _, _, _ = diag(tp, batch)
    (ThryE * ThryE).sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
c, t = collections.Counter(), collections.Counter()
for e in ev:
    c[e.name[:80]] += 1
    t[e.name[:80]] += e.device_time
print(f"{name} npts={npts}: {len(ev)} kernels, {sum(e.device_time for e in ev):.0f} us of device time per step")
for k, v in t.most_common(10):
    print(f"{c[k]:4d} {v:9.1f} us  {k}")
