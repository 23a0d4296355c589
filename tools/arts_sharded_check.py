#!/usr/bin/env python
"""ARTS wavelength-axis sharding on N GPUs: sharded diagnostic == single-GPU diagnostic, values and
gradients of the trainable leaves, for the 1V (table mode, halo point) and 2V (calc_in_2D) paths; device timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/arts_sharded_check.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
from tests.common import load_cfg
from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
from tsadar_b200.ts_params import ThomsonParams

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tab = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tsadar_b200", "data", "arts_angles.npz"))
sa = dict(sa=np.arange(19, 139.5, 0.5), weights=tab["weightMatrix"], angAxis=tab["angsFRED"])


def setup(name, npts, nvx=None):
    cfg = load_cfg(name)
    cfg["other"]["lamrangE"] = [cfg["data"]["fit_rng"]["forward_epw_start"], cfg["data"]["fit_rng"]["forward_epw_end"]]
    cfg["other"]["lamrangI"] = [cfg["data"]["fit_rng"]["forward_iaw_start"], cfg["data"]["fit_rng"]["forward_iaw_end"]]
    cfg["other"]["npts"] = npts
    cfg["other"]["extraoptions"]["spectype"] = "angular_full"
    if nvx:
        cfg["parameters"]["electron"]["fe"]["nvx"] = nvx
    for k in ("Te", "ne"):
        cfg["parameters"]["electron"][k]["active"] = True
    cfg["parameters"]["general"]["amp1"]["active"] = True
    cfg["parameters"]["general"]["lam"]["active"] = True
    n_lam = npts // 2
    batch = dict(i_data=np.ones((1024, n_lam)), e_data=np.ones((1024, n_lam)), noise_e=np.array([0.0]), noise_i=np.array([0.0]),
                 e_amps=np.array([1.0]), i_amps=np.array([1.0]))
    return cfg, batch


def run(cfg, batch, shard):
    diag = ThomsonScatteringDiagnostic(cfg, sa, shard_group=None if shard else False, force_shard=True)   # also below the size threshold
    tp = ThomsonParams(cfg["parameters"], num_params=1, batch=False, activate=True)
    rng = np.random.default_rng(0)

    def step():
        for t in tp.parameters():
            t.grad = None
        ThryE, _, _, _ = diag(tp, batch)
        cot = torch.tensor(rng.normal(size=tuple(ThryE.shape)), device=ThryE.device) if step.cot is None else step.cot
        step.cot = cot
        (ThryE * cot).sum().backward()
        return ThryE.detach()
    step.cot = None
    out = step()
    grads = [t.grad.clone() for t in tp.parameters()]
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return out, grads, float(ms)


ok = True
for name, npts, nvx in (("cfg_arts1v", 256, None), ("cfg_arts1v", 2048, None), ("cfg_arts2v", 64, 64), ("cfg_arts2v", 1024, 128)):
    cfg, batch = setup(name, npts, nvx)
    o1, g1, t1 = run(cfg, batch, False)
    o2, g2, t2 = run(cfg, batch, True)
    ev = float((o1 - o2).abs().max() / o1.abs().max())
    eg = max(float((a - b).abs().max() / (a.abs().max() + 1e-300)) for a, b in zip(g1, g2))
    # every rank must hold the same gradients
    gsum = torch.cat([g.reshape(-1) for g in g2]).clone()
    gmax = gsum.clone(); dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(gmax, gsum)) or float((gmax - gsum).abs().max()) <= 1e-12 * float(gmax.abs().max())
    if rank == 0:
        print(f"{name} npts={npts} nvx={nvx}: |dThryE|/max {ev:.2e}  |dgrad|/max {eg:.2e}  ranks agree {same}  "
              f"fwd+bwd 1 GPU {t1:.2f} ms -> {world} GPUs {t2:.2f} ms ({len(g2)} leaves)", flush=True)
    ok = ok and ev < 1e-9 and eg < 2e-5 and same      # gradients: FP32 PV sweeps + atomics, the north-star bar is 1e-4
dist.destroy_process_group()
if rank == 0:
    print("ARTS sharding check:", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
