#!/usr/bin/env python
"""Device timings of the reference's own named decks (BASELINE.json configs[0..3]), forward + VJP through
the C ABI, with the oracle port's CPU time on a bounded sample beside each:

  1d         table mode, W=5120, A=10, V=320 (f resampled on the 1024-node xi1 grid), EPW + IAW instances, B = 2 and B = 1024
  1d_series  same kernels, 80 lineouts, two ion species
  arts-1d    table mode, W=2048, A=241, one image: formfactor [1,2048,241] + angular weights + ATS stage through the diagnostic
  arts-2d    2V mode, W=1024, A=241, V=128: calc_in_2D (246 784 poles x 16 384 bicubic points), table cotangent

Called by bench.py (`run_named_configs`) -- under N ranks arts-2d runs wavelength-sharded over the ranks (strong scaling:
all-gather of the modlE slabs, all-reduce of the table cotangent, tsadar_b200/parallel.py); also a script:
    python tools/bench_configs.py"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

P9 = np.linspace(53.637560, 66.1191, 10)
SA_ARTS = np.arange(19, 139.5, 0.5)


def timeit(fn, n=5, warm=2, sync=None):
    for _ in range(warm):
        fn()
    if sync:
        sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def row(B, nI=1, dev="cuda"):
    p = np.zeros((B, 10 + 4 * nI))
    p[:, 0], p[:, 1], p[:, 2] = 0.6, 0.25, 526.5
    p[:, 7:10] = 1.0
    p[:, 10:14] = [40.0, 8.0, 0.2, 1.0]
    if nI == 2:
        p[:, 13] = 0.6
        p[:, 14:18] = [1.0, 1.0, 0.35, 0.4]
    return p


def _cpu_1d(nI):
    """oracle port, one lineout of the 1d deck: EPW + IAW instances, torch-f64 forward + autograd."""
    from oracle import np_oracle as O, torch_oracle as TO
    from tsadar_b200.synthetic import vgrid, super_gaussian_projected
    vx = vgrid(320)
    fe = super_gaussian_projected(vx, 2.5)
    t0 = time.perf_counter()
    for lam in ((319.7, 739.6), (523.1, 530.0)):
        leaves, p = TO.params_from_block(row(1, nI)[0], nI)
        fet = torch.tensor(fe, requires_grad=True)
        ff = TO.form_factor_1v(p, fet, vx, O.Grids(list(lam), 5120), P9, 1, 0.0)
        TO.modl_from_ff(ff, np.full(10, 0.1)).sum().backward()
    return time.perf_counter() - t0


def _cpu_arts1d():
    from oracle import np_oracle as O, torch_oracle as TO
    from tsadar_b200.synthetic import vgrid, super_gaussian_projected
    vx = vgrid(256)
    fe = super_gaussian_projected(vx, 2.5)
    t0 = time.perf_counter()
    leaves, p = TO.params_from_block(row(1)[0], 1)
    fet = torch.tensor(fe, requires_grad=True)
    ff = TO.form_factor_1v(p, fet, vx, O.Grids([400.0, 700.0], 2048), SA_ARTS, 1, 0.0)
    ff.sum().backward()
    return time.perf_counter() - t0


def _cpu_arts2d(npoles=64):
    """oracle port of calc_in_2D on a sample of `npoles` poles (2 wavelengths x npoles/2 angles), scaled to the image."""
    from oracle import np_oracle as O, torch_oracle as TO
    from tsadar_b200.synthetic import vgrid
    vx = vgrid(128)
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    DF = np.exp(-0.5 * (X**2 + Y**2)) / (2 * np.pi)
    leaves, p = TO.params_from_block(row(1)[0], 1)
    fet = torch.tensor(DF, requires_grad=True)
    t0 = time.perf_counter()
    ff = TO.form_factor_2d(p, fet, vx, O.Grids([400.0, 700.0], 2), SA_ARTS[: npoles // 2], 1, 0.0)
    ff.sum().backward()
    return (time.perf_counter() - t0) * (1024 * 241) / npoles


def run_named_configs(rank=0, world=1, dev=None, cpu=True):
    """-> dict for bench.py's JSON line.  Every rank must call it (collectives in the sharded arts-2d leg)."""
    import torch.distributed as dist
    from tsadar_b200.engine import FormFactorEngine, _FFPairFunction
    from tsadar_b200.synthetic import vgrid, super_gaussian_projected
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    cores = os.cpu_count() or 1
    out = {}
    # ---- 1d and 1d_series: per-rank lineout blocks, no collective (the ranks time the same local batch)
    for name, B, nI in (("1d_B2", 2, 1), ("1d_B1024", 1024, 1), ("1d_series_B80_I2", 80, 2)):
        vx = vgrid(320)
        fe = torch.tensor(np.tile(super_gaussian_projected(vx, 2.5), (B, 1)), device=dev)
        engE = FormFactorEngine((319.7, 739.6), 5120, 0.0, P9, np.full(10, 0.1), 1, nI, vx, mode="table")
        engI = FormFactorEngine((523.1, 530.0), 5120, 0.0, P9, np.full(10, 0.1), 1, nI, vx, mode="table")
        pr = torch.tensor(row(B, nI), device=dev)
        cot = torch.randn(B, 5120, dtype=torch.float64, device=dev)

        class _Ctx:   # the pair Function's forward / backward driven by hand (what FitModel.__call__ runs, minus autograd bookkeeping)
            def save_for_backward(self, *t):
                self.saved_tensors = t

        def step():
            c = _Ctx()
            _FFPairFunction.forward(c, engE, engI, pr, fe)
            _FFPairFunction.backward(c, cot, cot)
        ms = timeit(step)
        out[name] = {"shape": f"table mode, EPW+IAW windows (one tsff_ff_pair_fwd / _bwd), W=5120, A=10, I={nI}, B={B} lineouts per GPU", "fwd_vjp_ms": ms,
                     "lineouts_per_s": B * world / ms * 1e3}
        del engE, engI
    if cpu:
        torch.set_num_threads(cores)
        _cpu_1d(1)
        for key, nI in (("1d_B2", 1), ("1d_B1024", 1), ("1d_series_B80_I2", 2)):
            dt = _cpu_1d(nI)
            out[key]["cpu_port_lineouts_per_s"] = 1.0 / dt
            out[key]["cpu_sample"] = f"1 lineout (EPW + IAW instances), torch-f64 forward + autograd, {cores} cores"
    # ---- arts-1d: one image, the diagnostic chain (form factor, weight matrix, ATS stage), not sharded below the threshold
    from tests.common import load_cfg
    from tsadar_b200.thomson_diagnostic import ThomsonScatteringDiagnostic
    from tsadar_b200.ts_params import ThomsonParams
    tab = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tsadar_b200", "data", "arts_angles.npz"))
    sa = dict(sa=SA_ARTS, weights=tab["weightMatrix"], angAxis=tab["angsFRED"])

    def arts(name, npts, nvx, shard, graph=False):
        cfg = load_cfg(name)
        cfg["other"]["lamrangE"] = [cfg["data"]["fit_rng"]["forward_epw_start"], cfg["data"]["fit_rng"]["forward_epw_end"]]
        cfg["other"]["lamrangI"] = [cfg["data"]["fit_rng"]["forward_iaw_start"], cfg["data"]["fit_rng"]["forward_iaw_end"]]
        cfg["other"]["npts"] = npts
        cfg["other"]["extraoptions"]["spectype"] = "angular_full"
        if nvx:
            cfg["parameters"]["electron"]["fe"]["nvx"] = nvx
        n_lam = npts // 2
        batch = dict(i_data=np.ones((1024, n_lam)), e_data=np.ones((1024, n_lam)), noise_e=np.array([0.0]), noise_i=np.array([0.0]),
                     e_amps=np.array([1.0]), i_amps=np.array([1.0]))
        diag = ThomsonScatteringDiagnostic(cfg, sa, shard_group=None if shard else False)
        tp = ThomsonParams(cfg["parameters"], num_params=1, batch=False, activate=True)
        cot = [None]

        def step():
            for t in tp.parameters():
                t.grad = None
            ThryE, _, _, _ = diag(tp, batch)
            if cot[0] is None:
                g = torch.Generator(device=ThryE.device).manual_seed(0)
                cot[0] = torch.randn(ThryE.shape, dtype=torch.float64, device=ThryE.device, generator=g)
            (ThryE * cot[0]).sum().backward()
        ms = timeit(step, n=3, warm=2, sync=(dist.barrier if world > 1 else None))
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        graph_ms = None
        if graph and world == 1:
            # the same step captured once in a CUDA graph and replayed (what a fit loop does with fixed shapes: no per-step Python)
            try:
                cur = torch.cuda.current_stream()
                side = torch.cuda.Stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    for _ in range(3):
                        step()
                cur.wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step()
                graph_ms = timeit(g.replay, n=5, warm=2)
            except Exception as e:   # a host synchronisation inside the step: report it, keep the eager number
                graph_ms = f"capture failed: {type(e).__name__}: {str(e)[:120]}"
                torch.cuda.synchronize()
        return float(t), graph_ms

    ms, gms = arts("cfg_arts1v", 2048, None, world > 1, graph=True)     # a shard group is offered; FitModel.SHARD_MIN_POLES_1V declines it for this size
    out["arts-1d"] = {"shape": "table mode, one image: formfactor [1,2048,241] -> weights[1024,241] -> ATS stage, fwd + VJP of the fitted leaves",
                      "fwd_vjp_ms": ms, "images_per_s": 1e3 / ms, "fwd_vjp_cuda_graph_ms": gms,
                      "sharding": "none: 493 568 poles < FitModel.SHARD_MIN_POLES_1V, the collectives would cost more than they save (every rank evaluates the image)"}
    ms, _ = arts("cfg_arts2v", 1024, 128, world > 1)
    out["arts-2d"] = {"shape": "2V mode, one image: calc_in_2D on 246 784 poles x 128^2 bicubic points -> weights -> ATS stage, fwd + VJP",
                      "fwd_vjp_ms": ms, "images_per_s": 1e3 / ms,
                      "sharding": f"wavelength axis over {world} GPUs (all-gather of modlE slabs, all-reduce of the table cotangent)" if world > 1 else "none (1 GPU)",
                      "scaling": "strong"}
    if cpu:
        dt = _cpu_arts1d()
        out["arts-1d"]["cpu_port_images_per_s"] = 1.0 / dt
        out["arts-1d"]["cpu_sample"] = f"form factor [1,2048,241] forward + autograd only (no weights / ATS stage), torch-f64, {cores} cores"
        dt = _cpu_arts2d(64)
        out["arts-2d"]["cpu_port_images_per_s"] = 1.0 / dt
        out["arts-2d"]["cpu_sample"] = f"64 of the 246 784 poles (forward + autograd), scaled linearly, torch-f64, {cores} cores"
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(run_named_configs(), indent=1))
