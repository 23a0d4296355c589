#!/usr/bin/env python
"""bench.py -- lineouts/s of the form-factor forward + VJP on B200 (BASELINE.json metric).

Workload (config.workload = "synthetic_sweep", SURVEY.md 8d / BASELINE.json configs[4]): per lineout W = 1024
wavelengths on [400,700] nm, one scattering angle (60 deg), one ion species, f(v) on V = 4096 nodes (FP32 table),
direct-pole mode: 1024 x 4094 (omega, v) pairs per lineout and pass.  One "step" = for B lineouts per GPU:
    tsff_ff_fwd  (spectrum)  ->  tsff_loss_fwd_bwd (L2 vs a seeded target: the VJP seed)  ->  tsff_ff_bwd (params_bar, fe_bar)
Lineouts are independent, so N GPUs = N ranks each owning B lineouts (weak scaling); the only collective is the NCCL
all-reduce of the scalar loss.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--lineouts B] [--impl ours|reference]

--impl reference times the oracle restatement of the reference's algorithm (torch float64 forward + autograd, all host
threads) -- JAX is not installable in this image, so the reference itself cannot run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIRS_PER_LINEOUT = 1024 * 4094          # (omega, v) pairs per lineout and pass (N-2 quirk of ratintn)
FLOP_PER_PAIR_STEP = 19                  # fwd 7 + dI/dxi 5 + f-table adjoint 7
MUFU_PER_PAIR_STEP = 3
FLOP_PER_PAIR_FWD = 12                   # the forward pole sweep computes I and dI/dxi
FLOP_PER_PAIR_BWD = 7


# stdout must carry exactly ONE JSON line (the driver parses it).  Libraries write there too (NCCL prints its version
# banner on stdout when the first communicator is created), so file descriptor 1 is pointed at stderr for the whole run
# and the JSON line goes to a private duplicate of the original stdout.
_REAL_STDOUT = None


def _claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def _clock_sampler(path, stop):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return
    stop.wait()
    p.terminate()


def _parse_clocks(path, dev_index):
    sm, smax, reasons = [], 0.0, set()
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not f[0].isdigit() or int(f[0]) != dev_index:
                continue
            sm.append(float(f[1]))
            smax = max(smax, float(f[2]))
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    except Exception:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_step(params, fe, vx, target):
    """The reference-semantics CPU path: torch-f64 forward + autograd VJP of the same loss, per lineout."""
    import torch
    from oracle import np_oracle as O, torch_oracle as TO
    grids = oracle_step.grids
    tot = 0.0
    for b in range(params.shape[0]):
        leaves, p = TO.params_from_block(params[b], 1)
        fet = torch.tensor(fe[b].astype(np.float64), requires_grad=True)
        ff = TO.form_factor_direct(p, fet, vx, grids, np.array([60.0]), 1, 0.0)
        modl = TO.modl_from_ff(ff, np.array([1.0]))
        loss = torch.sum((torch.tensor(target[b]) - modl) ** 2) / target.shape[1]
        loss.backward()
        tot += float(loss.detach())
    return tot


def run_reference(args):
    import torch
    from oracle import np_oracle as O
    from tsadar_b200.synthetic import make_lineouts, LAM_RANGE, W_SYN, V_SYN
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nb = args.ref_lineouts
    params, fe, vx, _ = make_lineouts(nb, seed=42)
    oracle_step.grids = O.Grids(list(LAM_RANGE), W_SYN)
    target = np.zeros((nb, W_SYN))
    for _ in range(args.warmup_ref):
        oracle_step(params[:1], fe[:1], vx, target[:1])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(params, fe, vx, target)
    dt = time.perf_counter() - t0
    val = nb * args.steps / dt
    line = {
        "metric": "lineouts/sec (form-factor fwd+VJP)", "value": val, "unit": "lineouts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup_ref, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": "synthetic_sweep", "W": W_SYN, "V": V_SYN, "angles": 1, "ions": 1,
                   "lineouts_per_step": nb, "note": "oracle port of the reference algorithm (JAX not installable here)"},
        "cpu_baseline": {"value": val, "unit": "lineouts/s", "cores": cores, "kind": "port",
                         "sample": f"{nb} lineouts x {args.steps} steps, torch-f64 forward + autograd"},
        "e2e": {"value": val, "unit": "lineouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lineouts", type=int, default=16384, help="lineouts per GPU per step (measured: 4096 -> 1.74M, 8192 -> 1.79M, 16384 -> 1.82M lineouts/s: fewer partial waves)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-lineouts", type=int, default=8)
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--cpu-baseline-lineouts", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        args.steps = min(args.steps, 5)
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from tsadar_b200.engine import FormFactorEngine, loss_fwd_bwd, microbench
    from tsadar_b200.synthetic import make_lineouts, SA_SYN, LAM_RANGE, W_SYN, V_SYN

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, K, Wm = args.lineouts, args.steps, max(args.warmup, 3)
    # several ranks on one node: keep each rank's pinned staging buffers on its GPU's own NUMA node (the e2e leg moves
    # ~270 MB per step per GPU from host memory).  Not at N=1, where the cpu_baseline leg wants every host core.
    numa_cpus = None
    if world > 1:
        from tsadar_b200.parallel import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local)

    # ---- inputs (pinned host copies for the e2e leg, resident device copies for the kernel leg)
    params_h, fe_h, vx, _ = make_lineouts(B, seed=42 + rank)
    NP = params_h.shape[1]
    eng = FormFactorEngine(LAM_RANGE, W_SYN, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct")
    params_pin = torch.from_numpy(params_h).pin_memory()
    fe_pin = torch.from_numpy(fe_h).pin_memory()
    params_d, fe_d = params_pin.to(dev), fe_pin.to(dev)
    wq = torch.full((W_SYN,), 1.0 / W_SYN, dtype=torch.float64, device=dev)   # nanmean over the full window
    saved = torch.empty(eng.saved_bytes(B), dtype=torch.uint8, device=dev)
    pbar = torch.empty_like(params_d)
    fbar = torch.empty_like(fe_d)
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    # seeded target = the spectrum of slightly perturbed parameters (the reference's synthetic-inverse setup)
    pert = params_d.clone()
    pert[:, 0] *= 1.05
    pert[:, 1] *= 0.95
    target, _, _ = eng.forward(pert, fe_d, saved=saved)
    target = target.clone()
    unc = 1.0
    scale = 1.0 / (B * world)

    def step(p, f):
        modl, _, _ = eng.forward(p, f, saved=saved)
        _, tbar = loss_fwd_bwd(modl, target, wq, unc, scale, "l2", loss_out=loss, want_grad=True)
        if world > 1:
            dist.all_reduce(loss)
        eng.backward(p, f, saved, modl_bar=tbar, params_bar=pbar, fe_bar=fbar)

    launches_per_step = eng.launches_fwd() + 1 + eng.launches_bwd()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev_f = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev_b = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for _ in range(Wm):
        step(params_d, fe_d)
    barrier()

    # ---- clocks during the timed regions
    stop = threading.Event()
    clk_path = os.path.join(tempfile.gettempdir(), f"tsff_clocks_{rank}.csv")
    th = threading.Thread(target=_clock_sampler, args=(clk_path, stop), daemon=True)
    if rank == 0:
        th.start()
        time.sleep(0.3)

    # ---- leg 1: device-resident inputs (value)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        step(params_d, fe_d)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    # per-kernel durations of the two pole/node sweeps, measured on the launch stream in a separate pass so that the
    # event records do not perturb the timed region above
    eng.set_profile_events(ev_f, ev_b)
    tf = tb = 0.0
    nprof = min(K, 5)
    for _ in range(nprof):
        step(params_d, fe_d)
        torch.cuda.synchronize()
        tf += ev_f[0].elapsed_time(ev_f[1])
        tb += ev_b[0].elapsed_time(ev_b[1])
    eng.set_profile_events(None, None)
    tf, tb = tf / nprof, tb / nprof

    # ---- leg 2: end to end through the public call with HOST buffers: every step copies params + fe from pinned host
    # memory and reads loss + params_bar back.  The batch is cut into chunks that alternate between two streams (each
    # with its own engine = its own scratch), so the H2D copy of one chunk overlaps the kernels of the other.
    NCH = int(os.environ.get("TSFF_E2E_CHUNKS", "2"))   # measured on B200: 2 chunks 1.69M, 4 chunks 1.61M, 8 chunks 1.48M lineouts/s
    NCH = NCH if B % NCH == 0 and B >= 64 else 1
    Bc = B // NCH
    pbar_pin = torch.empty_like(params_pin).pin_memory()
    loss_pin = torch.zeros(NCH, dtype=torch.float64).pin_memory()
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    engs = [eng, FormFactorEngine(LAM_RANGE, W_SYN, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct")]
    ch = []
    for c in range(NCH):
        ch.append(dict(p=torch.empty((Bc, NP), dtype=torch.float64, device=dev), f=torch.empty((Bc, V_SYN), dtype=fe_d.dtype, device=dev),
                       saved=torch.empty(eng.saved_bytes(Bc), dtype=torch.uint8, device=dev), pbar=torch.empty((Bc, NP), dtype=torch.float64, device=dev),
                       fbar=torch.empty((Bc, V_SYN), dtype=fe_d.dtype, device=dev), loss=torch.zeros(1, dtype=torch.float64, device=dev),
                       tgt=target[c * Bc:(c + 1) * Bc].contiguous()))

    def step_e2e():
        for c in range(NCH):
            st, e, k = streams[c % 2], engs[c % 2], ch[c]
            with torch.cuda.stream(st):
                k["p"].copy_(params_pin[c * Bc:(c + 1) * Bc], non_blocking=True)
                k["f"].copy_(fe_pin[c * Bc:(c + 1) * Bc], non_blocking=True)
                modl, _, _ = e.forward(k["p"], k["f"], saved=k["saved"])
                _, tbar = loss_fwd_bwd(modl, k["tgt"], wq, unc, scale, "l2", loss_out=k["loss"], want_grad=True)
                e.backward(k["p"], k["f"], k["saved"], modl_bar=tbar, params_bar=k["pbar"], fe_bar=k["fbar"])
                pbar_pin[c * Bc:(c + 1) * Bc].copy_(k["pbar"], non_blocking=True)
                loss_pin[c:c + 1].copy_(k["loss"], non_blocking=True)

    def e2e_region(nsteps):
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cur)
        for _ in range(nsteps):
            step_e2e()
        for st in streams:
            cur.wait_stream(st)
        if world > 1:
            tot = loss_pin.sum().to(dev)   # the scalar loss all-reduce of the sharded fit
            dist.all_reduce(tot)

    e2e_region(2)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_region(K)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    stop.set()

    t = torch.tensor([ms_total, ms_e2e, tf, tb], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, tf, tb = [float(x) for x in t.cpu()]

    if rank == 0:
        ffma_peak = microbench(0)   # FFMA/s  (x2 = FLOP/s)
        mufu_peak = microbench(1)   # MUFU op/s
        value = B * world * K / (ms_total * 1e-3)
        e2e_val = B * world * K / (ms_e2e * 1e-3)
        pairs = B * PAIRS_PER_LINEOUT
        fwd_tf = FLOP_PER_PAIR_FWD * pairs / (tf * 1e-3) / 1e12
        bwd_tf = FLOP_PER_PAIR_BWD * pairs / (tb * 1e-3) / 1e12
        peak_tf = 2 * ffma_peak / 1e12
        nominal_tf = 148 * 128 * 2 * 1.965e9 / 1e12
        step_tf = value / world * PAIRS_PER_LINEOUT * FLOP_PER_PAIR_STEP / 1e12
        clocks = _parse_clocks(clk_path, local)
        # per-launch DRAM traffic and executed pipe utilisation of the dominant kernel come from the committed ncu capture
        # of this same command (profiles/ncu_latest.json, written by tools/ncu_summary.py json); scaled to this batch
        traffic, executed = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_latest.json")))
            traffic = prof["dram_bytes_per_launch"] * B / prof["lineouts_per_launch"]
            executed = {k: prof[k] for k in ("kernel", "issue_slots_busy_pct", "fma_pipe_pct", "xu_pipe_pct", "fp64_pipe_pct")}
            executed["thread_instructions_per_pole"] = prof["warp_instructions"] * 32 / (prof["lineouts_per_launch"] * W_SYN)
            executed["source"] = prof["source"]
        except Exception:
            pass
        # the HBM view of the same kernel (why "bound" is not "hbm"): algorithmic bytes per launch = the f tables read once
        # (FP32) + the pole / spectrum rows, against the measured copy bandwidth of MEASURED_PEAKS.json (driver-written;
        # fallback: the profiling guide's 7.7 TB/s nominal)
        hbm_peak, hbm_src = 7700.0, "nominal (B200_PROFILING.md fallback)"
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        alg_bytes = B * (V_SYN * 4 + W_SYN * 8 * 2 + NP * 8)
        hbm_view = {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (tf * 1e-3) / 1e9,
                    "traffic_gbs": (traffic / (tf * 1e-3) / 1e9) if traffic else None, "peak_gbs": hbm_peak,
                    "frac": alg_bytes / (tf * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                    "note": "traffic (ncu dram bytes) exceeds the algorithmic bytes because the sweep reads, once, the 59 KB "
                            "block-multipole blob k_direct_prep derives from each 16 KB f table, and writes the residuals the "
                            "adjoint needs; neither is re-read within the launch"}
        line = {
            "metric": "lineouts/sec (form-factor fwd+VJP)", "value": value, "unit": "lineouts/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 PV sweeps / f64 assembly", "data": "synthetic",
            "config": {"workload": "synthetic_sweep", "W": W_SYN, "V": V_SYN, "angles": 1, "ions": 1,
                       "lineouts_per_gpu": B, "pairs_per_lineout": PAIRS_PER_LINEOUT, "parallelism": f"lineouts x{world}",
                       "cache": f"working set {int((fe_d.numel()*4*2 + saved.numel() + B*W_SYN*8*3)/2**20)} MiB per step > 126 MiB L2"},
            "e2e": {"value": e2e_val, "unit": "lineouts/s", "h2d_bytes_per_step": int(params_pin.numel() * 8 + fe_pin.numel() * 4),
                    "d2h_bytes_per_step": int(pbar_pin.numel() * 8 + 8 * NCH), "ms_per_step": ms_e2e / K,
                    "pipeline": f"{NCH} chunks alternating on 2 streams (H2D of a chunk overlaps the kernels of the previous one)",
                    "host_cpus_bound_to_gpu_numa_node": numa_cpus},
            "gpu_launches": launches_per_step * K,
            "clocks": clocks,
            "roofline": {
                "bound": "fp32", "kernel": "k_direct_fwd (pole sweep: I and dI/dxi)", "achieved": fwd_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": fwd_tf / peak_tf, "traffic": traffic, "executed": executed,
                "note": "achieved = ALGORITHMIC flops (12 per (omega,v) pair for I and dI/dxi, SURVEY 8d) / kernel time; the block-multipole "
                        "sweep executes ~15x fewer instructions than that pairwise count, so frac may exceed 1: 'executed' (ncu) is the pipe view",
                "peak_source": "FFMA microbenchmark in this run (MEASURED_PEAKS.json has no FP32 entry); nominal %.1f" % nominal_tf,
                "frac_of_nominal": fwd_tf / nominal_tf, "ms_per_launch": tf, "hbm": hbm_view,
                "mufu": {"achieved_gops": pairs / (tf * 1e-3) / 1e9, "peak_gops": mufu_peak / 1e9,
                         "frac": pairs / (tf * 1e-3) / mufu_peak},
                "adjoint_kernel": {"kernel": "k_pv_nodes", "achieved": bwd_tf, "frac": bwd_tf / peak_tf, "ms_per_launch": tb,
                                   "mufu_frac": pairs / (tb * 1e-3) / mufu_peak},
                "step": {"algorithmic_tflops": step_tf, "frac": step_tf / peak_tf, "frac_of_nominal": step_tf / nominal_tf,
                         "mufu_frac": value / world * PAIRS_PER_LINEOUT * MUFU_PER_PAIR_STEP / mufu_peak},
            },
        }
        if not args.no_cpu_baseline:
            from oracle import np_oracle as O
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            nb = args.cpu_baseline_lineouts
            oracle_step.grids = O.Grids(list(LAM_RANGE), W_SYN)
            tgt = target[:nb].cpu().numpy()
            oracle_step(params_h[:1], fe_h[:1], vx, tgt[:1])
            t0 = time.perf_counter()
            oracle_step(params_h[:nb], fe_h[:nb], vx, tgt)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nb / dt, "unit": "lineouts/s", "cores": cores, "kind": "port",
                                    "sample": f"{nb} lineouts of the same workload, torch-f64 forward + autograd VJP (oracle/torch_oracle.py)"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
