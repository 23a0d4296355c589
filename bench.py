#!/usr/bin/env python
"""bench.py -- lineouts/s of the form-factor forward + VJP on B200 (BASELINE.json metric).

Workload (config.workload = "synthetic_sweep", SURVEY.md 8d / BASELINE.json configs[4]): per lineout W = 1024
wavelengths on [400,700] nm, one scattering angle (60 deg), one ion species, f(v) on V = 4096 nodes (FP32 table),
direct-pole mode: 1024 x 4094 (omega, v) pairs per lineout and pass.  One "step" = for B lineouts per GPU:
    tsff_ff_fwd  (spectrum)  ->  tsff_loss_fwd_bwd (L2 vs a seeded target: the VJP seed)  ->  tsff_ff_bwd (params_bar, fe_bar)
Lineouts are independent, so N GPUs = N ranks each owning B lineouts (weak scaling); the only collective is the NCCL
all-reduce of the scalar loss, once per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--lineouts B] [--impl ours|reference]

The JSON line carries, besides the contract's keys:
  roofline    EXECUTED FP32 flop of the dominant kernel (opcode counts of this binary from the committed ncu source-page
              capture, profiles/ncu_latest.json) / its CUDA-event duration measured here / the FFMA peak measured here.
              The pairwise-equivalent figure of SURVEY 8(d) (what an all-pairs sweep would need) is a separate key.
  parity      two lineouts of the timed batch against the float64 oracle (spectrum, gradients) after the timed region.
  sustained   the same step back to back for >= 2 s, with its own clock record.
  configs     the reference's named decks (1d, 1d_series, arts-1d, arts-2d): fwd+VJP ms on the device, the oracle-port
              CPU time of a bounded sample beside it; arts-2d runs wavelength-sharded over the N ranks (strong scaling).
--impl reference times the oracle restatement of the reference's algorithm (torch float64 forward + autograd, all host
threads) on a bounded sample of the same workload per step -- JAX is not installable in this image, so the reference
itself cannot run."""
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
FLOP_PER_PAIR_FWD = 12                   # the forward pole sweep computes I and dI/dxi
FLOP_PER_PAIR_BWD = 7
CPU_SAMPLE_LINEOUTS = 32                 # bounded CPU sample per step: the first lineouts of the rank-0 batch (seed 42)
METRIC = "lineouts/sec (form-factor fwd+VJP)"

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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons every 100 ms while a timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, tag):
        self.path = os.path.join(tempfile.gettempdir(), f"tsff_clocks_{tag}_{os.getpid()}.csv")
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.3)
        except Exception:
            self.p = None
        return self

    def stop(self, dev_index):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                pass
        sm, smax, pw, reasons = [], 0.0, 0.0, set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8 or not f[0].isdigit() or int(f[0]) != dev_index:
                    continue
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
                try:
                    pw = max(pw, float(f[3]))
                except ValueError:
                    pass
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        except Exception:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm), "power_w_max": pw}


# ---- the reference-semantics CPU path -----------------------------------------------------------------------------------
def oracle_step(params, fe, vx, target, grids, keep=None):
    """torch-f64 forward + autograd VJP of the same loss, per lineout (oracle/torch_oracle.py: the reference's complex-log
    ratintn over all 1024 x 4094 pairs).  keep: optional dict filled with modl / params_bar / fe_bar of every lineout."""
    import torch
    from oracle import torch_oracle as TO
    tot = 0.0
    B = params.shape[0]
    for b in range(B):
        leaves, p = TO.params_from_block(params[b], 1)
        fet = torch.tensor(fe[b].astype(np.float64), requires_grad=True)
        ff = TO.form_factor_direct(p, fet, vx, grids, np.array([60.0]), 1, 0.0)
        modl = TO.modl_from_ff(ff, np.array([1.0]))
        loss = torch.sum((torch.tensor(target[b]) - modl) ** 2) / target.shape[1]
        loss.backward()
        tot += float(loss.detach())
        if keep is not None:
            keep.setdefault("modl", []).append(modl.detach().numpy())
            keep.setdefault("pbar", []).append(leaves.grad.numpy().copy())
            keep.setdefault("fbar", []).append(fet.grad.numpy().copy())
    return tot


def workload_config(B, world):
    from tsadar_b200.synthetic import W_SYN, V_SYN
    ws_mib = int(B * (V_SYN * 4 * 2 + 2 * W_SYN * 8 + 19 * 8 + 3 * W_SYN * 8) / 2**20)   # tables + cotangents, residuals, spectra
    return {"workload": "synthetic_sweep", "W": W_SYN, "V": V_SYN, "angles": 1, "ions": 1, "lineouts_per_gpu": B,
            "pairs_per_lineout": PAIRS_PER_LINEOUT, "parallelism": f"lineouts x{world}",
            "cache": f"working set {ws_mib} MiB per step > 126 MiB L2 (inputs larger than L2, no flush needed)"}


def run_reference(args):
    """--impl reference: the oracle port on the host cores, rank 0 only.  Same metric / unit / config as our arm; a step is
    a bounded sample (CPU_SAMPLE_LINEOUTS lineouts of the rank-0 batch, seed 42 -- the sample cpu_baseline uses too)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import np_oracle as O
    from tsadar_b200.synthetic import make_lineouts, LAM_RANGE, W_SYN
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nb, K, Wm = args.cpu_lineouts, args.steps, args.warmup
    params, fe, vx, _ = make_lineouts(nb, seed=42)
    grids = O.Grids(list(LAM_RANGE), W_SYN)
    target = np.zeros((nb, W_SYN))
    for _ in range(Wm):
        oracle_step(params, fe, vx, target, grids)
    t0 = time.perf_counter()
    for _ in range(K):
        oracle_step(params, fe, vx, target, grids)
    dt = time.perf_counter() - t0
    val = nb * K / dt
    cfg = workload_config(args.lineouts, max(args.gpus, 1))
    line = {
        "metric": METRIC, "value": val, "unit": "lineouts/s", "n_gpus": args.gpus, "steps": K, "warmup": Wm,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference", "config": cfg,
        "cpu_baseline": {"value": val, "unit": "lineouts/s", "cores": cores, "kind": "port",
                         "sample": f"each step = the first {nb} lineouts of the workload's rank-0 batch (seed 42), torch-f64 forward + "
                                   f"autograd VJP (oracle/torch_oracle.py; JAX not installable here)"},
        "e2e": {"value": val, "unit": "lineouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---- executed-work counts of this binary (ncu source page), for the in-run roofline ----------------------------------------
def load_counts():
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_latest.json")))
    except Exception:
        return None, "profiles/ncu_latest.json missing"
    stale = None
    try:
        from tsadar_b200 import build as _b
        if prof.get("source_stamp") and prof["source_stamp"] != _b._stamp():
            stale = "kernel sources changed since the ncu capture: counts are from the previous binary"
    except Exception:
        pass
    return prof, stale


def kernel_view(prof, key, B, ms, ffma_peak, sm_mhz, n_sm):
    """Executed rates of one kernel: counts per lineout (ncu) x B / the duration measured in this run."""
    k = (prof or {}).get("kernels", {}).get(key)
    if not k or not ms:
        return None
    s = ms * 1e-3
    out = {"fp32_tflops": k["fp32_flop_per_lineout"] * B / s / 1e12, "fp64_tflops": k["fp64_flop_per_lineout"] * B / s / 1e12,
           "mufu_gops": k["mufu_per_lineout"] * B / s / 1e9, "dram_bytes": k.get("dram_bytes_per_lineout", 0.0) * B}
    out["frac"] = out["fp32_tflops"] * 1e12 / (2 * ffma_peak)
    if sm_mhz:
        out["issue_slot_frac"] = k["warp_inst_per_lineout"] * B / (s * sm_mhz * 1e6 * n_sm * 4)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lineouts", type=int, default=16384, help="lineouts per GPU per step (measured: 4096 -> 1.74M, 8192 -> 1.79M, 16384 -> 1.82M lineouts/s: fewer partial waves)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-lineouts", type=int, default=CPU_SAMPLE_LINEOUTS, help="lineouts per step of the CPU arm / the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the named-deck block")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
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
    params_h, fe_h, vx, m_syn = make_lineouts(B, seed=42 + rank)
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
    last = {}

    def step(p, f):
        modl, _, _ = eng.forward(p, f, saved=saved)
        _, tbar = loss_fwd_bwd(modl, target, wq, unc, scale, "l2", loss_out=loss, want_grad=True)
        if world > 1:
            dist.all_reduce(loss)
        eng.backward(p, f, saved, modl_bar=tbar, params_bar=pbar, fe_bar=fbar)
        last["modl"] = modl

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
    clk = ClockSampler("main").start() if rank == 0 else None

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
    # results of the device leg, kept for the parity block
    par_idx = [0, B - 1]
    got = {"modl": last["modl"][par_idx].cpu().numpy(), "pbar": pbar[par_idx].cpu().numpy(),
           "fbar": fbar[par_idx].double().cpu().numpy(), "target": target[par_idx].cpu().numpy()} if rank == 0 else None

    # ---- leg 2: end to end with HOST buffers, two entries:
    #  "params" (the line's `e2e`): where the reference's public call enters -- LossFunction.vg_loss(weights, batch)
    #      (loss_function.py:149-168): the normalised trainable leaves [B, 6] and the data batch [B, 1024] come from pinned
    #      host memory every step; tsff_params_fwd produces the parameter block and the f tables ON THE DEVICE (ThomsonParams
    #      + DLM1V), then form factor -> loss -> adjoint -> tsff_params_bwd; loss and d loss / d leaves go back to the host.
    #  "raw" (`e2e_raw_tables`): the raw operands of tsff_ff_fwd / _bwd -- params [B, 14] + fe [B, 4096] f32 in; loss,
    #      params_bar AND fe_bar out.
    # The batch is cut into chunks that alternate between two streams (each with its own engine = its own scratch), so the
    # copies of one chunk overlap the kernels of the other; the scalar loss is all-reduced every step.
    from tsadar_b200 import _ffi
    from tsadar_b200.ts_params import FusedThomsonParams
    NCH = int(os.environ.get("TSFF_E2E_CHUNKS", "2"))   # measured on B200: 2 chunks 1.69M, 4 chunks 1.61M, 8 chunks 1.48M lineouts/s
    NCH = NCH if B % NCH == 0 and B >= 64 else 1
    Bc = B // NCH
    # the producer side: the reference's 1d deck (tests/golden/cfg_1d.json = tests/configs/1d-*.yaml) on the 4096-node grid; the
    # active leaves (Te, ne, lam, amp1, amp2, m: those of test_1d_random.py) are set so that the physical values are the
    # synthetic lineouts' own
    par = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg_1d.json")))["parameters"]
    par["electron"]["fe"]["nvx"] = V_SYN
    fz = FusedThomsonParams(par, num_params=B, batch=True, activate=True, fe_dtype=torch.float32)
    phys = {("electron", "Te"): params_h[:, 0], ("electron", "ne"): params_h[:, 1], ("general", "lam"): params_h[:, 2],
            ("general", "amp1"): params_h[:, 7], ("general", "amp2"): params_h[:, 8], ("electron", "m"): m_syn}
    x_h = np.zeros((B, fz.NLA))
    for k, name in enumerate(fz.active_names):
        lo, sc = float(fz._cfg.shift[fz._keys.index(name)]), float(fz._cfg.scale[fz._keys.index(name)])
        u = (phys[name] - lo) / sc
        x_h[:, k] = np.log(u / (1.0 - u))
    x_pin = torch.from_numpy(x_h).pin_memory()
    xbar_pin = torch.empty_like(x_pin).pin_memory()
    tgt_pin = target.cpu().pin_memory()
    pbar_pin = torch.empty_like(params_pin).pin_memory()
    fbar_pin = torch.empty_like(fe_pin).pin_memory()
    loss_pin = torch.zeros(NCH, dtype=torch.float64).pin_memory()
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    engs = [eng, FormFactorEngine(LAM_RANGE, W_SYN, 0.0, SA_SYN, np.array([1.0]), 1, 1, vx, mode="direct")]
    loss_tot = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(2)]
    ch = []
    for c in range(NCH):
        ch.append(dict(p=torch.empty((Bc, NP), dtype=torch.float64, device=dev), f=torch.empty((Bc, V_SYN), dtype=fe_d.dtype, device=dev),
                       saved=torch.empty(eng.saved_bytes(Bc), dtype=torch.uint8, device=dev), pbar=torch.empty((Bc, NP), dtype=torch.float64, device=dev),
                       fbar=torch.empty((Bc, V_SYN), dtype=fe_d.dtype, device=dev),
                       xa=torch.empty((Bc, fz.NLA), dtype=torch.float64, device=dev), xab=torch.empty((Bc, fz.NLA), dtype=torch.float64, device=dev),
                       xs=fz.x_static[c * Bc:(c + 1) * Bc].contiguous(), tg=torch.empty((Bc, W_SYN), dtype=torch.float64, device=dev),
                       loss=[torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(2)],   # double-buffered by step parity
                       loss_ev=[torch.cuda.Event() for _ in range(2)],
                       tgt=target[c * Bc:(c + 1) * Bc].contiguous()))
    red_done = [torch.cuda.Event() for _ in range(2)]
    e2e_count = [0]
    L = _ffi.lib()

    def step_e2e(entry):
        par_ = e2e_count[0] & 1
        e2e_count[0] += 1
        for c in range(NCH):
            st, e, k = streams[c % 2], engs[c % 2], ch[c]
            sl = slice(c * Bc, (c + 1) * Bc)
            with torch.cuda.stream(st):
                if entry.startswith("params"):
                    k["xa"].copy_(x_pin[sl], non_blocking=True)
                    tg = k["tgt"]                               # the data batch is resident (fit.batch_to_device) ...
                    if entry == "params+data":                   # ... or re-sent with every call, as the reference's vg_loss does
                        k["tg"].copy_(tgt_pin[sl], non_blocking=True)
                        tg = k["tg"]
                    _ffi.check(L.tsff_params_fwd(fz._cfg_ref(), Bc, k["xa"].data_ptr(), k["xs"].data_ptr(), k["p"].data_ptr(), k["f"].data_ptr(),
                                                 st.cuda_stream))
                else:
                    k["p"].copy_(params_pin[sl], non_blocking=True)
                    k["f"].copy_(fe_pin[sl], non_blocking=True)
                    tg = k["tgt"]
                modl, _, _ = e.forward(k["p"], k["f"], saved=k["saved"])
                if world > 1 and e2e_count[0] > 2:
                    st.wait_event(red_done[par_])          # the reduce that read this loss slot two steps ago
                _, tbar = loss_fwd_bwd(modl, tg, wq, unc, scale, "l2", loss_out=k["loss"][par_], want_grad=True)
                k["loss_ev"][par_].record(st)
                e.backward(k["p"], k["f"], k["saved"], modl_bar=tbar, params_bar=k["pbar"], fe_bar=k["fbar"])
                if entry.startswith("params"):
                    _ffi.check(L.tsff_params_bwd(fz._cfg_ref(), Bc, k["xa"].data_ptr(), k["xs"].data_ptr(), k["pbar"].data_ptr(),
                                                 k["fbar"].data_ptr(), k["xab"].data_ptr(), st.cuda_stream))
                    xbar_pin[sl].copy_(k["xab"], non_blocking=True)
                else:
                    pbar_pin[sl].copy_(k["pbar"], non_blocking=True)
                    fbar_pin[sl].copy_(k["fbar"], non_blocking=True)
                if world == 1:
                    loss_pin[c:c + 1].copy_(k["loss"][par_], non_blocking=True)
        if world > 1:
            # the scalar loss of the sharded batch: ONE all-reduce per step, issued on the last chunk's stream as soon as both
            # chunks' loss kernels are done; the other stream runs on into the next step's copies
            st = streams[(NCH - 1) % 2]
            with torch.cuda.stream(st):
                for c in range(NCH - 1):
                    st.wait_event(ch[c]["loss_ev"][par_])
                torch.add(ch[0]["loss"][par_], ch[-1]["loss"][par_] if NCH > 1 else 0.0, out=loss_tot[par_])
                for c in range(1, NCH - 1):
                    loss_tot[par_].add_(ch[c]["loss"][par_])
                dist.all_reduce(loss_tot[par_])
                loss_pin[:1].copy_(loss_tot[par_], non_blocking=True)
                red_done[par_].record(st)

    def e2e_region(nsteps, entry):
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cur)
        for _ in range(nsteps):
            step_e2e(entry)
        for st in streams:
            cur.wait_stream(st)

    ms_e2e = {}
    for entry in ("params", "params+data", "raw"):
        e2e_count[0] = 0
        e2e_region(2, entry)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        e2e_region(K, entry)
        e3.record()
        barrier()
        ms_e2e[entry] = e2.elapsed_time(e3)
    clocks = clk.stop(local) if clk else None

    # ---- leg 3: sustained -- the device-resident step back to back for >= 2 s, its own clock record
    sustained = None
    if not args.no_sustained:
        # the step count must be THE SAME on every rank (each step all-reduces the loss): the slowest rank's timing decides, not this
        # rank's own (ranks whose ceil() differed by one step left the others waiting in a collective forever)
        from tsadar_b200.parallel import agreed_step_count
        n_sus = agreed_step_count(ms_total / K, args.sustained_seconds, K, device=dev)
        clk2 = ClockSampler("sustained").start() if rank == 0 else None
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record()
        for _ in range(n_sus):
            step(params_d, fe_d)
        s1.record()
        barrier()
        ms_sus = s0.elapsed_time(s1)
        sustained = {"steps": n_sus, "ms": ms_sus, "clocks": clk2.stop(local) if clk2 else None}

    t = torch.tensor([ms_total, ms_e2e["params"], ms_e2e["params+data"], ms_e2e["raw"], tf, tb, sustained["ms"] if sustained else 0.0],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e_p, ms_e2e_pd, ms_e2e_r, tf, tb, ms_sus = [float(x) for x in t.cpu()]

    # ---- the reference's named decks (every rank takes part: arts-2d is wavelength-sharded over the ranks)
    configs = None
    if not args.no_configs:
        try:
            from tools.bench_configs import run_named_configs
            configs = run_named_configs(rank, world, dev, cpu=(rank == 0 and world == 1 and not args.no_cpu_baseline))
        except Exception as exc:   # the headline line must not depend on the side block
            configs = {"error": f"{type(exc).__name__}: {exc}"}
        barrier()

    if rank == 0:
        ffma_peak = microbench(0)   # FFMA/s  (x2 = FLOP/s)
        mufu_peak = microbench(1)   # MUFU op/s
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        value = B * world * K / (ms_total * 1e-3)
        pairs = B * PAIRS_PER_LINEOUT
        peak_tf = 2 * ffma_peak / 1e12
        nominal_tf = n_sm * 128 * 2 * 1.965e9 / 1e12
        sm_mhz = (clocks or {}).get("sm_mhz")
        prof, stale = load_counts()
        kf = kernel_view(prof, "k_direct_fwd", B, tf, ffma_peak, sm_mhz, n_sm)
        kb = kernel_view(prof, "k_pv_nodes", B, tb, ffma_peak, sm_mhz, n_sm)
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        alg_bytes = B * (V_SYN * 4 + W_SYN * 8 * 2 + NP * 8)
        def e2e_block(ms, h2d, d2h, entry):
            return {"value": B * world * K / (ms * 1e-3), "unit": "lineouts/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms / K, "h2d_gbs_per_gpu": h2d / (ms / K * 1e-3) / 1e9,
                    "bound": "host<->device copies" if ms > 1.03 * ms_total else "kernels (copies hidden)", "entry": entry,
                    "pipeline": f"{NCH} chunks alternating on 2 streams; loss all-reduced every step",
                    "host_cpus_bound_to_gpu_numa_node": numa_cpus}
        e2e = e2e_block(ms_e2e_p, x_pin.numel() * 8, xbar_pin.numel() * 8 + 8 * NCH,
                        "LossFunction.vg_loss-style fit step: normalised leaves [B,6] f64 in from pinned host memory (the data batch is "
                        "resident on the device, fit.batch_to_device); tsff_params_fwd (ThomsonParams + DLM1V on the device) -> ff fwd -> "
                        "loss -> ff bwd -> tsff_params_bwd; loss + d loss / d leaves [B,6] out")
        e2e_pd = e2e_block(ms_e2e_pd, x_pin.numel() * 8 + tgt_pin.numel() * 8, xbar_pin.numel() * 8 + 8 * NCH,
                           "the same with the data batch [B,1024] f64 re-sent from pinned host memory every step, as the reference's "
                           "vg_loss(weights, batch) call does")
        e2e_raw = e2e_block(ms_e2e_r, params_pin.numel() * 8 + fe_pin.numel() * 4, pbar_pin.numel() * 8 + fbar_pin.numel() * 4 + 8 * NCH,
                            "raw C-ABI operands: params [B,14] f64 + fe [B,4096] f32 in; loss, params_bar, fe_bar out")
        roofline = {
            "bound": "fp32", "kernel": "k_direct_fwd<2,float,FP32,3> (pole sweep: I and dI/dxi, block-multipole form)",
            "achieved": kf["fp32_tflops"] if kf else None, "peak": peak_tf, "unit": "TFLOP/s", "frac": kf["frac"] if kf else None,
            "traffic": kf["dram_bytes"] if kf else None, "ms_per_launch": tf,
            "definition": "achieved = FP32 flop this kernel EXECUTES per launch (2 FFMA + 4 FFMA2 + FMUL + FADD + 2 FMUL2 + 2 FADD2 thread "
                          "instructions, ncu source page of this binary, x lineouts) / its CUDA-event duration in this run; peak = FFMA "
                          "microbenchmark in this run",
            "peak_source": "FFMA microbenchmark in this run (MEASURED_PEAKS.json has no FP32 entry); nominal %.1f" % nominal_tf,
            "counts_source": (prof or {}).get("source", "none"), "counts_stale": stale,
            "issue_slot_frac": kf.get("issue_slot_frac") if kf else None,
            "fp64_tflops_executed": kf["fp64_tflops"] if kf else None, "mufu_gops_executed": kf["mufu_gops"] if kf else None,
            "mufu_peak_gops": mufu_peak / 1e9,
            "pairwise_equivalent_tflops": FLOP_PER_PAIR_FWD * pairs / (tf * 1e-3) / 1e12,
            "pairwise_equivalent_frac": FLOP_PER_PAIR_FWD * pairs / (tf * 1e-3) / 1e12 / peak_tf,
            "pairwise_equivalent_note": "what an all-pairs sweep (12 flop per (omega,v) pair, SURVEY 8d) would need to match this time; "
                                        "NOT a roofline fraction -- the block-multipole form executes ~15x fewer instructions",
            "adjoint_kernel": "k_pv_nodes", "adjoint_ms_per_launch": tb,
            "adjoint_achieved": kb["fp32_tflops"] if kb else None, "adjoint_frac": kb["frac"] if kb else None,
            "adjoint_issue_slot_frac": kb.get("issue_slot_frac") if kb else None,
            "step_pairwise_equivalent_tflops": value / world * PAIRS_PER_LINEOUT * FLOP_PER_PAIR_STEP / 1e12,
            "hbm_algorithmic_bytes_per_launch": alg_bytes, "hbm_achieved_gbs": alg_bytes / (tf * 1e-3) / 1e9,
            "hbm_traffic_gbs": (kf["dram_bytes"] / (tf * 1e-3) / 1e9) if kf and kf["dram_bytes"] else None,
            "hbm_peak_gbs": hbm_peak, "hbm_frac": alg_bytes / (tf * 1e-3) / 1e9 / hbm_peak, "hbm_peak_source": hbm_src,
        }
        line = {
            "metric": METRIC, "value": value, "unit": "lineouts/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 PV sweeps / f64 assembly", "data": "synthetic",
            "config": workload_config(B, world),
            "e2e": e2e, "e2e_with_data_batch": e2e_pd, "e2e_raw_tables": e2e_raw,
            "gpu_launches": launches_per_step * K,
            "clocks": clocks,
            "roofline": roofline,
        }
        if sustained:
            line["sustained"] = {"value": B * world * sustained["steps"] / (ms_sus * 1e-3), "unit": "lineouts/s", "steps": sustained["steps"],
                                 "seconds": ms_sus * 1e-3, "clocks": sustained["clocks"]}
        if configs is not None:
            line["configs"] = configs
        # ---- parity of two lineouts of the timed batch against the float64 oracle, and the CPU baseline on the same sample
        from oracle import np_oracle as O
        grids = O.Grids(list(LAM_RANGE), W_SYN)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        keep = {}
        oracle_step(params_h[par_idx], fe_h[par_idx], vx, got["target"], grids, keep=keep)
        sp_mx = sp_pw = gp = gf = 0.0
        for i in range(len(par_idx)):
            ref = keep["modl"][i]
            sp_mx = max(sp_mx, float(np.abs(got["modl"][i] - ref).max() / np.abs(ref).max()))
            m = np.abs(ref) >= 1e-6 * np.abs(ref).max()
            sp_pw = max(sp_pw, float((np.abs(got["modl"][i] - ref)[m] / np.abs(ref)[m]).max()))
            rp = keep["pbar"][i] / (B * world)         # the oracle's loss is per lineout; the kernel's carries 1/B_total
            for k in (0, 1, 2, 11, 12):
                gp = max(gp, float(abs(got["pbar"][i][k] - rp[k]) / abs(rp[k])))
            rf = keep["fbar"][i] / (B * world)
            gf = max(gf, float(np.abs(got["fbar"][i] - rf).max() / np.abs(rf).max()))
        line["parity"] = {"lineouts_checked": par_idx, "spectrum_rel": sp_mx, "spectrum_pointwise_rel": sp_pw, "grad_rel": gp,
                          "fe_bar_rel": gf, "bars": {"spectrum": 1e-5, "grad": 1e-4},
                          "ok": bool(sp_mx <= 1e-5 and gp <= 1e-4 and gf <= 1e-4),
                          "oracle": "oracle/np_oracle.py + torch_oracle.py (float64, all pairs, autograd VJP of the same l2 loss)"}
        if not args.no_cpu_baseline and world == 1:
            nb = min(args.cpu_lineouts, B)
            tgt = target[:nb].cpu().numpy()
            t0 = time.perf_counter()
            oracle_step(params_h[:nb], fe_h[:nb], vx, tgt, grids)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nb / dt, "unit": "lineouts/s", "cores": cores, "kind": "port",
                                    "sample": f"the first {nb} lineouts of the timed batch (seed 42), torch-f64 forward + autograd VJP "
                                              f"(oracle/torch_oracle.py); --impl reference times the same sample per step"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
