#!/usr/bin/env python
"""Executed-work counts per kernel launch from an ncu report captured with `--set full --import-source on`:

    python tools/ncu_flops.py <X.ncu-rep> [kernel regex] [--json lineouts_per_launch]

Reads `ncu -i X.ncu-rep --page source --csv` (per-SASS-instruction "Predicated-On Thread Instructions Executed") and sums
by opcode: FP32 flop = 2 FFMA + 4 FFMA2 + FMUL + FADD + 2 FMUL2 + 2 FADD2 (a packed f32x2 instruction does two lanes),
FP64 flop = 2 DFMA + DMUL + DADD, MUFU ops, and the warp-level instruction total (issue slots).  These counts are properties
of the binary and the inputs, not of the clock, so bench.py divides them by the kernel time IT measures to get the executed
fraction of the FP32 roof (roofline.frac) and the issue-slot utilisation."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

FP32 = {"FFMA": 2, "FFMA2": 4, "FMUL": 1, "FADD": 1, "FMUL2": 2, "FADD2": 2}
FP64 = {"DFMA": 2, "DMUL": 1, "DADD": 1}


def parse(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True, errors="replace").stdout
    kernels, cur, hdr = [], None, None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"kernel": row[1], "ops": collections.Counter(), "warp_inst": 0, "thread_inst": 0}
            kernels.append(cur)
            hdr = None
            continue
        if row[0] == "Address":
            hdr = {h: i for i, h in enumerate(row)}
            continue
        if cur is None or hdr is None or len(row) < len(hdr) - 2:
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", row[hdr["Source"]])
        if not m:
            continue
        op = m.group(1).split(".")[0]
        try:
            ti = int(row[hdr["Predicated-On Thread Instructions Executed"]])
            wi = int(row[hdr["Instructions Executed"]])
        except ValueError:
            continue
        cur["ops"][op] += ti
        cur["warp_inst"] += wi
        cur["thread_inst"] += ti
    for k in kernels:
        o = k["ops"]
        k["fp32_flop"] = sum(o[n] * f for n, f in FP32.items())
        k["fp64_flop"] = sum(o[n] * f for n, f in FP64.items())
        k["mufu"] = o["MUFU"]
    return kernels


def raw_metrics(rep):
    """kernel name -> {metric: value} of the first launch of each kernel (raw page; byte units normalised)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                         errors="replace").stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
    res = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if name in res:
            continue
        d = {}
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                  "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                  "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                  "launch__registers_per_thread"):
            if m in idx and r[idx[m]] != "":
                d[m] = float(r[idx[m]].replace(",", "")) * scale.get(units[idx[m]], 1.0)
        res[name] = d
    return res


def latest(rep, lineouts, source):
    """profiles/ncu_latest.json: per-lineout executed counts of the benchmark step's kernels, for bench.py's in-run roofline."""
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from tsadar_b200 import build as _b
    raw = raw_metrics(rep)
    out = {"source": source, "source_stamp": _b._stamp(), "lineouts_per_launch": lineouts, "kernels": {}}
    for k in parse(rep):
        for key in ("k_direct_fwd", "k_direct_step", "k_pv_nodes", "k_direct_bwd_poles", "k_direct_prep"):
            if key in k["kernel"] and key not in out["kernels"]:
                r = next((v for n, v in raw.items() if key in n), {})
                out["kernels"][key] = {
                    "name": k["kernel"].split("(")[0][-60:], "fp32_flop_per_lineout": k["fp32_flop"] / lineouts,
                    "fp64_flop_per_lineout": k["fp64_flop"] / lineouts, "mufu_per_lineout": k["mufu"] / lineouts,
                    "warp_inst_per_lineout": k["warp_inst"] / lineouts,
                    "dram_bytes_per_lineout": (r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)) / lineouts,
                    "under_ncu": {"duration_us": r.get("gpu__time_duration.sum"), "issue_slots_busy_pct": r.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                  "fma_pipe_pct": r.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                                  "xu_pipe_pct": r.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                                  "fp64_pipe_pct": r.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                                  "warps_active_pct": r.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                                  "registers": r.get("launch__registers_per_thread")},
                    "opcode_thread_inst_per_lineout": {o: c / lineouts for o, c in k["ops"].most_common(12)}}
    print(json.dumps(out, indent=1))


def main():
    rep = sys.argv[1]
    if "--latest" in sys.argv:
        i = sys.argv.index("--latest")
        return latest(rep, int(sys.argv[i + 1]), sys.argv[i + 2] if len(sys.argv) > i + 2 else rep)
    rx = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "."
    ks = [k for k in parse(rep) if re.search(rx, k["kernel"])]
    if "--json" in sys.argv:
        n = int(sys.argv[sys.argv.index("--json") + 1])
        k = ks[0]
        print(json.dumps({"kernel": k["kernel"].split("(")[0], "lineouts_per_launch": n, "fp32_flop_per_lineout": k["fp32_flop"] / n,
                          "fp64_flop_per_lineout": k["fp64_flop"] / n, "mufu_per_lineout": k["mufu"] / n,
                          "warp_inst_per_lineout": k["warp_inst"] / n,
                          "opcode_thread_inst_per_lineout": {o: c / n for o, c in k["ops"].most_common(14)}}, indent=1))
        return
    for k in ks:
        print("== " + k["kernel"].split("(")[0][-90:])
        print(f"  warp instructions {k['warp_inst']:.4g}   thread instructions {k['thread_inst']:.4g}")
        print(f"  FP32 flop {k['fp32_flop']:.4g}   FP64 flop {k['fp64_flop']:.4g}   MUFU {k['mufu']:.4g}")
        print("  top opcodes: " + ", ".join(f"{o} {c:.3g}" for o, c in k["ops"].most_common(14)))


if __name__ == "__main__":
    main()
