#!/usr/bin/env python
"""Condense ncu outputs brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches <launches.csv>            # per-kernel count / mean ns / share of the step
    python tools/ncu_summary.py raw <raw.csv>                      # the fixed metric list below, one block per kernel

<raw.csv> comes from `ncu -i X.ncu-rep --page raw --csv`; <launches.csv> from the
`--metrics gpu__time_duration.sum --csv --log-file` pass of /opt/skills/guides/B200_PROFILING.md.
"""
import collections
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
    "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    d = collections.OrderedDict()
    for r in rows[start + 1:]:
        rec = dict(zip(hdr, r))
        if rec.get("Metric Name") != "gpu__time_duration.sum":
            continue
        d.setdefault(rec["Kernel Name"].split("(")[0][-70:], []).append(float(rec["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print(f"{'kernel':70s} {'n':>5s} {'mean_us':>10s} {'share':>7s}")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:70s} {len(v):5d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / tot:7.3f}")
    print(f"total {tot / 1e6:.3f} ms over {sum(len(v) for v in d.values())} launches (ncu: serialised, cold cache)")


def raw(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== " + r[idx["Kernel Name"]].split("(")[0][-90:] + f"   [launch id {r[idx['ID']]}]")
        for m in METRICS:
            if m in idx and r[idx[m]] != "":
                print(f"  {m:80s} {r[idx[m]]:>18s} {units[idx[m]]}")
        print()


def tojson(path, kernel_regex, lineouts):
    """python tools/ncu_summary.py json <raw.csv> <kernel regex> <lineouts per launch>  -> JSON for bench.py's roofline"""
    import json, re
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if re.search(kernel_regex, r[idx["Kernel Name"]]):
            def val(m):
                v = float(r[idx[m]].replace(",", ""))
                u = units[idx[m]]
                return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
            out = {"kernel": r[idx["Kernel Name"]].split("(")[0][-60:], "lineouts_per_launch": int(lineouts),
                   "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                   "issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "fma_pipe_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                   "xu_pipe_pct": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                   "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                   "warp_instructions": val("smsp__inst_executed.sum"),
                   "duration_us_under_ncu": val("gpu__time_duration.sum") * (1e-3 if units[idx["gpu__time_duration.sum"]] == "ns" else (1e3 if units[idx["gpu__time_duration.sum"]] == "ms" else 1.0)),
                   "source": path}
            print(json.dumps(out, indent=1))
            return
    raise SystemExit("kernel not found")


if __name__ == "__main__":
    if sys.argv[1] == "json":
        tojson(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
