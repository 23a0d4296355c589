#!/usr/bin/env python
"""Condense `ncu --page source --csv` (SASS view) into hot regions: consecutive instructions with a similar executed
count are merged into one line (count, #instr, share of all executed instructions, stall samples, opcode histogram).
    python tools/ncu_source.py <source.csv> [min_share]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].strip().isdigit(): continue
    ins.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)))
tot = sum(c for _, c, _ in ins); tots = sum(s for _, _, s in ins)
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
regions = []
cur = None
for k, (src, c, s) in enumerate(ins):
    if cur and c > 0 and 0.7 * cur["c"] <= c <= 1.43 * cur["c"]:
        cur["n"] += 1; cur["sum"] += c; cur["samples"] += s; cur["ops"][src.split()[0].split(".")[0] if not src.startswith("@") else src.split()[1].split(".")[0]] += 1; cur["end"] = k
    else:
        if cur: regions.append(cur)
        op = src.split()[0].split(".")[0] if not src.startswith("@") else src.split()[1].split(".")[0]
        cur = {"c": max(c, 1), "n": 1, "sum": c, "samples": s, "ops": collections.Counter({op: 1}), "start": k, "end": k}
regions.append(cur)
print(f"total warp-instructions {tot}, samples {tots}")
for r in regions:
    if r["sum"] / tot >= minshare:
        ops = " ".join(f"{o}:{n}" for o, n in r["ops"].most_common(8))
        print(f"[{r['start']:5d}-{r['end']:5d}] n={r['n']:4d} exec/instr={r['sum']//r['n']:9d} share={r['sum']/tot:6.3f} samples={r['samples']/max(tots,1):6.3f}  {ops}")
