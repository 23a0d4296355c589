#!/usr/bin/env python
"""Table-mode fwd + VJP timings over batch sizes / ion counts of the 1d deck shape (scratch tool): python tools/bench_series.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from bench_configs import timeit, row, P9
from tsadar_b200.engine import FormFactorEngine
from tsadar_b200.synthetic import vgrid, super_gaussian_projected
dev = torch.device("cuda")
for B, nI in ((2, 1), (8, 1), (80, 2), (80, 1), (256, 1), (1024, 1), (1024, 2)):
    vx = vgrid(320)
    fe = torch.tensor(np.tile(super_gaussian_projected(vx, 2.5), (B, 1)), device=dev)
    engE = FormFactorEngine((319.7, 739.6), 5120, 0.0, P9, np.full(10, 0.1), 1, nI, vx, mode="table")
    engI = FormFactorEngine((523.1, 530.0), 5120, 0.0, P9, np.full(10, 0.1), 1, nI, vx, mode="table")
    pr = torch.tensor(row(B, nI), device=dev)
    cot = torch.randn(B, 5120, dtype=torch.float64, device=dev)
    def step():
        for e in (engE, engI):
            modl, _, saved = e.forward(pr, fe)
            e.backward(pr, fe, saved, modl_bar=cot)
    print(f"1d B={B} I={nI}: {timeit(step):.3f} ms", flush=True)
