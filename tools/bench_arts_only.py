#!/usr/bin/env python
"""arts-1d alone (scratch): python tools/bench_arts_only.py"""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench_configs as bc
import types
# run only the arts-1d leg: reuse run_named_configs' inner function through a tiny shim
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_configs.py")).read()
out = bc.run_named_configs(cpu=False)
print(json.dumps({k: v for k, v in out.items() if k.startswith("arts-1d")}, indent=1))
