#!/usr/bin/env python
"""The named-deck timings alone (bench.py's `configs` block without the headline sweep): python tools/bench_arts_only.py [key-prefix]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench_configs as bc

prefix = sys.argv[1] if len(sys.argv) > 1 else "arts-1d"
out = bc.run_named_configs(cpu=False)
print(json.dumps({k: v for k, v in out.items() if k.startswith(prefix)}, indent=1))
