#!/usr/bin/env python
"""Time the IC-family kernels alone (bench.py's bench_ic / bench_tail) and print one JSON line.
BGX_IC_BULK=0 selects the element-wise staged tile I/O everywhere (A/B against the bulk-TMA tiles)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--batch-per-gpu", type=int, default=1 << 20)
args = ap.parse_args()
dev = torch.device("cuda:0")
out = {"ic_bulk": os.environ.get("BGX_IC_BULK", "1"), "ic_ala2": bench.bench_ic(args, dev), "ic_tail": bench.bench_tail(args, dev)}
print(json.dumps(out))
