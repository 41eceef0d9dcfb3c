#!/usr/bin/env python
"""Where BASELINE config 4 (builder-exact augmented Ala2 stack -> icdf maps -> global IC, 262,144 rows) spends its GPU
time: torch.profiler over a few passes, top CUDA kernels by total time."""
import os
import sys
import types

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda:0")
args = types.SimpleNamespace(steps=3)
orig = bench.bench_config4


def run():
    return orig(args, dev)


run()       # builds, warms up
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
