#!/usr/bin/env python
"""Where a KL training step spends its GPU time: torch.profiler over a few steps of bench.py's train step
(8-block Ala2 spline stack, 65536 rows), top CUDA kernels by total time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bgflow_b200.distributed import BucketedGradReducer

from bgflow_b200 import engine
if os.environ.get("BGX_BACKWARD_GEMM"):
    engine.config["backward_gemm"] = os.environ["BGX_BACKWARD_GEMM"]
dev = torch.device("cuda:0")
kind, dim, n_blocks, hidden, _, _ = bench.WORKLOADS["ala2_spline_d66_8blk"]
flow = bench.build_flow(kind, dim, n_blocks, hidden, dev)
red = BucketedGradReducer(flow)
opt = torch.optim.Adam(flow.parameters(), lr=1e-5)
rows = 65536


def step():
    red.zero_grad()
    z = torch.rand(rows, dim, device=dev)
    x, dlogp = flow(z)
    loss = (0.5 * ((x - 0.5) / 0.25).square().sum(-1, keepdim=True) - dlogp).mean()
    loss.backward()
    red.finish()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
