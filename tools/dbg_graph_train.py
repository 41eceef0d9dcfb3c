#!/usr/bin/env python
"""Debug aid: the bench's KL step (dim 66, 128-wide conditioners) as a replayed CUDA graph, small enough for
compute-sanitizer."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bgflow_b200.distributed import BucketedGradReducer, GraphedStep

dev = torch.device("cuda:0")
rows = int(os.environ.get("ROWS", 4096))
nblk = int(os.environ.get("BLOCKS", 2))
kind, dim, _, hidden, _, _ = bench.WORKLOADS["ala2_spline_d66_8blk"]
flow = bench.build_flow(kind, dim, nblk, hidden, dev)
red = BucketedGradReducer(flow)
opt = torch.optim.Adam(flow.parameters(), lr=1e-5, capturable=True)


def step():
    red.zero_grad()
    z = torch.rand(rows, dim, device=dev)
    x, dlogp = flow(z)
    loss = (0.5 * ((x - 0.5) / 0.25).square().sum(-1, keepdim=True) - dlogp).mean()
    loss.backward()
    red.finish()
    opt.step()
    return loss.detach()


with GraphedStep(step, list(flow.parameters()), warmup=2) as gs:
    for i in range(6):
        out = gs()
        torch.cuda.synchronize()
        print(i, float(out), flush=True)
print("ok")
