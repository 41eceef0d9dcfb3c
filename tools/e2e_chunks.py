#!/usr/bin/env python
"""End-to-end (pinned host in -> flow -> pinned host out) samples/s of bench.py's default workload against the
HostPipeline chunk size and stream count."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import bgflow_b200 as bg
from bgflow_b200.host import HostPipeline, bind_to_gpu_numa, wave_rows

dev = torch.device("cuda:0")
print(bind_to_gpu_numa(0))
kind, dim, n_blocks, hidden, _, _ = bench.WORKLOADS["ala2_spline_d66_8blk"]
flow = bench.build_flow(kind, dim, n_blocks, hidden, dev)
B = 1 << 20
z = torch.rand(B, dim).pin_memory()
prior = bg.UniformDistribution(torch.zeros(dim), torch.ones(dim)).to(dev)
w = wave_rows(dev, 1)
for streams, graph in ((3, False), (3, True)):
    for chunk in (w, 2 * w, 3 * w, 4 * w):
        pipe = HostPipeline(flow, dim, dim, B, dev, chunk_rows=chunk, n_streams=streams, prior=prior, use_graph=graph)
        for _ in range(3):
            pipe.run(z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            pipe.run(z)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        for _ in range(3):
            pipe.sample(B)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            pipe.sample(B)
        e1.record()
        torch.cuda.synchronize()
        ms_s = e0.elapsed_time(e1) / 10
        print(f"graph {int(graph)} streams {streams} chunk {chunk:7d} ({chunk / w:5.2f} waves): run {ms:6.2f} ms = {B / ms / 1e3:6.1f} M/s   sample {ms_s:6.2f} ms = {B / ms_s / 1e3:6.1f} M/s", flush=True)
        del pipe
