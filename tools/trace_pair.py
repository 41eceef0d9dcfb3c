#!/usr/bin/env python
"""Per-role wait-time totals of CTA 0 of the pair kernel (bgx_debug_set_trace): where the MMA issuer, the weight
producer and an epilogue warp spend their cycles.  BGX_PAIR_EPW / BGX_PAIR_DEBUG select the variant."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgflow_b200 as bg
from bgflow_b200 import _lib, engine

dev = "cuda:0"
lib = _lib.load()
engine.config["spline_kernel"] = os.environ.get("BGX_SPLINE_KERNEL", "pair")
B = 1 << 20
torch.manual_seed(0)
tr = bg.ConditionalSplineTransformer(bg.DenseNet([33, 128, 128, 825], activation=torch.nn.SiLU())).to(dev)
x = torch.rand(B, 33, device=dev)
y = torch.rand(B, 33, device=dev)
buf = torch.zeros(64, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3):
        tr.forward(x, y)
    torch.cuda.synchronize()
    lib.bgx_debug_set_trace(C.c_void_p(buf.data_ptr()), 64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr.forward(x, y)
    e1.record()
    torch.cuda.synchronize()
    lib.bgx_debug_set_trace(None, 0)
t = buf.cpu().tolist()
n_pairs = max(t[4], 1)
ev = n_pairs * 18
print(f"launch {e0.elapsed_time(e1):.3f} ms; CTA 0: {n_pairs} tile pairs, {ev} (unit, slot) events")
print(f"MMA warp  : total {t[0]:9d} cycles ({t[0] / ev:7.0f} / event); waiting weights {t[1]:9d} ({100 * t[1] / max(t[0], 1):4.1f} %), "
      f"A operand {t[2]:9d} ({100 * t[2] / max(t[0], 1):4.1f} %), pulled accumulator {t[3]:9d} ({100 * t[3] / max(t[0], 1):4.1f} %)")
print(f"producer  : total {t[12]:9d} cycles; waiting for a free stage {t[13]:9d} ({100 * t[13] / max(t[12], 1):4.1f} %)")
print(f"epilogue 0: total {t[8]:9d} cycles; waiting for accumulators {t[9]:9d} ({100 * t[9] / max(t[8], 1):4.1f} %)")
