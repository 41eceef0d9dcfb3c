#!/usr/bin/env python
"""Dump the in-kernel timeline of CTA 0 of the tensor-core spline kernel (debug aid).
    python tools/trace_tc.py [tiles_per_cta]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgflow_b200 as bg
from bgflow_b200 import _lib

ROLE = {0: "tma", 1: "mma", 2: "epi"}
EV = {(1, 5): "  issuing MMAs+commits (x64 cyc)", (1, 3): "  waited for weights (x64 cyc)", (1, 4): "  waited for accumulator (x64 cyc)", (1, 1): "operand ready -> issue layer", (1, 2): "issued (layer<16 / 16+chunk)",
      (2, 1): "hidden acc observed", (2, 2): "hidden handed over", (2, 3): "chunk acc observed",
      (2, 4): "chunk dims done", (2, 5): "next x staged", (2, 6): "y tile available"}


def main():
    tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dev = "cuda:0"
    lib = _lib.load()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B = int(os.environ.get('BGX_TC_MAX_CTAS', sms)) * 128 * tiles
    torch.manual_seed(0)
    tr = bg.ConditionalSplineTransformer(bg.DenseNet([33, 128, 128, 825], activation=torch.nn.SiLU())).to(dev)
    x = torch.rand(B, 33, device=dev)
    y = torch.rand(B, 33, device=dev)
    cap = 4096
    buf = torch.zeros(1 + 2 * cap, dtype=torch.int64, device=dev)
    with torch.no_grad():
        tr.forward(x, y)                       # warm-up (packs weights)
        torch.cuda.synchronize()
        lib.bgx_debug_set_trace(C.c_void_p(buf.data_ptr()), cap)
        tr.forward(x, y)
        torch.cuda.synchronize()
        lib.bgx_debug_set_trace(None, 0)
    t = buf.cpu().tolist()
    n = min(t[0], cap)
    ev = sorted((t[1 + 2 * i], t[2 + 2 * i]) for i in range(n))
    t0 = ev[0][0]
    last = {}
    for clk, code in ev:
        role, e, it, x_ = code >> 24, (code >> 16) & 0xff, (code >> 8) & 0xff, code & 0xff
        key = role
        dt = clk - last.get(key, clk)
        last[key] = clk
        print(f"{clk - t0:9d} (+{dt:7d}) {ROLE.get(role, role):4s} tile#{it} {str(EV.get((role, e), e)):32s} {x_}")


if __name__ == "__main__":
    main()
