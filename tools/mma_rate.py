#!/usr/bin/env python
"""tcgen05.mma kind::tf32 M=128 N=128 K=8 issue rate on resident operands (debug aid)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bgflow_b200 import _lib

lib = _lib.load()
dev = "cuda:0"
K = 128
A = torch.randn(128, K, device=dev)
W = torch.randn(128, K, device=dev)
scratch = torch.empty(128 * K, device=dev)
out = torch.empty(128, 128, device=dev)
for mode, name in ((0, "tf32 SS (A in smem)"), (1, "tf32 TS (A in TMEM)"), (3, "bf16 TS (A in TMEM)")):
    for reps in (16, 64, 256):
        st = torch.zeros(4, dtype=torch.int32, device=dev)
        rc = lib.bgx_tc_selftest(mode | (reps << 4), A.data_ptr(), W.data_ptr(), K, scratch.data_ptr(), out.data_ptr(),
                                 st.data_ptr(), None)
        torch.cuda.synchronize()
        n = reps * K // (16 if mode == 3 else 8)
        print(f"  bulk copies: {int(st[3])} x 16 KB landed after {int(st[2])} cycles")
        print(f"{name}: {n} MMAs in {int(st[1])} cycles -> {int(st[1]) / n:.1f} cycles/MMA (timeout={int(st[0])})")

print("--- bf16 TS variants (2048 MMAs)")
for label, bits in (("baseline", 0), ("commit every 12", 1 << 28), ("alternate A cols", 1 << 29),
                    ("concurrent tcgen05.ld", 1 << 30), ("all three", 7 << 28)):
    st = torch.zeros(4, dtype=torch.int32, device=dev)
    reps = 256
    word = 3 | (reps << 4) | bits
    if word >= 1 << 31:
        word -= 1 << 32
    rc = lib.bgx_tc_selftest(word, A.data_ptr(), W.data_ptr(), K, scratch.data_ptr(), out.data_ptr(), st.data_ptr(), None)
    torch.cuda.synchronize()
    n = reps * K // 16
    print(f"{label:24s}: {int(st[1]) / n:.1f} cycles/MMA (rc={rc}, timeout={int(st[0])})")

print("--- production issue pattern (warp-wide loop, elected lane, 3 products per k-step), 12 MMAs per rep")
for mode, name, cols in ((4, "bf16x3 TS N=128", 128), (5, "bf16x3 TS N=256", 256)):
    for label, bits in (("no commits", 0), ("commit every 24 MMAs", 1 << 28)):
        st = torch.zeros(4, dtype=torch.int32, device=dev)
        reps = 512
        word = mode | (reps << 4) | bits
        rc = lib.bgx_tc_selftest(word, A.data_ptr(), W.data_ptr(), K, scratch.data_ptr(), out.data_ptr(), st.data_ptr(), None)
        torch.cuda.synchronize()
        n = reps * 12
        cyc = int(st[1]) / n
        print(f"{name:18s} {label:22s}: {cyc:6.1f} cycles/MMA = {cyc * 128 / cols:6.1f} per 128x128x16 (rc={rc}, timeout={int(st[0])})")

print("--- SS mode (both operands from shared memory: the weight-gradient kernel bgx_gemm_tn) against TS mode, same issue pattern")
for mode, name in ((4, "bf16x3 TS N=128 (A in TMEM)"), (6, "bf16x3 SS N=128 (A in smem)")):
    st = torch.zeros(4, dtype=torch.int32, device=dev)
    reps = 512
    rc = lib.bgx_tc_selftest(mode | (reps << 4), A.data_ptr(), W.data_ptr(), K, scratch.data_ptr(), out.data_ptr(), st.data_ptr(), None)
    torch.cuda.synchronize()
    n = reps * 12
    print(f"{name:30s}: {int(st[1]) / n:6.1f} cycles per 128x128x16 MMA (rc={rc}, timeout={int(st[0])})")

print("--- does tcgen05.ld share the TMEM read port with the A-operand fetch of TS-mode MMAs?  (4 warps loop on tcgen05.ld.x32 of the accumulator)")
for mode, name, cols in ((4, "bf16x3 TS N=128", 128), (5, "bf16x3 TS N=256", 256)):
    for label, bits in (("alone", 0), ("+ concurrent tcgen05.ld", 1 << 30)):
        st = torch.zeros(4, dtype=torch.int32, device=dev)
        reps = 512
        word = mode | (reps << 4) | bits
        rc = lib.bgx_tc_selftest(word, A.data_ptr(), W.data_ptr(), K, scratch.data_ptr(), out.data_ptr(), st.data_ptr(), None)
        torch.cuda.synchronize()
        n = reps * 12
        cyc = int(st[1]) / n
        extra = f"; tcgen05.ld moved {int(st[3]) * 4 * 4096 / max(int(st[1]), 1):5.1f} B/cycle meanwhile" if bits else ""
        print(f"{name:18s} {label:26s}: {cyc:6.1f} cycles/MMA = {cyc * 128 / cols:6.1f} per 128x128x16 (rc={rc}, timeout={int(st[0])}){extra}")
