#!/usr/bin/env python
"""Device times of the training-path kernels at the KL step's shapes (B = 65536, 128-wide conditioner, 825 spline
parameters): bgx_linear (recompute + input gradients), bgx_gemm_tn (weight gradients) against torch fp32 / TF32."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bgflow_b200 import engine  # noqa: E402

dev = "cuda:0"
B = int(os.environ.get("B", 65536))


def timeit(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def kernel_us(f, name, n=5):
    """Average device time of the kernels whose name contains ``name`` over n calls of f (torch profiler)."""
    if os.environ.get("BGX_NO_TORCH_PROFILER"):      # under ncu (both want CUPTI)
        for _ in range(1 + n):
            f()
        return float("nan")
    from torch.profiler import ProfilerActivity, profile
    f()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            f()
        torch.cuda.synchronize()
    tot = sum(e.device_time_total for e in prof.key_averages() if name in e.key)
    return tot / n


print(f"B = {B}   (event times include the Python / launch overhead of the call; 'kernel' = device time of our kernel alone)")
for k, n in ((128, 828), (828, 128), (128, 128), (16, 128)):
    x = torch.randn(B, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    b = torch.randn(n, device=dev)
    f = engine.LinearTC()
    t_own = timeit(lambda: f(x, w, b))
    torch.backends.cuda.matmul.allow_tf32 = False
    t_fp32 = timeit(lambda: torch.addmm(b, x, w.t()))
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf32 = timeit(lambda: torch.addmm(b, x, w.t()))
    torch.backends.cuda.matmul.allow_tf32 = False
    gb = B * (k + n) * 4 / 1e9
    t_k = kernel_us(lambda: f(x, w, b), "linear_tc_kernel")
    print(f"linear  K={k:4d} N={n:4d}: bgx_linear {t_own * 1e3:7.1f} us, kernel {t_k:7.1f} us ({gb / t_k * 1e6:6.0f} GB/s)   torch fp32 {t_fp32 * 1e3:7.1f} us   torch tf32 {t_tf32 * 1e3:7.1f} us")
# spline transform backward at the same shapes (reads P, writes dP: 2 x B x 828 floats)
from bgflow_b200 import _lib  # noqa: E402
for stride in (828, 825):
    d_t, nb = 33, 8
    P = torch.randn(B, stride, device=dev)
    y = torch.rand(B, d_t, device=dev)
    go = torch.randn(B, d_t, device=dev)
    gd = torch.randn(B, 1, device=dev)
    ecol = torch.arange(3 * nb * d_t, 3 * nb * d_t + d_t, dtype=torch.int32, device=dev)
    fn = lambda: engine.spline_backward(P, y, go, gd, ecol, nb)
    t_own = timeit(fn)
    t_k = kernel_us(fn, "spline_backward_kernel")
    gb = 2 * B * 825 * 4 / 1e9
    print(f"spline_backward stride={stride}: {t_own * 1e3:7.1f} us, kernel {t_k:7.1f} us ({gb / t_k * 1e6:6.0f} GB/s)")
for n, k, ld in ((825, 128, 828), (128, 128, 128), (128, 16, 128)):
    g = torch.randn(B, ld, device=dev)
    h = torch.randn(B, k, device=dev)
    t_own = timeit(lambda: engine.gemm_tn(g, h, n))
    gv = g[:, :n]
    t_fp32 = timeit(lambda: (gv.t() @ h, gv.sum(0)))
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf32 = timeit(lambda: (gv.t() @ h, gv.sum(0)))
    torch.backends.cuda.matmul.allow_tf32 = False
    gb = B * (n + k) * 4 / 1e9
    t_k = kernel_us(lambda: engine.gemm_tn(g, h, n), "gemm_tn_kernel")
    print(f"gemm_tn N={n:4d} K={k:4d}: bgx_gemm_tn (+ slice sum) {t_own * 1e3:7.1f} us, kernel {t_k:7.1f} us ({gb / t_k * 1e6:6.0f} GB/s)   torch fp32 mm + sum {t_fp32 * 1e3:7.1f} us   tf32 {t_tf32 * 1e3:7.1f} us")
