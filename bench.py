#!/usr/bin/env python
"""Benchmark of the coupling-flow hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's engine on N B200s
    python bench.py --impl reference ...                     # the reference's CPU path (oracle port)

One "step" = one pass of the flow (forward + log|det J|, the sampling direction) over the
per-GPU batch of synthetic inputs.  Headline workload: the alanine-dipeptide RQ-spline stack of
BASELINE config 3 (D=66 split 33/33, 8 couplings, 8 bins, conditioner 33-128-128-825 SiLU) at
batch = 2^20 per GPU.  Multi-GPU = the same per-GPU batch on every rank (weak scaling, no
collective on the sample path).  Prints ONE JSON line on rank 0.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DRYRUN = os.environ.get("BGX_BENCH_DRYRUN") == "1"

def flops_bytes(kind, dim, hidden):
    """Algorithmic GEMM FLOPs and HBM bytes per sample per coupling block: FLOPs = 2 * sum(in * out)
    over the conditioner layers (x2 nets for affine), bytes = 4 * (D_c + 2 D_t + 2)."""
    d_c, d_t = dim // 2, dim - dim // 2
    dims = [d_c, *hidden, d_t * 25 if kind == "spline" else d_t]
    fl = 2 * sum(a * b for a, b in zip(dims[:-1], dims[1:])) * (1 if kind == "spline" else 2)
    return fl, 4 * (d_c + 2 * d_t + 2)


def _wl(kind, dim, n_blocks, hidden):
    return (kind, dim, n_blocks, hidden, *flops_bytes(kind, dim, hidden))


WORKLOADS = {
    # name: (kind, dim, n_blocks, hidden, flops/sample/block, algorithmic bytes/sample/block)
    "ala2_spline_d66_8blk": _wl("spline", 66, 8, (128, 128)),
    "ala2_affine_d66_8blk": _wl("affine", 66, 8, (128, 128, 128)),
    # BASELINE config 5 roofline sweep
    "spline_d384_8blk": _wl("spline", 384, 8, (128, 128)),
    "spline_d3072_8blk": _wl("spline", 3072, 8, (128, 128)),
    "affine_d384_8blk": _wl("affine", 384, 8, (128, 128, 128)),
    "affine_d3072_8blk": _wl("affine", 3072, 8, (128, 128, 128)),
}
assert WORKLOADS["ala2_spline_d66_8blk"][4:] == (252416, 404) and WORKLOADS["ala2_affine_d66_8blk"][4:] == (164864, 404)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_flow(kind, dim, n_blocks, hidden, device):
    """BASELINE configs 2/3: default nn.Linear init under torch.manual_seed(0)."""
    import bgflow_b200 as bg
    torch.manual_seed(0)
    d0 = dim // 2
    d1 = dim - d0
    layers = [bg.SplitFlow(d0)]
    for i in range(n_blocks):
        d_c, d_t = (d0, d1) if i % 2 == 0 else (d1, d0)
        if kind == "spline":
            tr = bg.ConditionalSplineTransformer(
                bg.DenseNet([d_c, *hidden, d_t * 25], activation=torch.nn.SiLU()), is_circular=False)
        else:
            tr = bg.AffineTransformer(
                shift_transformation=bg.DenseNet([d_c, *hidden, d_t], activation=torch.nn.ReLU()),
                scale_transformation=bg.DenseNet([d_c, *hidden, d_t], activation=torch.nn.ReLU()))
        layers += [bg.CouplingFlow(tr), bg.SwapFlow()]
    layers.append(bg.MergeFlow(d0))
    return bg.SequentialFlow(layers).to(device)


def oracle_blocks_from(flow):
    """The same parameters as oracle block dicts (CPU) for the reference / cpu_baseline legs."""
    import bgflow_b200 as bg
    from oracle import flows as of
    blocks = []
    for m in flow:
        if not isinstance(m, bg.CouplingFlow):
            continue
        t = m.transformer

        def mlp(net, act):
            lin = [l for l in net._layers if isinstance(l, torch.nn.Linear)]
            return of.MLP([l.weight.detach().cpu().float() for l in lin],
                          [l.bias.detach().cpu().float() for l in lin], act)
        if isinstance(t, bg.ConditionalSplineTransformer):
            blocks.append({"kind": "spline", "params_net": mlp(t._params_net, "silu")})
        else:
            blocks.append({"kind": "affine", "shift": mlp(t._shift_transformation, "relu"),
                           "scale": mlp(t._scale_transformation, "relu"),
                           "log_alpha": float(t._log_alpha.detach().cpu())})
    return blocks


def best_cpu_threads(blocks, kind, dim):
    """ATen's intra-op pool is slower with 128 threads than with 32 on these small ops: give the
    reference arm the thread count it is fastest with (probed on a 8192-row sample)."""
    from oracle import flows as of
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    z = torch.rand(8192, dim) if kind == "spline" else torch.randn(8192, dim)
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            of.coupling_stack(blocks, z, dim // 2)
            t0 = time.perf_counter()
            of.coupling_stack(blocks, z, dim // 2)
            t = time.perf_counter() - t0
            if t < best_t:
                best, best_t = c, t
    return best


def time_cpu_port(blocks, kind, dim, rows, reps, threads):
    """The reference's CPU PyTorch path, restated (oracle/flows.py), fp32, no_grad."""
    from oracle import flows as of
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1)
    z = torch.rand(rows, dim, generator=g) if kind == "spline" else torch.randn(rows, dim, generator=g)
    times = []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            of.coupling_stack(blocks, z, dim // 2)
            times.append(time.perf_counter() - t0)
    return times


def time_gpu_torch_port(blocks, kind, dim, rows, dev):
    """The reference's single-GPU PyTorch path, restated: the same op-by-op ATen sequence as the
    CPU baseline (oracle/flows.py) on CUDA tensors, fp32, no_grad, CUDA events."""
    from oracle import flows as of
    dblocks = []
    for b in blocks:
        nb = dict(b)
        for key in ("shift", "scale", "params_net"):
            if nb.get(key) is not None:
                m = nb[key]
                nb[key] = of.MLP([w.to(dev) for w in m.weights], [x.to(dev) for x in m.biases], m.act, m.periodic)
        dblocks.append(nb)
    g = torch.Generator().manual_seed(1)
    z = (torch.rand(rows, dim, generator=g) if kind == "spline" else torch.randn(rows, dim, generator=g)).to(dev)
    with torch.no_grad():
        for _ in range(2):
            of.coupling_stack(dblocks, z, dim // 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            of.coupling_stack(dblocks, z, dim // 2)
        e1.record()
        torch.cuda.synchronize()
    return rows * 3 / (e0.elapsed_time(e1) * 1e-3)


def bench_sweep(args, dev, run_timed, world):
    """BASELINE config 5: fwd + inv (+ log|det J|) roofline sweep over dim in {66, 384, 3072}, spline and affine
    8-block stacks at batch 2^20 per GPU, every block on a tcgen05 kernel (asserted through bgx_kernel_count).
    One entry per (kind, dim, direction): samples/s, ms per block launch, algorithmic TFLOP/s and GB/s with their
    fractions of the measured peaks."""
    import gc
    import bgflow_b200 as bg
    from bgflow_b200 import _lib
    pk = peaks()
    out = []
    for name in ("ala2_spline_d66_8blk", "spline_d384_8blk", "spline_d3072_8blk", "ala2_affine_d66_8blk",
                 "affine_d384_8blk", "affine_d3072_8blk"):
        kind, dim, n_blocks, hidden, flops_sb, bytes_sb = WORKLOADS[name]
        B = args.sweep_batch if dim <= 1024 else min(args.sweep_batch, args.sweep_batch_wide)
        flow = build_flow(kind, dim, n_blocks, hidden, dev)
        gen = torch.Generator(device=dev).manual_seed(1)
        z = (torch.rand(B, dim, generator=gen, device=dev) if kind == "spline"
             else torch.randn(B, dim, generator=gen, device=dev))
        with torch.no_grad():
            for direction in ("forward", "inverse"):
                inv = direction == "inverse"
                fn = (lambda: flow(z, inverse=True)) if inv else (lambda: flow(z))
                before = _lib.kernel_counts()
                for _ in range(2):
                    fn()
                steps = 3 if dim > 1024 else 5
                ms = run_timed(fn, steps) / steps
                used = {k: v - before[k] for k, v in _lib.kernel_counts().items() if v != before[k]}
                t_blk = ms * 1e-3 / n_blocks          # whole pass / blocks (split + merge copies included)
                out.append({"workload": name, "kind": kind, "dim": dim, "direction": direction, "batch_per_gpu": B,
                            "samples_per_s": B * world / (ms * 1e-3), "ms_per_pass": ms,
                            "tflops_algorithmic": flops_sb * B / t_blk / 1e12,
                            "tensor_frac": flops_sb * B / t_blk / 1e12 / pk["tf_sustained"],
                            "hbm_gbs_algorithmic": bytes_sb * B / t_blk / 1e9,
                            "hbm_frac": bytes_sb * B / t_blk / 1e9 / pk["hbm_gbs"],
                            "kernels": sorted(used)})
        del flow, z
        gc.collect()
        torch.cuda.empty_cache()
    return out



def bench_ic(args, dev):
    """Secondary measurement: the Z-matrix <-> Cartesian kernels on Ala2 (532 algorithmic bytes/sample)."""
    import bgflow_b200 as bg
    from bgflow_b200 import fixtures as oic
    B = args.batch_per_gpu
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    g = torch.Generator().manual_seed(1)
    xyz = (torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float32).reshape(1, -1)
           + 0.01 * torch.randn(B, 66, generator=g)).to(dev)
    out = {}
    pk = peaks()
    with torch.no_grad():
        ics = ic(xyz)[:-1]
        for name, fn in (("xyz_to_ic", lambda: ic(xyz)), ("ic_to_xyz", lambda: ic(*ics, inverse=True))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            gbs = 532 * B / (ms * 1e-3) / 1e9
            out[name] = {"samples_per_s": B / (ms * 1e-3), "ms": ms, "hbm_GBps_algorithmic": gbs,
                         "hbm_frac": gbs / pk["hbm_gbs"]}
    return out


def bench_tail(args, dev):
    """Secondary measurement: the builder tail on Ala2 — icdf maps of
    bonds / angles / torsions / augmented + IC -> Cartesian — as the reference structures it (one
    launch per field + the IC kernel) and fused (``fuse_domain_maps``: one multi-field launch + one
    mapped-IC kernel); the stand-alone CDF-map kernel over all five fields (640 algorithmic
    bytes/sample); the relative / mixed IC kernels."""
    import math
    import bgflow_b200 as bg
    from bgflow_b200 import fixtures as oic
    B = args.batch_per_gpu
    pk = peaks()
    one = lambda n, v=1.0: torch.full((n,), v, device=dev)
    inf = torch.tensor(math.inf, device=dev)
    m = {"bonds": bg.TruncatedNormalDistribution(one(21), one(21), torch.tensor(1e-5, device=dev), inf),
         "angles": bg.TruncatedNormalDistribution(one(20, 0.5), one(20), torch.tensor(1e-5, device=dev),
                                                  torch.tensor(1.0, device=dev)),
         "torsions": bg.SloppyUniform(torch.zeros(19, device=dev), one(19)),
         "fixed": torch.distributions.Normal(torch.zeros(9, device=dev), 20 * one(9)),
         "augmented": torch.distributions.Normal(torch.zeros(10, device=dev), one(10))}
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    layers = [bg.WrapFlow(bg.InverseFlow(bg.CDFTransform(m[n])), (i,))
              for i, n in enumerate(("bonds", "angles", "torsions", "augmented"))]
    layers += [bg.SetConstantFlow([4], [torch.zeros(1, 3, device=dev)]),
               bg.SetConstantFlow([5], [torch.tensor([0.5, 0.5, 0.5], device=dev)]),
               bg.WrapFlow(bg.InverseFlow(ic), indices=[0, 1, 2, 4, 5], out_indices=(0,))]
    tail = bg.SequentialFlow(layers).to(dev)
    fused = bg.fuse_domain_maps(tail)
    g = torch.Generator().manual_seed(1)
    us = [(torch.rand(B, w, generator=g) * 0.9 + 0.05).to(dev) for w in (21, 20, 19, 10)]
    u5 = us[:3] + [torch.rand(B, 9, generator=g).to(dev) * 0.9 + 0.05, us[3]]
    multi = bg.InverseFlow(bg.MultiCDFFlow([m[n] for n in ("bonds", "angles", "torsions", "fixed", "augmented")]))
    rel = bg.RelativeInternalCoordinateTransformation(oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK)
    x0 = torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float32).reshape(1, -1)
    xyz = (x0 + 0.01 * torch.randn(B, 66, generator=g)).to(dev)
    mixed = bg.MixedCoordinateTransformation(x0 + 0.02 * torch.randn(2000, 66, generator=g), oic.ALA2_RELATIVE_Z,
                                             oic.ALA2_RIGID_BLOCK, keepdims=9)
    out = {}

    def timed(name, fn, alg_bytes):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        gbs = alg_bytes * B / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "samples_per_s": B / (ms * 1e-3), "hbm_GBps_algorithmic": gbs,
                     "hbm_frac": gbs / pk["hbm_gbs"], "algorithmic_bytes_per_sample": alg_bytes}

    with torch.no_grad():
        timed("cdf_map_5_fields_79_cols", lambda: multi(*u5), 4 * (2 * 79 + 1))
        # tail: reads 70 uniforms, writes 66 coordinates + 10 augmented + dlogp
        timed("tail_per_field_launches", lambda: tail(*us), 4 * (70 + 76 + 1))
        timed("tail_fused", lambda: fused(*us), 4 * (70 + 76 + 1))
        rics = rel(xyz)[:-1]
        timed("relative_xyz_to_ic", lambda: rel(xyz), 4 * (2 * 66 + 1))
        timed("relative_ic_to_xyz", lambda: rel(*rics, inverse=True), 4 * (2 * 66 + 1))
        mics = mixed(xyz)[:-1]
        timed("mixed_xyz_to_ic", lambda: mixed(xyz), 4 * (66 + 60 + 1))
        timed("mixed_ic_to_xyz", lambda: mixed(*mics, inverse=True), 4 * (66 + 60 + 1))
    return out


def bench_config4(args, dev, rows=262144):
    """Secondary measurement: BASELINE config 4 end to end at batch 262144 — the builder-exact augmented
    Ala2 stack (tests/factory/test_generator_builder.py:45-66): uniform prior on (BONDS[21], ANGLES[20],
    TORSIONS[19, circular], AUGMENTED[10]) -> 14 multi-tensor spline couplings (WrapPeriodic conditioners on
    torsions) -> icdf maps of every field -> InverseFlow(GlobalIC) with constant origin / rotation."""
    import math
    import bgflow_b200 as bg
    from bgflow_b200 import fixtures as oic
    torch.manual_seed(0)
    width = {0: 21, 1: 20, 2: 19, 3: 10}

    def coupling(what, on):
        raw = sum(width[f] for f in on)
        periodic, col = [], 0
        for f in on:
            if f == 2:
                periodic += list(range(col, col + width[f]))
            col += width[f]
        circular = what == 2
        net = bg.DenseNet([raw + len(periodic), 128, 128, 24 * width[what] + (0 if circular else width[what])],
                          activation=torch.nn.SiLU())
        if periodic:
            net = bg.WrapPeriodic(net, left=0.0, right=1.0, indices=periodic)
        return bg.CouplingFlow(bg.ConditionalSplineTransformer(net, is_circular=circular),
                               transformed_indices=(what,), cond_indices=tuple(on))
    layers = []
    for _ in range(4):
        layers += [coupling(2, (3,)), coupling(3, (2,))]
    for _ in range(2):
        layers += [coupling(0, (1,)), coupling(1, (0,))]
    layers += [coupling(1, (2, 3)), coupling(0, (1, 2, 3))]
    one = lambda n, v=1.0: torch.full((n,), v, device=dev)
    marg = [bg.TruncatedNormalDistribution(one(21), one(21), torch.tensor(1e-5, device=dev), torch.tensor(math.inf, device=dev)),
            bg.TruncatedNormalDistribution(one(20, 0.5), one(20), torch.tensor(1e-5, device=dev), torch.tensor(1.0, device=dev)),
            bg.SloppyUniform(torch.zeros(19, device=dev), one(19)),
            torch.distributions.Normal(torch.zeros(10, device=dev), one(10))]
    layers += [bg.WrapFlow(bg.InverseFlow(bg.CDFTransform(m)), (i,)) for i, m in enumerate(marg)]
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    layers += [bg.SetConstantFlow([4], [torch.zeros(1, 3, device=dev)]),
               bg.SetConstantFlow([5], [torch.tensor([0.5, 0.5, 0.5], device=dev)]),
               bg.WrapFlow(bg.InverseFlow(ic), indices=[0, 1, 2, 4, 5], out_indices=(0,))]
    flow = bg.fuse_domain_maps(bg.SequentialFlow(layers).to(dev))
    g = torch.Generator().manual_seed(1)
    us = [torch.rand(rows, w, generator=g).to(dev) for w in (21, 20, 19, 10)]
    from bgflow_b200 import _lib
    with torch.no_grad():
        for _ in range(3):
            flow(*us)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            flow(*us)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    return {"samples_per_s": rows / (ms * 1e-3), "ms_per_pass": ms, "rows": rows,
            "kernel_launches_per_pass": (_lib.launch_count() - n0) // args.steps,
            "what": "uniform prior -> 14 multi-tensor spline couplings -> icdf maps -> global IC -> (xyz[66], aug[10]), "
                    "forward + log|det J|"}


def bench_train(args, flow, kind, dim, dev, run_timed, world, rows=65536):
    """BASELINE config 5, "KL-train grad allreduce": reverse-KL steps on a Gaussian target — kernel forward,
    recompute backward, per-coupling-block gradient buckets all-reduced over NCCL WHILE the backward of the earlier
    blocks still runs (``BucketedGradReducer``), Adam.  ``rows`` samples per rank per step.  The same step is also
    timed with the all-reduce after the backward and without any, which gives the exposed share of the collective."""
    from bgflow_b200.distributed import BucketedGradReducer
    red = BucketedGradReducer(flow)
    opt = torch.optim.Adam(flow.parameters(), lr=1e-5)

    def train_step(reduce=True):
        red.zero_grad()
        z = torch.rand(rows, dim, device=dev) if kind == "spline" else torch.randn(rows, dim, device=dev)
        x, dlogp = flow(z)
        loss = (0.5 * ((x - 0.5) / 0.25).square().sum(-1, keepdim=True) - dlogp).mean()
        loss.backward()
        if reduce:
            red.finish()
        opt.step()

    steps = max(3, args.steps // 2)
    res = {}
    from bgflow_b200 import engine as _engine
    gemm_modes = {}
    default_gm = _engine.backward_gemm_mode()
    for gm in ("fp32", "tf32", "bf16x3"):             # the other conditioner-backward variants (engine.config["backward_gemm"])
        old_gm = _engine.config["backward_gemm"]
        _engine.config["backward_gemm"] = gm
        red.overlap = True
        try:
            for _ in range(2):
                train_step()
            gemm_modes[gm] = run_timed(train_step, steps) / steps
        except Exception as exc:  # noqa: BLE001  (a torch build without mm(out_dtype=...) must not kill the bench)
            gemm_modes[gm] = f"unavailable: {type(exc).__name__}"
        _engine.config["backward_gemm"] = old_gm
    for mode in ("overlap", "after_backward", "no_allreduce"):
        red.overlap = mode == "overlap"
        fn = (lambda: train_step(False)) if mode == "no_allreduce" else train_step
        if mode == "no_allreduce":
            red.overlap = False
        for _ in range(3):
            fn()
        res[mode] = run_timed(fn, steps) / steps
    # the same step (default all-reduce policy) captured into ONE CUDA graph and replayed: what is left of the step
    # once the ~6 ms of Python that issue its ~300 kernels are gone
    graphed = None
    try:
        from bgflow_b200.distributed import GraphedStep
        red.overlap = "auto"
        opt = torch.optim.Adam(flow.parameters(), lr=1e-5, capturable=True)     # (train_step picks it up: no host syncs)
        with GraphedStep(train_step, list(flow.parameters()), warmup=2) as gs:
            for _ in range(4):
                gs()
            graphed = run_timed(gs, steps) / steps
    except Exception as exc:  # noqa: BLE001  (a capture that fails must not take the bench line with it)
        graphed = f"unavailable: {type(exc).__name__}: {exc}"[:200]
        torch.cuda.synchronize()
    red.remove()
    red.overlap = "auto"
    policy = "overlap" if red._overlapping() else "after_backward"      # the reducer's default for gradients of this size
    best = graphed if isinstance(graphed, float) else res[policy]
    out = {"samples_per_s": rows * world / (best * 1e-3), "ms_per_step": best,
           "issued": "one CUDA graph per step (distributed.GraphedStep)" if isinstance(graphed, float) else "eagerly from Python",
           "ms_per_step_eager": res[policy], "samples_per_s_eager": rows * world / (res[policy] * 1e-3), "rows_per_gpu": rows,
           "n_gpus": world, "allreduce_fp32_elems": red.n_elements, "allreduce_bytes": 4 * red.n_elements,
           "allreduce_mode": policy + " (BucketedGradReducer default for this gradient size)",
           "buckets": len(red.buckets), "ms_per_step_allreduce_overlapped": res["overlap"],
           "ms_per_step_allreduce_after_backward": res["after_backward"],
           "ms_per_step_no_allreduce": res["no_allreduce"], "ms_per_step_cuda_graph": graphed,
           "samples_per_s_cuda_graph": (rows * world / (graphed * 1e-3)) if isinstance(graphed, float) else None,
           "backward_gemm": default_gm,
           "ms_per_step_other_backward_gemm_modes": gemm_modes,
           "what": "fused-kernel forward; backward = conditioner re-run (bgx_linear) + bgx_spline_backward + input "
                   "gradients (bgx_linear) + weight / bias gradients (bgx_gemm_tn): own tcgen05 kernels, exact bf16 "
                   "operand splits, fp32 accumulation ('fp32' = the reference's semantics: torch autograd on cuBLAS "
                   "fp32 GEMMs); gradients of all blocks are views of one flat buffer: all-reduced with ONE NCCL "
                   "collective after the backward (default below 64 MB), or per coupling block from gradient hooks "
                   "while the rest of the backward runs (also timed: with 4 MB it loses, NCCL's CTAs take SMs from "
                   "the persistent one-CTA-per-SM kernels); Adam"}
    if world > 1:
        alone = res["after_backward"] - res["no_allreduce"]
        exposed = res["overlap"] - res["no_allreduce"]
        out["allreduce_ms_alone"] = alone                       # one collective after the backward
        out["allreduce_ms_exposed_when_overlapped"] = exposed   # per-block collectives launched from gradient hooks
        # share of the collective that overlapping hides; 0 when overlapping costs more than it hides (NCCL's CTAs
        # delay the persistent one-CTA-per-SM kernels: the reducer's default policy then reduces after the backward)
        out["overlap_fraction"] = max(0.0, 1.0 - exposed / alone) if alone > 1e-6 else None
    return out


_JSON_OUT = None


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    # stdout carries the ONE JSON line only: everything a library prints to file descriptor 1 (NCCL's version banner
    # under NCCL_DEBUG=VERSION, ...) goes to stderr; the line itself is written to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ala2_spline_d66_8blk", choices=sorted(WORKLOADS))
    ap.add_argument("--batch-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--cpu-sample-rows", type=int, default=65536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-full-pass", action="store_true",
                    help="reference arm: skip the single pass over the full per-GPU batch (about 20 s)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sweep-batch", type=int, default=1 << 20)
    ap.add_argument("--sweep-batch-wide", type=int, default=1 << 20, help="rows per GPU for the D = 3072 sweep points")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the KL training step (gradient all-reduce) measurement")
    ap.add_argument("--extras", action="store_true",
                    help="also time the IC kernels, the restated single-GPU PyTorch path and a KL training step "
                         "(adds to the JSON line)")
    args = ap.parse_args()
    if args.warmup < 3 and not DRYRUN:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kind, dim, n_blocks, hidden, flops_sb, bytes_sb = WORKLOADS[args.workload]
    metric = "flow samples/sec (fwd+log|detJ|)"
    config = {"workload": args.workload, "dim": dim, "couplings": n_blocks, "hidden": list(hidden),
              "n_bins": 8 if kind == "spline" else None, "batch_per_gpu": args.batch_per_gpu,
              "global_batch": args.batch_per_gpu * world, "direction": "forward (sampling)",
              "parallelism": f"batch-sharded x{world}, no collective on the sample path",
              "l2": "inputs (277 MB per step per GPU) exceed the 126 MB L2; no explicit flush"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        flow = build_flow(kind, dim, n_blocks, hidden, "cpu")
        blocks = oracle_blocks_from(flow)
        threads = best_cpu_threads(blocks, kind, dim)
        rows = args.cpu_sample_rows
        times = time_cpu_port(blocks, kind, dim, rows, args.warmup + args.steps, threads)[args.warmup:]
        total = sum(times)
        value = rows * len(times) / total
        sample = f"{rows} rows of the {args.workload} workload per step (chunk of the 2^20 batch)"
        # one pass over the FULL per-GPU batch of the b200 arm (2^20 rows as 16 chunks of 65536, what SURVEY 8d
        # prescribes when one call does not fit), so the two arms can be compared on the same configuration
        full = None
        if not args.no_cpu_full_pass:
            nchunk = max(1, args.batch_per_gpu // rows)
            t_full = sum(time_cpu_port(blocks, kind, dim, rows, nchunk, threads))
            full = {"rows": nchunk * rows, "chunks": nchunk, "seconds": t_full, "value": nchunk * rows / t_full,
                    "unit": "samples/s"}
        _emit(json.dumps({
            "impl": "reference", "metric": metric, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(config, cpu_rows_per_step=rows), "full_batch_pass": full,
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "host_cpus": os.cpu_count(),
                             "kind": "port", "sample": sample,
                             "note": "oracle/flows.py = op-for-op restatement of the reference's CPU PyTorch path "
                                     "(nflows spline restated); /root/reference is not on the GPU box"},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------ this repo's arm
    import torch.distributed as dist
    if world > 1:
        if DRYRUN:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    B = args.batch_per_gpu

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cpu" if DRYRUN else f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if DRYRUN:   # CPU plumbing test (tests/test_host_logic.py): sharding + barrier + max-over-ranks
        rows = [None] * world
        if world > 1:
            dist.barrier()
            dist.all_gather_object(rows, B)
        else:
            rows = [B]
        t0 = time.perf_counter()
        time.sleep(0.01 * (rank + 1))
        el = max_over_ranks(time.perf_counter() - t0)
        if rank == 0:
            _emit(json.dumps({"metric": metric, "value": B * world * args.steps / el, "unit": "samples/s",
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "dryrun": True,
                              "scaling": "weak", "shard_rows": rows, "config": config}))
        if world > 1:
            dist.destroy_process_group()
        return 0

    import bgflow_b200 as bg
    from bgflow_b200 import _lib, engine
    if os.environ.get("BGX_PRECISION"):
        engine.config["precision"] = os.environ["BGX_PRECISION"]
    if os.environ.get("BGX_FORCE_SIMT"):
        engine.config["force_simt"] = True
    if os.environ.get("BGX_SPLINE_KERNEL"):  # A/B switch
        engine.config["spline_kernel"] = os.environ["BGX_SPLINE_KERNEL"]
    config["kernel"] = ("fp32 SIMT" if engine.config["force_simt"] else
                        f"tcgen05 kind::f16, operands split into bf16 terms ({engine.config['precision']})")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    from bgflow_b200.host import HostPipeline, bind_to_gpu_numa, copy_ceiling
    numa = bind_to_gpu_numa(local_rank)      # before the pinned buffers exist: first touch lands on the GPU's node
    config["numa"] = numa
    flow = build_flow(kind, dim, n_blocks, hidden, dev)
    g = torch.Generator(device="cpu").manual_seed(1 + rank)
    z_host = (torch.rand(B, dim, generator=g) if kind == "spline" else torch.randn(B, dim, generator=g)).pin_memory()
    z = z_host.to(dev)
    couplings = [m for m in flow if isinstance(m, bg.CouplingFlow)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    with torch.no_grad():
        def step():
            return flow(z)
        for _ in range(args.warmup):
            step()
        # per-launch timing of the dominant kernel: events around every coupling call
        marks = []

        def pre(mod, inp):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append([e, None])

        def post(mod, inp, out):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks[-1][1] = e
        hooks = [h for m in couplings for h in (m.register_forward_pre_hook(pre), m.register_forward_hook(post))]
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = _lib.launch_count()
        ms = run_timed(step, args.steps)
        launches = _lib.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        for h in hooks:
            h.remove()
        kern_ms = [a.elapsed_time(b) for a, b in marks]
        value = B * world * args.steps / (ms * 1e-3)

        # ---- end to end: pinned host input -> device -> flow -> host result, every step
        e2e = None
        if not args.no_e2e:
            prior = (bg.UniformDistribution(torch.zeros(dim), torch.ones(dim)) if kind == "spline"
                     else bg.NormalDistribution(dim)).to(dev)
            pipe = HostPipeline(flow, dim_in=dim, dim_out=dim, max_rows=B, device=dev, prior=prior,
                                chunk_rows=int(os.environ.get("BGX_E2E_CHUNK", 0)) or None,     # default: 4 full waves
                                n_streams=int(os.environ.get("BGX_E2E_STREAMS", 3)))
            for _ in range(args.warmup):
                pipe.run(z_host)
            ms_e2e = run_timed(lambda: pipe.run(z_host), args.steps)
            h2d, d2h = B * dim * 4, B * dim * 4 + B * 4            # per GPU per step
            e2e = {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": "samples/s",
                   "h2d_bytes_per_step": world * h2d, "d2h_bytes_per_step": world * d2h,
                   "ms_per_step": ms_e2e / args.steps, "api": "bgflow_b200.host.HostPipeline.run (pinned host in/out)",
                   "chunk_rows": pipe.chunk, "streams": len(pipe.streams)}
            # what the host <-> device link gives THIS rank while every rank copies both ways at once
            barrier()
            cc = copy_ceiling(dev)
            barrier()
            t_copy = max(h2d / (cc["h2d_gbs_concurrent"] * 1e9), d2h / (cc["d2h_gbs_concurrent"] * 1e9))
            t_copy = max_over_ranks(t_copy)
            e2e["copy_ceiling"] = dict({k: round(v, 2) for k, v in cc.items()},
                                       note="pinned cudaMemcpyAsync, H2D and D2H concurrently, all ranks at once; GB/s of this rank",
                                       min_ms_per_step_from_copies=1e3 * t_copy,
                                       e2e_frac_of_copy_ceiling=1e3 * t_copy / (ms_e2e / args.steps))
            # the generator's sampling call: prior drawn on the device, nothing crosses PCIe host -> device
            for _ in range(args.warmup):
                pipe.sample(B)
            ms_s = run_timed(lambda: pipe.sample(B), args.steps)
            e2e["sample_call"] = {"value": B * world * args.steps / (ms_s * 1e-3), "unit": "samples/s",
                                  "h2d_bytes_per_step": 0, "d2h_bytes_per_step": world * d2h,
                                  "ms_per_step": ms_s / args.steps,
                                  "api": "HostPipeline.sample = BoltzmannGenerator.sample(n, with_dlogp=True): prior on the "
                                         "device (bg.py:115, normal.py:75-92), pinned host out",
                                  "frac_of_d2h_ceiling": (d2h / (cc["d2h_gbs"] * 1e9)) / (ms_s * 1e-3 / args.steps)}

    pk = peaks()
    traffic = None      # DRAM bytes per launch of the dominant kernel, from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    traffic_kernel = None
    if os.path.exists(tpath) and B == (1 << 20) and not engine.config["force_simt"] and dim == 66:
        tj = json.load(open(tpath)).get("spline_coupling_pair_kernel" if kind == "spline" else "affine_coupling_tc_kernel")
        if tj:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_kernel = tj["kernel"]
    avg_kern_s = 1e-3 * sum(kern_ms) / max(len(kern_ms), 1)
    roofline = {
        "bound": "tensor", "kernel": f"fused {kind} coupling block (conditioner GEMMs + transform + log-det)",
        "achieved": flops_sb * B / avg_kern_s / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
        "frac": flops_sb * B / avg_kern_s / 1e12 / pk["tf_sustained"], "traffic": traffic,
        "traffic_source": f"profiles/r2_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch of {traffic_kernel})",
        "peak_source": pk["source"] + " (dense bf16 cuBLAS, sustained); fp32-accurate math in the kernel",
        "avg_launch_ms": 1e3 * avg_kern_s, "launches_timed": len(kern_ms),
        "algorithmic_flops_per_launch": flops_sb * B, "algorithmic_bytes_per_launch": bytes_sb * B,
        "hbm": {"achieved": bytes_sb * B / avg_kern_s / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": bytes_sb * B / avg_kern_s / 1e9 / pk["hbm_gbs"]},
        "share_of_step": sum(kern_ms) / ms,
    }
    out = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "roofline": roofline,
           "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "impl": "b200"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        blocks = oracle_blocks_from(flow)
        threads = best_cpu_threads(blocks, kind, dim)
        rows = args.cpu_sample_rows
        times = time_cpu_port(blocks, kind, dim, rows, 3, threads)
        best = min(times[1:])
        out["cpu_baseline"] = {"value": rows / best, "unit": "samples/s", "cores": threads,
                               "host_cpus": os.cpu_count(), "kind": "port",
                               "sample": f"{rows} rows of the same workload, best of 2 after 1 warm-up"}
    if not args.no_sweep:
        sw = bench_sweep(args, dev, run_timed, world)
        if rank == 0:
            out["sweep"] = sw
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.extras:
        # cheap secondary lines every driver run carries (65536 rows): the restated single-GPU PyTorch path (the
        # north star's ">= 10x" denominator), the IC kernels and BASELINE config 4 end to end
        blocks = oracle_blocks_from(flow)
        out["secondary"] = {
            "gpu_pytorch_port": {"value": time_gpu_torch_port(blocks, kind, dim, 65536, dev), "unit": "samples/s",
                                 "sample": "65536 rows; oracle/flows.py op sequence on CUDA tensors"},
            "ic_ala2": bench_ic(args, dev), "config4_pipeline": bench_config4(args, dev)}
        out["secondary"]["speedup_vs_gpu_pytorch_port"] = value / out["secondary"]["gpu_pytorch_port"]["value"]
    if rank == 0 and world == 1 and args.extras:
        blocks = oracle_blocks_from(flow)
        out["extras"] = {
            "gpu_pytorch_port": {"value": time_gpu_torch_port(blocks, kind, dim, 65536, dev), "unit": "samples/s",
                                 "sample": "65536 rows; oracle/flows.py op sequence on CUDA tensors (the reference's "
                                           "single-GPU PyTorch path, restated)"},
            "ic_ala2": bench_ic(args, dev), "ic_tail": bench_tail(args, dev),
            "config4_pipeline": bench_config4(args, dev)}
    if not args.no_train:
        tr = bench_train(args, flow, kind, dim, dev, run_timed, world)
        if rank == 0:
            out["kl_train"] = tr
    if rank == 0:
        _emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
