"""Gradients through the fused coupling blocks (kernel forward, recompute backward) against the
oracle differentiated in fp64 on the CPU.  Tolerance: 2e-3 relative to the gradient scale."""

import numpy as np
import pytest
import torch

import bgflow_b200 as bg
from oracle import flows as of
from helpers import stack_from

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _oracle_grads(kind, blocks64, split, z, wx, wd, inverse):
    z = z.double().clone().requires_grad_(True)
    params = []
    for b in blocks64:
        for key in ("shift", "scale", "params_net"):
            if b.get(key) is not None:
                for t in b[key].weights + b[key].biases:
                    t.requires_grad_(True)
                    params.append(t)
    x, d = of.coupling_stack(blocks64, z, split, inverse=inverse)
    loss = (x * wx.double()).sum() + (d * wd.double()).sum()
    g = torch.autograd.grad(loss, [z, *params])
    return x.detach(), d.detach(), g[0], g[1:]


@pytest.mark.parametrize("kind", ["spline", "affine"])
@pytest.mark.parametrize("inverse", [False, True])
def test_input_and_parameter_gradients(kind, inverse):
    dim, nblk = 10, 2
    blocks, split = of.make_stack(kind, dim, nblk, hidden=(128, 128) if kind == "spline" else (24,), seed=2)
    blocks64, _ = of.make_stack(kind, dim, nblk, hidden=(128, 128) if kind == "spline" else (24,), seed=2,
                                dtype=torch.float64)
    flow = stack_from(blocks, split, DEV)
    g = torch.Generator().manual_seed(4)
    z = torch.rand(50, dim, generator=g) if kind == "spline" else torch.randn(50, dim, generator=g)
    wx, wd = torch.randn(50, dim, generator=g), torch.randn(50, 1, generator=g)
    x_ref, d_ref, gz_ref, gp_ref = _oracle_grads(kind, blocks64, split, z, wx, wd, inverse)

    zc = z.to(DEV).requires_grad_(True)
    x, d = flow(zc, inverse=inverse)
    assert x.requires_grad and d.requires_grad
    np.testing.assert_allclose(x.detach().cpu().double().numpy(), x_ref.numpy(), atol=2e-5, rtol=1e-4)
    loss = (x * wx.to(DEV)).sum() + (d * wd.to(DEV)).sum()
    loss.backward()
    scale = gz_ref.abs().max().item()
    np.testing.assert_allclose(zc.grad.cpu().double().numpy(), gz_ref.numpy(), atol=2e-3 * scale, rtol=2e-3)
    ours = []      # oracle order: per conditioner, all weights then all biases
    for m in flow.modules():
        if isinstance(m, bg.DenseNet):
            lin = [l for l in m._layers if isinstance(l, torch.nn.Linear)]
            ours += [l.weight.grad for l in lin] + [l.bias.grad for l in lin]
    assert len(ours) == len(gp_ref)
    for a, b in zip(ours, gp_ref):
        s = max(b.abs().max().item(), 1e-6)
        np.testing.assert_allclose(a.cpu().double().numpy(), b.numpy(), atol=3e-3 * s, rtol=3e-3)


@pytest.mark.parametrize("variant", ["shift_only", "scale_only", "preserve_volume", "shift_only_circular"])
@pytest.mark.parametrize("inverse", [False, True])
def test_affine_variants_backward(variant, inverse):
    """affine.py:41-47: NICE (no scale net: dlogp is a constant zero without a graph), scale-only,
    volume-preserving and circular shift-only transformers must all train (the recompute
    backward used to hand a graph-less dlogp to autograd.grad and raise)."""
    gen = torch.Generator().manual_seed(3)
    dim, split = 8, 4
    blocks, blocks64 = [], []
    for dt in (torch.float32, torch.float64):
        g2 = torch.Generator().manual_seed(11)
        out = blocks if dt == torch.float32 else blocks64
        for _ in range(2):
            b = {"kind": "affine", "log_alpha": -0.5,
                 "shift": of.make_mlp([4, 24, 4], "relu", g2, dt) if variant != "scale_only" else None,
                 "scale": of.make_mlp([4, 24, 4], "tanh", g2, dt) if variant in ("scale_only", "preserve_volume") else None,
                 "preserve_volume": variant == "preserve_volume", "is_circular": variant == "shift_only_circular"}
            out.append(b)
    flow = stack_from(blocks, split, DEV)
    z = torch.rand(37, dim, generator=gen)
    wx, wd = torch.randn(37, dim, generator=gen), torch.randn(37, 1, generator=gen)
    x_ref, d_ref, gz_ref, gp_ref = _oracle_grads("affine", blocks64, split, z, wx, wd, inverse)
    zc = z.to(DEV).requires_grad_(True)
    x, d = flow(zc, inverse=inverse)
    np.testing.assert_allclose(x.detach().cpu().double().numpy(), x_ref.numpy(), atol=2e-5, rtol=1e-4)
    ((x * wx.to(DEV)).sum() + (d * wd.to(DEV)).sum()).backward()
    s = max(gz_ref.abs().max().item(), 1e-6)
    np.testing.assert_allclose(zc.grad.cpu().double().numpy(), gz_ref.numpy(), atol=2e-3 * s, rtol=2e-3)
    ours = []
    for m in flow.modules():
        if isinstance(m, bg.DenseNet):
            lin = [l for l in m._layers if isinstance(l, torch.nn.Linear)]
            ours += [l.weight.grad for l in lin] + [l.bias.grad for l in lin]
    assert len(ours) == len(gp_ref)
    for a, b in zip(ours, gp_ref):
        sc = max(b.abs().max().item(), 1e-6)
        np.testing.assert_allclose(a.cpu().double().numpy(), b.numpy(), atol=3e-3 * sc, rtol=3e-3)


@pytest.mark.parametrize("kind", ["spline", "affine"])
@pytest.mark.parametrize("mode,tol", [("fp32", 3e-3), ("tf32", 3e-2), ("bf16x3", 3e-3), ("tcgen05", 3e-3)])   # (TF32: 2^-11 per product, ReLU kinks flip)
def test_backward_gemm_modes(kind, mode, tol):
    """engine.config["backward_gemm"]: the conditioner backward through torch autograd on cuBLAS fp32 or TF32 GEMMs, or
    written out explicitly with three bf16 tensor-core products of exact operand splits — on cuBLAS (``bf16x3``:
    ``_mlp_grad`` + ``bgx_split_bf16``) or on our own kernels (``tcgen05``: ``bgx_linear`` + ``bgx_gemm_tn``, what the
    default "auto" resolves to and the other tests cover); also the hand-written affine backward.  All against the
    fp64 oracle."""
    from bgflow_b200 import engine
    old = engine.config["backward_gemm"]
    engine.config["backward_gemm"] = mode
    try:
        dim, nblk = 10, 2
        hidden = (128, 128) if kind == "spline" or mode == "tcgen05" else (24, 24)
        blocks, split = of.make_stack(kind, dim, nblk, hidden=hidden, seed=2)
        blocks64, _ = of.make_stack(kind, dim, nblk, hidden=hidden, seed=2, dtype=torch.float64)
        flow = stack_from(blocks, split, DEV)
        g = torch.Generator().manual_seed(4)
        z = torch.rand(300, dim, generator=g) if kind == "spline" else torch.randn(300, dim, generator=g)
        wx, wd = torch.randn(300, dim, generator=g), torch.randn(300, 1, generator=g)
        for inverse in (False, True):
            x_ref, d_ref, gz_ref, gp_ref = _oracle_grads(kind, blocks64, split, z, wx, wd, inverse)
            flow.zero_grad()
            zc = z.to(DEV).requires_grad_(True)
            x, d = flow(zc, inverse=inverse)
            ((x * wx.to(DEV)).sum() + (d * wd.to(DEV)).sum()).backward()
            s = gz_ref.abs().max().item()
            np.testing.assert_allclose(zc.grad.cpu().double().numpy(), gz_ref.numpy(), atol=tol * s, rtol=tol)
            ours = []
            for m in flow.modules():
                if isinstance(m, bg.DenseNet):
                    lin = [l for l in m._layers if isinstance(l, torch.nn.Linear)]
                    ours += [l.weight.grad for l in lin] + [l.bias.grad for l in lin]
            assert len(ours) == len(gp_ref)
            for a, b in zip(ours, gp_ref):
                sc = max(b.abs().max().item(), 1e-6)
                a64 = a.cpu().double()
                if mode == "tf32":
                    # a TF32 pre-activation that lands on the other side of a ReLU kink switches one sample's whole
                    # contribution to a weight on or off: allow <= 1 % such elements, everything else within tol
                    bad = (a64 - b).abs() > 1.5 * tol * sc + 1.5 * tol * b.abs()
                    assert bad.double().mean().item() <= 0.01, (int(bad.sum()), bad.numel())
                else:
                    np.testing.assert_allclose(a64.numpy(), b.numpy(), atol=1.5 * tol * sc, rtol=1.5 * tol)
            if kind == "affine":      # log_alpha is not an oracle parameter: compare with the autograd (fp32) path
                la = [m for m in flow.modules() if isinstance(m, bg.AffineTransformer)]
                got = [t._log_alpha.grad.clone() for t in la]
                engine.config["backward_gemm"] = "fp32"
                flow.zero_grad()
                zc2 = z.to(DEV).requires_grad_(True)
                x2, d2 = flow(zc2, inverse=inverse)
                ((x2 * wx.to(DEV)).sum() + (d2 * wd.to(DEV)).sum()).backward()
                engine.config["backward_gemm"] = mode
                for a, t in zip(got, la):
                    torch.testing.assert_close(a, t._log_alpha.grad, atol=tol * float(t._log_alpha.grad.abs().max() + 1e-3), rtol=tol)
    finally:
        engine.config["backward_gemm"] = old


def test_kl_training_step_reduces_loss():
    """A few reverse-KL steps (bg.py:13-17) on a Gaussian target through kernel-forward blocks."""
    from bgflow_b200.distributed import kl_train_step
    torch.manual_seed(0)
    dim = 6
    layers = [bg.SplitFlow(3)]
    for _ in range(2):
        layers += [bg.CouplingFlow(bg.AffineTransformer(
            bg.DenseNet([3, 32, 3], activation=torch.nn.ReLU()), bg.DenseNet([3, 32, 3], activation=torch.nn.Tanh()))),
            bg.SwapFlow()]
    layers.append(bg.MergeFlow(3))
    flow = bg.SequentialFlow(layers).to(DEV)
    prior = bg.NormalDistribution(dim).to(DEV)
    target = bg.NormalDistribution(dim, mean=torch.full((dim,), 1.5)).to(DEV)
    gen = bg.BoltzmannGenerator(prior, flow, target)
    opt = torch.optim.Adam(gen.parameters(), lr=5e-3)
    losses = [float(kl_train_step(gen, opt, 4096)) for _ in range(60)]
    assert np.mean(losses[-5:]) < np.mean(losses[:5]) - 0.5, (losses[:5], losses[-5:])


@pytest.mark.parametrize("normalize", [True, False])
def test_internal_coordinate_gradients_match_oracle(normalize):
    """IC kernels forward, recompute backward: vector-Jacobian products of both directions against
    the oracle differentiated in fp64 (Energy.force / KLTrainer differentiate through the IC layer,
    trainers.py:158-175)."""
    from oracle import ic as oic
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z, normalize_angles=normalize)
    oplan = oic.make_plan(oic.ALA2_GLOBAL_Z)
    g = torch.Generator().manual_seed(8)
    xyz = torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float32).reshape(1, -1) + 0.01 * torch.randn(40, 66, generator=g)
    x64 = xyz.double().requires_grad_(True)
    ref = oic.xyz_to_ic(oplan, x64, normalize_angles=normalize)
    ws = [torch.randn(t.shape, generator=g) for t in ref]
    g_ref = torch.autograd.grad(sum((a * w.double()).sum() for a, w in zip(ref, ws)), x64)[0]
    xc = xyz.to(DEV).requires_grad_(True)
    outs = ic(xc)
    assert all(o.requires_grad for o in outs)
    sum((a * w.to(DEV)).sum() for a, w in zip(outs, ws)).backward()
    s = g_ref.abs().max().item()
    np.testing.assert_allclose(xc.grad.cpu().double().numpy(), g_ref.numpy(), atol=2e-3 * s, rtol=2e-3)

    ins64 = [t.detach().clone().requires_grad_(True) for t in ref[:5]]
    ref2 = oic.ic_to_xyz(oplan, *ins64, normalize_angles=normalize)
    ws = [torch.randn(t.shape, generator=g) for t in ref2]
    g_ref = torch.autograd.grad(sum((a * w.double()).sum() for a, w in zip(ref2, ws)), ins64)
    ins = [t.detach().float().to(DEV).requires_grad_(True) for t in ref[:5]]
    xyz_k, dlogp_k = ic(*ins, inverse=True)
    np.testing.assert_allclose(xyz_k.detach().cpu().double().numpy(), ref2[0].detach().numpy(), atol=1e-4)
    ((xyz_k * ws[0].to(DEV)).sum() + (dlogp_k * ws[1].to(DEV)).sum()).backward()
    for got, want in zip(ins, g_ref):
        s = max(want.abs().max().item(), 1e-6)
        np.testing.assert_allclose(got.grad.cpu().double().numpy(), want.numpy(), atol=3e-3 * s, rtol=3e-3)


def test_energy_gradient_reaches_flow_parameters_through_ic_layer():
    """z -> spline couplings (kernel) -> IC layer (kernel) -> xyz -> energy; parameter gradients
    against the same graph built from the oracle in fp64."""
    from oracle import ic as oic
    nb, na, nt = 21, 20, 19
    blocks, split = of.make_stack("spline", nb + na + nt, 2, hidden=(128, 128), seed=13)
    blocks64, _ = of.make_stack("spline", nb + na + nt, 2, hidden=(128, 128), seed=13, dtype=torch.float64)
    flow = stack_from(blocks, split, DEV)
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    oplan = oic.make_plan(oic.ALA2_GLOBAL_Z)
    g = torch.Generator().manual_seed(5)
    z = 0.25 + 0.5 * torch.rand(64, nb + na + nt, generator=g)
    x0 = torch.zeros(64, 1, 3)
    R = torch.full((64, 3), 0.5)
    w = torch.randn(64, 66, generator=g)

    def to_ics(y):      # spline outputs live in (0,1): bonds scaled to a physical range
        return 0.1 + 0.1 * y[:, :nb], 0.25 + 0.5 * y[:, nb:nb + na], y[:, nb + na:]

    params64 = []
    for b in blocks64:
        for t in b["params_net"].weights + b["params_net"].biases:
            t.requires_grad_(True)
            params64.append(t)
    y64, d64 = of.coupling_stack(blocks64, z.double(), split)
    xyz64, dic64 = oic.ic_to_xyz(oplan, *to_ics(y64), x0.double(), R.double())
    loss64 = (xyz64 * w.double()).sum() + (d64 + dic64).sum()
    g_ref = torch.autograd.grad(loss64, params64)

    y, d = flow(z.to(DEV))
    xyz, dic = ic(*to_ics(y), x0.to(DEV), R.to(DEV), inverse=True)
    loss = (xyz * w.to(DEV)).sum() + (d + dic).sum()
    np.testing.assert_allclose(loss.item(), loss64.item(), rtol=1e-4)
    loss.backward()
    ours = []
    for m in flow.modules():
        if isinstance(m, bg.DenseNet):
            lin = [l for l in m._layers if isinstance(l, torch.nn.Linear)]
            ours += [l.weight.grad for l in lin] + [l.bias.grad for l in lin]
    assert len(ours) == len(g_ref)
    for a, b in zip(ours, g_ref):
        s = max(b.abs().max().item(), 1e-6)
        np.testing.assert_allclose(a.cpu().double().numpy(), b.numpy(), atol=5e-3 * s, rtol=5e-3)


@pytest.mark.parametrize("d_t,n_bins,circular", [(7, 5, [True, False, True, False, False, True, False]),
                                                 (4, 12, True), (3, 20, False), (33, 8, False)])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("gemm", ["fp32", "auto"])
def test_spline_backward_kernel_against_torch_definition(d_t, n_bins, circular, inverse, gemm, monkeypatch):
    """bgx_spline_backward (hand-written chain rule) against autograd of the device-side torch
    definition, fp32 on both sides: circular / mixed masks, other bin counts, clamped inputs.

    With ``backward_gemm = "fp32"`` the recompute is the same cuBLAS GEMM as the torch definition's forward and the
    net's weights are scaled by 3 (saturated softmaxes, bins at their minimum width, gradients up to 1e4): every
    element must match.  The default tensor-core recompute carries 2^-17 per product (like the forward kernel), so a
    sample next to a knot can land in the neighbouring bin, where the gradient of the log-determinant jumps; under the
    x 3 weights one such sample dominates every element of the weight gradients, so this mode runs the well-conditioned
    net (weights x 1) and tolerates two (or 1 %) outlying elements per tensor."""
    from bgflow_b200 import _torch_math, engine
    monkeypatch.setitem(engine.config, "backward_gemm", gemm)
    torch.manual_seed(d_t * 100 + n_bins)
    n_nc = d_t - (sum(circular) if isinstance(circular, list) else (d_t if circular else 0))
    net = bg.DenseNet([6, 32, 3 * n_bins * d_t + n_nc], activation=torch.nn.SiLU()).to(DEV)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(3.0 if gemm == "fp32" else 1.0)
    tr = bg.ConditionalSplineTransformer(net, is_circular=circular)
    cond = torch.randn(257, 6, device=DEV, requires_grad=True)
    y = torch.rand(257, d_t, device=DEV)
    y[0, 0], y[1, -1] = -0.1, 1.2
    y.requires_grad_(True)
    wx, wd = torch.randn(257, d_t, device=DEV), torch.randn(257, 1, device=DEV)

    out, dlogp = tr.forward(cond, y, inverse=inverse)
    ((out * wx).sum() + (dlogp * wd).sum()).backward()
    got = [cond.grad.clone(), y.grad.clone()] + [p.grad.clone() for p in net.parameters()]
    cond.grad = y.grad = None
    net.zero_grad()
    out_t, dlogp_t = _torch_math.spline(tr, cond, y, inverse)
    torch.testing.assert_close(out, out_t, atol=2e-5, rtol=1e-4)
    ((out_t * wx).sum() + (dlogp_t * wd).sum()).backward()
    want = [cond.grad, y.grad] + [p.grad for p in net.parameters()]
    for a, b in zip(got, want):
        s = max(b.abs().max().item(), 1e-6)
        if gemm == "fp32":
            torch.testing.assert_close(a, b, atol=2e-3 * s, rtol=2e-3)
        else:
            bad = (a - b).abs() > 5e-3 * s + 5e-3 * b.abs()
            assert int(bad.sum()) <= max(2, 0.01 * bad.numel()), (int(bad.sum()), bad.numel())
    assert y.grad[0, 0] == 0 and got[1][0, 0] == 0          # clamped input: no gradient


@pytest.mark.parametrize("dims,act", [([10, 128, 128, 125], torch.nn.SiLU), ([33, 128, 825], torch.nn.ReLU), ([6, 32, 825], torch.nn.SiLU),
                                      ([200, 128, 64, 6], torch.nn.Tanh), ([5, 7], None), ([3, 24, 24, 5], torch.nn.SiLU)])
@pytest.mark.parametrize("batch", [1, 300, 4099])
def test_conditioner_recompute_and_backward_drivers(dims, act, batch):
    """bgx_mlp_forward_train / bgx_mlp_backward (one host call each for a whole DenseNet: bgx_linear, bgx_gemm_tn and
    the small activation / slice-sum kernels, on buffers padded to 4-float row strides) against fp64 autograd of
    dense.py's definition: widths off the 4 / 128 grid, a first layer wider than 128, a single-layer net."""
    from bgflow_b200 import _mlp_grad
    torch.manual_seed(sum(dims) + batch)
    net = bg.DenseNet(dims, activation=act() if act else None).to(DEV)
    assert _mlp_grad.tc_supported(net)
    x = torch.randn(batch, dims[0], device=DEV)
    w = torch.randn(batch, dims[-1], device=DEV)
    st = _mlp_grad.forward_tc(net, x)
    assert st["out_padded"].shape == (batch, (dims[-1] + 3) // 4 * 4)
    assert torch.count_nonzero(st["out_padded"][:, dims[-1]:]) == 0
    d_x, grads = _mlp_grad.backward_tc(st, w)
    net64 = bg.DenseNet(dims, activation=act() if act else None).double()
    net64.load_state_dict({k: v.double().cpu() for k, v in net.state_dict().items()})
    x64 = x.double().cpu().requires_grad_(True)
    out64 = net64(x64)
    ref = torch.autograd.grad(out64, [x64, *net64.parameters()], grad_outputs=w.double().cpu())
    sc = out64.abs().max().item()
    np.testing.assert_allclose(st["out"].cpu().double().numpy(), out64.detach().numpy(), atol=1e-4 * sc, rtol=1e-4)
    for a, b in zip([d_x, *grads], ref):
        assert a.shape == b.shape
        s = max(b.abs().max().item(), 1e-6)
        np.testing.assert_allclose(a.cpu().double().numpy(), b.numpy(), atol=1e-3 * s, rtol=3e-3)   # (a ReLU kink flip moves one sample)


def test_training_entry_points_reject_what_they_do_not_cover():
    """Error behaviour of the training-path C ABI: shapes outside the tensor-core linear kernel's cover are refused at
    pack time (the Python side then stays on torch autograd), inconsistent arguments come back as error codes, and a
    block with a conditioner the drivers do not cover still trains (through the autograd path) with the default mode."""
    import ctypes as C
    from bgflow_b200 import _lib, _mlp_grad, engine
    lib = _lib.load()
    wide = bg.DenseNet([200, 300, 5], activation=torch.nn.SiLU()).to(DEV)        # a 200 -> 300 layer: > 128 on both sides
    assert not _mlp_grad.tc_supported(wide)
    with pytest.raises(_lib.BgxError):
        engine.TrainNet().refresh([l.weight for l in wide._layers if isinstance(l, torch.nn.Linear)],
                                  [l.bias for l in wide._layers if isinstance(l, torch.nn.Linear)],
                                  [_lib.ACT_SILU, _lib.ACT_NONE])
    g = torch.zeros(64, 8, device=DEV)
    h = torch.zeros(64, 8, device=DEV)
    part = torch.zeros(4 * 128 * 129, device=DEV)
    args = (C.c_void_p(g.data_ptr()), 8, 8, C.c_void_p(h.data_ptr()), 8, 8)
    assert lib.bgx_gemm_tn(64, *args, 2, C.c_void_p(part.data_ptr()), None, None, None) != 0      # 2 slices > 1 batch tile
    assert lib.bgx_gemm_tn(64, *args, 0, C.c_void_p(part.data_ptr()), None, None, None) != 0
    assert lib.bgx_gemm_tn(64, C.c_void_p(g.data_ptr()), 4, 8, C.c_void_p(h.data_ptr()), 8, 8, 1,
                           C.c_void_p(part.data_ptr()), None, None, None) != 0                         # row stride < width
    assert lib.bgx_gemm_tn_slices(0, 8) == 0 and lib.bgx_gemm_tn_slices(64, 8) == 1
    # a spline block whose conditioner is too wide for the drivers: default mode falls back to autograd and still matches
    torch.manual_seed(0)
    net = bg.DenseNet([6, 200, 130, 3 * 8 * 4 + 4], activation=torch.nn.SiLU()).to(DEV)
    assert not _mlp_grad.tc_supported(net)
    tr = bg.ConditionalSplineTransformer(net, is_circular=False)
    cond = torch.randn(50, 6, device=DEV, requires_grad=True)
    y = torch.rand(50, 4, device=DEV, requires_grad=True)
    out, dl = tr.forward(cond, y)
    (out.sum() + dl.sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    assert torch.isfinite(cond.grad).all() and torch.isfinite(y.grad).all()


def test_training_step_as_a_cuda_graph():
    """distributed.GraphedStep: a whole reverse-KL step (sampling, kernel forward, tcgen05 backward, Adam) captured
    after a few eager calls and replayed — the replays keep training (weights are re-packed inside the graph: the loss
    keeps falling, parameters keep moving), and eager calls made after close() see the trained weights."""
    from bgflow_b200.distributed import BucketedGradReducer, GraphedStep
    torch.manual_seed(0)
    dim = 10
    blocks, split = of.make_stack("spline", dim, 2, hidden=(128, 128), seed=3)
    flow = stack_from(blocks, split, DEV)
    red = BucketedGradReducer(flow)
    opt = torch.optim.Adam(flow.parameters(), lr=3e-3, capturable=True)

    def step():
        red.zero_grad()
        z = torch.rand(2048, dim, device=DEV)
        x, dlogp = flow(z)
        loss = (0.5 * ((x - 0.5) / 0.1).square().sum(-1, keepdim=True) - dlogp).mean()
        loss.backward()
        red.finish()
        opt.step()
        return loss.detach()

    p0 = [p.detach().clone() for p in flow.parameters()]
    losses = []
    with GraphedStep(step, flow.parameters(), warmup=2) as gs:
        for _ in range(40):
            losses.append(float(gs()))
        assert gs.graph is not None
    assert np.mean(losses[-5:]) < np.mean(losses[:5]) - 0.5, losses[::5]
    assert all(not torch.equal(a, b) for a, b in zip(p0, flow.parameters()))
    # after close(): an eager forward runs on the trained weights (packed-weight caches were invalidated)
    z = torch.rand(512, dim, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    with torch.no_grad():
        x_fast, d_fast = flow(z)
    blocks_now = []
    ref = stack_from(blocks, split, DEV)
    ref.load_state_dict(flow.state_dict())
    with torch.no_grad():
        x_ref, d_ref = ref(z)
    assert torch.equal(x_fast, x_ref) and torch.equal(d_fast, d_ref)
    red.remove()


@pytest.mark.parametrize("kind", ["spline", "affine"])
def test_wide_block_gradients_on_the_training_kernels(kind):
    """BASELINE config 5 shapes through the training path: a D = 384 block (conditioner 192 -> 128 -> 128 -> 4800 / 192:
    a first layer k-tiled in two groups, a last layer of 38 passes, its transpose k-split 38 times, weight gradients over
    38 feature tiles) against the fp64-differentiated oracle, forward and inverse."""
    from bgflow_b200 import _mlp_grad
    dim = 384
    blocks, split = of.make_stack(kind, dim, 1, seed=5)
    blocks64, _ = of.make_stack(kind, dim, 1, seed=5, dtype=torch.float64)
    flow = stack_from(blocks, split, DEV)
    nets = [m for m in flow.modules() if isinstance(m, bg.DenseNet)]
    assert nets and all(_mlp_grad.tc_supported(n) for n in nets)
    g = torch.Generator().manual_seed(9)
    B = 300
    z = torch.rand(B, dim, generator=g) if kind == "spline" else torch.randn(B, dim, generator=g)
    wx, wd = torch.randn(B, dim, generator=g), torch.randn(B, 1, generator=g)
    for inverse in (False, True):
        x_ref, d_ref, gz_ref, gp_ref = _oracle_grads(kind, blocks64, split, z, wx, wd, inverse)
        flow.zero_grad()
        zc = z.to(DEV).requires_grad_(True)
        x, d = flow(zc, inverse=inverse)
        ((x * wx.to(DEV)).sum() + (d * wd.to(DEV)).sum()).backward()
        s = gz_ref.abs().max().item()
        np.testing.assert_allclose(zc.grad.cpu().double().numpy(), gz_ref.numpy(), atol=3e-3 * s, rtol=3e-3)
        ours = []
        for m in nets:
            lin = [l for l in m._layers if isinstance(l, torch.nn.Linear)]
            ours += [l.weight.grad for l in lin] + [l.bias.grad for l in lin]
        assert len(ours) == len(gp_ref)
        for a, b in zip(ours, gp_ref):
            sc = max(b.abs().max().item(), 1e-6)
            bad = (a.cpu().double() - b).abs() > 4.5e-3 * sc + 4.5e-3 * b.abs()
            assert int(bad.sum()) <= max(2, 0.002 * bad.numel()), (int(bad.sum()), bad.numel())    # (ReLU kinks / knot flips)
