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
