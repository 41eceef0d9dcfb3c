"""BASELINE.json's full sizes through size-independent properties (the oracle cannot run 2^20 rows
in seconds): forward -> inverse round trips, log-det antisymmetry, row independence (a row's result
does not depend on the batch it travels in: the full-batch output at sampled rows equals a launch
over those rows alone, bit for bit), and the fp64 oracle on a sampled subset of rows."""

import numpy as np
import pytest
import torch

import bgflow_b200 as bg
from oracle import flows as of, ic as oic
from helpers import config4_blocks, stack_from
from test_gpu_pipeline import build as build_config4, WIDTHS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("kind,batch,hidden", [("affine", 65536, (128, 128, 128)), ("spline", 65536, (128, 128)),
                                               ("spline", 1 << 20, (128, 128))])
def test_coupling_stacks_at_baseline_sizes(kind, batch, hidden):
    """Configs 2, 3 and 5: D = 66, 8 blocks, B = 65536 and 2^20."""
    blocks, split = of.make_stack(kind, 66, 8, hidden=hidden, seed=0)
    flow = stack_from(blocks, split, DEV)
    g = torch.Generator().manual_seed(1)
    z = (torch.rand(batch, 66, generator=g) if kind == "spline" else torch.randn(batch, 66, generator=g)).to(DEV)
    with torch.no_grad():
        x, dlogp = flow(z)
        zb, dinv = flow(x, inverse=True)
        assert torch.isfinite(x).all() and torch.isfinite(dlogp).all()
        torch.testing.assert_close(zb, z, atol=2e-4, rtol=1e-4)
        torch.testing.assert_close(dlogp + dinv, torch.zeros_like(dlogp), atol=2e-3, rtol=0)
        # row independence: 4096 sampled rows (a multiple of 4: same kernel) alone give the same bits
        idx = torch.randperm(batch, generator=g)[:4096].sort().values.to(DEV)
        xs, ds = flow(z[idx].contiguous())
        assert torch.equal(xs, x[idx]) and torch.equal(ds, dlogp[idx])
    # oracle in fp64 on 256 of those rows
    sub = idx[:256].cpu()
    blocks64, _ = of.make_stack(kind, 66, 8, hidden=hidden, seed=0, dtype=torch.float64)
    x_ref, d_ref = of.coupling_stack(blocks64, z[sub.to(DEV)].cpu().double(), split)
    np.testing.assert_allclose(x[sub.to(DEV)].cpu().double().numpy(), x_ref.numpy(), atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(dlogp[sub.to(DEV)].cpu().double().numpy(), d_ref.numpy(), atol=1e-3, rtol=1e-4)


def test_global_ic_at_two_to_the_twenty():
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    g = torch.Generator().manual_seed(2)
    B = 1 << 20
    xyz = (torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float32).reshape(1, -1) + 0.01 * torch.randn(B, 66, generator=g)).to(DEV)
    *ics, dlogp = ic(xyz)
    back, dinv = ic(*ics, inverse=True)
    torch.testing.assert_close(back, xyz, atol=1e-4, rtol=0)
    torch.testing.assert_close(dlogp + dinv, torch.zeros_like(dlogp), atol=1e-3, rtol=0)
    idx = torch.arange(0, B, 4099, device=DEV)
    ref = oic.xyz_to_ic(oic.make_plan(oic.ALA2_GLOBAL_Z), xyz[idx].cpu().double())
    for got, want in zip((*ics, dlogp), ref):
        np.testing.assert_allclose(got[idx].cpu().double().numpy(), want.numpy(), atol=1e-3 if want.shape[-1] == 1 else 1e-4)


def test_config4_at_262144():
    """Config 4 at its BASELINE batch: sampling direction then energy direction through the fused
    pipeline; the prior sample must come back (torsions modulo 1)."""
    flow = build_config4(config4_blocks(torch.float32), fuse=True)
    g = torch.Generator().manual_seed(4)
    us = [(torch.rand(262144, w, generator=g) * 0.96 + 0.02).to(DEV) for w in WIDTHS]
    with torch.no_grad():
        xyz, aug, dlogp = flow(*us)
        *back, dinv = flow(xyz, aug, inverse=True)
    assert torch.isfinite(xyz).all() and torch.isfinite(dlogp).all()
    for u, b in zip(us, back):
        e = (u - b).abs()
        if u.shape[1] == 19:
            e = torch.minimum(e, 1 - e)
        assert float(e.median()) < 1e-5 and float(e.quantile(0.999)) < 5e-3, (u.shape, float(e.max()))
    s = (dlogp + dinv).abs()
    assert float(s.median()) < 1e-3 and float(s.quantile(0.99)) < 5e-2


def test_training_kernels_at_two_to_the_twenty():
    """The training-path kernels at the bench batch (2^20 rows: 3.4 GB operands, byte offsets beyond 2^32): bgx_linear in
    both split modes and bgx_spline_backward against the same calls on a slice of the rows (rows are independent:
    bit-identical), bgx_gemm_tn against an fp64 product."""
    from bgflow_b200 import engine
    B, sub = 1 << 20, slice((1 << 20) - 4096, 1 << 20)          # the LAST rows: the largest offsets
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(B, 128, device=DEV, generator=g)
    w = torch.randn(828, 128, device=DEV, generator=g) / 128 ** 0.5
    b = torch.randn(828, device=DEV, generator=g)
    f = engine.LinearTC()
    y = f(x, w, b)                                               # K = 128 -> N = 828 (7 passes)
    assert torch.equal(y[sub], f(x[sub].contiguous(), w, b))
    ref = x[sub].double() @ w.double().t() + b.double()
    np.testing.assert_allclose(y[sub].double().cpu().numpy(), ref.cpu().numpy(), atol=3e-5 * ref.abs().max().item(), rtol=1e-4)
    f2 = engine.LinearTC()
    gx = f2(y, w.t().contiguous())                               # K = 828 -> N = 128 (7 groups, cp.async prefetch)
    assert torch.equal(gx[sub], f2(y[sub].contiguous(), w.t().contiguous()))
    dw, db = engine.gemm_tn(y, x, 825)                           # reduction over 2^20 rows
    ref_w = torch.zeros(825, 128, dtype=torch.float64, device=DEV)
    for lo in range(0, B, 1 << 17):
        ref_w += y[lo:lo + (1 << 17), :825].double().t() @ x[lo:lo + (1 << 17)].double()
    sw = (y[:, :825].abs().double().t() @ x.abs().double()).max().item()
    np.testing.assert_allclose(dw.double().cpu().numpy(), ref_w.cpu().numpy(), atol=3e-5 * sw, rtol=0)
    np.testing.assert_allclose(db.double().cpu().numpy(), y[:, :825].double().sum(0).cpu().numpy(),
                               atol=1e-5 * y[:, :825].abs().double().sum(0).max().item(), rtol=0)
    d_t, nb = 33, 8
    yy = torch.rand(B, d_t, device=DEV, generator=g)
    go = torch.randn(B, d_t, device=DEV, generator=g)
    gd = torch.randn(B, 1, device=DEV, generator=g)
    ecol = torch.arange(3 * nb * d_t, 3 * nb * d_t + d_t, dtype=torch.int32, device=DEV)
    dp, dy = engine.spline_backward(y, yy, go, gd, ecol, nb)     # 828-float rows: the vector path
    dp2, dy2 = engine.spline_backward(y[sub].contiguous(), yy[sub].contiguous(), go[sub].contiguous(), gd[sub].contiguous(), ecol, nb)
    assert torch.equal(dp[sub][:, :825], dp2[:, :825]) and torch.equal(dy[sub], dy2)
    assert torch.isfinite(dp[:, :825]).all() and torch.isfinite(dy).all()
