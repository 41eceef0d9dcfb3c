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
