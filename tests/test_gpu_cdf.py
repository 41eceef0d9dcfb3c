"""Parity of the IC-domain CDF-map kernel (``bgx_cdf_map``) and of the fused builder tail
(``bgx_ic_to_xyz_mapped`` / ``bgx_ic_from_xyz_mapped``) against the reference-generated golden
fixtures (tests/golden/cdf_maps.npz) and the oracle.

Stated tolerances: mapped values 2e-5 abs+rel against the reference's fp64 run on inputs away from
the eps-clamped edges (|dx/du| = 1/pdf amplifies the fp32 rounding of u in the tails: there the
bound is 4 ulp(u) / pdf); log-dets 1e-4 per row; fused tail: coordinates 1e-4, dlogp 1e-3 (the
IC kernels' own bars)."""

import math

import numpy as np
import pytest
import torch

import bgflow_b200 as bg
from bgflow_b200 import _lib
from oracle import cdf as ocdf, ic as oic
from conftest import load_golden
from helpers import marginals

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FIELDS = {"bonds": 21, "angles": 20, "torsions": 19, "fixed": 9, "augmented": 10}


def _t(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def _check_icdf(x, dlogp, u, g, name):
    u64 = u.cpu().double().numpy()
    bulk = (u64 > 1e-3) & (u64 < 1 - 1e-3)
    x64 = g[f"{name}_x_f64"]
    got = x.cpu().double().numpy()
    np.testing.assert_allclose(got[bulk], x64[bulk], rtol=2e-5, atol=2e-5, err_msg=name)
    rows = bulk.all(-1)
    np.testing.assert_allclose(dlogp.cpu().double().numpy()[rows], g[f"{name}_dlogp_f64"][rows], rtol=1e-4, atol=1e-4)
    assert torch.isfinite(x).all() and torch.isfinite(dlogp).all()       # eps-clamped edge rows included


@pytest.mark.parametrize("name", list(FIELDS))
def test_cdf_map_matches_reference_golden(name):
    g = load_golden("cdf_maps")
    flow = bg.InverseFlow(bg.CDFTransform(marginals()[name]))          # generator_builder.py:451
    u = _t(g[f"{name}_u_f32"])
    n0 = _lib.launch_count()
    x, dlogp = flow(u)
    assert _lib.launch_count() == n0 + 1
    assert x.shape == u.shape and dlogp.shape == (u.shape[0], 1)
    _check_icdf(x, dlogp, u, g, name)
    # cdf direction on the reference's fp32 x, against the oracle evaluated in fp64 on the same x
    xb = _t(g[f"{name}_x_f32"])
    ub, dlogpb = flow(xb, inverse=True)
    dist = ocdf.ic_marginals(FIELDS, torch.float64)[name]
    ref_u, ref_d = ocdf.cdf_transform(dist, xb.cpu().double())
    np.testing.assert_allclose(ub.cpu().double().numpy(), ref_u.numpy(), rtol=1e-5, atol=3e-7)
    rows = torch.isfinite(ref_d[:, 0]).numpy()
    np.testing.assert_allclose(dlogpb.cpu().double().numpy()[rows], ref_d.numpy()[rows], rtol=1e-5, atol=1e-4)


def test_custom_and_halfopen_truncated_normals():
    g = load_golden("cdf_maps")
    tn = bg.TruncatedNormalDistribution(_t(g["custom_mu_f32"]), _t(g["custom_sigma_f32"]), _t(g["custom_lower_f32"]),
                                        _t(g["custom_upper_f32"]))
    flow = bg.InverseFlow(bg.CDFTransform(tn))
    u = _t(g["custom_u_f32"])
    x, d = flow(u)
    _check_icdf(x, d, u, g, "custom")
    half = bg.TruncatedNormalDistribution(torch.zeros(5, device=DEV), lower_bound=-torch.tensor(math.inf, device=DEV))
    u = _t(g["halfopen_u_f32"])
    x, d = bg.InverseFlow(bg.CDFTransform(half))(u)
    _check_icdf(x, d, u, g, "halfopen")


@pytest.mark.parametrize("batch", [1, 127, 128, 129, 1000, 40000])
def test_multi_tensor_launch_round_trip_and_accumulation(batch):
    """All five fields in one launch (strided views included): round trip, log-det antisymmetry,
    running-dlogp accumulation, agreement with one launch per field."""
    m = marginals()
    names = list(FIELDS)
    g = torch.Generator().manual_seed(batch)
    big = torch.rand(batch, 100, generator=g).to(DEV) * 0.98 + 0.01
    us, col = [], 0
    for n in names:                                   # column slices of one wide tensor: row stride 100
        us.append(big[:, col:col + FIELDS[n]])
        col += FIELDS[n]
    multi = bg.InverseFlow(bg.MultiCDFFlow([m[n] for n in names]))
    n0 = _lib.launch_count()
    *xs, dlogp = multi(*us)
    assert _lib.launch_count() == n0 + 1
    acc = torch.randn(batch, 1, generator=g).to(DEV)
    *xs2, dlogp2 = multi(*us, _dlogp_acc=acc)
    torch.testing.assert_close(dlogp2, dlogp + acc, atol=1e-5, rtol=1e-6)
    total = 0
    for n, u, x in zip(names, us, xs):
        xi, di = bg.InverseFlow(bg.CDFTransform(m[n]))(u.contiguous())
        assert torch.equal(xi, x)
        total = total + di
    torch.testing.assert_close(dlogp, total, atol=1e-4, rtol=1e-6)
    *back, dinv = multi(*xs, inverse=True)
    for u, b in zip(us, back):
        torch.testing.assert_close(b, u, atol=2e-6, rtol=1e-5)
    torch.testing.assert_close(dlogp + dinv, torch.zeros_like(dlogp), atol=2e-3, rtol=0)


def test_wide_tensor_chunks_and_eps_none():
    """More columns than one shared-memory pass holds (88): chunked accumulation; eps=None = no clamps."""
    n = 300
    d = torch.distributions.Normal(torch.linspace(-1, 1, n, device=DEV), torch.linspace(0.5, 2, n, device=DEV))
    flow = bg.CDFTransform(d, eps=None)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(257, n, generator=g).to(DEV)
    u, dl = flow(x)
    ref_u, ref_d = ocdf.cdf_transform(ocdf.Normal(d.loc.cpu().double(), d.scale.cpu().double()), x.cpu().double(), eps=None)
    np.testing.assert_allclose(u.cpu().double().numpy(), ref_u.numpy(), rtol=3e-5, atol=3e-7)
    np.testing.assert_allclose(dl.cpu().double().numpy(), ref_d.numpy(), rtol=1e-5, atol=1e-3)
    assert bg.CDFTransform(d)(torch.empty(0, n, device=DEV))[0].shape == (0, n)


def test_generic_distribution_runs_reference_ops():
    d = torch.distributions.Laplace(torch.zeros(4, device=DEV), torch.ones(4, device=DEV))
    x = torch.randn(33, 4, device=DEV)
    u, dl = bg.CDFTransform(d)(x)
    torch.testing.assert_close(u, d.cdf(x).clamp(1e-7, 1 - 1e-7))
    torch.testing.assert_close(dl, d.log_prob(x).sum(-1, keepdim=True))


def test_input_gradients_and_learnable_marginal():
    m = marginals()
    u = (torch.rand(64, 21, device=DEV) * 0.9 + 0.05).requires_grad_(True)
    flow = bg.InverseFlow(bg.CDFTransform(m["bonds"]))
    x, dl = flow(u)
    wx, wd = torch.randn_like(x), torch.randn_like(dl)
    (gu,) = torch.autograd.grad((x * wx).sum() + (dl * wd).sum(), u)
    u64 = u.detach().cpu().double().requires_grad_(True)
    rx, rd = ocdf.cdf_transform(ocdf.ic_marginals(FIELDS, torch.float64)["bonds"], u64, inverse=True)
    (gr,) = torch.autograd.grad((rx * wx.cpu().double()).sum() + (rd * wd.cpu().double()).sum(), u64)
    np.testing.assert_allclose(gu.cpu().double().numpy(), gr.numpy(), rtol=2e-3, atol=2e-3 * float(gr.abs().max()))
    # the reference's own learnable case (tests/nn/flow/test_cdf.py:51-64)
    inp = torch.arange(0.1, 1.0, 0.1, device=DEV)[:, None].requires_grad_(True)
    tn = bg.TruncatedNormalDistribution(mu=torch.tensor([0.5], device=DEV), upper_bound=torch.tensor([1.0], device=DEV),
                                        is_learnable=True)
    out, dlogp = bg.InverseFlow(bg.CDFTransform(tn))(inp)
    assert out.mean().item() == pytest.approx(0.5, abs=1e-5)
    out.mean().backward(create_graph=True)
    dlogp.mean().backward()
    assert tn._mu.grad is not None


def test_distribution_transfer_and_constrain_flows():
    """tests/nn/flow/test_cdf.py:10-48 on the device."""
    src = torch.distributions.Normal(torch.zeros(2, device=DEV), torch.ones(2, device=DEV))
    tgt = torch.distributions.Normal(torch.ones(2, device=DEV), torch.ones(2, device=DEV))
    swap = bg.DistributionTransferFlow(src, tgt)
    out, dlogp = swap.forward(torch.zeros(2, 2, device=DEV))
    torch.testing.assert_close(out, torch.ones(2, 2, device=DEV), atol=1e-6, rtol=0)
    torch.testing.assert_close(dlogp, torch.zeros(2, 1, device=DEV), atol=1e-6, rtol=0)
    out2, dlogp = swap.forward(out, inverse=True)
    torch.testing.assert_close(out2, torch.zeros(2, 2, device=DEV), atol=1e-6, rtol=0)
    torch.manual_seed(1)
    flow = bg.ConstrainGaussianFlow(mu=torch.ones(10, device=DEV), lower_bound=1e-10)
    samples = ((1.0 + torch.randn(10, 10)) * 1000.).to(DEV)
    y, dlogp = flow.forward(samples)
    assert y.shape == (10, 10) and dlogp.shape == (10, 1) and (y >= 0.0).all() and (dlogp.sum() < 0.0).all()
    flow = bg.ConstrainGaussianFlow(mu=torch.ones(10, device=DEV), sigma=torch.ones(10, device=DEV),
                                    lower_bound=-1000., upper_bound=1000.)
    torch.manual_seed(1)
    samples = (1.0 + torch.randn(10, 10)).to(DEV)
    y, dlogp = flow.forward(samples)
    torch.testing.assert_close(samples, y, atol=1e-4, rtol=0)
    torch.testing.assert_close(dlogp, torch.zeros_like(dlogp), atol=1e-4, rtol=0)


# ------------------------------------------------------------------ fused builder tail

def _builder_tail(ic, m):
    layers = [bg.WrapFlow(bg.InverseFlow(bg.CDFTransform(m[n])), (i,))
              for i, n in enumerate(("bonds", "angles", "torsions", "augmented"))]
    layers += [bg.SetConstantFlow([4], [torch.zeros(1, 3, device=DEV)]),
               bg.SetConstantFlow([5], [torch.tensor([0.5, 0.5, 0.5], device=DEV)]),
               bg.WrapFlow(bg.InverseFlow(ic), indices=[0, 1, 2, 4, 5], out_indices=(0,))]
    return bg.SequentialFlow(layers).to(DEV)


@pytest.mark.parametrize("batch", [96, 1000])
def test_fused_tail_matches_unfused_and_oracle(batch):
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    m = marginals()
    tail = _builder_tail(ic, m)
    fused = bg.fuse_domain_maps(tail)
    g = torch.Generator().manual_seed(batch)
    us = [(torch.rand(batch, w, generator=g) * 0.9 + 0.05).to(DEV) for w in (21, 20, 19, 10)]
    # keep the geometry well conditioned: bonds 0.1..0.2 nm, angles 100..125 deg around the icdf pre-images
    x_ref, aug_ref, d_ref = tail(*us)
    n0 = _lib.launch_count()
    x, aug, d = fused(*us)
    assert _lib.launch_count() <= n0 + 3                      # one multi-cdf launch (aug) + the mapped-IC kernel (bulk tiles + tail rows)
    assert x.shape == (batch, 66) and aug.shape == (batch, 10) and d.shape == (batch, 1)
    torch.testing.assert_close(aug, aug_ref, atol=0, rtol=0)
    torch.testing.assert_close(x, x_ref, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(d, d_ref, atol=1e-3, rtol=1e-5)
    # oracle in fp64: icdf maps then IC -> xyz
    om = ocdf.ic_marginals({"bonds": 21, "angles": 20, "torsions": 19, "augmented": 10}, torch.float64)
    u64 = [u.cpu().double() for u in us]
    ics, dl = [], 0
    for u, n in zip(u64, ("bonds", "angles", "torsions", "augmented")):
        y, dd = ocdf.cdf_transform(om[n], u, inverse=True)
        ics.append(y)
        dl = dl + dd
    plan = oic.make_plan(oic.ALA2_GLOBAL_Z)
    x0 = torch.zeros(batch, 1, 3, dtype=torch.float64)
    R = torch.full((batch, 3), 0.5, dtype=torch.float64)
    xo, do = oic.ic_to_xyz(plan, ics[0], ics[1], ics[2], x0, R)
    np.testing.assert_allclose(x.cpu().double().numpy(), xo.numpy(), atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(d.cpu().double().numpy(), (dl + do).numpy(), atol=2e-3, rtol=1e-5)
    # energy direction through the fused tail
    b_us = fused(x, aug, inverse=True)
    r_us = tail(x_ref, aug_ref, inverse=True)
    for a, b in zip(b_us[:-1], r_us[:-1]):
        torch.testing.assert_close(a, b, atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(b_us[-1], r_us[-1], atol=2e-3, rtol=1e-4)
