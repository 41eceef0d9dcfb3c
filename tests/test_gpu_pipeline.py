"""BASELINE config 4 end to end: builder-exact augmented Ala2 stack — 14 multi-tensor
spline couplings (circular torsions, WrapPeriodic conditioners) -> icdf maps of every field ->
InverseFlow(GlobalInternalCoordinateTransformation) with origin 0 / rotation (0.5, 0.5, 0.5) —
through the fused kernels, against the oracle composed in fp64 on the CPU.

Tolerances: the 14 couplings stay inside the coupling bar (1e-4 on [0,1]-valued fields, 1e-3 on
dlogp); the icdf maps then amplify an input error by 1/pdf (up to ~30x for |z| < 2.5) and the
chain of 19 placements by the bond lengths (~1..3 here), so Cartesian coordinates are checked at
the median (2e-5) and at the 99th percentile (1e-3) rather than at the maximum."""

import numpy as np
import pytest
import torch

import bgflow_b200 as bg
from bgflow_b200 import _lib
from oracle import cdf as ocdf, flows as of, ic as oic
from helpers import config4_blocks, marginals, transformer_from

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
WIDTHS = (21, 20, 19, 10)
NAMES = ("bonds", "angles", "torsions", "augmented")


def build(blocks, fuse):
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    m = marginals()
    layers = [bg.CouplingFlow(transformer_from(b, DEV), transformed_indices=b["transformed"], cond_indices=b["cond"])
              for b in blocks]
    layers += [bg.WrapFlow(bg.InverseFlow(bg.CDFTransform(m[n])), (i,)) for i, n in enumerate(NAMES)]
    layers += [bg.SetConstantFlow([4], [torch.zeros(1, 3, device=DEV)]),
               bg.SetConstantFlow([5], [torch.tensor([0.5, 0.5, 0.5], device=DEV)]),
               bg.WrapFlow(bg.InverseFlow(ic), indices=[0, 1, 2, 4, 5], out_indices=(0,))]
    flow = bg.SequentialFlow(layers).to(DEV)
    return bg.fuse_domain_maps(flow) if fuse else flow


def oracle_pipeline(blocks64, us):
    xs, dlogp = list(us), 0
    for b in blocks64:
        xs, d = of.coupling_block(b, xs)
        dlogp = dlogp + d
    fields = list(xs)
    om = ocdf.ic_marginals(dict(zip(NAMES, WIDTHS)), torch.float64)
    for i, n in enumerate(NAMES):
        xs[i], d = ocdf.cdf_transform(om[n], xs[i], inverse=True)
        dlogp = dlogp + d
    B = xs[0].shape[0]
    xyz, d = oic.ic_to_xyz(oic.make_plan(oic.ALA2_GLOBAL_Z), xs[0], xs[1], xs[2],
                           torch.zeros(B, 1, 3, dtype=torch.float64), torch.full((B, 3), 0.5, dtype=torch.float64))
    return fields, xyz, xs[3], dlogp + d


@pytest.mark.parametrize("batch,fuse", [(257, False), (257, True), (4608, True)])
def test_config4_pipeline_matches_oracle(batch, fuse):
    blocks = config4_blocks(torch.float32)
    blocks64 = config4_blocks(torch.float64)
    flow = build(blocks, fuse)
    g = torch.Generator().manual_seed(batch)
    us = [torch.rand(batch, w, generator=g) for w in WIDTHS]               # the builder's uniform prior
    with torch.no_grad():
        flow(*(u.to(DEV) for u in us))                 # first call packs the conditioner weights
        n0 = _lib.launch_count()
        xyz, aug, dlogp = flow(*(u.to(DEV) for u in us))
    n_launch = _lib.launch_count() - n0
    assert xyz.shape == (batch, 66) and aug.shape == (batch, 10) and dlogp.shape == (batch, 1)
    fields, xyz_ref, aug_ref, dlogp_ref = oracle_pipeline(blocks64, [u.double() for u in us])
    # well-conditioned samples: every field value away from the eps-clamped ends of the icdf maps
    ok = torch.stack([((f > 2e-3) & (f < 1 - 2e-3)).all(-1) for f in fields]).all(0).numpy()
    assert ok.mean() > 0.6
    err = (xyz.cpu().double() - xyz_ref).abs().numpy()[ok]
    assert np.median(err) < 2e-5 and np.quantile(err, 0.99) < 1e-3, (np.median(err), np.quantile(err, 0.99), err.max())
    np.testing.assert_allclose(aug.cpu().double().numpy()[ok], aug_ref.numpy()[ok], atol=2e-3, rtol=1e-3)
    derr = (dlogp.cpu().double() - dlogp_ref).abs().numpy()[ok]
    assert np.median(derr) < 2e-4 and np.quantile(derr, 0.99) < 1e-2, (np.median(derr), derr.max())
    if fuse:
        # 14 couplings (+ segment gathers are torch copies) + one cdf launch (augmented) + one mapped-IC launch
        assert n_launch in (16, 17), n_launch      # 17: the IC kernel ran its full tiles through bulk copies + one tail launch


def test_config4_energy_direction_round_trip():
    """generator.energy(x): xyz -> ICs -> cdf maps -> inverse couplings; composed with the sampling
    direction it must return the prior sample and the negated log-det."""
    blocks = config4_blocks(torch.float32)
    flow = build(blocks, fuse=True)
    g = torch.Generator().manual_seed(3)
    us = [(torch.rand(2000, w, generator=g) * 0.9 + 0.05).to(DEV) for w in WIDTHS]
    with torch.no_grad():
        xyz, aug, dlogp = flow(*us)
        *back, dinv = flow(xyz, aug, inverse=True)
    for u, b, tol in zip(us, back, (2e-3, 2e-3, 2e-3, 2e-3)):
        e = (u - b).abs()
        if u.shape[1] == 19:
            e = torch.minimum(e, 1 - e)                      # torsions are circular
        assert float(e.median()) < 1e-5 and float(e.quantile(0.99)) < tol, (u.shape, float(e.max()))
    s = (dlogp + dinv).abs()
    assert float(s.median()) < 1e-3 and float(s.quantile(0.99)) < 5e-2


@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("chunk", [None, 1000, 4096])
def test_host_pipeline_matches_one_device_call(chunk, graph):
    """HostPipeline.run / .sample (pinned host buffers, chunks on a ring of streams — what bench.py's e2e number
    times) give the rows a single device-side call gives, whatever the chunking, eagerly and as a replayed CUDA graph;
    the default chunk is a whole number of kernel waves."""
    from bgflow_b200.host import HostPipeline, wave_rows
    dim, rows = 10, 5000
    blocks, split = of.make_stack("spline", dim, 2, hidden=(128, 128), seed=5)
    from helpers import stack_from
    flow = stack_from(blocks, split, DEV)
    prior = bg.UniformDistribution(torch.zeros(dim), torch.ones(dim)).to(DEV)
    pipe = HostPipeline(flow, dim, dim, rows, DEV, chunk_rows=chunk, prior=prior, with_energy=True, use_graph=graph)
    assert pipe.chunk == (chunk or wave_rows(DEV, 2 if graph else 4)) and wave_rows(DEV) % 256 == 0
    z = torch.rand(rows, dim, generator=torch.Generator().manual_seed(0)).pin_memory()
    with torch.no_grad():
        x_ref, d_ref = flow(z.to(DEV))
    for call in range(4):             # eager, capture + replay, replay, replay on changed host data
        if call == 3:
            z.copy_(torch.rand(rows, dim, generator=torch.Generator().manual_seed(1)))
            with torch.no_grad():
                x_ref, d_ref = flow(z.to(DEV))
        x, d = pipe.run(z)
        assert torch.equal(x, x_ref.cpu()) and torch.equal(d, d_ref.cpu())      # rows are independent: bit-exact
    assert (len(pipe._graphs) == 1) == graph
    seen = []
    for call in range(3):
        xs, ds, es = pipe.sample(rows)
        assert xs.shape == (rows, dim) and ds.shape == (rows, 1) and es.shape == (rows, 1)
        assert torch.isfinite(xs).all() and torch.isfinite(es).all()
        assert all(not torch.equal(xs, prev) for prev in seen)                  # every call draws new prior samples
        seen.append(xs.clone())
    with torch.no_grad():             # a parameter update invalidates the captured graph
        for p in flow.parameters():
            p.mul_(1.01)
        x_ref, d_ref = flow(z.to(DEV))
    for call in range(3):
        x, d = pipe.run(z)
        assert torch.equal(x, x_ref.cpu()) and torch.equal(d, d_ref.cpu())
    x2, d2 = pipe.run(z.clone())      # an unpinned buffer: no graph, same rows
    assert torch.equal(x2, x_ref.cpu())
    with pytest.raises(ValueError):
        pipe.run(torch.zeros(rows + 1, dim))
