"""CPU-side tests: tuple plumbing (mirrors the reference's tests/nn/flow/test_coupling.py,
test_sequential.py, test_inverted.py), z-matrix staging, the C ABI surface, and a 2-rank gloo
run of the batch sharding used by bench.py.  No kernel is launched here."""

import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import bgflow_b200 as bg
from bgflow_b200 import _lib, engine
from oracle import ic as oic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class DummyTransformer(bg.Transformer):
    """tests/nn/flow/test_coupling.py:58-70: records what it was called with."""

    def __init__(self):
        super().__init__()
        self.seen = None

    def _forward(self, x, y, **kwargs):
        self.seen = (x.shape, y.shape)
        return y + 1.0, torch.ones(*y.shape[:-1], 1)

    def _inverse(self, x, y, **kwargs):
        return y - 1.0, -torch.ones(*y.shape[:-1], 1)


def test_split_by_sizes_and_indices():
    x = torch.arange(20.0).reshape(2, 10)
    a, b, c, dlogp = bg.SplitFlow(2, 3)(x)
    assert a.shape == (2, 2) and b.shape == (2, 3) and c.shape == (2, 5) and dlogp.shape == (2, 1)
    y, d = bg.SplitFlow(2, 3)(a, b, c, inverse=True)
    assert torch.equal(y, x) and torch.equal(d, torch.zeros(2, 1))
    with pytest.raises(ValueError):
        bg.SplitFlow(8, 3)(x)
    split = bg.SplitFlow([0, 9], [1, 2, 3], [4, 5, 6, 7, 8])
    a, b, c, _ = split(x)
    assert torch.equal(a, x[:, [0, 9]])
    y, _ = split(a, b, c, inverse=True)
    assert torch.equal(y, x)
    with pytest.raises(ValueError):
        bg.SplitFlow([0, 1], [1, 2])(x)
    with pytest.raises(ValueError):
        bg.SplitFlow([0, 1], [2, 3])(x)


def test_coupling_multiple_tensors_generic_transformer():
    tr = DummyTransformer()
    flow = bg.CouplingFlow(tr, transformed_indices=(0, 2), cond_indices=(1, 3))
    xs = [torch.zeros(5, w) for w in (2, 3, 4, 1)]
    *ys, dlogp = flow(*xs)
    assert tr.seen == (torch.Size([5, 4]), torch.Size([5, 6]))
    assert [y.shape[-1] for y in ys] == [2, 3, 4, 1]
    assert torch.equal(ys[0], torch.ones(5, 2)) and torch.equal(ys[1], xs[1])
    *zs, dinv = flow(*ys, inverse=True)
    assert all(torch.equal(z, x) for z, x in zip(zs, xs)) and torch.equal(dlogp + dinv, torch.zeros(5, 1))
    with pytest.raises(ValueError):
        bg.CouplingFlow(tr, transformed_indices=(0, 1), cond_indices=(1,))


def test_sequential_inverse_swap_wrap_setconstant():
    tr = DummyTransformer()
    flow = bg.SequentialFlow([bg.SplitFlow(3), bg.CouplingFlow(tr), bg.SwapFlow(), bg.CouplingFlow(tr),
                              bg.SwapFlow(), bg.MergeFlow(3)])
    x = torch.randn(4, 7)
    y, dlogp = flow(x)
    assert y.shape == (4, 7) and dlogp.shape == (4, 1) and torch.equal(dlogp, 2 * torch.ones(4, 1))
    xb, dinv = flow(y, inverse=True)
    torch.testing.assert_close(xb, x)
    torch.testing.assert_close(dlogp + dinv, torch.zeros(4, 1))
    inv = bg.InverseFlow(flow)
    xb2, dinv2 = inv(y)
    torch.testing.assert_close(xb2, x)
    assert len(flow) == 6 and isinstance(flow[1], bg.CouplingFlow) and len(flow[1:3]) == 2
    a, b, d = bg.SwapFlow()(torch.zeros(2, 1), torch.ones(2, 2))
    assert a.shape == (2, 2) and d.shape == (2, 1)
    with pytest.warns(UserWarning):
        bg.SwapFlow()(torch.zeros(2, 1))
    wrapped = bg.WrapFlow(bg.SwapFlow(), indices=(0, 2))
    p, q, r, d = wrapped(torch.zeros(2, 1), torch.ones(2, 2), 2 * torch.ones(2, 3))
    assert p.shape == (2, 3) and q.shape == (2, 2) and r.shape == (2, 1)
    p2, q2, r2, _ = wrapped(p, q, r, inverse=True)
    assert p2.shape == (2, 1) and r2.shape == (2, 3)
    const = bg.SetConstantFlow([1], [torch.tensor([7.0, 8.0])])
    a, c, d = const(torch.zeros(3, 2))
    assert c.shape == (3, 2) and torch.equal(c[0], torch.tensor([7.0, 8.0])) and d.shape == (3, 1)
    (a2, d2) = const(a, c, inverse=True)
    assert a2.shape == (3, 2)


def test_transformers_refuse_cpu_tensors():
    tr = bg.AffineTransformer(bg.DenseNet([2, 4, 3], activation=torch.nn.ReLU()))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tr.forward(torch.zeros(4, 2), torch.zeros(4, 3))
    sp = bg.ConditionalSplineTransformer(bg.DenseNet([2, 4, 3 * 25]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sp.forward(torch.zeros(4, 2), torch.zeros(4, 3))
    with pytest.raises(ValueError):
        bg.AffineTransformer(bg.DenseNet([2, 3]), bg.DenseNet([2, 3]), is_circular=True)


def test_state_dict_layout_matches_reference_densenet():
    net = bg.DenseNet([3, 8, 8, 5], activation=torch.nn.SiLU())
    assert list(net.state_dict().keys()) == ["_layers.0.weight", "_layers.0.bias", "_layers.2.weight",
                                             "_layers.2.bias", "_layers.4.weight", "_layers.4.bias"]
    assert net(torch.zeros(2, 3)).shape == (2, 5)
    wp = bg.WrapPeriodic(bg.DenseNet([5, 4]), indices=[0, 2])
    assert wp(torch.zeros(2, 3)).shape == (2, 4)


def test_zplan_staging():
    plan = engine.ZPlan(oic.ALA2_GLOBAL_Z)
    assert plan.seeds == [0, 1, 2] and len(plan.rel) == 19 and sorted(plan.order) == list(range(19))
    placed = set(plan.seeds)
    for r in plan.order:
        i, j, k, l = plan.rel[r]
        assert {j, k, l} <= placed
        placed.add(i)
    ref = oic.make_plan(oic.ALA2_GLOBAL_Z)
    assert plan.seeds == ref.seeds and np.array_equal(plan.rel, ref.rel)
    with pytest.raises(ValueError):
        engine.ZPlan(np.array([[0, -1, -1, -1], [1, 0, -1, -1], [2, 1, 0, -1], [3, 4, 1, 0], [4, 3, 1, 0]]))


def test_c_abi_exports_every_declared_symbol():
    lib = _lib.load()                       # builds the .so if it is missing; raises if it cannot
    header = open(os.path.join(ROOT, "include", "bgflow_b200.h")).read()
    declared = set(re.findall(r"\b(bgx_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.bgx_version().startswith(b"bgflow_b200")
    assert lib.bgx_launch_count() >= 0
    # size-only pack query runs without a GPU: 33 -> 128 -> 128 -> 825 spline conditioner
    src = _lib.bgx_mlp()
    src.n_layers = 3
    for i, d in enumerate((33, 128, 128, 825)):
        src.dims[i] = d
    src.raw_width = 33
    lay = _lib.bgx_spline_layout()
    lay.d_t, lay.n_bins = 33, 8
    out = _lib.bgx_packed_mlp()
    assert lib.bgx_pack_mlp(ctypes.byref(src), ctypes.byref(lay), None, 0, ctypes.byref(out), None) == 0
    assert out.spline_dims_per_pass == 5 and out.spline_stride == 25 and out.N[2] == 7 * 128
    assert out.total_floats > 128 * 896
    src.dims[3] = 826                       # wrong width: spline.py:112-121 raises
    assert lib.bgx_pack_mlp(ctypes.byref(src), ctypes.byref(lay), None, 0, ctypes.byref(out), None) == -1


def test_two_rank_gloo_sharding():
    """bench.py's multi-GPU plumbing (rank-sharded batch, max-over-ranks timing) on 2 CPU ranks."""
    env = dict(os.environ, BGX_BENCH_DRYRUN="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "bench.py"),
           "--gpus", "2", "--steps", "2", "--warmup", "1"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["n_gpus"] == 2 and rec["dryrun"] is True and rec["scaling"] == "weak"
    assert rec["shard_rows"] == [rec["config"]["batch_per_gpu"]] * 2


def test_gradient_allreduce_two_ranks():
    """bgflow_b200.distributed.allreduce_gradients on 2 gloo ranks == single-process gradient."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "_allreduce_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "ALLREDUCE_OK world=2" in res.stdout


def test_shard_rows():
    from bgflow_b200.distributed import shard_rows
    assert [shard_rows(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_rows(1 << 20, 7, 8) == (7 << 17, 8 << 17)


@pytest.mark.parametrize("normalize", [True, False])
@pytest.mark.parametrize("mol", ["ala2", "chain12"])
def test_ic_backward_definition_matches_oracle_values_and_gradients(mol, normalize):
    """``_torch_math_ic`` (the differentiable definition the IC backward re-evaluates) against the
    oracle in fp64 on the CPU: values and vector-Jacobian products, both directions."""
    from oracle import ic as oic
    from bgflow_b200 import _torch_math_ic as tm
    from bgflow_b200.engine import ZPlan
    z = oic.ALA2_GLOBAL_Z if mol == "ala2" else oic.chain_z_matrix(12)
    n = len(z)
    plan, oplan = ZPlan(z, normalize_angles=normalize), oic.make_plan(z)
    g = torch.Generator().manual_seed(3)
    if mol == "ala2":
        xyz = torch.as_tensor(oic.ALA2_XYZ).reshape(1, -1) + 0.01 * torch.randn(9, 3 * n, generator=g, dtype=torch.float64)
    else:
        xyz = torch.randn(9, 3 * n, generator=g, dtype=torch.float64)
    xyz.requires_grad_(True)
    ref = oic.xyz_to_ic(oplan, xyz, normalize_angles=normalize)
    got = tm.ic_from_xyz(plan, xyz)
    ws = [torch.randn(t.shape, generator=g, dtype=torch.float64) for t in ref]
    g_ref = torch.autograd.grad(sum((a * w).sum() for a, w in zip(ref, ws)), xyz)[0]
    g_got = torch.autograd.grad(sum((a * w).sum() for a, w in zip(got, ws)), xyz)[0]
    for a, b in zip(ref, got):
        torch.testing.assert_close(b, a, atol=1e-11, rtol=0)
    torch.testing.assert_close(g_got, g_ref, atol=1e-9, rtol=1e-10)
    ins = [t.detach().clone().requires_grad_(True) for t in ref[:5]]
    ref2 = oic.ic_to_xyz(oplan, *ins, normalize_angles=normalize)
    got2 = tm.ic_to_xyz(plan, *ins)
    ws = [torch.randn(t.shape, generator=g, dtype=torch.float64) for t in ref2]
    g_ref = torch.autograd.grad(sum((a * w).sum() for a, w in zip(ref2, ws)), ins)
    g_got = torch.autograd.grad(sum((a * w).sum() for a, w in zip(got2, ws)), ins)
    for a, b in zip(ref2, got2):
        torch.testing.assert_close(b, a, atol=1e-11, rtol=0)
    for a, b in zip(g_ref, g_got):
        torch.testing.assert_close(b, a, atol=1e-9, rtol=1e-10)


# ------------------------------------------------------------------ SURVEY 8f rows (host side, CPU)

@pytest.mark.parametrize("keep", [None, 9, 15])
def test_relative_ic_backward_definition_matches_oracle(keep):
    """``_torch_math_ic.relic_*`` (what the relative / mixed IC backward re-evaluates) against the
    oracle in fp64: values and vector-Jacobian products, both directions."""
    from oracle import ic as oic
    from bgflow_b200 import _torch_math_ic as tm
    import bgflow_b200 as bg
    g = torch.Generator().manual_seed(5)
    x0 = torch.as_tensor(oic.ALA2_XYZ).reshape(1, -1)
    xyz = x0 + 0.01 * torch.randn(7, 66, generator=g, dtype=torch.float64)
    oplan = oic.make_rel_plan(oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK)
    if keep is None:
        layer = bg.RelativeInternalCoordinateTransformation(oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK)
        fwd = lambda x: oic.rel_xyz_to_ic(oplan, x)
        inv = lambda *a: oic.rel_ic_to_xyz(oplan, *a)
    else:
        data = x0 + 0.02 * torch.randn(300, 66, generator=g, dtype=torch.float64)
        layer = bg.MixedCoordinateTransformation(data, oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK, keepdims=keep)
        white = oic.Whitening(data.reshape(-1, 22, 3)[:, oic.ALA2_RIGID_BLOCK].reshape(-1, 15).numpy(), keepdims=keep)
        assert layer.dim_fixed == keep and abs(white.jacobian_xz - layer._plan.whitening["jacobian_xz"]) < 1e-9
        fwd = lambda x: oic.mixed_xyz_to_ic(oplan, white, x)
        inv = lambda *a: oic.mixed_ic_to_xyz(oplan, white, *a)
    plan = layer._plan
    xyz.requires_grad_(True)
    ref, got = fwd(xyz), tm.relic_from_xyz(plan, xyz)
    ws = [torch.randn(t.shape, generator=g, dtype=torch.float64) for t in ref]
    g_ref = torch.autograd.grad(sum((a * w).sum() for a, w in zip(ref, ws)), xyz)[0]
    g_got = torch.autograd.grad(sum((a * w).sum() for a, w in zip(got, ws)), xyz)[0]
    for a, b in zip(ref, got):
        torch.testing.assert_close(b, a, atol=1e-10, rtol=0)
    torch.testing.assert_close(g_got, g_ref, atol=1e-8, rtol=1e-9)
    ins = [t.detach().clone().requires_grad_(True) for t in ref[:4]]
    ref2, got2 = inv(*ins), tm.relic_to_xyz(plan, *ins)
    ws = [torch.randn(t.shape, generator=g, dtype=torch.float64) for t in ref2]
    g_ref = torch.autograd.grad(sum((a * w).sum() for a, w in zip(ref2, ws)), ins)
    g_got = torch.autograd.grad(sum((a * w).sum() for a, w in zip(got2, ws)), ins)
    for a, b in zip(ref2, got2):
        torch.testing.assert_close(b, a, atol=1e-10, rtol=0)
    for a, b in zip(g_ref, g_got):
        torch.testing.assert_close(b, a, atol=1e-8, rtol=1e-9)
    # reference properties (tests/nn/flow/crd_transform/test_ic.py:452-475)
    assert layer.dim_bonds == layer.dim_angles == layer.dim_torsions == 17
    assert (layer.bond_indices == oic.ALA2_RELATIVE_Z[:, :2]).all()
    assert (layer.fixed_atoms == oic.ALA2_RIGID_BLOCK).all()


def test_relplan_validation():
    from bgflow_b200.engine import RelPlan
    z = np.array([[3, 2, 1, 0], [4, 3, 2, 1]])
    plan = RelPlan(z, np.array([0, 1, 2]))
    assert plan.order == [0, 1] and plan.n_atoms == 5 and plan.fixed_width == 9
    with pytest.raises(ValueError):
        RelPlan(np.array([[3, 2, 1, 4], [4, 3, 2, 1]]), np.array([0, 1, 2]))     # cyclic: 3 needs 4 needs 3
    with pytest.raises(ValueError):
        RelPlan(np.array([[3, 2, 1, 0]]), np.array([0, 1, 2, 3]))                 # atom 3 twice


def test_cdf_host_mirrors_match_oracle():
    """TruncatedNormalDistribution / SloppyUniform mirrors and the torch CDFTransform definition
    (generic path + backward) against the oracle in fp64 on the CPU."""
    import bgflow_b200 as bg
    from bgflow_b200 import cdf as bcdf
    from oracle import cdf as ocdf
    g = torch.Generator().manual_seed(2)
    n = 6
    mu = 1 + torch.rand(n, generator=g, dtype=torch.float64)
    sigma = 0.2 + torch.rand(n, generator=g, dtype=torch.float64)
    lower = torch.rand(n, generator=g, dtype=torch.float64)
    upper = 4 * torch.ones(n, dtype=torch.float64)
    mine = bg.TruncatedNormalDistribution(mu, sigma, lower, upper)
    ref = ocdf.TruncatedNormal(mu, sigma, lower, upper)
    u = torch.rand(50, n, generator=g, dtype=torch.float64)
    for inverse in (True, False):
        x = u if inverse else ref.icdf(u)
        a = bcdf._torch_cdf_transform(mine, x, inverse, 1e-7)
        b = ocdf.cdf_transform(ref, x, inverse=inverse)
        torch.testing.assert_close(a[0], b[0], atol=1e-12, rtol=0)
        torch.testing.assert_close(a[1], b[1], atol=1e-12, rtol=0)
    s = mine.sample(1000)
    assert s.shape == (1000, n) and (s >= lower).all() and (s <= upper).all()
    assert torch.isfinite(mine.energy(s)).all()
    with pytest.raises(ValueError):
        mine.energy(torch.full((1, n), 5.0, dtype=torch.float64))
    su = bg.SloppyUniform(torch.zeros(3, dtype=torch.float64), 2 * torch.ones(3, dtype=torch.float64))
    ou = ocdf.Uniform(torch.zeros(3, dtype=torch.float64), 2 * torch.ones(3, dtype=torch.float64))
    x = torch.tensor([[0.5, -0.1, 2.0]], dtype=torch.float64)
    torch.testing.assert_close(su.cdf(x), ou.cdf(x))
    assert torch.equal(su.log_prob(x), ou.log_prob(x))
    # column specs for the kernel table
    cols = bcdf.marginal_columns(mine, n)
    assert len(cols) == n and cols[0][0] == _lib.DIST_TRUNCNORMAL and abs(cols[2][1] - float(mu[2])) < 1e-15
    assert bcdf.marginal_columns(torch.distributions.Normal(torch.zeros(2), torch.ones(2)), 2)[1] == (1, 0.0, 1.0, 0.0, 0.0)
    assert bcdf.marginal_columns(su, 3)[0][:3] == (3, 0.0, 2.0)
    assert bcdf.marginal_columns(torch.distributions.Beta(torch.ones(2), torch.ones(2)), 2) is None
    # kernel-backed marginals refuse CPU tensors like every other kernel-backed layer
    with pytest.raises(RuntimeError):
        bg.CDFTransform(mine)(torch.rand(4, n))
    # the C helper that derives the per-column constants runs on the host
    col = _lib.bgx_cdf_col()
    assert _lib.load().bgx_cdf_col_init(_lib.DIST_TRUNCNORMAL, 1.0, 1.0, 1e-5, float("inf"), ctypes.byref(col)) == 0
    assert abs(col.p[2] - 0.158657) < 1e-5 and abs(col.p[3] - 0.841343) < 1e-5
    assert _lib.load().bgx_cdf_col_init(_lib.DIST_TRUNCNORMAL, 1.0, -1.0, 0.0, 1.0, ctypes.byref(col)) == -1


def test_fuse_domain_maps_rewrites_builder_tail():
    """Structure only (no compute): the builder's tail (generator_builder.py:408-459) becomes
    [constants, one multi-tensor icdf launch for the non-IC fields, MappedICTail]."""
    import bgflow_b200 as bg
    from bgflow_b200 import cdf as bcdf
    from oracle import ic as oic
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    tn = lambda n, lo, hi: bg.TruncatedNormalDistribution(torch.ones(n), torch.ones(n), torch.tensor(lo), torch.tensor(hi))
    marg = [tn(21, 1e-5, float("inf")), tn(20, 1e-5, 1.0), bg.SloppyUniform(torch.zeros(19), torch.ones(19)),
            torch.distributions.Normal(torch.zeros(10), torch.ones(10))]
    head = bg.SwapFlow()
    layers = [head] + [bg.WrapFlow(bg.InverseFlow(bg.CDFTransform(m)), (i,)) for i, m in enumerate(marg)]
    layers += [bg.SetConstantFlow([4], [torch.zeros(1, 3)]), bg.SetConstantFlow([5], [torch.tensor([0.5, 0.5, 0.5])]),
               bg.WrapFlow(bg.InverseFlow(ic), indices=[0, 1, 2, 4, 5], out_indices=(0,))]
    fused = bg.fuse_domain_maps(bg.SequentialFlow(layers))
    kinds = [type(b).__name__ for b in fused]
    assert kinds == ["SwapFlow", "SetConstantFlow", "SetConstantFlow", "InverseFlow", "WrapFlow"]
    multi = fused[3]._delegate
    assert isinstance(multi, bcdf.MultiCDFFlow) and multi._indices == [3]
    tail = fused[4]._flow
    assert isinstance(tail, bcdf.MappedICTail) and tail._marginals[2] is marg[2]
    assert fused[4]._accumulates_dlogp and fused[3]._accumulates_dlogp
    # a generic (non kernel-backed) marginal on an IC field keeps the IC layer unfused
    layers[2] = bg.WrapFlow(bg.InverseFlow(bg.CDFTransform(torch.distributions.Beta(torch.ones(20), torch.ones(20)))), (1,))
    fused2 = bg.fuse_domain_maps(bg.SequentialFlow(layers))
    assert [type(b).__name__ for b in fused2] == ["SwapFlow", "SetConstantFlow", "SetConstantFlow", "InverseFlow", "WrapFlow"]
    assert isinstance(fused2[4]._flow, bg.InverseFlow) and fused2[3]._delegate._indices == [0, 1, 2, 3]
    # nothing to fuse: unchanged block list
    plain = bg.SequentialFlow([bg.SwapFlow(), bg.SwapFlow()])
    assert len(bg.fuse_domain_maps(plain)) == 2
