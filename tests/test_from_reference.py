"""``bgflow_b200.from_reference``: switching an existing reference object graph to the kernel-backed
mirror classes (CPU-only checks of structure, parameter sharing and plans; the kernels behind the
mirror classes are covered by the ``-m gpu`` parity tests).  Needs the reference checkout, which
exists in the build container only: skipped elsewhere."""

import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "bgflow")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    np.infty = np.inf                                  # what oracle/_stubs/shim/sitecustomize.py does
    for p in (os.path.join(ROOT, "oracle", "_stubs"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import bgflow
    return bgflow


def test_coupling_stack_shares_parameters(ref):
    import bgflow_b200 as bg
    torch.manual_seed(0)
    layers = [ref.SplitFlow(33)]
    for i in range(3):
        net = ref.DenseNet([33, 128, 128, 33 * 25], activation=torch.nn.SiLU())
        layers += [ref.CouplingFlow(ref.ConditionalSplineTransformer(net, is_circular=False)), ref.SwapFlow()]
    shift = ref.DenseNet([33, 64, 33], activation=torch.nn.ReLU())
    scale = ref.DenseNet([33, 64, 33], activation=torch.nn.ReLU())
    layers += [ref.CouplingFlow(ref.AffineTransformer(shift, scale)), ref.MergeFlow(33)]
    rflow = ref.SequentialFlow(layers)
    flow = bg.from_reference(rflow)
    assert [type(b).__name__ for b in flow] == [type(b).__name__ for b in rflow]
    assert all(type(b).__module__.startswith("bgflow_b200") for b in flow)
    # the very same parameter tensors, the same state_dict keys
    assert all(a is b for a, b in zip(flow.parameters(), rflow.parameters()))
    assert list(flow.state_dict()) == list(rflow.state_dict())
    t = flow[1].transformer
    assert isinstance(t, bg.ConditionalSplineTransformer) and isinstance(t._params_net, bg.DenseNet)
    assert t._params_net._layers is rflow[1].transformer._params_net._layers
    aff = flow[7].transformer
    assert isinstance(aff, bg.AffineTransformer) and aff._log_alpha is rflow[7].transformer._log_alpha
    # an optimiser step on the reference parameters is seen by the mirror's packed-weight cache key
    w = rflow[1].transformer._params_net._layers[0].weight
    v0 = w._version
    with torch.no_grad():
        w.add_(1.0)
    assert flow[1].transformer._params_net._layers[0].weight._version == v0 + 1
    # kernel-backed: refuses CPU tensors exactly like a natively built mirror flow
    with pytest.raises(RuntimeError):
        flow(torch.rand(4, 66))


def test_builder_tail_is_fused_and_plans_agree(ref):
    import bgflow_b200 as bg
    from bgflow_b200 import cdf as bcdf
    from oracle import ic as oic
    ic = ref.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    one = lambda n, v=1.0: torch.full((n,), v)
    marg = [ref.TruncatedNormalDistribution(one(21), one(21), torch.tensor(1e-5), torch.tensor(np.inf)),
            ref.TruncatedNormalDistribution(one(20, 0.5), one(20), torch.tensor(1e-5), torch.tensor(1.0)),
            ref.SloppyUniform(torch.zeros(19), one(19)),
            torch.distributions.Normal(torch.zeros(10), one(10))]
    net = ref.WrapPeriodic(ref.DenseNet([38, 128, 128, 250], activation=torch.nn.SiLU()), indices=list(range(19)))
    layers = [ref.CouplingFlow(ref.ConditionalSplineTransformer(net, is_circular=False), transformed_indices=(3,),
                               cond_indices=(2,))]
    layers += [ref.WrapFlow(ref.InverseFlow(ref.CDFTransform(m)), (i,)) for i, m in enumerate(marg)]
    layers += [ref.SetConstantFlow([4], [torch.zeros(1, 3)]), ref.SetConstantFlow([5], [torch.tensor([0.5, 0.5, 0.5])]),
               ref.WrapFlow(ref.InverseFlow(ic), indices=[0, 1, 2, 4, 5], out_indices=(0,))]
    rflow = ref.SequentialFlow(layers)
    flow = bg.from_reference(rflow)
    kinds = [type(b).__name__ for b in flow]
    assert kinds == ["CouplingFlow", "SetConstantFlow", "SetConstantFlow", "InverseFlow", "WrapFlow"]
    assert isinstance(flow[0].transformer._params_net, bg.WrapPeriodic)
    tail = flow[4]._flow
    assert isinstance(tail, bcdf.MappedICTail) and tail._marginals[0] is marg[0]
    assert flow[3]._delegate._indices == [3]
    plan = tail._ic._plan
    assert plan.seeds == [int(a) for a in ic._rel_ic.fixed_atoms]
    assert np.array_equal(plan.rel, ic.z_matrix) and plan.normalize_angles == ic.normalize_angles
    assert (tail._ic.bond_indices == ic.bond_indices).all() and (tail._ic.torsion_indices == ic.torsion_indices).all()
    # without fusing: block for block
    plain = bg.from_reference(rflow, fuse_tail=False)
    assert [type(b).__name__ for b in plain] == [type(b).__name__ for b in rflow]
    assert isinstance(plain[1]._flow._delegate, bg.CDFTransform)


def test_relative_mixed_generator_and_fallthrough(ref):
    import bgflow_b200 as bg
    from oracle import ic as oic
    g = torch.Generator().manual_seed(0)
    data = torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float32).reshape(1, -1) + 0.02 * torch.randn(200, 66, generator=g)
    rmixed = ref.MixedCoordinateTransformation(data, oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK, keepdims=9)
    mixed = bg.from_reference(rmixed)
    assert isinstance(mixed, bg.MixedCoordinateTransformation) and mixed.dim_fixed == 9
    w = mixed._plan.whitening
    np.testing.assert_array_equal(w["whiten"], rmixed._whiten.Twhiten.numpy())
    assert w["jacobian_xz"] == pytest.approx(float(rmixed._whiten.jacobian_xz))
    rrel = ref.RelativeInternalCoordinateTransformation(oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK, normalize_angles=False)
    rel = bg.from_reference(rrel)
    assert isinstance(rel, bg.RelativeInternalCoordinateTransformation) and not rel.normalize_angles
    assert (rel.fixed_atoms == oic.ALA2_RIGID_BLOCK).all()

    class Odd(ref.Flow):                                 # a flow this package knows nothing about
        def _forward(self, *xs, **kw):
            return (*xs, torch.zeros(xs[0].shape[0], 1))

        def _inverse(self, *xs, **kw):
            return (*xs, torch.zeros(xs[0].shape[0], 1))
    rflow = ref.SequentialFlow([Odd(), ref.SwapFlow()])
    flow = bg.from_reference(rflow)
    assert isinstance(flow[0], Odd) and isinstance(flow[1], bg.SwapFlow)
    a, b, d = flow(torch.ones(3, 2), torch.zeros(3, 2))    # reference block inside the mirror's SequentialFlow
    assert torch.equal(a, torch.zeros(3, 2)) and d.shape == (3, 1)
    with pytest.raises(NotImplementedError):
        bg.from_reference(rflow, strict=True)
    prior = ref.NormalDistribution(2)
    gen = bg.from_reference(ref.BoltzmannGenerator(prior, rflow, None))
    assert isinstance(gen, bg.BoltzmannGenerator) and gen.prior is prior and isinstance(gen.flow, bg.SequentialFlow)


def test_real_builder_generator_converts(ref):
    """BASELINE config 4 built by the REFERENCE's own BoltzmannGeneratorBuilder
    (tests/factory/test_generator_builder.py:45-66, OpenMM-free z-matrix, no target): every coupling
    becomes a kernel-backed one with the conditioner shapes SURVEY.md 8d lists, and the builder tail
    (icdf maps + constants + map to Cartesian) collapses to [constants, one multi-field icdf, MappedICTail]."""
    import bgflow_b200 as bg
    from bgflow_b200 import cdf as bcdf
    from bgflow import BoltzmannGeneratorBuilder, ShapeDictionary, BONDS, ANGLES, TORSIONS, AUGMENTED
    from oracle import ic as oic
    crd = ref.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    shape_info = ShapeDictionary.from_coordinate_transform(crd, dim_augmented=10)
    b = BoltzmannGeneratorBuilder(shape_info, target=None, device=torch.device("cpu"), dtype=torch.float32)
    for _ in range(4):
        b.add_condition(TORSIONS, on=AUGMENTED)
        b.add_condition(AUGMENTED, on=TORSIONS)
    for _ in range(2):
        b.add_condition(BONDS, on=ANGLES)
        b.add_condition(ANGLES, on=BONDS)
    b.add_condition(ANGLES, on=(TORSIONS, AUGMENTED))
    b.add_condition(BONDS, on=(ANGLES, TORSIONS, AUGMENTED))
    b.add_map_to_ic_domains()
    b.add_map_to_cartesian(crd)
    gen = b.build_generator()
    fast = bg.from_reference(gen)
    blocks = list(fast.flow)
    couplings = [blk for blk in blocks if isinstance(blk, bg.CouplingFlow)]
    assert len(couplings) == 14 and all(isinstance(c.transformer, bg.ConditionalSplineTransformer) for c in couplings)
    shapes = []
    for c in couplings:
        net = c.transformer._params_net
        inner = net.net if isinstance(net, bg.WrapPeriodic) else net
        assert isinstance(inner, bg.DenseNet)
        lin = [m for m in inner._layers if isinstance(m, torch.nn.Linear)]
        assert [l.out_features for l in lin[:-1]] == [128, 128]
        shapes.append((lin[0].in_features, lin[-1].out_features, isinstance(net, bg.WrapPeriodic)))
    assert shapes == ([(10, 456, False), (38, 250, True)] * 4 + [(20, 525, False), (21, 500, False)] * 2
                      + [(48, 500, True), (68, 525, True)])
    assert [type(x).__name__ for x in blocks[14:]] == ["SetConstantFlow", "SetConstantFlow", "InverseFlow", "WrapFlow"]
    assert isinstance(blocks[16]._delegate, bcdf.MultiCDFFlow) and blocks[16]._delegate._indices == [3]
    assert isinstance(blocks[17]._flow, bcdf.MappedICTail)
    assert all(a is b_ for a, b_ in zip(fast.flow.parameters(), gen.flow.parameters()))
    assert fast.prior is gen.prior


def test_plumbing_matches_reference_on_random_layouts(ref):
    """The mirror's tuple plumbing (WrapFlow routing, SplitFlow by sizes / indices, MergeFlow,
    SetConstantFlow, SwapFlow inside SequentialFlow, both directions) against the reference's own
    classes on random layouts — CPU tensors, generic inner flows, bit-for-bit."""
    import bgflow_b200 as bg
    rng = np.random.default_rng(0)

    class Scale(ref.Flow):                        # generic inner flow: multiplies tensor i by (i + 2)
        def _forward(self, *xs, **kw):
            return (*[x * (i + 2) for i, x in enumerate(xs)], torch.full((xs[0].shape[0], 1), float(len(xs))))

        def _inverse(self, *xs, **kw):
            return (*[x / (i + 2) for i, x in enumerate(xs)], torch.full((xs[0].shape[0], 1), -float(len(xs))))

    def same(a, b):
        assert len(a) == len(b)
        for u, v in zip(a, b):
            assert torch.equal(u, v), (u, v)

    for trial in range(40):
        n = int(rng.integers(2, 6))
        xs = [torch.randn(5, int(rng.integers(1, 4))) for _ in range(n)]
        k = int(rng.integers(1, n + 1))
        idx = [int(i) for i in rng.permutation(n)[:k]]
        out_idx = None if trial % 2 else [int(i) for i in rng.permutation(n)[:k]]
        r = ref.WrapFlow(Scale(), idx, out_idx)
        m = bg.WrapFlow(Scale(), idx, out_idx)
        same(r(*xs), m(*xs))
        ys = r(*xs)[:-1]
        same(r(*ys, inverse=True), m(*ys, inverse=True))
        # SequentialFlow of wrap + swap + set-constant, forward then inverse
        const = torch.randn(1, 2)
        pos = int(rng.integers(0, n + 1))
        rseq = ref.SequentialFlow([r, ref.SwapFlow(), ref.SetConstantFlow([pos], [const])])
        mseq = bg.SequentialFlow([m, bg.SwapFlow(), bg.SetConstantFlow([pos], [const])])
        fr, fm = rseq(*xs), mseq(*xs)
        same(fr, fm)
        same(rseq(*fr[:-1], inverse=True), mseq(*fm[:-1], inverse=True))
    for trial in range(20):
        d = int(rng.integers(4, 12))
        x = torch.randn(3, d)
        cuts = sorted(int(c) for c in rng.choice(np.arange(1, d), size=int(rng.integers(1, 3)), replace=False))
        sizes = [b - a for a, b in zip([0] + cuts, cuts + [d])]
        use = sizes if trial % 2 else sizes[:-1]               # the last size may be omitted
        same(ref.SplitFlow(*use)(x), bg.SplitFlow(*use)(x))
        parts = ref.SplitFlow(*use)(x)[:-1]
        same(ref.MergeFlow(*use)(*parts), bg.MergeFlow(*use)(*parts))
        perm = rng.permutation(d)
        groups = [[int(i) for i in perm[:cuts[0]]], [int(i) for i in perm[cuts[0]:]]]
        same(ref.SplitFlow(*groups)(x), bg.SplitFlow(*groups)(x))
        gparts = ref.SplitFlow(*groups)(x)[:-1]
        same(ref.SplitFlow(*groups)(*gparts, inverse=True), bg.SplitFlow(*groups)(*gparts, inverse=True))
