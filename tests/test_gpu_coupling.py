"""Parity of the fused coupling kernels (through the C ABI) against the oracle and the
reference-generated golden fixtures.  Tolerances (fp32 path, stated per SURVEY.md 8c):
  y: 1e-5 abs+rel per block, 1e-4 over an 8-block stack;  dlogp: 1e-3 abs over a stack.
"""

import numpy as np
import pytest
import torch

import bgflow_b200 as bg
from oracle import flows as of
from conftest import load_golden
from helpers import stack_from, transformer_from
from test_oracle_golden import _multi_blocks

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(params=["tc", "tc6", "simt"], autouse=True)
def kernel_mode(request):
    """Every parity test runs on the tensor-core kernels (spline and affine, where the shape is
    eligible; both precisions) and on the shape-general SIMT kernel."""
    from bgflow_b200 import engine
    old = dict(engine.config)
    engine.config.update(force_simt=(request.param == "simt"),
                         precision="bf16x6" if request.param == "tc6" else "bf16x3")
    yield request.param
    engine.check_pipeline_status(DEV)
    engine.config.update(old)


def _t(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def _cmp(a, b, atol, rtol=0.0, what=""):
    a = a.detach().cpu().double().numpy()
    b = np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, atol=atol, rtol=rtol, err_msg=what)


@pytest.mark.parametrize("name,kind", [
    ("affine_d66_8blk", "affine"), ("spline_d66_8blk", "spline"),
    ("affine_d10_3blk", "affine"), ("spline_d7_4blk", "spline")])
def test_stack_matches_reference_golden(name, kind):
    g = load_golden(name)
    dim, n_blocks, batch, seed = (int(v) for v in g["meta"][:4])
    hidden = tuple(int(v) for v in g["meta"][4:])
    blocks, split = of.make_stack(kind, dim, n_blocks, hidden=hidden, seed=seed)
    flow = stack_from(blocks, split, DEV)
    with torch.no_grad():
        x, dlogp = flow(_t(g["z_f32"]))
        assert x.shape == (batch, dim) and dlogp.shape == (batch, 1)
        # against the reference's fp32 run and its fp64 run (the truth)
        _cmp(x, g["x_f32"], 1e-4, 1e-4, "x vs ref fp32")
        _cmp(x, g["x_f64"], 1e-4, 1e-4, "x vs ref fp64")
        _cmp(dlogp, g["dlogp_f64"], 1e-3, 1e-4, "dlogp vs ref fp64")
        zi, dlogpi = flow(_t(g["x_f32"]), inverse=True)
        _cmp(zi, g["zi_f64"], 1e-4, 1e-4, "inverse")
        _cmp(dlogpi, g["dlogpi_f64"], 1e-3, 1e-4, "inverse dlogp")
        # round trip through our own forward output
        zb, dlb = flow(x, inverse=True)
        _cmp(zb, g["z_f64"], 2e-4, 1e-4, "round trip")
        _cmp(dlb + dlogp, np.zeros((batch, 1)), 1e-3, 0, "dlogp fwd + inv")


def test_multi_tensor_coupling_golden():
    g = load_golden("multi_tensor_coupling")
    blocks = _multi_blocks(torch.float32)
    layers = [bg.CouplingFlow(transformer_from(b, DEV), transformed_indices=b["transformed"],
                              cond_indices=b["cond"]) for b in blocks]
    flow = bg.SequentialFlow(layers)
    xs = [_t(g[f"in{i}_f32"]) for i in range(4)]
    with torch.no_grad():
        *ys, dlogp = flow(*xs)
        for i in range(4):
            _cmp(ys[i], g[f"out{i}_f64"], 5e-5, 1e-4, f"out{i}")
        _cmp(dlogp, g["dlogp_f64"], 1e-3, 1e-4)
        *zs, dlogpi = flow(*[_t(g[f"out{i}_f32"]) for i in range(4)], inverse=True)
        for i in range(4):
            _cmp(zs[i], g[f"back{i}_f64"], 5e-5, 1e-4, f"back{i}")
        _cmp(dlogpi, g["dlogpi_f64"], 1e-3, 1e-4)


def test_readme_config_golden():
    """BASELINE config 1 (README.md:54-96), run through the kernels instead of on the CPU."""
    g = load_golden("readme_doublewell")
    gen = torch.Generator().manual_seed(0)
    blk = {"kind": "affine", "shift": of.make_mlp([1, 4, 1], "relu", gen), "scale": of.make_mlp([1, 4, 1], "tanh", gen)}
    flow = bg.SequentialFlow([bg.SplitFlow(1), bg.CouplingFlow(transformer_from(blk, DEV)),
                              bg.InverseFlow(bg.SplitFlow(1))])
    prior = bg.NormalDistribution(2).to(DEV)
    gen_bg = bg.BoltzmannGenerator(prior, flow, None)
    with torch.no_grad():
        x, dlogp = flow(_t(g["z_f32"]))
        _cmp(x, g["x_f64"], 1e-5, 1e-5)
        _cmp(dlogp, g["dlogp_f64"], 1e-5, 1e-5)
        nll = gen_bg.energy(_t(g["x_f32"]))
        assert nll.shape == (1024, 1)
        _cmp(nll, g["nll_f64"], 1e-4, 1e-5)
        s = gen_bg.sample(1024)
        assert s.shape == (1024, 2)


@pytest.mark.parametrize("kind", ["affine", "spline"])
@pytest.mark.parametrize("batch", [1, 63, 64, 65, 1000])
def test_ragged_batches_against_oracle(kind, batch):
    dim = 13
    blocks, split = of.make_stack(kind, dim, 2, hidden=(40,), seed=5, n_bins=5)
    flow = stack_from(blocks, split, DEV)
    g = torch.Generator().manual_seed(batch)
    z = torch.rand(batch, dim, generator=g) if kind == "spline" else torch.randn(batch, dim, generator=g)
    blocks64, _ = of.make_stack(kind, dim, 2, hidden=(40,), seed=5, n_bins=5, dtype=torch.float64)
    x_ref, d_ref = of.coupling_stack(blocks64, z.double(), split)
    with torch.no_grad():
        x, d = flow(z.to(DEV))
    _cmp(x, x_ref, 2e-5, 1e-5)
    _cmp(d, d_ref, 2e-4, 1e-5)


def test_empty_batch():
    blocks, split = of.make_stack("spline", 6, 1, hidden=(8,), seed=1, n_bins=4)
    flow = stack_from(blocks, split, DEV)
    x, d = flow(torch.rand(0, 6, device=DEV))
    assert x.shape == (0, 6) and d.shape == (0, 1)


@pytest.mark.parametrize("dim,hidden", [(384, (128, 128)), (66, (256, 256)), (66, (100,)), (40, ())])
def test_wide_and_odd_shapes_against_oracle(dim, hidden):
    for kind in ("affine", "spline"):
        blocks, split = of.make_stack(kind, dim, 2, hidden=hidden, seed=9)
        flow = stack_from(blocks, split, DEV)
        g = torch.Generator().manual_seed(3)
        z = torch.rand(200, dim, generator=g) if kind == "spline" else torch.randn(200, dim, generator=g)
        blocks64, _ = of.make_stack(kind, dim, 2, hidden=hidden, seed=9, dtype=torch.float64)
        x_ref, d_ref = of.coupling_stack(blocks64, z.double(), split)
        with torch.no_grad():
            x, d = flow(z.to(DEV))
            zb, db = flow(x, inverse=True)
        _cmp(x, x_ref, 5e-5, 1e-4, f"{kind} x")
        _cmp(d, d_ref, 1e-3, 1e-4, f"{kind} dlogp")
        _cmp(zb, z.double(), 1e-4, 1e-4, f"{kind} round trip")


def test_spline_properties_like_reference_tests():
    # tests/nn/flow/transformer/test_spline.py:8-33: outputs inside (0,1); zero-parameter
    # network == identity (tests/factory/test_generator_builder.py:131-136)
    net = bg.DenseNet([3, 16, 2 * 25], activation=torch.nn.Tanh()).to(DEV)
    tr = bg.ConditionalSplineTransformer(net, is_circular=False)
    x = torch.randn(50, 3, device=DEV)
    y = torch.rand(50, 2, device=DEV)
    z, dlogp = tr.forward(x, y)
    assert z.shape == (50, 2) and dlogp.shape == (50, 1)
    assert (z > 0).all() and (z < 1).all()
    with torch.no_grad():
        for p in net.parameters():
            p.zero_()
    z, dlogp = tr.forward(x, y)
    torch.testing.assert_close(z, y, atol=1e-6, rtol=0)
    torch.testing.assert_close(dlogp, torch.zeros_like(dlogp), atol=1e-5, rtol=0)
    # wrong conditioner width -> RuntimeError like spline.py:112-121
    bad = bg.ConditionalSplineTransformer(bg.DenseNet([3, 8, 2 * 25 + 1]).to(DEV))
    with pytest.raises(RuntimeError):
        bad.forward(x, y)


def test_out_of_domain_inputs_are_clamped_and_counted():
    net = bg.DenseNet([2, 8, 3 * 25], activation=torch.nn.SiLU()).to(DEV)
    tr = bg.ConditionalSplineTransformer(net)
    x = torch.randn(10, 2, device=DEV)
    y = torch.rand(10, 3, device=DEV)
    y_bad = y.clone()
    y_bad[0, 0] = -0.25
    y_bad[3, 2] = 1.5
    z_bad, d_bad = tr.forward(x, y_bad, inverse=True)
    z_ok, d_ok = tr.forward(x, y_bad.clamp(0, 1), inverse=True)
    torch.testing.assert_close(z_bad, z_ok)
    torch.testing.assert_close(d_bad, d_ok)
    with pytest.warns(UserWarning):
        assert tr.out_of_domain_count() == 2


def test_affine_transformer_api_like_reference_tests():
    # tests/nn/flow/transformer/test_affine.py:12-42
    shift = bg.DenseNet([2, 4, 3], activation=torch.nn.ReLU()).to(DEV)
    tr = bg.AffineTransformer(shift, is_circular=True).to(DEV)
    x, y = torch.randn(7, 2, device=DEV), torch.rand(7, 3, device=DEV)
    out, dlogp = tr.forward(x, y)
    assert out.shape == (7, 3) and dlogp.shape == (7, 1)
    assert (out >= 0).all() and (out < 1).all() and (dlogp == 0).all()
    with pytest.raises(ValueError):
        bg.AffineTransformer(shift, scale_transformation=shift, is_circular=True)
    with pytest.raises(RuntimeError):
        tr.forward(x.cpu(), y.cpu())       # no CPU fallback, by design


def test_precision_modes_report(kernel_mode):
    """Both tensor-core precisions against the reference's fp64 run of the 8-block golden stack
    (the numbers quoted in DESIGN.md); both sit far inside the 1e-4 / 1e-3 parity tolerance."""
    from bgflow_b200 import engine
    if kernel_mode != "tc":
        pytest.skip("runs once")
    g = load_golden("spline_d66_8blk")
    blocks, split = of.make_stack("spline", 66, 8, hidden=(128, 128), seed=0)
    flow = stack_from(blocks, split, DEV)
    errs = {}
    for prec in ("bf16x3", "bf16x6"):
        engine.config["precision"] = prec
        with torch.no_grad():
            x, dlogp = flow(_t(g["z_f32"]))
        errs[prec] = ((x.cpu().double() - torch.from_numpy(g["x_f64"])).abs().max().item(),
                      (dlogp.cpu().double() - torch.from_numpy(g["dlogp_f64"])).abs().max().item())
    engine.config.update(force_simt=True)
    with torch.no_grad():
        x, dlogp = flow(_t(g["z_f32"]))
    errs["simt fp32"] = ((x.cpu().double() - torch.from_numpy(g["x_f64"])).abs().max().item(),
                         (dlogp.cpu().double() - torch.from_numpy(g["dlogp_f64"])).abs().max().item())
    print("max abs error vs reference fp64 (x, dlogp):", {k: (f"{v[0]:.2e}", f"{v[1]:.2e}") for k, v in errs.items()})
    for k, (ex, ed) in errs.items():
        assert ex < 1e-4 and ed < 1e-3, (k, ex, ed)


@pytest.mark.parametrize("kind", ["spline", "affine"])
@pytest.mark.parametrize("batch", [4, 132, 1000, 1001, 4099, 76036])
def test_tensor_core_kernels_partial_tiles(kind, batch, kernel_mode):
    """hidden = 128 stacks on ragged batches: the 2-CTA kernels (batch % 4 == 0, partial last tile;
    76036 rows = more than two tiles per CTA), the single-CTA kernels (other batches / bf16x6) and the
    SIMT kernel must all agree with the oracle."""
    blocks, split = of.make_stack(kind, 66, 3, hidden=(128, 128), seed=11)
    flow = stack_from(blocks, split, DEV)
    g = torch.Generator().manual_seed(batch)
    z = torch.rand(batch, 66, generator=g) if kind == "spline" else torch.randn(batch, 66, generator=g)
    blocks64, _ = of.make_stack(kind, 66, 3, hidden=(128, 128), seed=11, dtype=torch.float64)
    x_ref, d_ref = of.coupling_stack(blocks64, z.double(), split)
    with torch.no_grad():
        x, d = flow(z.to(DEV))
        zb, db = flow(x, inverse=True)
    _cmp(x, x_ref, 3e-5, 1e-4)
    _cmp(d, d_ref, 5e-4, 1e-4)
    _cmp(zb, z.double(), 1e-4, 1e-4)
    _cmp(db, -d_ref, 5e-4, 1e-4)


@pytest.mark.parametrize("batch", [300, 4096])
def test_builder_style_stack_hidden128_on_tensor_cores(batch, kernel_mode):
    """Builder-exact Ala2 couplings (TORSIONS[17, circular] <-> FIXED[9],
    BONDS[17] <-> ANGLES[17]; hidden (128,128) SiLU; WrapPeriodic conditioners on torsions): several
    state tensors, periodic inputs, circular splines -> the one-CTA kernel's generic I/O path (300 rows)
    or the gather-then-two-CTA path (4096 rows)."""
    g = torch.Generator().manual_seed(21)
    nb = 8
    T, F, Bd, A = 17, 9, 17, 17
    def net(d_in, d_out, periodic=None):
        m = of.make_mlp([d_in, 128, 128, d_out], "silu", g)
        m.periodic = periodic
        return m
    blocks = [
        {"kind": "spline", "transformed": (2,), "cond": (3,), "is_circular": True, "params_net": net(F, T * 3 * nb)},
        {"kind": "spline", "transformed": (3,), "cond": (2,), "params_net": net(2 * T, F * (3 * nb + 1), (list(range(T)), 0.0, 1.0))},
        {"kind": "spline", "transformed": (0,), "cond": (1,), "params_net": net(A, Bd * (3 * nb + 1))},
        {"kind": "spline", "transformed": (1,), "cond": (0, 2), "params_net": net(Bd + 2 * T, A * (3 * nb + 1), (list(range(Bd, Bd + T)), 0.0, 1.0))},
    ]
    layers = [bg.CouplingFlow(transformer_from(b, DEV), transformed_indices=b["transformed"], cond_indices=b["cond"])
              for b in blocks]
    flow = bg.SequentialFlow(layers)
    gd = torch.Generator().manual_seed(22)
    xs = [torch.rand(batch, w, generator=gd) for w in (Bd, A, T, F)]
    ref = [x.double() for x in xs]
    blocks64 = []
    for b in blocks:
        nb64 = dict(b)
        m = b["params_net"]
        nb64["params_net"] = of.MLP([w.double() for w in m.weights], [x.double() for x in m.biases], m.act, m.periodic)
        blocks64.append(nb64)
    dref = 0
    for b in blocks64:
        ref, d = of.coupling_block(b, ref)
        dref = dref + d
    with torch.no_grad():
        *ys, dlogp = flow(*[x.to(DEV) for x in xs])
        *zs, dinv = flow(*ys, inverse=True)
    for got, want in zip(ys, ref):
        _cmp(got, want, 3e-5, 1e-4)
    _cmp(dlogp, dref, 5e-4, 1e-4)
    for got, want in zip(zs, xs):
        _cmp(got, want.double(), 1e-4, 1e-4)


@pytest.mark.parametrize("batch", [1, 63, 64, 65, 1000, 70001])
def test_split_merge_kernel_matches_torch(batch):
    """SplitFlow / MergeFlow by sizes run as one launch each (bgx_split_merge): bit-exact copies."""
    from bgflow_b200 import _lib, engine
    g = torch.Generator().manual_seed(batch)
    x = torch.randn(batch, 66, generator=g).to(DEV)
    n0 = _lib.launch_count()
    a, b = bg.SplitFlow(33)._apply_tuple((x,), False)
    assert _lib.launch_count() == n0 + 1
    assert a.is_contiguous() and b.is_contiguous()
    assert torch.equal(a, x[:, :33]) and torch.equal(b, x[:, 33:])
    (back,) = bg.SplitFlow(33)._apply_tuple((b, a), True)
    assert torch.equal(back, torch.cat([b, a], dim=-1))
    # several parts, a strided source (column window of a wider tensor), leading batch dims
    wide = torch.randn(batch, 100, generator=g).to(DEV)
    src = wide[:, 7:7 + 40]
    parts = engine.split_cols(src, [1, 17, 22])
    for p, ref in zip(parts, torch.split(src, [1, 17, 22], dim=-1)):
        assert torch.equal(p, ref)
    assert torch.equal(engine.merge_cols(parts), src)
    x3 = x.reshape(1, batch, 66)
    *ys, d = bg.SplitFlow(10, 20)(x3)
    assert [tuple(y.shape) for y in ys] == [(1, batch, 10), (1, batch, 20), (1, batch, 36)] and d.shape == (1, batch, 1)
    p3 = bg.SplitFlow(10, 20)._apply_tuple((x3,), False)
    assert all(torch.equal(p, r) for p, r in zip(p3, torch.split(x3, [10, 20, 36], dim=-1)))
