"""The pair kernel (bgx_coupling_pair.cu): narrow (tiles in shared memory) and wide (BASELINE config 5:
D = 384 / 3072, tiles accessed in place) against the fp64 oracle and the reference goldens, both
directions, ragged / odd tile counts, and proof that the TENSOR-CORE kernel served the call
(``bgx_kernel_count``), not the SIMT fallback.

Tolerances as tests/test_gpu_coupling.py: y 1e-5 abs+rel per block, dlogp 1e-3 over a stack (the log-det of a
D_t = 1536 block sums 1536 terms: 2e-3 there)."""

import numpy as np
import pytest
import torch

from bgflow_b200 import _lib, engine
from oracle import flows as of
from conftest import load_golden
from helpers import stack_from

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _status_check():
    old = dict(engine.config)
    yield
    engine.check_pipeline_status(DEV)
    engine.config.clear()
    engine.config.update(old)


def _delta(before, after):
    return {k: after[k] - before[k] for k in after if after[k] != before[k]}


def _run_vs_oracle(dim, n_blocks, batch, seed, ytol, dtol, kind="spline", hidden=(128, 128)):
    blocks, split = of.make_stack(kind, dim, n_blocks, seed=seed, hidden=hidden)
    blocks64, _ = of.make_stack(kind, dim, n_blocks, seed=seed, hidden=hidden, dtype=torch.float64)
    flow = stack_from(blocks, split, DEV)
    g = torch.Generator().manual_seed(seed + batch)
    z = torch.rand(batch, dim, generator=g) if kind == "spline" else torch.randn(batch, dim, generator=g)
    x_ref, d_ref = of.coupling_stack(blocks64, z.double(), split)
    zi_ref, di_ref = of.coupling_stack(blocks64, z.double(), split, inverse=True)
    with torch.no_grad():
        x, d = flow(z.to(DEV))
        zi, di = flow(z.to(DEV), inverse=True)
    np.testing.assert_allclose(x.cpu().double().numpy(), x_ref.numpy(), atol=ytol, rtol=ytol, err_msg="forward x")
    np.testing.assert_allclose(d.cpu().double().numpy(), d_ref.numpy(), atol=dtol, rtol=1e-5, err_msg="forward dlogp")
    np.testing.assert_allclose(zi.cpu().double().numpy(), zi_ref.numpy(), atol=ytol, rtol=ytol, err_msg="inverse x")
    np.testing.assert_allclose(di.cpu().double().numpy(), di_ref.numpy(), atol=dtol, rtol=1e-5, err_msg="inverse dlogp")


@pytest.mark.parametrize("dim,batch", [(384, 300), (384, 1), (384, 129), (3072, 260), (66, 515), (130, 1000)])
def test_wide_mode_against_fp64_oracle(dim, batch):
    """Config 5 shapes: conditioner 192-128-128-4800 / 1536-128-128-38400 (layer 0 k-tiled in groups of 128,
    last layer 39 / 308 passes), ragged batches (not multiples of 4 or 128), odd tile counts."""
    engine.config["spline_kernel"] = "pair_wide" if dim <= 130 else "auto"
    before = _lib.kernel_counts()
    _run_vs_oracle(dim, 2, batch, seed=3, ytol=2e-5, dtol=2e-3 if dim > 1000 else 1e-3)
    got = _delta(before, _lib.kernel_counts())
    assert got == {"spline_pair_wide": 4}, got      # 2 blocks x 2 directions, all on the tensor-core kernel


@pytest.mark.parametrize("dim,batch,hidden", [(384, 300, (128, 128, 128)), (384, 1, (128, 128)), (3072, 131, (128, 128, 128)),
                                              (130, 1000, (128,)), (70, 515, (128, 128))])
def test_affine_wide_against_fp64_oracle(dim, batch, hidden):
    """RealNVP blocks of BASELINE config 5 on the affine pair kernel: shift and scale nets in the two
    tensor-memory slots of one tile, layer 0 k-tiled, last layer 2 / 12 passes; ragged batches, D_t not a multiple
    of 4 (scalar tile access) or of 32 (partial column groups)."""
    engine.config["affine_kernel"] = "pair"
    before = _lib.kernel_counts()
    _run_vs_oracle(dim, 2, batch, seed=7, ytol=5e-5, dtol=1e-3, kind="affine", hidden=hidden)
    got = _delta(before, _lib.kernel_counts())
    assert got == {"affine_pair_wide": 4}, got


@pytest.mark.parametrize("batch", [4, 128, 132, 256 + 128 + 8, 4096 + 4])
def test_narrow_mode_against_fp64_oracle(batch):
    """D = 66 (33 | 33): tiles staged in shared memory by bulk TMA; one / two / odd numbers of tiles."""
    engine.config["spline_kernel"] = "pair"
    before = _lib.kernel_counts()
    _run_vs_oracle(66, 3, batch, seed=5, ytol=2e-5, dtol=1e-3)
    got = _delta(before, _lib.kernel_counts())
    assert got == {"spline_pair": 6}, got


@pytest.mark.parametrize("mode", ["pair", "pair_wide"])
def test_reference_golden_on_the_pair_kernel(mode):
    engine.config["spline_kernel"] = mode
    g = load_golden("spline_d66_8blk")
    dim, n_blocks, batch, seed = (int(v) for v in g["meta"][:4])
    hidden = tuple(int(v) for v in g["meta"][4:])
    blocks, split = of.make_stack("spline", dim, n_blocks, hidden=hidden, seed=seed)
    flow = stack_from(blocks, split, DEV)
    before = _lib.kernel_counts()
    with torch.no_grad():
        x, dlogp = flow(torch.from_numpy(g["z_f32"]).to(DEV))
        zi, dlogpi = flow(torch.from_numpy(g["x_f32"]).to(DEV), inverse=True)
    got = _delta(before, _lib.kernel_counts())
    assert got == {"spline_pair_wide" if mode == "pair_wide" or batch % 4 else "spline_pair": 2 * n_blocks}, got
    np.testing.assert_allclose(x.cpu().double().numpy(), g["x_f64"], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(dlogp.cpu().double().numpy(), g["dlogp_f64"], atol=1e-3, rtol=1e-4)
    np.testing.assert_allclose(zi.cpu().double().numpy(), g["zi_f64"], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(dlogpi.cpu().double().numpy(), g["dlogpi_f64"], atol=1e-3, rtol=1e-4)


def test_pair_kernel_matches_two_cta_kernel():
    """Same operand splits, same MMA order, same epilogue arithmetic: the pair kernel (both modes) and the
    two-CTAs-per-SM kernel must produce the same transformed values bit for bit, over many tiles per CTA
    (B = 2^17 + 4: 1025 tiles); the log-det is summed over four dim shares instead of two (rounding only)."""
    blocks, split = of.make_stack("spline", 66, 2, seed=0)
    flow = stack_from(blocks, split, DEV)
    g = torch.Generator().manual_seed(9)
    z = torch.rand((1 << 17) + 4, 66, generator=g).to(DEV)
    out = {}
    with torch.no_grad():
        for mode in ("tc2", "pair", "pair_wide"):
            engine.config["spline_kernel"] = mode
            out[mode] = flow(z) + flow(z, inverse=True)
    for mode in ("pair", "pair_wide"):
        for k, (a, b) in enumerate(zip(out["tc2"], out[mode])):
            if k % 2 == 0:
                assert torch.equal(a, b), (mode, k, float((a - b).abs().max()))
            else:
                torch.testing.assert_close(a, b, atol=2e-5, rtol=0)


def test_out_of_domain_inputs_are_clamped_and_counted():
    """spline.py:145-155: inputs outside [left, right] are clamped (and counted on the device)."""
    import bgflow_b200 as bg
    blocks, split = of.make_stack("spline", 384, 1, seed=1)
    flow = stack_from(blocks, split, DEV)
    z = torch.rand(200, 384, generator=torch.Generator().manual_seed(2))
    z[5, 200] = 1.5        # second half = the transformed side of block 0
    z[7, 300] = -0.25
    with torch.no_grad():
        x, d = flow(z.to(DEV))
    tr = [m for m in flow.modules() if isinstance(m, bg.ConditionalSplineTransformer)][0]
    with pytest.warns(UserWarning):
        assert tr.out_of_domain_count() == 2
    blocks64, _ = of.make_stack("spline", 384, 1, seed=1, dtype=torch.float64)
    x_ref, d_ref = of.coupling_stack(blocks64, z.clamp(0.0, 1.0).double(), split)
    np.testing.assert_allclose(x.cpu().double().numpy(), x_ref.numpy(), atol=2e-5, rtol=2e-5)
    np.testing.assert_allclose(d.cpu().double().numpy(), d_ref.numpy(), atol=1e-3, rtol=1e-5)


@pytest.mark.parametrize("k,n", [(33, 128), (128, 128), (128, 825), (825, 128), (128, 33), (1536, 128), (128, 1536), (24, 24),
                                 (7, 200)])
@pytest.mark.parametrize("batch", [1, 300, 4096 + 3])
def test_linear_layer_on_tensor_cores(k, n, batch):
    """bgx_linear (training path): y = x W^T + b with exact two-term bf16 splits on tcgen05 against fp64; K k-tiled in
    128-input groups (N <= 128) or N in 128-column passes (K <= 128); odd widths, ragged batches, odd tile counts."""
    g = torch.Generator().manual_seed(k * 1000 + n)
    x = torch.randn(batch, k, generator=g)
    w = torch.randn(n, k, generator=g) / k ** 0.5
    b = torch.randn(n, generator=g)
    ref = x.double() @ w.double().t() + b.double()
    f = engine.LinearTC()
    y = f(x.to(DEV), w.to(DEV), b.to(DEV))
    scale = ref.abs().max().item()
    np.testing.assert_allclose(y.cpu().double().numpy(), ref.numpy(), atol=3e-5 * scale, rtol=1e-4)
    y0 = f(x.to(DEV), w.to(DEV))          # no bias
    np.testing.assert_allclose(y0.cpu().double().numpy(), (ref - b.double()).numpy(), atol=3e-5 * scale, rtol=1e-4)


def test_linear_layer_rejects_shapes_it_does_not_cover():
    f = engine.LinearTC()
    with pytest.raises(_lib.BgxError):
        f(torch.zeros(8, 200, device=DEV), torch.zeros(300, 200, device=DEV))


@pytest.mark.parametrize("batch,n,k,ldg", [(1, 5, 7, 5), (70, 5, 128, 8), (1000, 128, 10, 128), (4097, 300, 33, 300),
                                           (8192 + 5, 825, 128, 828), (65536, 513, 128, 513), (200, 1100, 64, 1100)])
def test_weight_gradient_on_tensor_cores(batch, n, k, ldg):
    """bgx_gemm_tn (training path): dW = g^T h and db = sum_b g, reduced over the batch on tcgen05 with exact two-term
    bf16 splits, against fp64; ragged batches, feature counts off the 128 / 8 grid, padded row strides (pad columns
    hold garbage the kernel must not read), one to three feature-tile groups."""
    gen = torch.Generator().manual_seed(batch + 7 * n + k)
    gbuf = torch.randn(batch, ldg, generator=gen)
    h = torch.randn(batch, k, generator=gen)
    g = gbuf[:, :n]
    ref_w = g.double().t() @ h.double()
    ref_b = g.double().sum(0)
    gd = gbuf.to(DEV)
    if ldg > n:
        gd[:, n:] = float("nan")
    dw, db = engine.gemm_tn(gd, h.to(DEV), n)
    assert dw.shape == (n, k) and db.shape == (n,)
    # products carry 2^-17 relative error each (the dropped g2 h2 term), sums run in fp32
    sw = (g.double().abs().t() @ h.double().abs()).max().item()
    np.testing.assert_allclose(dw.cpu().double().numpy(), ref_w.numpy(), atol=3e-5 * sw, rtol=0)
    np.testing.assert_allclose(db.cpu().double().numpy(), ref_b.numpy(), atol=1e-5 * g.double().abs().sum(0).max().item(), rtol=0)
    dw2, db2 = engine.gemm_tn(gd, h.to(DEV), n)          # fixed summation order: reproducible
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


def test_weight_gradient_rejects_wide_inputs():
    with pytest.raises(ValueError):
        engine.gemm_tn(torch.zeros(8, 16, device=DEV), torch.zeros(8, 200, device=DEV))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process():
    """One process driving two GPUs through the mirror's device guard: the launch helpers keep their once-flags
    (dynamic shared-memory limits, cluster occupancy, SM count) per device, so the first call on cuda:1 configures
    its kernels there too.  Same weights and inputs on both devices: bit-identical rows, and gradients that agree."""
    results = []
    for dev in ("cuda:0", "cuda:1"):
        outs = []
        for kind, dim in (("spline", 66), ("affine", 66), ("spline", 384), ("affine", 384)):
            blocks, split = of.make_stack(kind, dim, 2, seed=3)
            flow = stack_from(blocks, split, dev)
            g = torch.Generator().manual_seed(dim)
            z = (torch.rand(700, dim, generator=g) if kind == "spline" else torch.randn(700, dim, generator=g)).to(dev)
            z.requires_grad_(True)
            x, d = flow(z)
            (x.sum() + d.sum()).backward()
            outs += [x.detach().cpu(), d.detach().cpu(), z.grad.cpu()]
        w = torch.randn(300, 200, generator=torch.Generator().manual_seed(1))
        h = torch.randn(300, 64, generator=torch.Generator().manual_seed(2))
        dw, db = engine.gemm_tn(w.to(dev), h.to(dev))
        outs += [dw.cpu(), db.cpu()]
        engine.check_pipeline_status(dev)
        results.append(outs)
    for a, b in zip(*results):
        assert torch.equal(a, b)
