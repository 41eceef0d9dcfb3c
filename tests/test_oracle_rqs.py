"""Pins for the restated nflows rational-quadratic spline (oracle/flows.py).

nflows is a third-party, unpinned dependency of the reference and absent here, so the
restatement is anchored on (1) an independent on-disk port of the same nflows function
(transformers' VITS), (2) fp64 autograd derivatives, (3) round trips, and (4) the
reference's own property tests (tests/nn/flow/transformer/test_spline.py:8-69,
tests/factory/test_generator_builder.py:131-136)."""

import math

import pytest
import torch

from oracle import flows as of


def _rand_params(shape, k, seed, dtype=torch.float64, scale=2.0):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(*shape, k, generator=g, dtype=dtype) * scale
    h = torch.randn(*shape, k, generator=g, dtype=dtype) * scale
    s = torch.randn(*shape, k + 1, generator=g, dtype=dtype) * scale
    x = torch.rand(*shape, generator=g, dtype=dtype)
    return x, w, h, s


@pytest.mark.parametrize("reverse", [False, True])
def test_matches_independent_vits_port(reverse):
    vits = pytest.importorskip("transformers.models.vits.modeling_vits")
    tb = 3.0
    x, w, h, s = _rand_params((64, 9), 8, 0)
    x = (x * 2 - 1) * tb * 0.999
    ref_y, ref_lad = vits._rational_quadratic_spline(
        x, w, h, s, reverse=reverse, tail_bound=tb, min_bin_width=1e-3, min_bin_height=1e-3,
        min_derivative=1e-3)
    y, lad = of.rational_quadratic_spline(
        x, w, h, s, inverse=reverse, left=-tb, right=tb, bottom=-tb, top=tb,
        enable_identity_init=False)
    torch.testing.assert_close(y, ref_y, atol=1e-12, rtol=1e-12)
    torch.testing.assert_close(lad, ref_lad, atol=1e-11, rtol=1e-11)


@pytest.mark.parametrize("identity_init", [True, False])
@pytest.mark.parametrize("inverse", [False, True])
def test_logabsdet_is_the_derivative(inverse, identity_init):
    x, w, h, s = _rand_params((200,), 8, 1)
    x = x.clone().requires_grad_(True)
    y, lad = of.rational_quadratic_spline(x, w, h, s, inverse=inverse,
                                          enable_identity_init=identity_init)
    (dy,) = torch.autograd.grad(y.sum(), x)
    torch.testing.assert_close(torch.log(dy), lad, atol=1e-9, rtol=1e-9)
    assert (dy > 0).all()


def test_round_trip_and_range():
    x, w, h, s = _rand_params((500, 5), 8, 2)
    y, lad = of.rational_quadratic_spline(x, w, h, s, inverse=False)
    assert (y > 0).all() and (y < 1).all()           # test_spline.py:30-33
    xb, ladb = of.rational_quadratic_spline(y, w, h, s, inverse=True)
    torch.testing.assert_close(xb, x, atol=1e-12, rtol=1e-10)
    torch.testing.assert_close(lad + ladb, torch.zeros_like(lad), atol=1e-10, rtol=0)


def test_zero_parameters_is_identity():
    # enable_identity_init=True: softplus(0, beta) + min_d == 1 (nflows PR #65), which
    # tests/factory/test_generator_builder.py:131-136 relies on (atol=0.01)
    x = torch.rand(100, 4, dtype=torch.float64)
    z = torch.zeros(100, 4, 8, dtype=torch.float64)
    zs = torch.zeros(100, 4, 9, dtype=torch.float64)
    for inv in (False, True):
        y, lad = of.rational_quadratic_spline(x, z, z, zs, inverse=inv, enable_identity_init=True)
        torch.testing.assert_close(y, x, atol=1e-12, rtol=0)
        torch.testing.assert_close(lad, torch.zeros_like(lad), atol=1e-12, rtol=0)
    beta = math.log(2) / (1 - 1e-3)
    assert abs(1e-3 + math.log1p(1.0) / beta - 1.0) < 1e-15


def test_slope_periodicity_of_params():
    # tests/nn/flow/transformer/test_spline.py:36-69
    p = torch.randn(10, 3 * 8 * 6 + 3)
    circ = [True, False, True, False, True, False]
    w, h, s = of.spline_params(p, 6, circ)
    assert w.shape == (10, 6, 8) and s.shape == (10, 6, 9)
    assert torch.equal(s[:, [0, 2, 4], 0], s[:, [0, 2, 4], -1])
    assert not torch.equal(s[:, [1, 3, 5], 0], s[:, [1, 3, 5], -1])
    w, h, s = of.spline_params(torch.randn(10, 3 * 8 * 6), 6, True)
    assert torch.equal(s[..., 0], s[..., -1])
    with pytest.raises(RuntimeError):
        of.spline_params(torch.randn(10, 3 * 8 * 6 + 1), 6, False)


def test_boundaries_and_clamp():
    x, w, h, s = _rand_params((50,), 8, 3, dtype=torch.float32)
    x[:5] = 0.0
    x[5:10] = 1.0
    for inv in (False, True):
        y, lad = of.rational_quadratic_spline(x, w, h, s, inverse=inv)
        assert torch.isfinite(y).all() and torch.isfinite(lad).all()
        torch.testing.assert_close(y[:5], torch.zeros(5), atol=1e-6, rtol=0)
        torch.testing.assert_close(y[5:10], torch.ones(5), atol=1e-6, rtol=0)
