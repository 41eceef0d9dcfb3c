"""Self-consistency pins of the CDF-map oracle (oracle/cdf.py) beyond the reference goldens:
cdf(icdf(u)) = u, the log-det is the log of the map's own derivative (autograd in fp64), the
derived truncation constants against scipy, and the eps-clamp semantics of cdf.py:29-46."""

import math

import numpy as np
import pytest
import torch

from oracle import cdf as ocdf

T = lambda *v: torch.tensor(v, dtype=torch.float64)


def dists():
    return {
        "trunc": ocdf.TruncatedNormal(T(1.0, 0.5, -0.3), T(1.0, 0.4, 2.0), T(1e-5, 1e-5, -1.0), T(math.inf, 1.0, 0.5)),
        "normal": ocdf.Normal(T(0.0, -0.7, 3.0), T(20.0, 1.0, 0.1)),
        "uniform": ocdf.Uniform(T(0.0, -1.0, 2.0), T(1.0, 3.0, 2.5)),
    }


@pytest.mark.parametrize("name", ["trunc", "normal", "uniform"])
def test_round_trip_and_logdet_is_log_derivative(name):
    d = dists()[name]
    g = torch.Generator().manual_seed(0)
    u = (torch.rand(200, 3, generator=g, dtype=torch.float64) * 0.98 + 0.01).requires_grad_(True)
    x, dl = ocdf.cdf_transform(d, u, inverse=True)
    ub, dlb = ocdf.cdf_transform(d, x.detach())
    torch.testing.assert_close(ub, u.detach(), atol=1e-12, rtol=0)
    torch.testing.assert_close(dl + dlb, torch.zeros_like(dl), atol=1e-10, rtol=0)
    # elementwise map: d x_j / d u_j from one backward pass of sum(x)
    (jac,) = torch.autograd.grad(x.sum(), u)
    torch.testing.assert_close(torch.log(jac).sum(-1, keepdim=True), dl.detach(), atol=1e-9, rtol=0)


def test_truncation_constants_against_scipy():
    from scipy.stats import truncnorm
    mu, sigma, lo, hi = 0.5, 1.0, 1e-5, 1.0
    d = ocdf.TruncatedNormal(T(mu), T(sigma), T(lo), T(hi))
    ref = truncnorm((lo - mu) / sigma, (hi - mu) / sigma, loc=mu, scale=sigma)
    u = np.linspace(0.01, 0.99, 50)
    np.testing.assert_allclose(d.icdf(torch.from_numpy(u)[:, None])[:, 0].numpy(), ref.ppf(u), atol=1e-10)
    x = ref.ppf(u)
    np.testing.assert_allclose(d.cdf(torch.from_numpy(x)[:, None])[:, 0].numpy(), u, atol=1e-10)
    np.testing.assert_allclose(d.log_prob(torch.from_numpy(x)[:, None])[:, 0].numpy(), ref.logpdf(x), atol=1e-10)


def test_eps_clamps():
    d = dists()["uniform"]
    # outside the support: log_prob = -inf, clamped to -1/eps; cdf clamped into [eps, 1 - eps]
    x = T([-5.0, -1.0, 2.6], [0.5, 0.0, 2.25])
    u, dl = ocdf.cdf_transform(d, x)
    assert float(u[0, 0]) == 1e-7 and float(u[0, 2]) == 1 - 1e-7
    assert float(dl[0]) == pytest.approx(-2e7 - math.log(4.0), rel=1e-12)      # two columns outside + log(1/4)
    assert float(dl[1]) == pytest.approx(-math.log(1.0) - math.log(4.0) - math.log(0.5))
    # eps = None: no clamps at all
    u2, dl2 = ocdf.cdf_transform(d, x, eps=None)
    assert float(u2[0, 0]) == 0.0 and float(dl2[0]) == -math.inf
    # icdf clamps its argument first
    n = dists()["normal"]
    y, _ = ocdf.cdf_transform(n, T([0.0, 1.0, 0.5]), inverse=True)
    assert torch.isfinite(y).all()
