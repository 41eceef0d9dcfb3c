"""CPU checks of the kernels' ``__host__ __device__`` arithmetic (bgx_cdf_math.cuh compiled with g++
by this test: tests/native/hostmath.cpp) against scipy and the oracle.  The product never runs this
host build; it exists so that the arithmetic is verified before GPU time is spent."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import cdf as ocdf
from conftest import ROOT, load_golden

SRC = os.path.join(ROOT, "tests", "native", "hostmath.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "native", "_build")
KIND = {"normal": 1, "trunc": 2, "uniform": 3}


@pytest.fixture(scope="module")
def hm():
    os.makedirs(OUT_DIR, exist_ok=True)
    so = os.path.join(OUT_DIR, "hostmath.so")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", so,
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "bgflow_b200", "csrc")]
    subprocess.run(cmd, check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.hm_std_normal_icdf.restype = C.c_float
    lib.hm_std_normal_icdf.argtypes = [C.c_float]
    lib.hm_std_normal_cdf.restype = C.c_float
    lib.hm_std_normal_cdf.argtypes = [C.c_float]
    lib.hm_cdf_apply.restype = C.c_int
    lib.hm_cdf_apply.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_float,
                                 C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


EPS_LO, EPS_HI, LD_MIN = float(np.float32(1e-7)), float(np.float32(1.0 - 1e-7)), -1e7


def apply(lib, kind, a, b, lower, upper, x, inverse):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out, ld = np.empty_like(x), np.empty_like(x)
    rc = lib.hm_cdf_apply(KIND[kind], a, b, lower, upper, int(inverse), EPS_LO, EPS_HI, LD_MIN, x.size,
                          x.ctypes.data, out.ctypes.data, ld.ctypes.data)
    assert rc == 0
    return out, ld


def test_std_normal_icdf_against_scipy(hm):
    from scipy.special import ndtri, ndtr
    p = np.concatenate([np.linspace(1e-7, 1 - 1e-7, 20001), 10.0 ** -np.linspace(2, 30, 200)]).astype(np.float32)
    z = np.array([hm.hm_std_normal_icdf(float(v)) for v in p])
    ref = ndtri(p.astype(np.float64))
    assert np.max(np.abs(z - ref) / np.maximum(1.0, np.abs(ref))) < 6e-7
    zz = np.linspace(-12, 8, 4001).astype(np.float32)
    c = np.array([hm.hm_std_normal_cdf(float(v)) for v in zz])
    ref_c = ndtr(zz.astype(np.float64))
    # erfc amplifies the fp32 rounding of its argument by ~z^2 in the lower tail
    assert np.all(np.abs(c - ref_c) <= 3e-7 * np.maximum(1.0, zz.astype(np.float64) ** 2) * ref_c + 1e-38)


@pytest.mark.parametrize("name,kind,a,b,lower,upper", [
    ("bonds", "trunc", 1.0, 1.0, 1e-5, np.inf), ("angles", "trunc", 0.5, 1.0, 1e-5, 1.0),
    ("torsions", "uniform", 0.0, 1.0, 0, 0), ("fixed", "normal", 0.0, 20.0, 0, 0),
    ("augmented", "normal", 0.0, 1.0, 0, 0), ("halfopen", "trunc", 0.0, 1.0, -np.inf, np.inf)])
def test_cdf_math_matches_reference_golden(hm, name, kind, a, b, lower, upper):
    """The kernel arithmetic against the REFERENCE's fp64 outputs (tests/golden/cdf_maps.npz)."""
    g = load_golden("cdf_maps")
    u = g[f"{name}_u_f32"]
    x, ld = apply(hm, kind, a, b, lower, upper, u, inverse=True)
    x64 = g[f"{name}_x_f64"]
    # the fp64 reference saw the same fp32-representable inputs only approximately (u was drawn in
    # fp64): compare on the bulk, where d icdf / du is moderate
    bulk = (u > 1e-3) & (u < 1 - 1e-3)
    np.testing.assert_allclose(x[bulk], x64[bulk], rtol=2e-5, atol=2e-5)
    ld_rows = ld.reshape(u.shape).sum(-1)
    rows = bulk.all(-1)
    np.testing.assert_allclose(ld_rows[rows], g[f"{name}_dlogp_f64"][rows, 0], rtol=1e-4, atol=1e-4)
    # forward direction on the reference's fp32 x
    xb = g[f"{name}_x_f32"]
    ub, ldb = apply(hm, kind, a, b, lower, upper, xb, inverse=False)
    ref_u = ocdf.cdf_transform(_dist(kind, a, b, lower, upper, u.shape[1]), torch.from_numpy(xb).double())[0].numpy()
    np.testing.assert_allclose(ub[bulk], ref_u[bulk], rtol=1e-5, atol=2e-7)


def _dist(kind, a, b, lower, upper, n, dtype=torch.float64):
    t = lambda v: torch.full((n,), float(v), dtype=dtype)
    if kind == "trunc":
        return ocdf.TruncatedNormal(t(a), t(b), t(lower), t(upper))
    if kind == "normal":
        return ocdf.Normal(t(a), t(b))
    return ocdf.Uniform(t(a), t(b))


@pytest.mark.parametrize("kind,a,b,lower,upper", [
    ("trunc", 1.3, 0.4, 0.2, 4.0), ("trunc", 0.5, 1.0, 1e-5, 1.0), ("normal", -0.7, 2.5, 0, 0),
    ("uniform", -1.0, 3.0, 0, 0)])
def test_cdf_math_against_fp64_oracle(hm, kind, a, b, lower, upper):
    """Same fp32 inputs into the kernel arithmetic and into the oracle in fp64: both directions,
    round trip, log-det consistency, eps clamps at the edges."""
    rng = np.random.default_rng(3)
    u = rng.random(4096).astype(np.float32)
    u[:6] = [0.0, 1.0, 1e-9, 1 - 1e-9, 1e-5, 1 - 1e-5]
    d = _dist(kind, a, b, lower, upper, 1)
    x, ld = apply(hm, kind, a, b, lower, upper, u, inverse=True)
    xr, ldr = ocdf.cdf_transform(d, torch.from_numpy(u).double()[:, None], inverse=True)
    xr, ldr = xr[:, 0].numpy(), ldr[:, 0].numpy()
    # sensitivity of icdf to the fp32 rounding of Z*u + cdf_lo grows like 1/pdf in the tails
    sens = np.exp(ldr)                                    # = 1 / pdf(x) = |dx/du|
    tol = 3e-7 * sens * 2 + 2e-6 * (1 + np.abs(xr))
    assert np.all(np.abs(x - xr) <= tol)
    # d(ld)/du = z / sigma * dx/du
    ld_tol = 1e-5 + 1e-6 * sens * np.abs(xr - a) / b ** 2 + 1e-6 * np.abs(ldr)
    assert np.all(np.abs(ld - ldr) <= ld_tol), np.max(np.abs(ld - ldr) / ld_tol)
    # forward on the oracle's x (rounded to fp32)
    xf = xr.astype(np.float32)
    uu, ldf = apply(hm, kind, a, b, lower, upper, xf, inverse=False)
    ur, ldfr = ocdf.cdf_transform(d, torch.from_numpy(xf).double()[:, None])
    np.testing.assert_allclose(uu, ur[:, 0].numpy(), rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(ldf, ldfr[:, 0].numpy(), rtol=1e-6, atol=2e-6)


def test_cdf_edge_semantics(hm):
    # outside the support of a uniform: log_prob = -inf -> clamped to -1/eps (cdf.py:34-35)
    u, ld = apply(hm, "uniform", 0.0, 1.0, 0, 0, np.array([-0.5, 1.5, 1.0, 0.0, 0.25], np.float32), inverse=False)
    assert list(ld[:3]) == [LD_MIN, LD_MIN, LD_MIN] and ld[3] == 0.0 and ld[4] == 0.0
    np.testing.assert_array_equal(u, np.array([EPS_LO, EPS_HI, EPS_HI, EPS_LO, 0.25], np.float32))
    # icdf clamps its argument to [eps, 1 - eps]: finite everywhere
    x, ld = apply(hm, "trunc", 1.0, 1.0, 1e-5, np.inf, np.array([0.0, 1.0], np.float32), inverse=True)
    assert np.all(np.isfinite(x)) and np.all(np.isfinite(ld)) and x[0] >= 1e-5


# ------------------------------------------------------------------ register spline evaluation

def _rqs(hm_lib, root, fast, params, x, dom=(0.0, 1.0, 0.0, 1.0), identity_init=True):
    hm_lib.hm_rqs_eval.restype = C.c_int
    hm_lib.hm_rqs_eval.argtypes = [C.c_int, C.c_int] + [C.c_float] * 7 + [C.c_int, C.c_int] + [C.c_void_p] * 4
    params = np.ascontiguousarray(params, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    y, lad = np.empty_like(x), np.empty_like(x)
    rc = hm_lib.hm_rqs_eval(int(root), int(fast), *dom, 1e-3, 1e-3, 1e-3, int(identity_init), x.size,
                            params.ctypes.data, x.ctypes.data, y.ctypes.data, lad.ctypes.data)
    assert rc == 0
    return y, lad


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("root", [True, False])
@pytest.mark.parametrize("dom", [(0.0, 1.0, 0.0, 1.0), (-2.0, 3.0, -1.0, 0.5)])
def test_register_spline_evaluation_matches_oracle(hm, root, fast, dom):
    """``rqs_eval_reg`` (the tensor-core kernels' per-dim spline arithmetic: softmax knots from prefix
    sums, binary / linear bin search, root / direct branch, log-det) compiled for the host, against the
    fp64 oracle restatement of nflows' rational_quadratic_spline on the same fp32 parameters."""
    from oracle import flows as of
    rng = np.random.default_rng(7)
    n = 20000
    params = (rng.standard_normal((n, 25)) * rng.choice([0.3, 1.0, 3.0], size=(n, 1))).astype(np.float32)
    left, right, bottom, top = dom
    lo, hi = (bottom, top) if root else (left, right)
    x = (lo + (hi - lo) * rng.random(n)).astype(np.float32)
    x[:4] = [lo, hi, lo + 1e-7 * (hi - lo), hi - 1e-6 * (hi - lo)]
    y, lad = _rqs(hm, root, fast, params, x, dom)
    p64 = torch.from_numpy(params).double()
    yr, lr = of.rational_quadratic_spline(torch.from_numpy(x).double().clamp(lo, hi), p64[:, :8], p64[:, 8:16], p64[:, 16:],
                                          inverse=root, left=left, right=right, bottom=bottom, top=top)
    yr, lr = yr.numpy(), lr.numpy()
    span = (right - left) if root else (top - bottom)
    # a knot that moves by an fp32 rounding can flip the bin of an input sitting on it: y stays
    # continuous across the flip, the log-det does not -> compare the log-det where the bins agree
    # (parameters with std up to 3 give splines with slopes of 1e-3 .. 1e3: the few worst-conditioned
    # evaluations lose fp32 digits in the quadratic root; bound the tail and the bulk separately)
    err = np.abs(y - yr) / max(1.0, span)
    assert err.max() < 1e-4 and np.quantile(err, 0.999) < 5e-6, (err.max(), np.quantile(err, 0.999))
    close = np.abs(lad - lr) < 2e-4
    assert close.mean() > 0.995, close.mean()
    assert np.median(np.abs(lad - lr)) < 2e-6
    # outputs stay inside the image interval (tests/nn/flow/transformer/test_spline.py:30-33)
    olo, ohi = (left, right) if root else (bottom, top)
    assert (y >= olo - 1e-6).all() and (y <= ohi + 1e-6).all()


def test_register_spline_fast_and_reference_paths_agree(hm):
    """Binary bin search + MUFU-style log-det vs linear walk + logf: same bins, same values."""
    rng = np.random.default_rng(8)
    params = rng.standard_normal((5000, 25)).astype(np.float32) * 2
    x = rng.random(5000).astype(np.float32)
    for root in (True, False):
        y1, l1 = _rqs(hm, root, True, params, x)
        y0, l0 = _rqs(hm, root, False, params, x)
        np.testing.assert_array_equal(y1, y0)
        np.testing.assert_allclose(l1, l0, atol=2e-6, rtol=0)
    # zero parameters = identity (tests/factory/test_generator_builder.py:131-136, identity init)
    y, lad = _rqs(hm, True, True, np.zeros((100, 25), np.float32), np.linspace(0.01, 0.99, 100))
    np.testing.assert_allclose(y, np.linspace(0.01, 0.99, 100), atol=2e-6)
    np.testing.assert_allclose(lad, 0.0, atol=2e-6)


@pytest.mark.parametrize("root", [True, False])
@pytest.mark.parametrize("dom", [(0.0, 1.0, 0.0, 1.0), (-2.0, 3.0, -1.0, 5.0)])
def test_packed_two_dim_evaluation_equals_scalar_evaluation(hm, root, dom):
    """``rqs_eval_reg2`` (two transformed dims per thread in packed fp32 pairs, bgx_spline_reg2.cuh) performs
    the scalar evaluation's operations in the same order for each lane: on the host build both agree to the last
    bit, on every input including the domain edges."""
    rng = np.random.default_rng(11)
    n = 20000
    params = (rng.standard_normal((n, 25)) * rng.choice([0.3, 1.0, 3.0], size=(n, 1))).astype(np.float32)
    left, right, bottom, top = dom
    lo, hi = (bottom, top) if root else (left, right)
    x = (lo + (hi - lo) * rng.random(n)).astype(np.float32)
    x[:4] = [lo, hi, lo + 1e-7 * (hi - lo), hi - 1e-6 * (hi - lo)]
    y1, l1 = _rqs(hm, root, True, params, x, dom)
    hm.hm_rqs_eval2.restype = C.c_int
    hm.hm_rqs_eval2.argtypes = [C.c_int] + [C.c_float] * 7 + [C.c_int, C.c_int] + [C.c_void_p] * 4
    y2, l2 = np.empty_like(x), np.empty_like(x)
    rc = hm.hm_rqs_eval2(int(root), *dom, 1e-3, 1e-3, 1e-3, 1, n, params.ctypes.data, x.ctypes.data, y2.ctypes.data,
                         l2.ctypes.data)
    assert rc == 0
    np.testing.assert_array_equal(y1, y2)
    np.testing.assert_array_equal(l1, l2)
