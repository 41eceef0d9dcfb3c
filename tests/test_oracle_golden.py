"""Pin the CPU oracle to outputs of the reference itself (fixtures made by
tests/golden/make_golden.py from the unmodified reference)."""

import numpy as np
import pytest
import torch

from oracle import flows as of
from oracle import ic as oic
from conftest import load_golden

DT = {"f32": torch.float32, "f64": torch.float64}
# the oracle restates the reference op for op: fp64 agrees to rounding, fp32 to a few ulp
TOL = {"f32": dict(atol=2e-5, rtol=2e-5), "f64": dict(atol=1e-11, rtol=1e-11)}


def _close(a, b, tag, scale=1.0):
    t = TOL[tag]
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), atol=t["atol"] * scale, rtol=t["rtol"] * scale)


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name,kind,data", [
    ("affine_d66_8blk", "affine", "normal"), ("spline_d66_8blk", "spline", "uniform"),
    ("affine_d10_3blk", "affine", "normal"), ("spline_d7_4blk", "spline", "uniform")])
def test_stack_matches_reference(name, kind, data, tag):
    g = load_golden(name)
    meta = g["meta"]
    dim, n_blocks, batch, seed = (int(v) for v in meta[:4])
    hidden = tuple(int(v) for v in meta[4:])
    blocks, split = of.make_stack(kind, dim, n_blocks, hidden=hidden, seed=seed, dtype=DT[tag])
    if tag == "f32":
        tot = sum(float(t.double().abs().sum()) for b in blocks for k in ("shift", "scale", "params_net")
                  if b.get(k) is not None for t in b[k].weights + b[k].biases)
        assert abs(tot - float(g["param_checksum"])) < 1e-6 * tot, "seeded parameters drifted"
    z = torch.from_numpy(g[f"z_{tag}"])
    x, dlogp = of.coupling_stack(blocks, z, split)
    _close(x, g[f"x_{tag}"], tag, 4)
    _close(dlogp, g[f"dlogp_{tag}"], tag, 20)
    zi, dlogpi = of.coupling_stack(blocks, torch.from_numpy(g[f"x_{tag}"]), split, inverse=True)
    _close(zi, g[f"zi_{tag}"], tag, 4)
    _close(dlogpi, g[f"dlogpi_{tag}"], tag, 20)
    if kind == "affine":
        x2, d2 = of.coupling_stack(blocks, z, split, inverse=True)
        _close(x2, g[f"inv_x_{tag}"], tag, 4)
        _close(d2, g[f"inv_dlogp_{tag}"], tag, 20)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_readme_config(tag):
    g = load_golden("readme_doublewell")
    gen = torch.Generator().manual_seed(0)
    shift = of.make_mlp([1, 4, 1], "relu", gen, DT[tag])
    scale = of.make_mlp([1, 4, 1], "tanh", gen, DT[tag])
    blk = {"kind": "affine", "shift": shift, "scale": scale, "log_alpha": -1.0}
    z = torch.from_numpy(g[f"z_{tag}"])
    xs, dlogp = of.coupling_block(blk, [z[:, :1], z[:, 1:]])
    x = torch.cat(xs, -1)
    _close(x, g[f"x_{tag}"], tag)
    _close(dlogp, g[f"dlogp_{tag}"], tag)
    # NLL = prior.energy(z) - dlogp_inv (bg.py:20-22), prior = standard normal in 2-D
    zs, dinv = of.coupling_block(blk, [x[:, :1], x[:, 1:]], inverse=True)
    zz = torch.cat(zs, -1)
    nll = 0.5 * zz.pow(2).sum(-1, keepdim=True) + np.log(2 * np.pi) - dinv
    _close(nll, g[f"nll_{tag}"], tag, 4)


def _multi_blocks(dtype):
    g = torch.Generator().manual_seed(7)
    nb = 6
    blk_a = {"kind": "spline", "transformed": (2,), "cond": (0, 3), "is_circular": True,
             "params_net": of.make_mlp([11, 32, 32, 5 * 3 * nb], "silu", g, dtype)}
    net_b = of.make_mlp([2 * 5 + 4, 48, 13 * (3 * nb + 1)], "tanh", g, dtype)
    net_b.periodic = (list(range(5)), 0.0, 1.0)
    blk_b = {"kind": "spline", "transformed": (0, 1), "cond": (2, 3), "params_net": net_b}
    blk_c = {"kind": "affine", "transformed": (3,), "cond": (1,), "is_circular": True,
             "shift": of.make_mlp([6, 16, 4], "relu", g, dtype), "scale": None}
    blk_d = {"kind": "affine", "transformed": (1, 3), "cond": (0,), "preserve_volume": True,
             "shift": of.make_mlp([7, 24, 24, 10], "silu", g, dtype),
             "scale": of.make_mlp([7, 24, 24, 10], "silu", g, dtype), "log_alpha": -0.5}
    return [blk_a, blk_b, blk_c, blk_d]


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_multi_tensor_coupling(tag):
    g = load_golden("multi_tensor_coupling")
    blocks = _multi_blocks(DT[tag])
    xs = [torch.from_numpy(g[f"in{i}_{tag}"]) for i in range(4)]
    dlogp = 0
    for b in blocks:
        xs, d = of.coupling_block(b, xs)
        dlogp = dlogp + d
    for i in range(4):
        _close(xs[i], g[f"out{i}_{tag}"], tag, 4)
    _close(dlogp, g[f"dlogp_{tag}"], tag, 20)
    dlogpi = 0
    for b in reversed(blocks):
        xs, d = of.coupling_block(b, xs, inverse=True)
        dlogpi = dlogpi + d
    for i in range(4):
        _close(xs[i], g[f"back{i}_{tag}"], tag, 8)
    _close(dlogpi, g[f"dlogpi_{tag}"], tag, 20)


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", ["ic_ala2", "ic_ala2_raw", "ic_chain12"])
def test_ic_matches_reference(name, tag):
    g = load_golden(name)
    plan = oic.make_plan(g["z_matrix"])
    norm = bool(int(g["normalize"]))
    t = lambda k: torch.from_numpy(g[f"{k}_{tag}"])
    bonds, angles, torsions, x0, R, dlogp = oic.xyz_to_ic(plan, t("xyz"), normalize_angles=norm)
    # the fp32 reference gets its reference-frame log-det from a noisy 9x9 autograd
    # Jacobian (ic_helper.py:655-678, ~2e-5 noise) -> looser dlogp in fp32
    dl_scale = 50 if tag == "f32" else 100
    _close(bonds, t("bonds"), tag, 4)
    _close(angles, t("angles"), tag, 4)
    _close(torsions, t("torsions"), tag, 4)
    _close(x0, t("x0"), tag)
    _close(R, t("R"), tag, 4)
    _close(dlogp, t("dlogp"), tag, dl_scale)
    xyz, dinv = oic.ic_to_xyz(plan, t("bonds"), t("angles"), t("torsions"), t("x0"), t("R"),
                              normalize_angles=norm)
    _close(xyz, t("xyz_back"), tag, 10)
    _close(dinv, t("dlogp_inv"), tag, dl_scale)
    xyz, dgen = oic.ic_to_xyz(plan, t("gen_bonds"), t("gen_angles"), t("gen_torsions"), t("gen_x0"),
                              t("gen_R"), normalize_angles=norm)
    _close(xyz, t("gen_xyz"), tag, 10)
    _close(dgen, t("gen_dlogp"), tag, dl_scale)


def test_ala2_plan_matches_survey():
    plan = oic.make_plan(oic.ALA2_GLOBAL_Z)
    assert plan.seeds == [0, 1, 2]
    assert len(plan.rel) == 19 and sorted(plan.order) == list(range(19))


# ------------------------------------------------------------------ SURVEY 8f rows
from oracle import cdf as ocdf

CDF_FIELDS = {"bonds": 21, "angles": 20, "torsions": 19, "fixed": 9, "augmented": 10}


def _finite_close(a, b, tag, scale=1.0):
    a, b = np.asarray(a), np.asarray(b)
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    m = np.isfinite(b)
    _close(a[m], b[m], tag, scale)
    assert np.array_equal(a[~m], b[~m])          # same +-inf


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_cdf_maps_match_reference(tag):
    """CDFTransform over the builder's IC marginals, both directions, incl. the eps-clamped edges."""
    g = load_golden("cdf_maps")
    marg = ocdf.ic_marginals(CDF_FIELDS, DT[tag])
    t = lambda k: torch.from_numpy(g[f"{k}_{tag}"])
    custom = ocdf.TruncatedNormal(t("custom_mu"), t("custom_sigma"), t("custom_lower"), t("custom_upper"))
    zero = torch.zeros(5, dtype=DT[tag])
    halfopen = ocdf.TruncatedNormal(zero, torch.tensor(1.0), -torch.tensor(np.inf, dtype=DT[tag]),
                                    torch.tensor(np.inf, dtype=DT[tag]))
    cases = dict(marg, custom=custom, halfopen=halfopen)
    for name, dist in cases.items():
        x, d = ocdf.cdf_transform(dist, t(f"{name}_u"), inverse=True)
        _finite_close(x, t(f"{name}_x"), tag)
        _finite_close(d, t(f"{name}_dlogp"), tag, 4)
        u, db = ocdf.cdf_transform(dist, t(f"{name}_x"), inverse=False)
        _finite_close(u, t(f"{name}_ub"), tag)
        _finite_close(db, t(f"{name}_dlogpb"), tag, 4)


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", ["relic_ala2", "mixed_ala2_keep9", "mixed_ala2_keep15"])
def test_relative_ic_matches_reference(name, tag):
    g = load_golden(name)
    plan = oic.make_rel_plan(g["z_matrix"], g["fixed"])
    keep = int(g["keepdims"])
    t = lambda k: torch.from_numpy(g[f"{k}_{tag}"])
    if keep < 0:
        fwd = lambda x: oic.rel_xyz_to_ic(plan, x)
        inv = lambda *a: oic.rel_ic_to_xyz(plan, *a)
    else:
        data = torch.from_numpy(g["pca_data"]).to(DT[tag]).reshape(-1, 22, 3)[:, g["fixed"]].reshape(-1, 15)
        white = oic.Whitening(data.numpy(), keepdims=keep)
        fwd = lambda x: oic.mixed_xyz_to_ic(plan, white, x)
        inv = lambda *a: oic.mixed_ic_to_xyz(plan, white, *a)
    # fp32 PCA (numpy eigh in float32, pca.py:20-33) is only reproducible to ~1e-4 relative
    fs = 1
    bonds, angles, torsions, fixed, dlogp = fwd(t("xyz"))
    _close(bonds, t("bonds"), tag, 4)
    _close(angles, t("angles"), tag, 4)
    _close(torsions, t("torsions"), tag, 4)
    _close(fixed, t("fixed"), tag, 4 * fs)
    _close(dlogp, t("dlogp"), tag, 40 * fs)
    xyz, dinv = inv(t("bonds"), t("angles"), t("torsions"), t("fixed"))
    _close(xyz, t("xyz_back"), tag, 10 * fs)
    _close(dinv, t("dlogp_inv"), tag, 40 * fs)
    xyz, dgen = inv(t("gen_bonds"), t("gen_angles"), t("gen_torsions"), t("gen_fixed"))
    _close(xyz, t("gen_xyz"), tag, 10 * fs)
    _close(dgen, t("gen_dlogp"), tag, 40 * fs)
