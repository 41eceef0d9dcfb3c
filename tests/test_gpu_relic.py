"""Parity of the relative / mixed internal-coordinate kernels (``bgx_relic_to_xyz`` /
``bgx_relic_from_xyz``) against reference-generated golden fixtures and the oracle.
Tolerances as for the global transform (tests/nn/flow/crd_transform/test_ic.py:28-31):
coordinates / ICs 1e-4, dlogp 1e-3; whitened coordinates 1e-3 (they are divided by std ~ 0.02)."""

import numpy as np
import pytest
import torch

import bgflow_b200 as bg
from bgflow_b200 import _lib
from oracle import ic as oic
from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def _cmp(a, b, atol, what=""):
    np.testing.assert_allclose(a.detach().cpu().double().numpy(), np.asarray(b, dtype=np.float64),
                               atol=atol, rtol=1e-4, err_msg=what)


def _layer(g):
    keep = int(g["keepdims"])
    if keep < 0:
        return bg.RelativeInternalCoordinateTransformation(g["z_matrix"], g["fixed"])
    data = torch.from_numpy(g["pca_data"]).float()
    return bg.MixedCoordinateTransformation(data, g["z_matrix"], g["fixed"], keepdims=keep)


@pytest.mark.parametrize("name", ["relic_ala2", "mixed_ala2_keep9", "mixed_ala2_keep15"])
def test_matches_reference_golden(name):
    g = load_golden(name)
    ic = _layer(g)
    fix_tol = 1e-4 if int(g["keepdims"]) < 0 else 1e-3
    n0 = _lib.launch_count()
    bonds, angles, torsions, fixed, dlogp = ic(_t(g["xyz_f32"]))
    assert _lib.launch_count() <= n0 + 2          # full tiles through bulk-TMA copies + (possibly) one launch for the tail rows
    B = g["xyz_f32"].shape[0]
    assert bonds.shape == (B, 17) and fixed.shape == (B, ic.dim_fixed) and dlogp.shape == (B, 1)
    for got, key in ((bonds, "bonds"), (angles, "angles"), (torsions, "torsions")):
        _cmp(got, g[key + "_f64"], 1e-4, key)
    _cmp(fixed, g["fixed_f64"], fix_tol, "fixed")
    _cmp(dlogp, g["dlogp_f64"], 1e-3, "dlogp")
    xyz, dinv = ic(*(_t(g[k + "_f32"]) for k in ("bonds", "angles", "torsions", "fixed")), inverse=True)
    _cmp(xyz, g["xyz_back_f64"], 1e-4, "xyz_back")
    _cmp(dinv, g["dlogp_inv_f64"], 1e-3, "dlogp_inv")
    xyz, dgen = ic(*(_t(g["gen_" + k + "_f32"]) for k in ("bonds", "angles", "torsions", "fixed")), inverse=True)
    _cmp(xyz, g["gen_xyz_f64"], 1e-4, "gen_xyz")
    _cmp(dgen, g["gen_dlogp_f64"], 1e-3, "gen_dlogp")


@pytest.mark.parametrize("keep", [None, 15])
@pytest.mark.parametrize("batch", [1, 31, 32, 33, 127, 129, 5000])
def test_round_trip_ragged(batch, keep):
    """tests/nn/flow/crd_transform/test_ic.py:529-550 (inversion) on ragged batches; keepdims = 15 keeps
    the whitening invertible."""
    g = torch.Generator().manual_seed(batch)
    x0 = torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float32).reshape(1, -1)
    xyz = (x0 + 0.01 * torch.randn(batch, 66, generator=g)).to(DEV)
    if keep is None:
        ic = bg.RelativeInternalCoordinateTransformation(oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK)
    else:
        ic = bg.MixedCoordinateTransformation(x0 + 0.02 * torch.randn(400, 66, generator=g), oic.ALA2_RELATIVE_Z,
                                              oic.ALA2_RIGID_BLOCK, keepdims=keep)
    *ics, dlogp = ic(xyz)
    back, dinv = ic(*ics, inverse=True)
    torch.testing.assert_close(back, xyz, atol=1e-4, rtol=0)
    torch.testing.assert_close(dlogp + dinv, torch.zeros_like(dlogp), atol=1e-3, rtol=0)
    if keep is None:
        plan = oic.make_rel_plan(oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK)
        ref = oic.rel_xyz_to_ic(plan, xyz.cpu().double())
        for got, want in zip((*ics, dlogp), ref):
            np.testing.assert_allclose(got.cpu().double().numpy(), want.numpy(),
                                       atol=1e-3 if want.shape[-1] == 1 else 1e-4)


def test_small_chain_and_properties():
    """tests/nn/flow/crd_transform/test_ic.py:452-493: 5 atoms, fixed = (0, 1, 2)."""
    zmat = np.array([[3, 2, 1, 0], [4, 3, 2, 1]])
    fixed = np.array([0, 1, 2])
    ic = bg.RelativeInternalCoordinateTransformation(zmat, fixed)
    x = torch.randn(10, 15, generator=torch.Generator().manual_seed(0)).to(DEV)
    ics = ic.forward(x)
    assert ics[0].shape == (10, ic.dim_bonds) and ics[3].shape == (10, ic.dim_fixed) == (10, 9)
    assert np.allclose(zmat, ic.z_matrix) and np.allclose(fixed, ic.fixed_atoms) and ic.normalize_angles
    back, _ = ic(*ics[:-1], inverse=True)
    torch.testing.assert_close(back, x, atol=1e-4, rtol=0)
    data = torch.randn(1000, 15, generator=torch.Generator().manual_seed(1))
    mixed = bg.MixedCoordinateTransformation(data, zmat, fixed, keepdims=6)
    ics = mixed.forward(x)
    assert ics[3].shape == (10, 6) and mixed.dim_fixed == 6
    plan = oic.make_rel_plan(zmat, fixed)
    white = oic.Whitening(data.double().numpy()[:, :9], keepdims=6)
    ref = oic.mixed_xyz_to_ic(plan, white, x.cpu().double())
    for got, want in zip(ics, ref):
        np.testing.assert_allclose(got.cpu().double().numpy(), want.numpy(), atol=2e-3 if want.shape[-1] == 1 else 2e-4,
                                   rtol=1e-4)


def test_large_molecule_global_memory_path():
    n = 150
    z = oic.chain_z_matrix(n)[3:]
    fixed = np.array([0, 1, 2])
    g = torch.Generator().manual_seed(0)
    chain = torch.cumsum(torch.randn(n, 3, generator=g, dtype=torch.float64) * 0.6 + 0.5, dim=0)
    xyz = (chain.reshape(1, -1) + 0.02 * torch.randn(40, 3 * n, generator=g, dtype=torch.float64))
    ic = bg.RelativeInternalCoordinateTransformation(z, fixed)
    *ics, dlogp = ic(xyz.float().to(DEV))
    ref = oic.rel_xyz_to_ic(oic.make_rel_plan(z, fixed), xyz)
    for got, want in zip(ics, ref):
        np.testing.assert_allclose(got.cpu().double().numpy(), want.numpy(), atol=1e-3)
    np.testing.assert_allclose(dlogp.cpu().double().numpy(), ref[-1].numpy(), rtol=1e-5, atol=5e-2)
    back, dinv = ic(*ics, inverse=True)
    torch.testing.assert_close(back, xyz.float().to(DEV), atol=2e-2, rtol=0)


@pytest.mark.parametrize("keep", [None, 9])
def test_gradients_match_oracle(keep):
    g = torch.Generator().manual_seed(7)
    x0 = torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float64).reshape(1, -1)
    xyz64 = x0 + 0.01 * torch.randn(24, 66, generator=g, dtype=torch.float64)
    oplan = oic.make_rel_plan(oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK)
    if keep is None:
        ic = bg.RelativeInternalCoordinateTransformation(oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK)
        fwd = lambda x: oic.rel_xyz_to_ic(oplan, x)
    else:
        data = x0 + 0.02 * torch.randn(300, 66, generator=g, dtype=torch.float64)
        ic = bg.MixedCoordinateTransformation(data.float(), oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK, keepdims=keep)
        w = ic._plan.whitening
        white = oic.Whitening(np.zeros((4, 15)), keepdims=keep)            # reuse the layer's own fp32 PCA
        white.mean, white.whiten, white.blacken = (np.asarray(w[k], dtype=np.float64) for k in ("mean", "whiten", "blacken"))
        white.jacobian_xz = w["jacobian_xz"]
        fwd = lambda x: oic.mixed_xyz_to_ic(oplan, white, x)
    x = xyz64.float().to(DEV).requires_grad_(True)
    outs = ic(x)
    ws = [torch.randn(o.shape, generator=g, dtype=torch.float64) for o in outs]
    (gx,) = torch.autograd.grad(sum((o * w.float().to(DEV)).sum() for o, w in zip(outs, ws)), x)
    xr = xyz64.clone().requires_grad_(True)
    (gr,) = torch.autograd.grad(sum((o * w).sum() for o, w in zip(fwd(xr), ws)), xr)
    np.testing.assert_allclose(gx.cpu().double().numpy(), gr.numpy(), rtol=2e-3, atol=2e-3 * float(gr.abs().max()))
