"""Parity of the IC kernels against the reference-generated golden fixtures and the oracle.
Tolerances follow the reference's own cpu-fp32 row (tests/nn/flow/crd_transform/test_ic.py:28-31):
coordinates / ICs 1e-4, dlogp 1e-3."""

import numpy as np
import pytest
import torch

import bgflow_b200 as bg
from oracle import ic as oic
from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def _cmp(a, b, atol, what=""):
    np.testing.assert_allclose(a.detach().cpu().double().numpy(), np.asarray(b, dtype=np.float64),
                               atol=atol, rtol=1e-4, err_msg=what)


@pytest.mark.parametrize("name", ["ic_ala2", "ic_ala2_raw", "ic_chain12"])
def test_ic_matches_reference_golden(name):
    g = load_golden(name)
    ic = bg.GlobalInternalCoordinateTransformation(g["z_matrix"], normalize_angles=bool(int(g["normalize"])))
    bonds, angles, torsions, x0, R, dlogp = ic._forward(_t(g["xyz_f32"]))
    B = g["xyz_f32"].shape[0]
    assert x0.shape == (B, 1, 3) and R.shape == (B, 3) and dlogp.shape == (B, 1)
    for got, key in ((bonds, "bonds"), (angles, "angles"), (torsions, "torsions"), (x0, "x0"), (R, "R")):
        _cmp(got, g[key + "_f64"], 1e-4, key)
    _cmp(dlogp, g["dlogp_f64"], 1e-3, "dlogp")
    xyz, dinv = ic._inverse(*(_t(g[k + "_f32"]) for k in ("bonds", "angles", "torsions", "x0", "R")))
    _cmp(xyz, g["xyz_back_f64"], 1e-4, "xyz_back")
    _cmp(dinv, g["dlogp_inv_f64"], 1e-3, "dlogp_inv")
    xyz, dgen = ic._inverse(*(_t(g["gen_" + k + "_f32"]) for k in ("bonds", "angles", "torsions", "x0", "R")))
    _cmp(xyz, g["gen_xyz_f64"], 1e-4, "gen_xyz")
    _cmp(dgen, g["gen_dlogp_f64"], 1e-3, "gen_dlogp")


@pytest.mark.parametrize("batch", [1, 127, 128, 129, 5000])
def test_round_trip_ragged(batch):
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    g = torch.Generator().manual_seed(batch)
    xyz = torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float32).reshape(1, -1) + 0.01 * torch.randn(batch, 66, generator=g)
    xyz = xyz.to(DEV)
    *ics, dlogp = ic(xyz)
    back, dinv = ic(*ics, inverse=True)
    torch.testing.assert_close(back, xyz, atol=1e-4, rtol=0)        # test_ic.py:519-550 (atol 1e-3)
    torch.testing.assert_close(dlogp + dinv, torch.zeros_like(dlogp), atol=1e-3, rtol=0)
    plan = oic.make_plan(oic.ALA2_GLOBAL_Z)
    ref = oic.xyz_to_ic(plan, xyz.cpu().double())
    for got, want in zip((*ics, dlogp), ref):
        np.testing.assert_allclose(got.cpu().double().numpy(), want.numpy(), atol=1e-3 if want.shape[-1] == 1 and want.dim() == 2 else 1e-4)


def test_large_molecule_global_memory_path():
    n = 150                                   # 3N x 129 floats do not fit in shared memory
    z = oic.chain_z_matrix(n)
    g = torch.Generator().manual_seed(0)
    chain = torch.cumsum(torch.randn(n, 3, generator=g, dtype=torch.float64) * 0.6 + 0.5, dim=0)
    xyz = (chain.reshape(1, -1) + 0.02 * torch.randn(40, 3 * n, generator=g, dtype=torch.float64))
    ic = bg.GlobalInternalCoordinateTransformation(z)
    *ics, dlogp = ic(xyz.float().to(DEV))
    plan = oic.make_plan(z)
    ref = oic.xyz_to_ic(plan, xyz)
    # a random-walk chain has some nearly collinear triples: torsions there are ill-conditioned
    # (fp32 input rounding alone moves them by ~1e-4), hence 1e-3 here instead of 1e-4
    for got, want in zip(ics, ref):
        np.testing.assert_allclose(got.cpu().double().numpy(), want.numpy(), atol=1e-3)
    np.testing.assert_allclose(dlogp.cpu().double().numpy(), ref[-1].numpy(), rtol=1e-5, atol=5e-2)
    back, dinv = ic(*ics, inverse=True)
    # error accumulates along the 147-deep chain; the reference's own cuda-fp32 tolerance is 1e-2
    torch.testing.assert_close(back, xyz.float().to(DEV), atol=2e-2, rtol=0)


def test_properties_and_errors():
    ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
    assert (ic.dim_bonds, ic.dim_angles, ic.dim_torsions, ic.dim_fixed) == (21, 20, 19, 0)
    assert ic.bond_indices.shape == (21, 2) and ic.angle_indices.shape == (20, 3) and ic.torsion_indices.shape == (19, 4)
    assert len(ic.fixed_atoms) == 0 and ic.normalize_angles
    bad = oic.ALA2_GLOBAL_Z.copy()
    bad[5] = [5, 21, 1, 0]     # atom 21 is placed from 18, which needs 16 <- 14 <- 8 <- 6 <- 4; make a cycle
    bad[21] = [21, 5, 20, 16]
    with pytest.raises(ValueError):
        bg.GlobalInternalCoordinateTransformation(bad)
