"""Independent pin of the closed-form log-determinants the oracle (and the kernels) use for the
internal-coordinate transforms: the full 3N x 3N Jacobian of xyz -> (ICs) from autograd in fp64,
log|det| against the oracle's dlogp — global (reference frame + z-matrix), relative and mixed
transforms, normalised and raw angles, random tree-shaped z-matrices.  The reference forms per-atom
3x3 Jacobians and a 9x9 autograd Jacobian for the same numbers (ic.py:430-431, 503, ic_helper.py:655-678)."""

import math

import numpy as np
import pytest
import torch

from oracle import ic as oic


def random_global_z(n, rng):
    """Atom i >= 3 is placed from three distinct earlier atoms (a random tree-like z-matrix)."""
    z = [[0, -1, -1, -1], [1, 0, -1, -1], [2, 1, 0, -1]]
    for i in range(3, n):
        j, k, l = rng.choice(i, size=3, replace=False)
        z.append([i, int(j), int(k), int(l)])
    perm = rng.permutation(n)                       # relabel the atoms: seeds need not be 0, 1, 2
    return np.array([[perm[a] if a >= 0 else -1 for a in row] for row in z])


def logabsdet_jacobian(fn, x):
    jac = torch.autograd.functional.jacobian(fn, x)          # [out, in] for a single sample
    sign, ld = torch.linalg.slogdet(jac)
    return float(ld)


@pytest.mark.parametrize("normalize", [True, False])
@pytest.mark.parametrize("n_atoms,seed", [(4, 0), (6, 1), (9, 2)])
def test_global_ic_logdet_is_the_jacobian_determinant(n_atoms, seed, normalize):
    rng = np.random.default_rng(seed)
    z = random_global_z(n_atoms, rng)
    plan = oic.make_plan(z)
    x = torch.from_numpy(rng.standard_normal(3 * n_atoms) * 0.7)

    def fwd(v):
        b, a, t, x0, R, _ = oic.xyz_to_ic(plan, v[None], normalize_angles=normalize)
        return torch.cat([b[0], a[0], t[0], x0.reshape(-1), R[0]])

    dlogp = oic.xyz_to_ic(plan, x[None], normalize_angles=normalize)[-1]
    assert float(dlogp) == pytest.approx(logabsdet_jacobian(fwd, x), abs=1e-8)
    # inverse direction: the determinant of the inverse map at the image point
    ics = oic.xyz_to_ic(plan, x[None], normalize_angles=normalize)
    nb, na, nt = n_atoms - 1, n_atoms - 2, n_atoms - 3
    flat = fwd(x)

    def inv(u):
        b, a, t = u[:nb][None], u[nb:nb + na][None], u[nb + na:nb + na + nt][None]
        x0, R = u[-6:-3].reshape(1, 1, 3), u[-3:][None]
        return oic.ic_to_xyz(plan, b, a, t, x0, R, normalize_angles=normalize)[0][0]

    xyz, dinv = oic.ic_to_xyz(plan, *ics[:5], normalize_angles=normalize)
    # (the reference places the third seed atom with a float32 pi/2, ic_helper.py:464, restated as is:
    # the fp64 round trip closes to ~1e-7 only)
    torch.testing.assert_close(xyz[0], x, atol=5e-7, rtol=0)
    assert float(dinv) == pytest.approx(logabsdet_jacobian(inv, flat), abs=1e-6)
    assert float(dinv + dlogp) == pytest.approx(0.0, abs=1e-9)


@pytest.mark.parametrize("keep", [None, 9, 6])
def test_relative_and_mixed_logdet_is_the_jacobian_determinant(keep):
    rng = np.random.default_rng(5)
    n = 8
    fixed = np.array([4, 1, 6])
    others = [a for a in range(n) if a not in set(fixed.tolist())]
    placed, z = list(fixed), []
    for i in rng.permutation(others):
        j, k, l = rng.choice(placed, size=3, replace=False)
        z.append([int(i), int(j), int(k), int(l)])
        placed.append(int(i))
    z = np.array(z)[rng.permutation(len(z))]                # column order != placement order
    plan = oic.make_rel_plan(z, fixed)
    x = torch.from_numpy(rng.standard_normal(3 * n) * 0.7)
    if keep is None:
        def fwd(v):
            b, a, t, f, _ = oic.rel_xyz_to_ic(plan, v[None])
            return torch.cat([b[0], a[0], t[0], f[0]])
        dlogp = oic.rel_xyz_to_ic(plan, x[None])[-1]
        assert float(dlogp) == pytest.approx(logabsdet_jacobian(fwd, x), abs=1e-8)
        return
    data = rng.standard_normal((200, 9)) @ rng.standard_normal((9, 9)) * 0.3 + 1.0
    white = oic.Whitening(data, keepdims=keep)
    b, a, t, zf, dlogp = oic.mixed_xyz_to_ic(plan, white, x[None])
    assert zf.shape == (1, keep)
    if keep == 9:                                            # invertible whitening: square Jacobian
        def fwd(v):
            b_, a_, t_, f_, _ = oic.mixed_xyz_to_ic(plan, white, v[None])
            return torch.cat([b_[0], a_[0], t_[0], f_[0]])
        assert float(dlogp) == pytest.approx(logabsdet_jacobian(fwd, x), abs=1e-7)
        xyz, dinv = oic.mixed_ic_to_xyz(plan, white, b, a, t, zf)
        torch.testing.assert_close(xyz[0], x, atol=1e-8, rtol=0)
        assert float(dlogp + dinv) == pytest.approx(0.0, abs=1e-9)
    else:
        # keepdims < 9: the reference's convention is -sum(log std) of the kept components (pca.py:72)
        rel = oic.rel_xyz_to_ic(plan, x[None])[-1]
        assert float(dlogp - rel) == pytest.approx(-float(np.sum(np.log(white.std))), abs=1e-12)
