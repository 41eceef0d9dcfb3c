"""Oracle (CPU, torch) for the global Z-matrix <-> Cartesian transform.

TEST INFRASTRUCTURE — not product code.  Restates
bgflow/nn/flow/crd_transform/ic.py:25-125,386-513,633-716 and
bgflow/nn/flow/crd_transform/ic_helper.py:114-293,330-452,480-680 (paths relative to
/root/reference).  The geometry follows the reference op for op (including the
``eps`` clamps).  The log-determinants use the closed forms the survey verified
against the reference's Jacobian determinants in fp64:

    placed atom      log|det d(xyz)/d(b,a,t)| =  2 ln b + ln sin a
    reference frame  log|det|                 =  2 ln d01 + 2 ln d12 + ln sin a012

(the reference forms 3x3 Jacobians / a 9x9 autograd Jacobian for the same numbers);
``tests/test_oracle_golden.py`` pins this module to outputs of the reference itself.
"""

import math

import numpy as np
import torch

__all__ = ["ZPlan", "make_plan", "xyz_to_ic", "ic_to_xyz", "ALA2_GLOBAL_Z", "ALA2_XYZ", "chain_z_matrix",
           "RelPlan", "make_rel_plan", "rel_xyz_to_ic", "rel_ic_to_xyz", "Whitening", "mixed_xyz_to_ic",
           "mixed_ic_to_xyz", "ALA2_RIGID_BLOCK", "ALA2_RELATIVE_Z"]


# Fixture data of the reference's own test: tests/nn/flow/crd_transform/test_ic.py:64-89,91-116
ALA2_GLOBAL_Z = np.array([
    [0, -1, -1, -1], [1, 0, -1, -1], [2, 1, 0, -1], [3, 1, 0, 2], [4, 1, 0, 2], [5, 4, 1, 0],
    [6, 4, 1, 5], [7, 6, 4, 1], [8, 6, 4, 7], [9, 8, 6, 4], [10, 8, 6, 9], [14, 8, 6, 9],
    [11, 10, 8, 6], [12, 10, 8, 11], [13, 10, 8, 11], [15, 14, 8, 6], [16, 14, 8, 15],
    [17, 16, 14, 15], [18, 16, 17, 14], [19, 18, 16, 14], [20, 18, 19, 16], [21, 18, 20, 16],
])

ALA2_XYZ = np.array([
    [1.375, 1.25, 1.573], [1.312, 1.255, 1.662], [1.327, 1.306, 1.493], [1.377, 1.143, 1.549],
    [1.511, 1.31, 1.618], [1.606, 1.236, 1.63], [1.523, 1.441, 1.633], [1.445, 1.5, 1.607],
    [1.645, 1.515, 1.667], [1.703, 1.459, 1.74], [1.73, 1.53, 1.54], [1.792, 1.619, 1.554],
    [1.78, 1.439, 1.508], [1.663, 1.555, 1.457], [1.618, 1.646, 1.734], [1.509, 1.703, 1.709],
    [1.715, 1.705, 1.809], [1.798, 1.653, 1.831], [1.703, 1.847, 1.852], [1.801, 1.871, 1.892],
    [1.674, 1.911, 1.768], [1.631, 1.858, 1.933],
])


def chain_z_matrix(n_atoms):
    """Linear-chain global z-matrix (pattern of test_ic.py:302-304 extended to N atoms):
    atom i is placed from (i-1, i-2, i-3): worst-case dependency depth N-3."""
    z = []
    for i in range(n_atoms):
        z.append([i, i - 1 if i >= 1 else -1, i - 2 if i >= 2 else -1, i - 3 if i >= 3 else -1])
    return np.array(z)


class ZPlan:
    """Host-side description of a global z-matrix.

    seeds        the three atoms that define the reference frame (ic.py:94-97)
    rel          remaining rows ``(i, j, k, l)`` in their original order; this is the
                 column order of the bond/angle/torsion tensors (ic.py:376-378)
    order        a placement order of ``rel`` rows such that j, k, l are placed before i
                 (any order compatible with ic.py:25-91's stages gives identical results)
    n_atoms
    """

    def __init__(self, seeds, rel, order, n_atoms):
        self.seeds = seeds
        self.rel = rel
        self.order = order
        self.n_atoms = n_atoms


def make_plan(z_matrix):
    z = np.asarray(z_matrix)
    n_minus = np.sum(z == -1, axis=-1)
    # the rows with 3, 2 and 1 undefined references, in that order (ic.py:94-97)
    seed_rows = [int(np.where(n_minus == c)[0][0]) for c in (3, 2, 1)]
    seeds = [int(z[r, 0]) for r in seed_rows]
    rel = z[n_minus == 0]
    placed = set(seeds)
    remaining = list(range(len(rel)))
    order = []
    while remaining:
        stage = [r for r in remaining if all(int(a) in placed for a in rel[r, 1:])]
        if not stage:
            raise ValueError("z-matrix decomposition failed: atoms not reachable from the seed atoms")
        for r in stage:
            order.append(r)
        placed.update(int(rel[r, 0]) for r in stage)
        remaining = [r for r in remaining if r not in set(stage)]
    return ZPlan(seeds, rel, order, len(z))


def _norm(v, eps):
    return torch.linalg.norm(v, dim=-1, keepdim=True).clamp_min(eps)


def _angle(x1, x2, x3, eps):
    """ic_helper.py:168-210: angle at x2 between x1 and x3; returns (angle, sin(angle))."""
    r12 = x1 - x2
    r32 = x3 - x2
    c = ((r12 / _norm(r12, eps)) * (r32 / _norm(r32, eps))).sum(-1)
    c = c.clamp(-1.0 + eps, 1.0 - eps)
    return torch.acos(c), torch.sqrt(1.0 - c * c)


def _torsion(x1, x2, x3, x4, eps):
    """ic_helper.py:213-293."""
    b0 = x1 - x2
    b1 = x3 - x2
    b2 = x4 - x3
    u = b1 / _norm(b1, eps)
    v = b0 - (b0 * u).sum(-1, keepdim=True) * u
    w = b2 - (b2 * u).sum(-1, keepdim=True) * u
    x = (v * w).sum(-1)
    y = (torch.linalg.cross(u, v, dim=-1) * w).sum(-1)
    return torch.atan2(y, x)


def xyz_to_ic(plan, xyz, normalize_angles=True, eps=1e-7):
    """GlobalInternalCoordinateTransformation._forward (ic.py:633-676).

    xyz ``[B, 3N]`` -> bonds ``[B, N-1]``, angles ``[B, N-2]``, torsions ``[B, N-3]``,
    x0 ``[B, 1, 3]``, R ``[B, 3]``, dlogp ``[B, 1]``.
    """
    b = xyz.shape[0]
    x = xyz.reshape(b, -1, 3)
    rel = plan.rel
    xi, xj, xk, xl = (x[:, rel[:, c]] for c in range(4))
    bonds = _norm(xj - xi, eps)[..., 0]                       # ic_helper.py:148-165
    angles, sin_a = _angle(xi, xj, xk, eps)
    torsions = _torsion(xi, xj, xk, xl, eps)
    dlogp = -(2 * torch.log(bonds) + torch.log(sin_a)).sum(-1, keepdim=True)

    s0, s1, s2 = (x[:, s] for s in plan.seeds)
    d01 = _norm(s1 - s0, eps)                                 # [B,1]
    d12 = _norm(s2 - s1, eps)
    a012, sin012 = _angle(s0, s1, s2, eps)                   # [B]
    # tripod (ic_helper.py:114-138) + euler angles (ic_helper.py:330-341)
    e1 = (s1 - s0) / _norm(s1 - s0, eps)
    e2 = torch.linalg.cross(s2 - s0, e1, dim=-1)
    e2 = e2 / _norm(e2, eps)
    e3 = torch.linalg.cross(e2, e1, dim=-1)
    bx, by, bz = -e3, -e2, e1
    alpha = torch.atan2(bz[..., 0], -bz[..., 1])
    beta = bz[..., 2]
    gamma = torch.atan2(bx[..., 2], by[..., 2])
    dlogp = dlogp - (2 * torch.log(d01) + 2 * torch.log(d12) + torch.log(sin012)[:, None])

    a012 = a012[:, None]
    if normalize_angles:                                      # ic.py:100-111,193-203,424-427
        angles = angles / math.pi
        torsions = (torsions + math.pi) / (2 * math.pi)
        a012 = a012 / math.pi
        alpha = (alpha + math.pi) / (2 * math.pi)
        gamma = (gamma + math.pi) / (2 * math.pi)
        n_a = angles.shape[-1] + 1
        n_t = torsions.shape[-1] + 2
        dlogp = dlogp - n_a * math.log(math.pi) - n_t * math.log(2 * math.pi)

    bonds = torch.cat([d01, d12, bonds], dim=-1)
    angles = torch.cat([a012, angles], dim=-1)
    orientation = torch.stack([alpha, beta, gamma], dim=-1)
    return bonds, angles, torsions, s0[:, None, :], orientation, dlogp


def _place(p1, p2, p3, d, a, t, eps):
    """ic_helper.py:372-432 (position only)."""
    v1 = p1 - p2
    v2 = p1 - p3
    n = torch.linalg.cross(v1, v2, dim=-1)
    nn = torch.linalg.cross(v1, n, dim=-1)
    n = n / _norm(n, eps)
    nn = nn / _norm(nn, eps)
    v3 = n * -torch.sin(t) + nn * torch.cos(t)
    v3 = v3 / _norm(v3, eps)
    v1 = v1 / _norm(v1, eps)
    return p1 + v3 * d * torch.sin(a) - v1 * d * torch.cos(a)


def ic_to_xyz(plan, bonds, angles, torsions, x0, orientation, normalize_angles=True, eps=1e-7):
    """GlobalInternalCoordinateTransformation._inverse (ic.py:678-716).

    Returns xyz ``[B, 3N]`` and dlogp ``[B, 1]``.
    """
    b = bonds.shape[0]
    x0 = x0.reshape(-1, 1, 3).expand(b, 1, 3)[:, 0]
    alpha, beta, gamma = orientation[:, 0:1], orientation[:, 1:2], orientation[:, 2:3]
    const = 0.0
    if normalize_angles:                                      # ic.py:114-125,238-248,441-444
        angles = angles * math.pi
        torsions = torsions * (2 * math.pi) - math.pi
        alpha = alpha * (2 * math.pi) - math.pi
        gamma = gamma * (2 * math.pi) - math.pi
        const = angles.shape[-1] * math.log(math.pi) + (torsions.shape[-1] + 2) * math.log(2 * math.pi)
    d01, d12, a012 = bonds[:, 0:1], bonds[:, 1:2], angles[:, 0:1]
    rb, ra, rt = bonds[:, 2:], angles[:, 1:], torsions

    # ic_helper.py:344-368: R = Rz(alpha) Rx(acos(beta)) Rz(gamma)
    theta = torch.acos(beta)
    ca, sa, cb, sb, cg, sg = (f(v)[:, 0] for v in (alpha, theta, gamma) for f in (torch.cos, torch.sin))
    one, zero = torch.ones_like(ca), torch.zeros_like(ca)

    def mat(rows):
        return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)

    rz_a = mat([[ca, -sa, zero], [sa, ca, zero], [zero, zero, one]])
    rx_b = mat([[one, zero, zero], [zero, cb, -sb], [zero, sb, cb]])
    rz_g = mat([[cg, -sg, zero], [sg, cg, zero], [zero, zero, one]])
    rot = rz_a @ rx_b @ rz_g

    # ic_helper.py:511-540 + 455-477: p1 on the z axis, p2 placed with torsion pi/2 against the
    # helper point (0,-1,0).  The reference builds pi/2 as a float32 tensor (ic_helper.py:464), so
    # cos(t) = -4.37e-8 instead of 0 and p2 gets a ~4e-8 relative y component: restated as is.
    zeros = torch.zeros_like(d01)
    p1 = torch.cat([zeros, zeros, d01], dim=-1)
    t32 = float(np.float32(0.5 * np.pi))
    vx, vy = math.sin(t32), -math.cos(t32)
    vn = math.sqrt(vx * vx + vy * vy)
    ds = d12 * torch.sin(a012)
    p2 = torch.cat([ds * (vx / vn), ds * (vy / vn), d01 - d12 * torch.cos(a012)], dim=-1)
    x1 = (rot @ p1[..., None])[..., 0] + x0
    x2 = (rot @ p2[..., None])[..., 0] + x0
    dlogp = 2 * torch.log(d01) + 2 * torch.log(d12) + torch.log(torch.sin(a012)) + const

    pos = [None] * plan.n_atoms
    pos[plan.seeds[0]], pos[plan.seeds[1]], pos[plan.seeds[2]] = x0, x1, x2
    for r in plan.order:
        i, j, k, l = (int(v) for v in plan.rel[r])
        d, a, t = rb[:, r:r + 1], ra[:, r:r + 1], rt[:, r:r + 1]
        pos[i] = _place(pos[j], pos[k], pos[l], d, a, t, eps)
        dlogp = dlogp + 2 * torch.log(d) + torch.log(torch.sin(a))
    xyz = torch.stack(pos, dim=1).reshape(b, -1)
    return xyz, dlogp


# ------------------------------------------------------------------------------------------------
# Relative and mixed transforms
# ------------------------------------------------------------------------------------------------

class RelPlan:
    """Host-side description of a relative z-matrix (ic.py:268-384).

    fixed   atoms that stay Cartesian, in the column order of the ``x_fixed`` tensor (ic.py:419)
    rel     z-matrix rows ``(i, j, k, l)`` = column order of bonds / angles / torsions
    order   placement order of the rows (any order compatible with the stages of
            decompose_z_matrix, ic.py:25-91, gives identical results)
    """

    def __init__(self, fixed, rel, order, n_atoms):
        self.fixed = fixed
        self.rel = rel
        self.order = order
        self.n_atoms = n_atoms


def make_rel_plan(z_matrix, fixed_atoms):
    rel = np.asarray(z_matrix)
    fixed = [int(a) for a in np.asarray(fixed_atoms)]
    placed = set(fixed)
    remaining = list(range(len(rel)))
    order = []
    while remaining:
        stage = [r for r in remaining if all(int(a) in placed for a in rel[r, 1:])]
        if not stage:
            raise ValueError("z-matrix decomposition failed: atoms not reachable from the fixed atoms")
        order.extend(stage)
        placed.update(int(rel[r, 0]) for r in stage)
        remaining = [r for r in remaining if r not in set(stage)]
    return RelPlan(fixed, rel, order, len(rel) + len(fixed))


def rel_xyz_to_ic(plan, xyz, normalize_angles=True, eps=1e-7):
    """RelativeInternalCoordinateTransformation._forward (ic.py:386-433):
    xyz ``[B, 3N]`` -> bonds, angles, torsions ``[B, n_rel]``, x_fixed ``[B, 3 n_fixed]``, dlogp."""
    b = xyz.shape[0]
    x = xyz.reshape(b, -1, 3)
    rel = plan.rel
    xi, xj, xk, xl = (x[:, rel[:, c]] for c in range(4))
    bonds = _norm(xj - xi, eps)[..., 0]
    angles, sin_a = _angle(xi, xj, xk, eps)
    torsions = _torsion(xi, xj, xk, xl, eps)
    dlogp = -(2 * torch.log(bonds) + torch.log(sin_a)).sum(-1, keepdim=True)
    x_fixed = x[:, plan.fixed].reshape(b, -1)
    if normalize_angles:
        angles = angles / math.pi
        torsions = (torsions + math.pi) / (2 * math.pi)
        dlogp = dlogp - angles.shape[-1] * math.log(math.pi) - torsions.shape[-1] * math.log(2 * math.pi)
    return bonds, angles, torsions, x_fixed, dlogp


def rel_ic_to_xyz(plan, bonds, angles, torsions, x_fixed, normalize_angles=True, eps=1e-7):
    """RelativeInternalCoordinateTransformation._inverse (ic.py:435-513)."""
    b = x_fixed.shape[0]
    const = 0.0
    if normalize_angles:
        angles = angles * math.pi
        torsions = torsions * (2 * math.pi) - math.pi
        const = angles.shape[-1] * math.log(math.pi) + torsions.shape[-1] * math.log(2 * math.pi)
    xf = x_fixed.reshape(b, -1, 3)
    pos = [None] * plan.n_atoms
    for c, atom in enumerate(plan.fixed):
        pos[atom] = xf[:, c]
    dlogp = torch.zeros(b, 1, dtype=x_fixed.dtype) + const
    for r in plan.order:
        i, j, k, l = (int(v) for v in plan.rel[r])
        d, a, t = bonds[:, r:r + 1], angles[:, r:r + 1], torsions[:, r:r + 1]
        pos[i] = _place(pos[j], pos[k], pos[l], d, a, t, eps)
        dlogp = dlogp + 2 * torch.log(d) + torch.log(torch.sin(a))
    return torch.stack(pos, dim=1).reshape(b, -1), dlogp


class Whitening:
    """Static PCA whitening of the fixed block: WhitenFlow(X0, keepdims, whiten_inverse=False)
    (crd_transform/pca.py:10-34,37-107), computed in numpy float64 exactly like ``_pca``."""

    def __init__(self, x0, keepdims=None):
        x0 = np.asarray(x0, dtype=np.float64) if not isinstance(x0, np.ndarray) else x0
        if keepdims is None:
            keepdims = x0.shape[1]
        mean = x0.mean(axis=0)
        xm = x0 - mean
        cov = np.matmul(xm.T, xm) / (xm.shape[0] - 1.0)
        eigval, eigvec = np.linalg.eigh(cov)
        idx = np.argsort(eigval)[::-1][:keepdims]
        std = np.sqrt(eigval[idx])
        eigvec = eigvec[:, idx]
        self.mean = mean
        self.whiten = np.matmul(eigvec, np.diag(1.0 / std))      # [dim, keep]
        self.blacken = np.matmul(np.diag(std), eigvec.T)         # [keep, dim]
        self.std = std
        self.keepdims = keepdims
        self.jacobian_xz = -float(np.sum(np.log(std)))


def mixed_xyz_to_ic(plan, white, xyz, normalize_angles=True, eps=1e-7):
    """MixedCoordinateTransformation._forward (ic.py:836-860)."""
    bonds, angles, torsions, x_fixed, dlogp = rel_xyz_to_ic(plan, xyz, normalize_angles, eps)
    mean = torch.as_tensor(white.mean, dtype=xyz.dtype)
    tw = torch.as_tensor(white.whiten, dtype=xyz.dtype)
    z_fixed = torch.matmul(x_fixed - mean, tw)
    return bonds, angles, torsions, z_fixed, dlogp + white.jacobian_xz


def mixed_ic_to_xyz(plan, white, bonds, angles, torsions, z_fixed, normalize_angles=True, eps=1e-7):
    """MixedCoordinateTransformation._inverse (ic.py:862-884)."""
    mean = torch.as_tensor(white.mean, dtype=z_fixed.dtype)
    tb = torch.as_tensor(white.blacken, dtype=z_fixed.dtype)
    x_fixed = torch.matmul(z_fixed, tb) + mean
    xyz, dlogp = rel_ic_to_xyz(plan, bonds, angles, torsions, x_fixed, normalize_angles, eps)
    return xyz, dlogp - white.jacobian_xz


# fixture of the reference's own test: tests/nn/flow/crd_transform/test_ic.py:37-62
ALA2_RIGID_BLOCK = np.array([6, 8, 9, 10, 14])
ALA2_RELATIVE_Z = np.array([
    [0, 1, 4, 6], [1, 4, 6, 8], [2, 1, 4, 0], [3, 1, 4, 0], [4, 6, 8, 14], [5, 4, 6, 8], [7, 6, 8, 4],
    [11, 10, 8, 6], [12, 10, 8, 11], [13, 10, 8, 11], [15, 14, 8, 16], [16, 14, 8, 6], [17, 16, 14, 15],
    [18, 16, 14, 8], [19, 18, 16, 14], [20, 18, 16, 19], [21, 18, 16, 19],
])
