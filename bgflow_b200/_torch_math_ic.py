"""Differentiable device-side PyTorch definition of the global internal-coordinate transform.

Used ONLY by the backward pass of ``autograd.ic_to_xyz_with_grad`` / ``ic_from_xyz_with_grad``
(recompute-in-backward, like ``_torch_math`` for the coupling blocks): the forward always runs the
IC kernels of ``csrc/bgx_ic.cu``.  Never a forward path.

Geometry: bgflow/nn/flow/crd_transform/ic.py:386-513,633-716 and ic_helper.py:114-293,330-452,
480-680; log-determinants in closed form (``2 ln b + ln sin a`` per placed atom,
``2 ln d01 + 2 ln d12 + ln sin a012`` for the reference frame).  Atoms of one dependency stage are
placed together (``[B, k, 3]`` tensors), stages in sequence.
"""

import math

import numpy as np
import torch

__all__ = ["ic_to_xyz", "ic_from_xyz", "relic_to_xyz", "relic_from_xyz"]

_PI = math.pi
_TWO_PI = 2.0 * math.pi
# the reference's helper torsion for the third seed atom is a float32 pi/2 (ic_helper.py:464)
_T32 = float(np.float32(0.5 * np.pi))


def _stages(plan):
    """rel-row indices grouped by dependency depth (cached on the plan)."""
    cached = getattr(plan, "_torch_stages", None)
    if cached is not None:
        return cached
    depth = {int(s): 0 for s in plan.seeds}
    rows_by_depth = {}
    for r in plan.order:                        # plan.order is already topological
        i, j, k, l = (int(v) for v in plan.rel[r])
        d = 1 + max(depth[j], depth[k], depth[l])
        depth[i] = d
        rows_by_depth.setdefault(d, []).append(int(r))
    stages = [rows_by_depth[d] for d in sorted(rows_by_depth)]
    plan._torch_stages = stages
    return stages


def _unit(v, eps):
    return v / torch.linalg.norm(v, dim=-1, keepdim=True).clamp_min(eps)


def _cos_sin_angle(a, b, c, eps):
    """cos / sin of the angle at b (clamped like ic_helper.py:168-210)."""
    cos = (_unit(a - b, eps) * _unit(c - b, eps)).sum(-1).clamp(-1.0 + eps, 1.0 - eps)
    return cos, torch.sqrt(1.0 - cos * cos)


def ic_from_xyz(plan, xyz):
    """xyz ``[B, 3N]`` -> (bonds, angles, torsions, x0 ``[B,1,3]``, R ``[B,3]``, dlogp ``[B,1]``)."""
    eps = plan.eps
    B = xyz.shape[0]
    x = xyz.reshape(B, -1, 3)
    idx = torch.as_tensor(np.ascontiguousarray(plan.rel.T), device=x.device)          # [4, n_rel]
    pi, pj, pk, pl = (x.index_select(1, idx[c]) for c in range(4))
    rij = pj - pi
    bond = torch.linalg.norm(rij, dim=-1).clamp_min(eps)
    cos_a, sin_a = _cos_sin_angle(pi, pj, pk, eps)
    angle = torch.acos(cos_a)
    # torsion (i, j, k, l): project the two outer bonds on the plane normal to the j->k axis
    axis = _unit(pk - pj, eps)
    out0 = pi - pj
    out1 = pl - pk
    v = out0 - (out0 * axis).sum(-1, keepdim=True) * axis
    w = out1 - (out1 * axis).sum(-1, keepdim=True) * axis
    torsion = torch.atan2((torch.linalg.cross(axis, v, dim=-1) * w).sum(-1), (v * w).sum(-1))
    dlogp = -(2.0 * torch.log(bond) + torch.log(sin_a)).sum(-1, keepdim=True)

    s0, s1, s2 = (x[:, int(s)] for s in plan.seeds)
    d01 = torch.linalg.norm(s1 - s0, dim=-1, keepdim=True).clamp_min(eps)
    d12 = torch.linalg.norm(s2 - s1, dim=-1, keepdim=True).clamp_min(eps)
    cos0, sin0 = _cos_sin_angle(s0, s1, s2, eps)
    a012 = torch.acos(cos0)[:, None]
    ez = _unit(s1 - s0, eps)                                   # body z axis
    ey = -_unit(torch.linalg.cross(s2 - s0, ez, dim=-1), eps)  # body y axis
    ex = torch.linalg.cross(ey, ez, dim=-1)                    # body x axis (== -e3 of the tripod)
    alpha = torch.atan2(ez[:, 0], -ez[:, 1])
    beta = ez[:, 2]
    gamma = torch.atan2(ex[:, 2], ey[:, 2])
    dlogp = dlogp - (2.0 * torch.log(d01) + 2.0 * torch.log(d12) + torch.log(sin0)[:, None])
    if plan.normalize_angles:
        angle = angle / _PI
        torsion = (torsion + _PI) / _TWO_PI
        a012 = a012 / _PI
        alpha = (alpha + _PI) / _TWO_PI
        gamma = (gamma + _PI) / _TWO_PI
        dlogp = dlogp - (angle.shape[-1] + 1) * math.log(_PI) - (torsion.shape[-1] + 2) * math.log(_TWO_PI)
    bonds = torch.cat([d01, d12, bond], dim=-1)
    angles = torch.cat([a012, angle], dim=-1)
    return bonds, angles, torsion, s0[:, None, :], torch.stack([alpha, beta, gamma], dim=-1), dlogp


def ic_to_xyz(plan, bonds, angles, torsions, x0, R):
    """(bonds ``[B,N-1]``, angles ``[B,N-2]``, torsions ``[B,N-3]``, x0, R) -> xyz ``[B,3N]``, dlogp."""
    eps = plan.eps
    B = bonds.shape[0]
    origin = x0.reshape(-1, 3).expand(B, 3)
    R = R.reshape(-1, 3).expand(B, 3)
    alpha, beta, gamma = R[:, 0], R[:, 1], R[:, 2]
    const = 0.0
    if plan.normalize_angles:
        angles = angles * _PI
        torsions = torsions * _TWO_PI - _PI
        alpha = alpha * _TWO_PI - _PI
        gamma = gamma * _TWO_PI - _PI
        const = angles.shape[-1] * math.log(_PI) + (torsions.shape[-1] + 2) * math.log(_TWO_PI)
    d01, d12, a012 = bonds[:, 0], bonds[:, 1], angles[:, 0]
    dlogp = (2.0 * torch.log(bonds) ).sum(-1, keepdim=True) + torch.log(torch.sin(angles)).sum(-1, keepdim=True) + const

    # seed atoms in the body frame: s0 = 0, s1 = (0, 0, d01), s2 in the plane given by the helper torsion
    ca, sa = torch.cos(alpha), torch.sin(alpha)
    cg, sg = torch.cos(gamma), torch.sin(gamma)
    cb = beta
    sb = torch.sin(torch.acos(beta))
    # columns of Rz(alpha) Rx(theta) Rz(gamma), theta = acos(beta) (ic_helper.py:344-368)
    col_x = torch.stack([ca * cg - sa * cb * sg, sa * cg + ca * cb * sg, sb * sg], dim=-1)
    col_y = torch.stack([-ca * sg - sa * cb * cg, -sa * sg + ca * cb * cg, sb * cg], dim=-1)
    col_z = torch.stack([sa * sb, -ca * sb, cb], dim=-1)
    hx, hy = math.sin(_T32), -math.cos(_T32)
    hn = math.hypot(hx, hy)
    lateral = (d12 * torch.sin(a012))[:, None]
    along = (d01 - d12 * torch.cos(a012))[:, None]
    p1 = origin + col_z * d01[:, None]
    p2 = origin + col_x * (lateral * (hx / hn)) + col_y * (lateral * (hy / hn)) + col_z * along

    pos = {int(plan.seeds[0]): origin, int(plan.seeds[1]): p1, int(plan.seeds[2]): p2}
    rb, ra, rt = bonds[:, 2:], angles[:, 1:], torsions
    for rows in _stages(plan):
        sel = torch.as_tensor(rows, device=bonds.device)
        d = rb.index_select(1, sel)[..., None]
        a = ra.index_select(1, sel)[..., None]
        t = rt.index_select(1, sel)[..., None]
        q1 = torch.stack([pos[int(plan.rel[r, 1])] for r in rows], dim=1)
        q2 = torch.stack([pos[int(plan.rel[r, 2])] for r in rows], dim=1)
        q3 = torch.stack([pos[int(plan.rel[r, 3])] for r in rows], dim=1)
        v1 = q1 - q2
        normal = torch.linalg.cross(v1, q1 - q3, dim=-1)
        inplane = _unit(torch.linalg.cross(v1, normal, dim=-1), eps)
        normal = _unit(normal, eps)
        direction = _unit(inplane * torch.cos(t) - normal * torch.sin(t), eps)
        new = q1 + direction * (d * torch.sin(a)) - _unit(v1, eps) * (d * torch.cos(a))
        for c, r in enumerate(rows):
            pos[int(plan.rel[r, 0])] = new[:, c]
    xyz = torch.stack([pos[i] for i in range(plan.n_atoms)], dim=1).reshape(B, -1)
    return xyz, dlogp


# ---- relative / mixed transforms (backward definitions of bgx_relic_*) -----------------------------

def _whiten_tensors(plan, like):
    w = plan.whitening
    mk = lambda a: torch.as_tensor(np.asarray(a), dtype=like.dtype, device=like.device)
    return mk(w["mean"]), mk(w["whiten"]), mk(w["blacken"])


def relic_from_xyz(plan, xyz):
    """xyz ``[B, 3N]`` -> (bonds, angles, torsions ``[B, n_rel]``, fixed block, dlogp)
    (ic.py:386-433; with ``plan.whitening`` additionally pca.py:83-91)."""
    eps = plan.eps
    B = xyz.shape[0]
    x = xyz.reshape(B, -1, 3)
    idx = torch.as_tensor(np.ascontiguousarray(plan.rel.T), device=x.device)
    pi, pj, pk, pl = (x.index_select(1, idx[c]) for c in range(4))
    bond = torch.linalg.norm(pj - pi, dim=-1).clamp_min(eps)
    cos_a, sin_a = _cos_sin_angle(pi, pj, pk, eps)
    angle = torch.acos(cos_a)
    axis = _unit(pk - pj, eps)
    out0 = pi - pj
    out1 = pl - pk
    v = out0 - (out0 * axis).sum(-1, keepdim=True) * axis
    w = out1 - (out1 * axis).sum(-1, keepdim=True) * axis
    torsion = torch.atan2((torch.linalg.cross(axis, v, dim=-1) * w).sum(-1), (v * w).sum(-1))
    dlogp = -(2.0 * torch.log(bond) + torch.log(sin_a)).sum(-1, keepdim=True)
    if plan.normalize_angles:
        angle = angle / _PI
        torsion = (torsion + _PI) / _TWO_PI
        dlogp = dlogp - angle.shape[-1] * math.log(_PI) - torsion.shape[-1] * math.log(_TWO_PI)
    fixed = x.index_select(1, torch.as_tensor(plan.fixed, device=x.device)).reshape(B, -1)
    if plan.whitening is not None:
        mean, tw, _ = _whiten_tensors(plan, fixed)
        fixed = torch.matmul(fixed - mean, tw)
        dlogp = dlogp + float(plan.whitening["jacobian_xz"])
    return bond, angle, torsion, fixed, dlogp


def relic_to_xyz(plan, bonds, angles, torsions, fixed):
    """(ic.py:435-513; with ``plan.whitening`` additionally pca.py:93-99) -> xyz ``[B, 3N]``, dlogp."""
    eps = plan.eps
    B = bonds.shape[0]
    const = 0.0
    if plan.normalize_angles:
        angles = angles * _PI
        torsions = torsions * _TWO_PI - _PI
        const = angles.shape[-1] * math.log(_PI) + torsions.shape[-1] * math.log(_TWO_PI)
    dlogp = (2.0 * torch.log(bonds)).sum(-1, keepdim=True) + torch.log(torch.sin(angles)).sum(-1, keepdim=True) + const
    fixed = fixed.reshape(B, -1)
    if plan.whitening is not None:
        mean, _, tb = _whiten_tensors(plan, fixed)
        fixed = torch.matmul(fixed, tb) + mean
        dlogp = dlogp - float(plan.whitening["jacobian_xz"])
    xf = fixed.reshape(B, -1, 3)
    pos = {int(a): xf[:, c] for c, a in enumerate(plan.fixed)}
    for rows in _stages(plan):
        sel = torch.as_tensor(rows, device=bonds.device)
        d = bonds.index_select(1, sel)[..., None]
        a = angles.index_select(1, sel)[..., None]
        t = torsions.index_select(1, sel)[..., None]
        q1 = torch.stack([pos[int(plan.rel[r, 1])] for r in rows], dim=1)
        q2 = torch.stack([pos[int(plan.rel[r, 2])] for r in rows], dim=1)
        q3 = torch.stack([pos[int(plan.rel[r, 3])] for r in rows], dim=1)
        v1 = q1 - q2
        normal = torch.linalg.cross(v1, q1 - q3, dim=-1)
        inplane = _unit(torch.linalg.cross(v1, normal, dim=-1), eps)
        normal = _unit(normal, eps)
        direction = _unit(inplane * torch.cos(t) - normal * torch.sin(t), eps)
        new = q1 + direction * (d * torch.sin(a)) - _unit(v1, eps) * (d * torch.cos(a))
        for c, r in enumerate(rows):
            pos[int(plan.rel[r, 0])] = new[:, c]
    xyz = torch.stack([pos[i] for i in range(plan.n_atoms)], dim=1).reshape(B, -1)
    return xyz, dlogp
