"""Global internal-coordinate transform backed by the sm_100a IC kernels.

Mirrors ``GlobalInternalCoordinateTransformation`` (bgflow/nn/flow/crd_transform/ic.py:516-716):
same constructor, same read-only properties (consumed by the reference's
``ShapeDictionary.from_coordinate_transform``, factory/tensor_info.py:86-100), same direction
convention: ``_forward`` maps Cartesian -> (bonds, angles, torsions, x0, R), ``_inverse`` maps
back.  Host work (z-matrix staging, ic.py:25-97) happens once at construction in
``engine.ZPlan``; every call is ONE kernel with closed-form log-determinants.  When gradients are
requested the backward re-evaluates the transform with the device-side PyTorch definition in
``_torch_math_ic`` (recompute-in-backward, like the coupling blocks).
"""

import numpy as np
import torch

from . import autograd as _ag
from . import engine
from .flows import Flow

__all__ = ["GlobalInternalCoordinateTransformation"]


class GlobalInternalCoordinateTransformation(Flow):
    """z_matrix: ``(n_atoms, 4)`` int array; rows ``(i, j, k, l)`` place atom i by its bond to j,
    angle (i, j, k) and torsion (i, j, k, l); the three seed rows carry ``-1`` placeholders.

    ``raise_warnings`` / ``enforce_boundaries`` are accepted for API parity: the kernels always
    clamp with ``eps`` exactly like ``enforce_boundaries=True`` and never synchronise to warn.
    """

    _accumulates_dlogp = False

    def __init__(self, z_matrix, normalize_angles=True, eps: float = 1e-7, enforce_boundaries: bool = True,
                 raise_warnings: bool = True):
        super().__init__()
        if isinstance(z_matrix, torch.Tensor):
            z_matrix = z_matrix.cpu().numpy()
        if not enforce_boundaries:
            raise NotImplementedError("the IC kernels always enforce the eps boundaries")
        self._plan = engine.ZPlan(z_matrix, normalize_angles=normalize_angles, eps=eps)
        self._raise_warnings = raise_warnings

    # -- properties of ic.py:560-604
    @property
    def z_matrix(self):
        return self._plan.rel

    @property
    def fixed_atoms(self):
        return np.array([], dtype=np.int64)

    @property
    def dim_bonds(self):
        return len(self.z_matrix) + 2

    @property
    def dim_angles(self):
        return len(self.z_matrix) + 1

    @property
    def dim_torsions(self):
        return len(self.z_matrix)

    @property
    def dim_fixed(self):
        return 0

    @property
    def bond_indices(self):
        f = self._plan.seeds
        return np.vstack([np.array([[f[1], f[0]], [f[2], f[1]]]), self._plan.rel[:, :2]])

    @property
    def angle_indices(self):
        f = self._plan.seeds
        return np.vstack([np.array([[f[2], f[1], f[0]]]), self._plan.rel[:, :3]])

    @property
    def torsion_indices(self):
        return self._plan.rel[:, :4]

    @property
    def normalize_angles(self):
        return self._plan.normalize_angles

    def _forward(self, x, *args, **kwargs):
        """xyz ``[B, 3N]`` -> bonds ``[B, N-1]``, angles ``[B, N-2]``, torsions ``[B, N-3]``,
        x0 ``[B, 1, 3]``, R ``[B, 3]``, dlogp ``[B, 1]`` (ic.py:633-676)."""
        if torch.is_grad_enabled() and x.requires_grad:
            return _ag.ic_from_xyz_with_grad(self._plan, engine.ic_from_xyz, x)
        return engine.ic_from_xyz(self._plan, x)

    def _inverse(self, bonds, angles, torsions, x0, R, *args, **kwargs):
        """(ic.py:678-716) -> xyz ``[B, 3N]``, dlogp ``[B, 1]``."""
        ins = (bonds, angles, torsions, x0, R)
        if torch.is_grad_enabled() and any(t.requires_grad for t in ins):
            B = bonds.shape[0]
            x0 = x0.reshape(-1, 1, 3).expand(B, 1, 3)
            R = R.reshape(-1, 3).expand(B, 3)
            return _ag.ic_to_xyz_with_grad(self._plan, engine.ic_to_xyz, bonds, angles, torsions, x0, R)
        return engine.ic_to_xyz(self._plan, *ins)
