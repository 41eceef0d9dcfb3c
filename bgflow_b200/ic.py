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

__all__ = ["GlobalInternalCoordinateTransformation", "RelativeInternalCoordinateTransformation",
           "MixedCoordinateTransformation", "WhitenFlow"]


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


class RelativeInternalCoordinateTransformation(Flow):
    """Internal coordinates relative to a block of fixed atoms (ic.py:268-513): ``_forward`` maps
    xyz ``[B, 3N]`` -> (bonds, angles, torsions ``[B, n]``, x_fixed ``[B, 3 n_fixed]``, dlogp),
    ``_inverse`` maps back.  One kernel per call (``bgx_relic_from_xyz`` / ``bgx_relic_to_xyz``)."""

    _accumulates_dlogp = False

    def __init__(self, z_matrix, fixed_atoms, normalize_angles=True, eps: float = 1e-7,
                 enforce_boundaries: bool = True, raise_warnings: bool = True, _whitening=None):
        super().__init__()
        if isinstance(z_matrix, torch.Tensor):
            z_matrix = z_matrix.cpu().numpy()
        if isinstance(fixed_atoms, torch.Tensor):
            fixed_atoms = fixed_atoms.cpu().numpy()
        if not enforce_boundaries:
            raise NotImplementedError("the IC kernels always enforce the eps boundaries")
        self._plan = engine.RelPlan(z_matrix, fixed_atoms, normalize_angles=normalize_angles, eps=eps,
                                    whitening=_whitening)
        self._raise_warnings = raise_warnings

    z_matrix = property(lambda self: self._plan.rel)
    fixed_atoms = property(lambda self: self._plan.fixed)
    dim_bonds = property(lambda self: len(self._plan.rel))
    dim_angles = property(lambda self: len(self._plan.rel))
    dim_torsions = property(lambda self: len(self._plan.rel))
    dim_fixed = property(lambda self: 3 * len(self._plan.fixed))
    bond_indices = property(lambda self: self._plan.rel[:, :2])
    angle_indices = property(lambda self: self._plan.rel[:, :3])
    torsion_indices = property(lambda self: self._plan.rel[:, :4])
    normalize_angles = property(lambda self: self._plan.normalize_angles)

    def _forward(self, x, *args, **kwargs):
        x = x.reshape(x.shape[0], -1)
        if torch.is_grad_enabled() and x.requires_grad:
            return _ag.relic_from_xyz_with_grad(self._plan, engine.relic_from_xyz, x)
        return engine.relic_from_xyz(self._plan, x)

    def _inverse(self, bonds, angles, torsions, x_fixed, *args, **kwargs):
        x_fixed = x_fixed.reshape(x_fixed.shape[0], -1)
        ins = (bonds, angles, torsions, x_fixed)
        if torch.is_grad_enabled() and any(t.requires_grad for t in ins):
            return _ag.relic_to_xyz_with_grad(self._plan, engine.relic_to_xyz, *ins)
        return engine.relic_to_xyz(self._plan, *ins)


def _pca(x0, keepdims):
    """crd_transform/pca.py:10-34 (numpy, in the dtype of the data)."""
    mean = x0.mean(axis=0)
    xm = x0 - mean
    cov = np.matmul(xm.T, xm) / (xm.shape[0] - 1.0)
    eigval, eigvec = np.linalg.eigh(cov)
    idx = np.argsort(eigval)[::-1][:keepdims]
    std = np.sqrt(eigval[idx])
    eigvec = eigvec[:, idx]
    return mean, np.matmul(eigvec, np.diag(1.0 / std)), np.matmul(np.diag(std), eigvec.T), std


class WhitenFlow(Flow):
    """Static PCA whitening (crd_transform/pca.py:37-107).  A plain ``[B, d] x [d, k]`` library GEMM
    when used on its own; inside ``MixedCoordinateTransformation`` it is part of the IC kernels."""

    def __init__(self, X0, keepdims=None, whiten_inverse=True):
        super().__init__()
        if keepdims is None:
            keepdims = X0.shape[1]
        self.dim = X0.shape[1]
        self.keepdims = keepdims
        self.whiten_inverse = whiten_inverse
        mean, tw, tb, std = _pca(X0.detach().cpu().numpy(), keepdims)
        self.register_buffer("X0mean", torch.tensor(mean).to(X0))
        self.register_buffer("Twhiten", torch.tensor(tw).to(X0))
        self.register_buffer("Tblacken", torch.tensor(tb).to(X0))
        self.register_buffer("std", torch.tensor(std).to(X0))
        if torch.any(self.std <= 0):
            raise ValueError("Cannot construct whiten layer because trying to keep nonpositive eigenvalues.")
        self.jacobian_xz = -torch.sum(torch.log(self.std))

    def _whiten(self, x):
        return torch.matmul(x - self.X0mean, self.Twhiten), self.jacobian_xz * torch.ones((x.shape[0], 1)).to(x)

    def _blacken(self, x):
        return torch.matmul(x, self.Tblacken) + self.X0mean, -self.jacobian_xz * torch.ones((x.shape[0], 1)).to(x)

    def _forward(self, x, *args, **kwargs):
        return self._blacken(x) if self.whiten_inverse else self._whiten(x)

    def _inverse(self, x, *args, **kwargs):
        return self._whiten(x) if self.whiten_inverse else self._blacken(x)


class MixedCoordinateTransformation(RelativeInternalCoordinateTransformation):
    """Relative internal coordinates + PCA-whitened fixed block (ic.py:719-884): ``_forward`` maps
    xyz -> (bonds, angles, torsions, z_fixed ``[B, keepdims]``, dlogp).  The whitening GEMV runs inside
    the IC kernels; the PCA itself is host work done once here, in the dtype of ``data`` like the
    reference (pca.py:61-62)."""

    def __init__(self, data, z_matrix, fixed_atoms, keepdims=None, normalize_angles=True, eps: float = 1e-7,
                 enforce_boundaries: bool = True, raise_warnings: bool = True):
        if isinstance(fixed_atoms, torch.Tensor):
            fixed_atoms = fixed_atoms.cpu().numpy()
        fixed_atoms = np.asarray(fixed_atoms)
        n_data = data.shape[0]
        fixed = data.reshape(n_data, -1, 3)[:, fixed_atoms].reshape(n_data, -1)
        self_whiten = WhitenFlow(fixed, keepdims=keepdims, whiten_inverse=False)
        w = dict(mean=self_whiten.X0mean.cpu().numpy(), whiten=self_whiten.Twhiten.cpu().numpy(),
                 blacken=self_whiten.Tblacken.cpu().numpy(), jacobian_xz=float(self_whiten.jacobian_xz),
                 keepdims=int(self_whiten.keepdims))
        super().__init__(z_matrix, fixed_atoms, normalize_angles=normalize_angles, eps=eps,
                         enforce_boundaries=enforce_boundaries, raise_warnings=raise_warnings, _whitening=w)
        self._whiten = self_whiten

    @classmethod
    def from_whitening(cls, whiten, z_matrix, fixed_atoms, **kwargs):
        """Build the layer around an existing ``WhitenFlow`` (this package's or the reference's: anything
        with ``X0mean / Twhiten / Tblacken / jacobian_xz / keepdims``) instead of redoing the PCA."""
        w = dict(mean=whiten.X0mean.detach().cpu().numpy(), whiten=whiten.Twhiten.detach().cpu().numpy(),
                 blacken=whiten.Tblacken.detach().cpu().numpy(), jacobian_xz=float(whiten.jacobian_xz),
                 keepdims=int(whiten.keepdims))
        self = cls.__new__(cls)
        RelativeInternalCoordinateTransformation.__init__(self, z_matrix, fixed_atoms, _whitening=w, **kwargs)
        self._whiten = whiten
        return self

    dim_fixed = property(lambda self: self._whiten.keepdims)
