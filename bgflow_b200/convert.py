"""Switch an existing *reference* object graph (``bgflow.SequentialFlow``, ``bgflow.BoltzmannGenerator``,
...) over to this package: ``from_reference(obj)`` rebuilds the flow from the mirror classes of
this package while SHARING the reference's parameter tensors (the ``torch.nn.Linear`` modules inside
every ``DenseNet`` are reused, ``_log_alpha`` and learnable marginals too), so optimisers,
``state_dict`` keys and checkpoints keep working and both object graphs stay in sync.

The reference is recognised by class name and attributes, never imported: nothing here needs
``bgflow`` on the path.  What is recognised (reference file:line → mirror):

  SequentialFlow sequential.py:10-92, InverseFlow inverted.py:7-23, SplitFlow / MergeFlow / SwapFlow /
  CouplingFlow / WrapFlow / SetConstantFlow coupling.py:13-272, AffineTransformer affine.py:10-70,
  ConditionalSplineTransformer spline.py:14-204, DenseNet dense.py:9-48, WrapPeriodic
  periodic.py:7-37, CDFTransform cdf.py:13-46, Global / Relative / Mixed internal coordinates
  crd_transform/ic.py:268-884, BoltzmannGenerator bg.py:77-165.

Everything else (other flows, conditioners the kernels do not understand, priors, targets) is kept
as the reference object: it obeys the same ``forward(*xs, inverse=...)`` protocol, so the mirror's
``SequentialFlow`` simply calls it — the reference's "falls through to the reference
implementation" (SURVEY.md 8b).  ``strict=True`` raises instead.  The builder tail is fused
(``fuse_domain_maps``) unless ``fuse_tail=False``.
"""

import numpy as np
import torch

from . import cdf as _cdf
from . import flows as _flows
from . import ic as _ic
from . import nets as _nets
from . import transformers as _tr
from .bg import BoltzmannGenerator

__all__ = ["from_reference"]


def _name(obj):
    return type(obj).__name__


def _is_mirror(obj):
    return type(obj).__module__.startswith("bgflow_b200")


def _convert_net(net, strict):
    """Conditioner: reference DenseNet / WrapPeriodic(DenseNet) -> mirror classes sharing the layers."""
    if net is None or _is_mirror(net):
        return net
    n = _name(net)
    if n == "WrapPeriodic" and hasattr(net, "net") and hasattr(net, "indices"):
        return _nets.WrapPeriodic(_convert_net(net.net, strict), left=net.left, right=net.right, indices=net.indices)
    if n == "DenseNet" and isinstance(getattr(net, "_layers", None), torch.nn.Sequential):
        out = _nets.DenseNet.__new__(_nets.DenseNet)
        torch.nn.Module.__init__(out)
        out._layers = net._layers                   # the very same Linear / activation modules
        return out
    if strict:
        raise NotImplementedError(f"conditioner {n} has no kernel-backed mirror")
    return net          # the fused transformer will refuse it; CouplingFlow then needs the reference transformer


def _convert_transformer(t, strict):
    if _is_mirror(t):
        return t
    n = _name(t)
    if n == "AffineTransformer" and hasattr(t, "_log_alpha"):
        shift = _convert_net(t._shift_transformation, strict)
        scale = _convert_net(t._scale_transformation, strict)
        if all(x is None or _is_mirror(x) for x in (shift, scale)):
            out = _tr.AffineTransformer(shift, scale, preserve_volume=t._preserve_volume, is_circular=t._is_circular)
            out._log_alpha = t._log_alpha                   # shared parameter
            return out
    if n == "ConditionalSplineTransformer" and hasattr(t, "_params_net"):
        net = _convert_net(t._params_net, strict)
        if _is_mirror(net):
            circ = t._is_circular
            circ = bool(circ) if getattr(circ, "ndim", 0) == 0 else circ
            out = _tr.ConditionalSplineTransformer(net, is_circular=circ, left=t._left, right=t._right,
                                                   bottom=t._bottom, top=t._top)
            settings = getattr(t, "_default_settings", None)
            if settings is not None and hasattr(out, "_default_settings"):
                out._default_settings.update({k: v for k, v in settings.items() if k in out._default_settings})
            return out
    if strict:
        raise NotImplementedError(f"transformer {n} has no kernel-backed mirror")
    return t


def _global_z_matrix(ic):
    """Full z-matrix of a reference GlobalInternalCoordinateTransformation (ic.py:606-631 keeps the
    three seed atoms in ``_rel_ic.fixed_atoms`` and the remaining rows in ``_rel_ic.z_matrix``)."""
    rel = ic._rel_ic
    s0, s1, s2 = (int(a) for a in np.asarray(rel.fixed_atoms))
    seeds = np.array([[s0, -1, -1, -1], [s1, s0, -1, -1], [s2, s1, s0, -1]])
    return np.vstack([seeds, np.asarray(rel.z_matrix)])


def _convert_flow(f, strict):
    if _is_mirror(f) or not isinstance(f, torch.nn.Module):
        return f
    n = _name(f)
    if n == "SequentialFlow" and hasattr(f, "_blocks"):
        return _flows.SequentialFlow([_convert_flow(b, strict) for b in f._blocks])
    if n == "MergeFlow" and hasattr(f, "_delegate"):
        d = f._delegate
        if d._indices is None:
            return _flows.MergeFlow(*d._sizes, dim=d._split_dim)
        return _flows.InverseFlow(_convert_flow(d, strict))
    if n == "InverseFlow" and hasattr(f, "_delegate"):
        return _flows.InverseFlow(_convert_flow(f._delegate, strict))
    if n == "SplitFlow" and hasattr(f, "_split_dim"):
        args = f._sizes if f._indices is None else f._indices
        return _flows.SplitFlow(*args, dim=f._split_dim)
    if n == "SwapFlow":
        return _flows.SwapFlow()
    if n == "CouplingFlow" and hasattr(f, "transformer"):
        return _flows.CouplingFlow(_convert_transformer(f.transformer, strict), transformed_indices=f.transformed_indices,
                                   cond_indices=f.cond_indices, cat_dim=f.cat_dim)
    if n == "WrapFlow" and hasattr(f, "_flow"):
        return _flows.WrapFlow(_convert_flow(f._flow, strict), f._indices, f._out_indices)
    if n == "SetConstantFlow" and hasattr(f, "values"):
        return _flows.SetConstantFlow(list(f.indices), list(f.values), n_event_dims0=f.n_event_dims0)
    if n == "CDFTransform" and hasattr(f, "distribution"):
        return _cdf.CDFTransform(f.distribution, eps=f._eps)     # marginals are duck-typed: shared as they are
    if n == "GlobalInternalCoordinateTransformation" and hasattr(f, "_rel_ic"):
        rel = f._rel_ic
        return _ic.GlobalInternalCoordinateTransformation(_global_z_matrix(f), normalize_angles=rel._normalize_angles,
                                                          eps=rel._eps, enforce_boundaries=rel._enforce_boundaries,
                                                          raise_warnings=rel._raise_warnings)
    if n == "RelativeInternalCoordinateTransformation" and hasattr(f, "_z_matrix"):
        return _ic.RelativeInternalCoordinateTransformation(f._z_matrix, f._fixed_atoms, normalize_angles=f._normalize_angles,
                                                            eps=f._eps, enforce_boundaries=f._enforce_boundaries,
                                                            raise_warnings=f._raise_warnings)
    if n == "MixedCoordinateTransformation" and hasattr(f, "_whiten") and hasattr(f, "_rel_ic"):
        rel = f._rel_ic
        return _ic.MixedCoordinateTransformation.from_whitening(          # the PCA is taken over, not redone
            f._whiten, rel._z_matrix, rel._fixed_atoms, normalize_angles=rel._normalize_angles, eps=rel._eps,
            enforce_boundaries=rel._enforce_boundaries, raise_warnings=rel._raise_warnings)
    if strict:
        raise NotImplementedError(f"flow {n} has no mirror in bgflow_b200")
    return f


def from_reference(obj, strict=False, fuse_tail=True):
    """Return the kernel-backed equivalent of a reference flow / transformer / conditioner /
    ``BoltzmannGenerator`` (parameters shared, see the module docstring)."""
    if _name(obj) == "BoltzmannGenerator" and hasattr(obj, "_flow") and not _is_mirror(obj):
        return BoltzmannGenerator(obj._prior, from_reference(obj._flow, strict=strict, fuse_tail=fuse_tail), obj._target)
    if _name(obj) in ("AffineTransformer", "ConditionalSplineTransformer"):
        return _convert_transformer(obj, strict)
    if _name(obj) in ("DenseNet", "WrapPeriodic"):
        return _convert_net(obj, strict)
    out = _convert_flow(obj, strict)
    if fuse_tail and isinstance(out, _flows.SequentialFlow):
        out = _cdf.fuse_domain_maps(out)
    return out
