"""Transformers of coupling layers, backed by the fused sm_100a kernels.

Mirrors bgflow/nn/flow/transformer/base.py:7-16 (Transformer), affine.py:10-70
(AffineTransformer) and spline.py:14-204 (ConditionalSplineTransformer): same constructor
arguments, same ``_forward(x_cond, y) -> (y', dlogp)`` contract, same exceptions.  The
arithmetic (conditioner MLP, transform, log-det reduction) runs in ONE kernel per call
(``engine.affine_coupling`` / ``engine.spline_coupling``); nothing is evaluated op by op.
"""

import warnings

import numpy as np
import torch

from . import engine
from .autograd import fused_coupling_with_grad, needs_grad
from .flows import Flow
from .nets import DenseNet, MeanFreeDenseNet, WrapPeriodic

__all__ = ["Transformer", "AffineTransformer", "ConditionalSplineTransformer"]

DEFAULT_MIN_BIN_WIDTH = 1e-3
DEFAULT_MIN_BIN_HEIGHT = 1e-3
DEFAULT_MIN_DERIVATIVE = 1e-3


class Transformer(Flow):
    def __init__(self):
        super().__init__()

    def _forward(self, x, y, *args, **kwargs):
        raise NotImplementedError()

    def _inverse(self, x, y, *args, **kwargs):
        raise NotImplementedError()


def _net_spec(net, cond_width):
    """Describe a conditioner module to the kernels: (weights, biases, act_code, periodic) or
    None when the module is not a (WrapPeriodic-wrapped) DenseNet the kernels understand."""
    periodic = None
    if isinstance(net, WrapPeriodic):
        idx = net.periodic_indices(cond_width)
        periodic = (tuple(int(i) for i in idx), float(net.left), float(net.right), int(cond_width))
        net = net.net
    if type(net) is not DenseNet and not (isinstance(net, DenseNet) and type(net).forward is DenseNet.forward):
        return None
    if isinstance(net, MeanFreeDenseNet):
        return None
    mods = list(net._layers)
    weights, biases, acts = [], [], []
    for m in mods:
        if isinstance(m, torch.nn.Linear):
            weights.append(m.weight)
            biases.append(m.bias)
            acts.append(None)
        else:
            if not weights or acts[-1] is not None:
                return None
            acts[-1] = m
    if not weights or any(b is None for b in biases) or acts[-1] is not None:
        return None
    codes = {engine.activation_code(a) for a in acts[:-1]}
    if len(weights) > 1:
        if len(codes) != 1 or None in codes:
            return None
        code = codes.pop()
    else:
        code = 0
    return weights, biases, code, periodic


class _FusedMixin:
    """Shared parameter-packing logic."""

    def _packed(self, name, net, cond_width, spline=None):
        spec = _net_spec(net, cond_width)
        if spec is None:
            raise NotImplementedError(
                f"{type(self).__name__}: conditioner {type(net).__name__} is not a DenseNet / "
                "WrapPeriodic(DenseNet) with ReLU/SiLU/Tanh/no activation; the fused kernels cannot run it")
        weights, biases, code, periodic = spec
        cache = self.__dict__.setdefault("_pack_cache", {})
        pk = cache.get(name)
        if pk is None:
            pk = cache[name] = engine.PackedNet()
        return pk.refresh(weights, biases, code, periodic=periodic, spline=spline)


class AffineTransformer(_FusedMixin, Transformer):
    """RealNVP / NICE transformer (affine.py:10-70).

    ``y' = y * exp(log_sigma) + mu`` with ``mu = shift(x)``,
    ``log_sigma = tanh(scale(x)) * exp(log_alpha)``; ``dlogp = sum(log_sigma)``.
    """

    def __init__(self, shift_transformation=None, scale_transformation=None, init_downscale=1.0,
                 preserve_volume=False, is_circular=False):
        if scale_transformation is not None and is_circular:
            raise ValueError("Scaling is not compatible with periodicity.")
        super().__init__()
        self._shift_transformation = shift_transformation
        self._scale_transformation = scale_transformation
        self._log_alpha = torch.nn.Parameter(torch.zeros(1) - init_downscale)
        self._preserve_volume = preserve_volume
        self._is_circular = is_circular

    def _coupling(self, cond, tr, inverse=False, dlogp_acc=None, **kwargs):
        engine.require_cuda_fp32(*cond, *tr)
        if needs_grad([*cond, *tr], self):
            outs, dlogp = fused_coupling_with_grad(self, "affine", list(cond), list(tr), inverse)
            return outs, (dlogp if dlogp_acc is None else dlogp_acc + dlogp)
        return self._launch(cond, tr, inverse=inverse, dlogp_acc=dlogp_acc)

    def _launch(self, cond, tr, inverse=False, dlogp_acc=None):
        d_c = sum(t.shape[-1] for t in cond)
        d_t = sum(t.shape[-1] for t in tr)
        shift = scale = None
        if self._shift_transformation is not None:
            shift = self._packed("shift", self._shift_transformation, d_c)
            assert shift.N[shift.n_layers - 1] == d_t
        if self._scale_transformation is not None:
            scale = self._packed("scale", self._scale_transformation, d_c)
            assert scale.N[scale.n_layers - 1] == d_t
        return engine.affine_coupling(
            cond, tr, shift, scale, self._log_alpha_host(), inverse=inverse,
            preserve_volume=self._preserve_volume, is_circular=self._is_circular, dlogp_in=dlogp_acc)

    def _log_alpha_host(self):
        # the ABI takes log_alpha by value: read it back once per parameter version, not per call
        key = (self._log_alpha.data_ptr(), self._log_alpha._version)
        cached = self.__dict__.get("_log_alpha_cache")
        if cached is None or cached[0] != key:
            cached = (key, float(self._log_alpha.detach()))
            self.__dict__["_log_alpha_cache"] = cached
        return cached[1]

    def _forward(self, x, y, *cond, **kwargs):
        (out,), dlogp = self._coupling([x], [y], inverse=False)
        return out, dlogp

    def _inverse(self, x, y, *cond, **kwargs):
        (out,), dlogp = self._coupling([x], [y], inverse=True)
        return out, dlogp


class ConditionalSplineTransformer(_FusedMixin, Transformer):
    """Rational-quadratic spline transformer on ``[left, right) -> [bottom, top)``
    (spline.py:14-204; Durkan et al. 2019).  The number of bins is inferred from the width of
    ``params_net``'s output: ``3 * n_bins * y_dim + n_noncircular``.  As in the reference,
    ``_forward`` is the quadratic-root branch (sampling direction) and ``_inverse`` the direct
    evaluation; the nflows settings are ``min_bin_width = min_bin_height = min_derivative = 1e-3``
    and ``enable_identity_init=True``.
    """

    def __init__(self, params_net, is_circular=False, left=0.0, right=1.0, bottom=0.0, top=1.0):
        super().__init__()
        self._params_net = params_net
        self._is_circular = torch.as_tensor(is_circular, dtype=torch.bool)
        self._left = left
        self._right = right
        self._bottom = bottom
        self._top = top
        self._default_settings = {
            "min_bin_width": DEFAULT_MIN_BIN_WIDTH,
            "min_bin_height": DEFAULT_MIN_BIN_HEIGHT,
            "min_derivative": DEFAULT_MIN_DERIVATIVE,
            "enable_identity_init": True,
        }
        self._oob = None
        self._oob_seen = 0

    # -- circular bookkeeping (spline.py:190-204, with the intended non-circular count)
    def _circular_mask(self, y_dim):
        c = self._is_circular
        if c.dim() == 0:
            return [bool(c)] * y_dim
        if c.numel() != y_dim:
            raise RuntimeError(f"is_circular has {c.numel()} entries for {y_dim} transformed dims")
        return [bool(v) for v in c.tolist()]

    def _n_noncircular(self, y_dim):
        return y_dim - sum(self._circular_mask(y_dim))

    def _end_slope_cols(self, d_t, n_bins, device):
        """int32 device tensor: for every transformed dim the column of the conditioner output that
        holds its last slope (its own first slope if circular, else its S_last entry; spline.py:123-125)."""
        key = (d_t, n_bins, str(device))
        cache = self.__dict__.setdefault("_end_cols", {})
        if key not in cache:
            cols, nc = [], 0
            for d, circ in enumerate(self._circular_mask(d_t)):
                if circ:
                    cols.append(2 * n_bins * d_t + d * n_bins)
                else:
                    cols.append(3 * n_bins * d_t + nc)
                    nc += 1
            cache[key] = torch.tensor(cols, dtype=torch.int32, device=device)
        return cache[key]

    def _net_out_width(self):
        net = self._params_net.net if isinstance(self._params_net, WrapPeriodic) else self._params_net
        last = [m for m in net._layers if isinstance(m, torch.nn.Linear)][-1]
        return last.out_features

    def _coupling(self, cond, tr, inverse=False, dlogp_acc=None, **kwargs):
        engine.require_cuda_fp32(*cond, *tr)
        if needs_grad([*cond, *tr], self):
            outs, dlogp = fused_coupling_with_grad(self, "spline", list(cond), list(tr), inverse)
            return outs, (dlogp if dlogp_acc is None else dlogp_acc + dlogp)
        return self._launch(cond, tr, inverse=inverse, dlogp_acc=dlogp_acc)

    def _launch(self, cond, tr, inverse=False, dlogp_acc=None):
        d_c = sum(t.shape[-1] for t in cond)
        d_t = sum(t.shape[-1] for t in tr)
        mask = self._circular_mask(d_t)
        n_bins = self._net_out_width() // (3 * d_t)           # spline.py:112
        net = self._packed("params", self._params_net, d_c, spline=(d_t, n_bins, tuple(mask)))
        dev = tr[0].device
        if self._oob is None or self._oob.device != dev:
            self._oob = torch.zeros(1, dtype=torch.int32, device=dev)
            self._oob_seen = 0
        s = self._default_settings
        return engine.spline_coupling(
            cond, tr, net, n_bins, inverse=inverse, left=self._left, right=self._right,
            bottom=self._bottom, top=self._top, min_bin_width=s["min_bin_width"],
            min_bin_height=s["min_bin_height"], min_derivative=s["min_derivative"],
            identity_init=s["enable_identity_init"], oob_counter=self._oob, dlogp_in=dlogp_acc)

    def out_of_domain_count(self, warn=True):
        """Number of inputs clamped into ``[left, right]`` so far (synchronises).  The
        reference warns eagerly with a device->host sync per call (spline.py:145-155); here
        the kernel counts and the warning is raised lazily when this is queried."""
        if self._oob is None:
            return 0
        n = int(self._oob.item())
        if warn and n > self._oob_seen:
            warnings.warn(f"InputOutsideDomain: {n - self._oob_seen} inputs were clamped to "
                          f"[{self._left}, {self._right}]", UserWarning)
        self._oob_seen = n
        return n

    def _forward(self, x, y, *args, **kwargs):
        (out,), dlogp = self._coupling([x], [y], inverse=False)
        return out, dlogp

    def _inverse(self, x, y, *args, **kwargs):
        (out,), dlogp = self._coupling([x], [y], inverse=True)
        return out, dlogp
