"""Host-side mirror of bgflow's flow plumbing (same names, arguments and error behaviour).

Reference: bgflow/nn/flow/base.py:7-33 (Flow), sequential.py:10-92 (SequentialFlow),
inverted.py:7-23 (InverseFlow), coupling.py:13-272 (SplitFlow, MergeFlow, SwapFlow,
CouplingFlow, WrapFlow, SetConstantFlow).  Written from scratch: the tuple-of-tensors
contract is the same, but zero-Jacobian blocks do not allocate ``[B,1]`` zero tensors inside a
SequentialFlow, CouplingFlow hands the kernels a list of column segments instead of
``torch.cat`` copies, and the running ``dlogp`` is accumulated inside the fused kernels.
"""

import warnings
from collections.abc import Sequence

import numpy as np
import torch

__all__ = ["Flow", "SequentialFlow", "InverseFlow", "SplitFlow", "MergeFlow", "SwapFlow",
           "CouplingFlow", "WrapFlow", "SetConstantFlow"]


def _zero_dlogp(x):
    return torch.zeros(*x.shape[:-1], 1, dtype=x.dtype, device=x.device)


class Flow(torch.nn.Module):
    """``forward(*xs, inverse=False, **kwargs) -> (*ys, dlogp[..., 1])`` (base.py:17-33)."""

    #: blocks whose Jacobian is the identity advertise it, so SequentialFlow can skip the zeros
    _volume_preserving = False
    #: blocks that can add their log-det onto a running accumulator inside their kernel
    _accumulates_dlogp = False

    def __init__(self):
        super().__init__()

    def _forward(self, *xs, **kwargs):
        raise NotImplementedError()

    def _inverse(self, *xs, **kwargs):
        raise NotImplementedError()

    def forward(self, *xs, inverse=False, **kwargs):
        if inverse:
            return self._inverse(*xs, **kwargs)
        return self._forward(*xs, **kwargs)


class SequentialFlow(Flow):
    """Discrete stack of blocks (sequential.py:10-92); reversed order for ``inverse=True``."""

    def __init__(self, blocks):
        super().__init__()
        self._blocks = torch.nn.ModuleList(blocks)

    def forward(self, *xs, inverse=False, **kwargs):
        dlogp = None
        blocks = reversed(self._blocks) if inverse else self._blocks
        for block in blocks:
            if getattr(block, "_volume_preserving", False) and hasattr(block, "_apply_tuple"):
                xs = block._apply_tuple(xs, inverse)
                continue
            if getattr(block, "_accumulates_dlogp", False):
                *xs, dlogp = block(*xs, inverse=inverse, _dlogp_acc=dlogp, **kwargs)
                continue
            *xs, ddlogp = block(*xs, inverse=inverse, **kwargs)
            dlogp = ddlogp if dlogp is None else dlogp + ddlogp
        if dlogp is None:
            dlogp = _zero_dlogp(xs[0]) if len(self._blocks) else 0.0
        return (*xs, dlogp)

    def _forward(self, *args, **kwargs):
        return self.forward(*args, **kwargs, inverse=False)

    def _inverse(self, *args, **kwargs):
        return self.forward(*args, **kwargs, inverse=True)

    def trigger(self, function_name):
        results = [getattr(block, function_name)() for block in self._blocks
                   if hasattr(block, function_name) and callable(getattr(block, function_name))]
        if len(results) > 0 and all(res is not None for res in results):
            return torch.stack(results)
        return torch.zeros(0)

    def __iter__(self):
        return iter(self._blocks)

    def __getitem__(self, index):
        if isinstance(index, int):
            return self._blocks[index]
        indices = np.arange(len(self))[index]
        return SequentialFlow([self._blocks[i] for i in indices])

    def __len__(self):
        return len(self._blocks)


class InverseFlow(Flow):
    """Swap ``_forward`` / ``_inverse`` of a delegate (inverted.py:7-23)."""

    def __init__(self, delegate):
        super().__init__()
        self._delegate = delegate
        self._volume_preserving = getattr(delegate, "_volume_preserving", False) and hasattr(
            delegate, "_apply_tuple")
        self._accumulates_dlogp = getattr(delegate, "_accumulates_dlogp", False)

    def _forward(self, *xs, **kwargs):
        return self._delegate._inverse(*xs, **kwargs)

    def _inverse(self, *xs, **kwargs):
        return self._delegate._forward(*xs, **kwargs)

    def _apply_tuple(self, xs, inverse):
        return self._delegate._apply_tuple(xs, not inverse)


class SplitFlow(Flow):
    """Split one tensor into several along ``dim`` (coupling.py:13-104).

    ``sizes_or_indices``: ints = lengths (the last may be omitted and is inferred), or
    sequences of ints = the indices each output takes.
    """

    _volume_preserving = True

    def __init__(self, *sizes_or_indices, dim=-1):
        super().__init__()
        first = sizes_or_indices[0]
        if isinstance(first, Sequence) or isinstance(first, np.ndarray):
            self._sizes = None
            self._indices = [list(int(i) for i in idx) for idx in sizes_or_indices]
        else:
            self._sizes = tuple(int(s) for s in sizes_or_indices)
            self._indices = None
        self._split_dim = dim

    # -- tuple -> tuple, no dlogp
    def _apply_tuple(self, xs, inverse):
        if not inverse:
            (x,) = xs
            if self._kernel_ok(x):
                # dense halves let the coupling kernels move whole [128 x D] tiles with one bulk TMA
                # copy each (strided views would fall back to element-wise staging): one launch
                from . import engine
                n = x.shape[-1]
                rest = n - sum(self._sizes)
                if rest < 0:
                    raise ValueError(f"can't split x [{x.shape}] into sizes {self._sizes} along {self._split_dim}")
                return tuple(engine.split_cols(x, list(self._sizes) + ([rest] if rest > 0 else [])))
            parts = tuple(self._split(x))
            if x.is_cuda:
                parts = tuple(p.contiguous() for p in parts)
            return parts
        if all(self._kernel_ok(x) for x in xs):
            from . import engine
            return (engine.merge_cols(list(xs)),)
        return (self._merge(*xs),)

    def _kernel_ok(self, x):
        return (self._indices is None and x.is_cuda and x.dtype == torch.float32 and x.dim() >= 1
                and self._split_dim in (-1, x.dim() - 1) and len(self._sizes) < 8
                and not (torch.is_grad_enabled() and x.requires_grad))

    def _split(self, x):
        n = x.shape[self._split_dim]
        if self._indices is None:
            rest = n - sum(self._sizes)
            if rest < 0:
                raise ValueError(f"can't split x [{x.shape}] into sizes {self._sizes} along {self._split_dim}")
            sizes = list(self._sizes) + ([rest] if rest > 0 else [])
            return torch.split(x, sizes, dim=self._split_dim)
        self._check_cover(n, "split")
        return [x.index_select(self._split_dim, torch.as_tensor(idx, device=x.device)) for idx in self._indices]

    def _merge(self, *xs):
        if self._indices is None:
            return torch.cat(xs, dim=self._split_dim)
        n = sum(len(idx) for idx in self._indices)
        self._check_cover(n, "merge")
        shape = list(xs[0].shape)
        shape[self._split_dim] = n
        y = torch.empty(*shape, device=xs[0].device, dtype=xs[0].dtype)
        for x, idx in zip(xs, self._indices):
            y.index_copy_(self._split_dim % y.dim(), torch.as_tensor(idx, device=x.device), x)
        return y

    def _check_cover(self, n, what):
        seen = np.zeros(n, dtype=bool)
        for idx in self._indices:
            if seen[idx].any():
                raise ValueError(f"Cannot {what} tensor. Indices are overlapping.")
            seen[idx] = True
        if not seen.all():
            raise ValueError(f"{what} with indices missed indices {np.flatnonzero(~seen)}")

    def _forward(self, x, **kwargs):
        return (*self._split(x), self._dlogp(x))

    def _inverse(self, *xs, **kwargs):
        return self._merge(*xs), self._dlogp(xs[0])

    def _dlogp(self, x):
        shape = list(x.shape)
        shape[self._split_dim] = 1
        return torch.zeros(shape, dtype=x.dtype, device=x.device)


class MergeFlow(InverseFlow):
    """Shortcut for ``InverseFlow(SplitFlow(...))`` (coupling.py:107-110)."""

    def __init__(self, *sizes, dim=-1):
        super().__init__(SplitFlow(*sizes, dim=dim))


class SwapFlow(Flow):
    """Swap the first two tensors (coupling.py:113-130)."""

    _volume_preserving = True

    def __init__(self):
        super().__init__()

    def _apply_tuple(self, xs, inverse):
        if len(xs) == 1:
            warnings.warn("applying swapping on a single tensor has no effect")
            return tuple(xs)
        return (xs[1], xs[0], *xs[2:])

    def _forward(self, *xs, **kwargs):
        return (*self._apply_tuple(xs, False), _zero_dlogp(xs[0]))

    def _inverse(self, *xs, **kwargs):
        return (*self._apply_tuple(xs, True), _zero_dlogp(xs[0]))


class CouplingFlow(Flow):
    """Coupling layer (coupling.py:133-182).

    ``transformer(cond, inputs)`` transforms the tensors at ``transformed_indices`` conditioned
    on those at ``cond_indices``.  With the package's own transformers the concatenations of
    the reference (coupling.py:163-165) never materialise: the kernel reads the tensors as
    column segments and writes one output buffer that is returned as views.
    """

    _accumulates_dlogp = True

    def __init__(self, transformer, transformed_indices=(1,), cond_indices=(0,), cat_dim=-1):
        super().__init__()
        self.transformer = transformer
        self.transformed_indices = transformed_indices
        self.cond_indices = cond_indices
        invalid = np.intersect1d(self.transformed_indices, self.cond_indices)
        if len(invalid) > 0:
            raise ValueError(f"Indices {invalid} cannot be both transformed and conditioned on.")
        self.cat_dim = cat_dim

    def _run(self, x, inverse, _dlogp_acc=None, **kwargs):
        x = list(x)
        tr = [x[i] for i in self.transformed_indices]
        cond = [x[i] for i in self.cond_indices]
        fused = getattr(self.transformer, "_coupling", None)
        last_dim = all(self.cat_dim in (-1, t.dim() - 1) for t in tr + cond)
        if fused is not None and last_dim:
            ys, dlogp = fused(cond, tr, inverse=inverse, dlogp_acc=_dlogp_acc, **kwargs)
        else:
            # generic transformer: the reference's own cat / split plumbing
            lengths = [t.shape[self.cat_dim] for t in tr]
            inputs = torch.cat(tr, dim=self.cat_dim)
            cond_inputs = torch.cat(cond, dim=self.cat_dim)
            if inverse:
                y, dlogp = self.transformer.forward(cond_inputs, inputs, **kwargs, inverse=True)
            else:
                y, dlogp = self.transformer.forward(cond_inputs, inputs, **kwargs)
            ys = torch.split(y, lengths, self.cat_dim)
            if _dlogp_acc is not None:
                dlogp = _dlogp_acc + dlogp
        for i, yi in zip(self.transformed_indices, ys):
            x[i] = yi
        return (*x, dlogp)

    def _forward(self, *x, **kwargs):
        return self._run(x, False, **kwargs)

    def _inverse(self, *x, **kwargs):
        return self._run(x, True, **kwargs)


class WrapFlow(Flow):
    """Apply ``flow`` to the tensors at ``indices``; its outputs land at ``out_indices``
    (default: the same places).  Reference behaviour: coupling.py:185-222."""

    def __init__(self, flow, indices, out_indices=None):
        super().__init__()
        self._flow = flow
        self._indices = indices
        self._out_indices = indices if out_indices is None else out_indices
        # a wrapped kernel-backed flow still adds its log-det onto the running dlogp in-kernel
        self._accumulates_dlogp = getattr(flow, "_accumulates_dlogp", False)

    @staticmethod
    def _route(xs, take, put, run):
        taken = set(int(i) for i in take)
        rest = [x for pos, x in enumerate(xs) if pos not in taken]
        *ys, dlogp = run(*(xs[i] for i in take))
        # insert in ascending target position so that earlier inserts do not shift later ones
        for slot in np.argsort(put):
            rest.insert(put[slot], ys[slot])
        return (*rest, dlogp)

    def _forward(self, *xs, **kwargs):
        return self._route(xs, self._indices, self._out_indices,
                           lambda *inp: self._flow(*inp, **kwargs))

    def _inverse(self, *xs, **kwargs):
        return self._route(xs, self._out_indices, self._indices,
                           lambda *inp: self._flow(*inp, inverse=True, **kwargs))


class SetConstantFlow(Flow):
    """Forward: insert constant tensors at ``indices``; inverse: drop them.  Zero log-det.
    Reference behaviour: coupling.py:227-272 (constants are tiled over the batch shape of
    ``xs[0]``, whose first ``n_event_dims0`` dims are taken as batch dims)."""

    _PREFIX = "_values_"

    def __init__(self, indices, values, n_event_dims0=1):
        super().__init__()
        pairs = sorted(zip(indices, values), key=lambda iv: iv[0])
        self.indices = [i for i, _ in pairs]
        for n, (_, v) in enumerate(pairs):
            self.register_buffer(f"{self._PREFIX}{n}", v)
        self.n_event_dims0 = n_event_dims0

    @property
    def values(self):
        return [getattr(self, f"{self._PREFIX}{n}") for n in range(len(self.indices))]

    def _batch_zeros(self, ref):
        shape = list(ref.shape[:self.n_event_dims0])
        return shape, torch.zeros(shape + [1], device=ref.device, dtype=ref.dtype)

    def _forward(self, *xs, **kwargs):
        shape, dlogp = self._batch_zeros(xs[0])
        out = list(xs)
        for i, v in zip(self.indices, self.values):
            out.insert(i, v.repeat(*shape, *([1] * v.dim())))
        return (*out, dlogp)

    def _inverse(self, *xs, **kwargs):
        drop = set(self.indices)
        out = [x for pos, x in enumerate(xs) if pos not in drop]
        _, dlogp = self._batch_zeros(out[0])
        return (*out, dlogp)
