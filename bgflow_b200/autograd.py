"""Autograd bridge for the fused coupling kernels (training support).

Forward: always the fused CUDA kernel.  Backward: recompute-in-backward — the block is
re-evaluated with the differentiable device-side PyTorch definition in ``_torch_math`` and
differentiated by autograd (inputs AND conditioner parameters).  Nothing is saved besides the
block's inputs, so training memory stays ``O(B * D)`` per block instead of the reference's
``O(B * 825)`` parameter tensors; the backward itself is not fused yet.
"""

import torch

from . import _torch_math

__all__ = ["fused_coupling_with_grad", "needs_grad"]


def needs_grad(tensors, module):
    if not torch.is_grad_enabled():
        return False
    return any(t.requires_grad for t in tensors) or any(p.requires_grad for p in module.parameters())


class _FusedCoupling(torch.autograd.Function):
    @staticmethod
    def forward(ctx, transformer, kind, inverse, n_cond, n_tr, *tensors):
        cond = list(tensors[:n_cond])
        tr = list(tensors[n_cond:n_cond + n_tr])
        with torch.no_grad():
            outs, dlogp = transformer._launch(cond, tr, inverse=inverse, dlogp_acc=None)
        ctx.transformer, ctx.kind, ctx.inverse = transformer, kind, inverse
        ctx.n_cond, ctx.n_tr = n_cond, n_tr
        ctx.save_for_backward(*cond, *tr)
        return (*outs, dlogp)

    @staticmethod
    def backward(ctx, *grads):
        saved = ctx.saved_tensors
        t = ctx.transformer
        params = [p for p in t.parameters() if p.requires_grad]
        with torch.enable_grad():
            cond = [x.detach().requires_grad_(True) for x in saved[:ctx.n_cond]]
            tr = [x.detach().requires_grad_(True) for x in saved[ctx.n_cond:]]
            fn = _torch_math.spline if ctx.kind == "spline" else _torch_math.affine
            out, dlogp = fn(t, torch.cat(cond, dim=-1), torch.cat(tr, dim=-1), ctx.inverse)
            outs = torch.split(out, [x.shape[-1] for x in tr], dim=-1)
            gin = torch.autograd.grad([*outs, dlogp], [*cond, *tr, *params],
                                      grad_outputs=[g if g is not None else torch.zeros_like(o)
                                                    for g, o in zip(grads, [*outs, dlogp])],
                                      allow_unused=True)
        g_inputs = gin[:ctx.n_cond + ctx.n_tr]
        g_params = iter(gin[ctx.n_cond + ctx.n_tr:])
        full = [next(g_params) if p.requires_grad else None for p in t.parameters()]
        return (None, None, None, None, None, *g_inputs, *full)


def fused_coupling_with_grad(transformer, kind, cond, tr, inverse):
    """Kernel forward + recompute backward.  Returns (list of outputs, dlogp)."""
    params = list(transformer.parameters())
    res = _FusedCoupling.apply(transformer, kind, inverse, len(cond), len(tr), *cond, *tr, *params)
    return list(res[:-1]), res[-1]
