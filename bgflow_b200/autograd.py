"""Autograd bridge for the fused coupling and internal-coordinate kernels (training support,
SURVEY.md §8b/§8e).

Forward: always the fused CUDA kernel.  Backward: recompute-in-backward — nothing is saved besides
the block's inputs, so training memory stays ``O(B * D)`` per block instead of the reference's
``O(B * 825)`` parameter tensors.  Spline blocks: the conditioner is re-run and differentiated by
torch (dense GEMMs), the spline transform's chain rule is ONE hand-written kernel
(``bgx_spline_backward``).  Affine blocks and the IC layer are re-evaluated with the differentiable
device-side PyTorch definitions in ``_torch_math`` / ``_torch_math_ic``.
"""

import torch

from . import _torch_math, _torch_math_ic

__all__ = ["fused_coupling_with_grad", "needs_grad", "ic_to_xyz_with_grad", "ic_from_xyz_with_grad",
           "relic_to_xyz_with_grad", "relic_from_xyz_with_grad"]


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
            if ctx.kind == "spline":
                gin, g_tr = _spline_backward(t, cond, saved[ctx.n_cond:], params, grads, ctx.inverse)
            elif _affine_explicit_ok(t, params):
                gin, g_tr = _affine_backward(t, cond, saved[ctx.n_cond:], grads, ctx.inverse)
            else:
                from . import engine
                tr = [x.detach().requires_grad_(True) for x in saved[ctx.n_cond:]]
                with _matmul_mode(engine.backward_gemm_mode()):
                    out, dlogp = _torch_math.affine(t, torch.cat(cond, dim=-1), torch.cat(tr, dim=-1), ctx.inverse)
                    outs = torch.split(out, [x.shape[-1] for x in tr], dim=-1)
                    # only outputs that carry a graph and an upstream gradient: for a shift-only (NICE) or
                    # identity transformer dlogp is a constant zero without a grad_fn (affine.py:41-47)
                    pairs = [(o, g) for o, g in zip([*outs, dlogp], grads) if g is not None and o.requires_grad]
                    if pairs:
                        g_all = torch.autograd.grad([o for o, _ in pairs], [*cond, *tr, *params],
                                                    grad_outputs=[g for _, g in pairs], allow_unused=True)
                    else:
                        g_all = (None,) * (ctx.n_cond + ctx.n_tr + len(params))
                g_tr = g_all[ctx.n_cond:ctx.n_cond + ctx.n_tr]
                gin = (*g_all[:ctx.n_cond], *g_all[ctx.n_cond + ctx.n_tr:])
        g_cond = gin[:ctx.n_cond]
        g_params = iter(gin[ctx.n_cond:])
        full = [next(g_params) if p.requires_grad else None for p in t.parameters()]
        return (None, None, None, None, None, *g_cond, *g_tr, *full)


class _matmul_mode:
    """``backward_gemm == "tf32"``: cuBLAS TF32 tensor-core GEMMs inside the block (restored on exit)."""

    def __init__(self, mode):
        self.tf32 = mode == "tf32"

    def __enter__(self):
        if self.tf32:
            self.old = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        if self.tf32:
            torch.backends.cuda.matmul.allow_tf32 = self.old
        return False


def _affine_explicit_ok(t, params):
    """Plain RealNVP block (shift and scale DenseNets, no volume preservation / circular wrap), every parameter
    trainable: the backward below applies; anything else re-evaluates the torch definition under autograd."""
    from . import engine, _mlp_grad
    mode = engine.backward_gemm_mode()
    if mode not in ("bf16x3", "tcgen05") or t._preserve_volume or t._is_circular:
        return False
    sh, sc = t._shift_transformation, t._scale_transformation
    ok = _mlp_grad.tc_supported if mode == "tcgen05" else _mlp_grad.supported
    if sh is None or sc is None or not ok(sh) or not ok(sc):
        return False
    allp = list(t.parameters())
    return len(params) == len(allp) and len(allp) == 1 + len(list(sh.parameters())) + len(list(sc.parameters()))


@torch.no_grad()
def _affine_backward(t, cond, tr, grads, inverse):
    """Backward of ``y' = y exp(ls) + mu`` / ``y' = (y - mu) exp(-ls)`` with ``ls = tanh(s) exp(log_alpha)``
    (affine.py:35-70) written out by hand; both conditioners are re-run and differentiated on tensor cores
    (``_mlp_grad``).  Returns (grads of cond + parameters in ``t.parameters()`` order, grads of tr)."""
    from . import _mlp_grad
    widths = [x.shape[-1] for x in tr]
    y = torch.cat([x.detach() for x in tr], dim=-1) if len(tr) > 1 else tr[0].detach()
    x = torch.cat([c.detach() for c in cond], dim=-1) if len(cond) > 1 else cond[0].detach()
    lead, d_t = y.shape[:-1], y.shape[-1]
    y2, x2 = y.reshape(-1, d_t), x.reshape(-1, x.shape[-1])
    g_out = torch.cat([g if g is not None else torch.zeros_like(v) for g, v in zip(grads[:-1], tr)], dim=-1).reshape(-1, d_t)
    g_dl = grads[-1].reshape(-1, 1) if grads[-1] is not None else None
    from . import engine
    tc = engine.backward_gemm_mode() == "tcgen05"
    if tc:
        # recompute of both conditioners, elementwise chain rule and both conditioner backwards in ONE host call
        tn_mu, _ = _mlp_grad._train_net(t._shift_transformation)
        tn_s, _ = _mlp_grad._train_net(t._scale_transformation)
        d_x2, d_y, g_mu, g_s, d_la = engine.affine_block_backward(
            tn_mu, tn_s, t._log_alpha.detach(), x2, y2, g_out, g_dl.reshape(-1) if g_dl is not None else None, inverse)
        return _affine_pack_grads(t, cond, d_x2.reshape(*lead, x.shape[-1]), g_mu, g_s, d_la.reshape(t._log_alpha.shape)), \
            torch.split(d_y.reshape(*lead, d_t), widths, dim=-1)
    fwd, bwd = _mlp_grad.forward, _mlp_grad.backward
    st_mu = fwd(t._shift_transformation, x2)
    st_s = fwd(t._scale_transformation, x2)
    alpha = torch.exp(t._log_alpha.detach())
    th = torch.tanh(st_s["out"])
    ls = th * alpha
    if not inverse:
        e = torch.exp(ls)
        d_y = g_out * e
        d_mu = g_out
        d_ls = d_y * y2
        if g_dl is not None:
            d_ls = d_ls + g_dl
    else:
        e = torch.exp(-ls)
        d_y = g_out * e
        d_mu = -d_y
        d_ls = -d_y * (y2 - st_mu["out"])
        if g_dl is not None:
            d_ls = d_ls - g_dl
    d_log_alpha = (d_ls * ls).sum().reshape(t._log_alpha.shape)       # d ls / d log_alpha = ls
    d_s = d_ls * alpha * (1 - th * th)
    dx_mu, g_mu = bwd(st_mu, d_mu.contiguous())
    dx_s, g_s = bwd(st_s, d_s)
    d_x = (dx_mu + dx_s).reshape(*lead, x.shape[-1])
    return _affine_pack_grads(t, cond, d_x, g_mu, g_s, d_log_alpha), torch.split(d_y.reshape(*lead, d_t), widths, dim=-1)


def _affine_pack_grads(t, cond, d_x, g_mu, g_s, d_log_alpha):
    """(grads of cond, then the parameter gradients in ``t.parameters()`` order)."""
    g_cond = torch.split(d_x, [c.shape[-1] for c in cond], dim=-1)
    g_mu, g_s = list(g_mu), list(g_s)
    # parameter order of AffineTransformer.parameters(): registration order of the module's members
    named = dict(shift=g_mu, scale=g_s)
    order = []
    for name, _ in t.named_parameters():
        if name == "_log_alpha":
            order.append(d_log_alpha)
        elif name.startswith("_shift_transformation"):
            order.append(named["shift"].pop(0))
        elif name.startswith("_scale_transformation"):
            order.append(named["scale"].pop(0))
        else:
            raise RuntimeError(f"unexpected parameter {name} in an AffineTransformer")
    return (*g_cond, *order)


def _spline_backward(t, cond, tr, params, grads, inverse):
    """Spline block backward: conditioner recompute + its backward, the transform's chain rule in ONE kernel
    (``bgx_spline_backward``).  The conditioner GEMMs run on tensor cores (``_mlp_grad``: exact bf16 splits, three
    products, fp32 accumulation) for plain DenseNets, through torch autograd (fp32 cuBLAS) otherwise.
    Returns (grads of cond + params, grads of tr)."""
    from . import engine, _mlp_grad
    widths = [x.shape[-1] for x in tr]
    y = torch.cat([x.detach() for x in tr], dim=-1) if len(tr) > 1 else tr[0].detach()
    d_t = y.shape[-1]
    lead = y.shape[:-1]
    y2 = y.reshape(-1, d_t)
    g_out = torch.cat([g if g is not None else torch.zeros_like(x) for g, x in zip(grads[:-1], tr)], dim=-1)
    g_dl = grads[-1]
    st = t._default_settings
    out = {}

    def transform_backward(p2, n_out=None):
        k = (n_out or p2.shape[-1]) // (3 * d_t)
        d_p, out["d_y"] = engine.spline_backward(
            p2, y2, g_out.reshape(-1, d_t), g_dl.reshape(-1) if g_dl is not None else None,
            t._end_slope_cols(d_t, k, y.device), k, inverse=inverse, left=t._left, right=t._right, bottom=t._bottom,
            top=t._top, min_bin_width=st["min_bin_width"], min_bin_height=st["min_bin_height"],
            min_derivative=st["min_derivative"], identity_init=st["enable_identity_init"])
        return d_p

    net = t._params_net
    all_params = list(t.parameters())
    mode = engine.backward_gemm_mode()
    all_trainable = len(params) == len(all_params) and len(all_params) == len(list(net.parameters()))
    if mode == "tcgen05" and _mlp_grad.tc_supported(net) and all_trainable:
        # recompute, spline chain rule and conditioner backward in ONE host call (bgx_spline_coupling_backward)
        x = torch.cat([c.detach() for c in cond], dim=-1) if len(cond) > 1 else cond[0].detach()
        tn, lin = _mlp_grad._train_net(net)
        k = lin[-1].weight.shape[0] // (3 * d_t)
        d_x, out["d_y"], g_params = tn.spline_block_backward(
            x.reshape(-1, x.shape[-1]), y2, g_out.reshape(-1, d_t), g_dl.reshape(-1) if g_dl is not None else None,
            t._end_slope_cols(d_t, k, y.device), k, inverse=inverse, left=t._left, right=t._right, bottom=t._bottom,
            top=t._top, min_bin_width=st["min_bin_width"], min_bin_height=st["min_bin_height"],
            min_derivative=st["min_derivative"], identity_init=st["enable_identity_init"])
        g_cond = torch.split(d_x.reshape(*lead, x.shape[-1]), [c.shape[-1] for c in cond], dim=-1)
        gin = (*g_cond, *g_params)
    elif mode == "bf16x3" and _mlp_grad.supported(net) and all_trainable:
        x = torch.cat([c.detach() for c in cond], dim=-1) if len(cond) > 1 else cond[0].detach()
        d_x, g_params = _mlp_grad.forward_backward(net, x.reshape(-1, x.shape[-1]), transform_backward)
        g_cond = torch.split(d_x.reshape(*lead, x.shape[-1]), [c.shape[-1] for c in cond], dim=-1)
        gin = (*g_cond, *g_params)
    else:
        with _matmul_mode(engine.backward_gemm_mode()):
            p = net(torch.cat(cond, dim=-1) if len(cond) > 1 else cond[0])
            d_p = transform_backward(p.detach().reshape(-1, p.shape[-1]))
            gin = torch.autograd.grad(p, [*cond, *params], grad_outputs=d_p.reshape(p.shape), allow_unused=True)
    return gin, torch.split(out["d_y"].reshape(*lead, d_t), widths, dim=-1)


def fused_coupling_with_grad(transformer, kind, cond, tr, inverse):
    """Kernel forward + recompute backward.  Returns (list of outputs, dlogp)."""
    params = list(transformer.parameters())
    res = _FusedCoupling.apply(transformer, kind, inverse, len(cond), len(tr), *cond, *tr, *params)
    return list(res[:-1]), res[-1]


def _recompute_grads(fn, plan, saved, grads):
    """Re-evaluate ``fn(plan, *inputs)`` with the device-side torch definition and pull ``grads`` back."""
    with torch.enable_grad():
        ins = [x.detach().requires_grad_(True) for x in saved]
        outs = fn(plan, *ins)
        pairs = [(o, g) for o, g in zip(outs, grads) if g is not None]
        gin = torch.autograd.grad([o for o, _ in pairs], ins, grad_outputs=[g for _, g in pairs], allow_unused=True)
    return gin


class _ICToXYZ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, launch, bonds, angles, torsions, x0, R):
        with torch.no_grad():
            xyz, dlogp = launch(plan, bonds, angles, torsions, x0, R)
        ctx.plan = plan
        ctx.save_for_backward(bonds, angles, torsions, x0, R)
        return xyz, dlogp

    @staticmethod
    def backward(ctx, g_xyz, g_dlogp):
        gin = _recompute_grads(_torch_math_ic.ic_to_xyz, ctx.plan, ctx.saved_tensors, (g_xyz, g_dlogp))
        return (None, None, *gin)


class _ICFromXYZ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, launch, xyz):
        with torch.no_grad():
            outs = launch(plan, xyz)
        ctx.plan = plan
        ctx.save_for_backward(xyz)
        return outs

    @staticmethod
    def backward(ctx, *grads):
        gin = _recompute_grads(_torch_math_ic.ic_from_xyz, ctx.plan, ctx.saved_tensors, grads)
        return (None, None, *gin)


def ic_to_xyz_with_grad(plan, launch, bonds, angles, torsions, x0, R):
    """IC kernel forward, recompute backward.  ``launch`` is ``engine.ic_to_xyz``."""
    return _ICToXYZ.apply(plan, launch, bonds, angles, torsions, x0, R)


def ic_from_xyz_with_grad(plan, launch, xyz):
    return _ICFromXYZ.apply(plan, launch, xyz)


class _RelICToXYZ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, launch, bonds, angles, torsions, fixed):
        with torch.no_grad():
            xyz, dlogp = launch(plan, bonds, angles, torsions, fixed)
        ctx.plan = plan
        ctx.save_for_backward(bonds, angles, torsions, fixed)
        return xyz, dlogp

    @staticmethod
    def backward(ctx, g_xyz, g_dlogp):
        gin = _recompute_grads(_torch_math_ic.relic_to_xyz, ctx.plan, ctx.saved_tensors, (g_xyz, g_dlogp))
        return (None, None, *gin)


class _RelICFromXYZ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, launch, xyz):
        with torch.no_grad():
            outs = launch(plan, xyz)
        ctx.plan = plan
        ctx.save_for_backward(xyz)
        return outs

    @staticmethod
    def backward(ctx, *grads):
        gin = _recompute_grads(_torch_math_ic.relic_from_xyz, ctx.plan, ctx.saved_tensors, grads)
        return (None, None, *gin)


def relic_to_xyz_with_grad(plan, launch, bonds, angles, torsions, fixed):
    """Relative / mixed IC kernel forward, recompute backward.  ``launch`` is ``engine.relic_to_xyz``."""
    return _RelICToXYZ.apply(plan, launch, bonds, angles, torsions, fixed)


def relic_from_xyz_with_grad(plan, launch, xyz):
    return _RelICFromXYZ.apply(plan, launch, xyz)
