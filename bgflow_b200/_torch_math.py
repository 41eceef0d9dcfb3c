"""Differentiable PyTorch (device-side ATen) definitions of the coupling transformers.

Used ONLY by the backward pass of ``bgflow_b200.autograd`` (recompute-in-backward): the forward
of a coupling block always runs the fused CUDA kernel; when gradients are requested the block is
re-evaluated here on the GPU under ``torch.enable_grad()`` and differentiated by autograd.  This
is a stop-gap until the backward kernels exist — it is never a forward path and
never runs on the CPU.

Math: bgflow/nn/flow/transformer/affine.py:35-70, spline.py:87-188 and the rational-quadratic
spline of Durkan et al. 2019 (nflows semantics).
"""

import math

import torch

__all__ = ["affine", "spline"]


def affine(transformer, cond, y, inverse):
    t = transformer
    mu = t._shift_transformation(cond) if t._shift_transformation is not None else torch.zeros_like(y)
    if t._scale_transformation is not None:
        log_sigma = torch.tanh(t._scale_transformation(cond)) * torch.exp(t._log_alpha.to(cond))
        if t._preserve_volume:
            log_sigma = log_sigma - log_sigma.mean(dim=-1, keepdim=True)
    else:
        log_sigma = torch.zeros_like(y)
    if inverse:
        out = torch.exp(-log_sigma) * (y - mu)
        dlogp = (-log_sigma).sum(dim=-1, keepdim=True)
    else:
        out = torch.exp(log_sigma) * y + mu
        dlogp = log_sigma.sum(dim=-1, keepdim=True)
    if t._is_circular:
        out = out % 1.0
    return out, dlogp


def _knots(unnorm, minimum, lo, hi):
    k = unnorm.shape[-1]
    frac = minimum + (1 - minimum * k) * torch.softmax(unnorm, dim=-1)
    cum = torch.cumsum(frac, dim=-1)
    cum = torch.nn.functional.pad(cum, pad=(1, 0), value=0.0)
    cum = (hi - lo) * cum + lo
    first = torch.full_like(cum[..., :1], lo)
    last = torch.full_like(cum[..., :1], hi)
    cum = torch.cat([first, cum[..., 1:-1], last], dim=-1)
    return cum, cum[..., 1:] - cum[..., :-1]


def spline(transformer, cond, y, inverse):
    t = transformer
    d_t = y.shape[-1]
    p = t._params_net(cond)
    batch = p.shape[:-1]
    k = p.shape[-1] // (3 * d_t)
    mask = t._circular_mask(d_t)
    noncirc = [i for i, c in enumerate(mask) if not c]
    w, h, s, sl = torch.split(p, [k * d_t, k * d_t, k * d_t, len(noncirc)], dim=-1)
    w, h, s = (v.reshape(*batch, d_t, k) for v in (w, h, s))
    end = s[..., :1]
    if noncirc:
        end = end.clone()
        end[..., noncirc, 0] = sl
    slopes = torch.cat([s, end], dim=-1)
    st = t._default_settings
    left, right, bottom, top = t._left, t._right, t._bottom, t._top
    cw, widths = _knots(w, st["min_bin_width"], left, right)
    ch, heights = _knots(h, st["min_bin_height"], bottom, top)
    beta = math.log(2.0) / (1.0 - st["min_derivative"]) if st["enable_identity_init"] else 1.0
    deriv = st["min_derivative"] + torch.nn.functional.softplus(slopes, beta=beta)
    x = y.clamp(left, right)
    root_branch = not inverse                       # bgflow forward == nflows inverse=True
    locs = (ch if root_branch else cw).detach().clone()
    locs[..., -1] += 1e-6
    idx = ((x[..., None] >= locs).sum(dim=-1) - 1).clamp(0, k - 1)[..., None]

    def take(v):
        return v.gather(-1, idx)[..., 0]

    in_cw, in_w, in_ch, in_h = take(cw), take(widths), take(ch), take(heights)
    delta = take(heights / widths)
    d0, d1 = take(deriv), take(deriv[..., 1:])
    ss = d0 + d1 - 2 * delta
    if root_branch:
        q = x - in_ch
        a = q * ss + in_h * (delta - d0)
        b = in_h * d0 - q * ss
        c = -delta * q
        th = (2 * c) / (-b - torch.sqrt((b * b - 4 * a * c).clamp_min(0.0)))
        out = th * in_w + in_cw
    else:
        th = (x - in_cw) / in_w
    t1 = th * (1 - th)
    den = delta + ss * t1
    if not root_branch:
        out = in_ch + in_h * (delta * th * th + d0 * t1) / den
    num = delta * delta * (d1 * th * th + 2 * delta * t1 + d0 * (1 - th) ** 2)
    lad = torch.log(num) - 2 * torch.log(den)
    if root_branch:
        lad = -lad
    return out, lad.sum(dim=-1, keepdim=True)
