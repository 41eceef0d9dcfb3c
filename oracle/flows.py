"""Oracle (CPU, torch) for coupling blocks.  TEST INFRASTRUCTURE — not product code.

Every function restates what the cited reference lines compute, using plain
torch ops on whatever dtype the inputs have (fp64 = truth, fp32 = comparand and
CPU-baseline timing).  Citations are relative to /root/reference.

Conventions (bgflow/nn/flow/base.py:17-33): a block maps a tuple of tensors to
``(*ys, dlogp)`` with ``dlogp.shape == batch_shape + (1,)`` and ``dlogp`` the
log|det J| of the direction that was evaluated.
"""

import math

import torch

__all__ = [
    "dense_net", "wrap_periodic", "affine_transformer", "spline_params",
    "rational_quadratic_spline", "spline_transformer", "coupling_block",
    "coupling_stack", "MLP", "make_mlp", "make_stack",
]


# --------------------------------------------------------------------------- MLP

class MLP:
    """Plain container for a DenseNet's parameters (bgflow/nn/dense.py:10-45).

    ``weights[i]`` is ``nn.Linear.weight`` ``[out_i, in_i]``, ``biases[i]`` is
    ``[out_i]``; ``act`` in {"relu", "silu", "tanh", "none"} is applied after every
    layer but the last (dense.py:38-43).  ``periodic`` optionally holds the
    WrapPeriodic settings ``(indices, left, right)`` (bgflow/nn/periodic.py:23-37).
    """

    def __init__(self, weights, biases, act="relu", periodic=None):
        self.weights = list(weights)
        self.biases = list(biases)
        self.act = act
        self.periodic = periodic

    def to(self, dtype):
        return MLP([w.to(dtype) for w in self.weights], [b.to(dtype) for b in self.biases],
                   self.act, self.periodic)


_ACTS = {
    "relu": torch.relu,
    "silu": torch.nn.functional.silu,
    "tanh": torch.tanh,
    "none": lambda x: x,
}


def wrap_periodic(x, indices, left=0.0, right=1.0):
    """bgflow/nn/periodic.py:30-37: [cos.., sin.., others..] (cos block first)."""
    n = x.shape[-1]
    idx = list(indices)
    others = [i for i in range(n) if i not in set(idx)]
    y = x[..., idx]
    arg = 2 * math.pi * (y - left) / (right - left)
    return torch.cat([torch.cos(arg), torch.sin(arg), x[..., others]], dim=-1)


def dense_net(mlp, x):
    """bgflow/nn/dense.py:47-48 (+ periodic.py:30-37 when wrapped)."""
    if mlp.periodic is not None:
        idx, left, right = mlp.periodic
        x = wrap_periodic(x, idx, left, right)
    act = _ACTS[mlp.act]
    n = len(mlp.weights)
    for i, (w, b) in enumerate(zip(mlp.weights, mlp.biases)):
        x = torch.nn.functional.linear(x, w, b)
        if i < n - 1:
            x = act(x)
    return x


# ------------------------------------------------------------------------ affine

def affine_transformer(cond, y, shift=None, scale=None, log_alpha=-1.0, inverse=False,
                       preserve_volume=False, is_circular=False):
    """bgflow/nn/flow/transformer/affine.py:35-70."""
    if shift is not None:
        mu = dense_net(shift, cond)
    else:
        mu = torch.zeros_like(y)
    if scale is not None:
        alpha = math.exp(float(log_alpha))
        log_sigma = torch.tanh(dense_net(scale, cond)) * alpha
        if preserve_volume:
            log_sigma = log_sigma - log_sigma.mean(dim=-1, keepdim=True)
    else:
        log_sigma = torch.zeros_like(y)
    assert mu.shape[-1] == y.shape[-1] and log_sigma.shape[-1] == y.shape[-1]
    if not inverse:
        out = torch.exp(log_sigma) * y + mu
        dlogp = log_sigma.sum(dim=-1, keepdim=True)
    else:
        out = torch.exp(-log_sigma) * (y - mu)
        dlogp = (-log_sigma).sum(dim=-1, keepdim=True)
    if is_circular:
        out = out % 1.0
    return out, dlogp


# ------------------------------------------------------------------------ spline

def spline_params(p, d_t, is_circular=None):
    """bgflow/nn/flow/transformer/spline.py:87-126.

    ``p`` is the conditioner output ``[..., 3*K*d_t + n_nc]``; returns unnormalised
    ``widths [..., d_t, K]``, ``heights [..., d_t, K]``, ``slopes [..., d_t, K+1]``.
    ``is_circular``: None/False = none circular, True = all, or a bool sequence of
    length d_t.  For mixed masks the *intended* count (number of non-circular dims,
    factory/conditioner_factory.py:230-233) is used; it coincides with
    spline.py:190-204 in every case where the reference does not raise.
    """
    if is_circular is None or is_circular is False:
        circ = [False] * d_t
    elif is_circular is True:
        circ = [True] * d_t
    else:
        circ = [bool(c) for c in is_circular]
        assert len(circ) == d_t
    noncirc = [i for i, c in enumerate(circ) if not c]
    n_nc = len(noncirc)
    batch = p.shape[:-1]
    k = p.shape[-1] // (3 * d_t)
    if 3 * k * d_t + n_nc != p.shape[-1]:
        raise RuntimeError(
            f"params_net width {p.shape[-1]} != 3*{k}*{d_t} + {n_nc}")
    w, h, s, sl = torch.split(p, [k * d_t, k * d_t, k * d_t, n_nc], dim=-1)
    w = w.reshape(*batch, d_t, k)
    h = h.reshape(*batch, d_t, k)
    s = s.reshape(*batch, d_t, k)
    slopes = torch.cat([s, s[..., :1]], dim=-1)
    if n_nc:
        slopes[..., noncirc, k] = sl.reshape(*batch, n_nc)
    return w, h, slopes


def _softplus(x, beta):
    # torch.nn.functional.softplus semantics (threshold=20 on beta*x), written out
    bx = beta * x
    return torch.where(bx > 20.0, x, torch.log1p(torch.exp(torch.clamp(bx, max=20.0))) / beta)


def rational_quadratic_spline(inputs, unnormalized_widths, unnormalized_heights,
                              unnormalized_derivatives, inverse=False, left=0.0, right=1.0,
                              bottom=0.0, top=1.0, min_bin_width=1e-3, min_bin_height=1e-3,
                              min_derivative=1e-3, enable_identity_init=True, check_domain=False):
    """Restatement of ``nflows.transforms.splines.rational_quadratic_spline``.

    Third-party dependency of bgflow (spline.py:129-144, 160-175), unpinned
    (.github/workflows/CI.yml:44), absent from /root/reference.  Algorithm of
    Durkan et al. 2019 (Neural Spline Flows), as restated in SURVEY.md Appendix A.5.
    Shapes: inputs ``[...]``; widths/heights ``[..., K]``; derivatives ``[..., K+1]``.
    Returns ``(outputs, logabsdet)`` elementwise.
    """
    if check_domain and (inputs.min() < left or inputs.max() > right):
        raise ValueError("InputOutsideDomain")
    k = unnormalized_widths.shape[-1]
    if min_bin_width * k > 1.0 or min_bin_height * k > 1.0:
        raise ValueError("minimal bin width/height too large for the number of bins")

    def knots(unnorm, minimum, lo, hi):
        frac = torch.softmax(unnorm, dim=-1)
        frac = minimum + (1 - minimum * k) * frac
        cum = torch.cumsum(frac, dim=-1)
        cum = torch.nn.functional.pad(cum, pad=(1, 0), mode="constant", value=0.0)
        cum = (hi - lo) * cum + lo
        cum[..., 0] = lo
        cum[..., -1] = hi
        return cum, cum[..., 1:] - cum[..., :-1]

    cumwidths, widths = knots(unnormalized_widths, min_bin_width, left, right)
    cumheights, heights = knots(unnormalized_heights, min_bin_height, bottom, top)

    beta = math.log(2.0) / (1.0 - min_derivative) if enable_identity_init else 1.0
    derivatives = min_derivative + _softplus(unnormalized_derivatives, beta)

    locs = (cumheights if inverse else cumwidths).clone()
    locs[..., -1] += 1e-6
    bin_idx = (torch.sum(inputs[..., None] >= locs, dim=-1) - 1)[..., None]
    # inputs below the first knot would give -1; the reference raises before that
    bin_idx = bin_idx.clamp(0, k - 1)

    def take(t):
        return t.gather(-1, bin_idx)[..., 0]

    in_cw = take(cumwidths)
    in_w = take(widths)
    in_ch = take(cumheights)
    in_h = take(heights)
    in_delta = take(heights / widths)
    in_d = take(derivatives)
    in_d1 = take(derivatives[..., 1:])

    if inverse:
        q = inputs - in_ch
        s = in_d + in_d1 - 2 * in_delta
        a = q * s + in_h * (in_delta - in_d)
        b = in_h * in_d - q * s
        c = -in_delta * q
        disc = b.pow(2) - 4 * a * c
        root = (2 * c) / (-b - torch.sqrt(disc))
        outputs = root * in_w + in_cw
        t1 = root * (1 - root)
        den = in_delta + s * t1
        num = in_delta.pow(2) * (in_d1 * root.pow(2) + 2 * in_delta * t1 + in_d * (1 - root).pow(2))
        return outputs, -(torch.log(num) - 2 * torch.log(den))
    else:
        th = (inputs - in_cw) / in_w
        t1 = th * (1 - th)
        num_y = in_h * (in_delta * th.pow(2) + in_d * t1)
        den = in_delta + (in_d + in_d1 - 2 * in_delta) * t1
        outputs = in_ch + num_y / den
        num = in_delta.pow(2) * (in_d1 * th.pow(2) + 2 * in_delta * t1 + in_d * (1 - th).pow(2))
        return outputs, torch.log(num) - 2 * torch.log(den)


def spline_transformer(cond, y, params_net, inverse=False, is_circular=None, left=0.0, right=1.0,
                       bottom=0.0, top=1.0, enable_identity_init=True):
    """bgflow/nn/flow/transformer/spline.py:128-188.

    bgflow ``_forward`` calls nflows with ``inverse=True`` (root branch) and
    ``_inverse`` with ``inverse=False`` (direct branch).  Out-of-domain inputs are
    clamped (numerically identical to spline.py:145-155's catch-warn-clamp-retry).
    """
    p = dense_net(params_net, cond)
    w, h, s = spline_params(p, y.shape[-1], is_circular)
    yc = y.clamp(left, right)
    z, lad = rational_quadratic_spline(
        yc, w, h, s, inverse=not inverse, left=left, right=right, bottom=bottom, top=top,
        min_bin_width=1e-3, min_bin_height=1e-3, min_derivative=1e-3,
        enable_identity_init=enable_identity_init)
    return z, lad.sum(dim=-1, keepdim=True)


# ---------------------------------------------------------------------- plumbing

def coupling_block(block, xs, inverse=False):
    """bgflow/nn/flow/coupling.py:162-182: cat transformed / cond inputs, call the
    transformer, split back by the original widths.  ``block`` is a dict:
    ``{"kind": "affine"|"spline", "transformed": (..), "cond": (..), **transformer kwargs}``.
    """
    tr = tuple(block.get("transformed", (1,)))
    cd = tuple(block.get("cond", (0,)))
    lengths = [xs[i].shape[-1] for i in tr]
    y = torch.cat([xs[i] for i in tr], dim=-1)
    cond = torch.cat([xs[i] for i in cd], dim=-1)
    if block["kind"] == "affine":
        out, dlogp = affine_transformer(
            cond, y, block.get("shift"), block.get("scale"), block.get("log_alpha", -1.0),
            inverse=inverse, preserve_volume=block.get("preserve_volume", False),
            is_circular=block.get("is_circular", False))
    elif block["kind"] == "spline":
        out, dlogp = spline_transformer(
            cond, y, block["params_net"], inverse=inverse,
            is_circular=block.get("is_circular"),
            left=block.get("left", 0.0), right=block.get("right", 1.0),
            bottom=block.get("bottom", 0.0), top=block.get("top", 1.0),
            enable_identity_init=block.get("enable_identity_init", True))
    else:
        raise ValueError(block["kind"])
    outs = torch.split(out, lengths, dim=-1)
    xs = list(xs)
    for i, o in zip(tr, outs):
        xs[i] = o
    return xs, dlogp


def coupling_stack(blocks, x, split, inverse=False):
    """SplitFlow -> [CouplingFlow, SwapFlow]* -> MergeFlow over one ``[B, D]`` tensor.

    bgflow/nn/flow/sequential.py:49-59 (dlogp accumulation, reversed order on
    inverse), coupling.py:46-57 (split / cat), coupling.py:118-130 (swap).
    ``blocks`` is a list of block dicts; a SwapFlow follows every coupling
    (including the last, as in notebooks/alanine_dipeptide_basics.py:209-218).
    """
    xs = list(torch.split(x, [split, x.shape[-1] - split], dim=-1))
    dlogp = torch.zeros(*x.shape[:-1], 1, dtype=x.dtype, device=x.device)
    if not inverse:
        for blk in blocks:
            xs, d = coupling_block(blk, xs, inverse=False)
            dlogp = dlogp + d
            xs = [xs[1], xs[0]]
    else:
        for blk in reversed(blocks):
            xs = [xs[1], xs[0]]
            xs, d = coupling_block(blk, xs, inverse=True)
            dlogp = dlogp + d
    return torch.cat(xs, dim=-1), dlogp


# -------------------------------------------------------------- seeded factories

def make_mlp(dims, act, generator, dtype=torch.float32, weight_scale=1.0):
    """Default ``nn.Linear`` init (kaiming-uniform a=sqrt(5) == U(-1/sqrt(in), 1/sqrt(in))
    for weight and bias), drawn from ``generator`` so that the product-side tests can
    build bit-identical parameters without importing this module's internals."""
    ws, bs = [], []
    for d_in, d_out in zip(dims[:-1], dims[1:]):
        bound = 1.0 / math.sqrt(d_in)
        w = (torch.rand(d_out, d_in, generator=generator, dtype=torch.float64) * 2 - 1) * bound
        b = (torch.rand(d_out, generator=generator, dtype=torch.float64) * 2 - 1) * bound
        ws.append((w * weight_scale).to(dtype))
        bs.append(b.to(dtype))
    return MLP(ws, bs, act)


def make_stack(kind, dim, n_blocks, hidden=(128, 128), n_bins=8, seed=0, dtype=torch.float32,
               act=None, weight_scale=1.0):
    """The BASELINE configs 2/3 architecture: D split D/2 | D - D/2,
    ``n_blocks`` couplings alternating halves, seeded parameters."""
    g = torch.Generator().manual_seed(seed)
    d0 = dim // 2
    d1 = dim - d0
    blocks = []
    for i in range(n_blocks):
        d_c, d_t = (d0, d1) if i % 2 == 0 else (d1, d0)
        if kind == "affine":
            a = act or "relu"
            blocks.append({
                "kind": "affine",
                "shift": make_mlp([d_c, *hidden, d_t], a, g, dtype, weight_scale),
                "scale": make_mlp([d_c, *hidden, d_t], a, g, dtype, weight_scale),
                "log_alpha": -1.0,
            })
        else:
            a = act or "silu"
            blocks.append({
                "kind": "spline",
                "params_net": make_mlp([d_c, *hidden, d_t * (3 * n_bins + 1)], a, g, dtype,
                                       weight_scale),
            })
    return blocks, d0
