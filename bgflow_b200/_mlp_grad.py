"""Explicit recompute + backward of a ``DenseNet`` conditioner with tensor-core GEMMs (training path).

The reference differentiates ``DenseNet.forward`` (bgflow/nn/dense.py:47-48) with torch autograd: fp32 SIMT GEMMs on a
GPU.  Here the layer GEMMs of the recompute and of the backward (dW = g^T h, dh = g W) run as THREE bf16 tensor-core
products of exact two-term operand splits with fp32 accumulation (hi.hi + hi.lo + lo.hi: ~2^-16 relative per product,
the scheme of the fused forward kernels); the splits come from ``bgx_split_bf16``.  The GEMMs themselves are cuBLASLt's
(``torch.mm(..., out_dtype=float32)``) — plain library GEMMs, as the task allows; what is removed is the 14 ms of SIMT
GEMM time per 20 ms KL step that profiles/r2_train_profile.txt shows for the autograd version.
"""

import torch

from . import engine

__all__ = ["supported", "forward", "backward", "forward_backward", "forward_tc", "backward_tc", "tc_supported"]

_ACTS = (torch.nn.SiLU, torch.nn.ReLU, torch.nn.Tanh)


def _layers(net):
    mods = list(net._layers)
    lin, acts = [], []
    i = 0
    while i < len(mods):
        if not isinstance(mods[i], torch.nn.Linear):
            return None
        lin.append(mods[i])
        i += 1
        if i < len(mods) and not isinstance(mods[i], torch.nn.Linear):
            if not isinstance(mods[i], _ACTS):
                return None
            acts.append(mods[i])
            i += 1
        else:
            acts.append(None)
    if acts and acts[-1] is not None:      # no activation after the output layer (dense.py:40-41)
        return None
    return lin, acts


def supported(net):
    """A plain ``DenseNet`` (Linear / SiLU | ReLU | Tanh / ... / Linear) of fp32 CUDA parameters."""
    from .nets import DenseNet
    if type(net) is not DenseNet:
        return False
    la = _layers(net)
    return la is not None and all(l.weight.is_cuda and l.weight.dtype == torch.float32 for l in la[0])


def tc_supported(net):
    """``supported`` and every layer has at most 128 inputs or at most 128 outputs (what ``bgx_linear`` covers)."""
    return supported(net) and _tc_ok(_layers(net)[0])


def _mm3(a, b):
    """a = (a_hi, a_lo) [M, K], b = (b_hi, b_lo) [K, N] (bf16, any strides) -> fp32 [M, N] ~= a @ b."""
    f32 = torch.float32
    out = torch.mm(a[0], b[0], out_dtype=f32)
    out = torch.addmm(out, a[0], b[1], out_dtype=f32)
    return torch.addmm(out, a[1], b[0], out_dtype=f32)


def _t(pair):
    return pair[0].t(), pair[1].t()


def _act(mod, z):
    if mod is None:
        return z
    return mod(z)


def _act_grad(mod, z, g):
    """g * act'(z)."""
    if isinstance(mod, torch.nn.SiLU):
        s = torch.sigmoid(z)
        return g * (s * (1 + z * (1 - s)))
    if isinstance(mod, torch.nn.ReLU):
        return g * (z > 0)
    if isinstance(mod, torch.nn.Tanh):
        return g * (1 - torch.tanh(z) ** 2)
    return g


def _tc_ok(lin):
    return all(engine.LinearTC.supports(*l.weight.shape) for l in lin)


def _train_net(net):
    lin, acts = _layers(net)
    tn = net.__dict__.get("_bgx_train_net")
    if tn is None:
        tn = net.__dict__["_bgx_train_net"] = engine.TrainNet()
    tn.refresh([l.weight for l in lin], [l.bias for l in lin], [engine._ACT_CODES[type(a)] for a in acts])
    return tn, lin


@torch.no_grad()
def forward_tc(net, x):
    """``forward`` on our own tensor-core kernels, the whole net in one host call (``engine.TrainNet.forward`` =
    ``bgx_mlp_forward_train``: ``bgx_linear`` per layer, exact bf16 operand splits).  Every buffer is padded with zero
    columns to a multiple of 4 floats so that rows are 16-byte aligned: ``state["out_padded"]`` is
    ``[B, pad4(N)]``, ``state["out"]`` its first ``N = state["n_out"]`` columns."""
    tn, lin = _train_net(net)
    state = tn.forward(x)
    state["net"], state["tn"], state["lin"] = net, tn, lin
    return state


@torch.no_grad()
def backward_tc(state, d_out, need_dx=True):
    """``backward`` in one host call (``bgx_mlp_backward``): the input gradients dh = g W on ``bgx_linear``, the weight
    and bias gradients dW = g^T h, db = sum_b g (reductions over the batch) on ``bgx_gemm_tn``.  ``d_out`` is
    ``[B, N]``, or ``[B, pad4(N)]`` (the shape of ``state["out_padded"]``) with zero pad columns."""
    return state["tn"].backward(state, d_out, need_dx)


@torch.no_grad()
def forward(net, x):
    """Re-run ``net`` on ``x`` ``[B, d_in]`` keeping what the backward needs.  Returns a state whose ``["out"]`` is
    the network output ``[B, d_out]``."""
    lin, acts = _layers(net)
    split = engine.split_bf16
    hs = [split(x)]                       # operand splits of every layer input
    zs = []
    for i, (l, a) in enumerate(zip(lin, acts)):
        w = split(l.weight.detach())      # [out, in]
        z = _mm3(hs[-1], _t(w))
        z += l.bias.detach()
        zs.append((z, w))
        if i + 1 < len(lin):
            hs.append(split(_act(a, z)))
    return {"lin": lin, "acts": acts, "hs": hs, "zs": zs, "out": zs[-1][0]}


@torch.no_grad()
def backward(state, d_out, need_dx=True):
    """Gradients of ``sum(out * d_out)``: ``(d_x or None, [dW0, db0, dW1, db1, ...])``."""
    lin, acts, hs, zs = state["lin"], state["acts"], state["hs"], state["zs"]
    split = engine.split_bf16
    g = d_out
    grads = [None] * (2 * len(lin))
    d_x = None
    for i in range(len(lin) - 1, -1, -1):
        z, w = zs[i]
        gs = split(g)
        grads[2 * i] = _mm3(_t(gs), hs[i])            # dW = g^T h   [out, in]
        grads[2 * i + 1] = g.sum(dim=0)
        if i > 0:
            gh = _mm3(gs, w)                          # dh = g W     [B, in]
            g = _act_grad(acts[i - 1], zs[i - 1][0], gh)
        elif need_dx:
            d_x = _mm3(gs, w)
    return d_x, grads


def forward_backward(net, x, grad_fn, need_dx=True):
    """``forward`` + ``backward`` with ``d_out = grad_fn(out)``."""
    st = forward(net, x)
    return backward(st, grad_fn(st["out"]), need_dx)
