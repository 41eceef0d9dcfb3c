"""Build bgflow_b200 modules that carry exactly the oracle's seeded parameters."""

import torch

import bgflow_b200 as bg

_ACT = {"relu": torch.nn.ReLU, "silu": torch.nn.SiLU, "tanh": torch.nn.Tanh}


def densenet_from(mlp, device):
    dims = [mlp.weights[0].shape[1]] + [w.shape[0] for w in mlp.weights]
    net = bg.DenseNet(dims, activation=_ACT[mlp.act]() if mlp.act != "none" else None)
    linears = [m for m in net._layers if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for lin, w, b in zip(linears, mlp.weights, mlp.biases):
            lin.weight.copy_(w.float())
            lin.bias.copy_(b.float())
    if mlp.periodic is not None:
        idx, left, right = mlp.periodic
        net = bg.WrapPeriodic(net, left=left, right=right, indices=list(idx))
    return net.to(device)


def transformer_from(block, device):
    if block["kind"] == "affine":
        t = bg.AffineTransformer(
            shift_transformation=densenet_from(block["shift"], device) if block.get("shift") else None,
            scale_transformation=densenet_from(block["scale"], device) if block.get("scale") else None,
            preserve_volume=block.get("preserve_volume", False),
            is_circular=block.get("is_circular", False))
        with torch.no_grad():
            t._log_alpha.fill_(block.get("log_alpha", -1.0))
        return t.to(device)
    circ = block.get("is_circular")
    return bg.ConditionalSplineTransformer(
        densenet_from(block["params_net"], device), is_circular=False if circ is None else circ,
        left=block.get("left", 0.0), right=block.get("right", 1.0),
        bottom=block.get("bottom", 0.0), top=block.get("top", 1.0)).to(device)


def stack_from(blocks, split, device):
    layers = [bg.SplitFlow(split)]
    for blk in blocks:
        layers.append(bg.CouplingFlow(transformer_from(blk, device)))
        layers.append(bg.SwapFlow())
    layers.append(bg.MergeFlow(split))
    return bg.SequentialFlow(layers).to(device)


def config4_blocks(dtype=torch.float32, seed=11, hidden=(128, 128), n_bins=8):
    """BASELINE config 4, the builder-exact augmented Ala2 stack of
    tests/factory/test_generator_builder.py:45-66: state (BONDS[21], ANGLES[20], TORSIONS[19,
    circular], AUGMENTED[10]); 4 x (T <- Aug, Aug <- T), 2 x (B <- A, A <- B), A <- (T, Aug),
    B <- (A, T, Aug); conditioners DenseNet(in, 128, 128, out) SiLU, torsions enter conditioners
    through WrapPeriodic (cos / sin), 8-bin splines.  Returns oracle block dicts."""
    from oracle import flows as of
    g = torch.Generator().manual_seed(seed)
    B_, A_, T_, X_ = 0, 1, 2, 3
    width = {B_: 21, A_: 20, T_: 19, X_: 10}

    def block(what, on):
        d_t = width[what]
        raw = sum(width[f] for f in on)
        periodic, col = [], 0
        for f in on:
            if f == T_:
                periodic += list(range(col, col + width[f]))
            col += width[f]
        circular = what == T_
        n_out = 3 * n_bins * d_t + (0 if circular else d_t)
        net = of.make_mlp([raw + len(periodic), *hidden, n_out], "silu", g, dtype)
        if periodic:
            net.periodic = (periodic, 0.0, 1.0)
        return {"kind": "spline", "transformed": (what,), "cond": tuple(on), "params_net": net,
                "is_circular": True if circular else None}

    blocks = []
    for _ in range(4):
        blocks += [block(T_, (X_,)), block(X_, (T_,))]
    for _ in range(2):
        blocks += [block(B_, (A_,)), block(A_, (B_,))]
    blocks += [block(A_, (T_, X_)), block(B_, (A_, T_, X_))]
    return blocks


import math


def marginals(dev="cuda:0"):
    """InternalCoordinateMarginals defaults (factory/icmarginals.py:14-77) with this package's classes."""
    one = lambda n, v=1.0: torch.full((n,), v, device=dev)
    return {
        "bonds": bg.TruncatedNormalDistribution(one(21), one(21), torch.tensor(1e-5, device=dev),
                                                torch.tensor(math.inf, device=dev)),
        "angles": bg.TruncatedNormalDistribution(one(20, 0.5), one(20), torch.tensor(1e-5, device=dev),
                                                 torch.tensor(1.0, device=dev)),
        "torsions": bg.SloppyUniform(torch.zeros(19, device=dev), one(19)),
        "fixed": torch.distributions.Normal(torch.zeros(9, device=dev), 20 * one(9)),
        "augmented": torch.distributions.Normal(torch.zeros(10, device=dev), one(10)),
    }
