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
