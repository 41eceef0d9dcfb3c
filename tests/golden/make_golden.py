#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the REFERENCE itself.

Runs only in the build container (needs /root/reference):

    cd /root/repo && PYTHONPATH=oracle/_stubs/shim:oracle/_stubs:/root/reference:. \
        python tests/golden/make_golden.py

The reference modules (bgflow.SequentialFlow / CouplingFlow / AffineTransformer /
ConditionalSplineTransformer / GlobalInternalCoordinateTransformation /
BoltzmannGenerator) are imported unmodified; ``oracle/_stubs/shim`` restores
``numpy.infty`` and ``oracle/_stubs/nflows`` supplies the one third-party function
the reference cannot import here (see oracle/__init__.py).  Parameters come from
``oracle.flows.make_stack(seed=...)`` so that tests can rebuild them from the seed;
a float64 checksum of the parameters is stored to detect RNG drift.

Every fixture stores inputs and the reference's outputs in fp32 (the comparand) and
fp64 (the truth).
"""

import os
import sys

import numpy as np
import torch

import bgflow as bg
from bgflow.nn.periodic import WrapPeriodic

from oracle import flows as of
from oracle import ic as oic

HERE = os.path.dirname(os.path.abspath(__file__))
_ACT = {"relu": torch.nn.ReLU, "silu": torch.nn.SiLU, "tanh": torch.nn.Tanh}


def ref_densenet(mlp, dtype):
    dims = [mlp.weights[0].shape[1]] + [w.shape[0] for w in mlp.weights]
    net = bg.DenseNet(dims, activation=_ACT[mlp.act]() if mlp.act != "none" else None)
    net = net.to(dtype)
    linears = [m for m in net._layers if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for lin, w, b in zip(linears, mlp.weights, mlp.biases):
            lin.weight.copy_(w)
            lin.bias.copy_(b)
    if mlp.periodic is not None:
        idx, left, right = mlp.periodic
        net = WrapPeriodic(net, left=left, right=right, indices=list(idx))
    return net


def ref_transformer(block, dtype):
    if block["kind"] == "affine":
        t = bg.AffineTransformer(
            shift_transformation=ref_densenet(block["shift"], dtype) if block.get("shift") else None,
            scale_transformation=ref_densenet(block["scale"], dtype) if block.get("scale") else None,
            preserve_volume=block.get("preserve_volume", False),
            is_circular=block.get("is_circular", False)).to(dtype)
        with torch.no_grad():
            t._log_alpha.fill_(block.get("log_alpha", -1.0))
        return t
    circ = block.get("is_circular")
    if circ is None:
        circ = False
    return bg.ConditionalSplineTransformer(
        ref_densenet(block["params_net"], dtype), is_circular=torch.as_tensor(circ),
        left=block.get("left", 0.0), right=block.get("right", 1.0),
        bottom=block.get("bottom", 0.0), top=block.get("top", 1.0))


def ref_stack(blocks, split, dtype):
    layers = [bg.SplitFlow(split)]
    for blk in blocks:
        layers.append(bg.CouplingFlow(ref_transformer(blk, dtype)))
        layers.append(bg.SwapFlow())
    layers.append(bg.MergeFlow(split))
    return bg.SequentialFlow(layers)


def checksum(blocks):
    tot = 0.0
    for blk in blocks:
        for key in ("shift", "scale", "params_net"):
            if blk.get(key) is not None:
                for t in blk[key].weights + blk[key].biases:
                    tot += float(t.double().abs().sum())
    return tot


def gen_stack(name, kind, dim, n_blocks, hidden, batch, seed, data):
    out = {}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        blocks, split = of.make_stack(kind, dim, n_blocks, hidden=hidden, seed=seed, dtype=dtype)
        flow = ref_stack(blocks, split, dtype)
        g = torch.Generator().manual_seed(seed + 1)
        z = torch.randn(batch, dim, generator=g, dtype=torch.float64)
        if data == "uniform":
            z = torch.rand(batch, dim, generator=g, dtype=torch.float64)
        z = z.to(dtype)
        with torch.no_grad():
            x, dlogp = flow(z)
            zi, dlogpi = flow(x, inverse=True)
            # also the inverse direction evaluated on the raw input (NLL path on data)
            zi2, dlogpi2 = flow(z, inverse=True)
        out.update({f"z_{tag}": z.numpy(), f"x_{tag}": x.numpy(), f"dlogp_{tag}": dlogp.numpy(),
                    f"zi_{tag}": zi.numpy(), f"dlogpi_{tag}": dlogpi.numpy(),
                    f"inv_x_{tag}": zi2.numpy(), f"inv_dlogp_{tag}": dlogpi2.numpy()})
        if tag == "f32":
            out["param_checksum"] = np.float64(checksum(blocks))
    out["meta"] = np.array([dim, n_blocks, batch, seed] + list(hidden))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


def gen_readme():
    """Config 1: README.md:54-96 (D=2, one RealNVP block, ReLU shift / Tanh scale nets, B=1024)."""
    out = {}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        g = torch.Generator().manual_seed(0)
        shift = of.make_mlp([1, 4, 1], "relu", g, dtype)
        scale = of.make_mlp([1, 4, 1], "tanh", g, dtype)
        blk = {"kind": "affine", "shift": shift, "scale": scale, "log_alpha": -1.0}
        flow = bg.SequentialFlow([
            bg.SplitFlow(1), bg.CouplingFlow(ref_transformer(blk, dtype)), bg.InverseFlow(bg.SplitFlow(1))])
        prior = bg.NormalDistribution(2)
        target = bg.DoubleWellEnergy(2)
        gen = bg.BoltzmannGenerator(prior, flow, target).to(dtype)
        z = torch.randn(1024, 2, generator=torch.Generator().manual_seed(1), dtype=torch.float64).to(dtype)
        with torch.no_grad():
            x, dlogp = flow(z)
            nll = gen.energy(x)
            kld_like = target.energy(x) - dlogp       # bg.py:13-17 evaluated on this z
        out.update({f"z_{tag}": z.numpy(), f"x_{tag}": x.numpy(), f"dlogp_{tag}": dlogp.numpy(),
                    f"nll_{tag}": nll.numpy(), f"kl_{tag}": kld_like.numpy()})
    np.savez_compressed(os.path.join(HERE, "readme_doublewell.npz"), **out)
    print("readme", {k: v.shape for k, v in out.items()})


def gen_multi_tensor():
    """Builder-style coupling (coupling.py:162-182 with several tensors on each side,
    tests/nn/flow/test_coupling.py:86-126): state = (bonds[7], angles[6], torsions[5], aug[4]).
    Block A: circular spline on torsions conditioned on (bonds, aug) ;
    Block B: spline on (bonds, angles) conditioned on (torsions[circular->cos/sin], aug);
    Block C: affine on aug conditioned on (angles,), shift-only + circular;
    Block D: affine on (angles, aug) conditioned on bonds, volume preserving."""
    out = {}
    widths = [7, 6, 5, 4]
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        g = torch.Generator().manual_seed(7)
        nb = 6
        blk_a = {"kind": "spline", "transformed": (2,), "cond": (0, 3), "is_circular": True,
                 "params_net": of.make_mlp([11, 32, 32, 5 * 3 * nb], "silu", g, dtype)}
        net_b = of.make_mlp([2 * 5 + 4, 48, 13 * (3 * nb + 1)], "tanh", g, dtype)
        net_b.periodic = (list(range(5)), 0.0, 1.0)
        blk_b = {"kind": "spline", "transformed": (0, 1), "cond": (2, 3), "params_net": net_b}
        blk_c = {"kind": "affine", "transformed": (3,), "cond": (1,), "is_circular": True,
                 "shift": of.make_mlp([6, 16, 4], "relu", g, dtype), "scale": None}
        blk_d = {"kind": "affine", "transformed": (1, 3), "cond": (0,), "preserve_volume": True,
                 "shift": of.make_mlp([7, 24, 24, 10], "silu", g, dtype),
                 "scale": of.make_mlp([7, 24, 24, 10], "silu", g, dtype), "log_alpha": -0.5}
        blocks = [blk_a, blk_b, blk_c, blk_d]
        layers = [bg.CouplingFlow(ref_transformer(b, dtype), transformed_indices=b["transformed"],
                                  cond_indices=b["cond"]) for b in blocks]
        flow = bg.SequentialFlow(layers)
        gd = torch.Generator().manual_seed(8)
        xs = [torch.rand(48, w, generator=gd, dtype=torch.float64).to(dtype) for w in widths]
        with torch.no_grad():
            *ys, dlogp = flow(*xs)
            *zs, dlogpi = flow(*ys, inverse=True)
        for i in range(4):
            out[f"in{i}_{tag}"] = xs[i].numpy()
            out[f"out{i}_{tag}"] = ys[i].numpy()
            out[f"back{i}_{tag}"] = zs[i].numpy()
        out[f"dlogp_{tag}"] = dlogp.numpy()
        out[f"dlogpi_{tag}"] = dlogpi.numpy()
    np.savez_compressed(os.path.join(HERE, "multi_tensor_coupling.npz"), **out)
    print("multi_tensor", len(out))


def gen_ic(name, z_matrix, xyz0, batch, noise, seed, normalize=True):
    out = {"z_matrix": np.asarray(z_matrix)}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        ic = bg.GlobalInternalCoordinateTransformation(z_matrix, normalize_angles=normalize,
                                                       raise_warnings=False)
        g = torch.Generator().manual_seed(seed)
        x = torch.as_tensor(xyz0, dtype=torch.float64).reshape(1, -1)
        x = (x + noise * torch.randn(batch, x.shape[1], generator=g, dtype=torch.float64)).to(dtype)
        bonds, angles, torsions, x0, R, dlogp = ic._forward(x)
        xr, dlogp_inv = ic._inverse(bonds, angles, torsions, x0, R)
        vals = dict(xyz=x, bonds=bonds, angles=angles, torsions=torsions, x0=x0, R=R, dlogp=dlogp,
                    xyz_back=xr, dlogp_inv=dlogp_inv)
        # the sampling direction on "generated" ICs (uniform-ish perturbation of the data ICs)
        gi = torch.Generator().manual_seed(seed + 100)

        def jitter(t, s):
            return (t.double() + s * torch.randn(t.shape, generator=gi, dtype=torch.float64)).to(dtype)
        b2, a2, t2 = jitter(bonds, 0.003).abs(), jitter(angles, 0.01).clamp(0.05, 0.95), \
            jitter(torsions, 0.05) % 1.0 if normalize else jitter(torsions, 0.3)
        if not normalize:
            a2 = jitter(angles, 0.03).clamp(0.2, 2.9)
        R2 = torch.rand(batch, 3, generator=gi, dtype=torch.float64)
        R2[:, 1] = R2[:, 1] * 1.8 - 0.9
        if not normalize:
            R2[:, 0] = R2[:, 0] * 6 - 3
            R2[:, 2] = R2[:, 2] * 6 - 3
        R2 = R2.to(dtype)
        x02 = torch.randn(batch, 1, 3, generator=gi, dtype=torch.float64).to(dtype)
        x_gen, dlogp_gen = ic._inverse(b2, a2, t2, x02, R2)
        vals.update(gen_bonds=b2, gen_angles=a2, gen_torsions=t2, gen_x0=x02, gen_R=R2,
                    gen_xyz=x_gen, gen_dlogp=dlogp_gen)
        for k, v in vals.items():
            out[f"{k}_{tag}"] = v.detach().numpy()
    out["normalize"] = np.array(int(normalize))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, len(out))


def main():
    torch.set_num_threads(4)
    gen_readme()
    gen_stack("affine_d66_8blk", "affine", 66, 8, (128, 128, 128), 96, 0, "normal")
    gen_stack("spline_d66_8blk", "spline", 66, 8, (128, 128), 96, 0, "uniform")
    gen_stack("affine_d10_3blk", "affine", 10, 3, (16,), 33, 3, "normal")
    gen_stack("spline_d7_4blk", "spline", 7, 4, (24, 24), 33, 4, "uniform")
    gen_multi_tensor()
    gen_ic("ic_ala2", oic.ALA2_GLOBAL_Z, oic.ALA2_XYZ, 64, 0.01, 11, normalize=True)
    gen_ic("ic_ala2_raw", oic.ALA2_GLOBAL_Z, oic.ALA2_XYZ, 16, 0.01, 12, normalize=False)
    g = torch.Generator().manual_seed(5)
    chain = torch.cumsum(torch.randn(12, 3, generator=g, dtype=torch.float64) * 0.6 + 0.5, dim=0).numpy()
    gen_ic("ic_chain12", oic.chain_z_matrix(12), chain, 32, 0.02, 13, normalize=True)


if __name__ == "__main__":
    sys.exit(main())
