#!/usr/bin/env python
"""Run the SURVEY 8f kernels a few times on Ala2 at B=2^20 (profiling target): the multi-field
CDF map, the fused icdf + IC tail, the relative / mixed IC kernels and the split / merge copy."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgflow_b200 as bg
from bgflow_b200 import engine
from bgflow_b200 import fixtures as oic

B = 1 << 20
dev = "cuda"
one = lambda n, v=1.0: torch.full((n,), v, device=dev)
inf = torch.tensor(math.inf, device=dev)
m = {"bonds": bg.TruncatedNormalDistribution(one(21), one(21), torch.tensor(1e-5, device=dev), inf),
     "angles": bg.TruncatedNormalDistribution(one(20, 0.5), one(20), torch.tensor(1e-5, device=dev), torch.tensor(1.0, device=dev)),
     "torsions": bg.SloppyUniform(torch.zeros(19, device=dev), one(19)),
     "fixed": torch.distributions.Normal(torch.zeros(9, device=dev), 20 * one(9)),
     "augmented": torch.distributions.Normal(torch.zeros(10, device=dev), one(10))}
names = ("bonds", "angles", "torsions", "fixed", "augmented")
us = [(torch.rand(B, w) * 0.9 + 0.05).to(dev) for w in (21, 20, 19, 9, 10)]
multi = bg.InverseFlow(bg.MultiCDFFlow([m[n] for n in names]))
ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
tail = bg.MappedICTail(ic, [m["bonds"], m["angles"], m["torsions"]])
x0 = torch.zeros(1, 3, device=dev)
R = torch.full((1, 3), 0.5, device=dev)
rel = bg.RelativeInternalCoordinateTransformation(oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK)
ref = torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float32).reshape(1, -1)
xyz = (ref + 0.01 * torch.randn(B, 66)).to(dev)
mixed = bg.MixedCoordinateTransformation(ref + 0.02 * torch.randn(2000, 66), oic.ALA2_RELATIVE_Z, oic.ALA2_RIGID_BLOCK,
                                         keepdims=9)
z = torch.rand(B, 66, device=dev)
with torch.no_grad():
    for _ in range(3):
        *xs, d = multi(*us)
        back = multi(*xs, inverse=True)
        x, d = tail(us[0], us[1], us[2], x0, R)
        ics = tail(x, inverse=True)
        *r, d = rel(xyz)
        rel(*r, inverse=True)
        *q, d = mixed(xyz)
        mixed(*q, inverse=True)
        a, b = engine.split_cols(z, [33, 33])
        engine.merge_cols([a, b])
torch.cuda.synchronize()
print("ok", float((back[0] - us[0]).abs().max()))
