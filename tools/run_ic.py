#!/usr/bin/env python
"""Run both IC kernels a few times on Ala2 at B=2^20 (profiling target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgflow_b200 as bg
from bgflow_b200 import fixtures as oic

B = 1 << 20
ic = bg.GlobalInternalCoordinateTransformation(oic.ALA2_GLOBAL_Z)
xyz = (torch.as_tensor(oic.ALA2_XYZ, dtype=torch.float32).reshape(1, -1) + 0.01 * torch.randn(B, 66)).cuda()
for _ in range(3):
    *ics, d = ic(xyz)
    back, d2 = ic(*ics, inverse=True)
torch.cuda.synchronize()
print("max round-trip error", (back - xyz).abs().max().item())
