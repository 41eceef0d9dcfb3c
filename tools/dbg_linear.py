import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bgflow_b200 import engine
dev = "cuda:0"
for (b, k, n) in [(1, 128, 128), (300, 33, 128), (300, 825, 128)]:
    for variant in ("rand x, int w, no bias", "int x, rand w, no bias", "rand both, no bias", "rand both, bias", "rand both, bias, kept tensors"):
        g = torch.Generator().manual_seed(k * 1000 + n)
        x = torch.randn(b, k, generator=g) if "int x" not in variant else torch.randint(-3, 4, (b, k), generator=g).float()
        w = torch.randn(n, k, generator=g) / k ** 0.5 if "int w" not in variant else torch.randint(-3, 4, (n, k), generator=g).float()
        bias = torch.randn(n, generator=g) if ", bias" in variant else None
        ref = x.double() @ w.double().t() + (bias.double() if bias is not None else 0)
        f = engine.LinearTC()
        if "kept" in variant:
            xd, wd, bd = x.to(dev), w.to(dev), bias.to(dev)
            y = f(xd, wd, bd)
        else:
            y = f(x.to(dev), w.to(dev), bias.to(dev) if bias is not None else None)
        print(b, k, n, variant, "max err", (y.cpu().double() - ref).abs().max().item(), "scale", ref.abs().max().item())
engine.check_pipeline_status(dev)
