"""2-rank gloo worker for tests/test_host_logic.py::test_gradient_allreduce_two_ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bgflow_b200 import distributed as bd

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
full = torch.randn(16, 5, generator=torch.Generator().manual_seed(1))
lo, hi = bd.shard_rows(16, rank, world)
loss = net(full[lo:hi]).pow(2).sum() / 16          # shard of a mean over the global batch
loss.backward()
n = bd.allreduce_gradients(net.parameters(), average=False)
# reference: the same loss on the whole batch in one process
ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
ref.load_state_dict(net.state_dict())
(ref(full).pow(2).sum() / 16).backward()
err = max((p.grad - q.grad).abs().max().item() for p, q in zip(net.parameters(), ref.parameters()))
assert n == sum(p.numel() for p in net.parameters()), n
assert err < 1e-6, err

# the overlapped path: per-"block" buckets all-reduced from post-accumulate hooks while backward continues
net2 = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
net2.load_state_dict(net.state_dict())
red = bd.BucketedGradReducer(net2, bucket_modules=[net2[2], net2[0]], average=False)
assert red.overlap == "auto" and not red._overlapping()       # a few hundred bytes: one collective after the backward
red.overlap = True
for step in range(2):                      # twice: the buckets are reused, gradients must not pile up
    red.zero_grad()
    (net2(full[lo:hi]).pow(2).sum() / 16).backward()
    n2 = red.finish()
    err2 = max((p.grad - q.grad).abs().max().item() for p, q in zip(net2.parameters(), ref.parameters()))
    assert n2 == n and err2 < 1e-6, (step, n2, err2)
assert red.launched == 4, red.launched     # 2 buckets x 2 steps, each launched from a hook during backward
assert all(p.grad.data_ptr() >= b["flat"].data_ptr() for b in red.buckets[:1] for p in net2[2].parameters())
# the default for small gradients: nothing launched from the hooks, ONE all-reduce of the whole flat buffer in finish()
red.overlap = "auto"
red.zero_grad()
(net2(full[lo:hi]).pow(2).sum() / 16).backward()
assert red.finish() == n and red.launched == 4
err3 = max((p.grad - q.grad).abs().max().item() for p, q in zip(net2.parameters(), ref.parameters()))
assert err3 < 1e-6, err3
if rank == 0:
    print(f"ALLREDUCE_OK world={world} elements={n} err={err:.2e} bucketed_err={err2:.2e}")
dist.destroy_process_group()
