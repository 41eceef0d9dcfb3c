"""Multi-GPU helpers: one process per GPU (torchrun), batch sharded over ranks.

The sample / energy path needs no collective (every sample is independent; parameters are
replicated).  Training needs exactly one: the mean-over-samples loss makes gradients sums over
shards, so ``allreduce_gradients`` sums one flat fp32 bucket over all ranks (NCCL over NVLink on
GPUs; gloo in the CPU tests) between ``backward()`` and ``optimizer.step()`` — the hook the
reference's single-device ``KLTrainer.train`` (bgflow/nn/training/trainers.py:148-201) lacks.
"""

import torch
import torch.distributed as dist

__all__ = ["shard_rows", "allreduce_gradients", "kl_train_step"]


def shard_rows(n_total, rank=None, world=None):
    """Contiguous row range [lo, hi) of this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    per = (n_total + world - 1) // world
    return min(rank * per, n_total), min((rank + 1) * per, n_total)


def allreduce_gradients(parameters, average=True, group=None):
    """Sum (or average) the gradients of ``parameters`` over all ranks with ONE all-reduce of a
    flat fp32 bucket.  Parameters without a gradient contribute zeros (every rank must call this
    with the same parameter list).  Returns the number of elements reduced."""
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return int(flat.numel())


def kl_train_step(generator, optimizer, n_samples_per_rank, temperature=1.0):
    """One reverse-KL step of a data-parallel run: every rank draws its own samples
    (bg.py:13-17 ``kldiv``), gradients are all-reduced, all ranks take the same optimiser step.
    Returns the rank-local loss value (a tensor)."""
    optimizer.zero_grad(set_to_none=True)
    loss = generator.kldiv(n_samples_per_rank, temperature=temperature).mean()
    loss.backward()
    allreduce_gradients(generator.parameters())
    optimizer.step()
    return loss.detach()
