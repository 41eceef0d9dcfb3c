"""Multi-GPU helpers: one process per GPU (torchrun), batch sharded over ranks.

The sample / energy path needs no collective (every sample is independent; parameters are
replicated).  Training needs exactly one: the mean-over-samples loss makes gradients sums over
shards, so ``allreduce_gradients`` sums one flat fp32 bucket over all ranks (NCCL over NVLink on
GPUs; gloo in the CPU tests) between ``backward()`` and ``optimizer.step()`` — the hook the
reference's single-device ``KLTrainer.train`` (bgflow/nn/training/trainers.py:148-201) lacks.
"""

import torch
import torch.distributed as dist

__all__ = ["GraphedStep", "shard_rows", "allreduce_gradients", "kl_train_step", "BucketedGradReducer"]


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


class BucketedGradReducer:
    """Gradient all-reduce overlapped with the backward pass.

    The parameters of every bucket module (by default: every coupling block, i.e. every module with a
    ``transformer`` attribute, plus one bucket for whatever is left) share ONE flat fp32 buffer; their
    ``.grad`` tensors are views into it, so autograd accumulates in place and nothing is packed or copied.
    A post-accumulate hook counts a bucket's parameters; when the last one has its gradient — a block's
    backward is complete — the bucket's ``all_reduce(SUM, async_op=True)`` is launched.  NCCL runs it on its own
    stream (ordered after the work already queued), while the main stream goes on with the backward of the
    blocks earlier in the flow: the collectives of blocks N-1 .. 1 hide behind compute, only block 0's is exposed.
    ``finish()`` (before ``optimizer.step()``) waits for the handles and divides by the world size.

    Works on any backend / device (gloo on CPU in the tests); without a process group it only zeroes / averages."""

    def __init__(self, module, bucket_modules=None, group=None, average=True):
        self.group, self.average = group, average
        if bucket_modules is None:
            bucket_modules = [m for m in module.modules() if hasattr(m, "transformer")]
        seen, groups = set(), []
        for m in bucket_modules:
            ps = [p for p in m.parameters() if p.requires_grad and id(p) not in seen]
            seen.update(id(p) for p in ps)
            if ps:
                groups.append(ps)
        rest = [p for p in module.parameters() if p.requires_grad and id(p) not in seen]
        if rest:
            groups.append(rest)
        self.buckets, self._handles, self._hooks = [], [], []
        # all buckets are slices of ONE flat buffer (bucket starts 128-element aligned): zeroing, averaging and the
        # reduce-after-backward mode touch it with one kernel / one collective
        sizes = [(sum(p.numel() for p in ps) + 127) // 128 * 128 for ps in groups]
        dev = groups[0][0].device if groups else torch.device("cpu")
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self._n_params = sum(p.numel() for ps in groups for p in ps)
        start = 0
        for ps, size in zip(groups, sizes):
            flat = self.flat[start:start + size]
            start += size
            off = 0
            for p in ps:
                if p.dtype != torch.float32 or p.device != flat.device:
                    raise NotImplementedError("BucketedGradReducer: fp32 parameters on one device")
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            b = {"flat": flat, "n": len(ps), "ready": 0}
            self.buckets.append(b)
            for p in ps:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))
        self.launched = 0
        # True: every bucket's all-reduce is launched from its hook, while the backward of the earlier blocks runs.
        # False: nothing is launched from the hooks; ``finish`` reduces the whole flat buffer with ONE collective.
        # "auto" (default): overlap only when the gradients are large (>= ``overlap_min_bytes``).  Measured on 2 B200s
        # (profiles/): with the 4 MB of the 8-block Ala2 stack the collective is 0.3 ms of a 7.5 ms step, and
        # overlapping it COSTS 1.1 ms — the fused kernels are persistent grids of one CTA per SM with a static work
        # split, so every SM an NCCL kernel holds delays one CTA's whole share of a launch.
        self.overlap = "auto"
        self.overlap_min_bytes = 64 << 20

    def _overlapping(self):
        if self.overlap == "auto":
            return 4 * self.flat.numel() >= self.overlap_min_bytes
        return bool(self.overlap)

    @property
    def n_elements(self):
        """Gradient elements (the flat buffer pads every bucket to 128 elements on top of that)."""
        return self._n_params

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _make_hook(self, bucket):
        def hook(param):
            bucket["ready"] += 1
            if bucket["ready"] == bucket["n"]:
                bucket["ready"] = 0
                if self._overlapping() and self._active():
                    self._handles.append(dist.all_reduce(bucket["flat"], op=dist.ReduceOp.SUM, group=self.group,
                                                         async_op=True))
                    self.launched += 1
        return hook

    def zero_grad(self):
        """Use instead of ``optimizer.zero_grad()``: the gradients must stay views of the buckets."""
        self.flat.zero_()
        for b in self.buckets:
            b["ready"] = 0

    def finish(self):
        """Wait for the collectives launched during backward and average.  Returns the number of elements reduced."""
        if not self._overlapping() and self._active():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        for h in self._handles:
            h.wait()
        self._handles.clear()
        if self.average and self._active():
            self.flat.div_(dist.get_world_size(self.group))
        return self.n_elements

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()


class GraphedStep:
    """One training step — everything ``step_fn()`` launches: sampling, the flow's forward, the backward, the gradient
    all-reduce, the optimizer update — captured into ONE CUDA graph after ``warmup`` eager calls and replayed from then
    on.  A reverse-KL step of the 8-block Ala2 stack is ~300 kernels behind ~6 ms of Python; a replay is one launch.

    Requirements (the usual ones of whole-step capture): static shapes; gradients that stay where they are
    (``BucketedGradReducer`` keeps them as views of one buffer; use ``reducer.zero_grad()``, not ``set_to_none``);
    an optimizer that does not synchronise (``torch.optim.Adam(..., capturable=True)``); no host reads inside
    ``step_fn`` — return device tensors (e.g. the loss) and read them after ``__call__``: a replay refreshes the
    same tensors.  The re-packing of the updated weights for the kernels is part of the captured work, so replays
    always run on the current parameters; ``close()`` (or leaving the ``with`` block) bumps the parameters' version
    counters so that eager calls made afterwards re-pack too."""

    def __init__(self, step_fn, parameters, warmup=3):
        self.step_fn, self.params, self.warmup = step_fn, list(parameters), max(1, int(warmup))
        self.calls, self.graph, self.result = 0, None, None
        self._stream = None

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
            return self.result
        self.calls += 1
        dev = self.params[0].device
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        if self.calls <= self.warmup:
            # the eager calls run on the side stream the capture will use: autograd's gradient-accumulation nodes
            # remember the stream they were created on, and a capture must not touch the default stream
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                out = self.step_fn()
            cur.wait_stream(self._stream)
            return out
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream, capture_error_mode="thread_local"):
            self.result = self.step_fn()
        self.graph = g
        g.replay()              # the capture itself executes nothing
        return self.result

    def close(self):
        if self.graph is not None:
            with torch.no_grad():
                for p in self.params:
                    p.add_(0)        # version bump: parameter-keyed caches (packed weights) refresh on the next eager call
            self.graph = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def kl_train_step(generator, optimizer, n_samples_per_rank, temperature=1.0, reducer=None):
    """One reverse-KL step of a data-parallel run: every rank draws its own samples
    (bg.py:13-17 ``kldiv``), gradients are all-reduced, all ranks take the same optimiser step.
    With a ``BucketedGradReducer`` the all-reduce of every coupling block overlaps the rest of the backward;
    without one a single flat all-reduce follows it.  Returns the rank-local loss value (a tensor)."""
    if reducer is not None:
        reducer.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=True)
    loss = generator.kldiv(n_samples_per_rank, temperature=temperature).mean()
    loss.backward()
    if reducer is not None:
        reducer.finish()
    else:
        allreduce_gradients(generator.parameters())
    optimizer.step()
    return loss.detach()
