"""Host-buffer entry points: run a flow on data that lives in (pinned) host memory.

``HostPipeline.run(z_host)`` is the call ``bench.py`` times for its end-to-end number: the
batch is cut into chunks that are copied host->device, pushed through the flow and copied back
device->host on a small ring of CUDA streams, so the PCIe transfers of neighbouring chunks
overlap the kernels of the current one (the reference does ``x.to(device)`` / ``.cpu()`` around a
monolithic call, bgflow/bg.py:115-117).

``HostPipeline.sample(n)`` is the sampling call of a generator (bgflow/bg.py:105-135): the prior is
drawn ON THE DEVICE chunk by chunk (``prior.sample`` — bgflow/distribution/normal.py:75-92 draws on the
parameter's device too), so nothing crosses PCIe host->device; only samples, ``dlogp`` and (optionally) the
generator energy ``prior.energy(z) + dlogp`` (bg.py:124-126) come back.

``bind_to_gpu_numa`` pins the calling process to the CPUs of the GPU's NUMA node BEFORE pinned buffers are
allocated (first-touch places them on that node); ``copy_ceiling`` measures what the host<->device link
gives this process for concurrent H2D + D2H copies — the denominator of the end-to-end number.
"""

import os

import torch

__all__ = ["HostPipeline", "bind_to_gpu_numa", "copy_ceiling", "wave_rows"]


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index):
    """Restrict this process to the CPUs local to GPU ``device_index`` (sysfs ``local_cpulist`` of its PCI
    function).  Returns ``{"numa_node", "cpus", "bound"}``; a no-op (bound False) where sysfs does not say."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        with open(os.path.join(path, "numa_node")) as f:
            info["numa_node"] = int(f.read().strip())
        with open(os.path.join(path, "local_cpulist")) as f:
            text = f.read().strip()
        cpus = _parse_cpulist(text) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(cpus=text, bound=True)
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return info


def copy_ceiling(device, nbytes=256 << 20, reps=4):
    """Concurrent pinned H2D + D2H ``cudaMemcpyAsync`` on two streams: GB/s per direction for this
    process (call it on every rank at once to see what the box gives N GPUs together)."""
    device = torch.device(device)
    n = nbytes // 4
    h_in = torch.empty(n, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n, dtype=torch.float32).pin_memory()
    h_in.zero_()
    h_out.zero_()
    d_in = torch.empty(n, dtype=torch.float32, device=device)
    d_out = torch.zeros(n, dtype=torch.float32, device=device)
    s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    out = {}
    for mode in ("h2d", "d2h", "both"):
        ms = {}
        for _ in range(2):       # first pass = warm-up
            torch.cuda.synchronize(device)
            ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for k in ("h2d", "d2h")}
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    ev["h2d"][0].record()
                    for _ in range(reps):
                        d_in.copy_(h_in, non_blocking=True)
                    ev["h2d"][1].record()
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    ev["d2h"][0].record()
                    for _ in range(reps):
                        h_out.copy_(d_out, non_blocking=True)
                    ev["d2h"][1].record()
            torch.cuda.synchronize(device)
            for k in ("h2d", "d2h"):
                if mode in (k, "both"):
                    ms[k] = ev[k][0].elapsed_time(ev[k][1])
        for k, v in ms.items():
            out[f"{k}_gbs" + ("_concurrent" if mode == "both" else "")] = reps * nbytes / (v * 1e-3) / 1e9
    return out


def wave_rows(device, waves=4):
    """Rows of ``waves`` full waves of the fused coupling kernels: one persistent CTA per SM working on two 128-row
    tiles at a time, so a launch of 256 x SMs x waves rows keeps every SM busy to the end (a 2^17-row chunk is 3.46
    waves on 148 SMs: the fourth one runs half empty).  Measured on the 8-block Ala2 stack, 2^20 rows end to end:
    9.2 ms with 3- or 4-wave chunks, 9.5 ms with 2^17 rows, 12.8 ms with 2 waves (launch-bound on the host)."""
    sms = torch.cuda.get_device_properties(torch.device(device)).multi_processor_count
    return 256 * sms * int(waves)


class HostPipeline:
    """``use_graph``: the whole call — every chunk's copies and kernels on the ring of streams — is captured into ONE
    CUDA graph the second time it is made with the same host buffer and row count, and replayed from then on: a
    replay is one launch for the host instead of ~10 per coupling block and chunk, so the chunks can be short (the
    un-overlapped first copy in / last copy out shrink with them) without the host becoming the bottleneck."""

    def __init__(self, flow, dim_in, dim_out, max_rows, device, chunk_rows=None, n_streams=3,
                 inverse=False, prior=None, with_energy=False, use_graph=True):
        self.flow = flow
        self.device = torch.device(device)
        self.chunk = int(chunk_rows) if chunk_rows else wave_rows(self.device, 2 if use_graph else 4)
        self.chunk_eager = int(chunk_rows) if chunk_rows else wave_rows(self.device, 4)
        self.inverse = inverse
        self.prior = prior
        self.with_energy = with_energy
        self.use_graph = use_graph
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self.out = torch.empty(max_rows, dim_out, dtype=torch.float32).pin_memory()
        self.dlogp = torch.empty(max_rows, 1, dtype=torch.float32).pin_memory()
        self.energy = torch.empty(max_rows, 1, dtype=torch.float32).pin_memory() if with_energy else None
        self.dim_in = dim_in
        self._status = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._graphs, self._warm = {}, set()

    def _flow_version(self):
        """A graph holds the addresses of the packed parameters it was captured with: a parameter update (or move)
        must not replay it."""
        return tuple((p.data_ptr(), p._version) for p in self.flow.parameters())

    def _ranges(self, B, chunk):
        """Row ranges of the chunks of a batch of B rows.  (Uniform chunks: starting and ending with one- and
        two-wave chunks to shorten the un-overlapped first copy in / last copy out was measured SLOWER without graphs,
        10.1 vs 9.3 ms per 2^20 rows — a one-wave chunk is 0.26 ms of kernels, less than the host needs to issue it.)"""
        for lo in range(0, B, chunk):
            yield lo, min(B, lo + chunk)

    def _issue(self, body, B, chunk):
        """Fan out to the ring of streams, ``body(lo, hi)`` per chunk, fan in; the kernels' status flag comes back
        with the results."""
        from . import engine
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for n, (lo, hi) in enumerate(self._ranges(B, chunk)):
            with torch.cuda.stream(self.streams[n % len(self.streams)]):
                body(lo, hi)
        for s in self.streams:
            cur.wait_stream(s)
        self._status.copy_(engine.pipeline_status(self.device), non_blocking=True)

    def _launch(self, key, body, B):
        """``key`` None: always eager."""
        with torch.cuda.device(self.device):
            use_graph = self.use_graph and key is not None
            g = self._graphs.get(key) if use_graph else None
            if g is not None:
                g.replay()
            elif use_graph and key in self._warm:
                # second call with this key: everything lazy (packs, kernel attributes, allocator pools) is warm
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                # thread_local: other threads of the process (the NCCL watchdog of a multi-GPU job polls events) may
                # make CUDA calls while this thread captures
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._issue(body, B, self.chunk)
                while len(self._graphs) >= 4:          # every graph owns the device buffers of its chunks
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = g
                g.replay()
            else:
                if use_graph:
                    if len(self._warm) > 64:
                        self._warm.clear()
                    self._warm.add(key)
                self._issue(body, B, self.chunk if use_graph else self.chunk_eager)
            torch.cuda.current_stream(self.device).synchronize()      # the results are host memory: return them complete
        if int(self._status[0]) != 0:
            from . import _lib
            raise _lib.BgxError("tensor-core coupling kernel reported an internal pipeline timeout")

    @torch.no_grad()
    def run(self, z_host):
        """z_host: ``[B, dim_in]`` fp32 host tensor (pinned for asynchronous copies).
        Returns pinned host views ``(x [B, dim_out], dlogp [B, 1])``, valid until the next call."""
        B = z_host.shape[0]
        if B > self.out.shape[0] or z_host.shape[1] != self.dim_in:
            raise ValueError("input does not fit the pipeline's buffers")

        def body(lo, hi):
            x = z_host[lo:hi].to(self.device, non_blocking=True)
            y, d = self.flow(x, inverse=self.inverse)
            self.out[lo:hi].copy_(y, non_blocking=True)
            self.dlogp[lo:hi].copy_(d, non_blocking=True)

        # a graph holds the ADDRESS of the host buffer: key on it (and only capture pinned, contiguous inputs)
        key = ("run", z_host.data_ptr(), B, self._flow_version()) if z_host.is_pinned() and z_host.is_contiguous() else None
        self._launch(key, body, B)
        return self.out[:B], self.dlogp[:B]

    @torch.no_grad()
    def sample(self, n_samples, temperature=1.0):
        """``BoltzmannGenerator.sample(n, with_dlogp=True[, with_energy=True])`` (bg.py:105-135) into pinned host
        buffers: prior drawn on the device, flow, device->host copies of ``x``, ``dlogp`` (and the generator
        energy ``prior.energy(z) + dlogp``) — no host->device traffic.  Returns host views."""
        if self.prior is None:
            raise ValueError("HostPipeline.sample needs a prior (a module with .sample(n) on the device)")
        B = int(n_samples)
        if B > self.out.shape[0]:
            raise ValueError("n_samples does not fit the pipeline's buffers")

        def body(lo, hi):
            z = self.prior.sample(hi - lo, temperature=temperature)
            y, d = self.flow(z, temperature=temperature)
            self.out[lo:hi].copy_(y, non_blocking=True)
            self.dlogp[lo:hi].copy_(d, non_blocking=True)
            if self.with_energy:
                self.energy[lo:hi].copy_(self.prior.energy(z, temperature=temperature) + d, non_blocking=True)

        # (the torch generator is graph-safe: every replay draws new samples)
        self._launch(("sample", B, float(temperature), self._flow_version()), body, B)
        if self.with_energy:
            return self.out[:B], self.dlogp[:B], self.energy[:B]
        return self.out[:B], self.dlogp[:B]
