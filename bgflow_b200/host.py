"""Host-buffer entry point: run a flow on data that lives in (pinned) host memory.

``HostPipeline.run(z_host)`` is the call ``bench.py`` times for its end-to-end number: the
batch is cut into chunks that are copied host->device, pushed through the flow and copied back
device->host on a small ring of CUDA streams, so the PCIe transfers of neighbouring chunks
overlap the kernels of the current one (the reference does ``x.to(device)`` / ``.cpu()`` around a
monolithic call, bgflow/bg.py:115-117).
"""

import torch

__all__ = ["HostPipeline"]


class HostPipeline:
    def __init__(self, flow, dim_in, dim_out, max_rows, device, chunk_rows=1 << 17, n_streams=3,
                 inverse=False):
        self.flow = flow
        self.device = torch.device(device)
        self.chunk = int(chunk_rows)
        self.inverse = inverse
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self.out = torch.empty(max_rows, dim_out, dtype=torch.float32).pin_memory()
        self.dlogp = torch.empty(max_rows, 1, dtype=torch.float32).pin_memory()
        self.dim_in = dim_in

    @torch.no_grad()
    def run(self, z_host):
        """z_host: ``[B, dim_in]`` fp32 host tensor (pinned for asynchronous copies).
        Returns pinned host views ``(x [B, dim_out], dlogp [B, 1])``, valid until the next call."""
        B = z_host.shape[0]
        if B > self.out.shape[0] or z_host.shape[1] != self.dim_in:
            raise ValueError("input does not fit the pipeline's buffers")
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for n, lo in enumerate(range(0, B, self.chunk)):
            hi = min(B, lo + self.chunk)
            s = self.streams[n % len(self.streams)]
            with torch.cuda.stream(s):
                x = z_host[lo:hi].to(self.device, non_blocking=True)
                y, d = self.flow(x, inverse=self.inverse)
                self.out[lo:hi].copy_(y, non_blocking=True)
                self.dlogp[lo:hi].copy_(d, non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        cur.synchronize()          # the results are host memory: the call returns them complete
        return self.out[:B], self.dlogp[:B]
