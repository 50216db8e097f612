"""Evaluate a flow on HOST-resident samples with the copies overlapped with the kernels.

The reference keeps trajectories on the host and feeds the mapped coordinates to an external potential
(tfep/app/base.py:790-797), so in practice x arrives from, and y / log_det_J return to, host memory.  The
batch is cut into chunks that travel through three CUDA streams -- host->device copy, flow kernels,
device->host copy -- so that PCIe transfers in both directions hide behind the compute of neighbouring chunks.
"""

import torch


class HostPipeline:
    """Reusable pinned / device staging buffers for ``flow`` evaluated on batches of ``batch`` samples."""

    def __init__(self, flow, batch, n_features, device, n_chunks=4, dtype=torch.float32):
        self.flow, self.device = flow, torch.device(device)
        self.bounds = [(i * batch // n_chunks, (i + 1) * batch // n_chunks) for i in range(n_chunks)]
        self.bounds = [(a, b) for a, b in self.bounds if b > a]
        self.x_dev = torch.empty(batch, n_features, dtype=dtype, device=self.device)
        self.y_host = torch.empty(batch, n_features, dtype=dtype).pin_memory()
        self.ld_host = torch.empty(batch, dtype=dtype).pin_memory()
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)

    def __call__(self, x_host, inverse=False):
        """x_host: pinned (batch, n_features) tensor.  Returns pinned ``(y_host, ld_host)``; the copies are
        complete when the call returns (the current stream has been joined with the output stream)."""
        main = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(main)
        self.s_out.wait_stream(main)
        fn = self.flow.inverse if inverse else self.flow
        with torch.no_grad():
            for a, b in self.bounds:
                with torch.cuda.stream(self.s_in):
                    self.x_dev[a:b].copy_(x_host[a:b], non_blocking=True)
                    ready = self.s_in.record_event()
                main.wait_event(ready)
                y, ld = fn(self.x_dev[a:b])
                done = main.record_event()
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(done)
                    self.y_host[a:b].copy_(y, non_blocking=True)
                    self.ld_host[a:b].copy_(ld, non_blocking=True)
                    y.record_stream(self.s_out)
                    ld.record_stream(self.s_out)
        main.wait_stream(self.s_out)
        return self.y_host, self.ld_host
