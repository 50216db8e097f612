"""Evaluate a flow on HOST-resident samples with the copies overlapped with the kernels.

The reference keeps trajectories on the host and feeds the mapped coordinates to an external potential
(tfep/app/base.py:790-797), so in practice x arrives from, and y / log_det_J return to, host memory.  The
batch is cut into chunks that travel through three CUDA streams -- host->device copy, flow kernels,
device->host copy -- so that PCIe transfers in both directions hide behind the compute of neighbouring chunks;
with ``wait=False`` consecutive batches overlap as well (``depth`` staging sets on both sides): the upload
of batch g + 1 and the download of batch g - 1 run while batch g is in the kernels.
"""

import contextlib
import os

import torch


@contextlib.contextmanager
def gpu_numa_affinity(device):
    """While active, the calling thread runs on the CPU cores closest to ``device`` (NVML's ideal affinity: the NUMA node
    the GPU hangs off), so that pinned staging buffers allocated inside are local to the GPU's PCIe root.  The previous
    affinity is restored on exit; without NVML (or permission) this does nothing."""
    previous = None
    try:
        import pynvml
        pynvml.nvmlInit()
        index = torch.device(device).index or 0
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        if visible:
            entry = visible.split(',')[index].strip()
            handle = pynvml.nvmlDeviceGetHandleByUUID(entry) if entry.startswith('GPU-') else \
                pynvml.nvmlDeviceGetHandleByIndex(int(entry))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        previous = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
    except Exception:                      # no NVML, no permission, exotic topology: keep the inherited affinity
        previous = None
    try:
        yield
    finally:
        if previous is not None:
            os.sched_setaffinity(0, previous)


class FEPWorkConsumer:
    """On-device consumer of a mapped batch for (T)FEP: ``work = u_B(y) / kT - log|det J| (- u_A(x) / kT)`` (reference
    docs/intro_to_MTFEP.ipynb:563-569) and its estimator partial ``(max, sum exp)`` of ``-work``
    (tfep_b200.analysis.estimator.lse_partial), so that only 4 bytes per sample -- the generalized work the estimator and
    the bootstrap consume -- and one 16-byte pair travel back to the host instead of the mapped coordinates.
    ``potential(y) -> (batch,)`` stands for the target potential (synthetic analytic here: external engines are out of
    scope); ``reference_potential(x)`` is optional."""

    def __init__(self, potential, kT=1.0, reference_potential=None):
        self.potential, self.kT, self.reference_potential = potential, kT, reference_potential

    def __call__(self, x, y, log_det_J):
        from ..analysis.estimator import lse_partial
        work = self.potential(y) / self.kT - log_det_J
        if self.reference_potential is not None:
            work = work - self.reference_potential(x) / self.kT
        return work, lse_partial(work)


def harmonic_potential(k=1.0, mu=0.5):
    """u(y) = 1/2 sum_f k (y_f - mu)^2: the synthetic target potential of BASELINE.json (north_star)."""
    def u(y):
        return 0.5 * k * ((y - mu) ** 2).sum(dim=1)
    return u


class HostPipeline:
    """Reusable pinned / device staging buffers for ``flow`` evaluated on batches of ``batch`` samples.

    ``consumer(x_dev, y_dev, log_det_J_dev) -> tuple of device tensors``: what is downloaded instead of
    ``(y, log_det_J)`` (CUDA-graph mode, :meth:`step_graph`); e.g. :class:`FEPWorkConsumer` when the mapped samples feed
    the free-energy estimator and only the work values are needed on the host."""

    def __init__(self, flow, batch, n_features, device, n_chunks=4, dtype=torch.float32, depth=3, consumer=None):
        self.flow, self.device = flow, torch.device(device)
        self.consumer = consumer
        self.out_hosts = None               # per staging set: pinned buffers of the consumer's outputs
        self.bounds = [(i * batch // n_chunks, (i + 1) * batch // n_chunks) for i in range(n_chunks)]
        self.bounds = [(a, b) for a, b in self.bounds if b > a]
        # `depth` staging sets: a batch occupies upload, kernels and download one after the other, so three batches
        # in flight are needed to keep the two copy engines and the SMs busy at the same time
        self.depth = depth
        self.x_dev = [torch.empty(batch, n_features, dtype=dtype, device=self.device) for _ in range(depth)]
        with gpu_numa_affinity(self.device):
            self.y_hosts = [torch.empty(batch, n_features, dtype=dtype).pin_memory() for _ in range(depth)]
            self.ld_hosts = [torch.empty(batch, dtype=dtype).pin_memory() for _ in range(depth)]
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.generation = 0
        self._computed = [None] * depth      # event: kernels that read x_dev[i] have finished
        self._downloaded = [None] * depth    # event: y_hosts[i] / ld_hosts[i] are complete

    @property
    def y_host(self):
        """Output buffers of the most recent call."""
        return self.y_hosts[(self.generation - 1) % self.depth]

    @property
    def ld_host(self):
        return self.ld_hosts[(self.generation - 1) % self.depth]

    def __call__(self, x_host, inverse=False, wait=True):
        """x_host: pinned (batch, n_features) tensor.  Returns pinned ``(y_host, ld_host)``.

        ``wait=True``: the current stream has been joined with the output stream when the call returns (results
        valid after a synchronize of the current stream).  ``wait=False``: nothing is joined, so the next call
        may start uploading while this batch computes and downloads; results of this call are valid after
        :meth:`join` (after ``depth`` more calls the buffers are reused)."""
        if self.consumer is not None:
            raise NotImplementedError('a consumer is applied in CUDA-graph mode: use step_graph()')
        main = torch.cuda.current_stream(self.device)
        g = self.generation % self.depth
        self.generation += 1
        x_dev, y_host, ld_host = self.x_dev[g], self.y_hosts[g], self.ld_hosts[g]
        if self._computed[g] is not None:
            self.s_in.wait_event(self._computed[g])          # staging buffer free again
        if self._downloaded[g] is not None:
            self.s_out.wait_event(self._downloaded[g])
        fn = self.flow.inverse if inverse else self.flow
        with torch.no_grad():
            for a, b in self.bounds:
                with torch.cuda.stream(self.s_in):
                    x_dev[a:b].copy_(x_host[a:b], non_blocking=True)
                    ready = self.s_in.record_event()
                main.wait_event(ready)
                y, ld = fn(x_dev[a:b])
                done = main.record_event()
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(done)
                    y_host[a:b].copy_(y, non_blocking=True)
                    ld_host[a:b].copy_(ld, non_blocking=True)
                    y.record_stream(self.s_out)
                    ld.record_stream(self.s_out)
        self._computed[g] = done
        self._downloaded[g] = self.s_out.record_event()
        if wait:
            main.wait_stream(self.s_out)
        return y_host, ld_host

    def join(self):
        """Make the current stream wait for every download issued so far."""
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self.s_out)
        for st in getattr(self, '_graph_streams', ()):
            main.wait_stream(st)

    # -- CUDA-graph mode: the whole step (upload, kernels, download) replayed with one launch -----------------
    def step_graph(self, x_host, inverse=False):
        """Same work as ``__call__(x_host, wait=False)`` with the host cost of ONE graph launch per batch.

        One graph per staging buffer set is captured on its own stream and they are replayed in turn, so the
        upload / kernels / download of consecutive batches overlap while each batch stays ordered on its own
        stream.  ``x_host`` must be the same pinned tensor at every call (its address is part of the graph; a
        different tensor triggers a re-capture).  Results: ``y_host`` / ``ld_host`` after :meth:`join`."""
        key = (x_host.data_ptr(), bool(inverse))
        if getattr(self, '_graph_key', None) != key:
            self._capture(x_host, inverse)
            self._graph_key = key
        g = self.generation % self.depth
        self.generation += 1
        with torch.cuda.stream(self._graph_streams[g]):
            self._graphs[g].replay()
        if self.consumer is not None:
            return tuple(self.out_hosts[g])
        return self.y_hosts[g], self.ld_hosts[g]

    @property
    def last_stream(self):
        """The stream the most recent :meth:`step_graph` call was replayed on (an event recorded there fires when that
        step's download has finished)."""
        return self._graph_streams[(self.generation - 1) % self.depth]

    @property
    def outputs_host(self):
        """Pinned outputs of the most recent :meth:`step_graph` call (valid after :meth:`join`)."""
        g = (self.generation - 1) % self.depth
        return tuple(self.out_hosts[g]) if self.consumer is not None else (self.y_hosts[g], self.ld_hosts[g])

    def _capture(self, x_host, inverse):
        fn = self.flow.inverse if inverse else self.flow
        main = torch.cuda.current_stream(self.device)
        self._graph_streams = [torch.cuda.Stream(self.device) for _ in range(self.depth)]
        self._graphs = []
        with torch.no_grad():
            y, ld = fn(self.x_dev[0])                 # warm-up outside capture: packs weights, allocates workspaces
            if self.consumer is not None:
                outs = self.consumer(self.x_dev[0], y, ld)
                with gpu_numa_affinity(self.device):
                    self.out_hosts = [[torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
                                      for _ in range(self.depth)]
        torch.cuda.synchronize(self.device)
        for g in range(self.depth):
            st = self._graph_streams[g]
            st.wait_stream(main)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(st), torch.no_grad():
                with torch.cuda.graph(graph, stream=st):
                    self.x_dev[g].copy_(x_host, non_blocking=True)
                    y, ld = fn(self.x_dev[g])
                    if self.consumer is not None:
                        for host, out in zip(self.out_hosts[g], self.consumer(self.x_dev[g], y, ld)):
                            host.copy_(out, non_blocking=True)
                    else:
                        self.y_hosts[g].copy_(y, non_blocking=True)
                        self.ld_hosts[g].copy_(ld, non_blocking=True)
            self._graphs.append(graph)
        torch.cuda.synchronize(self.device)
