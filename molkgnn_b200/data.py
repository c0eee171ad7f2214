"""Host -> device staging of collated batches.

The reference feeds ``MolGCN`` from a PyG ``DataLoader`` (``entry.py`` / ``data.py``: pinned memory, worker processes);
here the only part of that on the hot path is the copy of the four collated tensors (x, p, edge_index, edge_attr) from
pinned host memory to the GPU.  ``DevicePrefetcher`` issues that copy on a side stream, so the inputs of step i+1 cross
PCIe/NVLink-C2C while the kernels of step i run; torch is only the allocator and the stream/event plumbing.
"""
from __future__ import annotations

import torch

BATCH_KEYS = ("x", "p", "edge_index", "edge_attr")


class DevicePrefetcher(object):
    """Stages the NEXT batch while the current one computes: the host -> device copies of its tensors and (optionally) its
    degree-bucket plan run on a side stream.  The plan's one host round trip (bucket sizes) then waits only for that
    side stream -- whose work finished a step ago -- instead of draining the compute stream, so the host can queue steps
    ahead of the GPU."""

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        # `depth` batches may be in flight (put() called `depth` steps ahead of get(), like a DataLoader's prefetch_factor):
        # one side stream per slot, because finishing a plan synchronises the host with ITS stream -- with one stream the host
        # would also wait for the copies of the batch staged after it.  With depth 1 the host waits ~0.4 ms per step for the
        # 22 MB copy that could only start when the step before the previous one had finished.
        # (Each side stream has its own pool in torch's caching allocator: the first few batches of a run allocate their staging
        # and plan buffers with cudaMalloc -- ~2 ms each, once per stream and size; a loop that is timed from its first step
        # should have run a handful of steps before, cf. bench.py.)
        self.depth = max(1, int(depth))
        self.streams = [torch.cuda.Stream(self.device) for _ in range(self.depth)]
        self.stream = self.streams[0]
        self._n = 0

    def put(self, host_batch=None, device_batch=None, build_plan=False, store_batch=None):
        """Starts staging a batch given as a dict of (pinned) host tensors, of tensors already on the device, or as
        ``store_batch=(MoleculeStore, ids)`` (assembled on the GPU from the HBM-resident store: the only host->device copy is
        the id list); returns a handle for ``get``.  With ``build_plan`` the counting half of the GPU bucket pass is queued
        behind the copies."""
        from .plan import BucketPlan
        # Memory protocol: everything staged here is allocated from the side stream's pool and later used on the compute
        # stream WITHOUT record_stream (recorded blocks come back late, the pool keeps growing with synchronising
        # cudaMallocs).  Instead the side stream first waits for everything queued on the compute stream so far: a block
        # freed by the host (its last compute-stream use was queued before this point) is then safe to reuse here.
        stream = self.streams[self._n % self.depth]
        self._n += 1
        stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(stream):
            if store_batch is not None:
                dev = store_batch[0].collate(store_batch[1])
            else:
                dev = device_batch if host_batch is None else {k: v.to(self.device, non_blocking=True)
                                                                for k, v in host_batch.items()}
            plan = None
            if build_plan:
                plan = BucketPlan.begin_from_edge_index(dev["edge_index"], dev["p"], dev["edge_attr"], dev["x"].shape[0])
            ev = torch.cuda.Event()
            ev.record(stream)
        return dev, ev, plan, host_batch is not None or store_batch is not None, stream

    def get(self, handle):
        """Finishes the plan (if any), makes the current stream wait for the staged batch and hands the tensors over to it.
        Returns the device tensors, or (tensors, plan) if the handle carries a plan."""
        dev, ev, plan, owned, stream = handle
        cur = torch.cuda.current_stream(self.device)
        if plan is not None:
            with torch.cuda.stream(stream):
                plan.finish()                      # bucket sizes to the host + assignment kernels, still on the side stream
                ev = torch.cuda.Event()
                ev.record(stream)
        cur.wait_event(ev)
        if plan is not None:
            return dev, plan
        return dev
