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
    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)

    def put(self, host_batch):
        """Starts the copy of a dict of (pinned) host tensors; returns a handle for ``get``."""
        with torch.cuda.stream(self.stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in host_batch.items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def get(self, handle):
        """Makes the current stream wait for the copy and hands the device tensors over to it."""
        dev, ev = handle
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev.values():
            t.record_stream(cur)
        return dev
