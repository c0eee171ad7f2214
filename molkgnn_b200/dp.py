"""Molecule-sharded data parallelism for the conv stack.

Molecules are independent graphs, so the forward needs no communication: every rank runs the full kernels on its own
shard (parameters replicated).  The only exchange is ONE all-reduce per step of a flat fp32 bucket holding the
kernel-parameter gradients (0.49 MB for the base stack): NCCL over NVLink 5 / NVSwitch on GPUs, gloo in the CPU tests.
Parameters that never receive a gradient (p_support, length/angle weights) are not in the bucket.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

GRAD_PARAM_SUFFIXES = ("x_center", "x_support", "edge_attr_support", "support_attr_sc_weight",
                       "center_attr_sc_weight", "edge_attr_support_sc_weight")


def shard_bounds(num_nodes_per_molecule, world):
    """Contiguous molecule ranges balanced by atom count -> list of (first_molecule, last_molecule_exclusive)."""
    import numpy as np
    sizes = np.asarray(num_nodes_per_molecule, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(sizes)])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, total * r / world, side="left")))
    cuts.append(len(sizes))
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


class OneShotAllReduce(object):
    """One-shot all-reduce over NVLink peer memory (csrc/oneshot.cu): every rank reads the other ranks' copies of the flat
    gradient buffer straight out of their exchange buffers and sums them in rank order -- one kernel per step instead of the
    2 (W - 1) hops of a ring, for a buffer that is latency sized.  ``create`` returns None (on EVERY rank) when the exchange
    cannot be set up on some rank (no peer access, IPC refused), so that the caller falls back to NCCL consistently."""

    def __init__(self, handle, world, cap_floats):
        self.handle, self.world, self.cap_floats = handle, world, cap_floats

    @classmethod
    def create(cls, numel, group=None):
        import ctypes as C
        from . import _lib
        if not (dist.is_initialized() and torch.cuda.is_available()):
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or world > 16:
            return None
        dev = torch.device("cuda", torch.cuda.current_device())
        L = _lib.lib()
        handle, mine, ok = C.c_void_p(), torch.zeros(64, dtype=torch.uint8), 1
        # the slot size must be the same on every rank (validated collectively, once)
        sizes = torch.tensor([int(numel), -int(numel)], dtype=torch.int64, device=dev)
        dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=group)
        if int(sizes[0].item()) != -int(sizes[1].item()):
            return None
        try:
            _lib.check(L.molkgnn_oneshot_create(rank, world, 4 * int(numel) + 64, C.byref(handle)))
            buf = (C.c_ubyte * 64)()
            _lib.check(L.molkgnn_oneshot_ipc_handle(handle, buf))
            mine = torch.tensor(list(buf), dtype=torch.uint8)
        except Exception:
            ok = 0
        # every rank learns every handle (and whether every rank got this far)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        gathered = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(gathered, mine.to(dev), group=group)
        if int(flag.item()) == 1:
            try:
                allh = torch.stack(gathered).cpu().contiguous()
                _lib.check(L.molkgnn_oneshot_open(handle, C.c_void_p(allh.data_ptr())))
            except Exception:
                ok = 0
        flag = torch.tensor([ok if int(flag.item()) == 1 else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) != 1:
            if handle.value:
                L.molkgnn_oneshot_destroy(handle)
            return None
        return cls(handle, world, int(numel) + 16)

    def allreduce(self, flat, average=True):
        from . import _lib
        if flat.numel() > self.cap_floats or flat.dtype != torch.float32 or not flat.is_contiguous():
            # rank-local failure: the peers time out on this rank's flag, poison their buffers with NaN and raise at their next
            # check() -- they never reduce without this rank.  create() validated the slot size collectively, so with the same
            # model on every rank this cannot be reached on one rank only.
            raise _lib.MolKGNNError("OneShotAllReduce: buffer larger than the exchange slot or not a contiguous fp32 tensor")
        _lib.check(_lib.lib().molkgnn_oneshot_allreduce(self.handle, _lib.ptr(flat), flat.numel(), 1 if average else 0,
                                                        _lib.stream_ptr()))
        return flat

    def check(self):
        """raises if a peer's flag timed out in some step since the last check (synchronises the device)"""
        from . import _lib
        e = _lib.lib().molkgnn_oneshot_error(self.handle)
        if e:
            raise _lib.MolKGNNError(f"OneShotAllReduce: the flag of rank {e - 1} did not arrive within the time limit")

    def check_async(self):
        """The periodic check of the step loop, without a device synchronisation: queues a copy of the error word into pinned host
        memory behind the work on the current stream and examines the copy queued by the PREVIOUS call (its event completed
        long ago).  An error therefore surfaces one period late -- the gradients were NaN-poisoned in the meantime."""
        from . import _lib
        prev = self.__dict__.get("_pending_err")
        if prev is not None:
            buf, ev = prev
            ev.synchronize()
            e = int(buf.item())
            if e:
                _lib.lib().molkgnn_oneshot_error(self.handle)       # (clears the word)
                raise _lib.MolKGNNError(f"OneShotAllReduce: the flag of rank {e - 1} did not arrive within the time limit")
        else:
            buf = torch.zeros(1, dtype=torch.int32).pin_memory()
        _lib.check(_lib.lib().molkgnn_oneshot_error_async(self.handle, buf.data_ptr(), _lib.stream_ptr()))
        ev = torch.cuda.Event()
        ev.record()
        self._pending_err = (buf, ev)

    def close(self):
        from . import _lib
        if self.handle is not None and self.handle.value:
            _lib.lib().molkgnn_oneshot_destroy(self.handle)
            self.handle = None


def _native_flat_of(module):
    flats = [m._stack.last_grad_flat for m in module.modules()
             if getattr(m, "_stack", None) is not None and getattr(m._stack, "last_grad_flat", None) is not None]
    return flats[0] if len(flats) == 1 else None


class GradBucket(object):
    """Flat all-reduce bucket over the gradients of the kernel parameters of a module."""

    def __init__(self, module, world=None, average=True, group=None, oneshot=None):
        self.params = [p for n, p in module.named_parameters()
                       if p.requires_grad and n.rsplit(".", 1)[-1] in GRAD_PARAM_SUFFIXES]
        self.numel = sum(p.numel() for p in self.params)
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.average = average
        self.group = group
        self.flat = None
        self.module = module
        # one-shot NVLink all-reduce of the native flat buffer: the default on NCCL process groups since the 8-GPU measurement
        # (profiles/r02j_bench_8gpu*.json, 4096 molecules per GPU: 1.376 ms per step against 1.539 ms with ncclAllReduce, i.e.
        # 93 % against 83 % of 8x the one-GPU rate; tools/dp_check.py: sharded gradients == full-batch gradients on 8 GPUs for
        # both paths).  MOLKGNN_DP_ONESHOT=0 / oneshot=False selects NCCL, which is also the fallback whenever the peer-memory
        # exchange cannot be set up on some rank.
        import os
        if oneshot is None:
            oneshot = os.environ.get("MOLKGNN_DP_ONESHOT", "1") != "0"
        self.oneshot = None
        self.check_every = int(os.environ.get("MOLKGNN_DP_CHECK_EVERY", "64"))
        self._steps = 0
        # the one-shot kernel is a latency play: every rank reads EVERY rank's copy (world x the bytes of a ring all-reduce), so
        # large buckets (the wide model: 12.7 MB of kernel gradients) go through NCCL
        max_bytes = int(os.environ.get("MOLKGNN_DP_ONESHOT_MAX_BYTES", str(4 << 20)))
        if 4 * self.numel > max_bytes:
            oneshot = False
        if oneshot and self.world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl":
            self.oneshot = OneShotAllReduce.create(2 * self.numel + 4096, group)   # room for the native buffer's padding

    def check(self):
        """Raises if the one-shot exchange ever timed out (a rank that died or raised before its launch leaves the others
        with NaN-poisoned gradients, csrc/oneshot.cu).  Synchronises the device: call it where the loop synchronises anyway
        (logging, before a checkpoint); ``allreduce`` itself polls the error word every ``check_every`` steps without synchronising
        (``OneShotAllReduce.check_async``: an error surfaces one period late)."""
        if self.oneshot is not None:
            self.oneshot.check()

    def _native_flat(self, ps):
        flat = _native_flat_of(self.module)
        if flat is None or len(ps) != len(self.params):
            return None
        lo = flat.data_ptr()
        hi = lo + flat.numel() * flat.element_size()
        # every p.grad must still be the view the backward handed out (autograd keeps it when .grad was None before); the
        # views are laid out in parameter order, so the first and the last one bracket the rest
        for p in (ps[0], ps[-1], ps[len(ps) // 2]):
            a = p.grad.data_ptr()
            if p.grad.device != flat.device or a < lo or a + p.grad.numel() * 4 > hi:
                return None
        return flat

    def allreduce(self):
        """sum (or mean) the gradients over ranks; afterwards every p.grad is a view into the reduced flat bucket"""
        ps = [p for p in self.params if p.grad is not None]
        if not ps:
            return None
        # fast path: the native backward wrote every kernel-parameter gradient of the stack into ONE flat buffer and the
        # p.grad tensors are views of it (functional.MolGCNFn.backward) -> reduce that buffer in place, no gather/scatter
        flat = self._native_flat(ps)
        if flat is not None:
            if self.world > 1 and self.oneshot is not None and flat.is_cuda:
                self.oneshot.allreduce(flat, self.average)
                self._steps += 1
                if self.check_every > 0 and self._steps % self.check_every == 0:
                    self.oneshot.check_async()      # no device synchronisation inside the step loop
            elif self.world > 1:
                if self.average and flat.is_cuda:
                    # NCCL averages inside the collective: no separate scaling kernel behind it
                    dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
                else:
                    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                    if self.average:
                        flat.div_(self.world)
            self.flat = flat
            return flat
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                flat.div_(self.world)
        off = 0
        for p in ps:
            n = p.numel()
            p.grad = flat[off:off + n].view_as(p)
            off += n
        self.flat = flat
        return flat
