#!/usr/bin/env python
"""Benchmark of the MolKGNN conv stack (bucket pass + conv/propagate fwd + bwd) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--molecules B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic molecules: GPU degree-bucket pass, MolGCN forward
(3 layers, kernels 10/20/30/50, x 28-dim, bond 7-dim) and backward down to grad_x and all kernel-parameter gradients,
+ NCCL all-reduce of the flat kernel-gradient bucket when N > 1 (molecules sharded per rank, weak scaling).
Workload = BASELINE.json configs[1]: batch 4096 molecules per GPU.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own unmodified modules (oracle/_ref, staged by
tools/make_oracle_ref.py; kind "reference") on the host cores -- the oracle port (kind "port") only where they are not staged.
`--molecules 65536` = one GPU's share of BASELINE configs[3]; `--forward-only` = configs[4] (inference sweep, h returned to the host).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_BASE = (10, 20, 30, 50)
NUM_LAYERS = 3
X_DIM, EDGE_DIM = 28, 7
METRIC = "molecules/sec fwd+bwd MolKGNN conv (1/2/4/8 B200, % HBM roofline) vs host CPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--molecules", type=int, default=4096, help="molecules per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=None, help="synthetic-batch seed (default 1000 * rank, SURVEY 8(d))")
    ap.add_argument("--wide", action="store_true",
                    help="BASELINE configs[2]: wide kernel sets 40/80/120/200 per degree, 5 layers (the streamed-operand tcgen05 kernels)")
    ap.add_argument("--forward-only", action="store_true",
                    help="BASELINE configs[4] (inference sweep): bucket pass + forward under torch.no_grad; e2e returns h [N,K] to the host")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules (oracle/_ref, staged by tools/make_oracle_ref.py) on the host cores; where they
# are not staged, the oracle port of the same algorithm
# ---------------------------------------------------------------------------------------------------------------------
def cpu_arm(kind=None, batch=16, seed=123):
    """-> (kind, one(i), cores): `one(i)` runs MolGCN fwd + backward on the i-th README-shape mini-batch (16 molecules).

    kind "reference": the UNMODIFIED reference modules (models/MolKGNN/{kernels,KernelLayer}.py, imported from oracle/_ref
    through the torch_geometric stand-in of tests/stubs), exactly BASELINE.md 2.  The 20 per-degree tensors the reference
    precomputes OFFLINE (pre_transform, data.py:19-30) are built outside the timed region.
    kind "port": oracle/molkgnn_oracle.py (vectorised restatement, ~35x faster than the reference's Python loops).
    The reference is super-linear in batch size (O(n4^2) chirality loop, ~30 MB/molecule of autograd state), so
    batch 16 is the configuration it is quoted on and the only one that is feasible on a host."""
    from molkgnn_b200 import synth
    from oracle import molkgnn_oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_oracle_ref
    if kind is None:
        kind = "reference" if make_oracle_ref.available() else "port"
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    pool = [synth.make_batch(batch, seed=seed + i) for i in range(8)]
    if kind == "reference":
        import warnings
        warnings.filterwarnings("ignore", message="Using torch.cross without specifying the dim")
        MolGCN = make_oracle_ref.load()["KernelLayer"].MolGCN
        net = MolGCN(num_layers=NUM_LAYERS, num_kernel1_1hop=L_BASE[0], num_kernel2_1hop=L_BASE[1], num_kernel3_1hop=L_BASE[2],
                     num_kernel4_1hop=L_BASE[3], num_kernel1_Nhop=L_BASE[0], num_kernel2_Nhop=L_BASE[1],
                     num_kernel3_Nhop=L_BASE[2], num_kernel4_Nhop=L_BASE[3], x_dim=X_DIM, p_dim=3, edge_attr_dim=EDGE_DIM)
        kws = []
        for b in pool:
            bk = orc.bucket_pass(b["edge_index"], b["x"].shape[0], b["p"], b["edge_attr"])    # offline in the reference
            kw = dict(edge_index=torch.from_numpy(b["edge_index"]), edge_attr=torch.from_numpy(b["edge_attr"]),
                      p=torch.from_numpy(b["p"]), save_score=False)
            for d in range(1, 5):
                for k, v in bk[d].items():
                    kw[f"{k}_deg{d}"] = torch.from_numpy(v)
            kws.append(kw)

        def one(i):
            b, kw = pool[i % len(pool)], kws[i % len(pool)]
            x = torch.from_numpy(b["x"]).clone().requires_grad_(True)
            h = net(x=x, **kw)
            h.sum().backward()
            net.zero_grad(set_to_none=True)
    else:
        params = orc.init_molgcn_params(NUM_LAYERS, L_BASE, L_BASE, X_DIM, requires_grad=True)

        def one(i):
            b = pool[i % len(pool)]
            N = b["x"].shape[0]
            bk = orc.buckets_to_torch(orc.bucket_pass(b["edge_index"], N, b["p"], b["edge_attr"]))
            x = torch.from_numpy(b["x"]).clone().requires_grad_(True)
            h = orc.molgcn_forward(params, x, torch.from_numpy(b["edge_index"]), bk)
            h.sum().backward()
    return kind, one, cores


def cpu_stream(seconds=None, batches=None, batch=16, warm=2, kind=None):
    """-> (molecules, seconds, cores, kind) of a stream of batch-16 mini-batches through the CPU arm"""
    kind, one, cores = cpu_arm(kind, batch)
    for i in range(warm):
        one(i)
    t0 = time.perf_counter()
    n = 0
    while True:
        one(n)
        n += 1
        el = time.perf_counter() - t0
        if (batches is not None and n >= batches) or (seconds is not None and el >= seconds):
            break
    return n * batch, el, cores, kind


CPU_SAMPLE = {"reference": "UNMODIFIED reference models/MolKGNN/{kernels,KernelLayer}.py (oracle/_ref) MolGCN fwd + autograd bwd, torch CPU, "
                           "all host threads",
              "port": "oracle/molkgnn_oracle.py (vectorised restatement; oracle/_ref not staged on this box), torch CPU, all host threads"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on all host threads.  A step = a bounded sample of
    the workload: 2 mini-batches of 16 molecules (the reference runs ~25 molecules/s, so 20 steps take ~30 s)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, one, cores = cpu_arm()
    per_step = 2 if kind == "reference" else 8
    for i in range(args.warmup * per_step):
        one(i)
    t0 = time.perf_counter()
    for i in range(args.steps * per_step):
        one(i)
    el = time.perf_counter() - t0
    v = args.steps * per_step * 16 / el
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "molecules/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.molecules, args.forward_only),
        "cpu_baseline": {"value": v, "unit": "molecules/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} steps x {per_step} mini-batches x 16 molecules (README batch shape), "
                                   + CPU_SAMPLE[kind]},
        "e2e": {"value": v, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(B, fwd_only=False):
    which = ("BASELINE configs[2] (wide kernel sets): MolGCN conv stack " + ("forward only" if fwd_only else "fwd+bwd")
             if tuple(L_BASE) != (10, 20, 30, 50) else
             "BASELINE configs[4] (inference sweep): MolGCN conv stack forward only" if fwd_only else
             "BASELINE configs[1]: MolGCN conv stack fwd+bwd" if B == 4096 else
             "BASELINE configs[3] (one GPU's share of the sharded training step): MolGCN conv stack fwd+bwd" if B == 65536 else
             "MolGCN conv stack fwd+bwd")
    return {"workload": f"{which}, {NUM_LAYERS} layers, kernels {'/'.join(str(v) for v in L_BASE)} "
                        f"(1-hop and N-hop), node_dim 28, edge_dim 7, batch {B} synthetic 3D molecules per GPU "
                        "(18-32 atoms, degrees 1-4), GPU degree-bucket pass included",
            "molecules_per_gpu": B, "layers": NUM_LAYERS, "kernels": list(L_BASE), "forward_only": bool(fwd_only),
            "cpu_arms": "the CPU arms (cpu_baseline, --impl reference) time a stream of batch-16 mini-batches of the same model and "
                        "generator: the reference is super-linear in batch size and cannot run this batch on a host (SURVEY 6)",
            "pipeline": "the GPU bucket pass of step i+2 is queued on a side stream while step i computes (one pass per step, "
                        "inside the timed region)",
            "l2": "no explicit flush: per-step working set (activations+gradients of 3 layers, ~0.5 GB) exceeds the 126 MB L2"}


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = 0.05
        for window in (0.05, 0.5):               # (a timed region shorter than nvidia-smi's period: the lines next to it, also under load)
            for ts, line in self.rows:
                if ts < t0 - window or ts > t1 + window:
                    continue
                f = [s.strip() for s in line.split(",")]
                try:
                    sm.append(float(f[0]))
                    mx = float(f[1])
                except Exception:
                    continue
                for nme, val in zip(names, f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            if sm:
                break
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window_s": window}


def main():
    args = parse()
    if args.wide:                                # configs[2]: the same protocol on the wide model
        global L_BASE, NUM_LAYERS
        L_BASE, NUM_LAYERS = (40, 80, 120, 200), 5
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE line (the JSON): until it is printed, file descriptor 1 points at stderr, so that banners
    # native libraries write there (NCCL prints "NCCL version ..." at NCCL_DEBUG=VERSION and =WARN) cannot precede it
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from molkgnn_b200 import build as mkbuild
    if rank == 0:
        mkbuild.build()
    if world > 1:
        dist.barrier()
    import molkgnn_b200 as mk
    from molkgnn_b200 import synth, roofline, _lib, functional as Fn
    from molkgnn_b200.dp import GradBucket

    B = args.molecules
    fwd_only = args.forward_only
    batch = synth.make_batch(B, seed=1000 * rank if args.seed is None else args.seed)   # shard of this rank
    N, E = batch["x"].shape[0], batch["edge_index"].shape[1]
    host = {k: torch.from_numpy(batch[k]).pin_memory() for k in ("x", "p", "edge_index", "edge_attr")}
    devt = {k: v.to(dev) for k, v in host.items()}
    torch.manual_seed(0)
    net = mk.MolGCN(NUM_LAYERS, *L_BASE, *L_BASE, x_dim=X_DIM, p_dim=3, edge_attr_dim=EDGE_DIM).to(dev)
    K = sum(L_BASE)
    wout = torch.randn(N, K, device=dev)
    bucket = GradBucket(net, world) if world > 1 else None

    from molkgnn_b200.data import DevicePrefetcher
    pf = DevicePrefetcher(dev)

    def step(t, plan=None):
        """bucket pass (here, or already staged one step ahead on the prefetcher's stream: `plan`) + fwd + bwd
        (+ all-reduce) on device tensors `t`"""
        if fwd_only:
            with torch.no_grad():
                return net(x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False, plan=plan)
        x = t["x"].detach().requires_grad_(True)
        h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False, plan=plan)
        h.backward(wout)                     # dL/dh handed in directly: no torch arithmetic inside the timed region
        if bucket is not None and not os.environ.get("MOLKGNN_BENCH_NO_ALLREDUCE"):   # (diagnostic switch)
            bucket.allreduce()
        return h

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    grad_params = [p for p in net.parameters()]

    def zero_grads():
        for p in grad_params:                    # net.zero_grad(set_to_none=True) without the module-tree walk
            p.grad = None

    def run_steps(k):
        """k steps on the device-resident batch.  Every step runs its own GPU bucket pass; it is queued one step ahead on
        the prefetcher's side stream, so its host round trip (bucket sizes) never drains the compute stream."""
        q = [pf.put(device_batch=devt, build_plan=True) for _ in range(min(pf.depth, k))]
        for i in range(k):
            t, plan = pf.get(q.pop(0))
            if i + pf.depth < k:
                q.append(pf.put(device_batch=devt, build_plan=True))
            step(t, plan)
            zero_grads()

    # nvidia-smi needs a few hundred ms before its first line: the sampler starts BEFORE the warm-up, so that it is streaming
    # (one line per 20 ms) when the ~0.2 s timed region begins; stop() keeps the lines inside the timed region
    sampler = ClockSampler(local) if rank == 0 else None
    # ---- warm-up (also: find the dominant kernel with the event profiler) ----
    step(devt)                                   # the unpipelined path once (plan built inside the forward)
    zero_grads()
    # (warm-up runs have an ODD number of steps, and there are two of them: the prefetcher alternates between two side streams, each
    # with its own pool in torch's caching allocator, and the start of a run needs one more cached segment on the stream its fourth
    # batch lands on -- after two odd runs both pools have it; otherwise the timed run pays a 2 ms cudaMalloc in its second step)
    run_steps(max(args.warmup, 3) | 1)
    Fn.profile_start()
    run_steps(3)
    breakdown = Fn.profile_stop()
    breakdown.pop("bucket_count", None)      # contains the host round trip of the bucket sizes, not a kernel time
    top = max(breakdown, key=lambda k: breakdown[k][1])
    # Host hygiene of a latency-bound loop: everything allocated so far (torch, the model, the batch) moves to the permanent
    # generation, so that a full garbage collection inside the timed region does not walk it -- with N ranks coupled by the
    # all-reduce, one rank's multi-millisecond collection pause is every rank's pause.
    import gc
    gc.collect()
    gc.freeze()

    # ---- timed region: device-resident inputs ----
    l0 = _lib.lib().molkgnn_launch_count()
    Fn.profile_start([top])
    sync_all()
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps)                        # K bucket passes + K forward/backward passes between the two events
    e1.record()
    sync_all()
    t_wall1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    top_timed = Fn.profile_stop()
    launches = (_lib.lib().molkgnn_launch_count() - l0) / args.steps
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end: host (pinned) -> device copies and a device -> host read of the result inside the timed region ----
    # Every step copies its own inputs from pinned host memory (K copies for K steps, all inside the timed region) and reads
    # its loss back; the copy of step i+1 is issued on the prefetcher's side stream before step i's loss is waited for, the
    # way a pinned-memory DataLoader feeds the reference's training loop.
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    # inference sweep (configs[4]): what returns to the host per step is the per-MOLECULE result, global_add_pool(h) [B, K]
    # (MolKGNNNet.py:144-146 pools before anything leaves the model), through the native deterministic segmented sum
    h_host = [torch.empty(B, K, dtype=torch.float32).pin_memory() for _ in range(2)] if fwd_only else None
    ptr_dev = torch.from_numpy(batch["ptr"]).to(dev)
    batch_dev = torch.from_numpy(batch["batch"]).to(dev)

    def e2e_run(k):
        q = [pf.put(host_batch=host, build_plan=True) for _ in range(min(pf.depth, k))]   # prefetch depth 2, as a DataLoader's
        pending = None
        for i in range(k):
            t, plan = pf.get(q.pop(0))
            if i + pf.depth < k:
                q.append(pf.put(host_batch=host, build_plan=True))
            h = step(t, plan)
            if fwd_only:                          # inference: the pooled result [B, K] returns to the host
                buf = h_host[i & 1]
                buf.copy_(mk.global_add_pool(h, batch_dev, ptr_dev), non_blocking=True)
            else:
                loss = (h.detach() * wout).sum()
                buf = loss_host[i & 1]
                buf.copy_(loss, non_blocking=True)    # D2H read of the step's result into pinned memory ...
            ev = torch.cuda.Event()
            ev.record()
            zero_grads()
            if pending is not None:               # ... consumed on the host one step later (asynchronous logging), so
                pending[1].synchronize()          # the host can queue step i+1 while step i still runs
                float(pending[0].view(-1)[0])
            pending = (buf, ev)
        pending[1].synchronize()
        float(pending[0].view(-1)[0])

    e2e_run(max(args.warmup, 3) | 1)             # two odd warm-up runs: both side-stream pools of the allocator are warm (see above)
    e2e_run(3)
    sync_all()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    # ---- end to end with the dataset RESIDENT in HBM (molkgnn_b200.store: packed molecule store + GPU batcher): the step's host
    # input is the list of molecule ids of the batch (pinned), the batch is assembled on the GPU -- what the 180 GB are for ----
    ms_store = None
    if not os.environ.get("MOLKGNN_BENCH_NO_STORE"):
        from molkgnn_b200.store import MoleculeStore
        ptr_h = batch["ptr"]
        eptr = np.searchsorted(batch["edge_index"][0], ptr_h)            # edges are grouped molecule by molecule
        local_ei = batch["edge_index"] - np.repeat(ptr_h[:-1], np.diff(eptr))[None, :]
        store = MoleculeStore(devt["x"], devt["p"], torch.from_numpy(local_ei).to(dev), devt["edge_attr"],
                              torch.from_numpy(ptr_h).to(dev), torch.from_numpy(eptr).to(dev))
        gen = torch.Generator().manual_seed(rank)
        id_bufs = [torch.randperm(B, generator=gen).pin_memory() for _ in range(4)]

        def store_run(k):
            pending = None
            # H2D: B int64 ids per step; gather + rebasing on the GPU
            q = [pf.put(store_batch=(store, id_bufs[j & 3]), build_plan=True) for j in range(min(pf.depth, k))]
            for i in range(k):
                bt, plan = pf.get(q.pop(0))
                if i + pf.depth < k:
                    q.append(pf.put(store_batch=(store, id_bufs[(i + pf.depth) & 3]), build_plan=True))
                h = step(bt, plan)
                if fwd_only:
                    buf = h_host[i & 1]
                    buf.copy_(mk.global_add_pool(h, bt["batch"], bt["ptr"]), non_blocking=True)
                else:
                    loss = (h.detach() * wout).sum()
                    buf = loss_host[i & 1]
                    buf.copy_(loss, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                zero_grads()
                if pending is not None:
                    pending[1].synchronize()
                    float(pending[0].view(-1)[0])
                pending = (buf, ev)
            pending[1].synchronize()
            float(pending[0].view(-1)[0])

        store_run(3)
        store_run(3)
        sync_all()
        e0.record()
        store_run(args.steps)
        e1.record()
        sync_all()
        ms_store = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_store], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_store = float(t.item())

    # ---- multi-GPU self-check (outside the timed regions): sharded CUDA gradients after the all-reduce == full-batch gradients ----
    dp_selfcheck = None
    if world > 1 and not os.environ.get("MOLKGNN_BENCH_NO_SELFCHECK"):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import dp_check
        try:
            dp_selfcheck = dp_check.run_check(64 * world)
        except Exception as e:                       # reported, never fatal for the bench line
            dp_selfcheck = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes per launch / measured launch duration) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path))
        peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    deg = np.bincount(batch["edge_index"][0], minlength=N)
    n = [int((deg == d).sum()) for d in range(1, 5)]
    fwd_b, bwd_b = roofline.stack_bytes(N, E, n, X_DIM, L_BASE, L_BASE, NUM_LAYERS, EDGE_DIM)
    fwd_f, bwd_f = roofline.stack_flops(E, n, X_DIM, L_BASE, L_BASE, NUM_LAYERS, EDGE_DIM)
    kbytes = roofline.kernel_bytes_per_step(top, N, E, n, X_DIM, L_BASE, L_BASE, NUM_LAYERS, EDGE_DIM)
    cnt, kms = top_timed.get(top, (0, 0.0))
    per_launch_ms = kms / max(cnt, 1)
    launches_per_step = cnt / args.steps
    achieved = (kbytes / max(launches_per_step, 1e-9)) / (per_launch_ms * 1e-3) / 1e9 if cnt else 0.0
    if fwd_only:
        fwd_b = roofline.stack_bytes(N, E, n, X_DIM, L_BASE, L_BASE, NUM_LAYERS, EDGE_DIM, training=False)[0]
        bwd_b, bwd_f = 0, 0
    step_gbs = (fwd_b + bwd_b) * args.steps / (ms * 1e-3) / 1e9
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/traffic.json), or null
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if top in tj and B == 4096:
            traffic, traffic_src = tj[top]["bytes_per_launch"], tj[top]["source"]
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms_per_launch": per_launch_ms,
            "kernel_share_of_step": kms / ms if ms else None,
            "algorithmic_bytes_per_launch": kbytes / max(launches_per_step, 1e-9),
            "step_algorithmic_bytes": fwd_b + bwd_b, "step_achieved_gbs": step_gbs, "step_frac": step_gbs / peak,
            "step_tflops_fp32": (fwd_f + bwd_f) * args.steps / (ms * 1e-3) / 1e12,
            "breakdown_ms_per_step": {k: v[1] / 3 for k, v in sorted(breakdown.items())}}
    line = {
        "metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(B, fwd_only),
        "clocks": clocks,
        "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "molecules/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": B * K * 4 if fwd_only else 4, "ms_per_step": ms_e2e / args.steps,
                "api": "molkgnn_b200.MolGCN.forward/backward; every step's x/p/edge_index/edge_attr copied from pinned host "
                       "memory and bucketed (molkgnn_b200.data.DevicePrefetcher: side streams, two steps ahead), every step's loss "
                       "copied to pinned host memory and consumed by the host one step later; all K copies and K reads are "
                       "inside the timed region"},
        "e2e_store": None if ms_store is None else {
            "value": world * B * args.steps / (ms_store * 1e-3), "unit": "molecules/s", "h2d_bytes_per_step": 8 * B,
            "d2h_bytes_per_step": B * K * 4 if fwd_only else 4, "ms_per_step": ms_store / args.steps,
            "api": "molkgnn_b200.store.MoleculeStore.collate(ids) -> MolGCN.forward/backward: the dataset is packed in HBM once, "
                   "every step copies only its (shuffled) molecule ids from pinned host memory, assembles the batch on the GPU "
                   "(csrc/collate.cu), runs the bucket pass inside the forward and reads the loss back"},
        "gpu_launches": launches,
        "roofline": roof,
        "nodes_per_gpu": N, "edges_per_gpu": E,
    }
    if dp_selfcheck is not None:
        line["dp_selfcheck"] = dp_selfcheck
    if bucket is not None:
        if bucket.oneshot is not None:
            bucket.oneshot.check()               # raises if a peer's flag ever timed out
        line["allreduce"] = ("one-shot kernel over NVLink peer memory (csrc/oneshot.cu)" if bucket.oneshot is not None
                                       else "ncclAllReduce (ReduceOp.AVG) of the flat gradient buffer")
    if world == 1 and not args.no_cpu_baseline:
        mol, el, cores, kind = cpu_stream(seconds=args.cpu_seconds)
        line["cpu_baseline"] = {"value": mol / el, "unit": "molecules/s", "cores": cores, "kind": kind,
                                "sample": f"{mol} molecules as batch-16 mini-batches (fwd+bwd) in {el:.1f} s, " + CPU_SAMPLE[kind]}
        if kind == "reference":                  # second figure: the vectorised port of the same algorithm (the stricter baseline)
            mol2, el2, _, _ = cpu_stream(seconds=min(4.0, args.cpu_seconds), kind="port")
            line["cpu_baseline"]["port_value"] = mol2 / el2
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
