"""The BASELINE.json configurations that are NOT the bench line, measured on one GPU with the same protocol as bench.py
(CUDA events on the launching stream, device-resident inputs, GPU bucket pass of every step inside the timed region):

  c3  wide kernel sets 40/80/120/200, 5 layers, 4096 molecules, fwd+bwd   (configs[2])
  c4  base model, 65 536 molecules per GPU, fwd+bwd                        (one GPU's share of configs[3])
  c5  base model, 65 536-molecule chunks, forward only under no_grad       (one GPU's share of configs[4]); also end to end
      with every chunk copied from pinned host memory and the result read back

One JSON line per configuration.  Usage: python tools/configs_bench.py [c3] [c4] [c5] [--steps K]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c3": dict(L=(40, 80, 120, 200), layers=5, mol=4096, fwd_only=False, steps=5),
    "c4": dict(L=(10, 20, 30, 50), layers=3, mol=65536, fwd_only=False, steps=20),
    "c5": dict(L=(10, 20, 30, 50), layers=3, mol=65536, fwd_only=True, steps=20),
}
_batches = {}


def make_batch(mol):
    """4096-molecule shards with different seeds, concatenated (the generator is a per-molecule Python loop)."""
    from molkgnn_b200 import synth
    if mol not in _batches:
        parts = [synth.make_batch(min(4096, mol - s), seed=7 + s) for s in range(0, mol, 4096)]
        off, ei = 0, []
        for b in parts:
            ei.append(b["edge_index"] + off)
            off += b["x"].shape[0]
        _batches[mol] = {"x": np.concatenate([b["x"] for b in parts]), "p": np.concatenate([b["p"] for b in parts]),
                         "edge_attr": np.concatenate([b["edge_attr"] for b in parts]),
                         "edge_index": np.concatenate(ei, axis=1)}
    return _batches[mol]


def run(name, steps=None):
    import molkgnn_b200 as mk
    from molkgnn_b200 import roofline, _lib, functional as Fn
    from molkgnn_b200.data import DevicePrefetcher
    c = CONFIGS[name]
    steps = steps or c["steps"]
    dev = torch.device("cuda", 0)
    batch = make_batch(c["mol"])
    N, E = batch["x"].shape[0], batch["edge_index"].shape[1]
    host = {k: torch.from_numpy(batch[k]).pin_memory() for k in ("x", "p", "edge_index", "edge_attr")}
    devt = {k: v.to(dev) for k, v in host.items()}
    torch.manual_seed(0)
    L = c["L"]
    net = mk.MolGCN(c["layers"], *L, *L, x_dim=28, p_dim=3, edge_attr_dim=7).to(dev)
    wout = torch.randn(N, sum(L), device=dev)
    pf = DevicePrefetcher(dev)
    params = list(net.parameters())

    def step(t, plan):
        if c["fwd_only"]:
            with torch.no_grad():
                return net(x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False, plan=plan)
        x = t["x"].detach().requires_grad_(True)
        h = net(x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"], p=t["p"], save_score=False, plan=plan)
        h.backward(wout)
        for p in params:
            p.grad = None
        return h

    def loop(k, from_host):
        kw = {"host_batch": host} if from_host else {"device_batch": devt}
        nxt = pf.put(build_plan=True, **kw)
        out = None
        for i in range(k):
            t, plan = pf.get(nxt)
            if i + 1 < k:
                nxt = pf.put(build_plan=True, **kw)
            h = step(t, plan)
            if from_host:
                out = float(h.detach()[:, 0].sum())       # device -> host read of the chunk's result
        return out

    def timed(k, from_host):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        loop(k, from_host)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    loop(3, False)
    Fn.profile_start()
    loop(2, False)
    breakdown = {k: v[1] / 2 for k, v in sorted(Fn.profile_stop().items()) if k != "bucket_count"}
    pc0 = Fn.path_counts()
    l0 = _lib.lib().molkgnn_launch_count()
    ms = timed(steps, False)
    launches = (_lib.lib().molkgnn_launch_count() - l0) / steps
    loop(2, True)
    ms_e2e = timed(steps, True)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    deg = np.bincount(batch["edge_index"][0], minlength=N)
    n = [int((deg == d).sum()) for d in range(1, 5)]
    fwd_b, bwd_b = roofline.stack_bytes(N, E, n, 28, L, L, c["layers"], 7)
    fwd_f, bwd_f = roofline.stack_flops(E, n, 28, L, L, c["layers"], 7)
    by = fwd_b if c["fwd_only"] else fwd_b + bwd_b
    fl = fwd_f if c["fwd_only"] else fwd_f + bwd_f
    print(json.dumps({
        "config": name, "pass": "forward only (no_grad)" if c["fwd_only"] else "fwd+bwd", "kernels": list(L),
        "layers": c["layers"], "molecules": c["mol"], "nodes": N, "edges": E, "steps": steps,
        "ms_per_step": ms, "molecules_per_s": c["mol"] / (ms * 1e-3),
        "e2e_ms_per_step": ms_e2e, "e2e_molecules_per_s": c["mol"] / (ms_e2e * 1e-3),
        "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()),
        "algorithmic_bytes_per_step": by, "achieved_gbs": by / (ms * 1e-3) / 1e9, "hbm_frac": by / (ms * 1e-3) / 1e9 / peak,
        "tflops_fp32": fl / (ms * 1e-3) / 1e12, "gpu_launches": launches, "path_counts": str(pc0),
        "breakdown_ms_per_step": breakdown, "peak_gbs": peak}), flush=True)
    del net, devt, host, wout
    torch.cuda.empty_cache()


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in CONFIGS] or list(CONFIGS)
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else None
    for nm in names:
        run(nm, steps)
