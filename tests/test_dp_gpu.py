"""Multi-GPU check of the one-shot NVLink all-reduce (csrc/oneshot.cu, molkgnn_b200.dp.OneShotAllReduce) against NCCL.
Needs >= 2 GPUs on the node (skipped otherwise): launches tools/oneshot_check.py under torchrun, which runs 64 back-to-back
steps with skewed ranks and compares with ncclAllReduce."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_oneshot_allreduce_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "oneshot_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["created"] and out["world"] == n
    assert out["bitwise_identical_across_ranks"]
    # the sum runs in rank order, NCCL's in ring/tree order: equal up to fp32 rounding of a W-term sum
    assert out["max_rel_vs_nccl"] < 1e-6 and out["avg_rel"] < 1e-6


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_cuda_gradients_after_allreduce_equal_full_batch():
    """Sum over ranks of the shard gradients (CUDA kernels, GradBucket.allreduce over NCCL and over the one-shot NVLink kernel)
    == gradients of the full batch on one GPU, to 1e-5 (another summation order over molecules)."""
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["world"] == n
    assert out["nccl"]["max_rel_err_vs_full_batch"] < 1e-5
    assert out["oneshot"]["max_rel_err_vs_full_batch"] < 1e-5
