"""Known-answer tests of the tcgen05 plumbing (csrc/tc.cuh): the shared-memory / instruction descriptor encodings that the
tensor-core conv kernels rely on, checked against a float64 product."""
import ctypes as C
import json
import os

import pytest
import torch

from molkgnn_b200 import _lib

pytestmark = pytest.mark.gpu


def run(N, K, a_mn, b_mn, swap=0, seed=0):
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(seed)
    A = (torch.randn(128, K, generator=g) * 0.5).half()
    B = (torch.randn(N, K, generator=g) * 0.5).half()
    ref = A.double() @ B.double().T
    Ad = (A.T if a_mn == 1 else A).contiguous().to(dev)
    Bd = (B.T if b_mn else B).contiguous().to(dev)
    D = torch.full((128, N), float("nan"), device=dev)
    _lib.check(_lib.lib().molkgnn_tc_selftest(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), N, K, a_mn, b_mn, swap,
                                              _lib.stream_ptr()))
    torch.cuda.synchronize()
    return float((D.double().cpu() - ref).abs().max()), float(ref.abs().max())


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("N,K", [(16, 16), (64, 64), (112, 112), (208, 128), (256, 32)])
def test_umma_known_answer(N, K, a_mn, b_mn):
    err, mag = run(N, K, a_mn, b_mn)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/tc_selftest.jsonl", "a") as f:
        f.write(json.dumps(dict(N=N, K=K, a_mn=a_mn, b_mn=b_mn, err=err, mag=mag)) + "\n")
    assert err < 1e-3 * max(mag, 1.0), f"UMMA mismatch: max err {err} (|ref| max {mag})"


@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("N,K", [(16, 16), (128, 112), (112, 128), (256, 64)])
def test_umma_a_operand_in_tensor_memory(N, K, b_mn):
    """A operand read from TMEM (a_mn = 2: lane = row, fp16 pairs packed per 32-bit column)."""
    err, mag = run(N, K, 2, b_mn)
    assert err < 1e-3 * max(mag, 1.0), f"UMMA (A in TMEM) mismatch: max err {err} (|ref| max {mag})"


@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("N,K", [(16, 16), (128, 112), (80, 32), (128, 128)])
def test_umma_a_operand_copied_to_tensor_memory_by_tcgen05_cp(N, K, b_mn):
    """A staged in shared memory (K-major core-matrix layout), moved to TMEM with tcgen05.cp.128x256b per K step, then read by
    the MMA as its TMEM operand (a_mn = 3) -- the operand path of the layer-fused forward."""
    err, mag = run(N, K, 3, b_mn)
    assert err < 1e-3 * max(mag, 1.0), f"UMMA (A via tcgen05.cp) mismatch: max err {err} (|ref| max {mag})"
