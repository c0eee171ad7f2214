"""Diagnostic: does tcgen05 kind::f16 honour fp16 SUBNORMAL inputs (needed by the unscaled hi/lo operand split)?
   Prints the max error of A*B^T against float64 for operands drawn in the subnormal range."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molkgnn_b200 import _lib

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
N, K = 64, 64
for name, sa, sb in (("normal", 0.5, 0.5), ("A subnormal", 2.0 ** -17, 0.5), ("B subnormal", 0.5, 2.0 ** -17),
                     ("both subnormal", 2.0 ** -17, 2.0 ** -17)):
    A = (torch.randn(128, K, generator=g) * sa).half()
    B = (torch.randn(N, K, generator=g) * sb).half()
    ref = A.double() @ B.double().T
    D = torch.full((128, N), float("nan"), device=dev)
    Ad, Bd = A.contiguous().to(dev), B.contiguous().to(dev)
    _lib.check(_lib.lib().molkgnn_tc_selftest(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), N, K, 0, 0, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    err = float((D.double().cpu() - ref).abs().max())
    print(json.dumps(dict(case=name, max_err=err, ref_max=float(ref.abs().max()), rel=err / float(ref.abs().max()),
                          frac_sub_A=float(((A.abs() < 6.1e-5) & (A != 0)).float().mean()))), flush=True)
