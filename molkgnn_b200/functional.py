"""Autograd glue between the PyTorch modules and the CUDA kernels behind the C-ABI.

Torch provides device memory, the current stream and the autograd graph; every FLOP of the conv path runs in
``libmolkgnn_b200.so``.  There is no CPU/PyTorch fallback on purpose.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ptr, stream_ptr, check
from .plan import BucketPlan

# optional CUDA-event profiler (bench.py): name -> list of (start, end) events recorded on the launching stream
_PROF = None
_PROF_NAMES = None


def profile_start(names=None):
    global _PROF, _PROF_NAMES
    _PROF, _PROF_NAMES = {}, (None if names is None else set(names))


def profile_stop():
    """-> {name: (launches, total_ms)}; synchronises."""
    global _PROF
    torch.cuda.synchronize()
    out = {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in (_PROF or {}).items()}
    _PROF = None
    return out


class _timed(object):
    def __init__(self, name):
        self.on = _PROF is not None and (_PROF_NAMES is None or name in _PROF_NAMES)
        self.name = name

    def __enter__(self):
        if self.on:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()

    def __exit__(self, *a):
        if self.on:
            self.e.record()
            _PROF.setdefault(self.name, []).append((self.s, self.e))

PARAMS_PER_DEGREE = ("x_center", "x_support", "edge_attr_support", "p_support", "support_attr_sc_weight",
                     "center_attr_sc_weight", "edge_attr_support_sc_weight")


def _rup4(v):
    return (int(v) + 3) // 4 * 4


class LayerPack(object):
    """ctypes ``molkgnn_layer_t`` + the packed (normalised) kernel-set workspace of one KernelSetConv layer.

    ``params``: list of 4 (degree 1..4) dicts/objects with the KernelConv parameter tensors, or None for a degree
    without kernels."""

    def __init__(self, params, F, Fe, device):
        L = _lib.lib()
        self.F, self.Fp, self.Fe = int(F), _rup4(F), int(Fe)
        if not (1 <= self.Fe <= 8):
            raise _lib.MolKGNNError(f"edge_attr_dim {Fe} not supported (1..8)")
        self.L = []
        c = _lib.Layer()
        c.F, c.Fp, c.Fe = self.F, self.Fp, self.Fe
        self._keep = []
        koff = 0
        self.packed = []
        for d in range(4):
            prm = params[d]
            Ld = 0 if prm is None else int(prm["x_center"].shape[0])
            self.L.append(Ld)
            c.L[d] = Ld
            c.koff[d] = koff
            koff += Ld
            if Ld:
                ts = {}
                for k in PARAMS_PER_DEGREE:
                    t = prm[k].detach()
                    if t.dtype != torch.float32 or not t.is_cuda:
                        raise _lib.MolKGNNError(f"kernel parameter {k} must be a float32 CUDA tensor")
                    ts[k] = t.contiguous()
                    self._keep.append(ts[k])
                if ts["x_support"].shape != (Ld, d + 1, self.F) or ts["x_center"].shape != (Ld, self.F):
                    raise _lib.MolKGNNError(f"degree {d + 1}: kernel shapes do not match node_attr_dim={self.F}")
                if ts["edge_attr_support"].shape != (Ld, d + 1, self.Fe):
                    raise _lib.MolKGNNError(f"degree {d + 1}: edge_attr_support shape mismatch")
                c.x_center[d] = ts["x_center"].data_ptr()
                c.x_support[d] = ts["x_support"].data_ptr()
                c.edge_attr_support[d] = ts["edge_attr_support"].data_ptr()
                c.p_support[d] = ts["p_support"].data_ptr()
                c.w_support[d] = ts["support_attr_sc_weight"].data_ptr()
                c.w_center[d] = ts["center_attr_sc_weight"].data_ptr()
                c.w_edge[d] = ts["edge_attr_support_sc_weight"].data_ptr()
            pk = torch.empty(max(int(L.molkgnn_packed_floats(d + 1, Ld, self.Fp)), 4), dtype=torch.float32, device=device)
            self.packed.append(pk)
            c.packed[d] = pk.data_ptr()
        c.K = koff
        nb = int(L.molkgnn_tile_img_bytes(C.byref(c)))
        self.tile_img = torch.empty(nb, dtype=torch.uint8, device=device) if nb > 0 else None
        c.tile_img = self.tile_img.data_ptr() if nb > 0 else None
        self.K, self.Kp = koff, _rup4(koff)
        self.koff = [int(c.koff[d]) for d in range(4)]
        self.c = c
        self.device = device

    def pack(self):
        with _timed("param_pack"):
            check(_lib.lib().molkgnn_param_pack(C.byref(self.c), stream_ptr()))
        return self


def _i64x4(v):
    a = (C.c_int64 * 4)()
    for d in range(4):
        a[d] = int(v[d])
    return a


def pad_norm(x, Fp):
    """(x padded to [N,Fp] (x itself if already so), row norms)."""
    N, F = x.shape
    x = x.float()
    if x.stride(1) != 1:
        x = x.contiguous()
    ldx = x.stride(0)
    norm = torch.empty(N, dtype=torch.float32, device=x.device)
    if F == Fp and ldx % 4 == 0 and x.data_ptr() % 16 == 0:
        out = x
    else:
        out = torch.empty(N, Fp, dtype=torch.float32, device=x.device)
    with _timed("pad_norm"):
        check(_lib.lib().molkgnn_pad_norm(ptr(x), N, F, ldx, ptr(out), out.stride(0), ptr(norm), stream_ptr()))
    return out, norm


def x_images(plan: BucketPlan, pack: LayerPack, x, xnorm):
    """Normalised fp16 (hi, lo) images of the activations in tile order (tensor-core operand of the tile kernels), or
    None when the plan carries no molecule tiles / the layer is not eligible."""
    L = _lib.lib()
    one = int(L.molkgnn_tile_ximg_bytes(C.byref(plan.c), C.byref(pack.c))) if pack.tile_img is not None else 0
    if one <= 0:
        return None
    ximg = torch.empty(plan.n_tiles * one, dtype=torch.uint8, device=x.device)
    with _timed("x_images"):
        check(L.molkgnn_tile_ximg_build(C.byref(plan.c), C.byref(pack.c), ptr(x), x.stride(0), ptr(xnorm), ptr(ximg),
                                        stream_ptr()))
    return ximg


def conv_forward(plan: BucketPlan, pack: LayerPack, x, xnorm, is_last, dense=False, argmax_in=None, want_free=False,
                 ximg=None):
    """-> (sc, argmax_used_u8, argmax_free_u8 or None).  sc compact [sum n_d L_d] or dense zero-filled [N,Kp]."""
    scoff, tot = plan.scoff(pack.L)
    dev = x.device
    argmax = torch.empty(max(tot, 1), dtype=torch.uint8, device=dev)
    free = torch.empty(max(tot, 1), dtype=torch.uint8, device=dev) if want_free else None
    counter = torch.empty(8, dtype=torch.int32, device=dev)
    if dense:
        sc = torch.zeros(plan.N, pack.Kp, dtype=torch.float32, device=dev)
        ld = pack.Kp
    else:
        sc = torch.empty(max(tot, 1), dtype=torch.float32, device=dev)
        ld = 0
    if argmax_in is not None:
        argmax_in = argmax_in.to(device=dev, dtype=torch.uint8).contiguous()
        if argmax_in.numel() != max(tot, 1) and argmax_in.numel() != tot:
            raise _lib.MolKGNNError("argmax_in has the wrong size")
    with _timed("conv_fwd"):
        check(_lib.lib().molkgnn_conv_fwd(C.byref(plan.c), C.byref(pack.c), ptr(x), x.stride(0), ptr(xnorm),
                                          1 if is_last else 0, ptr(sc), 1 if dense else 0, ld, _i64x4(scoff),
                                          ptr(argmax), ptr(free), ptr(argmax_in), ptr(counter), ptr(ximg),
                                          stream_ptr()))
    return sc, argmax, free


def propagate_forward(plan: BucketPlan, pack: LayerPack, sc, next_pack=None):
    """-> (h [N,Kp], ||h_i||, fp16 images of h for ``next_pack`` or None)"""
    L = _lib.lib()
    scoff, _ = plan.scoff(pack.L)
    h = torch.empty(plan.N, pack.Kp, dtype=torch.float32, device=sc.device)
    hnorm = torch.empty(plan.N, dtype=torch.float32, device=sc.device)
    ximg = None
    if next_pack is not None and next_pack.tile_img is not None and pack.Kp <= 112 and next_pack.Fp == pack.Kp:
        one = int(L.molkgnn_tile_ximg_bytes(C.byref(plan.c), C.byref(next_pack.c)))
        if one > 0:
            ximg = torch.empty(plan.n_tiles * one, dtype=torch.uint8, device=sc.device)
    with _timed("propagate_fwd"):
        check(L.molkgnn_propagate_fwd(C.byref(plan.c), C.byref(pack.c), ptr(sc), _i64x4(scoff), ptr(h), pack.Kp,
                                      ptr(hnorm), ptr(ximg), stream_ptr()))
    return h, hnorm, ximg


def path_counts():
    a = (C.c_int64 * 4)()
    _lib.lib().molkgnn_path_counts(a)
    return dict(fwd_tile=int(a[0]), fwd_other=int(a[1]), bwd_tile=int(a[2]), bwd_other=int(a[3]))


def conv_backward(plan: BucketPlan, pack: LayerPack, x, xnorm, grad, grad_mode, argmax, need_gx, need_gparams,
                  ximg=None):
    """-> (gx [N,Fp] or None, list of 4 dicts of parameter grads or None)."""
    L = _lib.lib()
    dev = x.device
    scoff, tot = plan.scoff(pack.L)
    coef = torch.empty(max(tot, 1), dtype=torch.float32, device=dev)
    partials = torch.empty(int(L.molkgnn_conv_bwd_partial_floats(C.byref(plan.c), C.byref(pack.c))),
                           dtype=torch.float32, device=dev)
    gx = torch.empty(plan.N, pack.Fp, dtype=torch.float32, device=dev) if need_gx else None
    grads, gc = None, None
    if need_gparams:
        gc = _lib.LayerGrads()
        grads = []
        for d in range(4):
            Ld = pack.L[d]
            if not Ld:
                grads.append(None)
                continue
            g = dict(x_center=torch.empty(Ld, pack.F, dtype=torch.float32, device=dev),
                     x_support=torch.empty(Ld, d + 1, pack.F, dtype=torch.float32, device=dev),
                     edge_attr_support=torch.empty(Ld, d + 1, pack.Fe, dtype=torch.float32, device=dev),
                     w=torch.empty(3, dtype=torch.float32, device=dev))
            gc.x_center[d] = g["x_center"].data_ptr()
            gc.x_support[d] = g["x_support"].data_ptr()
            gc.edge_attr_support[d] = g["edge_attr_support"].data_ptr()
            gc.w_support[d] = g["w"].data_ptr()
            gc.w_center[d] = g["w"].data_ptr() + 4
            gc.w_edge[d] = g["w"].data_ptr() + 8
            grads.append(g)
    if grad.stride(1) != 1 or grad.data_ptr() % 16:
        grad = grad.contiguous()
    scratch = None
    if ximg is not None:
        scratch = torch.empty(plan.N * ((pack.Fp + 15) // 16 * 16), dtype=torch.float32, device=dev)
    # one C call runs all three kernels; under the profiler they are issued separately so each can be timed
    for name, ph in ((("conv_bwd", 7),) if _PROF is None else (("bwd_w", 1), ("param_finalize", 2), ("bwd_x", 4))):
        with _timed(name):
            check(L.molkgnn_conv_bwd(C.byref(plan.c), C.byref(pack.c), ptr(x), x.stride(0), ptr(xnorm), ptr(grad),
                                     grad.stride(0), grad_mode, ptr(argmax), _i64x4(scoff), ptr(coef), ptr(partials),
                                     ptr(gx), pack.Fp if need_gx else 0, C.byref(gc) if gc is not None else None, ph,
                                     ptr(ximg), ptr(scratch), stream_ptr()))
    return gx, grads


def _flatten_param_grads(grads, pack, needs):
    """grads of one layer -> flat tuple in PARAMS_PER_DEGREE order for the 4 degrees (None where absent)."""
    out = []
    for d in range(4):
        g = grads[d] if grads is not None else None
        if g is None:
            out += [None] * 7
            continue
        out += [g["x_center"], g["x_support"], g["edge_attr_support"], None, g["w"][0], g["w"][1], g["w"][2]]
    return out


class KernelSetConvFn(torch.autograd.Function):
    """One KernelSetConv layer -> dense sim_sc [N,K]  (BaseKernelSetConv.forward, kernels.py:610-751)."""

    @staticmethod
    def forward(ctx, x, plan, params, Fe, is_last, argmax_in, aux, *flat):
        pack = LayerPack(params, x.shape[1], Fe, x.device).pack()
        xp, xnorm = pad_norm(x.detach(), pack.Fp)
        ximg = x_images(plan, pack, xp, xnorm)
        sc, argmax, free = conv_forward(plan, pack, xp, xnorm, is_last, dense=True, argmax_in=argmax_in,
                                        want_free=aux is not None, ximg=ximg)
        ctx.ximg = ximg
        if aux is not None:
            aux["argmax"], aux["argmax_free"] = argmax, free
        ctx.plan, ctx.pack = plan, pack
        ctx.save_for_backward(xp, xnorm, argmax)
        ctx.need_gx = x.requires_grad
        ctx.F = x.shape[1]
        return sc[:, :pack.K]

    @staticmethod
    def backward(ctx, grad_sc):
        xp, xnorm, argmax = ctx.saved_tensors
        pack = ctx.pack
        need_gp = any(ctx.needs_input_grad[7:])
        gx, grads = conv_backward(ctx.plan, pack, xp, xnorm, grad_sc.float(), 0, argmax, ctx.need_gx, need_gp,
                                  ximg=ctx.ximg)
        flat = _flatten_param_grads(grads, pack, ctx.needs_input_grad[7:])
        return (gx[:, :ctx.F] if gx is not None else None, None, None, None, None, None, None, *flat)


class MolGCNFn(torch.autograd.Function):
    """The whole conv stack: for every layer conv -> propagate (MolGCN.forward, KernelLayer.py:107-120)."""

    @staticmethod
    def forward(ctx, x, plan, layer_params, Fe, argmax_in, aux, *flat):
        dev = x.device
        nl = len(layer_params)
        F = x.shape[1]
        packs, saved, ximgs = [], [], []
        for params in layer_params:
            packs.append(LayerPack(params, F, Fe, dev).pack())
            F = packs[-1].K
        h, hnorm, ximg = None, None, None
        for i, pack in enumerate(packs):
            if i == 0:
                h, hnorm = pad_norm(x.detach(), pack.Fp)
            if ximg is None:
                ximg = x_images(plan, pack, h, hnorm)
            sc, argmax, free = conv_forward(plan, pack, h, hnorm, i == nl - 1, dense=False,
                                            argmax_in=None if argmax_in is None else argmax_in[i],
                                            want_free=aux is not None, ximg=ximg)
            if aux is not None:
                aux.setdefault("argmax", []).append(argmax)
                aux.setdefault("argmax_free", []).append(free)
                aux.setdefault("sc", []).append(sc)
            saved += [h, hnorm, argmax]
            ximgs.append(ximg)
            h, hnorm, ximg = propagate_forward(plan, pack, sc, packs[i + 1] if i + 1 < nl else None)
        ctx.plan, ctx.packs, ctx.ximgs = plan, packs, ximgs
        ctx.save_for_backward(*saved)
        ctx.need_gx = x.requires_grad
        ctx.F0 = x.shape[1]
        return h[:, :packs[-1].K]

    @staticmethod
    def backward(ctx, grad_h):
        saved = ctx.saved_tensors
        packs = ctx.packs
        nl = len(packs)
        needs = ctx.needs_input_grad[6:]
        g = grad_h.float()
        flat_all = [None] * (nl * 28)
        for i in range(nl - 1, -1, -1):
            xp, xnorm, argmax = saved[3 * i:3 * i + 3]
            need_gp = any(needs[28 * i:28 * (i + 1)])
            need_gx = ctx.need_gx if i == 0 else True
            gx, grads = conv_backward(ctx.plan, packs[i], xp, xnorm, g, 1, argmax, need_gx, need_gp, ximg=ctx.ximgs[i])
            flat_all[28 * i:28 * (i + 1)] = _flatten_param_grads(grads, packs[i], None)
            g = gx
        return (g[:, :ctx.F0] if g is not None else None, None, None, None, None, None, *flat_all)


def flat_params(params):
    """list (4 degrees) of dict-like -> flat tuple of tensors in PARAMS_PER_DEGREE order (None for absent degrees)."""
    out = []
    for prm in params:
        for k in PARAMS_PER_DEGREE:
            out.append(None if prm is None else prm[k])
    return out
