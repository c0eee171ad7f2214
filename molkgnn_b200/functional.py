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
    """Starts the library's CUDA-event profiler (one event pair per launch group, recorded on the launching stream by the
    native code, include/molkgnn_b200.h molkgnn_profile_enable) and the Python-side scopes of the per-layer API."""
    global _PROF, _PROF_NAMES
    _PROF, _PROF_NAMES = {}, (None if names is None else set(names))
    # the native profiler filters on ONE name (what bench.py needs inside its timed region) or records everything
    _lib.lib().molkgnn_profile_only(list(names)[0].encode() if names is not None and len(names) == 1 else None)
    _lib.lib().molkgnn_profile_enable(1)


def profile_stop():
    """-> {name: (launches, total_ms)}; synchronises."""
    global _PROF
    torch.cuda.synchronize()
    out = {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in (_PROF or {}).items()}
    _PROF = None
    buf = C.create_string_buffer(8192)
    n = _lib.lib().molkgnn_profile_read(buf, 8192)
    _lib.lib().molkgnn_profile_enable(0)
    if n < 0:
        raise _lib.MolKGNNError(_lib.lib().molkgnn_last_error().decode())
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        if name not in out:                      # the native scopes are the finer ones; Python scopes win on a clash
            out[name] = (int(cnt), float(ms))
    return out


class _timed(object):
    """Python-side scope; only used for names the native profiler does not record itself."""
    NATIVE = {"param_pack", "pad_norm", "x_images", "conv_fwd", "propagate_fwd", "bwd_w", "bwd_x", "param_finalize",
              "conv_bwd", "bucket_build"}

    def __init__(self, name):
        self.on = _PROF is not None and name not in self.NATIVE and (_PROF_NAMES is None or name in _PROF_NAMES)
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

    def __init__(self, params, F, Fe, device, c=None):
        L = _lib.lib()
        self.F, self.Fp, self.Fe = int(F), _rup4(F), int(Fe)
        if not (1 <= self.Fe <= 8):
            raise _lib.MolKGNNError(f"edge_attr_dim {Fe} not supported (1..8)")
        self.L = []
        if c is None:
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
                                          ptr(argmax), ptr(free), ptr(argmax_in), ptr(counter), ptr(ximg), None,
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
    coef = torch.empty(int(L.molkgnn_conv_bwd_coef_floats(C.byref(plan.c), C.byref(pack.c))), dtype=torch.float32, device=dev)
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
    for name, ph in (("conv_bwd", 7),):
        with _timed(name):
            check(L.molkgnn_conv_bwd(C.byref(plan.c), C.byref(pack.c), ptr(x), x.stride(0), ptr(xnorm), ptr(grad),
                                     grad.stride(0), grad_mode, ptr(argmax), _i64x4(scoff), ptr(coef), ptr(partials),
                                     ptr(gx), pack.Fp if need_gx else 0, C.byref(gc) if gc is not None else None, ph,
                                     ptr(ximg), ptr(scratch), None, stream_ptr()))
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
        gx, grads = conv_backward(ctx.plan, pack, xp, xnorm, grad_sc.float().contiguous(), 0, argmax, ctx.need_gx, need_gp,
                                  ximg=ctx.ximg)
        flat = _flatten_param_grads(grads, pack, ctx.needs_input_grad[7:])
        return (gx[:, :ctx.F] if gx is not None else None, None, None, None, None, None, None, *flat)


class StackPack(object):
    """The layers of one MolGCN as a contiguous ``molkgnn_layer_t[nl]`` + their packed workspaces.  Built once per module
    and reused while the parameter storages stay in place (in-place optimizer updates keep them)."""

    def __init__(self, layer_params, F0, Fe, device):
        self.nl = len(layer_params)
        if not (1 <= self.nl <= _lib.MAX_LAYERS):
            raise _lib.MolKGNNError(f"{self.nl} layers not supported (1..{_lib.MAX_LAYERS})")
        self.arr = (_lib.Layer * self.nl)()
        self.packs = []
        F = F0
        for i, params in enumerate(layer_params):
            self.packs.append(LayerPack(params, F, Fe, device, c=self.arr[i]))
            F = self.packs[-1].K
        self.F0, self.Fe, self.device = F0, Fe, device
        self.key = self.make_key(layer_params, F0, Fe, device)
        # flat parameter-gradient views, in flat_params() order; offsets come from the native layout (plan independent)
        self.last_grad_flat = None
        self.packed_all = False

    @staticmethod
    def make_key(layer_params, F0, Fe, device):
        k = [F0, Fe, str(device)]
        for params in layer_params:
            for prm in params:
                if prm is None:
                    k.append(0)
                else:
                    k += [prm[n].data_ptr() if prm[n].is_contiguous() else object() for n in PARAMS_PER_DEGREE]
        return tuple(k)

    def prepack(self):
        """Queues the parameter packing of every layer (normalised kernel rows, tensor-core images) on the current stream."""
        # layers that run the molecule-tile kernels do not need the bucket-order tensor-core images (what = 1 | 4); should
        # the plan turn out untiled, MolGCNFn.forward lets the native forward pack everything again
        self.packed_all = any(pk.tile_img is None for pk in self.packs) or _lib.lib().molkgnn_get_fwd_path() < 2
        check(_lib.lib().molkgnn_param_pack_layers(self.arr, self.nl, 7 if self.packed_all else 5, stream_ptr()))

    def layout(self, plan: BucketPlan, flags):
        lay = _lib.StackLayout()
        check(_lib.lib().molkgnn_stack_layout(C.byref(plan.c), self.arr, self.nl, flags, C.byref(lay)))
        return lay

    def persistent_grad_flat(self, lay, dev):
        """The stack's own flat gradient buffer (direct mode): allocated once, rewritten by every backward that finds no
        parameter holding a gradient."""
        f = self.__dict__.get("_pflat")
        if f is None or f.numel() != lay.grad_floats or f.device != dev:
            f = torch.empty(lay.grad_floats, dtype=torch.float32, device=dev)
            self._pflat = f
            self._pviews = None
        return f

    def persistent_grad_views(self, lay):
        v = self.__dict__.get("_pviews")
        if v is None:
            v = self.grad_views(lay, self._pflat)
            self._pviews = v
        return v

    def grad_views(self, lay, flat):
        """Views of the flat gradient buffer in flat_params() order (None for parameters without a gradient)."""
        specs = self.__dict__.get("_view_specs")
        if specs is None:                  # (offset, shape, stride) per parameter; the g_* offsets do not depend on the plan
            specs = []
            for i, pk in enumerate(self.packs):
                for d in range(4):
                    Ld = pk.L[d]
                    if not Ld:
                        specs += [None] * 7
                        continue
                    oc, os_, oe, ow = (lay.g_x_center[i][d], lay.g_x_support[i][d], lay.g_edge_attr_support[i][d],
                                       lay.g_w[i][d])
                    specs += [(oc, (Ld, pk.F), (pk.F, 1)), (os_, (Ld, d + 1, pk.F), ((d + 1) * pk.F, pk.F, 1)),
                              (oe, (Ld, d + 1, pk.Fe), ((d + 1) * pk.Fe, pk.Fe, 1)), None, (ow, (), ()), (ow + 1, (), ()),
                              (ow + 2, (), ())]
            self._view_specs = specs
        as_strided = flat.as_strided
        return [None if sp is None else as_strided(sp[1], sp[2], sp[0]) for sp in specs]


FLAG_KEEP_SC, FLAG_WANT_FREE, FLAG_PACKED, FLAG_INFERENCE = 1, 2, 4, 8


class MolGCNFn(torch.autograd.Function):
    """The whole conv stack: for every layer conv -> propagate (MolGCN.forward, KernelLayer.py:107-120); one native call
    for the forward and one for the backward (csrc/stack.cu)."""

    @staticmethod
    def forward(ctx, x, plan, stack, argmax_in, aux, *flat):
        # direct mode (see MolGCN.forward / DESIGN.md 5): `flat` is ONE anchor tensor (or nothing), the parameters themselves are in
        # stack.direct_params and backward() writes their .grad itself instead of returning 72 tensors to the autograd engine
        ctx.direct = getattr(stack, "direct_params", None) if len(flat) <= 1 else None
        L = _lib.lib()
        dev = x.device
        nl = stack.nl
        plan.finish()                                # bucket sizes to the host (the packing queued by the caller runs meanwhile)
        flags = (FLAG_KEEP_SC | FLAG_WANT_FREE) if aux is not None else 0
        if stack.packed_all or plan.n_tiles > 0:
            flags |= FLAG_PACKED
        if not any(ctx.needs_input_grad):            # torch.no_grad() / nothing requires a gradient: inference
            flags |= FLAG_INFERENCE
        lay = stack.layout(plan, flags)
        xc = x.detach()
        if xc.dtype != torch.float32:
            xc = xc.float()
        if xc.stride(1) != 1:
            xc = xc.contiguous()
        ws = torch.empty(lay.fwd_bytes, dtype=torch.uint8, device=dev)
        Kl, Kpl = stack.packs[-1].K, stack.packs[-1].Kp
        h = torch.empty(plan.N, Kpl, dtype=torch.float32, device=dev)
        forced, keep = None, []
        if argmax_in is not None:
            forced = (C.c_void_p * nl)()
            for i in range(nl):
                if argmax_in[i] is None:
                    continue
                t = argmax_in[i].to(device=dev, dtype=torch.uint8).contiguous()
                if t.numel() != max(lay.sc_elems[i], 1) and t.numel() != lay.sc_elems[i]:
                    raise _lib.MolKGNNError("argmax_in has the wrong size")
                keep.append(t)
                forced[i] = t.data_ptr()
        tile_fwd = (C.c_int32 * nl)()
        check(L.molkgnn_stack_fwd(C.byref(plan.c), stack.arr, nl, C.byref(lay), flags, ptr(xc), xc.stride(0), ptr(ws),
                                  ptr(h), Kpl, forced, tile_fwd, stream_ptr()))
        ctx.tile_fwd = tile_fwd
        if aux is not None:
            for i in range(nl):
                n = lay.sc_elems[i]
                aux.setdefault("argmax", []).append(ws[lay.argmax[i]:lay.argmax[i] + max(n, 1)])
                aux.setdefault("argmax_free", []).append(ws[lay.argmax_free[i]:lay.argmax_free[i] + max(n, 1)])
                aux.setdefault("sc", []).append(ws[lay.sc[i]:lay.sc[i] + 4 * max(n, 1)].view(torch.float32))
        ctx.plan, ctx.stack, ctx.lay = plan, stack, lay
        ctx.versions = getattr(stack, "versions", None)
        ctx.save_for_backward(ws)
        ctx.need_gx = x.requires_grad
        ctx.F0 = x.shape[1]
        return h[:, :Kl]

    @staticmethod
    def backward(ctx, grad_h):
        (ws,) = ctx.saved_tensors
        if ws.device.index != torch.cuda.current_device():
            with torch.cuda.device(ws.device):          # autograd may call from another current device
                return MolGCNFn._backward(ctx, grad_h, ws)
        return MolGCNFn._backward(ctx, grad_h, ws)

    @staticmethod
    def _backward(ctx, grad_h, ws):
        L = _lib.lib()
        plan, stack, lay = ctx.plan, ctx.stack, ctx.lay
        if ctx.versions != getattr(stack, "versions", None):
            raise _lib.MolKGNNError(
                "MolGCN backward: the kernel parameters were modified in place and re-packed by a later forward of the same "
                "module before this backward ran (forward A, optimizer.step(), forward B, backward A).  The packed kernel rows "
                "are one workspace per module; run backward A before the parameters change, or use a second module copy.")
        dev = ws.device
        g = grad_h
        if g.dtype != torch.float32:
            g = g.float()
        if g.stride(1) != 1 or g.data_ptr() % 16:
            g = g.contiguous()
        need_gp = any(ctx.needs_input_grad[5:])
        direct = ctx.direct
        accumulate = False
        if direct is not None and need_gp:
            # Direct mode.  If no parameter holds a gradient (zero_grad(set_to_none=True) / first step) the gradients are written
            # into the stack's PERSISTENT flat buffer and every p.grad becomes its cached view of it: no allocation, no 72 new
            # view tensors, no 72 AccumulateGrad nodes per step.  The buffer is overwritten by the next such backward -- the
            # semantics of DDP's gradient_as_bucket_view=True.  If some parameter still holds a gradient (accumulation over
            # micro-batches, a second backward through the same graph) a fresh buffer is used and added to the existing .grad.
            accumulate = any(t is not None and t.grad is not None for t in direct)
        bs = torch.empty(lay.bwd_bytes, dtype=torch.uint8, device=dev)
        if not need_gp:
            gflat = None
        elif direct is not None and not accumulate:
            gflat = stack.persistent_grad_flat(lay, dev)
        else:
            gflat = torch.empty(lay.grad_floats, dtype=torch.float32, device=dev)
        Fp0 = stack.packs[0].Fp
        gx = torch.empty(plan.N, Fp0, dtype=torch.float32, device=dev) if ctx.need_gx else None
        check(L.molkgnn_stack_bwd(C.byref(plan.c), stack.arr, stack.nl, C.byref(lay), ptr(ws), ptr(bs), ptr(g), g.stride(0),
                                  ptr(gx), ptr(gflat), ctx.tile_fwd, stream_ptr()))
        stack.last_grad_flat = gflat                # dp.GradBucket all-reduces this buffer in place
        gx_out = gx[:, :ctx.F0] if gx is not None else None
        if direct is not None:
            if need_gp:
                views = stack.persistent_grad_views(lay) if not accumulate else stack.grad_views(lay, gflat)
                for t, v in zip(direct, views):
                    if t is None or v is None or not t.requires_grad:
                        continue
                    if t.grad is None:
                        t.grad = v
                    else:
                        t.grad.add_(v)
            return (gx_out, None, None, None, None) + ((None,) if len(ctx.needs_input_grad) > 5 else ())
        flat_all = stack.grad_views(lay, gflat) if need_gp else [None] * (stack.nl * 28)
        return (gx_out, None, None, None, None, *flat_all)


def flat_params(params):
    """list (4 degrees) of dict-like -> flat tuple of tensors in PARAMS_PER_DEGREE order (None for absent degrees)."""
    out = []
    for prm in params:
        for k in PARAMS_PER_DEGREE:
            out.append(None if prm is None else prm[k])
    return out
