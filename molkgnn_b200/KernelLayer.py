"""Drop-in for the reference's ``models/MolKGNN/KernelLayer.py`` (MolGCN, lines 8-123).

Same constructor and ``forward(**kwargs)`` protocol; internally the whole stack (per layer: fused conv of the four
degree buckets, then the neighbour sum that PyG's ``propagate`` performed) runs as CUDA kernels, with the degree
buckets rebuilt on the GPU from ``edge_index`` in one pass per batch.
"""
from __future__ import annotations

import os

import torch
from torch.nn import Module, ModuleList

from .functional import MolGCNFn, StackPack, PARAMS_PER_DEGREE
from .kernels import KernelSetConv
from .plan import BucketPlan


class MolGCN(Module):
    def __init__(self, num_layers=5, num_kernel1_1hop=0, num_kernel2_1hop=0, num_kernel3_1hop=0, num_kernel4_1hop=0,
                 num_kernel1_Nhop=0, num_kernel2_Nhop=0, num_kernel3_Nhop=0, num_kernel4_Nhop=0, x_dim=5, p_dim=3,
                 edge_attr_dim=1):
        super(MolGCN, self).__init__()
        self.num_layers = num_layers
        if num_layers < 1:
            raise Exception('at least one convolution layer is needed')
        self.layers = ModuleList()
        self.num_kernels_list = []
        if (num_kernel1_1hop is not None) and (num_kernel2_1hop is not None) and (
                num_kernel3_1hop is not None) and (num_kernel4_1hop is not None):
            kernel_layer = KernelSetConv(num_kernel1_1hop, num_kernel2_1hop, num_kernel3_1hop, num_kernel4_1hop,
                                         D=p_dim, node_attr_dim=x_dim, edge_attr_dim=edge_attr_dim)
            num_kernels = num_kernel1_1hop + num_kernel2_1hop + num_kernel3_1hop + num_kernel4_1hop
        else:
            raise Exception('MolGCN: num_kernel1-4 need to be specified')
        self.layers.append(kernel_layer)
        self.num_kernels_list.append(num_kernels)
        for i in range(num_layers - 1):
            kernel_layer = KernelSetConv(L1=num_kernel1_Nhop, L2=num_kernel2_Nhop, L3=num_kernel3_Nhop,
                                         L4=num_kernel4_Nhop, D=p_dim, node_attr_dim=self.num_kernels(i),
                                         edge_attr_dim=edge_attr_dim)
            self.layers.append(kernel_layer)
            self.num_kernels_list.append(kernel_layer.get_num_kernel())
        self.edge_attr_dim = edge_attr_dim

    def num_kernels(self, layer):
        return self.num_kernels_list[layer]

    def __getstate__(self):
        # the native layer descriptors (ctypes) are a cache: never pickled / deep-copied with the module
        st = super(MolGCN, self).__getstate__() if hasattr(super(MolGCN, self), '__getstate__') else self.__dict__.copy()
        st = dict(st)
        st.pop('_stack', None)
        st.pop('_param_slots', None)
        return st

    def build_plan(self, edge_index, p, edge_attr, num_nodes):
        """GPU degree-bucket pass for one collated batch (replaces the offline pre-transform, wrapper.py:559-672)."""
        return BucketPlan.from_edge_index(edge_index, p, edge_attr, num_nodes)

    def forward(self, *argv, **kwargv):
        if len(argv) != 0:
            raise Exception('Kernel does not take positional argument, use keyword argument instead. e.g. '
                            'model(data=data)')
        x = kwargv['x']
        if x.is_cuda and x.device.index != torch.cuda.current_device():
            # the native calls take torch's CURRENT stream and allocate through torch: run them on the tensors' device
            with torch.cuda.device(x.device):
                return self.forward(**kwargv)
        edge_index = kwargv['edge_index']
        edge_attr = kwargv['edge_attr']
        p = kwargv['p']
        save_score = kwargv.get('save_score', False)
        # Reference protocol (KernelLayer.py:63-101): the 20 precomputed per-degree tensors arrive as kwargs.  The index tensors
        # (selected_index_deg*, nei_index_deg*) are rebuilt on the GPU from edge_index (bit-exact, tests/test_bucket_gpu.py);
        # the DATA tensors are honoured: the conv reads its bond rows from nei_edge_attr_deg* and the degree-4 coordinates from
        # p_focal_deg4 / nei_p_deg4 exactly like the reference (kernels.py:679, 356) -- NOT from `edge_attr`, which MolKGNNNet
        # batch-normalises before the call (MolKGNNNet.py:116-119).  Without those kwargs (callers that never ran the
        # pre-transform) the rows are gathered from `edge_attr` (or `raw_edge_attr=` if given), i.e. what the pre-transform
        # would have stored (wrapper.py:578-593).
        raw_edge_attr = kwargv.get('raw_edge_attr', None)
        plan = kwargv.get('plan', None)
        if plan is None:
            ref_rows = BucketPlan.ref_rows_from_kwargs(kwargv) if raw_edge_attr is None else None
            # first half of the GPU bucket pass: counting kernels queued; the host round trip for the bucket sizes happens
            # inside MolGCNFn.forward, after the host-side preparation below and the parameter packing were queued
            plan = BucketPlan.begin_from_edge_index(edge_index, p, edge_attr if raw_edge_attr is None else raw_edge_attr,
                                                    x.shape[0], ref_rows=ref_rows)
        if any(layer._has_mixed_sets() for layer in self.layers):
            # A layer that carries fixed (predefined, requires_grad=False) AND trainable kernels for a degree
            # (kernels.py:452-516, 702-715) has two sets of mixing weights per degree: it runs through the layer modules
            # (two passes of the same kernels per layer) and the native neighbour sum, layer by layer.
            return self._forward_layerwise(x, plan.finish(), save_score, kwargv)
        # Host-side fast path: the 84 parameter tensors are fetched straight from the modules' _parameters dicts (the
        # slots are collected once; nn.Module.__getattr__ per parameter and step is what made a step host bound), and the
        # native layer descriptors are rebuilt only when a parameter tensor moved (new storage, device, dtype).
        slots = self.__dict__.get('_param_slots')
        if slots is None:
            slots = []
            for layer in self.layers:
                for prm_owner in layer._degree_owners():
                    for k in PARAMS_PER_DEGREE:
                        slots.append(None if prm_owner is None else (prm_owner._parameters, k))
            self.__dict__['_param_slots'] = slots
        flat = [None if sl is None else sl[0][sl[1]] for sl in slots]
        key = (x.shape[1], x.device) + tuple(0 if t is None else t.data_ptr() for t in flat)
        stack = self.__dict__.get('_stack')
        if stack is None or stack.key != key:
            layer_params = [layer._degree_params() for layer in self.layers]
            for t in flat:
                if t is not None and not t.is_contiguous():
                    raise Exception('MolGCN: kernel parameters must be contiguous tensors')
            stack = StackPack(layer_params, x.shape[1], self.edge_attr_dim, x.device)
            stack.key = key
            self.__dict__['_stack'] = stack
        stack.prepack()
        # the packed (normalised) kernel rows are ONE workspace per module, rewritten by every forward: remember which
        # parameter versions it holds, so that a backward running after a later forward with modified parameters is refused
        # instead of silently differentiating with the wrong kernels (MolGCNFn.backward)
        stack.versions = tuple(0 if t is None else t._version for t in flat)
        aux = kwargv.get('aux', None)
        if save_score and aux is None:
            aux = {}
        # Parameter gradients: by default NOT through 72 autograd edges (72 view tensors, output checks and AccumulateGrad nodes per
        # step cost ~0.3 ms of host time, a third of the step's host budget and what the 8-GPU runs were bound by) -- the backward
        # writes / accumulates p.grad itself (functional.MolGCNFn, direct mode).  `loss.backward()` and optimizers see the same
        # .grad tensors; torch.autograd.grad w.r.t. the kernel parameters and tensor hooks on them need the autograd edges:
        # set `module.direct_param_grads = False` (or MOLKGNN_PARAM_GRADS=autograd); parameters that carry hooks switch it off
        # by themselves.
        direct = self.__dict__.get('direct_param_grads')
        if direct is None:
            direct = os.environ.get('MOLKGNN_PARAM_GRADS', 'direct') != 'autograd'
            self.__dict__['direct_param_grads'] = direct
        if direct:
            for t in flat:
                if t is not None and (t._backward_hooks or getattr(t, '_post_accumulate_grad_hooks', None)):
                    direct = False
                    break
        if direct:
            stack.direct_params = flat
            args = ()
            if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in flat):
                anchor = self.__dict__.get('_anchor')
                if anchor is None or anchor.device != x.device:
                    anchor = torch.zeros((), device=x.device, requires_grad=True)   # makes the graph reach the backward
                    self.__dict__['_anchor'] = anchor
                args = (anchor,)
            h = MolGCNFn.apply(x, plan, stack, kwargv.get('argmax_in', None), aux, *args)
        else:
            stack.direct_params = None
            h = MolGCNFn.apply(x, plan, stack, kwargv.get('argmax_in', None), aux, *flat)
        if save_score:
            # KernelLayer.py:117 hands save_score to every layer, each of which dumps its [N, K] score matrix (kernels.py:749-750,
            # 594-608: scores.csv, rewritten layer after layer).  The stack keeps the compact scores; expand and dump them here.
            for i, layer in enumerate(self.layers):
                layer.save_score(self._dense_scores(plan, stack.packs[i], aux['sc'][i]))
        return h

    def _forward_layerwise(self, x, plan, save_score, kwargv):
        """MolGCN.forward (KernelLayer.py:107-120) layer by layer: sim_sc = layer(...), h = propagate(edge_index, sim_sc)."""
        if kwargv.get('argmax_in', None) is not None or kwargv.get('aux', None) is not None:
            raise NotImplementedError('argmax_in / aux are not available for mixed fixed+trainable kernel sets')
        # neighbour sum in edge order from the plan's in-lists (deterministic: <= 4 gathered rows added one after the other)
        src = plan.in_src.long()
        ok = (src >= 0).unsqueeze(-1)
        src = src.clamp_min(0)
        h = x
        for i, layer in enumerate(self.layers):
            sim_sc = layer(is_last_layer=(i == self.num_layers - 1), x=h, plan=plan, save_score=save_score,
                           **{f'{k}_deg{d}': None for d in range(1, 5)
                              for k in ('selected_index', 'nei_index', 'p_focal', 'nei_p', 'nei_edge_attr')})
            h = torch.zeros_like(sim_sc)
            for t in range(4):
                h = h + torch.where(ok[:, t], sim_sc.index_select(0, src[:, t]), torch.zeros((), device=x.device))
        return h

    def save_kernels(self, dir, file_name):
        """GNNModel.save_kernels (model.py:417-431): the first layer's trainable kernel sets, keys '{deg-1}.{param}' as read
        by analyses/atom_encoder/kernel_reader.py:86."""
        import os
        if not os.path.exists(dir):
            os.mkdir(dir)
        torch.save(self.layers[0].trainable_kernelconv_set.state_dict(), dir + file_name)

    def save_kernellayer(self, path, time_stamp):
        """MolKGNNNet.save_kernellayer (MolKGNNNet.py:61-67): one state dict per layer."""
        for i, layer in enumerate(self.layers):
            torch.save(layer.state_dict(), f'{path}/{time_stamp}_{i}th_layer.pth')

    @staticmethod
    def _dense_scores(plan, pack, sc_compact):
        """compact per-degree score blocks -> the reference's sim_sc [N, K] (kernels.py:743-747)."""
        out = torch.zeros(plan.N, pack.K, dtype=torch.float32, device=sc_compact.device)
        scoff, _ = plan.scoff(pack.L)
        for d in range(4):
            n, L = plan.n[d], pack.L[d]
            if n == 0 or L == 0:
                continue
            rows = plan.sel[plan.boff[d]:plan.boff[d] + n].long()
            out[rows, pack.koff[d]:pack.koff[d] + L] = sc_compact[scoff[d]:scoff[d] + n * L].view(n, L)
        return out

    def message(self, sim_sc_j):
        return sim_sc_j
