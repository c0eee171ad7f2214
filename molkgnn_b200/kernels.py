"""Drop-in modules for the reference's ``models/MolKGNN/kernels.py``: same class names, constructor signatures,
parameter names / shapes / init order (so reference checkpoints and seeds carry over), same forward protocol and
error behaviour -- but the arithmetic runs in the sm_100a kernels behind ``libmolkgnn_b200.so``.

Reference map
  KernelConv          /root/reference/models/MolKGNN/kernels.py:9-448
  BaseKernelSetConv   kernels.py:451-751
  KernelSetConv       kernels.py:754-781
"""
from __future__ import annotations

import os

import torch
from torch.nn import Module, ModuleList
from torch.nn.parameter import Parameter

from . import _lib
from .functional import KernelSetConvFn, flat_params, PARAMS_PER_DEGREE
from .plan import BucketPlan


class _Bag(object):
    """Minimal attribute bag standing in for torch_geometric.data.Data (the reference uses it only as such)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def _param_dict(kc):
    return {k: getattr(kc, k) for k in PARAMS_PER_DEGREE}


class KernelConv(Module):
    def __init__(self, L=None, D=None, num_supports=None, node_attr_dim=None, edge_attr_dim=None, init_kernel=None,
                 requires_grad=True, init_length_sc_weight=0.2, init_angle_sc_weight=0.2,
                 init_center_attr_sc_weight=0.2, init_support_attr_sc_weight=0.2,
                 init_edge_attr_support_sc_weight=0.2, weight_requires_grad=True):
        super(KernelConv, self).__init__()
        if init_kernel is None:
            if (L is None) or (D is None) or (num_supports is None) or (node_attr_dim is None) or (
                    edge_attr_dim is None):
                raise Exception('either number of kernels L, convolution dimention D, number of support num_supports '
                                'or feature dimension node_attr_dim is not specified')
            # same RNG order as kernels.py:50-53
            init_kernel = _Bag(x_center=torch.randn(L, node_attr_dim),
                               x_support=torch.randn(L, num_supports, node_attr_dim),
                               edge_attr_support=torch.randn(L, num_supports, edge_attr_dim),
                               p_support=torch.randn(L, num_supports, D))
        self.num_kernels = init_kernel.x_center.shape[0]
        self.x_center = Parameter(init_kernel.x_center, requires_grad=requires_grad)
        self.x_support = Parameter(init_kernel.x_support, requires_grad=requires_grad)
        self.edge_attr_support = Parameter(init_kernel.edge_attr_support, requires_grad=requires_grad)
        self.p_support = Parameter(init_kernel.p_support, requires_grad=requires_grad)
        self.length_sc_weight = Parameter(torch.tensor(init_length_sc_weight), requires_grad=weight_requires_grad)
        self.angle_sc_weight = Parameter(torch.tensor(init_angle_sc_weight), requires_grad=weight_requires_grad)
        self.center_attr_sc_weight = Parameter(torch.tensor(init_center_attr_sc_weight),
                                               requires_grad=weight_requires_grad)
        self.support_attr_sc_weight = Parameter(torch.tensor(init_support_attr_sc_weight),
                                                requires_grad=weight_requires_grad)
        self.edge_attr_support_sc_weight = Parameter(torch.tensor(init_edge_attr_support_sc_weight),
                                                     requires_grad=weight_requires_grad)

    def get_num_kernels(self):
        return self.num_kernels

    def forward(self, is_last_layer, **kwargv):
        """One degree bucket, reference protocol (kernels.py:428-448): ``data=`` with x_focal/p_focal/x_neighbor/
        p_neighbor/edge_attr_neighbor, or those five as keyword arguments.  Returns [L, n] (kernel-major)."""
        if len(kwargv) == 1:
            d = kwargv['data']
            x_focal, p_focal, x_neighbor = d.x_focal, d.p_focal, d.x_neighbor
            p_neighbor, edge_attr_neighbor = d.p_neighbor, d.edge_attr_neighbor
        else:
            x_focal, p_focal, x_neighbor = kwargv['x_focal'], kwargv['p_focal'], kwargv['x_neighbor']
            p_neighbor, edge_attr_neighbor = kwargv['p_neighbor'], kwargv['edge_attr_neighbor']
        if p_focal.shape[-1] != self.p_support.shape[-1]:
            raise Exception(f'data coordinates is of {p_focal.shape[-1]}D, but the kernel is '
                            f'{self.p_support.shape[-1]}D')
        n, deg, F = x_neighbor.shape
        if deg != self.x_support.shape[1]:
            raise Exception(f'neighbourhood has degree {deg} but the kernel has {self.x_support.shape[1]} supports')
        dev = x_focal.device
        # node table = [focal rows ; neighbour rows]; the bucket of this degree indexes into it
        table = torch.cat([x_focal, x_neighbor.reshape(n * deg, F)], dim=0)
        sel = [None] * 4
        nei = [None] * 4
        pf, pn, ea = [None] * 4, [None] * 4, [None] * 4
        sel[deg - 1] = torch.arange(n, device=dev, dtype=torch.int64)
        nei[deg - 1] = torch.arange(n * deg, device=dev, dtype=torch.int64) + n
        pf[deg - 1], pn[deg - 1], ea[deg - 1] = p_focal, p_neighbor, edge_attr_neighbor
        plan = BucketPlan.from_reference_tensors(n * (deg + 1), sel, nei, pf, pn, ea)
        params = [None] * 4
        params[deg - 1] = _param_dict(self)
        sc = KernelSetConvFn.apply(table, plan, params, self.edge_attr_support.shape[-1], bool(is_last_layer), None,
                                   None, *flat_params(params))
        return sc[:n].t()


class BaseKernelSetConv(Module):
    def __init__(self, fixed_kernelconv1=None, fixed_kernelconv2=None, fixed_kernelconv3=None,
                 fixed_kernelconv4=None, trainable_kernelconv1=None, trainable_kernelconv2=None,
                 trainable_kernelconv3=None, trainable_kernelconv4=None):
        super(BaseKernelSetConv, self).__init__()
        fixed = [fixed_kernelconv1, fixed_kernelconv2, fixed_kernelconv3, fixed_kernelconv4]
        train = [trainable_kernelconv1, trainable_kernelconv2, trainable_kernelconv3, trainable_kernelconv4]
        self.fixed_kernelconv_set = ModuleList(fixed)
        self.trainable_kernelconv_set = ModuleList(train)
        self.num_fixed_kernel_list = [k.get_num_kernels() if k is not None else None for k in fixed]
        self.num_trainable_kernel_list = [k.get_num_kernels() if k is not None else None for k in train]
        self.num_kernel_list = [(f or 0) + (t or 0) for f, t in
                                zip(self.num_fixed_kernel_list, self.num_trainable_kernel_list)]

    # ---- kept for API parity with the reference helpers -------------------------------------------------------
    def get_reorder_index(self, index):
        return torch.sort(index, dim=0)[1]

    def save_score(self, sc):
        """kernels.py:594-608: dump the [N,K] score matrix with kernel names as CSV (host side, pandas)."""
        import pandas as pd
        root = 'customized_kernels'
        sc_np = sc.cpu().detach().numpy()
        headers = []
        for i, file in enumerate(os.listdir(root)):
            headers += list(pd.read_csv(root + '/' + file)['name'])
            headers += ['std_kernel'] * self.num_trainable_kernel_list[i]
        pd.DataFrame(sc_np, columns=headers).transpose().to_csv('scores.csv')

    def _has_mixed_sets(self):
        return any(f is not None and t is not None for f, t in zip(self.fixed_kernelconv_set, self.trainable_kernelconv_set))

    def _degree_owners(self):
        """Per degree: the KernelConv module that owns the parameters (None for a degree without kernels)."""
        out = []
        for d in range(4):
            f, t = self.fixed_kernelconv_set[d], self.trainable_kernelconv_set[d]
            if f is not None and t is not None:
                raise NotImplementedError('the native stack driver takes one KernelConv per degree; a layer with fixed AND '
                                          'trainable kernels runs through BaseKernelSetConv.forward (two passes)')
            out.append(t if f is None else f)
        return out

    def _degree_params(self):
        """Per degree: concatenation [fixed ; trainable] kernels (kernels.py:702-715) as one parameter dict."""
        out = []
        for d in range(4):
            f, t = self.fixed_kernelconv_set[d], self.trainable_kernelconv_set[d]
            if f is None and t is None:
                out.append(None)
            elif f is None or t is None:
                out.append(_param_dict(t if f is None else f))
            else:
                raise NotImplementedError('one parameter set per degree expected here; BaseKernelSetConv.forward runs a layer '
                                          'with fixed AND trainable kernels as two passes')
        return out

    def forward(self, is_last_layer, *argv, **kwargv):
        if len(argv) != 0:
            raise Exception('Kernel does not take positional argument, use keyword argument instead. e.g. '
                            'model(data=data)')
        if 'data' in kwargv:   # reference: len(kwargv) == 2, i.e. data= and save_score= (kernels.py:622)
            g = lambda k: getattr(kwargv['data'], k)  # noqa: E731
        else:
            g = lambda k: kwargv[k]  # noqa: E731
        x = g('x')
        save_score = kwargv['save_score']
        if x.is_cuda and x.device.index != torch.cuda.current_device():
            with torch.cuda.device(x.device):        # the native calls take torch's CURRENT stream: run on the tensors' device
                return self.forward(is_last_layer, **kwargv)
        sel = [g(f'selected_index_deg{d}') for d in range(1, 5)]
        nei = [g(f'nei_index_deg{d}') for d in range(1, 5)]
        pf = [g(f'p_focal_deg{d}') for d in range(1, 5)]
        pn = [g(f'nei_p_deg{d}') for d in range(1, 5)]
        ea = [g(f'nei_edge_attr_deg{d}') for d in range(1, 5)]
        fixed, train = list(self.fixed_kernelconv_set), list(self.trainable_kernelconv_set)
        for d in range(4):
            if sel[d] is not None and sel[d].numel() and fixed[d] is None and train[d] is None:
                raise Exception(f'kernels.py::BaseKernelSet:both fixed and trainable kernelconv_set are None for '
                                f'degree {d + 1}')
        plan = kwargv.get('plan', None)
        if plan is None:
            plan = BucketPlan.from_reference_tensors(x.shape[0], sel, nei, pf, pn, ea)

        def run(params):
            Fe = next(int(p['edge_attr_support'].shape[-1]) for p in params if p is not None)
            return KernelSetConvFn.apply(x, plan, params, Fe, bool(is_last_layer), kwargv.get('argmax_in', None),
                                         kwargv.get('aux', None), *flat_params(params))

        if any(f is not None and t is not None for f, t in zip(fixed, train)):
            # fixed AND trainable kernels for a degree (kernels.py:702-715): every KernelConv carries its own three mixing
            # weights, so the two sets run as two passes of the same kernels and the score columns of a degree are laid
            # side by side as [fixed ; trainable]
            if kwargv.get('argmax_in', None) is not None or kwargv.get('aux', None) is not None:
                raise NotImplementedError('argmax_in / aux are not available for mixed fixed+trainable kernel sets')
            pf_, pt_ = [None if f is None else _param_dict(f) for f in fixed], [None if t is None else _param_dict(t)
                                                                               for t in train]
            sc_f, sc_t = run(pf_), run(pt_)
            cols, of, ot = [], 0, 0
            for d in range(4):
                Lf, Lt = self.num_fixed_kernel_list[d] or 0, self.num_trainable_kernel_list[d] or 0
                cols += [sc_f[:, of:of + Lf], sc_t[:, ot:ot + Lt]]
                of, ot = of + Lf, ot + Lt
            sc = torch.cat(cols, dim=1)
        else:
            sc = run(self._degree_params())
        if save_score:
            self.save_score(sc)
        return sc


class KernelSetConv(BaseKernelSetConv):
    """Convolution with kernels of degree 1 to 4 (kernels.py:754-781)."""

    def __init__(self, L1, L2, L3, L4, D, node_attr_dim, edge_attr_dim):
        self.L = [L1, L2, L3, L4]
        kcs = [KernelConv(L=L, D=D, num_supports=d, node_attr_dim=node_attr_dim, edge_attr_dim=edge_attr_dim)
               for d, L in enumerate(self.L, start=1)]
        super(KernelSetConv, self).__init__(trainable_kernelconv1=kcs[0], trainable_kernelconv2=kcs[1],
                                            trainable_kernelconv3=kcs[2], trainable_kernelconv4=kcs[3])

    def get_num_kernel(self):
        return sum(self.L)
