"""molkgnn_b200: B200-native (sm_100a) MolKGNN molecular-kernel convolution behind the reference's module API."""
from .kernels import KernelConv, BaseKernelSetConv, KernelSetConv  # noqa: F401
from .KernelLayer import MolGCN  # noqa: F401
from .plan import BucketPlan, ToXAndPAndEdgeAttrForDeg  # noqa: F401
from ._lib import MolKGNNError  # noqa: F401
from .MolKGNNNet import MolKGNNNet, global_add_pool  # noqa: F401
from .store import MoleculeStore, StoreLoader, load_split  # noqa: F401

__all__ = ["KernelConv", "BaseKernelSetConv", "KernelSetConv", "MolGCN", "BucketPlan", "ToXAndPAndEdgeAttrForDeg",
           "MolKGNNError", "MolKGNNNet", "global_add_pool", "MoleculeStore", "StoreLoader", "load_split"]
