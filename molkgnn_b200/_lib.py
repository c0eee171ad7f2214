"""ctypes binding of the C-ABI in include/molkgnn_b200.h.  No CPU fallback: if the shared library is missing or a
call fails this raises -- the product path is the CUDA path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MOLKGNN_B200_LIB: alternative build of the same library (the -DMK_PHASE_CLOCKS profiling variant), never a fallback
LIB_PATH = os.environ.get("MOLKGNN_B200_LIB") or os.path.join(HERE, "libmolkgnn_b200.so")

i32, i64, vp, fp = C.c_int32, C.c_int64, C.c_void_p, C.c_void_p


class Plan(C.Structure):
    _fields_ = [("N", i32), ("E", i32), ("n", i32 * 4), ("boff", i32 * 4), ("eoff", i32 * 4),
                ("deg", vp), ("pos", vp), ("sel", vp), ("nei", vp), ("nei_eid", vp), ("ehat", vp), ("tsign", vp),
                ("in_cnt", vp), ("in_src", vp), ("in_j", vp),
                ("tile_start", vp), ("n_tiles", i32), ("tile_max_nodes", i32), ("tile_max_deg", i32 * 4),
                ("tile_meta", vp), ("ehat_node", vp), ("node_tile", vp), ("tile_order", vp), ("tile_grid", i32)]


class Layer(C.Structure):
    _fields_ = [("F", i32), ("Fp", i32), ("Fe", i32), ("L", i32 * 4), ("koff", i32 * 4), ("K", i32),
                ("x_center", vp * 4), ("x_support", vp * 4), ("edge_attr_support", vp * 4), ("p_support", vp * 4),
                ("w_support", vp * 4), ("w_center", vp * 4), ("w_edge", vp * 4), ("packed", vp * 4),
                ("tile_img", vp)]


class LayerGrads(C.Structure):
    _fields_ = [("x_center", vp * 4), ("x_support", vp * 4), ("edge_attr_support", vp * 4),
                ("w_support", vp * 4), ("w_center", vp * 4), ("w_edge", vp * 4)]


MAX_LAYERS = 16


class StackLayout(C.Structure):
    _fields_ = [("fwd_bytes", i64), ("bwd_bytes", i64), ("grad_floats", i64),
                ("h", i64 * MAX_LAYERS), ("hnorm", i64 * (MAX_LAYERS + 1)), ("ximg", i64 * MAX_LAYERS),
                ("sc", i64 * MAX_LAYERS), ("argmax", i64 * MAX_LAYERS), ("argmax_free", i64 * MAX_LAYERS),
                ("argmax_tile", i64 * MAX_LAYERS),
                ("counter", i64), ("sc_elems", i64 * MAX_LAYERS), ("scoff", (i64 * 4) * MAX_LAYERS),
                ("coef", i64), ("partials", i64), ("scratch", i64), ("gx", i64 * 2),
                ("g_x_center", (i64 * 4) * MAX_LAYERS), ("g_x_support", (i64 * 4) * MAX_LAYERS),
                ("g_edge_attr_support", (i64 * 4) * MAX_LAYERS), ("g_w", (i64 * 4) * MAX_LAYERS),
                ("partials_alt", i64)]


EXPORTS = {
    "molkgnn_last_error": (C.c_char_p, []),
    "molkgnn_version": (C.c_int, []),
    "molkgnn_num_sms": (C.c_int, []),
    "molkgnn_launch_count": (i64, []),
    "molkgnn_bucket_scratch_bytes": (i64, [i32, i32]),
    "molkgnn_tile_meta_bytes": (i64, []),
    "molkgnn_struct_bytes": (i64, [i32]),
    "molkgnn_set_tile_order": (C.c_int, [C.c_int]),
    "molkgnn_tile_ximg_bytes": (i64, [C.POINTER(Plan), C.POINTER(Layer)]),
    "molkgnn_tile_ximg_build": (C.c_int, [C.POINTER(Plan), C.POINTER(Layer), vp, i32, vp, vp, vp]),
    "molkgnn_tile_ximg_build_raw": (C.c_int, [C.POINTER(Plan), C.POINTER(Layer), vp, i32, vp, vp, vp, vp]),
    "molkgnn_bucket_build": (C.c_int, [C.POINTER(Plan), vp, vp, i32, vp, i32, vp, vp]),
    "molkgnn_bucket_build_begin": (C.c_int, [C.POINTER(Plan), vp, vp, i32, vp, i32, vp, vp]),
    "molkgnn_bucket_build_finish": (C.c_int, [C.POINTER(Plan), vp, vp, i32, vp, i32, vp, vp]),
    "molkgnn_bucket_build_finish_ref": (C.c_int, [C.POINTER(Plan), vp, vp, i32, vp * 4, i64 * 4, i32, vp, vp, vp, vp]),
    "molkgnn_bucket_export": (C.c_int, [C.POINTER(Plan), i32, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp]),
    "molkgnn_plan_from_buckets": (C.c_int, [C.POINTER(Plan), vp * 4, vp * 4, vp * 4, vp * 4, i32, vp * 4, i32, vp]),
    "molkgnn_collate": (C.c_int, [vp, i32, i64, vp, vp, vp, i32, vp, i32, vp, i32, vp, i64, vp, i32, vp, vp, vp, vp, vp, vp, i64,
                                  vp, vp, vp, vp, vp]),
    "molkgnn_segment_sum": (C.c_int, [vp, i32, i32, vp, i32, vp, vp]),
    "molkgnn_pad_norm": (C.c_int, [vp, i32, i32, i32, vp, i32, vp, vp]),
    "molkgnn_packed_floats": (i64, [i32, i32, i32]),
    "molkgnn_param_pack": (C.c_int, [C.POINTER(Layer), vp]),
    "molkgnn_param_pack_layers": (C.c_int, [C.POINTER(Layer), i32, i32, vp]),
    "molkgnn_tile_img_bytes": (i64, [C.POINTER(Layer)]),
    "molkgnn_conv_fwd_smem_bytes": (i64, [C.POINTER(Layer)]),
    "molkgnn_conv_fwd": (C.c_int, [C.POINTER(Plan), C.POINTER(Layer), vp, i32, vp, i32, vp, i32, i32, i64 * 4, vp, vp,
                                   vp, vp, vp, vp, vp]),
    "molkgnn_tile_argmax_bytes": (i64, [C.POINTER(Plan), C.POINTER(Layer)]),
    "molkgnn_propagate_fwd": (C.c_int, [C.POINTER(Plan), C.POINTER(Layer), vp, i64 * 4, vp, i32, vp, vp, vp]),
    "molkgnn_conv_bwd_partial_floats": (i64, [C.POINTER(Plan), C.POINTER(Layer)]),
    "molkgnn_conv_bwd_coef_floats": (i64, [C.POINTER(Plan), C.POINTER(Layer)]),
    "molkgnn_set_fwd_path": (C.c_int, [C.c_int]),
    "molkgnn_get_fwd_path": (C.c_int, []),
    "molkgnn_tc_selftest": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "molkgnn_conv_bwd": (C.c_int, [C.POINTER(Plan), C.POINTER(Layer), vp, i32, vp, vp, i32, i32, vp, i64 * 4, vp, vp,
                                   vp, i32, C.POINTER(LayerGrads), i32, vp, vp, vp, vp]),
    "molkgnn_set_bwd_path": (C.c_int, [C.c_int]),
    "molkgnn_path_counts": (None, [i64 * 4]),
    "molkgnn_stack_layout": (C.c_int, [C.POINTER(Plan), C.POINTER(Layer), i32, i32, C.POINTER(StackLayout)]),
    "molkgnn_stack_fwd": (C.c_int, [C.POINTER(Plan), C.POINTER(Layer), i32, C.POINTER(StackLayout), i32, vp, i32, vp, vp,
                                    i32, C.POINTER(vp), C.POINTER(i32), vp]),
    "molkgnn_stack_bwd": (C.c_int, [C.POINTER(Plan), C.POINTER(Layer), i32, C.POINTER(StackLayout), vp, vp, vp, i32, vp,
                                    vp, C.POINTER(i32), vp]),
    "molkgnn_oneshot_create": (C.c_int, [i32, i32, i64, C.POINTER(vp)]),
    "molkgnn_oneshot_ipc_handle": (C.c_int, [vp, vp]),
    "molkgnn_oneshot_open": (C.c_int, [vp, vp]),
    "molkgnn_oneshot_allreduce": (C.c_int, [vp, vp, i64, i32, vp]),
    "molkgnn_oneshot_error": (C.c_int, [vp]),
    "molkgnn_oneshot_error_async": (C.c_int, [vp, vp, vp]),
    "molkgnn_oneshot_destroy": (C.c_int, [vp]),
    "molkgnn_profile_enable": (C.c_int, [C.c_int]),
    "molkgnn_profile_only": (C.c_int, [C.c_char_p]),
    "molkgnn_profile_read": (C.c_int, [C.c_char_p, C.c_int]),
}

_lib = None


class MolKGNNError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MolKGNNError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -m molkgnn_b200.build). "
                "molkgnn_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise MolKGNNError(lib().molkgnn_last_error().decode())


def ptr(t):
    """device (or host) pointer of a torch tensor, None -> NULL"""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (raw C getter: this runs several times per step)."""
    import torch
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
