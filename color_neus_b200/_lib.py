"""ctypes binding of libcneus.so (C ABI declared in include/cneus.h).

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no CPU or
PyTorch fallback: if the library is missing or an entry point fails, the caller gets an exception.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CNEUS_LIB", os.path.join(_HERE, "libcneus.so"))  # CNEUS_LIB: profiling builds (tools/)

ABI_VERSION = 1
MAX_SDF_LIN, MAX_COLOR_LIN, MAX_RELIGHT_LIN = 12, 8, 8
COLOR_MODES = {"idr": 0, "no_view_dir": 1, "no_normal": 2}


class CneusError(RuntimeError):
    pass


class NetDesc(C.Structure):
    _fields_ = [
        ("sdf_n_lin", C.c_int32), ("sdf_d_hidden", C.c_int32), ("sdf_d_out", C.c_int32), ("sdf_multires", C.c_int32),
        ("sdf_skip", C.c_int32), ("sdf_scale", C.c_float),
        ("color_mode", C.c_int32), ("color_n_lin", C.c_int32), ("color_d_hidden", C.c_int32),
        ("color_d_feature", C.c_int32), ("color_multires_view", C.c_int32), ("color_squeeze_out", C.c_int32),
        ("has_relight", C.c_int32), ("relight_n_layers", C.c_int32), ("relight_y_in_layer", C.c_int32),
        ("relight_d_hidden", C.c_int32), ("relight_multires_view", C.c_int32), ("relight_include_grad", C.c_int32),
        ("relight_inv_sigmoid", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


class Linear(C.Structure):
    _fields_ = [("weight_g", C.c_void_p), ("weight_v", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_int32),
                ("in_", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("sdf", Linear * MAX_SDF_LIN), ("color", Linear * MAX_COLOR_LIN), ("relight_in", Linear),
                ("relight_mlp", Linear * MAX_RELIGHT_LIN)]


RENDER_OUT_FIELDS = ["color_fine", "global_color", "weight_sum", "weight_max", "depth", "weights", "cdf",
                     "inside_sphere", "gradients", "delta_relight", "sdf", "sampled_color", "global_sampled", "alpha",
                     "mid_z", "dists", "scalars"]


class RenderOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in RENDER_OUT_FIELDS]


_lib = None


def lib():
    """Load libcneus.so once; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CneusError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no fallback path")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
    dp = C.POINTER(NetDesc)
    sig = {
        "cneus_abi_version": (C.c_int, []),
        "cneus_last_error": (C.c_char_p, []),
        "cneus_device_sm_count": (C.c_int, []),
        "cneus_packed_bytes": (sz, [dp]),
        "cneus_pack_weights": (C.c_int, [dp, C.POINTER(Params), vp, sz, vp]),
        "cneus_workspace_bytes": (sz, [dp, i64, i32, i64]),
        "cneus_sdf_forward": (C.c_int, [dp, vp, vp, i64, vp, i32, vp, sz, vp]),
        "cneus_sdf_gradient": (C.c_int, [dp, vp, vp, i64, vp, vp, sz, vp]),
        "cneus_color_forward": (C.c_int, [dp, vp, vp, vp, vp, vp, i64, vp, vp, sz, vp]),
        "cneus_relight_forward": (C.c_int, [dp, vp, vp, vp, vp, vp, i64, vp, vp, vp, sz, vp]),
        "cneus_up_sample": (C.c_int, [vp, vp, vp, vp, i64, i32, i32, f32, vp, vp, vp]),
        "cneus_cat_z_vals": (C.c_int, [dp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp, vp, sz, vp]),
        "cneus_sample_z": (C.c_int, [dp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp, sz, vp]),
        "cneus_render_core": (C.c_int, [dp, vp, vp, vp, vp, vp, i64, i32, f32, f32, C.POINTER(RenderOut), vp, sz, vp]),
        "cneus_sdf_grid": (C.c_int, [dp, vp, vp, vp, vp, i32, i64, i64, vp, vp, sz, vp]),
        "cneus_vertex_color": (C.c_int, [dp, vp, vp, i64, vp, vp, sz, vp]),
        "cneus_profile_enable": (None, [C.c_int]),
        "cneus_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
        "cneus_launch_count": (C.c_int64, []),
        "cneus_force_simt": (None, [C.c_int]),
        "cneus_embed": (C.c_int, [vp, i64, i32, i32, vp, vp]),
        "cneus_backward_chunk_rays": (None, [C.c_int]),
        "cneus_backward_fused_recompute": (None, [C.c_int]),
        "cneus_gen_rays": (C.c_int, [vp, i32, vp, i32, i32, vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "cneus_gemm_test_workspace_bytes": (sz, []),
        "cneus_gemm_test": (C.c_int, [C.c_int, vp, vp, vp, i64, i64, i64, i64, i64, i64, vp, C.c_int, vp, i64, C.c_int, C.c_int,
                                      vp, sz, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if L.cneus_abi_version() != ABI_VERSION:
        raise CneusError(f"libcneus.so ABI {L.cneus_abi_version()} != expected {ABI_VERSION}")
    _lib = L
    return L


EXPORTED = ["cneus_abi_version", "cneus_last_error", "cneus_device_sm_count", "cneus_packed_bytes",
            "cneus_pack_weights", "cneus_workspace_bytes", "cneus_sdf_forward", "cneus_sdf_gradient",
            "cneus_color_forward", "cneus_relight_forward", "cneus_up_sample", "cneus_cat_z_vals", "cneus_sample_z",
            "cneus_render_core", "cneus_sdf_grid", "cneus_vertex_color", "cneus_profile_enable", "cneus_profile_read",
            "cneus_launch_count", "cneus_force_simt", "cneus_backward_workspace_bytes",
            "cneus_render_backward", "cneus_tc_prof_enable", "cneus_tc_prof_read", "cneus_tc_prof_read_types", "cneus_gemm_test_workspace_bytes",
            "cneus_gemm_test", "cneus_gen_rays", "cneus_clip_adam_workspace_bytes", "cneus_clip_adam_step",
            "cneus_loss_workspace_bytes", "cneus_neus_loss", "cneus_mc_workspace_bytes", "cneus_mc_count", "cneus_mc_emit",
            "cneus_mc_tables", "cneus_gather_pixels_u8", "cneus_backward_chunk_rays",
            "cneus_backward_fused_recompute", "cneus_embed"]


def check(rc, what):
    if rc != 0:
        msg = lib().cneus_last_error()
        raise CneusError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise CneusError("expected a contiguous float32 CUDA tensor, got "
                         f"{type(t).__name__} {getattr(t, 'dtype', None)} cuda={getattr(t, 'is_cuda', None)}")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
