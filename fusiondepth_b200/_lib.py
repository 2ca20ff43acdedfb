"""ctypes binding of libfusiondepth_b200.so (the C ABI in include/fusiondepth_b200.h).

There is no CPU fallback: every operator of the package goes through this library and fails
loudly if it cannot be loaded or if a call returns non-zero.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_long, c_size_t, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libfusiondepth_b200.so")

_lib = None


class FusionDepthLibraryError(RuntimeError):
    pass


class PhotolossDesc(ctypes.Structure):
    _fields_ = [
        ("B", c_int), ("H", c_int), ("W", c_int),
        ("color", (c_void_p * 4) * 3),
        ("disp", c_void_p * 4),
        ("K", c_void_p), ("inv_K", c_void_p),
        ("T", c_void_p * 2),
        ("noise", c_void_p * 4),
        ("beam", c_void_p),
        ("min_depth", c_float), ("max_depth", c_float),
        ("smoothness", c_float),
        ("si_thresh", c_float), ("si_var", c_float),
        ("si_scales", c_int),
        ("si_pred_mul", c_float), ("si_tgt_mul", c_float), ("si_lo", c_float), ("si_weight", c_float),
        ("sel", c_void_p),
        ("out_depth", c_void_p * 4),
        ("out_color", (c_void_p * 2) * 4),
        ("out_to_optimise", c_void_p * 4),
    ]


class Segment(ctypes.Structure):
    _fields_ = [("a", c_void_p), ("b", c_void_p), ("C", c_int), ("up", c_int)]


_I, _L, _F, _P = c_int, c_long, c_float, c_void_p
_D = ctypes.c_double

_SIGNATURES = {
    "fd_last_error": (c_char_p, []),
    "fd_version": (c_int, []),
    "fd_launch_count": (c_long, []),
    "fd_lidar_workspace_bytes": (c_size_t, [_I, _I, _I]),
    "fd_lidar_depth_map": (c_int, [_P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _P, _P]),
    "fd_lidar_pool_scale": (c_int, [_P, _I, _I, _I, _P, _P]),
    "fd_two_channel": (c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "fd_photoloss_workspace_bytes": (c_size_t, [_I, _I, _I]),
    "fd_photoloss_fwd": (c_int, [POINTER(PhotolossDesc), _P, _P, _P]),
    "fd_photoloss_bwd": (c_int, [POINTER(PhotolossDesc), _P, POINTER(c_void_p * 4), _P, _P, _P, _P]),
    "fd_upsample_bilinear_fwd": (c_int, [_P, _P, _L, _I, _I, _I, _I, _P]),
    "fd_upsample_bilinear_bwd": (c_int, [_P, _P, _L, _I, _I, _I, _I, _P]),
    "fd_backproject_fwd": (c_int, [_P, _P, _P, _I, _I, _I, _P]),
    "fd_backproject_bwd": (c_int, [_P, _P, _P, _I, _I, _I, _P]),
    "fd_project3d_fwd": (c_int, [_P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "fd_project3d_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _P]),
    "fd_grid_sample_border_fwd": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "fd_grid_sample_border_bwd": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "fd_pose_matrix_fwd": (c_int, [_P, _P, _I, _P, _I, _P]),
    "fd_pose_matrix_bwd": (c_int, [_P, _P, _I, _P, _P, _P, _I, _P]),
    "fd_cat_xy": (c_int, [_P, _P, _P, _I, _I, _I, _P]),
    "fd_ssim_fwd": (c_int, [_P, _P, _P, _L, _I, _I, _P]),
    "fd_ssim_bwd": (c_int, [_P, _P, _P, _P, _L, _I, _I, _P, _P]),
    "fd_refine_pack_workspace_bytes": (c_size_t, [_I, _I, _I]),
    "fd_refine_pack": (c_int, [_P, _P, _P, POINTER(c_void_p * 4), _I, _I, _I, _I, _I, _I, _I, _F, _F,
                               POINTER(c_void_p * 4), _P, _P, _P]),
    "fd_masked_median_workspace_bytes": (c_size_t, [_I, _I, _I]),
    "fd_depth_errors_workspace_bytes": (c_size_t, [_I, _I, _I]),
    "fd_depth_errors": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _I, _F, _F, _I, _I, _F, _F, _P, _P, _P]),
    "fd_masked_median": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P]),
    "fd_prep_input": (c_int, [_P, _P, _I, _I, _I, _I, _F, _F, _P]),
    "fd_stem_im2col": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "fd_pad_rows": (c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "fd_conv2d_fwd": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fd_conv2d_dgrad": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fd_conv2d_wgrad": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fd_weight_transpose": (c_int, [_P, _P, _I, _I, _I, _P]),
    "fd_conv2d_tc_supported": (c_int, [_I, _I]),
    "fd_tf32_split": (c_int, [_P, _P, _L, _P]),
    "fd_weight_transpose_split": (c_int, [_P, _P, _P, _I, _I, _I, _P]),
    "fd_weight_transpose_split_batched": (c_int, [_P, _P, _P, _P, _I, _L, _P]),
    "fd_conv2d_fwd_tc": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fd_conv2d_fwd_tc_stats": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "fd_conv2d_dgrad_tc": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fd_conv2d_wgrad_tc": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fd_act_bwd": (c_int, [_P, _P, _P, _P, _L, _I, _I, _P]),
    "fd_conv2d_cout1_supported": (c_int, [_I, _I, _I, _I, _I]),
    "fd_conv2d_cout1_fwd": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "fd_conv2d_cout1_dgrad": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "fd_conv2d_cout1_wgrad": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "fd_debug_set_conv_trace": (c_int, [_P]),
    "fd_conv2d_c16_supported": (c_int, [_I, _I, _I, _I, _I]),
    "fd_conv2d_c16_fwd": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "fd_conv2d_c16_dgrad": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "fd_conv2d_c16_wgrad": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "fd_bn_workspace_bytes": (c_size_t, [_I]),
    "fd_bn_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _I, _F, _F, _I, _P, _P, _P, _P, _L, _I, _F, _I, _P]),
    "fd_bn_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _L, _I, _I, _P]),
    "fd_bn_bwd_xmask_ok": (c_int, [_L, _I]),
    "fd_bn_bwd_xmask": (c_int, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _L, _I, _I, _P]),
    "fd_maxpool3x3s2_fwd": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "fd_maxpool3x3s2_bwd": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "fd_assemble_fwd": (c_int, [POINTER(Segment), _I, _P, _I, _I, _I, _I, _P]),
    "fd_assemble_bwd": (c_int, [_P, POINTER(c_void_p), POINTER(c_int), POINTER(c_int), _I, _I, _I, _I, _I, _P]),
    "fd_assemble_bwd2": (c_int, [_P, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_int), _I, _I, _I, _I,
                                 _I, _P]),
    "fd_add": (c_int, [_P, _P, _P, _L, _P]),
    "fd_add_relu": (c_int, [_P, _P, _P, _L, _P]),
    "fd_mean_hw_fwd": (c_int, [_P, _P, _I, _I, _I, _F, _P]),
    "fd_mean_hw_bwd": (c_int, [_P, _P, _I, _I, _I, _F, _P]),
    "fd_adam_step": (c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _P, _F, _P]),
    "fd_lanczos_ksize": (c_int, [_I, _I]),
    "fd_lanczos_coeffs": (c_int, [_I, _I, _P, _P]),
    "fd_resize_lanczos_u8": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P]),
    "fd_color_jitter_u8": (c_int, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "fd_image_to_tensor": (c_int, [_P, _P, _I, _I, _I, _P]),
    "fd_gdc_select": (c_int, [_P, _P, _I, _I, _P, _D, _D, _P, _P, _P]),
    "fd_gdc_knn": (c_int, [_P, _I, _I, _P, _P]),
    "fd_gdc_weights": (c_int, [_P, _P, _I, _I, _D, _P, _P]),
    "fd_gdc_rhs": (c_int, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "fd_gdc_apply": (c_int, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "fd_gdc_apply_t": (c_int, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "fd_gdc_dot": (c_int, [_P, _P, _I, _P, _P]),
    "fd_gdc_cg_update": (c_int, [_P, _P, _P, _P, _I, _P, _P]),
    "fd_gdc_cg_dir": (c_int, [_P, _P, _I, _P, _P]),
}

EXPORTS = tuple(_SIGNATURES)


def load(path: str = SO_PATH):
    """Loads the shared library and types every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FusionDepthLibraryError(
            "fusiondepth_b200: %s not found -- build it with `python -m fusiondepth_b200.build` "
            "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)         # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, who: str = ""):
    if rc != 0:
        msg = load().fd_last_error().decode(errors="replace")
        raise FusionDepthLibraryError("%s failed (rc=%d): %s" % (who or "fusiondepth_b200", rc, msg))


def launch_count() -> int:
    return int(load().fd_launch_count())
