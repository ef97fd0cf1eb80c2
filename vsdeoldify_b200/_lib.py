"""ctypes binding of libhavc_b200.so (see include/havc_b200.h).

The library is the product: there is no CPU / torch fallback.  Importing this module never touches
the GPU; `lib()` raises `HavcLibraryError` if the shared object is missing (run `__graft_entry__.build()`
or `vsdeoldify_b200/csrc/build.sh`).
"""
from __future__ import annotations

import ctypes as C
import os

HAVC_F16, HAVC_BF16, HAVC_F32 = 0, 1, 2
HAVC_MAX_TAPS = 16

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhavc_b200.so")


class HavcLibraryError(RuntimeError):
    pass


class HavcError(RuntimeError):
    pass


class ActView(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("C", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("B", C.c_int32), ("P", C.c_int32),
        ("stride_w", C.c_int64), ("stride_h", C.c_int64), ("stride_b", C.c_int64), ("stride_p", C.c_int64),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("src0", ActView), ("src1", ActView),
        ("weight", C.c_void_p),
        ("w_rows", C.c_int32), ("w_taps", C.c_int32), ("w_cin", C.c_int32), ("w_batches", C.c_int32),
        ("w_c1_off", C.c_int32),
        ("n_taps", C.c_int32),
        ("tap_dh", C.c_int8 * HAVC_MAX_TAPS), ("tap_dw", C.c_int8 * HAVC_MAX_TAPS),
        ("tap_p", C.c_int8 * HAVC_MAX_TAPS), ("tap_wi", C.c_int8 * HAVC_MAX_TAPS),
        ("out_B", C.c_int32), ("out_H", C.c_int32), ("out_W", C.c_int32),
        ("box_w", C.c_int32), ("box_h", C.c_int32), ("box_b", C.c_int32),
        ("a_batched", C.c_int32), ("b_batched", C.c_int32),
        ("BN", C.c_int32), ("N_total", C.c_int32),
        ("bias", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("relu1", C.c_int32), ("relu2", C.c_int32),
        ("residual", C.c_void_p),
        ("res_stride_w", C.c_int64), ("res_stride_h", C.c_int64), ("res_stride_b", C.c_int64),
        ("out", C.c_void_p),
        ("out_dtype", C.c_int32),
        ("out_stride_w", C.c_int64), ("out_stride_h", C.c_int64), ("out_stride_b", C.c_int64),
        ("up", C.c_int32), ("oy", C.c_int32), ("ox", C.c_int32),
        ("shuffle", C.c_int32), ("group_n", C.c_int32), ("c_store", C.c_int32),
        ("dtype", C.c_int32),
        ("src1_single_tap", C.c_int32), ("src1_wi", C.c_int32),
        ("split_n", C.c_int32),
        ("out2", C.c_void_p),
        ("out2_stride_w", C.c_int64), ("out2_stride_h", C.c_int64), ("out2_stride_b", C.c_int64),
        ("c_store2", C.c_int32),
        ("residual2", C.c_void_p),
        ("res2_stride_w", C.c_int64), ("res2_stride_h", C.c_int64), ("res2_stride_b", C.c_int64),
        ("head_w", C.c_void_p), ("head_out", C.c_void_p),
        ("head_stride_w", C.c_int64), ("head_stride_h", C.c_int64), ("head_stride_b", C.c_int64),
        ("tma_store", C.c_int32),
        ("leaky1", C.c_float),
        ("pair", C.c_int32),
        ("src0_lo", C.c_void_p), ("src1_lo", C.c_void_p), ("weight_lo", C.c_void_p), ("residual_lo", C.c_void_p),
        ("out_lo", C.c_void_p),
        ("blur", C.c_int32),
    ]


class HueRanges(C.Structure):
    _fields_ = [("n", C.c_int32), ("lo_deg", C.c_double * 8), ("hi_deg", C.c_double * 8)]


class PeriodicPlan(C.Structure):
    """havc_periodic_plan (include/havc_b200.h)."""
    _fields_ = [("ratio", C.c_int32), ("taps", C.c_int32), ("offset", C.c_int32), ("lo", C.c_int32), ("hi", C.c_int32), ("w", C.c_float * 72)]


# every symbol include/havc_b200.h declares: name -> (restype, argtypes)
_SIGNATURES = {
    "havc_last_error": (C.c_char_p, []),
    "havc_version": (C.c_int, []),
    "havc_launch_count": (C.c_int64, []),
    "havc_launch_count_reset": (None, []),
    "havc_conv_gemm": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "havc_im2col_small": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 11 + [C.c_void_p]),
    "havc_maxpool3x3s2": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p]),
    "havc_phase_split": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    "havc_affine_act": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "havc_blur2x2": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "havc_softmax_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p]),
    "havc_resample_v_f32_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p]),
    "havc_resample_h_periodic": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.POINTER(PeriodicPlan), C.c_void_p]),
    "havc_post_horizontal_periodic": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                C.c_void_p, C.c_int, C.c_int, C.POINTER(PeriodicPlan), C.c_void_p]),
    "havc_resample_h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_void_p]),
    "havc_pre_vertical": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "havc_gray_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
    "havc_head": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "havc_blend_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_void_p]),
    "havc_resample_v": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_void_p]),
    "havc_post_horizontal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "havc_frame_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "havc_chroma_stabilizer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                         C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "havc_red_fix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "havc_luma_masked_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double,
                                         C.c_double, C.c_float, C.c_void_p]),
    "havc_adaptive_luma_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double,
                                           C.c_double, C.c_double, C.c_double, C.c_void_p]),
    "havc_restore_color_gradient": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double,
                                              C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_double, C.c_int,
                                              C.c_void_p]),
    "havc_gray_mask_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "havc_restore_color": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double,
                                     C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "havc_average_frames_u8": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                         C.c_longlong, C.c_void_p]),
    "havc_adjust_chroma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(HueRanges), C.c_double, C.c_int,
                                     C.c_double, C.c_int, C.c_void_p]),
    "havc_chroma_tweak": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                    C.POINTER(HueRanges), C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                                    C.c_void_p]),
    "havc_image_tweak": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                   C.POINTER(HueRanges), C.c_void_p, C.c_void_p]),
    "havc_pil_resample_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p]),
    "havc_zhang_pre": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "havc_eccv_head": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "havc_zhang_tanh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_void_p]),
    "havc_bilinear_ab": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "havc_zhang_post": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "havc_chroma_post_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "havc_select_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "havc_vs_merge_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_void_p]),
    "havc_luma_adjusted_levels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double,
                                            C.c_double, C.c_double, C.c_double, C.c_void_p]),
    "havc_zimg_rgb_to_yuv420p8": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "havc_zimg_tweak_yuv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "havc_zimg_yuv420p8_to_rgb": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "havc_zimg_gray8_to_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "havc_zimg_inverse_matrix": (C.c_int, [C.c_void_p, C.c_int]),
}

_lib = None


def exported_symbols():
    return list(_SIGNATURES)


def lib():
    """Load libhavc_b200.so (once) and attach prototypes.  Fails loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HavcLibraryError(
            f"{LIB_PATH} not found: the CUDA library is the product and there is no fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    try:
        l = C.CDLL(LIB_PATH)
    except OSError as e:  # e.g. libcudart missing
        raise HavcLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(l, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().havc_last_error().decode("utf-8", "replace")
        raise HavcError(f"{what or 'havc call'} failed ({rc}): {msg}")
