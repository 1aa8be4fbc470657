"""ctypes binding of the C ABI in include/roo_b200.h (kangaroo_b200/lib/libroo_b200.so).

There is no fallback: if the CUDA library is missing or fails to load, importing raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libroo_b200.so")

OK = 0
WIN_9x7, WIN_11x11, WIN_16x16 = 0, 1, 2
WORDS = {WIN_9x7: 1, WIN_11x11: 2, WIN_16x16: 4}
IMG_U8, IMG_F32 = 0, 1
POPC32_COMPAT, POPC64 = 0, 1
FP_DEFAULT, FP_REFERENCE, FP_IEEE = 0, 1, 2
TUNE_HSWEEP, TUNE_INSWEEP_COST, TUNE_STRIP_CTAS_PER_SM, TUNE_GUIDED_SCRATCH_MIB, TUNE_SOLO_GEOMETRY = 0, 1, 2, 3, 4
VOL_U16, VOL_F32, VOL_I32, VOL_U32, VOL_U8, VOL_ELEM = 0, 1, 2, 3, 4, 5
DISP_I8, DISP_F32 = 0, 1
PROF_KINDS = ("census", "cost", "sweep", "wta", "lrcheck", "vgroup") + tuple(f"pass{i}" for i in range(8))


def disp_padded(max_disp: int) -> int:
    """internal disparity count: maxDisp rounded up to 32 / 64 / 128 / 256 / 512"""
    return 32 if max_disp <= 32 else (64 if max_disp <= 64 else (128 if max_disp <= 128 else (256 if max_disp <= 256 else 512)))


class RooImage(C.Structure):
    """roo_image_t == roo::Image<T> (Image.h:617-620)."""
    _fields_ = [("pitch", C.c_size_t), ("ptr", C.c_void_p), ("w", C.c_size_t), ("h", C.c_size_t)]


class RooVolume(C.Structure):
    """roo_volume_t == roo::Volume<T> (Volume.h:363-369)."""
    _fields_ = [("pitch", C.c_size_t), ("ptr", C.c_void_p), ("w", C.c_size_t), ("h", C.c_size_t),
                ("img_pitch", C.c_size_t), ("d", C.c_size_t)]


class PipelineParams(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("max_disp", C.c_int), ("window", C.c_int), ("popc_mode", C.c_int),
                ("P1", C.c_float), ("P2", C.c_float), ("img_scale", C.c_float),
                ("dohoriz", C.c_int), ("dovert", C.c_int), ("doreverse", C.c_int), ("dodiag", C.c_int),
                ("subpix", C.c_int), ("lrcheck", C.c_int), ("lr_maxdiff", C.c_float),
                ("max_batch", C.c_int), ("keep_volume", C.c_int), ("fuse_vertical", C.c_int),
                ("median_size", C.c_int), ("median_maxbad", C.c_int), ("median_iters", C.c_int),
                ("fp_mode", C.c_int), ("filtgrad_threshold", C.c_float)]


# every symbol include/roo_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
_IMG, _VOL, _S = _P(RooImage), _P(RooVolume), C.c_void_p
SYMBOLS = {
    "roo_b200_version": (C.c_char_p, []),
    "roo_status_string": (C.c_char_p, [C.c_int]),
    "roo_launch_count": (C.c_ulonglong, []),
    "roo_set_ieee_division": (None, [C.c_int]),
    "roo_set_tuning": (C.c_int, [C.c_int, C.c_int]),
    "roo_census": (C.c_int, [_IMG, _IMG, C.c_int, C.c_int, _S]),
    "roo_census_stereo": (C.c_int, [_IMG, _IMG, _IMG, C.c_int, _S]),
    "roo_census_stereo_volume": (C.c_int, [_VOL, _IMG, _IMG, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _S]),
    "roo_sgm": (C.c_int, [_VOL, _VOL, C.c_int, _IMG, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                          C.c_int, _S]),
    "roo_costvol_minimum": (C.c_int, [_IMG, C.c_int, _VOL, C.c_int, C.c_uint, _S]),
    "roo_costvol_minimum_elem": (C.c_int, [_IMG, _VOL, _S]),
    "roo_costvol_minimum_subpix": (C.c_int, [_IMG, _VOL, C.c_uint, C.c_float, _S]),
    "roo_filter_disp_grad": (C.c_int, [_IMG, _IMG, C.c_float, _S]),
    "roo_costvol_minimum_square_penalty_subpix": (C.c_int, [_IMG, _VOL, _IMG, C.c_uint, C.c_float, C.c_float, C.c_float, _S]),
    "roo_bilateral_filter_joint": (C.c_int, [_IMG, _IMG, _IMG, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint, _S]),
    "roo_elementwise_multiply": (C.c_int, [_IMG, _IMG, _IMG, C.c_float, C.c_float, _S]),
    "roo_elementwise_division": (C.c_int, [_IMG, _IMG, _IMG, C.c_float, C.c_float, C.c_float, C.c_float, _S]),
    "roo_elementwise_square": (C.c_int, [_IMG, _IMG, C.c_float, C.c_float, _S]),
    "roo_elementwise_multiply_add": (C.c_int, [_IMG, _IMG, _IMG, _IMG, C.c_float, C.c_float, C.c_float, _S]),
    "roo_box_filter": (C.c_int, [_IMG, _IMG, C.c_int, _S]),
    "roo_guided_filter_volume": (C.c_int, [_VOL, _IMG, C.c_int, C.c_float, C.c_int, _S]),
    "roo_release_scratch": (C.c_int, []),
    "roo_dense_stereo": (C.c_int, [_IMG, C.c_int, _IMG, _IMG, C.c_int, C.c_float, C.c_int, _S]),
    "roo_bilateral_filter_volume": (C.c_int, [_VOL, _VOL, _IMG, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint, C.c_int, _S]),
    "roo_dense_stereo_subpixel_refine": (C.c_int, [_IMG, _IMG, _IMG, _IMG, _S]),
    "roo_left_right_check_f32": (C.c_int, [_IMG, _IMG, C.c_float, C.c_float, _S]),
    "roo_left_right_check_i8": (C.c_int, [_IMG, _IMG, C.c_int, C.c_int, _S]),
    "roo_elementwise_scale_bias": (C.c_int, [_IMG, _IMG, C.c_int, C.c_float, C.c_float, _S]),
    "roo_box_half": (C.c_int, [_IMG, _IMG, C.c_int, _S]),
    "roo_create_matlab_lookup_table": (C.c_int, [_IMG] + [C.c_float] * 6 + [_S]),
    "roo_create_matlab_lookup_table_homography": (C.c_int, [_IMG] + [C.c_float] * 6 + [_P(C.c_float), _S]),
    "roo_warp": (C.c_int, [_IMG, _IMG, _IMG, _S]),
    "roo_disp2depth": (C.c_int, [_IMG, _IMG, C.c_float, C.c_float, C.c_float, _S]),
    "roo_disparity_image_to_vbo": (C.c_int, [_IMG, _IMG, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _S]),
    "roo_costvol_from_stereo_truncated_abs_and_grad": (C.c_int, [_VOL, _IMG, _IMG, C.c_float, C.c_float, C.c_float,
                                                                 C.c_float, _S]),
    "roo_median_filter_reject_negative": (C.c_int, [_IMG, _IMG, C.c_int, C.c_int, _S]),
    "roo_engine_create": (C.c_int, [_P(C.c_void_p), _P(PipelineParams)]),
    "roo_engine_destroy": (C.c_int, [C.c_void_p]),
    "roo_engine_scratch_bytes": (C.c_size_t, [C.c_void_p]),
    "roo_engine_set_front_end": (C.c_int, [C.c_void_p, C.c_int, _IMG, _IMG]),
    "roo_engine_run_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _S]),
    "roo_engine_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "roo_engine_submit_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _P(C.c_longlong)]),
    "roo_engine_wait": (C.c_int, [C.c_void_p, C.c_longlong]),
    "roo_engine_export_volume": (C.c_int, [C.c_void_p, C.c_int, _VOL, _S]),
    "roo_engine_export_census": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _IMG, _S]),
    "roo_split_engine_create": (C.c_int, [_P(C.c_void_p), _P(PipelineParams), _P(C.c_int), C.c_int]),
    "roo_split_engine_destroy": (C.c_int, [C.c_void_p]),
    "roo_split_engine_strip_count": (C.c_int, [C.c_void_p]),
    "roo_split_engine_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "roo_split_engine_last_stats": (C.c_int, [C.c_void_p, _P(C.c_float), _P(C.c_ulonglong)]),
    "roo_multi_engine_create": (C.c_int, [_P(C.c_void_p), _P(PipelineParams), _P(C.c_int), C.c_int]),
    "roo_multi_engine_destroy": (C.c_int, [C.c_void_p]),
    "roo_multi_engine_device_count": (C.c_int, [C.c_void_p]),
    "roo_multi_engine_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "roo_engine_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "roo_engine_get_profile": (C.c_int, [C.c_void_p, _P(C.c_double), _P(C.c_longlong)]),
    "roo_engine_debug_counters": (C.c_int, [C.c_void_p, _P(C.c_ulonglong), C.c_int, C.c_int]),
}


class RooError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: {status_string(code)} ({code})")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Loads libroo_b200.so; raises if it is absent (no CPU path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C kangaroo_b200/csrc` "
                              "(or __graft_entry__.build()); kangaroo_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def status_string(code: int) -> str:
    return lib().roo_status_string(code).decode()


def check(code: int, what: str) -> None:
    if code != OK:
        raise RooError(code, what)


def launch_count() -> int:
    return int(lib().roo_launch_count())
