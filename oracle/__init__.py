"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of the CPU oracle (oracle/kangaroo_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (kangaroo_b200/) never does.

Arrays use numpy's C order: images are (H, W[, words]) and volumes are (D, H, W) -- the same
"x fastest, then y, then d" order as roo::Image / roo::Volume (Image.h:617-620, Volume.h:363-369).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkangaroo_oracle.so")

WIN_9x7, WIN_11x11, WIN_16x16 = 0, 1, 2
WORDS = {WIN_9x7: 1, WIN_11x11: 2, WIN_16x16: 4}
IMG_U8, IMG_F32 = 0, 1
POPC32_COMPAT, POPC64 = 0, 1
VOL_U16, VOL_F32, VOL_I32, VOL_U32, VOL_U8, VOL_ELEM = 0, 1, 2, 3, 4, 5
DISP_I8, DISP_F32 = 0, 1

COSTVOLELEM = np.dtype([("n", np.int32), ("sum", np.float32)])  # CostVolElem.h:10-19

_VOL_TYPES = {np.dtype(np.uint16): VOL_U16, np.dtype(np.float32): VOL_F32, np.dtype(np.int32): VOL_I32,
              np.dtype(np.uint32): VOL_U32, np.dtype(np.uint8): VOL_U8, COSTVOLELEM: VOL_ELEM}


class KoImage(C.Structure):
    _fields_ = [("pitch", C.c_size_t), ("ptr", C.c_void_p), ("w", C.c_size_t), ("h", C.c_size_t)]


class KoVolume(C.Structure):
    _fields_ = [("pitch", C.c_size_t), ("ptr", C.c_void_p), ("w", C.c_size_t), ("h", C.c_size_t),
                ("img_pitch", C.c_size_t), ("d", C.c_size_t)]


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc -O2 -fopenmp -ffp-contract=off). Building is not using."""
    src = os.path.join(_HERE, "kangaroo_oracle.c")
    hdr = os.path.join(_HERE, "kangaroo_oracle.h")
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        P = C.POINTER
        L.ko_num_threads.restype = C.c_int
        L.ko_set_num_threads.argtypes = [C.c_int]
        L.ko_census.argtypes = [P(KoImage), P(KoImage), C.c_int, C.c_int]
        L.ko_census_stereo.argtypes = [P(KoImage), P(KoImage), P(KoImage), C.c_int]
        L.ko_census_stereo_volume.argtypes = [P(KoVolume), P(KoImage), P(KoImage), C.c_int, C.c_int, C.c_int,
                                              C.c_float, C.c_int]
        L.ko_sgm.argtypes = [P(KoVolume), P(KoVolume), C.c_int, P(KoImage), C.c_int, C.c_int, C.c_float, C.c_float,
                             C.c_int, C.c_int, C.c_int, C.c_int]
        L.ko_costvol_minimum.argtypes = [P(KoImage), C.c_int, P(KoVolume), C.c_int, C.c_uint]
        L.ko_costvol_minimum_elem.argtypes = [P(KoImage), P(KoVolume)]
        L.ko_costvol_minimum_subpix.argtypes = [P(KoImage), P(KoVolume), C.c_uint, C.c_float, P(KoImage)]
        L.ko_dense_stereo_subpixel_refine.argtypes = [P(KoImage)] * 5
        L.ko_left_right_check_f32.argtypes = [P(KoImage), P(KoImage), C.c_float, C.c_float]
        L.ko_costvol_minimum_square_penalty_subpix.argtypes = [P(KoImage), P(KoVolume), P(KoImage), C.c_uint, C.c_float, C.c_float,
                                                               C.c_float, P(KoImage)]
        L.ko_filter_disp_grad.argtypes = [P(KoImage), P(KoImage), P(KoImage), C.c_float]
        L.ko_bilateral_filter_joint.argtypes = [P(KoImage), P(KoImage), P(KoImage), C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
        L.ko_left_right_check_i8.argtypes = [P(KoImage), P(KoImage), C.c_int, C.c_int]
        L.ko_dense_stereo.argtypes = [P(KoImage), P(KoImage), P(KoImage), C.c_int, C.c_int, C.c_float, C.c_int]
        L.ko_elementwise.argtypes = [C.c_int] + [P(KoImage)] * 4 + [C.c_float] * 4
        L.ko_box_filter.argtypes = [P(KoImage), P(KoImage), C.c_int]
        L.ko_guided_filter_volume.argtypes = [P(KoVolume), P(KoImage), C.c_int, C.c_float, C.c_int]
        L.ko_elementwise_scale_bias.argtypes = [P(KoImage), P(KoImage), C.c_int, C.c_float, C.c_float]
        L.ko_box_half.argtypes = [P(KoImage), P(KoImage), C.c_int]
        L.ko_disp2depth.argtypes = [P(KoImage), P(KoImage), C.c_float, C.c_float, C.c_float]
        L.ko_disparity_image_to_vbo.argtypes = [P(KoImage), P(KoImage)] + [C.c_float] * 5
        L.ko_median_filter_reject_negative.argtypes = [P(KoImage), P(KoImage), C.c_int, C.c_int]
        L.ko_warp.argtypes = [P(KoImage), P(KoImage), P(KoImage)]
        L.ko_create_matlab_lookup_table.argtypes = [P(KoImage)] + [C.c_float] * 6
        L.ko_create_matlab_lookup_table_h.argtypes = [P(KoImage)] + [C.c_float] * 6 + [P(C.c_float)]
        L.ko_costvol_abs_and_grad.argtypes = [P(KoVolume), P(KoImage), P(KoImage), C.c_float, C.c_float, C.c_float, C.c_float]
        L.ko_hamming.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ko_hamming.restype = C.c_uint
        L.ko_pipeline_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_void_p, C.c_void_p]
        L.ko_pipeline_u8.restype = C.c_int
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().ko_num_threads())


def set_num_threads(n: int) -> None:
    lib().ko_set_num_threads(int(n))


def use_all_cores() -> int:
    """All host cores this process may run on (torchrun sets OMP_NUM_THREADS=1 for multi-rank launches)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    set_num_threads(n)
    return num_threads()


def _img(a: np.ndarray) -> KoImage:
    """(H, W) or (H, W, words) C-contiguous-in-x array -> ko_image (row pitch from strides)."""
    assert a.ndim in (2, 3)
    if a.ndim == 3:
        assert a.strides[2] == a.itemsize and a.strides[1] == a.itemsize * a.shape[2]
    else:
        assert a.strides[1] == a.itemsize
    return KoImage(a.strides[0], a.ctypes.data, a.shape[1], a.shape[0])


def _vol(a: np.ndarray) -> KoVolume:
    assert a.ndim == 3 and a.strides[2] == a.itemsize
    return KoVolume(a.strides[1], a.ctypes.data, a.shape[2], a.shape[1], a.strides[0], a.shape[0])


def _img_type(a: np.ndarray) -> int:
    if a.dtype == np.uint8:
        return IMG_U8
    if a.dtype == np.float32:
        return IMG_F32
    raise TypeError(a.dtype)


def census(img: np.ndarray, window: int) -> np.ndarray:
    h, w = img.shape
    out = np.zeros((h, w, WORDS[window]), np.uint64)
    lib().ko_census(C.byref(_img(out)), C.byref(_img(img)), window, _img_type(img))
    return out


def hamming(p: np.ndarray, q: np.ndarray, popc_mode: int = POPC32_COMPAT) -> int:
    p = np.ascontiguousarray(p, np.uint64).ravel()
    q = np.ascontiguousarray(q, np.uint64).ravel()
    return int(lib().ko_hamming(p.ctypes.data, q.ctypes.data, p.size, popc_mode))


def census_stereo(left: np.ndarray, right: np.ndarray, max_disp: int) -> np.ndarray:
    h, w = left.shape[:2]
    disp = np.zeros((h, w), np.int8)
    lib().ko_census_stereo(C.byref(_img(disp)), C.byref(_img(left)), C.byref(_img(right)), max_disp)
    return disp


def census_stereo_volume(left: np.ndarray, right: np.ndarray, max_disp: int, sd: float, vol_dtype=np.float32,
                         popc_mode: int = POPC32_COMPAT, depth: int | None = None,
                         fill: float = 0.0) -> np.ndarray:
    h, w, words = left.shape
    vol = np.full((depth or max_disp, h, w), fill, np.dtype(vol_dtype))
    lib().ko_census_stereo_volume(C.byref(_vol(vol)), C.byref(_img(left)), C.byref(_img(right)), words,
                                  _VOL_TYPES[vol.dtype], max_disp, sd, popc_mode)
    return vol


def sgm(vol_c: np.ndarray, left: np.ndarray, max_disp: int, p1: float, p2: float, dohoriz=True, dovert=True,
        doreverse=True, dodiag=False) -> np.ndarray:
    vol_h = np.empty(vol_c.shape, np.float32)
    lib().ko_sgm(C.byref(_vol(vol_h)), C.byref(_vol(vol_c)), _VOL_TYPES[vol_c.dtype], C.byref(_img(left)),
                 _img_type(left), max_disp, p1, p2, int(dohoriz), int(dovert), int(doreverse), int(dodiag))
    return vol_h


def costvol_minimum(vol: np.ndarray, max_disp: int, disp_dtype=np.float32) -> np.ndarray:
    d, h, w = vol.shape
    disp = np.zeros((h, w), np.dtype(disp_dtype))
    dt = DISP_I8 if disp.dtype == np.int8 else DISP_F32
    lib().ko_costvol_minimum(C.byref(_img(disp)), dt, C.byref(_vol(vol)), _VOL_TYPES[vol.dtype], max_disp)
    return disp


def costvol_minimum_elem(vol: np.ndarray) -> np.ndarray:
    d, h, w = vol.shape
    disp = np.zeros((h, w), np.float32)
    lib().ko_costvol_minimum_elem(C.byref(_img(disp)), C.byref(_vol(vol)))
    return disp


def costvol_minimum_subpix(vol: np.ndarray, max_disp: int, sd: float):
    d, h, w = vol.shape
    disp = np.zeros((h, w), np.float32)
    mask = np.zeros((h, w), np.uint8)
    lib().ko_costvol_minimum_subpix(C.byref(_img(disp)), C.byref(_vol(vol)), max_disp, sd, C.byref(_img(mask)))
    return disp, mask


def costvol_minimum_square_penalty_subpix(vol: np.ndarray, lastd: np.ndarray, max_disp: int, sd: float, lam: float, theta: float):
    d, h, w = vol.shape
    out = np.zeros((h, w), np.float32)
    mask = np.zeros((h, w), np.uint8)
    lastd = np.ascontiguousarray(lastd, np.float32)
    lib().ko_costvol_minimum_square_penalty_subpix(C.byref(_img(out)), C.byref(_vol(vol)), C.byref(_img(lastd)), max_disp, sd,
                                                   lam, theta, C.byref(_img(mask)))
    return out, mask


def filter_disp_grad(grad_src: np.ndarray, img_in: np.ndarray, threshold: float) -> np.ndarray:
    """FilterDispGrad with the gradient taken of grad_src (what the output image held before the call)."""
    grad_src = np.ascontiguousarray(grad_src, np.float32)
    img_in = np.ascontiguousarray(img_in, np.float32)
    out = np.zeros_like(grad_src)
    lib().ko_filter_disp_grad(C.byref(_img(out)), C.byref(_img(grad_src)), C.byref(_img(img_in)), threshold)
    return out


def bilateral_filter_joint(img_in: np.ndarray, guide: np.ndarray, gs: float, gr: float, gc: float, size: int) -> np.ndarray:
    img_in = np.ascontiguousarray(img_in, np.float32)
    guide = np.ascontiguousarray(guide)
    out = np.zeros_like(img_in)
    lib().ko_bilateral_filter_joint(C.byref(_img(out)), C.byref(_img(img_in)), C.byref(_img(guide)),
                                    IMG_U8 if guide.dtype == np.uint8 else IMG_F32, gs, gr, gc, size)
    return out


def dense_stereo(left: np.ndarray, right: np.ndarray, max_disp: int, accept_thresh: float, score_rad: int, signed: bool = False) -> np.ndarray:
    """roo::DenseStereo<{unsigned char, char}, unsigned char> on tightly packed (h, w) uint8 images."""
    left, right = np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(right, np.uint8)
    out = np.zeros(left.shape, np.int8 if signed else np.uint8)
    lib().ko_dense_stereo(C.byref(_img(out)), C.byref(_img(left)), C.byref(_img(right)), 1 if signed else 0, max_disp, accept_thresh, score_rad)
    return out


EW_MULTIPLY, EW_DIVISION, EW_SQUARE, EW_MULTIPLY_ADD = 0, 1, 2, 3


def elementwise(op: int, a: np.ndarray, b=None, c=None, s0: float = 1.0, s1: float = 0.0, s2: float = 1.0, s3: float = 0.0) -> np.ndarray:
    """cu_operations.cu:91-181 on float images: Multiply s0*(a*b)+s1, Division s2*(a+s0)/(b+s1)+s3, Square (s0*a*a)+s1,
    MultiplyAdd s0*a*b + s1*c + s2."""
    a = np.ascontiguousarray(a, np.float32)
    out = np.zeros_like(a)
    ib = C.byref(_img(np.ascontiguousarray(b, np.float32))) if b is not None else None
    ic = C.byref(_img(np.ascontiguousarray(c, np.float32))) if c is not None else None
    lib().ko_elementwise(op, C.byref(_img(out)), C.byref(_img(a)), ib, ic, s0, s1, s2, s3)
    return out


def box_filter(img_in: np.ndarray, rad: int) -> np.ndarray:
    img_in = np.ascontiguousarray(img_in, np.float32)
    out = np.zeros_like(img_in)
    lib().ko_box_filter(C.byref(_img(out)), C.byref(_img(img_in)), rad)
    return out


def guided_filter_volume(vol: np.ndarray, guide: np.ndarray, rad: int, eps: float, max_disp: int | None = None) -> np.ndarray:
    """The applications' guided filtering of a (D, h, w) cost volume (stereo2/main.cpp:392-405); returns the filtered copy."""
    out = np.array(vol, np.float32, order="C", copy=True)
    guide = np.ascontiguousarray(guide, np.float32)
    lib().ko_guided_filter_volume(C.byref(_vol(out)), C.byref(_img(guide)), rad, eps, out.shape[0] if max_disp is None else max_disp)
    return out


def dense_stereo_subpixel_refine(disp: np.ndarray, left: np.ndarray, right: np.ndarray):
    h, w = disp.shape
    out = np.zeros((h, w), np.float32)
    mask = np.zeros((h, w), np.uint8)
    lib().ko_dense_stereo_subpixel_refine(C.byref(_img(out)), C.byref(_img(disp)), C.byref(_img(left)),
                                          C.byref(_img(right)), C.byref(_img(mask)))
    return out, mask


def left_right_check_f32(disp_l: np.ndarray, disp_r: np.ndarray, sd: float = -1.0, max_diff: float = 0.5):
    out = np.array(disp_l, np.float32, copy=True)
    lib().ko_left_right_check_f32(C.byref(_img(out)), C.byref(_img(disp_r)), sd, max_diff)
    return out


def left_right_check_i8(disp_l: np.ndarray, disp_r: np.ndarray, sd: int = -1, max_diff: int = 0):
    out = np.array(disp_l, np.int8, copy=True)
    lib().ko_left_right_check_i8(C.byref(_img(out)), C.byref(_img(disp_r)), sd, max_diff)
    return out


PIX_U8, PIX_F32, PIX_U16 = 0, 1, 2
_PIX = {np.dtype(np.uint8): PIX_U8, np.dtype(np.float32): PIX_F32, np.dtype(np.uint16): PIX_U16}


def elementwise_scale_bias(a: np.ndarray, s: float, offset: float = 0.0) -> np.ndarray:
    b = np.zeros(a.shape, np.float32)
    lib().ko_elementwise_scale_bias(C.byref(_img(b)), C.byref(_img(a)), _PIX[a.dtype], s, offset)
    return b


def box_half(img: np.ndarray) -> np.ndarray:
    h, w = img.shape
    out = np.zeros((h // 2, w // 2), img.dtype)
    lib().ko_box_half(C.byref(_img(out)), C.byref(_img(img)), _PIX[img.dtype])
    return out


def disp2depth(disp: np.ndarray, fu: float, baseline: float, min_disp: float = 0.0) -> np.ndarray:
    out = np.zeros(disp.shape, np.float32)
    lib().ko_disp2depth(C.byref(_img(disp)), C.byref(_img(out)), fu, baseline, min_disp)
    return out


def disparity_image_to_vbo(disp: np.ndarray, baseline: float, fu: float, fv: float, u0: float, v0: float) -> np.ndarray:
    h, w = disp.shape
    vbo = np.zeros((h, w, 4), np.float32)
    lib().ko_disparity_image_to_vbo(C.byref(_img(vbo)), C.byref(_img(disp)), baseline, fu, fv, u0, v0)
    return vbo


def costvol_abs_and_grad(left: np.ndarray, right: np.ndarray, depth: int, sd: float, alpha: float = 0.9,
                         r1: float = 0.03, r2: float = 0.008) -> np.ndarray:
    h, w = left.shape
    vol = np.zeros((depth, h, w), np.float32)
    lib().ko_costvol_abs_and_grad(C.byref(_vol(vol)), C.byref(_img(left)), C.byref(_img(right)), sd, alpha, r1, r2)
    return vol


def create_matlab_lookup_table(w: int, h: int, fu, fv, u0, v0, k1, k2) -> np.ndarray:
    lut = np.zeros((h, w, 2), np.float32)
    lib().ko_create_matlab_lookup_table(C.byref(_img(lut)), fu, fv, u0, v0, k1, k2)
    return lut


def create_matlab_lookup_table_h(w: int, h: int, fu, fv, u0, v0, k1, k2, H_on) -> np.ndarray:
    lut = np.zeros((h, w, 2), np.float32)
    Hc = (C.c_float * 9)(*[float(x) for x in np.asarray(H_on).ravel()])
    lib().ko_create_matlab_lookup_table_h(C.byref(_img(lut)), fu, fv, u0, v0, k1, k2, Hc)
    return lut


def warp(img: np.ndarray, lookup: np.ndarray) -> np.ndarray:
    h, w = lookup.shape[:2]
    out = np.zeros((h, w), np.uint8)
    lib().ko_warp(C.byref(_img(out)), C.byref(_img(img)), C.byref(_img(np.ascontiguousarray(lookup, np.float32))))
    return out


def median_filter_reject_negative(img: np.ndarray, size: int, maxbad: int) -> np.ndarray:
    out = np.zeros(img.shape, np.float32)
    lib().ko_median_filter_reject_negative(C.byref(_img(out)), C.byref(_img(img)), size, maxbad)
    return out


def pipeline_u8(left: np.ndarray, right: np.ndarray, max_disp: int, window: int = WIN_9x7,
                popc_mode: int = POPC32_COMPAT, p1: float = 0.01, p2: float = 0.02, dohoriz=True, dovert=True,
                doreverse=True, dodiag=False, subpix=False, lrcheck=False, lr_maxdiff: float = 1.0,
                want_volume: bool = False):
    """applications/stereo2/main.cpp:375-454 on a pair of (H, W) uint8 images."""
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    h, w = left.shape
    disp = np.zeros((h, w), np.float32)
    vol = np.zeros((max_disp, h, w), np.float32) if want_volume else None
    rc = lib().ko_pipeline_u8(left.ctypes.data, right.ctypes.data, w, h, max_disp, window, popc_mode, p1, p2,
                              int(dohoriz), int(dovert), int(doreverse), int(dodiag), int(subpix), int(lrcheck),
                              lr_maxdiff, disp.ctypes.data, vol.ctypes.data if want_volume else None)
    if rc != 0:
        raise MemoryError("ko_pipeline_u8 failed")
    return (disp, vol) if want_volume else disp
