"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference kernels (oracle/_ref/libkangaroo_ref.so,
built by oracle/Makefile from /root/reference/src) on the current CUDA device.

numpy in, numpy out; device memory comes from torch.  Works only for w <= 1024 and h <= 1024
(the reference launches one thread per row/column element in a single block).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libkangaroo_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(SO)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(SO)
        z, p, i, f, u = C.c_size_t, C.c_void_p, C.c_int, C.c_float, C.c_uint
        L.kref_census.argtypes = [p, z, p, z, z, z, i, i]
        L.kref_census_stereo.argtypes = [p, z, p, p, z, z, z, i]
        L.kref_census_stereo_volume.argtypes = [p, z, z, z, p, p, z, z, z, i, i, i, f]
        L.kref_sgm.argtypes = [p, p, z, z, z, z, p, z, z, z, z, i, i, f, f, i, i, i]
        L.kref_costvol_minimum.argtypes = [p, z, p, z, z, z, z, z, i, i, u]
        L.kref_costvol_minimum_elem.argtypes = [p, z, p, z, z, z, z, z]
        L.kref_costvol_minimum_subpix.argtypes = [p, z, p, z, z, z, z, z, u, f]
        L.kref_dense_stereo_subpixel_refine.argtypes = [p, z, p, p, p, z, z, z]
        L.kref_left_right_check_f32.argtypes = [p, p, z, z, z, f, f]
        L.kref_left_right_check_i8.argtypes = [p, p, z, z, z, i, i]
        L.kref_elementwise_scale_bias.argtypes = [p, z, p, z, z, z, i, f, f]
        L.kref_box_half.argtypes = [p, z, p, z, z, z, i]
        L.kref_disp2depth.argtypes = [p, p, z, z, z, f, f, f]
        L.kref_disparity_image_to_vbo.argtypes = [p, z, p, z, z, z, f, f, f, f, f]
        L.kref_median_reject_negative.argtypes = [p, p, z, z, z, i, i]
        L.kref_warp.argtypes = [p, z, p, z, z, z, p, z, z, z]
        L.kref_create_matlab_lookup_table.argtypes = [p, z, z, z, f, f, f, f, f, f]
        L.kref_create_matlab_lookup_table_h.argtypes = [p, z, z, z, f, f, f, f, f, f, C.POINTER(C.c_float)]
        L.kref_costvol_abs_and_grad.argtypes = [p, z, z, z, p, p, z, z, z, f, f, f, f]
        _lib = L
    return _lib


def _dev(a: np.ndarray):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()


def _back(t, dtype, shape) -> np.ndarray:
    return t.cpu().numpy().view(dtype).reshape(shape).copy()


def _ck(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"reference {what} failed: code {rc}")


def census(img: np.ndarray, window: int) -> np.ndarray:
    import torch
    h, w = img.shape
    words = {0: 1, 1: 2, 2: 4}[window]
    d_in = _dev(img)
    d_out = torch.zeros(h * w * words * 8, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_census(d_out.data_ptr(), w * words * 8, d_in.data_ptr(), w * img.itemsize, w, h, window,
                          0 if img.dtype == np.uint8 else 1), "Census")
    return _back(d_out, np.uint64, (h, w, words))


def census_stereo(left: np.ndarray, right: np.ndarray, max_disp: int) -> np.ndarray:
    import torch
    h, w = left.shape[:2]
    dl, dr = _dev(left), _dev(right)
    out = torch.zeros(h * w, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_census_stereo(out.data_ptr(), w, dl.data_ptr(), dr.data_ptr(), w * 8, w, h, max_disp),
        "CensusStereo")
    return _back(out, np.int8, (h, w))


def census_stereo_volume(left, right, max_disp: int, sd: float, vol_dtype=np.float32, depth=None, fill=0.0):
    import torch
    h, w, words = left.shape
    vol_dtype = np.dtype(vol_dtype)
    depth = depth or max_disp
    dl, dr = _dev(left), _dev(right)
    vol = _dev(np.full((depth, h, w), fill, vol_dtype))
    _ck(lib().kref_census_stereo_volume(vol.data_ptr(), w * vol_dtype.itemsize, w * h * vol_dtype.itemsize, depth,
                                        dl.data_ptr(), dr.data_ptr(), w * words * 8, w, h, words,
                                        1 if vol_dtype == np.float32 else 0, max_disp, sd), "CensusStereoVolume")
    return _back(vol, vol_dtype, (depth, h, w))


def sgm(vol_c: np.ndarray, left: np.ndarray, max_disp: int, p1: float, p2: float, dohoriz=True, dovert=True,
        doreverse=True) -> np.ndarray:
    import torch
    d, h, w = vol_c.shape
    elem = vol_c.dtype.itemsize == 8
    dc, dleft = _dev(vol_c), _dev(left)
    dh = torch.full((d * h * w * 4,), 0x7F, dtype=torch.uint8, device="cuda")  # garbage: SGM must memset
    cs = vol_c.dtype.itemsize
    _ck(lib().kref_sgm(dh.data_ptr(), dc.data_ptr(), w * 4, w * h * 4, w * cs, w * h * cs, dleft.data_ptr(),
                       w * left.itemsize, w, h, d, 1 if elem else 0, max_disp, p1, p2, int(dohoriz), int(dovert),
                       int(doreverse)), "SemiGlobalMatching")
    return _back(dh, np.float32, (d, h, w))


_VT = {np.dtype(np.float32): 0, np.dtype(np.int32): 1, np.dtype(np.uint32): 2, np.dtype(np.uint16): 3,
       np.dtype(np.uint8): 4}


def costvol_minimum(vol: np.ndarray, max_disp: int, disp_dtype=np.float32) -> np.ndarray:
    import torch
    d, h, w = vol.shape
    disp_dtype = np.dtype(disp_dtype)
    dv = _dev(vol)
    out = torch.zeros(h * w * disp_dtype.itemsize, dtype=torch.uint8, device="cuda")
    vs = vol.dtype.itemsize
    _ck(lib().kref_costvol_minimum(out.data_ptr(), w * disp_dtype.itemsize, dv.data_ptr(), w * vs, w * h * vs, w, h,
                                   d, 0 if disp_dtype == np.int8 else 1, _VT[vol.dtype], max_disp), "CostVolMinimum")
    return _back(out, disp_dtype, (h, w))


def costvol_minimum_elem(vol: np.ndarray) -> np.ndarray:
    import torch
    d, h, w = vol.shape
    dv = _dev(vol)
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_costvol_minimum_elem(out.data_ptr(), w * 4, dv.data_ptr(), w * 8, w * h * 8, w, h, d),
        "CostVolMinimum(elem)")
    return _back(out, np.float32, (h, w))


def costvol_minimum_subpix(vol: np.ndarray, max_disp: int, sd: float) -> np.ndarray:
    import torch
    d, h, w = vol.shape
    dv = _dev(vol)
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_costvol_minimum_subpix(out.data_ptr(), w * 4, dv.data_ptr(), w * 4, w * h * 4, w, h, d, max_disp,
                                          sd), "CostVolMinimumSubpix")
    return _back(out, np.float32, (h, w))


def dense_stereo_subpixel_refine(disp: np.ndarray, left: np.ndarray, right: np.ndarray, pad: int = 8):
    """The reference reads outside the images near the borders (Q8): the three inputs are embedded in
    zero-padded parents so that those reads stay inside the allocation."""
    import torch
    h, w = disp.shape

    def emb(a):
        p = np.zeros((h + 2 * pad, w + 512), np.uint8)
        p[pad:pad + h, 256:256 + w] = a
        return p

    pitch = w + 512
    off = pad * pitch + 256
    dd, dl, dr = _dev(emb(disp)), _dev(emb(left)), _dev(emb(right))
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_dense_stereo_subpixel_refine(out.data_ptr(), w * 4, dd.data_ptr() + off, dl.data_ptr() + off,
                                                dr.data_ptr() + off, pitch, w, h), "DenseStereoSubpixelRefine")
    return _back(out, np.float32, (h, w))


def left_right_check_f32(disp_l, disp_r, sd=-1.0, max_diff=0.5) -> np.ndarray:
    h, w = disp_l.shape
    dl, dr = _dev(disp_l.astype(np.float32)), _dev(disp_r.astype(np.float32))
    _ck(lib().kref_left_right_check_f32(dl.data_ptr(), dr.data_ptr(), w * 4, w, h, sd, max_diff), "LeftRightCheck")
    return _back(dl, np.float32, (h, w))


def left_right_check_i8(disp_l, disp_r, sd=-1, max_diff=0) -> np.ndarray:
    h, w = disp_l.shape
    dl, dr = _dev(disp_l.astype(np.int8)), _dev(disp_r.astype(np.int8))
    _ck(lib().kref_left_right_check_i8(dl.data_ptr(), dr.data_ptr(), w, w, h, sd, max_diff), "LeftRightCheck<char>")
    return _back(dl, np.int8, (h, w))


# ---- callers either side of the path (front end / back end)

_PIX = {np.dtype(np.uint8): 0, np.dtype(np.float32): 1, np.dtype(np.uint16): 2}


def elementwise_scale_bias(a: np.ndarray, s: float, offset: float = 0.0) -> np.ndarray:
    import torch
    h, w = a.shape
    da = _dev(a)
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_elementwise_scale_bias(out.data_ptr(), w * 4, da.data_ptr(), w * a.itemsize, w, h, _PIX[a.dtype], s,
                                          offset), "ElementwiseScaleBias")
    return _back(out, np.float32, (h, w))


def box_half(img: np.ndarray) -> np.ndarray:
    import torch
    h, w = img.shape
    assert h % 2 == 0 and w % 2 == 0
    di = _dev(img)
    out = torch.zeros((h // 2) * (w // 2) * img.itemsize, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_box_half(out.data_ptr(), (w // 2) * img.itemsize, di.data_ptr(), w * img.itemsize, w // 2, h // 2,
                            _PIX[img.dtype]), "BoxHalf")
    return _back(out, img.dtype, (h // 2, w // 2))


def disp2depth(disp: np.ndarray, fu: float, baseline: float, min_disp: float = 0.0) -> np.ndarray:
    import torch
    h, w = disp.shape
    di = _dev(disp)
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_disp2depth(di.data_ptr(), out.data_ptr(), w * 4, w, h, fu, baseline, min_disp), "Disp2Depth")
    return _back(out, np.float32, (h, w))


def disparity_image_to_vbo(disp: np.ndarray, baseline: float, fu: float, fv: float, u0: float, v0: float) -> np.ndarray:
    import torch
    h, w = disp.shape
    di = _dev(disp)
    out = torch.zeros(h * w * 16, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_disparity_image_to_vbo(out.data_ptr(), w * 16, di.data_ptr(), w * 4, w, h, baseline, fu, fv, u0, v0),
        "DisparityImageToVbo")
    return _back(out, np.float32, (h, w, 4))


def median_filter_reject_negative(img: np.ndarray, size: int, maxbad: int) -> np.ndarray:
    """MedianFilterRejectNegative{5x5,7x7,9x9}, out of place (the applications call it in place, which races)."""
    import torch
    h, w = img.shape
    di = _dev(img)
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_median_reject_negative(out.data_ptr(), di.data_ptr(), w * 4, w, h, size, maxbad),
        "MedianFilterRejectNegative")
    return _back(out, np.float32, (h, w))


def costvol_minimum_square_penalty_subpix(vol: np.ndarray, lastd: np.ndarray, max_disp: int, sd: float, lam: float,
                                          theta: float) -> np.ndarray:
    import torch
    d, h, w = vol.shape
    dv, dd = _dev(vol), _dev(np.ascontiguousarray(lastd, np.float32))
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    lib().kref_costvol_minimum_square_penalty_subpix.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                                 C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint,
                                                                 C.c_float, C.c_float, C.c_float]
    _ck(lib().kref_costvol_minimum_square_penalty_subpix(out.data_ptr(), dv.data_ptr(), w * 4, w * h * 4, d, dd.data_ptr(), w * 4,
                                                         w, h, max_disp, sd, lam, theta), "CostVolMinimumSquarePenaltySubpix")
    return _back(out, np.float32, (h, w))


def filter_disp_grad(grad_src: np.ndarray, img_in: np.ndarray, threshold: float, margin: int = 16) -> np.ndarray:
    """FilterDispGrad OUT OF PLACE: the output image is pre-filled with grad_src (whose central differences the kernel
    reads), the values come from img_in.  Both live inside a larger zero-filled allocation because the kernel reads one
    pixel beyond every border unguarded; returns the (h, w) interior."""
    import torch
    h, w = grad_src.shape
    assert h % 16 == 0 and w % 16 == 0

    def emb(a):
        big = np.zeros((h + 2 * margin, w + 2 * margin), np.float32)
        big[margin:-margin, margin:-margin] = a
        return big
    W = w + 2 * margin
    do, di = _dev(emb(grad_src)), _dev(emb(img_in))
    off = (margin * W + margin) * 4
    lib().kref_filter_disp_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float]
    _ck(lib().kref_filter_disp_grad(do.data_ptr() + off, di.data_ptr() + off, W * 4, w, h, threshold), "FilterDispGrad")
    big = _back(do, np.float32, (h + 2 * margin, W))
    return np.ascontiguousarray(big[margin:-margin, margin:-margin])


def bilateral_filter_joint(img_in: np.ndarray, guide: np.ndarray, gs: float, gr: float, gc: float, size: int) -> np.ndarray:
    import torch
    h, w = img_in.shape
    di, dg = _dev(np.ascontiguousarray(img_in, np.float32)), _dev(np.ascontiguousarray(guide))
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    lib().kref_bilateral_joint.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t,
                                           C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_uint]
    _ck(lib().kref_bilateral_joint(out.data_ptr(), di.data_ptr(), w * 4, dg.data_ptr(), w * guide.itemsize,
                                   0 if guide.dtype == np.uint8 else 1, w, h, gs, gr, gc, size), "BilateralFilter")
    return _back(out, np.float32, (h, w))


def warp(img: np.ndarray, lookup: np.ndarray) -> np.ndarray:
    """roo::Warp: lookup is (h, w, 2) float32 = (x, y) sample positions inside [0, W-2] x [0, H-2] of img."""
    import torch
    h, w = lookup.shape[:2]
    ih, iw = img.shape
    di, dl = _dev(img), _dev(lookup)
    out = torch.zeros(h * w, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_warp(out.data_ptr(), w, di.data_ptr(), iw, iw, ih, dl.data_ptr(), w * 8, w, h), "Warp")
    return _back(out, np.uint8, (h, w))


def costvol_abs_and_grad(left: np.ndarray, right: np.ndarray, depth: int, sd: float, alpha: float, r1: float, r2: float,
                         margin: int = 1):
    """CostVolumeFromStereoTruncatedAbsAndGrad on float images embedded in a zero margin: the kernel reads row[x-1] /
    row[x+1] unguarded (Image.h:367-372), so the images are the interior of a larger allocation."""
    import torch
    h, w = left.shape

    def emb(a):
        big = np.zeros((h + 2 * margin, w + 2 * margin), np.float32)
        big[margin:-margin, margin:-margin] = a
        return big
    bl, br = emb(left), emb(right)
    dl, dr = _dev(bl), _dev(br)
    pitch = (w + 2 * margin) * 4
    off = margin * pitch + margin * 4
    vol = torch.zeros(depth * h * w * 4, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_costvol_abs_and_grad(vol.data_ptr(), w * 4, w * h * 4, depth, dl.data_ptr() + off, dr.data_ptr() + off,
                                        pitch, w, h, sd, alpha, r1, r2), "CostVolumeFromStereoTruncatedAbsAndGrad")
    return _back(vol, np.float32, (depth, h, w))


def create_matlab_lookup_table(w: int, h: int, fu, fv, u0, v0, k1, k2) -> np.ndarray:
    import torch
    out = torch.zeros(h * w * 8, dtype=torch.uint8, device="cuda")
    _ck(lib().kref_create_matlab_lookup_table(out.data_ptr(), w * 8, w, h, fu, fv, u0, v0, k1, k2), "CreateMatlabLookupTable")
    return _back(out, np.float32, (h, w, 2))


def create_matlab_lookup_table_h(w: int, h: int, fu, fv, u0, v0, k1, k2, H_on) -> np.ndarray:
    import torch
    out = torch.zeros(h * w * 8, dtype=torch.uint8, device="cuda")
    Hc = (C.c_float * 9)(*[float(x) for x in np.asarray(H_on).ravel()])
    _ck(lib().kref_create_matlab_lookup_table_h(out.data_ptr(), w * 8, w, h, fu, fv, u0, v0, k1, k2, Hc),
        "CreateMatlabLookupTable(H)")
    return _back(out, np.float32, (h, w, 2))


def box_filter(img_in: np.ndarray, rad: int) -> np.ndarray:
    """roo::BoxFilter<float,float,float> (cu_integral_image.h:26-38) of a dense (h, w) float32 image."""
    import torch
    h, w = img_in.shape
    di = _dev(np.ascontiguousarray(img_in, np.float32))
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    lib().kref_box_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
    _ck(lib().kref_box_filter(out.data_ptr(), di.data_ptr(), w, h, rad), "BoxFilter")
    return _back(out, np.float32, (h, w))


def guided_filter_volume(vol: np.ndarray, guide: np.ndarray, rad: int, eps: float) -> np.ndarray:
    """The applications' per-slice ComputeCovariance + GuidedFilter loop (stereo2/main.cpp:392-405) over a dense (D, h, w)
    float32 volume, guide = (h, w) float32."""
    D, h, w = vol.shape
    dv, dg = _dev(np.ascontiguousarray(vol, np.float32)), _dev(np.ascontiguousarray(guide, np.float32))
    lib().kref_guided_filter_volume.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_float]
    _ck(lib().kref_guided_filter_volume(dv.data_ptr(), dg.data_ptr(), w, h, D, rad, eps), "GuidedFilter")
    return _back(dv, np.float32, (D, h, w))


def elementwise(op: int, a: np.ndarray, b: np.ndarray | None, c: np.ndarray | None, s0=1.0, s1=0.0, s2=1.0, s3=0.0) -> np.ndarray:
    """op 0 Multiply(s0 = scalar, s1 = offset), 1 Division(s0 = sa, s1 = sb, s2 = scalar, s3 = offset), 2 Square(s0, s1),
    3 MultiplyAdd(s0 = sab, s1 = sc, s2 = offset) on dense float32 images (cu_operations.cu:85-190)."""
    import torch
    h, w = a.shape
    da = _dev(np.ascontiguousarray(a, np.float32))
    db = _dev(np.ascontiguousarray(b, np.float32)) if b is not None else da
    dc = _dev(np.ascontiguousarray(c, np.float32)) if c is not None else da
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    lib().kref_elementwise.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_size_t] * 2 + [C.c_float] * 4
    _ck(lib().kref_elementwise(op, out.data_ptr(), da.data_ptr(), db.data_ptr(), dc.data_ptr(), w, h, s0, s1, s2, s3), "Elementwise")
    return _back(out, np.float32, (h, w))


def dense_stereo(left: np.ndarray, right: np.ndarray, max_disp: int, accept_thresh: float, score_rad: int, signed: bool = False) -> np.ndarray:
    """roo::DenseStereo<{unsigned char, char}, unsigned char> (cu_dense_stereo.cu:209-253,376-406) on dense (h, w) uint8 images.
    Candidates left of the image read the bytes that precede the row (raw access): the previous row's tail in these tightly
    packed buffers -- inside the allocation, since the kernel only scores rows y >= 2 rad + 1."""
    import torch
    h, w = left.shape
    dl, dr = _dev(np.ascontiguousarray(left, np.uint8)), _dev(np.ascontiguousarray(right, np.uint8))
    out = torch.zeros(h * w, dtype=torch.uint8, device="cuda")
    lib().kref_dense_stereo.argtypes = [C.c_void_p] * 3 + [C.c_size_t] * 2 + [C.c_int, C.c_int, C.c_float, C.c_int]
    _ck(lib().kref_dense_stereo(out.data_ptr(), dl.data_ptr(), dr.data_ptr(), w, h, 1 if signed else 0, max_disp, accept_thresh, score_rad),
        "DenseStereo")
    return _back(out, np.int8 if signed else np.uint8, (h, w))
