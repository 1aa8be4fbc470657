"""Python mirror of the reference's operator surface for the census / SGM path.

Same names and argument meaning as namespace roo (include/kangaroo/cu_census.h:12-38,
cu_semi_global_matching.h:10-12, cu_dense_stereo.h:13-47,81-85); every call goes through the C ABI
(include/roo_b200.h) into the CUDA library.  torch is used only to own device memory and streams.
Like the reference launchers the calls are asynchronous; unlike them a failure raises RooError.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import (DISP_F32, DISP_I8, IMG_F32, IMG_U8, POPC32_COMPAT, POPC64, VOL_ELEM, VOL_F32, VOL_I32, VOL_U16,
                   VOL_U32, VOL_U8, WIN_9x7, WIN_11x11, WIN_16x16, WORDS, check, lib)

COSTVOLELEM = np.dtype([("n", np.int32), ("sum", np.float32)])  # CostVolElem.h:10-19
ULONG = np.dtype(np.uint64)                                     # unsigned long
ULONG2 = np.dtype((np.uint64, (2,)))                            # ulong2
ULONG4 = np.dtype((np.uint64, (4,)))                            # ulong4

_VOL_TYPES = {np.dtype(np.uint16): VOL_U16, np.dtype(np.float32): VOL_F32, np.dtype(np.int32): VOL_I32,
              np.dtype(np.uint32): VOL_U32, np.dtype(np.uint8): VOL_U8, COSTVOLELEM: VOL_ELEM}


def _stream(stream) -> int:
    if stream is None:
        return torch.cuda.current_stream().cuda_stream
    return stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)


def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class Image:
    """roo::Image<T, TargetDevice, Manage>: a pitched device image (Image.h:43-44, 77-83)."""

    def __init__(self, w: int, h: int, dtype, pitch: int | None = None, device=None):
        self.dtype = np.dtype(dtype)
        self.w, self.h = int(w), int(h)
        row = self.w * self.dtype.itemsize
        self.pitch = int(pitch) if pitch is not None else _align(row, 512)  # cudaMallocPitch-like
        assert self.pitch >= row
        self.buf = torch.zeros(self.pitch * self.h, dtype=torch.uint8, device=device or "cuda")

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()

    def c(self) -> capi.RooImage:
        return capi.RooImage(self.pitch, self.ptr, self.w, self.h)

    def sub_image(self, x: int, y: int, w: int, h: int) -> "Image":
        """Image::SubImage (Image.h): a view that keeps the parent pitch."""
        v = object.__new__(Image)
        v.dtype, v.w, v.h, v.pitch = self.dtype, w, h, self.pitch
        off = y * self.pitch + x * self.dtype.itemsize
        v.buf = self.buf[off:]
        return v

    @classmethod
    def from_numpy(cls, a: np.ndarray, pitch: int | None = None, device=None) -> "Image":
        a = np.ascontiguousarray(a)
        h, w = a.shape[:2]
        dt = a.dtype if a.ndim == 2 else np.dtype((a.dtype, (a.shape[2],)))
        im = cls(w, h, dt, pitch, device)
        im.upload(a)
        return im

    def upload(self, a: np.ndarray) -> None:
        row = self.w * self.dtype.itemsize
        host = np.zeros((self.h, self.pitch), np.uint8)
        host[:, :row] = np.ascontiguousarray(a).view(np.uint8).reshape(self.h, row)
        self.buf[: self.pitch * self.h].copy_(torch.from_numpy(host.reshape(-1)))

    def numpy(self) -> np.ndarray:
        row = self.w * self.dtype.itemsize
        host = self.buf[: self.pitch * (self.h - 1) + row].cpu().numpy()
        full = np.zeros(self.pitch * self.h, np.uint8)
        full[: host.size] = host
        rows = np.ascontiguousarray(full.reshape(self.h, self.pitch)[:, :row])
        base = self.dtype.base
        out = rows.view(base)
        if self.dtype.shape:
            return out.reshape(self.h, self.w, *self.dtype.shape).copy()
        return out.reshape(self.h, self.w).copy()


class Volume:
    """roo::Volume<T, TargetDevice, Manage> (Volume.h:21-60): d outermost, x fastest."""

    def __init__(self, w: int, h: int, d: int, dtype, pitch: int | None = None, device=None):
        self.dtype = np.dtype(dtype)
        self.w, self.h, self.d = int(w), int(h), int(d)
        row = self.w * self.dtype.itemsize
        self.pitch = int(pitch) if pitch is not None else _align(row, 512)
        self.img_pitch = self.pitch * self.h  # Memory.h:70-78
        self.buf = torch.zeros(self.img_pitch * self.d, dtype=torch.uint8, device=device or "cuda")

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()

    def c(self) -> capi.RooVolume:
        return capi.RooVolume(self.pitch, self.ptr, self.w, self.h, self.img_pitch, self.d)

    @classmethod
    def from_numpy(cls, a: np.ndarray, pitch: int | None = None, device=None) -> "Volume":
        d, h, w = a.shape
        v = cls(w, h, d, a.dtype, pitch, device)
        v.upload(a)
        return v

    def upload(self, a: np.ndarray) -> None:
        row = self.w * self.dtype.itemsize
        host = np.zeros((self.d * self.h, self.pitch), np.uint8)
        host[:, :row] = np.ascontiguousarray(a).view(np.uint8).reshape(self.d * self.h, row)
        self.buf.copy_(torch.from_numpy(host.reshape(-1)))

    def fill_bytes(self, v: int) -> None:
        self.buf.fill_(v)

    def numpy(self) -> np.ndarray:
        row = self.w * self.dtype.itemsize
        rows = np.ascontiguousarray(self.buf.cpu().numpy().reshape(self.d * self.h, self.pitch)[:, :row])
        return rows.view(self.dtype).reshape(self.d, self.h, self.w).copy()


def _img_type(im: Image) -> int:
    if im.dtype == np.uint8:
        return IMG_U8
    if im.dtype == np.float32:
        return IMG_F32
    raise TypeError(f"image type {im.dtype}")


def _window_of(census: Image) -> int:
    words = census.dtype.itemsize // 8
    return {1: WIN_9x7, 2: WIN_11x11, 4: WIN_16x16}[words]


# ---------------------------------------------------------------------------------------------------
# operators (reference names)
# ---------------------------------------------------------------------------------------------------

def Census(census: Image, img: Image, stream=None) -> None:
    """roo::Census (cu_census.h:13-23): the descriptor type of `census` picks the window."""
    check(lib().roo_census(C.byref(census.c()), C.byref(img.c()), _window_of(census), _img_type(img), _stream(stream)),
          "Census")


def CensusStereo(disp: Image, left: Image, right: Image, maxDisp: int, stream=None) -> None:
    check(lib().roo_census_stereo(C.byref(disp.c()), C.byref(left.c()), C.byref(right.c()), maxDisp, _stream(stream)),
          "CensusStereo")


def CensusStereoVolume(vol: Volume, left: Image, right: Image, maxDisp: int, sd: float, popc_mode: int = POPC32_COMPAT,
                       stream=None) -> None:
    check(lib().roo_census_stereo_volume(C.byref(vol.c()), C.byref(left.c()), C.byref(right.c()),
                                         left.dtype.itemsize // 8, _VOL_TYPES[vol.dtype], maxDisp, sd, popc_mode,
                                         _stream(stream)), "CensusStereoVolume")


def SemiGlobalMatching(volH: Volume, volC: Volume, left: Image, maxDisp: int, P1: float, P2: float, dohoriz: bool,
                       dovert: bool, doreverse: bool, dodiag: bool = False, stream=None) -> None:
    check(lib().roo_sgm(C.byref(volH.c()), C.byref(volC.c()), _VOL_TYPES[volC.dtype], C.byref(left.c()),
                        _img_type(left), maxDisp, P1, P2, int(dohoriz), int(dovert), int(doreverse), int(dodiag),
                        _stream(stream)), "SemiGlobalMatching")


def CostVolMinimum(disp: Image, vol: Volume, maxDisp: int | None = None, stream=None) -> None:
    """roo::CostVolMinimum<Tdisp,Tvol>(disp, vol, maxDisp) and CostVolMinimum(Image<float>, Volume<CostVolElem>)."""
    if vol.dtype == COSTVOLELEM:
        check(lib().roo_costvol_minimum_elem(C.byref(disp.c()), C.byref(vol.c()), _stream(stream)), "CostVolMinimum")
        return
    dt = DISP_I8 if disp.dtype == np.int8 else DISP_F32
    check(lib().roo_costvol_minimum(C.byref(disp.c()), dt, C.byref(vol.c()), _VOL_TYPES[vol.dtype], maxDisp,
                                    _stream(stream)), "CostVolMinimum")


def CostVolMinimumSubpix(disp: Image, vol: Volume, maxDisp: int, sd: float, stream=None) -> None:
    check(lib().roo_costvol_minimum_subpix(C.byref(disp.c()), C.byref(vol.c()), maxDisp, sd, _stream(stream)),
          "CostVolMinimumSubpix")


def CostVolMinimumSquarePenaltySubpix(imga: Image, vol: Volume, imgd: Image, maxDisp: int, sd: float, lam: float, theta: float,
                                      stream=None) -> None:
    """roo::CostVolMinimumSquarePenaltySubpix (cu_dense_stereo.h:87-89)."""
    check(lib().roo_costvol_minimum_square_penalty_subpix(C.byref(imga.c()), C.byref(vol.c()), C.byref(imgd.c()), maxDisp, sd,
                                                          lam, theta, _stream(stream)), "CostVolMinimumSquarePenaltySubpix")


def ElementwiseMultiply(c: Image, a: Image, b: Image, scalar: float = 1.0, offset: float = 0.0, stream=None) -> None:
    """roo::ElementwiseMultiply<float,float,float,float>: c = scalar*(a*b) + offset (cu_operations.h:22-23)."""
    check(lib().roo_elementwise_multiply(C.byref(c.c()), C.byref(a.c()), C.byref(b.c()), scalar, offset, _stream(stream)),
          "ElementwiseMultiply")


def ElementwiseDivision(c: Image, a: Image, b: Image, sa: float = 0.0, sb: float = 0.0, scalar: float = 1.0, offset: float = 0.0,
                        stream=None) -> None:
    """roo::ElementwiseDivision<float,...>: c = scalar*(a+sa)/(b+sb) + offset (cu_operations.h:26-27)."""
    check(lib().roo_elementwise_division(C.byref(c.c()), C.byref(a.c()), C.byref(b.c()), sa, sb, scalar, offset, _stream(stream)),
          "ElementwiseDivision")


def ElementwiseSquare(b: Image, a: Image, scalar: float = 1.0, offset: float = 0.0, stream=None) -> None:
    """roo::ElementwiseSquare<float,float,float>: b = scalar*a*a + offset (cu_operations.h:30-31)."""
    check(lib().roo_elementwise_square(C.byref(b.c()), C.byref(a.c()), scalar, offset, _stream(stream)), "ElementwiseSquare")


def ElementwiseMultiplyAdd(d: Image, a: Image, b: Image, c: Image, sab: float = 1.0, sc: float = 1.0, offset: float = 0.0,
                           stream=None) -> None:
    """roo::ElementwiseMultiplyAdd<float,...>: d = sab*a*b + sc*c + offset (cu_operations.h:34-35)."""
    check(lib().roo_elementwise_multiply_add(C.byref(d.c()), C.byref(a.c()), C.byref(b.c()), C.byref(c.c()), sab, sc, offset,
                                             _stream(stream)), "ElementwiseMultiplyAdd")


def BoxFilter(out: Image, inp: Image, scratch, rad: int, stream=None) -> None:
    """roo::BoxFilter<float,float,float>(out, in, scratch, rad) (cu_integral_image.h:26-38).  `scratch` is accepted for the
    reference's signature and not used."""
    check(lib().roo_box_filter(C.byref(out.c()), C.byref(inp.c()), rad, _stream(stream)), "BoxFilter")


def ComputeMeanVarience(varI: Image, meanII: Image, meanI: Image, I: Image, scratch, rad: int, stream=None) -> None:
    """roo::ComputeMeanVarience<float,float,float> (cu_integral_image.h:42-54), spelled as the reference spells it."""
    BoxFilter(meanI, I, scratch, rad, stream)
    ElementwiseSquare(varI, I, stream=stream)
    BoxFilter(meanII, varI, scratch, rad, stream)
    ElementwiseMultiplyAdd(varI, meanI, meanI, meanII, -1.0, stream=stream)


def ComputeCovariance(covIP: Image, meanIP: Image, meanP: Image, P: Image, meanI: Image, I: Image, scratch, rad: int,
                      stream=None) -> None:
    """roo::ComputeCovariance (cu_integral_image.h:56-68)."""
    BoxFilter(meanP, P, scratch, rad, stream)
    ElementwiseMultiply(covIP, I, P, stream=stream)
    BoxFilter(meanIP, covIP, scratch, rad, stream)
    ElementwiseMultiplyAdd(covIP, meanI, meanP, meanIP, -1.0, stream=stream)


def GuidedFilter(q: Image, covIP: Image, varI: Image, meanP: Image, meanI: Image, I: Image, scratch, tmp1: Image, tmp2: Image,
                 tmp3: Image, rad: int, eps: float, stream=None) -> None:
    """roo::GuidedFilter (cu_integral_image.h:72-93): a = cov/(var+eps), b = meanP - a meanI, q = mean(a) I + mean(b)."""
    a, b, meana, meanb = tmp1, tmp2, tmp3, tmp1
    ElementwiseDivision(a, covIP, varI, 0.0, eps, stream=stream)
    BoxFilter(meana, a, scratch, rad, stream)
    ElementwiseMultiplyAdd(b, a, meanI, meanP, -1.0, stream=stream)
    BoxFilter(meanb, b, scratch, rad, stream)
    ElementwiseMultiplyAdd(q, meana, I, meanb, stream=stream)


def GuidedFilterVolume(vol: Volume, I: Image, rad: int, eps: float, maxDisp: int, stream=None) -> None:
    """The applications' loop over the slices of a cost volume (stereo2/main.cpp:392-405: ComputeMeanVarience once, then
    ComputeCovariance + GuidedFilter per slice, in place) as a handful of launches over all slices."""
    check(lib().roo_guided_filter_volume(C.byref(vol.c()), C.byref(I.c()), rad, eps, maxDisp, _stream(stream)), "GuidedFilterVolume")


def DenseStereo(dDisp: Image, dCamLeft: Image, dCamRight: Image, maxDisp: int, acceptThresh: float, score_rad: int, stream=None) -> None:
    """roo::DenseStereo<{unsigned char, char}, unsigned char> (cu_dense_stereo.h:24-28): TDisp follows dDisp's dtype."""
    if dDisp.dtype not in (np.uint8, np.int8):
        raise TypeError("DenseStereo: disparity image must be uint8 or int8")
    check(lib().roo_dense_stereo(C.byref(dDisp.c()), 1 if dDisp.dtype == np.int8 else 0, C.byref(dCamLeft.c()), C.byref(dCamRight.c()),
                                 int(maxDisp), acceptThresh, int(score_rad), _stream(stream)), "DenseStereo")


def release_scratch() -> None:
    """Return the scratch the box / guided filters keep between calls to the driver (current device)."""
    check(lib().roo_release_scratch(), "release_scratch")


def FilterDispGrad(dOut: Image, dIn: Image, threshold: float, stream=None) -> None:
    """roo::FilterDispGrad (cu_dense_stereo.h:101-103); dOut may be dIn, as in the applications."""
    check(lib().roo_filter_disp_grad(C.byref(dOut.c()), C.byref(dIn.c()), threshold, _stream(stream)), "FilterDispGrad")


def BilateralFilter(dOut: Image, dIn: Image, dImg: Image, gs: float, gr: float, gc: float, size: int, stream=None) -> None:
    """roo::BilateralFilter<float,float,{uchar,float}>(dOut, dIn, dImg, gs, gr, gc, size) (cu_bilateral.h:18-22)."""
    it = capi.IMG_U8 if dImg.dtype == np.uint8 else capi.IMG_F32
    check(lib().roo_bilateral_filter_joint(C.byref(dOut.c()), C.byref(dIn.c()), C.byref(dImg.c()), it, gs, gr, gc, size,
                                           _stream(stream)), "BilateralFilter")


def BilateralFilterVolume(vOut: Volume, vIn: Volume, dImg: Image, gs: float, gr: float, gc: float, size: int, maxDisp: int,
                          stream=None) -> None:
    """Every slice d < maxDisp of a cost volume through the joint bilateral filter in ONE launch (the applications loop over
    the slices on the host, stereo2/main.cpp:407-421)."""
    it = capi.IMG_U8 if dImg.dtype == np.uint8 else capi.IMG_F32
    check(lib().roo_bilateral_filter_volume(C.byref(vOut.c()), C.byref(vIn.c()), C.byref(dImg.c()), it, gs, gr, gc, size, maxDisp,
                                            _stream(stream)), "BilateralFilterVolume")


def DenseStereoSubpixelRefine(dDispOut: Image, dDisp: Image, dCamLeft: Image, dCamRight: Image, stream=None) -> None:
    check(lib().roo_dense_stereo_subpixel_refine(C.byref(dDispOut.c()), C.byref(dDisp.c()), C.byref(dCamLeft.c()),
                                                 C.byref(dCamRight.c()), _stream(stream)), "DenseStereoSubpixelRefine")


def LeftRightCheck(dispL: Image, dispR: Image, sd=-1, maxDiff=None, stream=None) -> None:
    """roo::LeftRightCheck: char overload (sd=-1, maxDiff=0) or float overload (sd=-1, maxDiff=0.5)."""
    if dispL.dtype == np.int8:
        check(lib().roo_left_right_check_i8(C.byref(dispL.c()), C.byref(dispR.c()), int(sd),
                                            0 if maxDiff is None else int(maxDiff), _stream(stream)), "LeftRightCheck")
    else:
        check(lib().roo_left_right_check_f32(C.byref(dispL.c()), C.byref(dispR.c()), float(sd),
                                             0.5 if maxDiff is None else float(maxDiff), _stream(stream)),
              "LeftRightCheck")


# ---- callers either side of the path: front end (N3) and back end (N2) of SURVEY.md section 8f

FLOAT4 = np.dtype((np.float32, (4,)))   # float4
PIX_U8, PIX_F32, PIX_U16 = 0, 1, 2
_PIX_TYPES = {np.dtype(np.uint8): PIX_U8, np.dtype(np.float32): PIX_F32, np.dtype(np.uint16): PIX_U16}


def ElementwiseScaleBias(b: Image, a: Image, s: float, offset: float = 0.0, stream=None) -> None:
    """roo::ElementwiseScaleBias<float,{uchar,ushort,float},float> (cu_operations.h:14-15): b = s*a + offset."""
    check(lib().roo_elementwise_scale_bias(C.byref(b.c()), C.byref(a.c()), _PIX_TYPES[a.dtype], s, offset,
                                           _stream(stream)), "ElementwiseScaleBias")


def BoxHalf(out: Image, in_: Image, stream=None) -> None:
    """roo::BoxHalf<uchar,uint,uchar> / <float,float,float> (reduce.h:7-8)."""
    if out.dtype != in_.dtype or in_.dtype not in (np.dtype(np.uint8), np.dtype(np.float32)):
        raise TypeError("BoxHalf: unsigned char or float images of the same type (reduce.h:7-8 instantiations)")
    check(lib().roo_box_half(C.byref(out.c()), C.byref(in_.c()), _PIX_TYPES[in_.dtype], _stream(stream)), "BoxHalf")


def BoxReduce(pyramid, stream=None) -> None:
    """roo::BoxReduce(Pyramid) (reduce.h:35-46): level l = BoxHalf(level l-1); `pyramid` is a list of Images."""
    for lvl in range(1, len(pyramid)):
        BoxHalf(pyramid[lvl], pyramid[lvl - 1], stream)


FLOAT2 = np.dtype((np.float32, (2,)))   # float2


def CreateMatlabLookupTable(lookup: Image, fu: float, fv: float, u0: float, v0: float, k1: float, k2: float, H_on=None,
                            stream=None) -> None:
    """roo::CreateMatlabLookupTable (cu_lookup_warp.cu:32-38, and :77-83 with a 3x3 homography H_on); lookup is FLOAT2."""
    if H_on is None:
        check(lib().roo_create_matlab_lookup_table(C.byref(lookup.c()), fu, fv, u0, v0, k1, k2, _stream(stream)),
              "CreateMatlabLookupTable")
    else:
        Hc = (C.c_float * 9)(*[float(x) for x in np.asarray(H_on).ravel()])
        check(lib().roo_create_matlab_lookup_table_homography(C.byref(lookup.c()), fu, fv, u0, v0, k1, k2, Hc,
                                                              _stream(stream)), "CreateMatlabLookupTable")


def Warp(out: Image, in_: Image, lookup: Image, stream=None) -> None:
    """roo::Warp (cu_lookup_warp.cu:96-106): rectify `in_` through a FLOAT2 lookup table (bilinear)."""
    check(lib().roo_warp(C.byref(out.c()), C.byref(in_.c()), C.byref(lookup.c()), _stream(stream)), "Warp")


def Disp2Depth(dIn: Image, dOut: Image, fu: float, fBaseline: float, fMinDisp: float = 0.0, stream=None) -> None:
    """roo::Disp2Depth (cu_depth_tools.h:11)."""
    check(lib().roo_disp2depth(C.byref(dIn.c()), C.byref(dOut.c()), fu, fBaseline, fMinDisp, _stream(stream)),
          "Disp2Depth")


def DisparityImageToVbo(dVbo: Image, dDisp: Image, baseline: float, fu: float, fv: float, u0: float, v0: float,
                        stream=None) -> None:
    """roo::DisparityImageToVbo (cu_dense_stereo.cu:633-646); dVbo is an Image of FLOAT4."""
    check(lib().roo_disparity_image_to_vbo(C.byref(dVbo.c()), C.byref(dDisp.c()), baseline, fu, fv, u0, v0,
                                           _stream(stream)), "DisparityImageToVbo")


def CostVolumeFromStereoTruncatedAbsAndGrad(dvol: Volume, dimgl: Image, dimgr: Image, sd: float, alpha: float, r1: float,
                                            r2: float, stream=None) -> None:
    """roo::CostVolumeFromStereoTruncatedAbsAndGrad (cu_dense_stereo.h:66); alpha and r1 are ignored as in the reference."""
    check(lib().roo_costvol_from_stereo_truncated_abs_and_grad(C.byref(dvol.c()), C.byref(dimgl.c()), C.byref(dimgr.c()),
                                                               sd, alpha, r1, r2, _stream(stream)),
          "CostVolumeFromStereoTruncatedAbsAndGrad")


def _median(size):
    def f(dOut: Image, dIn: Image, maxbad: int = 100, stream=None) -> None:
        check(lib().roo_median_filter_reject_negative(C.byref(dOut.c()), C.byref(dIn.c()), size, maxbad, _stream(stream)),
              f"MedianFilterRejectNegative{size}x{size}")
    f.__doc__ = f"roo::MedianFilterRejectNegative{size}x{size} (cu_median.h:19-32), out of place."
    return f


MedianFilterRejectNegative5x5 = _median(5)
MedianFilterRejectNegative7x7 = _median(7)
MedianFilterRejectNegative9x9 = _median(9)


def set_ieee_division(on: bool) -> None:
    lib().roo_set_ieee_division(int(on))


def set_tuning(knob: int, value: int) -> None:
    """Development knobs for A/B measurements (roo_set_tuning); results never depend on them."""
    check(lib().roo_set_tuning(int(knob), int(value)), "roo_set_tuning")


# ---------------------------------------------------------------------------------------------------
# fused engine
# ---------------------------------------------------------------------------------------------------

def _fuse_flag(fuse_vertical) -> int:
    """None: the engine decides per group (roo_pipeline_params_t.fuse_vertical = 0); True: always; False: never."""
    return 0 if fuse_vertical is None else (1 if fuse_vertical else -1)


class StereoEngine:
    """The whole per-frame path (stereo2/main.cpp:375-454) on engine-owned scratch, batched."""

    def __init__(self, w: int, h: int, max_disp: int, window: int = WIN_9x7, popc_mode: int = POPC32_COMPAT,
                 P1: float = 0.01, P2: float = 0.02, img_scale: float = 1.0 / 255.0, dohoriz=True, dovert=True,
                 doreverse=True, dodiag=False, subpix=False, lrcheck=False, lr_maxdiff: float = 1.0,
                 max_batch: int = 1, keep_volume: bool = False, fuse_vertical: bool | None = None, median_size: int = 0,
                 median_maxbad: int = 100, median_iters: int = 1, fp_mode: int = capi.FP_DEFAULT,
                 filtgrad_threshold: float = 0.0):
        self.params = capi.PipelineParams(w, h, max_disp, window, popc_mode, P1, P2, np.float32(img_scale),
                                          int(dohoriz), int(dovert), int(doreverse), int(dodiag), int(subpix),
                                          int(lrcheck), lr_maxdiff, max_batch, int(keep_volume),
                                          _fuse_flag(fuse_vertical), median_size, median_maxbad, median_iters, fp_mode,
                                          filtgrad_threshold)
        self.w, self.h, self.max_disp = w, h, max_disp
        self._h = C.c_void_p()
        check(lib().roo_engine_create(C.byref(self._h), C.byref(self.params)), "roo_engine_create")

    def close(self) -> None:
        if self._h:
            lib().roo_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def scratch_bytes(self) -> int:
        return int(lib().roo_engine_scratch_bytes(self._h))

    def set_front_end(self, level: int = 0, lookup_left: Image | None = None, lookup_right: Image | None = None) -> None:
        """Raw frames of (w << level, h << level) in: [Warp through the FLOAT2 tables] -> BoxHalf x level -> the path."""
        self._luts = (lookup_left, lookup_right)   # keep the tables alive
        check(lib().roo_engine_set_front_end(self._h, level,
                                             C.byref(lookup_left.c()) if lookup_left is not None else None,
                                             C.byref(lookup_right.c()) if lookup_right is not None else None),
              "roo_engine_set_front_end")
        self._fe_level = level

    def _in_shape(self):
        """(h, w) of the frames a run takes: the working size, or the raw size when a front end is set."""
        lvl = getattr(self, "_fe_level", 0)
        return (self.h << lvl, self.w << lvl)

    def _check_io(self, left, right, disp, cuda: bool) -> int:
        """The C library takes raw pointers: shapes, dtypes, devices and contiguity are checked here."""
        n = int(left.shape[0]) if left.dim() == 3 else -1
        ih, iw = self._in_shape()
        for name, t, dt, shape in (("left", left, torch.uint8, (n, ih, iw)), ("right", right, torch.uint8, (n, ih, iw)),
                                   ("disp", disp, torch.float32, (n, self.h, self.w))):
            if t is None:
                continue
            if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.is_cuda != cuda:
                raise ValueError(f"{name}: expected a contiguous {'CUDA' if cuda else 'host'} {dt} tensor of shape {shape}, "
                                 f"got {tuple(t.shape)} {t.dtype} {'cuda' if t.is_cuda else 'cpu'}")
        if cuda and (right.device != left.device or (disp is not None and disp.device != left.device)):
            raise ValueError("left, right and disp must live on the same device")
        return n

    def run_device(self, left: torch.Tensor, right: torch.Tensor, disp: torch.Tensor | None = None, stream=None):
        """left/right: (n, h, w) uint8 CUDA tensors (contiguous); returns (n, h, w) float32."""
        n = self._check_io(left, right, disp, cuda=True)
        if disp is None:
            disp = torch.empty((n, self.h, self.w), dtype=torch.float32, device=left.device)
        check(lib().roo_engine_run_device(self._h, left.data_ptr(), right.data_ptr(), disp.data_ptr(), n,
                                          _stream(stream)), "roo_engine_run_device")
        return disp

    def run_host(self, left: torch.Tensor, right: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
        """Host (ideally pinned) tensors in, host tensor out; H2D + compute + D2H, synchronous."""
        self._check_io(left, right, disp, cuda=False)
        check(lib().roo_engine_run_host(self._h, left.data_ptr(), right.data_ptr(), disp.data_ptr(), left.shape[0]),
              "roo_engine_run_host")
        return disp

    def submit_host(self, left: torch.Tensor, right: torch.Tensor, disp: torch.Tensor) -> int:
        """Asynchronous run_host for one group (<= max_batch pairs): returns a ticket for wait().  The tensors must
        stay alive (and should be pinned) until wait(ticket) returns."""
        self._check_io(left, right, disp, cuda=False)
        t = C.c_longlong(-1)
        check(lib().roo_engine_submit_host(self._h, left.data_ptr(), right.data_ptr(), disp.data_ptr(), left.shape[0],
                                           C.byref(t)), "roo_engine_submit_host")
        return int(t.value)

    def wait(self, ticket: int) -> None:
        check(lib().roo_engine_wait(self._h, ticket), "roo_engine_wait")

    def set_profiling(self, on: bool) -> None:
        check(lib().roo_engine_set_profiling(self._h, int(on)), "roo_engine_set_profiling")

    def get_profile(self) -> dict:
        """{kind: (total_ms, launches)} since profiling was switched on; synchronise the stream first."""
        n = len(capi.PROF_KINDS)
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        check(lib().roo_engine_get_profile(self._h, ms, cnt), "roo_engine_get_profile")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(capi.PROF_KINDS)}

    def export_volume(self, slot: int = 0, depth: int | None = None) -> Volume:
        v = Volume(self.w, self.h, depth or self.max_disp, np.float32)
        check(lib().roo_engine_export_volume(self._h, slot, C.byref(v.c()), _stream(None)), "roo_engine_export_volume")
        return v

    def export_census(self, slot: int, side: int) -> Image:
        words = WORDS[self.params.window]
        im = Image(self.w, self.h, np.dtype((np.uint64, (words,))) if words > 1 else ULONG)
        check(lib().roo_engine_export_census(self._h, slot, side, C.byref(im.c()), _stream(None)),
              "roo_engine_export_census")
        return im


class MultiGpuStereoEngine:
    """The batch sharded across the GPUs of the box by pair index: one engine + one host thread per device
    inside the C++ library, no collective (roo_multi_engine_*).  Host tensors in, host tensor out."""

    def __init__(self, w: int, h: int, max_disp: int, devices=None, **kw):
        proto = StereoEngine.__new__(StereoEngine)
        defaults = dict(window=WIN_9x7, popc_mode=POPC32_COMPAT, P1=0.01, P2=0.02, img_scale=1.0 / 255.0, dohoriz=True,
                        dovert=True, doreverse=True, dodiag=False, subpix=False, lrcheck=False, lr_maxdiff=1.0,
                        max_batch=1, keep_volume=False, fuse_vertical=None, median_size=0, median_maxbad=100,
                        median_iters=1, fp_mode=capi.FP_DEFAULT, filtgrad_threshold=0.0)
        defaults.update(kw)
        d = defaults
        self.params = capi.PipelineParams(w, h, max_disp, d["window"], d["popc_mode"], d["P1"], d["P2"],
                                          np.float32(d["img_scale"]), int(d["dohoriz"]), int(d["dovert"]),
                                          int(d["doreverse"]), int(d["dodiag"]), int(d["subpix"]), int(d["lrcheck"]),
                                          d["lr_maxdiff"], d["max_batch"], int(d["keep_volume"]),
                                          _fuse_flag(d["fuse_vertical"]), d["median_size"], d["median_maxbad"],
                                          d["median_iters"], d["fp_mode"], d["filtgrad_threshold"])
        del proto
        self.w, self.h = w, h
        self._h = C.c_void_p()
        if devices is None:
            arr, n = None, 0
        else:
            arr, n = (C.c_int * len(devices))(*devices), len(devices)
        check(lib().roo_multi_engine_create(C.byref(self._h), C.byref(self.params), arr, n), "roo_multi_engine_create")

    @property
    def device_count(self) -> int:
        return int(lib().roo_multi_engine_device_count(self._h))

    def run_host(self, left: torch.Tensor, right: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
        n = int(left.shape[0]) if left.dim() == 3 else -1
        for name, t, dt, shape in (("left", left, torch.uint8, (n, self.h, self.w)), ("right", right, torch.uint8, (n, self.h, self.w)),
                                   ("disp", disp, torch.float32, (n, self.h, self.w))):
            if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.is_cuda:
                raise ValueError(f"{name}: expected a contiguous host {dt} tensor of shape {shape}, got {tuple(t.shape)} {t.dtype}")
        check(lib().roo_multi_engine_run_host(self._h, left.data_ptr(), right.data_ptr(), disp.data_ptr(),
                                              left.shape[0]), "roo_multi_engine_run_host")
        return disp

    def close(self) -> None:
        if self._h:
            lib().roo_multi_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SplitStereoEngine:
    """ONE pair split into row strips across GPUs (roo_split_engine_*): the paths that travel in y hand their state from
    strip to strip through peer memory over NVLink; everything else is strip-local.  Host tensors in, host tensor out."""

    def __init__(self, w: int, h: int, max_disp: int, devices=None, **kw):
        defaults = dict(window=WIN_9x7, popc_mode=POPC32_COMPAT, P1=0.01, P2=0.02, img_scale=1.0 / 255.0, dohoriz=True,
                        dovert=True, doreverse=True, dodiag=False, subpix=False, lrcheck=False, lr_maxdiff=1.0,
                        keep_volume=False, fp_mode=capi.FP_DEFAULT)
        defaults.update(kw)
        d = defaults
        self.params = capi.PipelineParams(w, h, max_disp, d["window"], d["popc_mode"], d["P1"], d["P2"],
                                          np.float32(d["img_scale"]), int(d["dohoriz"]), int(d["dovert"]),
                                          int(d["doreverse"]), int(d["dodiag"]), int(d["subpix"]), int(d["lrcheck"]),
                                          d["lr_maxdiff"], 1, int(d["keep_volume"]), -1, 0, 100, 1, d["fp_mode"], 0.0)
        self.w, self.h = w, h
        self._h = C.c_void_p()
        if devices is None:
            arr, n = None, 0
        else:
            arr, n = (C.c_int * len(devices))(*devices), len(devices)
        check(lib().roo_split_engine_create(C.byref(self._h), C.byref(self.params), arr, n), "roo_split_engine_create")

    @property
    def strip_count(self) -> int:
        return int(lib().roo_split_engine_strip_count(self._h))

    def run_host(self, left: torch.Tensor, right: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
        for name, t, dt in (("left", left, torch.uint8), ("right", right, torch.uint8), ("disp", disp, torch.float32)):
            if tuple(t.shape) != (self.h, self.w) or t.dtype != dt or not t.is_contiguous() or t.is_cuda:
                raise ValueError(f"{name}: expected a contiguous host {dt} tensor of shape {(self.h, self.w)}")
        check(lib().roo_split_engine_run_host(self._h, left.data_ptr(), right.data_ptr(), disp.data_ptr()),
              "roo_split_engine_run_host")
        return disp

    def last_stats(self):
        """(device milliseconds of the last frame, bytes handed between strips through peer memory)"""
        ms, nb = C.c_float(0), C.c_ulonglong(0)
        check(lib().roo_split_engine_last_stats(self._h, C.byref(ms), C.byref(nb)), "roo_split_engine_last_stats")
        return float(ms.value), int(nb.value)

    def close(self) -> None:
        if self._h:
            lib().roo_split_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
