// roo.hpp -- header-only C++ shim that re-creates the reference's operator surface for the census /
// semi-global-matching path on top of the C ABI in roo_b200.h.
//
// A translation unit of an application written against Kangaroo
//   #include <kangaroo/cu_census.h> / cu_semi_global_matching.h / cu_dense_stereo.h
// can include this header instead and link libroo_b200.so: the free functions below have the reference's
// names, argument order and meaning, and the same explicit instantiation set
// (cu_census.cu:309-314, cu_semi_global_matching.cu:88-89, cu_dense_stereo.cu:54-60).
//
// roo::Image / roo::Volume here are minimal non-owning views with the reference's member layout
// (Image.h:617-620, Volume.h:363-369) -- when this header is used NEXT to the real kangaroo headers, define
// ROO_B200_USE_KANGAROO_TYPES before including it and the real roo::Image / roo::Volume are used instead
// (they are layout-compatible with roo_image_t / roo_volume_t).
//
// Like the reference launchers the operators return void, are asynchronous and report nothing
// (cu_census.cu:180-220 never check the launch); define ROO_B200_THROW to turn a non-zero status into a
// std::runtime_error.  Every call goes to the stream set with roo::SetStream() (default: the legacy default
// stream, which is what the reference uses).
#pragma once

#include <cstddef>
#include <cstdio>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>

#include <vector_types.h>   // ulong2, ulong4 (CUDA toolkit)

#include "../roo_b200.h"

#ifdef ROO_B200_USE_KANGAROO_TYPES
// the reference's own data model: roo::Image / roo::Volume (and their Target / Management policies) and CostVolElem
#include <kangaroo/Image.h>
#include <kangaroo/Volume.h>
#include <kangaroo/CostVolElem.h>
#endif

namespace roo {

#ifndef ROO_B200_USE_KANGAROO_TYPES
struct TargetDevice {};
struct DontManage {};
// include/kangaroo/CostVolElem.h:10-19
struct alignas(8) CostVolElem { int n; float sum; };

template <typename T, typename Target = TargetDevice, typename Management = DontManage>
struct Image {
    Image() : pitch(0), ptr(nullptr), w(0), h(0) {}
    Image(T* p, size_t w_, size_t h_) : pitch(sizeof(T) * w_), ptr(p), w(w_), h(h_) {}
    Image(T* p, size_t w_, size_t h_, size_t pitch_) : pitch(pitch_), ptr(p), w(w_), h(h_) {}
    Image SubImage(size_t x, size_t y, size_t width, size_t height) const {
        return Image((T*)((unsigned char*)ptr + y * pitch) + x, width, height, pitch);
    }
    size_t pitch;
    T* ptr;
    size_t w;
    size_t h;
};

template <typename T, typename Target = TargetDevice, typename Management = DontManage>
struct Volume {
    Volume() : pitch(0), ptr(nullptr), w(0), h(0), img_pitch(0), d(0) {}
    // (the reference's 4-argument constructor leaves d unset, Volume.h:55-59; this one sets it)
    Volume(T* p, size_t w_, size_t h_, size_t d_) : pitch(sizeof(T) * w_), ptr(p), w(w_), h(h_), img_pitch(sizeof(T) * w_ * h_), d(d_) {}
    Volume(T* p, size_t w_, size_t h_, size_t d_, size_t pitch_) : pitch(pitch_), ptr(p), w(w_), h(h_), img_pitch(pitch_ * h_), d(d_) {}
    Volume(T* p, size_t w_, size_t h_, size_t d_, size_t pitch_, size_t img_pitch_) : pitch(pitch_), ptr(p), w(w_), h(h_), img_pitch(img_pitch_), d(d_) {}
    Image<T, Target, DontManage> ImageXY(size_t z) const {
        return Image<T, Target, DontManage>((T*)((unsigned char*)ptr + z * img_pitch), w, h, pitch);
    }
    size_t pitch;
    T* ptr;
    size_t w;
    size_t h;
    size_t img_pitch;
    size_t d;
};
#endif  // ROO_B200_USE_KANGAROO_TYPES

static_assert(sizeof(Image<float>) == sizeof(roo_image_t), "roo::Image must be layout-compatible with roo_image_t");
static_assert(sizeof(Volume<float>) == sizeof(roo_volume_t), "roo::Volume must be layout-compatible with roo_volume_t");

namespace b200 {
inline void*& stream_slot() { static thread_local void* s = nullptr; return s; }
inline void done(int status, const char* what) {
#ifdef ROO_B200_THROW
    if (status != ROO_OK) throw std::runtime_error(std::string(what) + ": " + roo_status_string(status));
#else
    (void)status; (void)what;
#endif
}
template <typename T> roo_image_t c(const Image<T>& i) { return roo_image_t{i.pitch, (void*)i.ptr, i.w, i.h}; }
template <typename T> roo_volume_t c(const Volume<T>& v) { return roo_volume_t{v.pitch, (void*)v.ptr, v.w, v.h, v.img_pitch, v.d}; }
template <typename T> struct words;
template <> struct words<unsigned long> { static constexpr int n = 1, win = ROO_WIN_9x7; };
template <> struct words<ulong2> { static constexpr int n = 2, win = ROO_WIN_11x11; };
template <> struct words<ulong4> { static constexpr int n = 4, win = ROO_WIN_16x16; };
template <typename T> struct voltype;
template <> struct voltype<float> { static constexpr int v = ROO_VOL_F32; };
template <> struct voltype<int> { static constexpr int v = ROO_VOL_I32; };
template <> struct voltype<unsigned int> { static constexpr int v = ROO_VOL_U32; };
template <> struct voltype<unsigned short> { static constexpr int v = ROO_VOL_U16; };
template <> struct voltype<unsigned char> { static constexpr int v = ROO_VOL_U8; };
template <> struct voltype<CostVolElem> { static constexpr int v = ROO_VOL_ELEM; };
template <typename T> struct imgtype;
template <> struct imgtype<unsigned char> { static constexpr int v = ROO_IMG_U8; };
template <> struct imgtype<float> { static constexpr int v = ROO_IMG_F32; };
}  // namespace b200

// extension: the stream every operator below launches on (thread-local; default 0 = legacy default stream)
inline void SetStream(void* cuda_stream) { b200::stream_slot() = cuda_stream; }

// ---- cu_census.h:12-23 -----------------------------------------------------------------------------
inline void Census(Image<unsigned long> census, Image<unsigned char> img) { auto a = b200::c(census), b = b200::c(img); b200::done(roo_census(&a, &b, ROO_WIN_9x7, ROO_IMG_U8, b200::stream_slot()), "Census"); }
inline void Census(Image<ulong2> census, Image<unsigned char> img) { auto a = b200::c(census), b = b200::c(img); b200::done(roo_census(&a, &b, ROO_WIN_11x11, ROO_IMG_U8, b200::stream_slot()), "Census"); }
inline void Census(Image<ulong4> census, Image<unsigned char> img) { auto a = b200::c(census), b = b200::c(img); b200::done(roo_census(&a, &b, ROO_WIN_16x16, ROO_IMG_U8, b200::stream_slot()), "Census"); }
inline void Census(Image<unsigned long> census, Image<float> img) { auto a = b200::c(census), b = b200::c(img); b200::done(roo_census(&a, &b, ROO_WIN_9x7, ROO_IMG_F32, b200::stream_slot()), "Census"); }
inline void Census(Image<ulong2> census, Image<float> img) { auto a = b200::c(census), b = b200::c(img); b200::done(roo_census(&a, &b, ROO_WIN_11x11, ROO_IMG_F32, b200::stream_slot()), "Census"); }
inline void Census(Image<ulong4> census, Image<float> img) { auto a = b200::c(census), b = b200::c(img); b200::done(roo_census(&a, &b, ROO_WIN_16x16, ROO_IMG_F32, b200::stream_slot()), "Census"); }

// ---- cu_census.h:33 ----------------------------------------------------------------------------------
inline void CensusStereo(Image<char> disp, Image<unsigned long> left, Image<unsigned long> right, int maxDisp) {
    auto d = b200::c(disp), l = b200::c(left), r = b200::c(right);
    b200::done(roo_census_stereo(&d, &l, &r, maxDisp, b200::stream_slot()), "CensusStereo");
}

// ---- cu_census.h:36-38; instantiated for Tvol in {unsigned short, float} x T in {unsigned long, ulong2, ulong4}
template <typename Tvol, typename T>
inline void CensusStereoVolume(Volume<Tvol> vol, Image<T> left, Image<T> right, int maxDisp, float sd) {
    static_assert(std::is_same<Tvol, float>::value || std::is_same<Tvol, unsigned short>::value, "Tvol");
    auto v = b200::c(vol); auto l = b200::c(left), r = b200::c(right);
    b200::done(roo_census_stereo_volume(&v, &l, &r, b200::words<T>::n, b200::voltype<Tvol>::v, maxDisp, sd,
                                        ROO_POPC32_COMPAT, b200::stream_slot()), "CensusStereoVolume");
}

// ---- cu_semi_global_matching.h:10-12; instantiated <float,CostVolElem,unsigned char> and <float,float,float>
template <typename TH, typename TC, typename Timg>
inline void SemiGlobalMatching(Volume<TH> volH, Volume<TC> volC, Image<Timg> left, int maxDisp, float P1, float P2,
                               bool dohoriz, bool dovert, bool doreverse) {
    static_assert(std::is_same<TH, float>::value, "TH must be float");
    auto h = b200::c(volH); auto cvol = b200::c(volC); auto l = b200::c(left);
    b200::done(roo_sgm(&h, &cvol, b200::voltype<TC>::v, &l, b200::imgtype<Timg>::v, maxDisp, P1, P2, dohoriz, dovert,
                       doreverse, /*dodiag=*/0, b200::stream_slot()), "SemiGlobalMatching");
}
// extension: the same call with the four diagonal paths
template <typename TH, typename TC, typename Timg>
inline void SemiGlobalMatching8(Volume<TH> volH, Volume<TC> volC, Image<Timg> left, int maxDisp, float P1, float P2,
                                bool dohoriz, bool dovert, bool doreverse) {
    auto h = b200::c(volH); auto cvol = b200::c(volC); auto l = b200::c(left);
    b200::done(roo_sgm(&h, &cvol, b200::voltype<TC>::v, &l, b200::imgtype<Timg>::v, maxDisp, P1, P2, dohoriz, dovert,
                       doreverse, /*dodiag=*/1, b200::stream_slot()), "SemiGlobalMatching8");
}

// ---- cu_dense_stereo.h:13-15; instantiated (char,{float,int,uint,ushort,uchar}) and (float,{float,ushort})
template <typename Tdisp, typename Tvol>
inline void CostVolMinimum(Image<Tdisp> disp, Volume<Tvol> vol, unsigned maxDisp) {
    static_assert(std::is_same<Tdisp, char>::value || std::is_same<Tdisp, float>::value, "Tdisp");
    auto d = b200::c(disp); auto v = b200::c(vol);
    b200::done(roo_costvol_minimum(&d, std::is_same<Tdisp, char>::value ? ROO_DISP_I8 : ROO_DISP_F32, &v,
                                   b200::voltype<Tvol>::v, maxDisp, b200::stream_slot()), "CostVolMinimum");
}
// ---- cu_dense_stereo.h:81-82
inline void CostVolMinimum(Image<float> disp, Volume<CostVolElem> vol) {
    auto d = b200::c(disp); auto v = b200::c(vol);
    b200::done(roo_costvol_minimum_elem(&d, &v, b200::stream_slot()), "CostVolMinimum");
}
// ---- cu_dense_stereo.h:84-85
inline void CostVolMinimumSubpix(Image<float> disp, Volume<float> vol, unsigned maxDisp, float sd) {
    auto d = b200::c(disp); auto v = b200::c(vol);
    b200::done(roo_costvol_minimum_subpix(&d, &v, maxDisp, sd, b200::stream_slot()), "CostVolMinimumSubpix");
}
// ---- cu_dense_stereo.h:87-89
inline void CostVolMinimumSquarePenaltySubpix(Image<float> imga, Volume<float> vol, Image<float> imgd, unsigned maxDisp, float sd,
                                              float lambda, float theta) {
    auto a = b200::c(imga), d = b200::c(imgd);
    auto v = b200::c(vol);
    b200::done(roo_costvol_minimum_square_penalty_subpix(&a, &v, &d, maxDisp, sd, lambda, theta, b200::stream_slot()),
               "CostVolMinimumSquarePenaltySubpix");
}
// ---- cu_bilateral.h:18-22 (the overload the applications run on cost-volume slices) and its whole-volume form
template <typename To, typename Ti, typename Ti2>
inline void BilateralFilter(Image<To> dOut, const Image<Ti> dIn, const Image<Ti2> dImg, float gs, float gr, float gc, unsigned size) {
    static_assert(std::is_same<To, float>::value && std::is_same<Ti, float>::value, "float in, float out");
    auto o = b200::c(dOut), i2 = b200::c(dIn), g = b200::c(dImg);
    b200::done(roo_bilateral_filter_joint(&o, &i2, &g, b200::imgtype<Ti2>::v, gs, gr, gc, size, b200::stream_slot()), "BilateralFilter");
}
template <typename Ti2>
inline void BilateralFilterVolume(Volume<float> vOut, Volume<float> vIn, const Image<Ti2> dImg, float gs, float gr, float gc,
                                  unsigned size, int maxDisp) {
    auto o = b200::c(vOut), i2 = b200::c(vIn);
    auto g = b200::c(dImg);
    b200::done(roo_bilateral_filter_volume(&o, &i2, &g, b200::imgtype<Ti2>::v, gs, gr, gc, size, maxDisp, b200::stream_slot()),
               "BilateralFilterVolume");
}
// ---- cu_dense_stereo.h:24-28 (instantiated for <unsigned char, unsigned char> and <char, unsigned char>, cu_dense_stereo.cu:405-406)
template <typename TDisp, typename TImg>
inline void DenseStereo(Image<TDisp> dDisp, const Image<TImg> dCamLeft, const Image<TImg> dCamRight, TDisp maxDisp, float acceptThresh,
                        int score_rad) {
    static_assert(std::is_same<TImg, unsigned char>::value && (std::is_same<TDisp, unsigned char>::value || std::is_same<TDisp, char>::value),
                  "DenseStereo<{unsigned char, char}, unsigned char>");
    auto d = b200::c(dDisp), l = b200::c(dCamLeft), r = b200::c(dCamRight);
    b200::done(roo_dense_stereo(&d, std::is_same<TDisp, char>::value ? 1 : 0, &l, &r, (int)maxDisp, acceptThresh, score_rad,
                                b200::stream_slot()), "DenseStereo");
}
// ---- cu_operations.h:22-35, the float instantiations the guided filter is composed of
template <typename Tout, typename Tin1, typename Tin2, typename Tup>
inline void ElementwiseMultiply(Image<Tout> c, Image<Tin1> a, Image<Tin2> b, Tup scalar = 1, Tup offset = 0) {
    static_assert(std::is_same<Tout, float>::value && std::is_same<Tin1, float>::value && std::is_same<Tin2, float>::value, "float images");
    auto ci = b200::c(c), ai = b200::c(a), bi = b200::c(b);
    b200::done(roo_elementwise_multiply(&ci, &ai, &bi, (float)scalar, (float)offset, b200::stream_slot()), "ElementwiseMultiply");
}
template <typename Tout, typename Tin1, typename Tin2, typename Tup>
inline void ElementwiseDivision(Image<Tout> c, const Image<Tin1> a, const Image<Tin2> b, Tup sa = 0, Tup sb = 0, Tup scalar = 1, Tup offset = 0) {
    static_assert(std::is_same<Tout, float>::value && std::is_same<Tin1, float>::value && std::is_same<Tin2, float>::value, "float images");
    auto ci = b200::c(c), ai = b200::c(a), bi = b200::c(b);
    b200::done(roo_elementwise_division(&ci, &ai, &bi, (float)sa, (float)sb, (float)scalar, (float)offset, b200::stream_slot()),
               "ElementwiseDivision");
}
template <typename Tout, typename Tin, typename Tup>
inline void ElementwiseSquare(Image<Tout> b, const Image<Tin> a, Tup scalar = 1, Tup offset = 0) {
    static_assert(std::is_same<Tout, float>::value && std::is_same<Tin, float>::value, "float images");
    auto bi = b200::c(b), ai = b200::c(a);
    b200::done(roo_elementwise_square(&bi, &ai, (float)scalar, (float)offset, b200::stream_slot()), "ElementwiseSquare");
}
template <typename Tout, typename Tin1, typename Tin2, typename Tin3, typename Tup>
inline void ElementwiseMultiplyAdd(Image<Tout> d, const Image<Tin1> a, const Image<Tin2> b, const Image<Tin3> c, Tup sab = 1, Tup sc = 1,
                                   Tup offset = 0) {
    static_assert(std::is_same<Tout, float>::value && std::is_same<Tin1, float>::value && std::is_same<Tin2, float>::value &&
                  std::is_same<Tin3, float>::value, "float images");
    auto di = b200::c(d), ai = b200::c(a), bi = b200::c(b), ci = b200::c(c);
    b200::done(roo_elementwise_multiply_add(&di, &ai, &bi, &ci, (float)sab, (float)sc, (float)offset, b200::stream_slot()),
               "ElementwiseMultiplyAdd");
}
// ---- cu_integral_image.h:26-38.  Scratch is part of the reference's signature; this library brings its own.
template <typename Tout, typename Tin, typename TSum>
inline void BoxFilter(Image<Tout> out, Image<Tin> in, Image<unsigned char> /*scratch*/, int rad) {
    static_assert(std::is_same<Tout, float>::value && std::is_same<Tin, float>::value && std::is_same<TSum, float>::value, "float images");
    auto o = b200::c(out), i2 = b200::c(in);
    b200::done(roo_box_filter(&o, &i2, rad, b200::stream_slot()), "BoxFilter");
}
// ---- cu_integral_image.h:42-54
template <typename Tout, typename Tin, typename TSum>
inline void ComputeMeanVarience(Image<Tout> varI, Image<Tout> meanII, Image<Tout> meanI, const Image<Tin> I, Image<unsigned char> Scratch, int rad) {
    BoxFilter<float, float, float>(meanI, I, Scratch, rad);
    ElementwiseSquare<float, float, float>(varI, I);                       // I.*I, parked in the output image
    BoxFilter<float, float, float>(meanII, varI, Scratch, rad);
    ElementwiseMultiplyAdd<float, float, float, float, float>(varI, meanI, meanI, meanII, -1);
}
// ---- cu_integral_image.h:56-68
inline void ComputeCovariance(Image<float> covIP, Image<float> meanIP, Image<float> meanP, const Image<float> P, const Image<float> meanI,
                              const Image<float> I, Image<unsigned char> Scratch, int rad) {
    BoxFilter<float, float, float>(meanP, P, Scratch, rad);
    ElementwiseMultiply<float, float, float, float>(covIP, I, P);          // I.*p, parked in the output image
    BoxFilter<float, float, float>(meanIP, covIP, Scratch, rad);
    ElementwiseMultiplyAdd<float, float, float, float, float>(covIP, meanI, meanP, meanIP, -1);
}
// ---- cu_integral_image.h:72-93 (tmp1 holds a, then mean_b; tmp2 holds b; tmp3 holds mean_a)
inline void GuidedFilter(Image<float> q, const Image<float> covIP, const Image<float> varI, const Image<float> meanP, const Image<float> meanI,
                         const Image<float> I, Image<unsigned char> Scratch, Image<float> tmp1, Image<float> tmp2, Image<float> tmp3,
                         int rad, float eps) {
    ElementwiseDivision<float, float, float, float>(tmp1, covIP, varI, 0, eps);
    BoxFilter<float, float, float>(tmp3, tmp1, Scratch, rad);
    ElementwiseMultiplyAdd<float, float, float, float, float>(tmp2, tmp1, meanI, meanP, -1);
    BoxFilter<float, float, float>(tmp1, tmp2, Scratch, rad);
    ElementwiseMultiplyAdd<float, float, float, float, float>(q, tmp3, I, tmp1);
}
// The applications' loop (stereo2/main.cpp:392-405: ComputeMeanVarience once, then ComputeCovariance + GuidedFilter per
// slice, in place) over the first maxDisp slices in a handful of launches.
inline void GuidedFilterVolume(Volume<float> vol, const Image<float> I, int rad, float eps, int maxDisp) {
    auto v = b200::c(vol);
    auto g = b200::c(I);
    b200::done(roo_guided_filter_volume(&v, &g, rad, eps, maxDisp, b200::stream_slot()), "GuidedFilterVolume");
}
// ---- cu_dense_stereo.h:101-103 (dOut may be dIn, as in stereo2/main.cpp:457)
inline void FilterDispGrad(Image<float> dOut, Image<float> dIn, float threshold) {
    auto o = b200::c(dOut), i2 = b200::c(dIn);
    b200::done(roo_filter_disp_grad(&o, &i2, threshold, b200::stream_slot()), "FilterDispGrad");
}
// ---- cu_dense_stereo.h:45-47
inline void DenseStereoSubpixelRefine(Image<float> dDispOut, const Image<unsigned char> dDisp,
                                      const Image<unsigned char> dCamLeft, const Image<unsigned char> dCamRight) {
    auto o = b200::c(dDispOut), d = b200::c(dDisp), l = b200::c(dCamLeft), r = b200::c(dCamRight);
    b200::done(roo_dense_stereo_subpixel_refine(&o, &d, &l, &r, b200::stream_slot()), "DenseStereoSubpixelRefine");
}
// ---- cu_dense_stereo.h:37-41
inline void LeftRightCheck(Image<char> dispL, Image<char> dispR, int sd = -1, int maxDiff = 0) {
    auto l = b200::c(dispL), r = b200::c(dispR);
    b200::done(roo_left_right_check_i8(&l, &r, sd, maxDiff, b200::stream_slot()), "LeftRightCheck");
}
inline void LeftRightCheck(Image<float> dispL, Image<float> dispR, float sd = -1, float maxDiff = 0.5) {
    auto l = b200::c(dispL), r = b200::c(dispR);
    b200::done(roo_left_right_check_f32(&l, &r, sd, maxDiff, b200::stream_slot()), "LeftRightCheck");
}

// ---- callers either side of the path: cu_operations.h:14-15, reduce.h:7-8, cu_depth_tools.h:11,
//      cu_dense_stereo.h (DisparityImageToVbo) -- same names, template parameters and defaults as the reference
namespace b200 {
template <typename T> struct pix_type;
template <> struct pix_type<unsigned char> { static constexpr int value = ROO_PIX_U8; };
template <> struct pix_type<unsigned short> { static constexpr int value = ROO_PIX_U16; };
template <> struct pix_type<float> { static constexpr int value = ROO_PIX_F32; };
}  // namespace b200
template <typename Tout, typename Tin, typename Tup>
inline void ElementwiseScaleBias(Image<Tout> b, const Image<Tin> a, float s, Tup offset = 0) {
    static_assert(std::is_same<Tout, float>::value && std::is_same<Tup, float>::value,
                  "ElementwiseScaleBias: <float, {unsigned char, unsigned short, float}, float> (cu_operations.cu:260-262)");
    auto cb = b200::c(b), ca = b200::c(a);
    b200::done(roo_elementwise_scale_bias(&cb, &ca, b200::pix_type<Tin>::value, s, offset, b200::stream_slot()),
               "ElementwiseScaleBias");
}
template <typename To, typename UpType, typename Ti>
inline void BoxHalf(Image<To> out, const Image<Ti> in) {
    static_assert(std::is_same<To, Ti>::value && (std::is_same<Ti, unsigned char>::value || std::is_same<Ti, float>::value),
                  "BoxHalf: <unsigned char, unsigned int, unsigned char> or <float, float, float> (cu_resample.cu:80-81)");
    auto co = b200::c(out), ci = b200::c(in);
    b200::done(roo_box_half(&co, &ci, b200::pix_type<Ti>::value, b200::stream_slot()), "BoxHalf");
}
inline void CreateMatlabLookupTable(Image<float2> lookup, float fu, float fv, float u0, float v0, float k1, float k2) {
    auto cl = b200::c(lookup);
    b200::done(roo_create_matlab_lookup_table(&cl, fu, fv, u0, v0, k1, k2, b200::stream_slot()), "CreateMatlabLookupTable");
}
// the overload with a homography (cu_lookup_warp.cu:77-83); H_on = 9 floats, row-major, e.g. Mat<float,9>::m
inline void CreateMatlabLookupTable(Image<float2> lookup, float fu, float fv, float u0, float v0, float k1, float k2,
                                    const float (&H_on)[9]) {
    auto cl = b200::c(lookup);
    b200::done(roo_create_matlab_lookup_table_homography(&cl, fu, fv, u0, v0, k1, k2, H_on, b200::stream_slot()),
               "CreateMatlabLookupTable");
}
inline void Warp(Image<unsigned char> out, const Image<unsigned char> in, const Image<float2> lookup) {
    auto co = b200::c(out), ci = b200::c(in), cl = b200::c(lookup);
    b200::done(roo_warp(&co, &ci, &cl, b200::stream_slot()), "Warp");
}
inline void Disp2Depth(Image<float> dIn, const Image<float> dOut, float fu, float fBaseline, float fMinDisp = 0.0) {
    auto ci = b200::c(dIn), co = b200::c(dOut);
    b200::done(roo_disp2depth(&ci, &co, fu, fBaseline, fMinDisp, b200::stream_slot()), "Disp2Depth");
}
inline void DisparityImageToVbo(Image<float4> dVbo, const Image<float> dDisp, float baseline, float fu, float fv,
                                float u0, float v0) {
    auto cv = b200::c(dVbo), cd = b200::c(dDisp);
    b200::done(roo_disparity_image_to_vbo(&cv, &cd, baseline, fu, fv, u0, v0, b200::stream_slot()), "DisparityImageToVbo");
}

// ---- cu_dense_stereo.h:66
inline void CostVolumeFromStereoTruncatedAbsAndGrad(Volume<float> dvol, Image<float> dimgl, Image<float> dimgr, float sd,
                                                    float alpha, float r1, float r2) {
    auto v = b200::c(dvol);
    auto l = b200::c(dimgl), r = b200::c(dimgr);
    b200::done(roo_costvol_from_stereo_truncated_abs_and_grad(&v, &l, &r, sd, alpha, r1, r2, b200::stream_slot()),
               "CostVolumeFromStereoTruncatedAbsAndGrad");
}
// ---- cu_median.h:19-32 (dOut may alias dIn as in the applications: the library then filters through a temporary)
inline void MedianFilterRejectNegative5x5(Image<float> dOut, Image<float> dIn, int maxbad = 100) {
    auto o = b200::c(dOut), i = b200::c(dIn);
    b200::done(roo_median_filter_reject_negative(&o, &i, 5, maxbad, b200::stream_slot()), "MedianFilterRejectNegative5x5");
}
inline void MedianFilterRejectNegative7x7(Image<float> dOut, Image<float> dIn, int maxbad) {
    auto o = b200::c(dOut), i = b200::c(dIn);
    b200::done(roo_median_filter_reject_negative(&o, &i, 7, maxbad, b200::stream_slot()), "MedianFilterRejectNegative7x7");
}
inline void MedianFilterRejectNegative9x9(Image<float> dOut, Image<float> dIn, int maxbad) {
    auto o = b200::c(dOut), i = b200::c(dIn);
    b200::done(roo_median_filter_reject_negative(&o, &i, 9, maxbad, b200::stream_slot()), "MedianFilterRejectNegative9x9");
}

// ---- on-disk outputs (extra/SavePPM.h:20-39; applications/stereo/main.cpp:400-410): host pointers, tightly packed rows
template <typename T>
inline bool SavePXM(const std::string& filename, const T* host, size_t w, size_t h, size_t pitch_bytes,
                    const std::string& ppm_type = "P5", int num_colors = 255) {
    FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "%s\n%zu %zu\n%d\n", ppm_type.c_str(), w, h, num_colors);
    for (size_t r = 0; r < h; ++r) std::fwrite(reinterpret_cast<const char*>(host) + r * pitch_bytes, sizeof(T), w, f);
    return std::fclose(f) == 0;
}
inline bool SavePDM(const std::string& filename, const float* host, size_t cols, size_t rows) {
    FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "P7\n%zu %zu\n4294967295\n", cols, rows);
    std::fwrite(host, sizeof(float), cols * rows, f);
    return std::fclose(f) == 0;
}

// ---- extension: the fused per-frame engine (census -> cost -> SGM -> WTA/subpixel -> LR check) --------
class StereoEngine {
public:
    explicit StereoEngine(const roo_pipeline_params_t& p) : e_(nullptr) {
        const int rc = roo_engine_create(&e_, &p);
        if (rc != ROO_OK) throw std::runtime_error(std::string("roo_engine_create: ") + roo_status_string(rc));
    }
    ~StereoEngine() { if (e_) roo_engine_destroy(e_); }
    StereoEngine(const StereoEngine&) = delete;
    StereoEngine& operator=(const StereoEngine&) = delete;
    // n tightly packed (h x w) uint8 pairs in device memory -> n (h x w) float disparity images
    void RunDevice(const uint8_t* left, const uint8_t* right, float* disp, int n, void* stream = nullptr) {
        b200::done(roo_engine_run_device(e_, left, right, disp, n, stream), "roo_engine_run_device");
    }
    // the same with host buffers (pinned recommended): upload, compute and download are pipelined
    void RunHost(const uint8_t* left, const uint8_t* right, float* disp, int n) {
        b200::done(roo_engine_run_host(e_, left, right, disp, n), "roo_engine_run_host");
    }
    // raw frames in: rectify through the lookup tables (or none) and BoxReduce `level` times before the path
    void SetFrontEnd(int level, const Image<float2>* lookupLeft = nullptr, const Image<float2>* lookupRight = nullptr) {
        roo_image_t l{}, r{};
        if (lookupLeft && lookupRight) { l = b200::c(*lookupLeft); r = b200::c(*lookupRight); }
        b200::done(roo_engine_set_front_end(e_, level, lookupLeft ? &l : nullptr, lookupRight ? &r : nullptr),
                   "roo_engine_set_front_end");
    }
    // streaming form: enqueue one group (n <= max_batch) and return a ticket; Wait(ticket) blocks until its
    // disparities are in `disp`.  Two groups may be in flight (copies overlap the other group's kernels).
    long long SubmitHost(const uint8_t* left, const uint8_t* right, float* disp, int n) {
        long long t = -1;
        b200::done(roo_engine_submit_host(e_, left, right, disp, n, &t), "roo_engine_submit_host");
        return t;
    }
    void Wait(long long ticket) { b200::done(roo_engine_wait(e_, ticket), "roo_engine_wait"); }
    roo_engine_t* handle() { return e_; }
private:
    roo_engine_t* e_;
};

}  // namespace roo
