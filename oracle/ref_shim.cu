// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin extern "C" shim over the UNMODIFIED reference operators so that tests and
// the golden-vector generator can call them through ctypes.  This file contains no
// reference code: it only includes the reference's public headers
// (include/kangaroo/cu_census.h, cu_semi_global_matching.h, cu_dense_stereo.h) and
// forwards plain pointers to the roo:: free functions.  It is compiled together
// with /root/reference/src/{cu_census,cu_semi_global_matching,cu_dense_stereo,cu_operations,
// cu_resample,cu_depth_tools,cu_median,cu_lookup_warp,cu_bilateral,cu_integral_image}.cu (from where they lie) into oracle/_ref/libkangaroo_ref.so
// by oracle/Makefile.
//
// The reference kernels launch one thread per pixel of a row/column in ONE block
// (cu_census.cu:304-306, cu_semi_global_matching.cu:69-83), so every entry point
// here refuses w > 1024 or h > 1024 instead of letting the launch fail silently.
#include <cuda_runtime.h>
#include <kangaroo/cu_census.h>
#include <kangaroo/cu_semi_global_matching.h>
#include <kangaroo/cu_dense_stereo.h>
#include <kangaroo/cu_operations.h>
#include <kangaroo/Pyramid.h>
#include <kangaroo/reduce.h>
#include <kangaroo/cu_depth_tools.h>
#include <kangaroo/cu_median.h>
#include <kangaroo/cu_lookup_warp.h>
#include <kangaroo/cu_bilateral.h>
#include <kangaroo/cu_integral_image.h>
#include <vector>

namespace {
template <typename T>
roo::Image<T> img(void* p, size_t pitch, size_t w, size_t h) {
    return roo::Image<T>((T*)p, w, h, pitch);
}
template <typename T>
roo::Volume<T> vol(void* p, size_t pitch, size_t img_pitch, size_t w, size_t h, size_t d) {
    return roo::Volume<T>((T*)p, w, h, d, pitch, img_pitch);
}
int finish() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}
bool too_big(size_t w, size_t h) { return w > 1024 || h > 1024; }

// Padded device image for the integral-image operators: BoxFilter (cu_integral_image.h:26-38) borrows its OUTPUT image as
// storage for the transposed row sums (out.AlignedImage<TSum>(in.h, in.w)), which only fits when
// align16(4 h) * w <= h * pitch -- hence the slack in the pitch.
struct PadImg {
    float* p = nullptr;
    size_t pitch = 0, w = 0, h = 0;
    PadImg(size_t w_, size_t h_) : w(w_), h(h_) {
        pitch = (4 * w + 16 * ((w + h - 1) / h) + 31) / 16 * 16;
        cudaMalloc((void**)&p, pitch * h);
        cudaMemset(p, 0, pitch * h);
    }
    ~PadImg() { cudaFree(p); }
    PadImg(const PadImg&) = delete;
    PadImg& operator=(const PadImg&) = delete;
    roo::Image<float> im() const { return roo::Image<float>(p, w, h, pitch); }
    void put(const void* dense) { cudaMemcpy2D(p, pitch, dense, 4 * w, 4 * w, h, cudaMemcpyDeviceToDevice); }
    void get(void* dense) const { cudaMemcpy2D(dense, 4 * w, p, pitch, 4 * w, h, cudaMemcpyDeviceToDevice); }
};
struct ScratchBytes {
    unsigned char* p = nullptr;
    size_t n = 0;
    ScratchBytes(size_t w, size_t h) {
        n = ((4 * w + 15) / 16 * 16) * h + ((4 * h + 15) / 16 * 16) * w + 256;
        cudaMalloc((void**)&p, n);
    }
    ~ScratchBytes() { cudaFree(p); }
    roo::Image<unsigned char> im() const { return roo::Image<unsigned char>(p, n, 1, n); }
};
}  // namespace

extern "C" {

// window: 0 = 9x7 -> unsigned long, 1 = "11x11" -> ulong2, 2 = "16x16" -> ulong4
// in_type: 0 = unsigned char, 1 = float
int kref_census(void* out, size_t out_pitch, void* in, size_t in_pitch, size_t w, size_t h,
                int window, int in_type) {
    if (in_type == 0) {
        roo::Image<unsigned char> i = img<unsigned char>(in, in_pitch, w, h);
        if (window == 0) roo::Census(img<unsigned long>(out, out_pitch, w, h), i);
        else if (window == 1) roo::Census(img<ulong2>(out, out_pitch, w, h), i);
        else if (window == 2) roo::Census(img<ulong4>(out, out_pitch, w, h), i);
        else return -1;
    } else if (in_type == 1) {
        roo::Image<float> i = img<float>(in, in_pitch, w, h);
        if (window == 0) roo::Census(img<unsigned long>(out, out_pitch, w, h), i);
        else if (window == 1) roo::Census(img<ulong2>(out, out_pitch, w, h), i);
        else if (window == 2) roo::Census(img<ulong4>(out, out_pitch, w, h), i);
        else return -1;
    } else return -1;
    return finish();
}

int kref_census_stereo(void* disp, size_t disp_pitch, void* l, void* r, size_t c_pitch,
                       size_t w, size_t h, int maxDisp) {
    if (too_big(w, h)) return -2;
    roo::CensusStereo(img<char>(disp, disp_pitch, w, h), img<unsigned long>(l, c_pitch, w, h),
                      img<unsigned long>(r, c_pitch, w, h), maxDisp);
    return finish();
}

// words: 1/2/4 (unsigned long / ulong2 / ulong4); vol_type: 0 = unsigned short, 1 = float
int kref_census_stereo_volume(void* v, size_t v_pitch, size_t v_img_pitch, size_t d, void* l, void* r,
                              size_t c_pitch, size_t w, size_t h, int words, int vol_type,
                              int maxDisp, float sd) {
    if (too_big(w, h)) return -2;
#define KREF_CSV(TV, TC)                                                                          \
    roo::CensusStereoVolume<TV, TC>(vol<TV>(v, v_pitch, v_img_pitch, w, h, d),                    \
                                    img<TC>(l, c_pitch, w, h), img<TC>(r, c_pitch, w, h), maxDisp, sd)
    if (vol_type == 1) {
        if (words == 1) KREF_CSV(float, unsigned long);
        else if (words == 2) KREF_CSV(float, ulong2);
        else if (words == 4) KREF_CSV(float, ulong4);
        else return -1;
    } else if (vol_type == 0) {
        if (words == 1) KREF_CSV(unsigned short, unsigned long);
        else if (words == 2) KREF_CSV(unsigned short, ulong2);
        else if (words == 4) KREF_CSV(unsigned short, ulong4);
        else return -1;
    } else return -1;
#undef KREF_CSV
    return finish();
}

// volc_type: 0 = float (left image float), 1 = CostVolElem (left image unsigned char)
int kref_sgm(void* vh, void* vc, size_t h_pitch, size_t h_img_pitch, size_t c_pitch, size_t c_img_pitch,
             void* left, size_t left_pitch, size_t w, size_t h, size_t d, int volc_type, int maxDisp,
             float P1, float P2, int dohoriz, int dovert, int doreverse) {
    if (too_big(w, h)) return -2;
    if (volc_type == 0)
        roo::SemiGlobalMatching<float, float, float>(vol<float>(vh, h_pitch, h_img_pitch, w, h, d),
                                                     vol<float>(vc, c_pitch, c_img_pitch, w, h, d),
                                                     img<float>(left, left_pitch, w, h), maxDisp, P1, P2,
                                                     dohoriz != 0, dovert != 0, doreverse != 0);
    else if (volc_type == 1)
        roo::SemiGlobalMatching<float, roo::CostVolElem, unsigned char>(
            vol<float>(vh, h_pitch, h_img_pitch, w, h, d),
            vol<roo::CostVolElem>(vc, c_pitch, c_img_pitch, w, h, d),
            img<unsigned char>(left, left_pitch, w, h), maxDisp, P1, P2, dohoriz != 0, dovert != 0,
            doreverse != 0);
    else return -1;
    return finish();
}

// disp_type: 0 = char, 1 = float; vol_type: 0 float, 1 int, 2 unsigned, 3 unsigned short, 4 unsigned char
// NOTE the reference kernel is unguarded (cu_dense_stereo.cu:25-43): only call with w%32==0 && h%32==0.
int kref_costvol_minimum(void* disp, size_t disp_pitch, void* v, size_t v_pitch, size_t v_img_pitch,
                         size_t w, size_t h, size_t d, int disp_type, int vol_type, unsigned maxDisp) {
    if ((w % 32) || (h % 32)) return -3;
#define KREF_CVM(TD, TV)                                                                          \
    roo::CostVolMinimum<TD, TV>(img<TD>(disp, disp_pitch, w, h), vol<TV>(v, v_pitch, v_img_pitch, w, h, d), maxDisp)
    if (disp_type == 0) {
        if (vol_type == 0) KREF_CVM(char, float);
        else if (vol_type == 1) KREF_CVM(char, int);
        else if (vol_type == 2) KREF_CVM(char, unsigned int);
        else if (vol_type == 3) KREF_CVM(char, unsigned short);
        else if (vol_type == 4) KREF_CVM(char, unsigned char);
        else return -1;
    } else if (disp_type == 1) {
        if (vol_type == 0) KREF_CVM(float, float);
        else if (vol_type == 3) KREF_CVM(float, unsigned short);
        else return -1;
    } else return -1;
#undef KREF_CVM
    return finish();
}

int kref_costvol_minimum_elem(void* disp, size_t disp_pitch, void* v, size_t v_pitch, size_t v_img_pitch,
                              size_t w, size_t h, size_t d) {
    if ((w % 32) || (h % 32)) return -3;
    roo::CostVolMinimum(img<float>(disp, disp_pitch, w, h),
                        vol<roo::CostVolElem>(v, v_pitch, v_img_pitch, w, h, d));
    return finish();
}

int kref_costvol_minimum_subpix(void* disp, size_t disp_pitch, void* v, size_t v_pitch, size_t v_img_pitch,
                                size_t w, size_t h, size_t d, unsigned maxDisp, float sd) {
    roo::CostVolMinimumSubpix(img<float>(disp, disp_pitch, w, h),
                              vol<float>(v, v_pitch, v_img_pitch, w, h, d), maxDisp, sd);
    return finish();
}

int kref_dense_stereo_subpixel_refine(void* out, size_t out_pitch, void* disp, void* l, void* r,
                                      size_t u8_pitch, size_t w, size_t h) {
    roo::DenseStereoSubpixelRefine(img<float>(out, out_pitch, w, h), img<unsigned char>(disp, u8_pitch, w, h),
                                   img<unsigned char>(l, u8_pitch, w, h), img<unsigned char>(r, u8_pitch, w, h));
    return finish();
}

int kref_left_right_check_f32(void* dl, void* dr, size_t pitch, size_t w, size_t h, float sd, float maxDiff) {
    roo::LeftRightCheck(img<float>(dl, pitch, w, h), img<float>(dr, pitch, w, h), sd, maxDiff);
    return finish();
}

int kref_left_right_check_i8(void* dl, void* dr, size_t pitch, size_t w, size_t h, int sd, int maxDiff) {
    roo::LeftRightCheck(img<char>(dl, pitch, w, h), img<char>(dr, pitch, w, h), sd, maxDiff);
    return finish();
}

// ---- callers either side of the path (front end / back end)
// in_type: 0 = unsigned char, 1 = float, 2 = unsigned short
int kref_elementwise_scale_bias(void* b, size_t b_pitch, void* a, size_t a_pitch, size_t w, size_t h, int in_type,
                                float s, float offset) {
    if (in_type == 0) roo::ElementwiseScaleBias<float, unsigned char, float>(img<float>(b, b_pitch, w, h), img<unsigned char>(a, a_pitch, w, h), s, offset);
    else if (in_type == 1) roo::ElementwiseScaleBias<float, float, float>(img<float>(b, b_pitch, w, h), img<float>(a, a_pitch, w, h), s, offset);
    else if (in_type == 2) roo::ElementwiseScaleBias<float, unsigned short, float>(img<float>(b, b_pitch, w, h), img<unsigned short>(a, a_pitch, w, h), s, offset);
    else return -1;
    return finish();
}

// pix_type: 0 = unsigned char, 1 = float; (w, h) = size of the OUTPUT image, the input is (2w, 2h)
int kref_box_half(void* out, size_t out_pitch, void* in, size_t in_pitch, size_t w, size_t h, int pix_type) {
    if (pix_type == 0) roo::BoxHalf<unsigned char, unsigned int, unsigned char>(img<unsigned char>(out, out_pitch, w, h), img<unsigned char>(in, in_pitch, 2 * w, 2 * h));
    else if (pix_type == 1) roo::BoxHalf<float, float, float>(img<float>(out, out_pitch, w, h), img<float>(in, in_pitch, 2 * w, 2 * h));
    else return -1;
    return finish();
}

int kref_disp2depth(void* in, void* out, size_t pitch, size_t w, size_t h, float fu, float baseline, float minDisp) {
    roo::Disp2Depth(img<float>(in, pitch, w, h), img<float>(out, pitch, w, h), fu, baseline, minDisp);
    return finish();
}

int kref_disparity_image_to_vbo(void* vbo, size_t vbo_pitch, void* disp, size_t disp_pitch, size_t w, size_t h,
                                float baseline, float fu, float fv, float u0, float v0) {
    roo::DisparityImageToVbo(img<float4>(vbo, vbo_pitch, w, h), img<float>(disp, disp_pitch, w, h), baseline, fu, fv, u0, v0);
    return finish();
}

// size: 5, 7 or 9 (MedianFilterRejectNegative{5x5,7x7,9x9}); out and in must be different images (the kernels read
// neighbours that other blocks write when called in place)
int kref_median_reject_negative(void* out, void* in, size_t pitch, size_t w, size_t h, int size, int maxbad) {
    if (out == in) return -3;
    if (size == 5) roo::MedianFilterRejectNegative5x5(img<float>(out, pitch, w, h), img<float>(in, pitch, w, h), maxbad);
    else if (size == 7) roo::MedianFilterRejectNegative7x7(img<float>(out, pitch, w, h), img<float>(in, pitch, w, h), maxbad);
    else if (size == 9) roo::MedianFilterRejectNegative9x9(img<float>(out, pitch, w, h), img<float>(in, pitch, w, h), maxbad);
    else return -1;
    return finish();
}

// rectification warp: out (w x h) = bilinear sample of in (in_w x in_h) at lookup(x, y); lookup holds float2
int kref_warp(void* out, size_t out_pitch, void* in, size_t in_pitch, size_t in_w, size_t in_h, void* lookup,
              size_t lookup_pitch, size_t w, size_t h) {
    roo::Warp(img<unsigned char>(out, out_pitch, w, h), img<unsigned char>(in, in_pitch, in_w, in_h),
              img<float2>(lookup, lookup_pitch, w, h));
    return finish();
}

// N4: alternative matching cost (cu_dense_stereo.cu:820-848); the kernel has no bounds test: w, h, d multiples of 8 only
int kref_costvol_abs_and_grad(void* v, size_t v_pitch, size_t v_img_pitch, size_t d, void* l, void* r, size_t i_pitch,
                              size_t w, size_t h, float sd, float alpha, float r1, float r2) {
    if (w % 8 || h % 8 || d % 8) return -4;
    roo::CostVolumeFromStereoTruncatedAbsAndGrad(vol<float>(v, v_pitch, v_img_pitch, w, h, d), img<float>(l, i_pitch, w, h),
                                                 img<float>(r, i_pitch, w, h), sd, alpha, r1, r2);
    return finish();
}

// cu_dense_stereo.cu:122-174
int kref_costvol_minimum_square_penalty_subpix(void* imga, void* v, size_t v_pitch, size_t v_img_pitch, size_t d, void* imgd,
                                               size_t i_pitch, size_t w, size_t h, unsigned maxDisp, float sd, float lambda, float theta) {
    roo::CostVolMinimumSquarePenaltySubpix(img<float>(imga, i_pitch, w, h), vol<float>(v, v_pitch, v_img_pitch, w, h, d),
                                           img<float>(imgd, i_pitch, w, h), maxDisp, sd, lambda, theta);
    return finish();
}

// cu_dense_stereo.cu:793-812: reads the central differences of OUT (its previous contents) and the values of IN; called
// out of place here (out != in), on images that are the interior of a larger allocation (the kernel reads x-1, x+1, y-1, y+1
// unguarded).  w and h must be multiples of 16 (gcd grid, launch_utils.h:61-65).
int kref_filter_disp_grad(void* out, void* in, size_t pitch, size_t w, size_t h, float threshold) {
    if (out == in) return -3;
    roo::FilterDispGrad(img<float>(out, pitch, w, h), img<float>(in, pitch, w, h), threshold);
    return finish();
}

// cu_bilateral.cu:110-155; img_type: 0 = unsigned char, 1 = float guide image
int kref_bilateral_joint(void* out, void* in, size_t pitch, void* gimg, size_t g_pitch, int img_type, size_t w, size_t h, float gs,
                         float gr, float gc, unsigned size) {
    if (img_type == 0) roo::BilateralFilter<float, float, unsigned char>(img<float>(out, pitch, w, h), img<float>(in, pitch, w, h), img<unsigned char>(gimg, g_pitch, w, h), gs, gr, gc, size);
    else roo::BilateralFilter<float, float, float>(img<float>(out, pitch, w, h), img<float>(in, pitch, w, h), img<float>(gimg, g_pitch, w, h), gs, gr, gc, size);
    return finish();
}

// cu_integral_image.h:26-38 (PrefixSumRows -> Transpose -> PrefixSumRows -> BoxFilterIntegralImage); dense w x h float
// device buffers in and out.  The scan kernel runs one block of nextpow2(w)/2 threads per row: w, h <= 2048.
int kref_box_filter(void* out, void* in, size_t w, size_t h, int rad) {
    if (w > 2048 || h > 2048 || w < 2 || h < 2) return -2;
    PadImg o(w, h), i(w, h);
    ScratchBytes sc(w, h);
    i.put(in);
    roo::BoxFilter<float, float, float>(o.im(), i.im(), sc.im(), rad);
    const int e = finish();
    o.get(out);
    return e ? e : finish();
}

// The applications' guided filtering of a cost volume (applications/stereo2/main.cpp:392-405): ComputeMeanVarience once
// per guide image, then ComputeCovariance + GuidedFilter per disparity slice, in place.  v: dense D x h x w floats.
int kref_guided_filter_volume(void* v, void* guide, size_t w, size_t h, size_t D, int rad, float eps) {
    if (w > 2048 || h > 2048 || w < 2 || h < 2) return -2;
    PadImg I(w, h), varI(w, h), meanI(w, h), P(w, h), t0(w, h), t1(w, h), t2(w, h), t3(w, h), t4(w, h);
    ScratchBytes sc(w, h);
    I.put(guide);
    roo::ComputeMeanVarience<float, float, float>(varI.im(), t0.im(), meanI.im(), I.im(), sc.im(), rad);
    for (size_t d = 0; d < D; ++d) {
        float* slice = (float*)v + d * w * h;
        P.put(slice);
        roo::ComputeCovariance(t0.im(), t2.im(), t1.im(), P.im(), meanI.im(), I.im(), sc.im(), rad);
        roo::GuidedFilter(P.im(), t0.im(), varI.im(), t1.im(), meanI.im(), I.im(), sc.im(), t2.im(), t3.im(), t4.im(), rad, eps);
        const int e = finish();
        if (e) return e;
        P.get(slice);
    }
    return finish();
}

// cu_operations.cu:85-165,170-190: the float elementwise operators the guided filter is composed of.
// op: 0 = Multiply (c = s0*(a*b) + s1), 1 = Division (c = s2*(a+s0)/(b+s1) + s3), 2 = Square (c = s0*a*a + s1),
//     3 = MultiplyAdd (d = s0*a*b + s1*c + s2)
int kref_elementwise(int op, void* out, void* a, void* b, void* c, size_t w, size_t h, float s0, float s1, float s2, float s3) {
    const size_t p = 4 * w;
    switch (op) {
    case 0: roo::ElementwiseMultiply<float, float, float, float>(img<float>(out, p, w, h), img<float>(a, p, w, h), img<float>(b, p, w, h), s0, s1); break;
    case 1: roo::ElementwiseDivision<float, float, float, float>(img<float>(out, p, w, h), img<float>(a, p, w, h), img<float>(b, p, w, h), s0, s1, s2, s3); break;
    case 2: roo::ElementwiseSquare<float, float, float>(img<float>(out, p, w, h), img<float>(a, p, w, h), s0, s1); break;
    case 3: roo::ElementwiseMultiplyAdd<float, float, float, float, float>(img<float>(out, p, w, h), img<float>(a, p, w, h), img<float>(b, p, w, h), img<float>(c, p, w, h), s0, s1, s2); break;
    default: return -1;
    }
    return finish();
}

// cu_dense_stereo.cu:209-253,376-406: the direct block matcher.  disp_type 0 = unsigned char, 1 = char disparities; dense
// w x h images.  One block of w threads per row: w <= 1024.  maxDisp == 255 (127 for char) never terminates in the
// reference (the candidate counter wraps before it exceeds the bound) and is refused here.
int kref_dense_stereo(void* disp, void* l, void* r, size_t w, size_t h, int disp_type, int maxDisp, float acceptThresh, int score_rad) {
    if (w > 1024 || score_rad < 0 || score_rad > 7) return -2;
    if (disp_type == 0) {
        if (maxDisp < 0 || maxDisp >= 255) return -2;
        roo::DenseStereo<unsigned char, unsigned char>(img<unsigned char>(disp, w, w, h), img<unsigned char>(l, w, w, h), img<unsigned char>(r, w, w, h),
                                                       (unsigned char)maxDisp, acceptThresh, score_rad);
    } else {
        if (maxDisp <= -128 || maxDisp >= 127) return -2;
        roo::DenseStereo<char, unsigned char>(img<char>(disp, w, w, h), img<unsigned char>(l, w, w, h), img<unsigned char>(r, w, w, h), (char)maxDisp,
                                              acceptThresh, score_rad);
    }
    return finish();
}

int kref_create_matlab_lookup_table(void* lookup, size_t pitch, size_t w, size_t h, float fu, float fv, float u0, float v0,
                                    float k1, float k2) {
    roo::CreateMatlabLookupTable(img<float2>(lookup, pitch, w, h), fu, fv, u0, v0, k1, k2);
    return finish();
}

int kref_create_matlab_lookup_table_h(void* lookup, size_t pitch, size_t w, size_t h, float fu, float fv, float u0, float v0,
                                      float k1, float k2, const float* H_on) {
    roo::Mat<float, 9> H;
    for (int i = 0; i < 9; ++i) H[i] = H_on[i];
    roo::CreateMatlabLookupTable(img<float2>(lookup, pitch, w, h), fu, fv, u0, v0, k1, k2, H);
    return finish();
}

}  // extern "C"
