// Cost-volume filters of the applications' alternative aggregation.
//
// roo::BilateralFilter<float,float,Timg>(dOut, dIn, dImg, gs, gr, gc, size) -- the joint bilateral filter with spatial,
// range and guide-image weights (include/kangaroo/cu_bilateral.h:18-22; src/cu_bilateral.cu:110-155).  The applications
// filter a cost volume with it one disparity slice per launch, after a device-to-device copy of the slice
// (applications/stereo2/main.cpp:407-421: a host loop of `maxdisp` copies and launches).  Here the whole volume is ONE
// launch: a thread owns one pixel and DCH consecutive slices, so the spatial weight and the guide-image weight of a tap
// -- which do not depend on the slice -- are computed once per tap instead of once per tap and slice.
//
// Arithmetic = the reference's -use_fast_math SASS, operation for operation (results bit-identical to its kernel):
//   rS = MUFU.RCP((gs+gs)*gs), sw = MUFU.EX2(((float)(r*r+c*c) * -rS) * log2e), rw = EX2(((d * -d) * rR) * log2e) with
//   d = p - q, cw likewise on the guide image, w = cw * (sw * rw), sumw = w + sumw, sum = FFMA(q, w, sum),
//   out = sumw == 0 ? p : sum * MUFU.RCP(sumw) (div.approx.ftz); every operation flushes denormals (.ftz).
#include "common.cuh"
#include "ftz.cuh"
#include "kernels.cuh"

namespace roo_b200 {

__device__ __forceinline__ float ex2_ftz(float a) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
constexpr float BIL_LOG2E = 1.4426950216293334961f;   // the constant in the reference's SASS

// exp(-(v*v) / (2 g g)) the reference's way: EX2(((v * -v) * rcp) * log2e)
__device__ __forceinline__ float gauss_w(float v, float rcp2gg) { return ex2_ftz(fmul_ftz(fmul_ftz(fmul_ftz(v, -v), rcp2gg), BIL_LOG2E)); }

constexpr int BIL_TX = 128, BIL_DCH = 8;

template <typename Timg>
__global__ void __launch_bounds__(BIL_TX)
bilateral_joint_volume_kernel(Vol<float> out, Vol<float> in, Img<Timg> img, float gs, float gr, float gc, int size, int nd) {
    const int x = blockIdx.x * BIL_TX + threadIdx.x, y = blockIdx.y, dbase = blockIdx.z * BIL_DCH;
    if (x >= out.w) return;
    const int w = in.w, h = in.h;
    const float rS = rcp_approx_ftz(fmul_ftz(fadd_ftz(gs, gs), gs));
    const float rR = rcp_approx_ftz(fmul_ftz(fadd_ftz(gr, gr), gr));
    const float rC = rcp_approx_ftz(fmul_ftz(fadd_ftz(gc, gc), gc));
    const float pc = (float)img(x, y);
    float p[BIL_DCH], sum[BIL_DCH], sumw[BIL_DCH];
#pragma unroll
    for (int k = 0; k < BIL_DCH; ++k) { p[k] = dbase + k < nd ? in(x, y, dbase + k) : 0.0f; sum[k] = 0.0f; sumw[k] = 0.0f; }
    for (int r = -size; r <= size; ++r) {
        const int yy = clampi(y + r, 0, h - 1);
        for (int c = -size; c <= size; ++c) {
            const int xx = clampi(x + c, 0, w - 1);
            const float sw = ex2_ftz(fmul_ftz(fmul_ftz((float)(r * r + c * c), -rS), BIL_LOG2E));
            const float cw = gauss_w(fadd_ftz(pc, -(float)img(xx, yy)), rC);
#pragma unroll
            for (int k = 0; k < BIL_DCH; ++k) {
                if (dbase + k < nd) {
                    const float q = in(xx, yy, dbase + k);
                    const float rw = gauss_w(fadd_ftz(p[k], -q), rR);
                    const float wgt = fmul_ftz(cw, fmul_ftz(sw, rw));
                    sumw[k] = fadd_ftz(wgt, sumw[k]);
                    sum[k] = ffma_ftz(q, wgt, sum[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < BIL_DCH; ++k)
        if (dbase + k < nd) out(x, y, dbase + k) = sumw[k] == 0.0f ? p[k] : ref_div<false>(sum[k], sumw[k]);
}

static int launch_bilateral(const roo_volume_t& out, const roo_volume_t& in, const roo_image_t& img, int img_type, float gs,
                            float gr, float gc, int size, int nd, cudaStream_t st) {
    dim3 grid(cdiv((int)out.w, BIL_TX), (unsigned)out.h, (unsigned)cdiv(nd, BIL_DCH));
    if (img_type == ROO_IMG_U8)
        bilateral_joint_volume_kernel<unsigned char><<<grid, BIL_TX, 0, st>>>(Vol<float>(out), Vol<float>(in), Img<unsigned char>(img), gs, gr, gc, size, nd);
    else
        bilateral_joint_volume_kernel<float><<<grid, BIL_TX, 0, st>>>(Vol<float>(out), Vol<float>(in), Img<float>(img), gs, gr, gc, size, nd);
    count_launch();
    return launch_status();
}

}  // namespace roo_b200

using namespace roo_b200;

static bool overlap(const void* a, size_t na, const void* b, size_t nb) {
    const char *pa = (const char*)a, *pb = (const char*)b;
    return pa < pb + nb && pb < pa + na;
}

extern "C" int roo_bilateral_filter_joint(const roo_image_t* out, const roo_image_t* in, const roo_image_t* img, int img_type,
                                          float gs, float gr, float gc, unsigned size, void* stream) {
    if (img_type != ROO_IMG_U8 && img_type != ROO_IMG_F32) return ROO_ERR_INVALID_ARGUMENT;
    if (!valid_image(out, 4) || !valid_image(in, 4) || !valid_image(img, img_type == ROO_IMG_U8 ? 1 : 4)) return ROO_ERR_INVALID_ARGUMENT;
    if (out->w != in->w || out->h != in->h || img->w != in->w || img->h != in->h || size > 64) return ROO_ERR_INVALID_ARGUMENT;
    if (overlap(out->ptr, out->pitch * out->h, in->ptr, in->pitch * in->h)) return ROO_ERR_INVALID_ARGUMENT;   // a tap would read filtered values
    const roo_volume_t vo{out->pitch, out->ptr, out->w, out->h, out->pitch * out->h, 1}, vi{in->pitch, in->ptr, in->w, in->h, in->pitch * in->h, 1};
    return launch_bilateral(vo, vi, *img, img_type, gs, gr, gc, (int)size, 1, as_stream(stream));
}

extern "C" int roo_bilateral_filter_volume(const roo_volume_t* out, const roo_volume_t* in, const roo_image_t* img, int img_type,
                                           float gs, float gr, float gc, unsigned size, int maxDisp, void* stream) {
    if (img_type != ROO_IMG_U8 && img_type != ROO_IMG_F32) return ROO_ERR_INVALID_ARGUMENT;
    if (!valid_volume(out, 4) || !valid_volume(in, 4) || !valid_image(img, img_type == ROO_IMG_U8 ? 1 : 4)) return ROO_ERR_INVALID_ARGUMENT;
    if (out->w != in->w || out->h != in->h || img->w != in->w || img->h != in->h || size > 64) return ROO_ERR_INVALID_ARGUMENT;
    if (maxDisp <= 0) return ROO_OK;
    if ((size_t)maxDisp > in->d || (size_t)maxDisp > out->d) return ROO_ERR_INVALID_ARGUMENT;
    if (overlap(out->ptr, out->img_pitch * out->d, in->ptr, in->img_pitch * in->d)) return ROO_ERR_INVALID_ARGUMENT;
    return launch_bilateral(*out, *in, *img, img_type, gs, gr, gc, (int)size, maxDisp, as_stream(stream));
}
