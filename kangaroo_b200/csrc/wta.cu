// Winner-takes-all, parabola refinement and left-right consistency on the reference's pitched
// layouts (granular drop-in operators).
//
// Replaces, from src/cu_dense_stereo.cu of the reference: KernCostVolMinimum (:25-43, :735-755),
// KernCostVolMinimumSubpix (:66-109), KernLeftRightCheck (:512-532), KernDenseStereoSubpixelRefine
// (:580-619, score patch_score.h:257-298).  All kernels here are bounds-guarded (the reference's
// CostVolMinimum is not, Q5) and take a stream.  In the fused engine the WTA/parabola step does not
// run as a kernel at all: it is the epilogue of the last aggregation sweep (sgm.cu).
#include "common.cuh"
#include "kernels.cuh"

namespace roo_b200 {

constexpr int WTA_TX = 128;

template <typename Tvol>
__device__ __forceinline__ Tvol ldvol(const Vol<Tvol>& v, int x, int y, int d) { return v(x, y, d); }

// ---- CostVolMinimum<Tdisp,Tvol> -------------------------------------------------------------------
template <typename Tdisp, typename Tvol>
__global__ void __launch_bounds__(WTA_TX)
costvol_min_kernel(Img<Tdisp> disp, Vol<Tvol> vol, unsigned maxDispVal) {
    const int x = blockIdx.x * WTA_TX + threadIdx.x, y = blockIdx.y;
    if (x >= disp.w) return;
    Tdisp bestd = 0;
    Tvol bestc = vol(x, y, 0);
    const int maxDisp = (int)min(maxDispVal, (unsigned)(x + 1));
    for (int d = 1; d < maxDisp; ++d) {
        const Tvol c = vol(x, y, d);
        if (c < bestc) { bestc = c; bestd = (Tdisp)d; }  // Tdisp = char wraps past 127 like the reference
    }
    disp(x, y) = bestd;
}

template <typename Tdisp, typename Tvol>
static int cvm_launch(const roo_image_t* disp, const roo_volume_t* vol, unsigned maxDisp, cudaStream_t st) {
    dim3 grid(cdiv((int)disp->w, WTA_TX), (unsigned)disp->h);
    costvol_min_kernel<Tdisp, Tvol><<<grid, WTA_TX, 0, st>>>(Img<Tdisp>(*disp), Vol<Tvol>(*vol), maxDisp);
    count_launch();
    return launch_status();
}

// ---- CostVolMinimum(Image<float>, Volume<CostVolElem>) ---------------------------------------------
template <bool IEEE>
__global__ void __launch_bounds__(WTA_TX)
costvol_min_elem_kernel(Img<float> disp, Vol<roo_costvolelem_t> vol) {
    const int x = blockIdx.x * WTA_TX + threadIdx.x, y = blockIdx.y;
    if (x >= disp.w) return;
    float bestd = 0.0f, bestc = 1E30f;
    for (int d = 0; d < vol.d; ++d) {
        const roo_costvolelem_t e = vol(x, y, d);
        const float c = ref_div<IEEE>(e.sum, (float)e.n);  // n == 0: inf/NaN never wins (cu_dense_stereo.cu:748)
        if (c < bestc) { bestc = c; bestd = (float)d; }
    }
    disp(x, y) = bestd;
}

// ---- CostVolMinimumSubpix ---------------------------------------------------------------------------
template <bool IEEE>
__global__ void __launch_bounds__(WTA_TX)
costvol_min_subpix_kernel(Img<float> disp, Vol<float> vol, unsigned maxDispVal, int sdi) {
    const int x = blockIdx.x * WTA_TX + threadIdx.x, y = blockIdx.y;
    if (x >= disp.w) return;
    int bestd = 0;
    float bestc = 1E10f;
    for (int d = 0; d < (int)maxDispVal; ++d) {
        const int xr = x + sdi * d;
        if (0 <= xr && xr < vol.w) {
            const float c = vol(x, y, d);
            if (c < bestc) { bestc = c; bestd = d; }
        }
    }
    float out = (float)bestd;
    const int bestxr = x + sdi * bestd;
    if (0 < bestxr && bestxr < vol.w - 1 && bestd + 1 < vol.d) {   // bestd+1 == vol.d: the reference reads out of bounds
        const float sl = vol(x, y, max(bestd - 1, 0));             // float->unsigned saturation in the reference (Q7)
        const float sr = vol(x, y, bestd + 1);
        const float sub = parabola_vertex<IEEE>((float)bestd, bestc, sl, sr);
        if ((float)(bestd - 1) < sub && sub < (float)(bestd + 1)) out = sub;
    }
    disp(x, y) = out;
}

// ---- CostVolMinimumSquarePenaltySubpix (cu_dense_stereo.cu:122-174) --------------------------------------
// argmin_d of  (lastd - d)^2 / (2 theta) + lambda * vol(x,y,d)  (the coupling step of the applications' variational
// refinement, stereo/main.cpp:376), then the same parabola as CostVolMinimumSubpix on the penalised costs.
// Reference SASS (fast-math): inv2theta = MUFU.RCP(theta + theta); c = FFMA(ddif, inv2theta * ddif, lambda * vol).
template <bool IEEE>
__device__ __forceinline__ float sqpen_cost(float lastd, float d, float inv2theta, float lambda, float v) {
    const float ddif = __fadd_rn(lastd, -d);
    if (IEEE) return __fadd_rn(__fmul_rn(__fmul_rn(inv2theta, ddif), ddif), __fmul_rn(lambda, v));
    return __fmaf_rn(ddif, __fmul_rn(inv2theta, ddif), __fmul_rn(v, lambda));
}
template <bool IEEE>
__global__ void __launch_bounds__(WTA_TX)
costvol_min_sqpen_subpix_kernel(Img<float> imga, Vol<float> vol, Img<float> imgd, unsigned maxDispVal, int sdi, float lambda,
                                float theta) {
    const int x = blockIdx.x * WTA_TX + threadIdx.x, y = blockIdx.y;
    if (x >= imga.w) return;
    const float lastd = imgd(x, y);
    const float inv2theta = IEEE ? __fdiv_rn(1.0f, __fmul_rn(2.0f, theta)) : rcp_approx_ftz(__fadd_rn(theta, theta));
    int bestd = 0;
    float bestc = sqpen_cost<IEEE>(lastd, 0.0f, inv2theta, lambda, vol(x, y, 0));
    for (int d = 1; d < (int)maxDispVal; ++d) {
        const int xr = x + sdi * d;
        if (0 <= xr && xr < vol.w) {
            const float c = sqpen_cost<IEEE>(lastd, (float)d, inv2theta, lambda, vol(x, y, d));
            if (c < bestc) { bestc = c; bestd = d; }
        }
    }
    float out = (float)bestd;
    const int bestxr = x + sdi * bestd;
    if (0 < bestxr && bestxr < vol.w - 1 && bestd + 1 < vol.d) {   // bestd+1 == vol.d: the reference reads out of bounds (Q7)
        const float dl = (float)(bestd - 1), dr = (float)(bestd + 1);
        const float sl = sqpen_cost<IEEE>(lastd, dl, inv2theta, lambda, vol(x, y, max(bestd - 1, 0)));   // index saturates at 0 (Q7)
        const float sr = sqpen_cost<IEEE>(lastd, dr, inv2theta, lambda, vol(x, y, bestd + 1));
        const float sub = parabola_vertex<IEEE>((float)bestd, bestc, sl, sr);
        if (dl < sub && sub < dr) out = sub;
    }
    imga(x, y) = out;
}

// ---- FilterDispGrad (cu_dense_stereo.cu:793-812) ----------------------------------------------------------
// out(x,y) = |central-difference gradient of G|^2 < threshold ? in(x,y) : -1, G = the contents of `out` BEFORE the call
// (the applications call it in place, main.cpp:457, where the reference races with itself; here G is a snapshot).
// Reference SASS: dx = (G(x+1,y) - G(x-1,y)) * 0.5, dy likewise, m = FFMA(dx, dx, dy * dy), valid = !(m >= threshold).
// The reference reads outside the image on the border pixels (undefined); here out-of-image neighbours clamp to the edge.
__global__ void __launch_bounds__(WTA_TX)
filter_disp_grad_kernel(Img<float> out, Img<float> grad, Img<float> in, float threshold, size_t out_pair, size_t grad_pair,
                        size_t in_pair) {
    const int x = blockIdx.x * WTA_TX + threadIdx.x, y = blockIdx.y;
    if (x >= out.w) return;
    out.ptr += blockIdx.z * out_pair; grad.ptr += blockIdx.z * grad_pair; in.ptr += blockIdx.z * in_pair;   // bytes between pairs
    const float* row = grad.row(y);
    const float dx = __fmul_rn(__fadd_rn(row[min(x + 1, grad.w - 1)], -row[max(x - 1, 0)]), 0.5f);
    const float dy = __fmul_rn(__fadd_rn(grad(x, min(y + 1, grad.h - 1)), -grad(x, max(y - 1, 0))), 0.5f);
    const float m = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
    out(x, y) = m < threshold ? in(x, y) : -1.0f;
}

// ---- LeftRightCheck -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(WTA_TX)
lr_check_f32_kernel(char* dispL, size_t pitchL, size_t batchL, const char* dispR, size_t pitchR, size_t batchR, int w,
                    float sd, float maxDiff) {
    const int x = blockIdx.x * WTA_TX + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    float* pl = reinterpret_cast<float*>(dispL + (size_t)blockIdx.z * batchL + (size_t)y * pitchL) + x;
    const float* rrow = reinterpret_cast<const float*>(dispR + (size_t)blockIdx.z * batchR + (size_t)y * pitchR);
    const float dl = *pl;
    const float xr = __fmaf_rn(sd, dl, (float)x);  // x + sd*dl, contracted like the reference build
    const float nanv = __int_as_float(0x7fffffff);
    if (0.0f <= xr && xr < (float)w) {
        const float dr = rrow[(int)xr];
        if (!isfinite(dr) || fabsf(dl - dr) > maxDiff) *pl = nanv;
    } else {
        *pl = nanv;  // also the NaN-dl case: both comparisons are false
    }
}

__global__ void __launch_bounds__(WTA_TX)
lr_check_i8_kernel(Img<signed char> dispL, Img<signed char> dispR, float sd, float maxDiff) {
    const int x = blockIdx.x * WTA_TX + threadIdx.x, y = blockIdx.y;
    if (x >= dispL.w) return;
    const signed char dl = dispL(x, y);
    // `const char xr = x + sd*dl`: cvt.rzi.s32.f32 then truncation to 8 bits (wraps for x > 127)
    const signed char xr = (signed char)(int)__fmaf_rn(sd, (float)dl, (float)x);
    if (0 <= xr && (int)xr < dispR.w) {
        const signed char dr = dispR((int)xr, y);
        // InvalidValue<char>::IsValid(v) == !v while Value() == 0 (InvalidValue.h:50-59, Q10)
        if (dr != 0 || (float)abs((int)dl - (int)dr) > maxDiff) dispL(x, y) = 0;
    } else {
        dispL(x, y) = 0;
    }
}

int launch_lr_check_f32(float* dispL, size_t pitchL, const float* dispR, size_t pitchR, int w, int h, int batch,
                        size_t batchL, size_t batchR, float sd, float maxDiff, cudaStream_t st) {
    dim3 grid(cdiv(w, WTA_TX), h, batch);
    lr_check_f32_kernel<<<grid, WTA_TX, 0, st>>>((char*)dispL, pitchL, batchL, (const char*)dispR, pitchR, batchR, w, sd,
                                                 maxDiff);
    count_launch();
    return launch_status();
}

// ---- DenseStereoSubpixelRefine -------------------------------------------------------------------------
// SANDPatchScore<float,2,ImgAccessRaw> (patch_score.h:257-298) on unsigned char images.
__device__ __forceinline__ float sand5x5(const Img<unsigned char>& i1, int x1, int y1, const Img<unsigned char>& i2,
                                         int x2, int y2) {
    float sum1 = 0.0f, sum2 = 0.0f;
#pragma unroll
    for (int r = -2; r <= 2; ++r)
#pragma unroll
        for (int c = -2; c <= 2; ++c) {
            sum1 += (float)i1(x1 + c, y1 + r);
            sum2 += (float)i2(x2 + c, y2 + r);
        }
    const float mean1 = __fdividef(sum1, 25.0f), mean2 = __fdividef(sum2, 25.0f);
    float sad = 0.0f;
#pragma unroll
    for (int r = -2; r <= 2; ++r)
#pragma unroll
        for (int c = -2; c <= 2; ++c) {
            const float a = (float)i1(x1 + c, y1 + r), b = (float)i2(x2 + c, y2 + r);
            sad += fabsf((a - mean1) - (b - mean2));
        }
    return sad;
}

__global__ void __launch_bounds__(WTA_TX)
subpixel_refine_kernel(Img<float> out, Img<unsigned char> disp, Img<unsigned char> left, Img<unsigned char> right) {
    const int x = blockIdx.x * WTA_TX + threadIdx.x, y = blockIdx.y;
    const int w = disp.w, h = disp.h;
    if (x >= w) return;
    const int bestDisp = disp(x, y);
    const float nanv = __int_as_float(0x7fffffff);
    float res = nanv;
    // guard where the reference reads outside the images (Q8)
    const bool inside = y >= 2 && y + 2 < h && x >= 2 && x + 2 < w && x - bestDisp - 3 >= 0 && x - bestDisp + 3 < w;
    if (inside) {
        const float d1 = (float)(bestDisp + 1), d2 = (float)bestDisp, d3 = (float)(bestDisp - 1);
        const float s1 = sand5x5(left, x, y, right, x - (bestDisp + 1), y);
        const float s2 = sand5x5(left, x, y, right, x - bestDisp, y);
        const float s3 = sand5x5(left, x, y, right, x - (bestDisp - 1), y);
        const float denom = (d1 - d2) * (d1 - d3) * (d2 - d3);
        const float A = __fdividef(d3 * (s2 - s1) + d2 * (s1 - s3) + d1 * (s3 - s2), denom);
        const float B = __fdividef(d3 * d3 * (s1 - s2) + d2 * d2 * (s3 - s1) + d1 * d1 * (s2 - s3), denom);
        const float newDisp = __fdividef(-B, 2.0f * A);
        if (d3 < newDisp && newDisp < d1) res = newDisp;
    }
    out(x, y) = res;
}

}  // namespace roo_b200

using namespace roo_b200;

extern "C" int roo_costvol_minimum(const roo_image_t* disp, int disp_type, const roo_volume_t* vol, int vol_type,
                                   unsigned maxDisp, void* stream) {
    static const size_t vsz[] = {2, 4, 4, 4, 1};
    if (disp_type != ROO_DISP_I8 && disp_type != ROO_DISP_F32) return ROO_ERR_INVALID_ARGUMENT;
    if (vol_type < ROO_VOL_U16 || vol_type > ROO_VOL_U8) return ROO_ERR_INVALID_ARGUMENT;
    if (!valid_image(disp, disp_type == ROO_DISP_I8 ? 1 : 4) || !valid_volume(vol, vsz[vol_type])) return ROO_ERR_INVALID_ARGUMENT;
    if (disp->w != vol->w || disp->h != vol->h || maxDisp > vol->d) return ROO_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    // instantiation set of the reference (cu_dense_stereo.cu:54-60)
    if (disp_type == ROO_DISP_I8) {
        switch (vol_type) {
            case ROO_VOL_F32: return cvm_launch<signed char, float>(disp, vol, maxDisp, st);
            case ROO_VOL_I32: return cvm_launch<signed char, int>(disp, vol, maxDisp, st);
            case ROO_VOL_U32: return cvm_launch<signed char, unsigned>(disp, vol, maxDisp, st);
            case ROO_VOL_U16: return cvm_launch<signed char, unsigned short>(disp, vol, maxDisp, st);
            case ROO_VOL_U8: return cvm_launch<signed char, unsigned char>(disp, vol, maxDisp, st);
        }
    } else {
        if (vol_type == ROO_VOL_F32) return cvm_launch<float, float>(disp, vol, maxDisp, st);
        if (vol_type == ROO_VOL_U16) return cvm_launch<float, unsigned short>(disp, vol, maxDisp, st);
    }
    return ROO_ERR_UNSUPPORTED;
}

extern "C" int roo_costvol_minimum_elem(const roo_image_t* disp, const roo_volume_t* vol, void* stream) {
    if (!valid_image(disp, 4) || !valid_volume(vol, 8) || disp->w != vol->w || disp->h != vol->h) return ROO_ERR_INVALID_ARGUMENT;
    dim3 grid(cdiv((int)disp->w, WTA_TX), (unsigned)disp->h);
    if (g_ieee_div.load())
        costvol_min_elem_kernel<true><<<grid, WTA_TX, 0, as_stream(stream)>>>(Img<float>(*disp), Vol<roo_costvolelem_t>(*vol));
    else
        costvol_min_elem_kernel<false><<<grid, WTA_TX, 0, as_stream(stream)>>>(Img<float>(*disp), Vol<roo_costvolelem_t>(*vol));
    count_launch();
    return launch_status();
}

extern "C" int roo_costvol_minimum_subpix(const roo_image_t* disp, const roo_volume_t* vol, unsigned maxDisp, float sd,
                                          void* stream) {
    if (!valid_image(disp, 4) || !valid_volume(vol, 4) || disp->w != vol->w || disp->h != vol->h || maxDisp > vol->d)
        return ROO_ERR_INVALID_ARGUMENT;
    if (sd != -1.0f && sd != 1.0f) return ROO_ERR_UNSUPPORTED;
    dim3 grid(cdiv((int)disp->w, WTA_TX), (unsigned)disp->h);
    const int sdi = sd < 0 ? -1 : 1;
    if (g_ieee_div.load())
        costvol_min_subpix_kernel<true><<<grid, WTA_TX, 0, as_stream(stream)>>>(Img<float>(*disp), Vol<float>(*vol), maxDisp, sdi);
    else
        costvol_min_subpix_kernel<false><<<grid, WTA_TX, 0, as_stream(stream)>>>(Img<float>(*disp), Vol<float>(*vol), maxDisp, sdi);
    count_launch();
    return launch_status();
}

extern "C" int roo_costvol_minimum_square_penalty_subpix(const roo_image_t* imga, const roo_volume_t* vol, const roo_image_t* imgd,
                                                         unsigned maxDisp, float sd, float lambda, float theta, void* stream) {
    if (!valid_image(imga, 4) || !valid_image(imgd, 4) || !valid_volume(vol, 4) || imga->w != vol->w || imga->h != vol->h ||
        imgd->w != vol->w || imgd->h != vol->h || maxDisp > vol->d)
        return ROO_ERR_INVALID_ARGUMENT;
    if (sd != -1.0f && sd != 1.0f) return ROO_ERR_UNSUPPORTED;
    dim3 grid(cdiv((int)imga->w, WTA_TX), (unsigned)imga->h);
    const int sdi = sd < 0 ? -1 : 1;
    if (g_ieee_div.load())
        costvol_min_sqpen_subpix_kernel<true><<<grid, WTA_TX, 0, as_stream(stream)>>>(Img<float>(*imga), Vol<float>(*vol), Img<float>(*imgd), maxDisp, sdi, lambda, theta);
    else
        costvol_min_sqpen_subpix_kernel<false><<<grid, WTA_TX, 0, as_stream(stream)>>>(Img<float>(*imga), Vol<float>(*vol), Img<float>(*imgd), maxDisp, sdi, lambda, theta);
    count_launch();
    return launch_status();
}

namespace roo_b200 {
// out(x,y) from the gradient of `grad` (must not overlap out) and the values of `in` (may be `grad`)
int launch_filter_disp_grad(const roo_image_t& out, const roo_image_t& grad, const roo_image_t& in, float threshold, cudaStream_t st,
                            int batch, size_t out_pair, size_t grad_pair, size_t in_pair) {
    dim3 grid(cdiv((int)out.w, WTA_TX), (unsigned)out.h, (unsigned)batch);
    filter_disp_grad_kernel<<<grid, WTA_TX, 0, st>>>(Img<float>(out), Img<float>(grad), Img<float>(in), threshold, out_pair,
                                                     grad_pair, in_pair);
    count_launch();
    return launch_status();
}
}  // namespace roo_b200

extern "C" int roo_filter_disp_grad(const roo_image_t* out, const roo_image_t* in, float threshold, void* stream) {
    if (!valid_image(out, 4) || !valid_image(in, 4) || out->w != in->w || out->h != in->h) return ROO_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    // the gradient is taken of what `out` holds when the call is made: snapshot it (stream-ordered temporary), then write
    const size_t tp = out->w * sizeof(float);
    float* snap = nullptr;
    ROO_CUDA_TRY(cudaMallocAsync((void**)&snap, tp * out->h, st));
    int rc = (int)cudaMemcpy2DAsync(snap, tp, out->ptr, out->pitch, tp, out->h, cudaMemcpyDeviceToDevice, st);
    const roo_image_t g{tp, snap, out->w, out->h};
    // `in` aliasing `out` (the applications' call): its values are the snapshot's
    const char *ob = (const char*)out->ptr, *ib = (const char*)in->ptr;
    const bool alias = ob < ib + in->pitch * in->h && ib < ob + out->pitch * out->h;
    if (alias && (in->ptr != out->ptr || in->pitch != out->pitch)) rc = rc ? rc : ROO_ERR_INVALID_ARGUMENT;   // partial overlap
    if (rc == 0) rc = launch_filter_disp_grad(*out, g, alias ? g : *in, threshold, st);
    const cudaError_t fe = cudaFreeAsync(snap, st);
    return rc != 0 ? rc : (int)fe;
}

extern "C" int roo_dense_stereo_subpixel_refine(const roo_image_t* out, const roo_image_t* disp, const roo_image_t* left,
                                                const roo_image_t* right, void* stream) {
    if (!valid_image(out, 4) || !valid_image(disp, 1) || !valid_image(left, 1) || !valid_image(right, 1))
        return ROO_ERR_INVALID_ARGUMENT;
    if (out->w != disp->w || out->h != disp->h || left->w != disp->w || left->h != disp->h || right->w != disp->w ||
        right->h != disp->h)
        return ROO_ERR_INVALID_ARGUMENT;
    dim3 grid(cdiv((int)disp->w, WTA_TX), (unsigned)disp->h);
    subpixel_refine_kernel<<<grid, WTA_TX, 0, as_stream(stream)>>>(Img<float>(*out), Img<unsigned char>(*disp),
                                                                   Img<unsigned char>(*left), Img<unsigned char>(*right));
    count_launch();
    return launch_status();
}

extern "C" int roo_left_right_check_f32(const roo_image_t* dispL, const roo_image_t* dispR, float sd, float maxDiff,
                                        void* stream) {
    if (!valid_image(dispL, 4) || !valid_image(dispR, 4) || dispL->h > dispR->h) return ROO_ERR_INVALID_ARGUMENT;
    // the reference bounds xr by dispR.w
    if (dispR->w != dispL->w) return ROO_ERR_INVALID_ARGUMENT;
    return launch_lr_check_f32((float*)dispL->ptr, dispL->pitch, (const float*)dispR->ptr, dispR->pitch, (int)dispL->w,
                               (int)dispL->h, 1, 0, 0, sd, maxDiff, as_stream(stream));
}

extern "C" int roo_left_right_check_i8(const roo_image_t* dispL, const roo_image_t* dispR, int sd, int maxDiff,
                                       void* stream) {
    if (!valid_image(dispL, 1) || !valid_image(dispR, 1) || dispL->h > dispR->h) return ROO_ERR_INVALID_ARGUMENT;
    dim3 grid(cdiv((int)dispL->w, WTA_TX), (unsigned)dispL->h);
    lr_check_i8_kernel<<<grid, WTA_TX, 0, as_stream(stream)>>>(Img<signed char>(*dispL), Img<signed char>(*dispR),
                                                               (float)sd, (float)maxDiff);
    count_launch();
    return launch_status();
}
