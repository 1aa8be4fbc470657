// The direct block matcher: roo::DenseStereo<{unsigned char, char}, unsigned char>
// (include/kangaroo/cu_dense_stereo.h:24-28; src/cu_dense_stereo.cu:209-253,376-406; scores include/kangaroo/patch_score.h:81-100
// for score_rad 0 and :257-298, SANDPatchScore<float,rad,ImgAccessRaw>, for score_rad 1..7).  No application calls it; it is
// here for completeness of the operator set of cu_dense_stereo.h.
//
// The reference runs one block of w threads per image row (w <= 1024) and every thread reads its (2 rad + 1)^2 patch and, per
// candidate, the right image's patch straight from global memory.  Here a CTA owns 128 pixels of a row and stages the rows
// y-rad..y+rad of both images in shared memory once: the left window (128 + 2 rad columns) and the right window that all its
// candidates can reach (up to 128 + 254 + 2 rad columns) -- every byte is then read from DRAM once per CTA row instead of once
// per candidate and tap -- and any width works.
//
// Results: the patch sums are exact integers, so only the order of the absolute-difference accumulation (row-major, as in the
// source) and the two divisions matter; default fp mode = the reference's div.approx forms (bit-identical to its kernel),
// IEEE mode = the CPU oracle.  Raw access is kept: a candidate may address up to 3 rad + 1 columns left of the image, i.e. the
// bytes that precede the row in memory (the previous row's tail for tightly packed images) -- inside the image's
// allocation, because only rows y >= 2 rad + 1 are scored.  Pixels at x >= maxDisp + rad never do.
#include "common.cuh"
#include "kernels.cuh"

namespace roo_b200 {

constexpr int DS_TX = 128, DS_MAXD = 254;

template <int RAD, bool IEEE>
__global__ void __launch_bounds__(DS_TX)
dense_stereo_kernel(Img<signed char> disp, Img<unsigned char> left, Img<unsigned char> right, int maxDispVal, float acceptThresh) {
    constexpr int W = 2 * RAD + 1, AREA = W * W, LW = DS_TX + 2 * RAD, RW = DS_TX + DS_MAXD + 2 * RAD;
    __shared__ unsigned char sl[W][LW], sr[W][RW];
    const int w = left.w, h = left.h, y = blockIdx.y, x0 = blockIdx.x * DS_TX, x = x0 + threadIdx.x;
    if (y < W || y >= h - W) {                      // border rows: InvalidValue<TD>::Value() == 0
        if (x < w) disp(x, y) = 0;
        return;
    }
    const int minD = min(maxDispVal, 0), maxD = max(0, maxDispVal);
    const int rstart = max(x0 - maxD, -W) - RAD;    // leftmost right-image column any candidate of this CTA can touch
    const int rend = min(x0 + DS_TX - 1 - minD, w - W) + RAD;
    const long long lbytes = (long long)left.pitch * h, rbytes = (long long)right.pitch * h;
    for (int i = threadIdx.x; i < W * LW; i += DS_TX) {
        const int rr = i / LW, cc = i - rr * LW;
        const long long off = (long long)(y - RAD + rr) * (long long)left.pitch + (x0 - RAD + cc);
        sl[rr][cc] = off >= 0 && off < lbytes ? (unsigned char)left.ptr[off] : 0;
    }
    const int rw = rend - rstart + 1;
    for (int i = threadIdx.x; i < W * rw; i += DS_TX) {
        const int rr = i / rw, cc = i - rr * rw;
        const long long off = (long long)(y - RAD + rr) * (long long)right.pitch + (rstart + cc);
        sr[rr][cc] = off >= 0 && off < rbytes ? (unsigned char)right.ptr[off] : 0;
    }
    __syncthreads();
    if (x >= w) return;
    int bestDisp = 0;
    if (W <= x && x < w - W) {
        float bestScore = 1e36f, sndBestScore = 1e37f;
        int sndBestDisp = 0;
        const int lo = max(minD, -((w - W) - x)), hi = min(maxD, x + W);
        const int lx = threadIdx.x + RAD;           // this pixel's column in the left window
        // SANDPatchScore, the reference build's way (SASS of KernDenseStereo<.., SANDPatchScore<float,rad>>): sum / area has become
        // sum * RC with RC = the float nearest 1/area, contracted into the subtractions -- (i1 - mean1) = FFMA(sum1, -RC, i1),
        // -(i2 - mean2) = FFMA(sum2, RC, -i2) -- so the means are never rounded on their own.  IEEE mode: source order.
        constexpr float RC = 1.0f / (float)AREA;
        float s1f = 0.0f, mean1 = 0.0f;
        if (RAD > 0) {
            int s1 = 0;
#pragma unroll
            for (int r = 0; r < W; ++r)
#pragma unroll
                for (int c = -RAD; c <= RAD; ++c) s1 += sl[r][lx + c];
            s1f = (float)s1;                        // the patch sums are exact integers in fp32, whatever the order
            mean1 = __fdiv_rn(s1f, (float)AREA);
        }
        for (int c = lo; c <= hi; ++c) {
            const int rx = x - c - rstart;          // the candidate's column in the right window
            float score;
            if (RAD == 0) {
                const float diff = (float)((int)sl[0][lx] - (int)sr[0][rx]);
                score = diff * diff;
            } else {
                // rows unrolled only for small patches: with all (2 rad + 1)^2 taps visible the compiler keeps the left
                // patch in registers across candidates (225 of them at rad 7) and spills
                int s2 = 0;
#pragma unroll(RAD <= 2 ? W : 1)
                for (int r = 0; r < W; ++r)
#pragma unroll
                    for (int k = -RAD; k <= RAD; ++k) s2 += sr[r][rx + k];
                const float s2f = (float)s2, mean2 = IEEE ? __fdiv_rn(s2f, (float)AREA) : 0.0f;
                score = 0.0f;
#pragma unroll(RAD <= 2 ? W : 1)
                for (int r = 0; r < W; ++r)
#pragma unroll
                    for (int k = -RAD; k <= RAD; ++k) {
                        const float a = (float)sl[r][lx + k], b = (float)sr[r][rx + k];
                        const float t = IEEE ? __fadd_rn(__fadd_rn(a, -mean1), -__fadd_rn(b, -mean2))
                                             : __fadd_rn(__fmaf_rn(s1f, -RC, a), __fmaf_rn(s2f, RC, -b));
                        score = __fadd_rn(score, fabsf(t));
                    }
            }
            if (score < bestScore) {
                sndBestDisp = bestDisp; sndBestScore = bestScore;
                bestDisp = c; bestScore = score;
            } else if (score <= sndBestScore) {
                sndBestDisp = c; sndBestScore = score;
            }
        }
        if (abs(bestDisp - sndBestDisp) > 1) {
            const float cd = ref_div<IEEE>(__fadd_rn(sndBestScore, -bestScore), bestScore);
            if (cd < acceptThresh) bestDisp = 0;
        }
    }
    disp(x, y) = (signed char)bestDisp;
}

template <int RAD>
static int launch_dense_rad(const roo_image_t& disp, const roo_image_t& l, const roo_image_t& r, int maxDisp, float thr, cudaStream_t st) {
    dim3 grid(cdiv((int)disp.w, DS_TX), (unsigned)disp.h);
    if (g_ieee_div.load())
        dense_stereo_kernel<RAD, true><<<grid, DS_TX, 0, st>>>(Img<signed char>(disp), Img<unsigned char>(l), Img<unsigned char>(r), maxDisp, thr);
    else
        dense_stereo_kernel<RAD, false><<<grid, DS_TX, 0, st>>>(Img<signed char>(disp), Img<unsigned char>(l), Img<unsigned char>(r), maxDisp, thr);
    count_launch();
    return launch_status();
}

}  // namespace roo_b200

using namespace roo_b200;

extern "C" int roo_dense_stereo(const roo_image_t* disp, int disp_signed, const roo_image_t* left, const roo_image_t* right, int maxDisp,
                                float acceptThresh, int score_rad, void* stream) {
    if (!valid_image(disp, 1) || !valid_image(left, 1) || !valid_image(right, 1)) return ROO_ERR_INVALID_ARGUMENT;
    if (disp->w != left->w || disp->h != left->h || right->w != left->w || right->h != left->h) return ROO_ERR_INVALID_ARGUMENT;
    if (score_rad < 0 || score_rad > 7) return ROO_ERR_INVALID_ARGUMENT;      // the reference launches nothing for other radii
    // TD maxDisp = 255 (127 for char) never terminates in the reference: its TD candidate counter wraps before exceeding it
    if (disp_signed ? (maxDisp < -128 || maxDisp > 126) : (maxDisp < 0 || maxDisp > DS_MAXD)) return ROO_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    switch (score_rad) {
    case 0: return launch_dense_rad<0>(*disp, *left, *right, maxDisp, acceptThresh, st);
    case 1: return launch_dense_rad<1>(*disp, *left, *right, maxDisp, acceptThresh, st);
    case 2: return launch_dense_rad<2>(*disp, *left, *right, maxDisp, acceptThresh, st);
    case 3: return launch_dense_rad<3>(*disp, *left, *right, maxDisp, acceptThresh, st);
    case 4: return launch_dense_rad<4>(*disp, *left, *right, maxDisp, acceptThresh, st);
    case 5: return launch_dense_rad<5>(*disp, *left, *right, maxDisp, acceptThresh, st);
    case 6: return launch_dense_rad<6>(*disp, *left, *right, maxDisp, acceptThresh, st);
    default: return launch_dense_rad<7>(*disp, *left, *right, maxDisp, acceptThresh, st);
    }
}
