// Census transform and Hamming matching cost (granular operators on the reference's pitched
// layouts + the engine's internal disparity-innermost byte cost volume).
//
// Replaces src/cu_census.cu of the reference: KernCensus9x7 (:18-46), KernCensus11x11 (:52-110),
// KernCensus16x16 (:116-177), KernCensusStereo (:226-259), KernCensusStereoVolume (:272-299).
// Nothing here is derived from those kernels' structure: the window is staged once per CTA in
// shared memory (clamp-to-edge applied while staging), descriptors are built from registers and
// stored with one 8/16/32-byte vector store per pixel, rows are processed by independent CTAs
// (no w <= 1024 limit), and every launch takes an explicit stream and a batch dimension.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"

namespace roo_b200 {

// ------------------------------------------------------------------------------------------------
// Census
// ------------------------------------------------------------------------------------------------
template <int WIN> struct CensusGeom;
template <> struct CensusGeom<ROO_WIN_9x7>   { static constexpr int CX0 = -4, CX1 = 4, RY0 = -3, RY1 = 3, WORDS = 1; };
template <> struct CensusGeom<ROO_WIN_11x11> { static constexpr int CX0 = -5, CX1 = 5, RY0 = -5, RY1 = 5, WORDS = 2; };
template <> struct CensusGeom<ROO_WIN_16x16> { static constexpr int CX0 = -4, CX1 = 3, RY0 = -8, RY1 = 7, WORDS = 4; };

constexpr int CENSUS_TX = 32, CENSUS_TY = 8;

template <typename Tin, int WIN>
__global__ void __launch_bounds__(CENSUS_TX * CENSUS_TY)
census_kernel(char* __restrict__ out, size_t out_pitch, size_t out_batch, const char* __restrict__ in, size_t in_pitch,
              size_t in_batch, int w, int h) {
    using G = CensusGeom<WIN>;
    constexpr int SW = CENSUS_TX + G::CX1 - G::CX0;
    constexpr int SH = CENSUS_TY + G::RY1 - G::RY0;
    __shared__ Tin tile[SH][SW + 1];

    in += (size_t)blockIdx.z * in_batch;
    out += (size_t)blockIdx.z * out_batch;
    const int x0 = blockIdx.x * CENSUS_TX, y0 = blockIdx.y * CENSUS_TY;
    const int tid = threadIdx.y * CENSUS_TX + threadIdx.x;

    // stage the window tile, clamp-to-edge (Image.h:297-303 GetWithClampedRange)
    for (int i = tid; i < SH * SW; i += CENSUS_TX * CENSUS_TY) {
        const int ty = i / SW, tx = i - ty * SW;
        const int gx = clampi(x0 + tx + G::CX0, 0, w - 1);
        const int gy = clampi(y0 + ty + G::RY0, 0, h - 1);
        tile[ty][tx] = *reinterpret_cast<const Tin*>(in + (size_t)gy * in_pitch + (size_t)gx * sizeof(Tin));
    }
    __syncthreads();

    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    const int cx = threadIdx.x - G::CX0, cy = threadIdx.y - G::RY0;  // centre inside the tile
    const Tin p = tile[cy][cx];

    if (WIN == ROO_WIN_9x7) {
        // bit (r+3)*9 + (c+4)  (cu_census.cu:29-41)
        unsigned long long o = 0;
#pragma unroll
        for (int r = -3; r <= 3; ++r)
#pragma unroll
            for (int c = -4; c <= 4; ++c)
                o |= (unsigned long long)(tile[cy + r][cx + c] < p) << ((r + 3) * 9 + (c + 4));
        *reinterpret_cast<unsigned long long*>(out + (size_t)y * out_pitch + (size_t)x * 8) = o;
    } else if (WIN == ROO_WIN_11x11) {
        // x: rows -5..-1 (bits 0..54) + row 0 c=-5..0 (55..60); y: row 0 c=1..5 (0..4) + rows 1..5 (5..59)
        unsigned long long ox = 0, oy = 0;
#pragma unroll
        for (int r = -5; r < 0; ++r)
#pragma unroll
            for (int c = -5; c <= 5; ++c)
                ox |= (unsigned long long)(tile[cy + r][cx + c] < p) << ((r + 5) * 11 + (c + 5));
#pragma unroll
        for (int c = -5; c <= 0; ++c) ox |= (unsigned long long)(tile[cy][cx + c] < p) << (55 + c + 5);
#pragma unroll
        for (int c = 1; c <= 5; ++c) oy |= (unsigned long long)(tile[cy][cx + c] < p) << (c - 1);
#pragma unroll
        for (int r = 1; r <= 5; ++r)
#pragma unroll
            for (int c = -5; c <= 5; ++c)
                oy |= (unsigned long long)(tile[cy + r][cx + c] < p) << (5 + (r - 1) * 11 + (c + 5));
        *reinterpret_cast<ulonglong2*>(out + (size_t)y * out_pitch + (size_t)x * 16) = make_ulonglong2(ox, oy);
    } else {
        // four words of 4 rows x 8 cols, low 32 bits each (cu_census.cu:126-174)
        unsigned o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned acc = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = -4; c < 4; ++c)
                    acc |= (unsigned)(tile[cy - 8 + 4 * k + r][cx + c] < p) << (r * 8 + (c + 4));
            o[k] = acc;
        }
        struct { unsigned long long x, y, z, w; } v = {o[0], o[1], o[2], o[3]};
        char* dst = out + (size_t)y * out_pitch + (size_t)x * 32;  // ulong4 is 16-byte aligned
        reinterpret_cast<ulonglong2*>(dst)[0] = make_ulonglong2(v.x, v.y);
        reinterpret_cast<ulonglong2*>(dst)[1] = make_ulonglong2(v.z, v.w);
    }
}

template <typename Tin>
static int census_dispatch(char* out, size_t out_pitch, size_t out_batch, const char* in, size_t in_pitch,
                           size_t in_batch, int w, int h, int batch, int window, cudaStream_t st) {
    dim3 block(CENSUS_TX, CENSUS_TY), grid(cdiv(w, CENSUS_TX), cdiv(h, CENSUS_TY), batch);
    switch (window) {
        case ROO_WIN_9x7:
            census_kernel<Tin, ROO_WIN_9x7><<<grid, block, 0, st>>>(out, out_pitch, out_batch, in, in_pitch, in_batch, w, h);
            break;
        case ROO_WIN_11x11:
            census_kernel<Tin, ROO_WIN_11x11><<<grid, block, 0, st>>>(out, out_pitch, out_batch, in, in_pitch, in_batch, w, h);
            break;
        case ROO_WIN_16x16:
            census_kernel<Tin, ROO_WIN_16x16><<<grid, block, 0, st>>>(out, out_pitch, out_batch, in, in_pitch, in_batch, w, h);
            break;
        default: return ROO_ERR_INVALID_ARGUMENT;
    }
    count_launch();
    return launch_status();
}

int launch_census(char* out, size_t out_pitch, size_t out_batch, const char* in, size_t in_pitch, size_t in_batch,
                  int w, int h, int batch, int window, int in_type, cudaStream_t st) {
    if (in_type == ROO_IMG_U8)
        return census_dispatch<unsigned char>(out, out_pitch, out_batch, in, in_pitch, in_batch, w, h, batch, window, st);
    if (in_type == ROO_IMG_F32)
        return census_dispatch<float>(out, out_pitch, out_batch, in, in_pitch, in_batch, w, h, batch, window, st);
    return ROO_ERR_INVALID_ARGUMENT;
}

// ------------------------------------------------------------------------------------------------
// CensusStereo: direct Hamming WTA on unsigned long descriptors (cu_census.cu:226-259)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
census_stereo_kernel(Img<signed char> disp, Img<unsigned long long> left, Img<unsigned long long> right, int maxDispVal) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= disp.w) return;
    const unsigned long long p = left(x, y);
    const unsigned long long* __restrict__ rrow = right.row(y);
    unsigned bestScore = 0xFFFFF;
    int bestDisp = 0;  // InvalidValue<char>::Value()
    int minDisp = max(min(maxDispVal, 0), x - (left.w - 1));
    int maxDisp = min(max(0, maxDispVal), x);
    for (int d = minDisp; d < maxDisp; ++d) {
        const unsigned score = hamming_word<false>(p, rrow[x - d]);
        if (score < bestScore) { bestScore = score; bestDisp = d; }
    }
    disp(x, y) = (signed char)bestDisp;
}

// ------------------------------------------------------------------------------------------------
// CensusStereoVolume into the reference's d-outermost volume (cu_census.cu:272-299)
// One CTA per 128-pixel row segment; the right-image descriptors the segment can reach
// (x + sd*d, d < maxDisp) are staged in shared memory once; each store instruction writes 32
// consecutive x of one disparity slice (128 B for float).
// ------------------------------------------------------------------------------------------------
constexpr int CSV_TX = 128;

template <int WORDS, typename Tvol, bool POPC64>
__global__ void __launch_bounds__(CSV_TX)
census_stereo_volume_kernel(Vol<Tvol> vol, const char* __restrict__ left, const char* __restrict__ right,
                            size_t l_pitch, size_t r_pitch, int rw, int maxDisp, int sdi) {
    extern __shared__ unsigned long long cache_r[];  // [(CSV_TX + maxDisp - 1) * WORDS], word-major
    const int x0 = blockIdx.x * CSV_TX, y = blockIdx.y;
    const int x = x0 + threadIdx.x;
    const int span = CSV_TX + maxDisp - 1;
    // right-image x range touched by this segment: sd=-1: [x0-(maxDisp-1), x0+TX) ; sd=+1: [x0, x0+TX+maxDisp-1)
    const int rx0 = sdi < 0 ? x0 - (maxDisp - 1) : x0;
    const unsigned long long* rrow = reinterpret_cast<const unsigned long long*>(right + (size_t)y * r_pitch);
    for (int i = threadIdx.x; i < span; i += CSV_TX) {
        const int gx = rx0 + i;
        const bool ok = gx >= 0 && gx < rw;
#pragma unroll
        for (int k = 0; k < WORDS; ++k) cache_r[k * span + i] = ok ? rrow[(size_t)gx * WORDS + k] : 0ull;
    }
    __syncthreads();
    if (x >= vol.w) return;
    unsigned long long p[WORDS];
    const unsigned long long* lrow = reinterpret_cast<const unsigned long long*>(left + (size_t)y * l_pitch);
#pragma unroll
    for (int k = 0; k < WORDS; ++k) p[k] = lrow[(size_t)x * WORDS + k];
    const float inv_bits = 1.0f / (float)(WORDS * 64);  // power of two: exact, equals the reference's divide
    for (int d = 0; d < maxDisp; ++d) {
        const int xd = x + sdi * d;
        float score = 0.5f;
        if (xd >= 0 && xd < rw) {
            const int i = xd - rx0;
            unsigned hd = 0;
#pragma unroll
            for (int k = 0; k < WORDS; ++k) hd += hamming_word<POPC64>(p[k], cache_r[k * span + i]);
            score = (float)hd * inv_bits;
        }
        vol(x, y, d) = (Tvol)score;  // Tvol = unsigned short truncates to 0 (reference quirk Q2)
    }
}

template <int WORDS, typename Tvol>
static int csv_launch(const roo_volume_t* vol, const roo_image_t* l, const roo_image_t* r, int maxDisp, int sdi,
                      int popc_mode, cudaStream_t st) {
    dim3 grid(cdiv((int)l->w, CSV_TX), (unsigned)l->h), block(CSV_TX);
    const size_t smem = (size_t)(CSV_TX + maxDisp - 1) * WORDS * 8;
    if (smem > 200 * 1024) return ROO_ERR_UNSUPPORTED;
    auto kern = popc_mode == ROO_POPC64 ? census_stereo_volume_kernel<WORDS, Tvol, true>
                                        : census_stereo_volume_kernel<WORDS, Tvol, false>;
    if (smem > 48 * 1024) ROO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(Vol<Tvol>(*vol), (const char*)l->ptr, (const char*)r->ptr, l->pitch, r->pitch, (int)r->w,
                                    maxDisp, sdi);
    count_launch();
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Engine-internal raw Hamming cost: C8[pair][y][x][DP] bytes, disparity innermost, so that one warp
// of an aggregation sweep reads its 32*DPL disparities of a pixel with one coalesced load.
// Stores the raw count h; the sweeps multiply by 1/bits (exact).  d > x (no right pixel) stores
// bits/2, the reference's 0.5.
// One CTA per (row, 128-pixel segment): the descriptors the segment touches are staged in shared memory
// once (coalesced); a thread owns 4 consecutive disparities of one pixel (one 32-bit store) and consecutive
// threads write consecutive words, so the 128*DP-byte output of the CTA is one contiguous coalesced stream.
// ------------------------------------------------------------------------------------------------
constexpr int COST_TX = 128;   // pixels per CTA

// WT = the part of a descriptor word the popcount looks at: the reference's 32-bit __popc on 64-bit words
// (hamming_distance.h:40-62) only ever sees the low half, so the compat mode stages 32-bit words.
template <int WORDS, bool POPC64>
__global__ void __launch_bounds__(256)
cost_u8_kernel(unsigned char* __restrict__ c8, const unsigned long long* __restrict__ cl,
               const unsigned long long* __restrict__ cr, int w, int h, int DP, int maxDisp) {
    using WT = typename std::conditional<POPC64, unsigned long long, unsigned>::type;
    extern __shared__ __align__(16) unsigned char cost_smem[];
    const int span = COST_TX + DP - 1;
    WT* s_r = reinterpret_cast<WT*>(cost_smem);                 // [WORDS][span] right descriptors x0-(DP-1) .. x0+TX-1
    WT* s_l = s_r + WORDS * span;                               // [WORDS][COST_TX] left descriptors
    unsigned* s_out = reinterpret_cast<unsigned*>(cost_smem + (size_t)WORDS * (span + COST_TX) * 8);   // [COST_TX][groups+1]
    const int x0 = blockIdx.x * COST_TX, y = blockIdx.y;
    const size_t rowoff = ((size_t)blockIdx.z * h + y) * (size_t)w;
    const int rx0 = x0 - (DP - 1);
    for (int i = threadIdx.x; i < span * WORDS; i += 256) {
        const int px = i / WORDS, k = i - px * WORDS;
        const int gx = rx0 + px;
        s_r[k * span + px] = (gx >= 0 && gx < w) ? (WT)cr[(rowoff + gx) * WORDS + k] : (WT)0;
    }
    for (int i = threadIdx.x; i < COST_TX * WORDS; i += 256) {
        const int px = i / WORDS, k = i - px * WORDS;
        s_l[k * COST_TX + px] = (x0 + px < w) ? (WT)cl[(rowoff + x0 + px) * WORDS + k] : (WT)0;
    }
    __syncthreads();
    // lanes run over pixels (consecutive descriptors in shared memory: conflict-free); a thread keeps ONE pixel
    // (its left descriptor stays in registers) and walks the disparity groups, packing 4 consecutive disparities
    // into one word of a padded output tile; the tile then leaves coalesced in 16-byte stores
    const int groups = DP >> 2;                           // 32-bit words per pixel (8, 16, 32 or 64)
    const int lg = 31 - __clz(groups);
    const int gpad = groups + 1;
    auto popc = [](WT v) -> unsigned { return POPC64 ? (unsigned)__popcll((unsigned long long)v) : (unsigned)__popc((unsigned)v); };
    // every pixel of the segment sees its whole disparity range: no per-disparity tests
    const bool full = x0 >= DP - 1 && maxDisp == DP;
    {
        const int px = threadIdx.x & (COST_TX - 1);
        WT p[WORDS];
#pragma unroll
        for (int k = 0; k < WORDS; ++k) p[k] = s_l[k * COST_TX + px];
        const int x = x0 + px;
        unsigned* orow = s_out + px * gpad;
        for (int g = threadIdx.x / COST_TX; g < groups; g += 256 / COST_TX) {
            const WT* q = s_r + px + (DP - 1) - g * 4;     // x-d in s_r for d = 4g; d+1 is one element lower
            unsigned hd[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hd[j] = 0;
#pragma unroll
                for (int k = 0; k < WORDS; ++k) hd[j] += popc(p[k] ^ q[k * span - j]);
                if (!full) {
                    const int d = g * 4 + j;
                    if (!(d < maxDisp && d <= x)) hd[j] = WORDS * 32;   // 0.5 * bits: no right pixel (the reference's 0.5)
                }
            }
            orow[g] = hd[0] | (hd[1] << 8) | (hd[2] << 16) | (hd[3] << 24);
        }
    }
    __syncthreads();
    uint4* out = reinterpret_cast<uint4*>(c8 + (rowoff + x0) * DP);   // DP bytes per pixel: 16-byte aligned
    const int npx = min(COST_TX, w - x0);
    for (int i = threadIdx.x; i < npx * (groups >> 2); i += 256) {   // consecutive threads -> consecutive 16 bytes
        const int px = i >> (lg - 2), g = (i & ((groups >> 2) - 1)) << 2;
        const unsigned* t = s_out + px * gpad + g;
        out[i] = make_uint4(t[0], t[1], t[2], t[3]);
    }
}

int launch_cost_u8(unsigned char* c8, const void* cl, const void* cr, int w, int h, int batch, int DP, int maxDisp,
                   int words, int popc_mode, cudaStream_t st) {
    dim3 grid(cdiv(w, COST_TX), h, batch), block(256);
    const auto* l = (const unsigned long long*)cl;
    const auto* r = (const unsigned long long*)cr;
    const bool p64 = popc_mode == ROO_POPC64;
    const size_t smem = (size_t)words * (COST_TX + DP - 1 + COST_TX) * 8 + (size_t)COST_TX * (DP / 4 + 1) * 4;
#define ROO_COST(W)                                                                                        \
    do {                                                                                                   \
        auto kern = p64 ? cost_u8_kernel<W, true> : cost_u8_kernel<W, false>;                              \
        if (smem > 48 * 1024) ROO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, block, smem, st>>>(c8, l, r, w, h, DP, maxDisp);                                      \
    } while (0)
    if (words == 1) ROO_COST(1);
    else if (words == 2) ROO_COST(2);
    else if (words == 4) ROO_COST(4);
    else return ROO_ERR_INVALID_ARGUMENT;
#undef ROO_COST
    count_launch();
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Engine: disparity straight from the descriptors, i.e. CensusStereoVolume(vol, self, other, maxDisp, sd)
// followed by CostVolMinimum<float,float> or CostVolMinimumSubpix(disp, vol, maxDisp, sd) without ever
// materialising vol.  Used for the right-reference disparity of the LR check (sd = +1,
// stereo2/main.cpp:385,432,435: vol[1] is never aggregated) and for the left one when no SGM path is
// enabled (sd = -1).  The raw costs are h/bits with out-of-image taps = 0.5.
// ------------------------------------------------------------------------------------------------
template <int WORDS, bool POPC64, bool IEEE>
__global__ void __launch_bounds__(128)
census_wta_kernel(float* __restrict__ disp, const unsigned long long* __restrict__ cself,
                  const unsigned long long* __restrict__ cother, int w, int h, int maxDisp, int subpix, int sdi) {
    // The 128 pixels of a block compare against 128 + maxDisp - 1 consecutive descriptors of the other image's
    // row: staged once in shared memory (taps outside the image are never read).  The search runs on integers:
    // cost = count / bits is monotonic in the count and the out-of-image value 0.5 is exactly bits/2 counts, so
    // "first strict minimum" = min over (count << 9 | d) -- five instructions per candidate.
    extern __shared__ unsigned long long s_o[];
    const int x0 = blockIdx.x * 128, x = x0 + threadIdx.x, y = blockIdx.y;
    const size_t rowoff = ((size_t)blockIdx.z * h + y) * (size_t)w;
    const unsigned long long* orow = cother + rowoff * WORDS;
    const int span = 128 + maxDisp - 1;
    const int base = sdi > 0 ? x0 : x0 - (maxDisp - 1);   // image x of s_o[0]
    for (int i = threadIdx.x; i < span * WORDS; i += 128) {
        const int xi = base + i / WORDS;
        s_o[i] = (xi >= 0 && xi < w) ? orow[(size_t)xi * WORDS + i % WORDS] : 0ull;
    }
    __syncthreads();
    if (x >= w) return;
    const unsigned long long* sp = cself + (rowoff + x) * WORDS;
    unsigned long long p[WORDS];
#pragma unroll
    for (int k = 0; k < WORDS; ++k) p[k] = sp[k];
    constexpr int HALF = WORDS * 32;                      // 0.5 * bits
    const float inv_bits = 1.0f / (float)(WORDS * 64);
    auto count = [&](int d) -> int {                      // Hamming count of candidate d; bits/2 outside the image
        const int xd = x + sdi * d;
        if (xd < 0 || xd >= w) return HALF;
        const unsigned long long* q = s_o + (size_t)(xd - base) * WORDS;
        unsigned hd = 0;
#pragma unroll
        for (int k = 0; k < WORDS; ++k) hd += hamming_word<POPC64>(p[k], q[k]);
        return (int)hd;
    };
    const int dvalid = sdi < 0 ? x + 1 : w - x;            // candidates d < dvalid have their tap inside the image
    // CostVolMinimum (cu_dense_stereo.cu:25-43) searches d < min(maxDisp, x+1), slices whose tap is outside hold 0.5;
    // CostVolMinimumSubpix (cu_dense_stereo.cu:66-109) searches the d < maxDisp whose tap is inside.
    const int md = subpix ? min(maxDisp, dvalid) : min(maxDisp, x + 1);
    const int din = min(md, dvalid);
    unsigned best = 0xffffffffu;
    const unsigned long long* q = s_o + (size_t)(x - base) * WORDS;
    for (int d = 0; d < din; ++d, q += sdi * WORDS) {
        unsigned hd = 0;
#pragma unroll
        for (int k = 0; k < WORDS; ++k) hd += hamming_word<POPC64>(p[k], q[k]);
        best = min(best, (hd << 9) | (unsigned)d);
    }
    if (din < md) best = min(best, ((unsigned)HALF << 9) | (unsigned)din);   // the first out-of-image slice (0.5)
    const int bestd = (int)(best & 511u);
    float out = (float)bestd;
    if (subpix) {
        const float bestc = (float)(best >> 9) * inv_bits;
        const int bestxr = x + sdi * bestd;
        if (0 < bestxr && bestxr < w - 1 && bestd + 1 < maxDisp) {
            const float sl = (float)count(max(bestd - 1, 0)) * inv_bits;  // GPU float->unsigned saturation of bestd-1 (Q7)
            const float sr = (float)count(bestd + 1) * inv_bits;
            const float sub = parabola_vertex<IEEE>((float)bestd, bestc, sl, sr);
            if ((float)(bestd - 1) < sub && sub < (float)(bestd + 1)) out = sub;
        }
    }
    disp[rowoff + x] = out;
}

int launch_census_wta(float* disp, const void* cself, const void* cother, int w, int h, int batch, int maxDisp,
                      int words, int popc_mode, int subpix, int sdi, int ieee_mode, cudaStream_t st) {
    dim3 grid(cdiv(w, 128), h, batch), block(128);
    const auto* a = (const unsigned long long*)cself;
    const auto* b = (const unsigned long long*)cother;
    const bool p64 = popc_mode == ROO_POPC64, ieee = ieee_mode != 0;
    const size_t smem = (size_t)(128 + maxDisp - 1) * words * 8;
#define ROO_RW(W, P, I) census_wta_kernel<W, P, I><<<grid, block, smem, st>>>(disp, a, b, w, h, maxDisp, subpix, sdi)
#define ROO_RW2(W)                                                                 \
    if (p64) { if (ieee) ROO_RW(W, true, true); else ROO_RW(W, true, false); }     \
    else { if (ieee) ROO_RW(W, false, true); else ROO_RW(W, false, false); }
    if (words == 1) { ROO_RW2(1) }
    else if (words == 2) { ROO_RW2(2) }
    else if (words == 4) { ROO_RW2(4) }
    else return ROO_ERR_INVALID_ARGUMENT;
#undef ROO_RW2
#undef ROO_RW
    count_launch();
    return launch_status();
}

}  // namespace roo_b200

using namespace roo_b200;

extern "C" int roo_census(const roo_image_t* census, const roo_image_t* img, int window, int in_type, void* stream) {
    if (window < 0 || window > 2 || (in_type != ROO_IMG_U8 && in_type != ROO_IMG_F32)) return ROO_ERR_INVALID_ARGUMENT;
    const size_t words = window == ROO_WIN_9x7 ? 1 : (window == ROO_WIN_11x11 ? 2 : 4);
    if (!valid_image(img, in_type == ROO_IMG_U8 ? 1 : 4) || !valid_image(census, words * 8)) return ROO_ERR_INVALID_ARGUMENT;
    if (census->w != img->w || census->h != img->h) return ROO_ERR_INVALID_ARGUMENT;
    return launch_census((char*)census->ptr, census->pitch, 0, (const char*)img->ptr, img->pitch, 0, (int)img->w,
                         (int)img->h, 1, window, in_type, as_stream(stream));
}

extern "C" int roo_census_stereo(const roo_image_t* disp, const roo_image_t* left, const roo_image_t* right, int maxDisp,
                                 void* stream) {
    if (!valid_image(disp, 1) || !valid_image(left, 8) || !valid_image(right, 8)) return ROO_ERR_INVALID_ARGUMENT;
    if (left->w != disp->w || left->h != disp->h || right->w != disp->w || right->h != disp->h)
        return ROO_ERR_INVALID_ARGUMENT;
    dim3 grid(cdiv((int)disp->w, 256), (unsigned)disp->h), block(256);
    census_stereo_kernel<<<grid, block, 0, as_stream(stream)>>>(Img<signed char>(*disp), Img<unsigned long long>(*left),
                                                                Img<unsigned long long>(*right), maxDisp);
    count_launch();
    return launch_status();
}

extern "C" int roo_census_stereo_volume(const roo_volume_t* vol, const roo_image_t* left, const roo_image_t* right,
                                        int words, int vol_type, int maxDisp, float sd, int popc_mode, void* stream) {
    if (words != 1 && words != 2 && words != 4) return ROO_ERR_INVALID_ARGUMENT;
    if (vol_type != ROO_VOL_F32 && vol_type != ROO_VOL_U16) return ROO_ERR_INVALID_ARGUMENT;
    if (sd != -1.0f && sd != 1.0f) return ROO_ERR_UNSUPPORTED;  // Q12: only integer unit steps are exact
    if (!valid_volume(vol, vol_type == ROO_VOL_F32 ? 4 : 2) || !valid_image(left, (size_t)words * 8) ||
        !valid_image(right, (size_t)words * 8))
        return ROO_ERR_INVALID_ARGUMENT;
    if (vol->w != left->w || vol->h != left->h || right->h != left->h)
        return ROO_ERR_INVALID_ARGUMENT;
    if (maxDisp <= 0) return ROO_OK;  // reference: empty loop
    if ((size_t)maxDisp > vol->d) return ROO_ERR_INVALID_ARGUMENT;
    const int sdi = sd < 0 ? -1 : 1;
    cudaStream_t st = as_stream(stream);
#define ROO_CSV(W)                                                                         \
    return vol_type == ROO_VOL_F32 ? csv_launch<W, float>(vol, left, right, maxDisp, sdi, popc_mode, st) \
                                   : csv_launch<W, unsigned short>(vol, left, right, maxDisp, sdi, popc_mode, st)
    if (words == 1) { ROO_CSV(1); }
    if (words == 2) { ROO_CSV(2); }
    ROO_CSV(4);
#undef ROO_CSV
}
