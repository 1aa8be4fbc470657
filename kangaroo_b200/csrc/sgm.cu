// Semi-global-matching aggregation for sm_100a.
//
// Replaces src/cu_semi_global_matching.cu:21-89 of the reference (one CTA per launch, one thread per
// scanline, previous row re-read from global memory, stride-`pitch` accesses on the horizontal paths).
//
// Design (DESIGN.md "SGM sweep"):
//  * internal aggregate H[pair][y][x][DP], fp32, disparity innermost: a pixel's disparities are one
//    contiguous 32*DPL*4-byte run, so EVERY path direction (vertical, horizontal, diagonal) moves
//    whole 128..1024-byte coalesced runs, and the horizontal paths stream memory linearly;
//  * one warp per scanline, lane l owns disparities [l*DPL, (l+1)*DPL); the previous pixel's row of
//    H never leaves registers; d-1 / d+1 neighbours cross lanes with two shuffles per pixel;
//  * min over disparities of the path cost = ONE redux.sync.min.f32 (CREDUX.MIN.F32, new on sm_100)
//    after a per-lane min -- not a 5-step shuffle tree;
//  * the loads of step r+PF (aggregate, cost, image pixel) are issued before step r is computed
//    (register ring), because the recurrence itself is a serial chain of `pathlen` steps;
//  * the last sweep can carry the winner-takes-all / parabola epilogue and skip writing H.
//
// Numerics: identical operation order to the reference kernel.  `lastBestCr + P2/(1+|dI|)` is one
// fma(rcp.approx, P2, lastBestCr) exactly like the SASS of the reference's -use_fast_math build, so the
// aggregate is bit-identical to the reference kernels'; with roo_set_ieee_division(1) it is an IEEE
// divide and add, bit-identical to the CPU oracle.
#include "common.cuh"
#include "kernels.cuh"

namespace roo_b200 {

constexpr int SWEEP_WARPS = 4;   // warps (= scanlines) per CTA
constexpr int SWEEP_PF = 4;      // prefetch distance in path steps
constexpr float SGM_MAX_ERROR = 1E30f;

template <int N> struct VecF;
template <> struct VecF<1> { using T = float; };
template <> struct VecF<2> { using T = float2; };
template <> struct VecF<4> { using T = float4; };

template <int DPL>
__device__ __forceinline__ void load_f(float (&v)[DPL], const float* p) {
    if constexpr (DPL == 8) {
        const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if constexpr (DPL == 4) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else if constexpr (DPL == 2) {
        const float2 a = *reinterpret_cast<const float2*>(p);
        v[0] = a.x; v[1] = a.y;
    } else {
        v[0] = *p;
    }
}
template <int DPL>
__device__ __forceinline__ void store_f(float* p, const float (&v)[DPL]) {
    if constexpr (DPL == 8) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else if constexpr (DPL == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (DPL == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        *p = v[0];
    }
}

// raw cost of one step as loaded (converted to float at use)
template <int DPL, int COST> struct RawCost;
template <int DPL> struct RawCost<DPL, COST_F32> {
    float v[DPL];
    __device__ __forceinline__ void load(const void* base, size_t idx) { load_f<DPL>(v, (const float*)base + idx); }
    __device__ __forceinline__ float get(int j, float) const { return v[j]; }
};
template <int DPL> struct RawCost<DPL, COST_U8> {
    unsigned w[(DPL + 3) / 4];
    __device__ __forceinline__ void load(const void* base, size_t idx) {
        const unsigned char* p = (const unsigned char*)base + idx;
        if constexpr (DPL == 8) { const uint2 t = *reinterpret_cast<const uint2*>(p); w[0] = t.x; w[1] = t.y; }
        else if constexpr (DPL == 4) w[0] = *reinterpret_cast<const unsigned*>(p);
        else if constexpr (DPL == 2) w[0] = *reinterpret_cast<const unsigned short*>(p);
        else w[0] = *p;
    }
    __device__ __forceinline__ float get(int j, float scale) const {
        return (float)((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) * scale;  // count * 1/bits: exact
    }
};

template <int DPL, int COST>
struct Stage {
    float hin[DPL];
    RawCost<DPL, COST> c;
    float pix;
};

struct Scanline { int x0, y0, len; };

__device__ __forceinline__ Scanline scanline_of(int s, int w, int h, int dx, int dy) {
    Scanline sl;
    if (dx == 0) { sl.x0 = s; sl.y0 = dy > 0 ? 0 : h - 1; sl.len = h; }
    else if (dy == 0) { sl.y0 = s; sl.x0 = dx > 0 ? 0 : w - 1; sl.len = w; }
    else {
        // a scanline starts at every pixel of the entry row, then of the entry column (same order as the oracle)
        if (s < w) { sl.x0 = s; sl.y0 = dy > 0 ? 0 : h - 1; }
        else { sl.x0 = dx > 0 ? 0 : w - 1; sl.y0 = dy > 0 ? (s - w + 1) : (h - 1 - (s - w + 1)); }
        const int lenx = dx > 0 ? (w - sl.x0) : (sl.x0 + 1);
        const int leny = dy > 0 ? (h - sl.y0) : (sl.y0 + 1);
        sl.len = min(lenx, leny);
    }
    return sl;
}

template <int DPL, int COST, int EPI>
__global__ void __launch_bounds__(SWEEP_WARPS * 32)
sgm_sweep_kernel(const SweepArgs a, const int n_scan, const int ieee) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * SWEEP_WARPS + (threadIdx.x >> 5);
    if (s >= n_scan) return;
    const int pair = blockIdx.y;
    const Scanline sl = scanline_of(s, a.w, a.h, a.dx, a.dy);

    float* __restrict__ H = a.H + (size_t)pair * a.h_pair;
    const void* Cbase = COST == COST_F32 ? (const void*)((const float*)a.C + (size_t)pair * a.c_pair)
                                         : (const void*)((const unsigned char*)a.C + (size_t)pair * a.c_pair);
    const char* __restrict__ img = a.img + (size_t)pair * a.img_pair;
    const bool first = a.first != 0;
    const int d0 = lane * DPL;
    const int DP = a.DP;

    auto elem_index = [&](int r) -> size_t {
        const int x = sl.x0 + r * a.dx, y = sl.y0 + r * a.dy;
        return ((size_t)y * a.w + x) * DP + d0;
    };
    auto load_stage = [&](Stage<DPL, COST>& st, int r) {
        const int x = sl.x0 + r * a.dx, y = sl.y0 + r * a.dy;
        const size_t idx = ((size_t)y * a.w + x) * DP + d0;
        if (!first) load_f<DPL>(st.hin, H + idx);
        st.c.load(Cbase, idx);
        const char* prow = img + (size_t)y * a.img_pitch;
        st.pix = a.img_type == ROO_IMG_U8 ? (float)((const unsigned char*)prow)[x] * a.img_scale
                                          : ((const float*)prow)[x];
    };

    Stage<DPL, COST> ring[SWEEP_PF];
#pragma unroll
    for (int k = 0; k < SWEEP_PF; ++k)
        if (k < sl.len) load_stage(ring[k], k);

    float hp[DPL];          // previous pixel's H row, +inf where d >= its disparity range
    float lastBest = 0.0f;  // reference: lastBestCr starts at 0, NOT at the first pixel's minimum
    float last_c = 0.0f;
    const float INF = __int_as_float(0x7f800000);

    for (int r0 = 0; r0 < sl.len; r0 += SWEEP_PF) {
#pragma unroll
        for (int k = 0; k < SWEEP_PF; ++k) {
            const int r = r0 + k;
            if (r >= sl.len) break;
            Stage<DPL, COST> cur = ring[k];
            if (r + SWEEP_PF < sl.len) load_stage(ring[k], r + SWEEP_PF);

            const int x = sl.x0 + r * a.dx, y = sl.y0 + r * a.dy;
            const int maxDisp = min(a.maxDisp, x + 1);
            float hnew[DPL];
            float best = SGM_MAX_ERROR;
            if (r == 0) {
                // start pixel: volH += volC (cu_semi_global_matching.cu:31-35)
#pragma unroll
                for (int j = 0; j < DPL; ++j) {
                    const float hin = first ? 0.0f : cur.hin[j];
                    const bool in = d0 + j < maxDisp;
                    hnew[j] = in ? hin + cur.c.get(j, a.cost_scale) : hin;
                    hp[j] = in ? hnew[j] : INF;
                }
            } else {
                const float diff = last_c - cur.pix;
                const float denom = 1.0f + fabsf(diff);
                float up = __shfl_up_sync(0xffffffffu, hp[DPL - 1], 1);
                float dn = __shfl_down_sync(0xffffffffu, hp[0], 1);
                if (lane == 0) up = INF;    // d-1 < 0
                if (lane == 31) dn = INF;   // d+1 beyond the padded range
                const float base = ieee ? sgm_p2_base<true>(lastBest, a.P2, denom) : sgm_p2_base<false>(lastBest, a.P2, denom);
#pragma unroll
                for (int j = 0; j < DPL; ++j) {
                    const float hm = j > 0 ? hp[j - 1] : up;
                    const float hq = j < DPL - 1 ? hp[j + 1] : dn;
                    float CM = fminf(base, hp[j]);
                    CM = fminf(CM, hm + a.P1);
                    CM = fminf(CM, hq + a.P1);
                    const float Cr = (CM + cur.c.get(j, a.cost_scale)) - lastBest;
                    const float hin = first ? 0.0f : cur.hin[j];
                    const bool in = d0 + j < maxDisp;
                    if (in) best = fminf(best, Cr);
                    hnew[j] = in ? hin + Cr : hin;
                }
#pragma unroll
                for (int j = 0; j < DPL; ++j) hp[j] = (d0 + j < maxDisp) ? hnew[j] : INF;
                lastBest = warp_min_f32(best);
            }
            last_c = cur.pix;

            const size_t idx = ((size_t)y * a.w + x) * DP + d0;
            if (EPI != EPI_WTA_ONLY) store_f<DPL>(H + idx, hnew);

            if (EPI != EPI_NONE) {
                // winner-takes-all over d < min(maxDisp, x+1): first (lowest-d) minimum
                float lc = INF;
                int ld = 0;
#pragma unroll
                for (int j = 0; j < DPL; ++j) {
                    const float v = hp[j];  // == hnew in range, +inf outside
                    if (v < lc) { lc = v; ld = d0 + j; }
                }
                const float m = warp_min_f32(lc);
                const unsigned ball = __ballot_sync(0xffffffffu, lc == m);
                const int win = __ffs(ball) - 1;
                int bestd = __shfl_sync(0xffffffffu, ld, win);
                float bestc = m;
                float out;
                if (!a.subpix) {
                    out = (float)bestd;  // CostVolMinimum<float,float> (cu_dense_stereo.cu:25-43)
                } else {
                    // CostVolMinimumSubpix, sd = -1 (cu_dense_stereo.cu:66-109): bestc starts at 1e10
                    if (!(bestc < 1E10f)) { bestc = 1E10f; bestd = 0; }
                    out = (float)bestd;
                    const int bestxr = x - bestd;
                    if (0 < bestxr && bestxr < a.w - 1 && bestd + 1 < a.maxDisp) {
                        const int dl = max(bestd - 1, 0);  // float->unsigned saturation in the reference (Q7)
                        const int dr = bestd + 1;
                        float slc = 0.0f, src = 0.0f;
#pragma unroll
                        for (int j = 0; j < DPL; ++j) {
                            if (d0 + j == dl) slc = hnew[j];
                            if (d0 + j == dr) src = hnew[j];
                        }
                        const float sl_ = __shfl_sync(0xffffffffu, slc, dl / DPL);
                        const float sr_ = __shfl_sync(0xffffffffu, src, dr / DPL);
                        const float sub = ieee ? parabola_vertex<true>((float)bestd, bestc, sl_, sr_)
                                               : parabola_vertex<false>((float)bestd, bestc, sl_, sr_);
                        if ((float)(bestd - 1) < sub && sub < (float)(bestd + 1)) out = sub;
                    }
                }
                if (lane == 0) a.disp[(size_t)pair * a.disp_pair + (size_t)y * a.w + x] = out;
            }
        }
    }
}

template <int DPL, int COST>
static void sweep_launch_epi(const SweepArgs& a, int n_scan, dim3 grid, cudaStream_t st) {
    const int ieee = g_ieee_div.load();
    if (a.epi == EPI_NONE) sgm_sweep_kernel<DPL, COST, EPI_NONE><<<grid, SWEEP_WARPS * 32, 0, st>>>(a, n_scan, ieee);
    else if (a.epi == EPI_WTA_WRITE) sgm_sweep_kernel<DPL, COST, EPI_WTA_WRITE><<<grid, SWEEP_WARPS * 32, 0, st>>>(a, n_scan, ieee);
    else sgm_sweep_kernel<DPL, COST, EPI_WTA_ONLY><<<grid, SWEEP_WARPS * 32, 0, st>>>(a, n_scan, ieee);
}

template <int DPL>
static void sweep_launch_cost(const SweepArgs& a, int n_scan, dim3 grid, cudaStream_t st) {
    if (a.cost_kind == COST_F32) sweep_launch_epi<DPL, COST_F32>(a, n_scan, grid, st);
    else sweep_launch_epi<DPL, COST_U8>(a, n_scan, grid, st);
}

int launch_sweep(const SweepArgs& a, cudaStream_t st) {
    const int n_scan = a.dx == 0 ? a.w : (a.dy == 0 ? a.h : a.w + a.h - 1);
    dim3 grid(cdiv(n_scan, SWEEP_WARPS), a.batch);
    switch (a.DP) {
        case 32: sweep_launch_cost<1>(a, n_scan, grid, st); break;
        case 64: sweep_launch_cost<2>(a, n_scan, grid, st); break;
        case 128: sweep_launch_cost<4>(a, n_scan, grid, st); break;
        case 256: sweep_launch_cost<8>(a, n_scan, grid, st); break;
        default: return ROO_ERR_UNSUPPORTED;
    }
    count_launch();
    return launch_status();
}

int sgm_directions(int dohoriz, int dovert, int doreverse, int dodiag, int dxs[8], int dys[8]) {
    int n = 0;
    auto add = [&](int dx, int dy) { dxs[n] = dx; dys[n] = dy; ++n; };
    if (dovert) add(0, 1);                                   // cu_semi_global_matching.cu:72
    if (dodiag) { add(1, 1); add(-1, 1); }
    if (dovert && doreverse) add(0, -1);                     // :74
    if (dodiag && doreverse) { add(-1, -1); add(1, -1); }
    if (dohoriz) {
        add(1, 0);                                           // :81
        if (doreverse) add(-1, 0);                           // :83
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// Layout adapters for the granular roo_sgm(): roo::Volume (d outermost, x fastest) <-> internal
// (d innermost).  32(x) x 32(d) tiles through shared memory; both sides coalesced.
// ------------------------------------------------------------------------------------------------
template <typename Tsrc>
__device__ __forceinline__ float cost_as_float(const Tsrc& v, int ieee);
template <> __device__ __forceinline__ float cost_as_float<float>(const float& v, int) { return v; }
// CostVolElem::operator float (CostVolElem.h:12-15): n > 0 ? sum / n : 1e30
template <> __device__ __forceinline__ float cost_as_float<roo_costvolelem_t>(const roo_costvolelem_t& e, int ieee) {
    if (e.n <= 0) return 1E30f;
    return ieee ? ref_div<true>(e.sum, (float)e.n) : ref_div<false>(e.sum, (float)e.n);
}

template <typename Tsrc>
__global__ void __launch_bounds__(256)
vol_to_internal_kernel(float* __restrict__ dst, Vol<Tsrc> src, int DP, int maxDisp, int ieee) {
    __shared__ float tile[32][33];
    const int x0 = blockIdx.x * 32, dd0 = blockIdx.y * 32, y = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int d = dd0 + ty + 8 * k, x = x0 + tx;
        float v = 0.0f;
        if (d < maxDisp && x < src.w) v = cost_as_float<Tsrc>(src(x, y, d), ieee);
        tile[ty + 8 * k][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x0 + ty + 8 * k, d = dd0 + tx;
        if (x < src.w) dst[((size_t)y * src.w + x) * DP + d] = tile[tx][ty + 8 * k];
    }
}

__global__ void __launch_bounds__(256)
internal_to_vol_kernel(Vol<float> dst, const float* __restrict__ src, int DP, int maxDisp) {
    __shared__ float tile[32][33];
    const int x0 = blockIdx.x * 32, dd0 = blockIdx.y * 32, y = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x0 + ty + 8 * k, d = dd0 + tx;
        tile[ty + 8 * k][tx] = x < dst.w ? src[((size_t)y * dst.w + x) * DP + d] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int d = dd0 + ty + 8 * k, x = x0 + tx;
        if (d < maxDisp && x < dst.w) dst(x, y, d) = tile[tx][ty + 8 * k];
    }
}

int launch_internal_to_vol(const roo_volume_t* dst, const float* src, int DP, int maxDisp, cudaStream_t st) {
    dim3 grid(cdiv((int)dst->w, 32), DP / 32, (unsigned)dst->h);
    internal_to_vol_kernel<<<grid, 256, 0, st>>>(Vol<float>(*dst), src, DP, maxDisp);
    count_launch();
    return launch_status();
}

}  // namespace roo_b200

using namespace roo_b200;

extern "C" int roo_sgm(const roo_volume_t* volH, const roo_volume_t* volC, int volc_type, const roo_image_t* left,
                       int img_type, int maxDisp, float P1, float P2, int dohoriz, int dovert, int doreverse,
                       int dodiag, void* stream) {
    if (volc_type != ROO_VOL_F32 && volc_type != ROO_VOL_ELEM) return ROO_ERR_INVALID_ARGUMENT;
    if (img_type != ROO_IMG_U8 && img_type != ROO_IMG_F32) return ROO_ERR_INVALID_ARGUMENT;
    if (!valid_volume(volH, 4) || !valid_volume(volC, volc_type == ROO_VOL_F32 ? 4 : 8) ||
        !valid_image(left, img_type == ROO_IMG_U8 ? 1 : 4))
        return ROO_ERR_INVALID_ARGUMENT;
    if (volH->w != volC->w || volH->h != volC->h || left->w != volC->w || left->h != volC->h)
        return ROO_ERR_INVALID_ARGUMENT;
    if (maxDisp > 256) return ROO_ERR_UNSUPPORTED;
    if ((size_t)maxDisp > volH->d || (size_t)maxDisp > volC->d) return ROO_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    // volH.Memset(0) (cu_semi_global_matching.cu:68; Volume.h:78-81 clears pitch*h*d bytes)
    if (volH->img_pitch == volH->pitch * volH->h) {
        ROO_CUDA_TRY(cudaMemsetAsync(volH->ptr, 0, volH->pitch * volH->h * volH->d, st));
    } else {
        ROO_CUDA_TRY(cudaMemset2DAsync(volH->ptr, volH->img_pitch, 0, volH->pitch * volH->h, volH->d, st));
    }
    int dxs[8], dys[8];
    const int ndir = sgm_directions(dohoriz, dovert, doreverse, dodiag, dxs, dys);
    if (ndir == 0 || maxDisp <= 0) return ROO_OK;

    const int w = (int)volC->w, h = (int)volC->h, DP = disp_padded(maxDisp);
    const size_t n = (size_t)w * h * DP;
    float* scratch = nullptr;  // [Ci | Hi], stream-ordered pool memory
    ROO_CUDA_TRY(cudaMallocAsync((void**)&scratch, 2 * n * sizeof(float), st));
    float* Ci = scratch;
    float* Hi = scratch + n;
    const int ieee = g_ieee_div.load();
    dim3 tgrid(cdiv(w, 32), DP / 32, h);
    if (volc_type == ROO_VOL_F32)
        vol_to_internal_kernel<float><<<tgrid, 256, 0, st>>>(Ci, Vol<float>(*volC), DP, maxDisp, ieee);
    else
        vol_to_internal_kernel<roo_costvolelem_t><<<tgrid, 256, 0, st>>>(Ci, Vol<roo_costvolelem_t>(*volC), DP, maxDisp, ieee);
    count_launch();
    int rc = launch_status();

    SweepArgs a{};
    a.H = Hi; a.h_pair = n; a.C = Ci; a.c_pair = n;
    a.img = (const char*)left->ptr; a.img_pitch = left->pitch; a.img_pair = 0; a.img_type = img_type;
    a.img_scale = 1.0f; a.cost_scale = 1.0f;
    a.w = w; a.h = h; a.DP = DP; a.maxDisp = maxDisp; a.batch = 1;
    a.P1 = P1; a.P2 = P2; a.cost_kind = COST_F32; a.epi = EPI_NONE; a.subpix = 0; a.disp = nullptr; a.disp_pair = 0;
    for (int i = 0; i < ndir && rc == 0; ++i) {
        a.dx = dxs[i]; a.dy = dys[i]; a.first = i == 0;
        rc = launch_sweep(a, st);
    }
    if (rc == 0) rc = launch_internal_to_vol(volH, Hi, DP, maxDisp, st);
    cudaError_t fe = cudaFreeAsync(scratch, st);
    return rc != 0 ? rc : (int)fe;
}
