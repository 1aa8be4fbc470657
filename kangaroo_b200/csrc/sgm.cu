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
#include "sgm_step.cuh"

#include <type_traits>

namespace roo_b200 {

constexpr int SWEEP_WARPS = 4;   // warps (= scanlines) per CTA

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

// prefetch depth in path steps (stages of one pixel each, staged global -> shared by cp.async): under load a
// DRAM access takes ~3000 SM cycles on B200, so each SM needs 60-80 KB in flight to stream at HBM speed
__host__ __device__ constexpr int sweep_pfs(int DPL, int CE) { return DPL >= 8 ? 4 : (DPL == 4 && CE == 4 ? 4 : 8); }
// one prefetched path step: [aggregate row][cost row, or with in-sweep cost the 32*DPL right-image census words R(x-d)]
// [intensity, left census word]
template <int DPL, int COST> __host__ __device__ constexpr int sweep_cost_bytes() {
    return COST == COST_CEN32 ? 32 * DPL * 4 : 32 * DPL * RawCost<DPL, COST>::ELEM;
}
template <int DPL, int COST> __host__ __device__ constexpr int sweep_stage_bytes() { return 32 * DPL * 4 + sweep_cost_bytes<DPL, COST>() + 16; }

template <int DPL, int COST, int EPI, bool FIRST, bool IEEE>
__global__ void __launch_bounds__(SWEEP_WARPS * 32)
sgm_sweep_kernel(const SweepArgs a, const int n_scan) {
    constexpr int DP = 32 * DPL;
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    constexpr bool CEN = COST == COST_CEN32;                      // cost recomputed from census words (no cost volume)
    constexpr int PFS = sweep_pfs(DPL, CEN ? 4 : CE);
    constexpr int STAGE_B = sweep_stage_bytes<DPL, COST>();
    constexpr int CB = sweep_cost_bytes<DPL, COST>();
    extern __shared__ __align__(16) unsigned char sweep_smem[];   // [SWEEP_WARPS][PFS][STAGE_B]

    const int lane = threadIdx.x & 31;
    // warp index through a shuffle: provably warp-uniform, so the branches below are uniform branches
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int s = blockIdx.x * SWEEP_WARPS + warp;
    if (s >= n_scan) return;
    const int pair = blockIdx.y;
    const int w = a.w, dx = a.dx, M = a.maxDisp, subpix = a.subpix;
    const float P1 = a.P1, P2 = a.P2, cscale = a.cost_scale;
    const Scanline sl = scanline_of(s, w, a.h, dx, a.dy);
    const int len = sl.len;
    const int d0 = lane * DPL;

    // element (x0,y0,d0) and the per-step strides; every access below is pointer + running offset
    const size_t e0 = ((size_t)sl.y0 * w + sl.x0) * DP + d0;
    const ptrdiff_t pstep = (ptrdiff_t)a.dy * w + dx;      // pixels per path step
    const ptrdiff_t estep = pstep * DP;                    // elements per path step
    float* hst = a.H + (size_t)pair * a.h_pair + e0;                                  // store cursor
    const float* hld = hst;                                                           // prefetch cursors
    const char* cld = (const char*)a.C + ((size_t)pair * a.c_pair + e0) * CE;
    const float* ild = a.img + (size_t)pair * a.img_pair + (size_t)sl.y0 * w + sl.x0;
    // in-sweep cost: this lane's right-image descriptors R(x - d0 - j) (u64 each, the low word is used; the arrays are
    // padded, positions left of the image give masked garbage) and the left descriptor of the pixel
    const unsigned long long* rld = CEN ? a.cenR + (size_t)pair * a.cen_pair + (size_t)sl.y0 * w + sl.x0 - d0 : nullptr;
    const unsigned long long* lld = CEN ? a.cenL + (size_t)pair * a.cen_pair + (size_t)sl.y0 * w + sl.x0 : nullptr;
    float* dst = (EPI != EPI_NONE) ? a.disp + (size_t)pair * a.disp_pair + (size_t)sl.y0 * w + sl.x0 : nullptr;

    // Prefetch: step r+PFS-1 is copied global -> shared (asynchronously, no registers) while step r is computed.
    // Every lane copies and later reads its own bytes; only the intensity (lane 0) needs the __syncwarp.
    const unsigned pfBase = (unsigned)__cvta_generic_to_shared(sweep_smem) + warp * PFS * STAGE_B;
    auto issue_step = [&](int rl) {
        if (rl < len) {
            const unsigned sd = pfBase + ((unsigned)rl & (PFS - 1)) * STAGE_B;
            if (!FIRST) cp_async_bytes<DPL * 4>(sd + lane * DPL * 4, hld);
            if (CEN) {
#pragma unroll
                for (int j = 0; j < DPL; ++j) cp_async_bytes<4>(sd + DP * 4 + (lane * DPL + j) * 4, reinterpret_cast<const unsigned*>(rld - j));
                if (lane == 1) cp_async_bytes<4>(sd + DP * 4 + CB + 4, reinterpret_cast<const unsigned*>(lld));
                rld += pstep; lld += pstep;
            } else {
                cp_async_bytes<DPL * (CE > 0 ? CE : 1)>(sd + DP * 4 + lane * DPL * CE, cld);
            }
            if (lane == 0) cp_async_bytes<4>(sd + DP * 4 + CB, ild);
            hld += estep; cld += estep * CE; ild += pstep;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int k = 0; k < PFS - 1; ++k) issue_step(k);

    float hp[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) hp[j] = ROO_INF;   // path start: no previous pixel
    float lastBest = 0.0f, last_c = 0.0f;
    int x = sl.x0;
    // ---- row-strip split: continue a path that comes in through the strip's entry row
    constexpr int REC = DP + 4;
    const int y_entry = a.dy > 0 ? 0 : a.h - 1, y_exit = a.dy > 0 ? a.h - 1 : 0;
    bool continued = false;
    if (a.strip_import != nullptr && a.dy != 0 && sl.y0 == y_entry) {
        const int xin = sl.x0 - dx;                  // the pixel of the upstream strip's exit row this path comes from
        if (xin >= 0 && xin < w) {
            const float* rec = a.strip_import + ((size_t)pair * w + xin) * REC;
            if (lane == 0) {
                int seq;
                do { asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seq) : "l"(rec + DP + 3) : "memory"); } while (seq != a.strip_seq);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < DPL; ++j) hp[j] = __ldcg(rec + d0 + j);
            lastBest = __ldcg(rec + DP);
            last_c = __ldcg(rec + DP + 1);
            continued = true;
        }
    }
    // all lanes in range iff x >= xf (possible only when maxDisp fills the padded range)
    const int xf = (M == DP) ? DP - 1 : 0x3fffffff;

    auto step = [&](auto masked_tag, int r) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        const unsigned stg = pfBase + ((unsigned)r & (PFS - 1)) * STAGE_B;
        float hin[DPL], hnew[DPL], best, pix;
        if (!FIRST) lds_vec<DPL>(hin, stg + lane * DPL * 4);
        float craw[DPL];
        if (CEN) {
            unsigned rw[DPL], lw;
#pragma unroll
            for (int j = 0; j < DPL; ++j) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rw[j]) : "r"(stg + DP * 4 + (lane * DPL + j) * 4));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lw) : "r"(stg + DP * 4 + CB + 4));
#pragma unroll
            for (int j = 0; j < DPL; ++j) craw[j] = (float)__popc(lw ^ rw[j]);
        } else {
            RawCost<DPL, COST> rc;
            rc.lds(stg + DP * 4 + lane * DPL * CE);
#pragma unroll
            for (int j = 0; j < DPL; ++j) craw[j] = rc.raw(j);
        }
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pix) : "r"(stg + DP * 4 + CB));
        // start pixel: `volH += volC`, lastBestCr = 0 (cu_semi_global_matching.cu:31-35) == a step with P2 = 0
        const float p2 = (r == 0 && !continued) ? 0.0f : P2;
        const float denom = 1.0f + fabsf(last_c - pix);
        const int lim = MASKED ? min(M, x + 1) - d0 : 0;
        sgm_step<DPL, MASKED, FIRST, IEEE>(hp, lastBest, denom, P1, p2, craw, cscale, hin, lim, lane, hnew, hp, best);
        lastBest = (r == 0 && !continued) ? 0.0f : best;
        last_c = pix;
        if (EPI != EPI_WTA_ONLY) store_f<DPL>(hst, hnew);
        hst += estep;
        if (EPI != EPI_NONE) {
            const float out = wta_epilogue<DPL, IEEE>(hp, lane, x, w, M, subpix);
            if (lane == 0) *dst = out;
            dst += pstep;
        }
        x += dx;
    };

#pragma unroll 1
    for (int r = 0; r < len; ++r) {
        issue_step(r + PFS - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(PFS - 1) : "memory");   // step r's stage has landed
        __syncwarp();
        if (x >= xf) step(std::false_type{}, r);
        else step(std::true_type{}, r);
    }
    // ---- row-strip split: hand the state of a path that leaves through the strip's exit row to the downstream strip
    if (a.strip_export != nullptr && a.dy != 0 && sl.y0 + a.dy * (len - 1) == y_exit) {
        const int xout = x - dx;                     // x of the last pixel
        float* rec = a.strip_export + ((size_t)pair * w + xout) * REC;
#pragma unroll
        for (int j = 0; j < DPL; ++j) rec[d0 + j] = hp[j];
        if (lane == 0) { rec[DP] = lastBest; rec[DP + 1] = last_c; }
        __syncwarp();
        if (lane == 0) {
            __threadfence_system();
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(rec + DP + 3), "r"(a.strip_seq) : "memory");
        }
    }
}

template <int DPL, int COST, int EPI>
static void sweep_launch3(const SweepArgs& a, int n_scan, dim3 grid, cudaStream_t st) {
    const bool ieee = a.ieee != 0;
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    size_t smem = (size_t)SWEEP_WARPS * sweep_pfs(DPL, COST == COST_CEN32 ? 4 : CE) * sweep_stage_bytes<DPL, COST>();
    // Row-strip split: all scanlines of a sweep advance in lock step when every CTA is resident, so a strip would export
    // its states only at the very end and the strips of one sweep would run one after the other.  Asking for more shared
    // memory than the kernel needs caps the CTAs per SM: the grid then runs in waves (lowest scanlines first on every
    // strip), the first wave's states are exported early and the downstream strip starts while this one is still working.
    const int cap = g_strip_ctas_per_sm.load(std::memory_order_relaxed);
    if (cap > 0 && (a.strip_import != nullptr || a.strip_export != nullptr)) {
        const size_t per_cta = (size_t)(227 * 1024) / cap - 1024;   // 1 KB per CTA is reserved by the runtime
        if (per_cta > smem) smem = per_cta;
    }
#define ROO_SWEEP(F, I)                                                                                  \
    do {                                                                                                 \
        auto kern = sgm_sweep_kernel<DPL, COST, EPI, F, I>;                                              \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        kern<<<grid, SWEEP_WARPS * 32, smem, st>>>(a, n_scan);                                           \
    } while (0)
    if (a.first) { if (ieee) ROO_SWEEP(true, true); else ROO_SWEEP(true, false); }
    else { if (ieee) ROO_SWEEP(false, true); else ROO_SWEEP(false, false); }
#undef ROO_SWEEP
}

template <int DPL, int COST>
static void sweep_launch_epi(const SweepArgs& a, int n_scan, dim3 grid, cudaStream_t st) {
    if (a.epi == EPI_NONE) sweep_launch3<DPL, COST, EPI_NONE>(a, n_scan, grid, st);
    else if (a.epi == EPI_WTA_WRITE) sweep_launch3<DPL, COST, EPI_WTA_WRITE>(a, n_scan, grid, st);
    else sweep_launch3<DPL, COST, EPI_WTA_ONLY>(a, n_scan, grid, st);
}

template <int DPL>
static void sweep_launch_cost(const SweepArgs& a, int n_scan, dim3 grid, cudaStream_t st) {
    if (a.cost_kind == COST_F32) sweep_launch_epi<DPL, COST_F32>(a, n_scan, grid, st);
    else if (a.cost_kind == COST_CEN32) sweep_launch_epi<DPL, COST_CEN32>(a, n_scan, grid, st);
    else sweep_launch_epi<DPL, COST_U8>(a, n_scan, grid, st);
}

std::atomic<int> g_use_hsweep{1};
// measured on 8 B200, one 3840x2160x256 pair (ms at 2 / 4 / 8 strips): no cap 18.1 / 13.2 / 10.7, 2 CTAs per SM 17.8 / 12.8 / 10.2,
// 3: 17.0 / 12.0 / 9.7, 5: 17.6 / 12.4 / 10.2
std::atomic<int> g_strip_ctas_per_sm{3};

int launch_sweep(const SweepArgs& a, cudaStream_t st) {
    if (a.dy == 0 && g_use_hsweep.load(std::memory_order_relaxed)) return launch_hsweep(a, st);
    const int n_scan = a.dx == 0 ? a.w : (a.dy == 0 ? a.h : a.w + a.h - 1);
    dim3 grid(cdiv(n_scan, SWEEP_WARPS), a.batch);
    switch (a.DP) {
        case 32: sweep_launch_cost<1>(a, n_scan, grid, st); break;
        case 64: sweep_launch_cost<2>(a, n_scan, grid, st); break;
        case 128: sweep_launch_cost<4>(a, n_scan, grid, st); break;
        case 256: sweep_launch_cost<8>(a, n_scan, grid, st); break;
        case 512: sweep_launch_cost<16>(a, n_scan, grid, st); break;
        default: return ROO_ERR_UNSUPPORTED;
    }
    count_launch();
    return launch_status();
}

// Adaptive-P2 intensity image as tightly packed fp32 [pair][y][x]: u8 * scale (stereo2/main.cpp:376 uses
// 1/255; scale 1 reproduces the uchar instantiation's integer difference exactly) or a pitched fp32 copy.
template <typename Tin>
__global__ void __launch_bounds__(256)
image_to_f32_kernel(float* __restrict__ dst, const char* __restrict__ src, size_t pitch, size_t src_pair, int w, int h,
                    float scale) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const Tin v = reinterpret_cast<const Tin*>(src + (size_t)blockIdx.z * src_pair + (size_t)y * pitch)[x];
    dst[((size_t)blockIdx.z * h + y) * w + x] = sizeof(Tin) == 1 ? (float)v * scale : (float)v;
}

int launch_image_to_f32(float* dst, const void* src, size_t pitch, size_t src_pair, int img_type, int w, int h,
                        int batch, float scale, cudaStream_t st) {
    dim3 grid(cdiv(w, 256), h, batch);
    if (img_type == ROO_IMG_U8)
        image_to_f32_kernel<unsigned char><<<grid, 256, 0, st>>>(dst, (const char*)src, pitch, src_pair, w, h, scale);
    else
        image_to_f32_kernel<float><<<grid, 256, 0, st>>>(dst, (const char*)src, pitch, src_pair, w, h, scale);
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

SgmPlan sgm_plan(int dohoriz, int dovert, int doreverse, int dodiag, int fuse) {
    SgmPlan pl{};
    auto add = [&](int fused, int dx, int dy) { pl.pass[pl.n++] = SgmPass{fused, dx, dy}; };
    // a fused group needs the vertical path of that travel direction, and must not be the last pass
    // (the winner-takes-all epilogue rides on a single-path sweep)
    const bool f = fuse && dodiag && dovert && dohoriz;
    if (f) add(1, 0, 1);
    else {
        if (dovert) add(0, 0, 1);                                // cu_semi_global_matching.cu:72
        if (dodiag) { add(0, 1, 1); add(0, -1, 1); }
    }
    if (doreverse) {
        if (f) add(1, 0, -1);
        else {
            if (dovert) add(0, 0, -1);                           // :74
            if (dodiag) { add(0, -1, -1); add(0, 1, -1); }
        }
    }
    if (dohoriz) {
        add(0, 1, 0);                                            // :81
        if (doreverse) add(0, -1, 0);                            // :83
    }
    return pl;
}

int launch_pass(SweepArgs a, const SgmPass& pass, float* edge, int* progress, cudaStream_t st) {
    if (pass.fused) return launch_vgroup(a, pass.dy > 0 ? 1 : 0, edge, progress, st);
    a.dx = pass.dx; a.dy = pass.dy;
    return launch_sweep(a, st);
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
    if (maxDisp > ROO_MAX_DISP) return ROO_ERR_UNSUPPORTED;
    if ((size_t)maxDisp > volH->d || (size_t)maxDisp > volC->d) return ROO_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    // volH.Memset(0) (cu_semi_global_matching.cu:68; Volume.h:78-81 clears pitch*h*d bytes)
    if (volH->img_pitch == volH->pitch * volH->h) {
        ROO_CUDA_TRY(cudaMemsetAsync(volH->ptr, 0, volH->pitch * volH->h * volH->d, st));
    } else {
        ROO_CUDA_TRY(cudaMemset2DAsync(volH->ptr, volH->img_pitch, 0, volH->pitch * volH->h, volH->d, st));
    }
    const SgmPlan plan = sgm_plan(dohoriz, dovert, doreverse, dodiag, maxDisp <= ROO_MAX_DISP_FUSED ? 1 : 0);
    if (plan.n == 0 || maxDisp <= 0) return ROO_OK;

    const int w = (int)volC->w, h = (int)volC->h, DP = disp_padded(maxDisp);
    const size_t n = (size_t)w * h * DP;
    bool fused = false;
    for (int i = 0; i < plan.n; ++i) fused |= plan.pass[i].fused != 0;
    const size_t edge_n = fused ? vgroup_edge_floats(w, h, DP) : 0;
    const size_t flag_n = fused ? (size_t)vgroup_bands(w, h, DP) + 2 : 0;   // + the CTA ticket counter
    // [Ci | Hi | fp32 image | edge rows | flags], stream-ordered pool memory; every segment 256-byte aligned
    auto al = [](size_t nfloats) { return (nfloats + 63) / 64 * 64; };
    const size_t img_n = al((size_t)w * h);
    float* scratch = nullptr;
    ROO_CUDA_TRY(cudaMallocAsync((void**)&scratch, (2 * al(n) + img_n + al(edge_n) + al(flag_n)) * sizeof(float), st));
    float* Ci = scratch;
    float* Hi = Ci + al(n);
    float* imgf = Hi + al(n);
    float* edge = imgf + img_n;
    int* flags = reinterpret_cast<int*>(edge + al(edge_n));
    const int ieee = g_ieee_div.load();
    dim3 tgrid(cdiv(w, 32), DP / 32, h);
    if (volc_type == ROO_VOL_F32)
        vol_to_internal_kernel<float><<<tgrid, 256, 0, st>>>(Ci, Vol<float>(*volC), DP, maxDisp, ieee);
    else
        vol_to_internal_kernel<roo_costvolelem_t><<<tgrid, 256, 0, st>>>(Ci, Vol<roo_costvolelem_t>(*volC), DP, maxDisp, ieee);
    count_launch();
    int rc = launch_status();
    if (rc == 0) rc = launch_image_to_f32(imgf, left->ptr, left->pitch, 0, img_type, w, h, 1, 1.0f, st);

    SweepArgs a{};
    a.H = Hi; a.h_pair = n; a.C = Ci; a.c_pair = n;
    a.img = imgf; a.img_pair = 0; a.cost_scale = 1.0f;
    a.w = w; a.h = h; a.DP = DP; a.maxDisp = maxDisp; a.batch = 1;
    a.ieee = ieee;
    a.P1 = P1; a.P2 = P2; a.cost_kind = COST_F32; a.epi = EPI_NONE; a.subpix = 0; a.disp = nullptr; a.disp_pair = 0;
    for (int i = 0; i < plan.n && rc == 0; ++i) {
        a.first = i == 0;
        rc = launch_pass(a, plan.pass[i], edge, flags, st);
    }
    if (rc == 0) rc = launch_internal_to_vol(volH, Hi, DP, maxDisp, st);
    cudaError_t fe = cudaFreeAsync(scratch, st);
    return rc != 0 ? rc : (int)fe;
}
