// Horizontal aggregation sweeps (paths (+1,0) and (-1,0); src/cu_semi_global_matching.cu:79-84 of the
// reference launches them as ONE block of h threads with stride-`pitch` accesses).
//
// On the internal layout H[pair][y][x][DP] a horizontal scanline is ONE contiguous run of memory
// (w * DP * 4 bytes), so this kernel does not move its rows with per-lane loads at all:
//  * one warp per scanline; lane l owns disparities [l*DPL, (l+1)*DPL), previous pixel's row in registers,
//    d-1 / d+1 neighbours by two shuffles, min over d by one redux.sync (as in sgm.cu);
//  * the aggregate and cost rows of CH consecutive pixels (5 KB) are ONE chunk, copied global -> shared by the
//    bulk-copy engine (cp.async.bulk, SASS UBLKCP) that ONE elected lane issues, completion counted in bytes on an
//    mbarrier (expect_tx / try_wait.parity).  NST chunks per warp are in flight; no lane spends an issue slot on a
//    load, an address or a cp.async group -- the sweep was bound by instruction issue, not by HBM, once the
//    winner-takes-all epilogue rides on it;
//  * intensities for the adaptive P2: lane k keeps the pixel 32*blk + k of the scanline (one coalesced load per 32
//    pixels, the next block already in flight); a step takes its value with one shuffle;
//  * direction is a template parameter: every address inside a chunk is an immediate offset.
// Numerics: sgm_step() -- identical to the generic sweep, the reference and the oracle.
#include "common.cuh"
#include "kernels.cuh"
#include "sgm_step.cuh"

#include <type_traits>

namespace roo_b200 {

#ifndef HS_WARPS_N
#define HS_WARPS_N 4
#endif
// measured on B200, 1280x720x128 x16 (right + left pass, ms): 2.5 KB x 4 chunks 4.49, 2.5 KB x 3 4.41, 1.25 KB x 4 4.83,
// 1.25 KB x 8 4.94, 5 KB x 3 4.36, 5 KB x 2 4.28 -- the bulk-copy engine likes few large copies
#ifndef HS_NST_N
#define HS_NST_N 2
#endif
#ifndef HS_CHUNK_BYTES
#define HS_CHUNK_BYTES 5120
#endif
constexpr int HS_WARPS = HS_WARPS_N;   // scanlines per CTA
constexpr int HS_NST = HS_NST_N;       // chunks in flight per warp

// pixels per chunk: ~5 KB of aggregate + cost per bulk copy pair
template <int DPL, int COST> __host__ __device__ constexpr int hs_chunk() {
    constexpr int px_bytes = 32 * DPL * (4 + RawCost<DPL, COST>::ELEM);
    constexpr int n = HS_CHUNK_BYTES / px_bytes;
    return n >= 16 ? 16 : (n >= 8 ? 8 : (n >= 4 ? 4 : (n >= 2 ? 2 : 1)));
}

__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned sdst, const void* gsrc, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "HS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HS_DONE;\n\t"
        "bra HS_WAIT;\n"
        "HS_DONE:\n\t}"
        ::"r"(mbar), "r"(parity) : "memory");
}

// DX = +1: path (+1,0), travel index t == x;  DX = -1: path (-1,0), t == w-1-x.
// SUBPIX (epilogue only): compile-time, so that the eight pixels of a chunk are ONE straight-line block in the plain
// winner-takes-all case (no uniform branch per pixel: the census window rotates by register renaming, not by moves)
template <int DPL, int COST, int EPI, bool FIRST, bool IEEE, int DX, bool SUBPIX>
__global__ void __launch_bounds__(HS_WARPS * 32)
sgm_hsweep_kernel(const SweepArgs a) {
    constexpr int DP = 32 * DPL;
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    constexpr int CH = hs_chunk<DPL, COST>();
    constexpr int NST = HS_NST;
    constexpr unsigned HROW = DP * 4, CROW = DP * CE;            // bytes of one pixel's aggregate / cost row
    constexpr unsigned STAGE_B = CH * (HROW + CROW);             // [CH aggregate rows][CH cost rows], ascending x
    constexpr bool CEN = COST == COST_CEN32;                     // cost recomputed from census words, nothing staged for it
    constexpr bool STAGED = !FIRST || CROW > 0;                  // anything to prefetch at all?
    extern __shared__ __align__(128) unsigned char hs_smem[];    // [HS_WARPS][NST][STAGE_B], then the mbarriers

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int y = blockIdx.x * HS_WARPS + warp;
    const int pair = blockIdx.y;
    const int w = a.w, M = a.maxDisp;
    constexpr int subpix = SUBPIX ? 1 : 0;
    const float P1 = a.P1, P2 = a.P2, cscale = a.cost_scale;

    const unsigned sbase = (unsigned)__cvta_generic_to_shared(hs_smem) + warp * NST * STAGE_B;
    const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(hs_smem) + HS_WARPS * NST * STAGE_B + warp * NST * 8;
    // per-warp row of DP floats for the epilogue's parabola taps (after the mbarriers; only when an epilogue runs)
    const unsigned wta_scratch = EPI != EPI_NONE ? (unsigned)__cvta_generic_to_shared(hs_smem) + HS_WARPS * NST * STAGE_B +
                                                       HS_WARPS * NST * 8 + warp * DP * 4 : 0u;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(mbar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async proxy
    }
    __syncwarp();
    if (y >= a.h) return;

    // scanline bases (element (0, y, lane's first disparity))
    const size_t row0 = (size_t)y * w;
    float* const Hrow = a.H + (size_t)pair * a.h_pair + row0 * DP;
    const char* const Crow = (const char*)a.C + ((size_t)pair * a.c_pair + row0 * DP) * CE;
    const float* const Irow = a.img + (size_t)pair * a.img_pair + row0;
    float* const Drow = (EPI != EPI_NONE) ? a.disp + (size_t)pair * a.disp_pair + row0 : nullptr;

    const int nchunks = (w + CH - 1) / CH;
    // chunk k covers travel indices [k*CH, k*CH + n): x ascending from xlo = (DX > 0 ? k*CH : w - k*CH - n);
    // its rows go to the stage slots [s0, s0 + n) with s0 = (DX > 0 ? 0 : CH - n), so that travel index i always
    // sits in slot (DX > 0 ? i : CH-1-i)
    auto issue = [&](int k) {   // lane 0 only
        const int t0 = k * CH, n = min(CH, w - t0);
        const int xlo = DX > 0 ? t0 : w - t0 - n, s0 = DX > 0 ? 0 : CH - n;
        const unsigned st = sbase + (unsigned)(k % NST) * STAGE_B, mb = mbar0 + 8 * (unsigned)(k % NST);
        mbar_expect_tx(mb, (unsigned)n * ((FIRST ? 0u : HROW) + CROW));
        if (!FIRST) bulk_g2s(st + s0 * HROW, Hrow + (size_t)xlo * DP, (unsigned)n * HROW, mb);
        if (CROW > 0) bulk_g2s(st + CH * HROW + s0 * CROW, Crow + (size_t)xlo * DP * CE, (unsigned)n * CROW, mb);
    };
    if (STAGED && lane == 0)
        for (int k = 0; k < NST && k < nchunks; ++k) issue(k);

    // ---- COST_CEN32: sliding window of right-image census words in registers -------------------------------
    // rw[j] = low word of R(x - d0 - j, y) for this lane's disparities.  A step to the next pixel shifts the window by
    // one disparity: one shuffle between neighbouring lanes plus ONE new word per warp (d = 0 entering at lane 0 going
    // right, d = DP-1 entering at lane 31 going left).  New words and the left image's words come out of per-32-pixel
    // register blocks like the intensities.  Words of pixels left of the image are 0: their disparities are masked.
    const unsigned long long* const Lc = CEN ? a.cenL + (size_t)pair * a.cen_pair + row0 : nullptr;
    const unsigned long long* const Rc = CEN ? a.cenR + (size_t)pair * a.cen_pair + row0 : nullptr;
    auto cen_word = [&](const unsigned long long* row, int x) -> unsigned {
        return (x >= 0 && x < w) ? __ldg(reinterpret_cast<const unsigned*>(row + x)) : 0u;   // little endian: low half first
    };
    // travel index tau <-> x: DX > 0: x = tau;  DX < 0: x = w-1-tau.  The word entering the window at step t is
    // R(x) going right (tau = t) and R(x - (DP-1)) going left (tau = t + DP-1).
    constexpr int RSHIFT = DX > 0 ? 0 : DP - 1;
    auto gatherL = [&](int blk) { const int t = 32 * blk + lane; return cen_word(Lc, DX > 0 ? t : w - 1 - t); };
    auto gatherR = [&](int blk) { const int t = 32 * blk + lane; return cen_word(Rc, DX > 0 ? t : w - 1 - t); };
    unsigned lcur = 0, lnxt = 0, rcur = 0, rnxt = 0, rw[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) rw[j] = 0;
    if (CEN) {
        lcur = gatherL(0); lnxt = gatherL(1);
        rcur = gatherR(RSHIFT >> 5); rnxt = gatherR((RSHIFT >> 5) + 1);
        if (DX < 0) {   // window of the position one step before the start (x = w): R(w - d), d >= 1
#pragma unroll
            for (int j = 0; j < DPL; ++j) rw[j] = cen_word(Rc, w - lane * DPL - j);
        }
    }

    // intensities in travel order: lane k holds pixel t = 32*blk + k
    auto gather = [&](int blk) {
        const int t = 32 * blk + lane;
        return t < w ? __ldg(Irow + (DX > 0 ? t : w - 1 - t)) : 0.0f;
    };
    float icur = gather(0), inxt = gather(1);

    float hp[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) hp[j] = ROO_INF;   // path start: no previous pixel
    float lastBest = 0.0f, last_c = 0.0f;
    const int d0 = lane * DPL;
    const int xf = (M == DP) ? DP - 1 : 0x3fffffff;   // every lane in range iff x >= xf

    auto chunk = [&](auto masked_tag, auto full_tag, int k, unsigned st) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        constexpr bool FULL = decltype(full_tag)::value;
        const int t0 = k * CH, n = FULL ? CH : min(CH, w - t0);
        const int x0 = DX > 0 ? t0 : w - 1 - t0;    // x of travel index t0
        float* const hst = Hrow + (size_t)x0 * DP + d0;
        const bool first = k == 0;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (FULL || i < n) {
                const unsigned slot = DX > 0 ? i : CH - 1 - i;
                float hin[DPL], hnew[DPL], craw[DPL], best;
                if (!FIRST) lds_vec<DPL>(hin, st + slot * HROW + lane * DPL * 4);
                RawCost<DPL, COST> rc;
                if (!CEN) {
                    rc.lds(st + CH * HROW + slot * CROW + lane * DPL * CE);
#pragma unroll
                    for (int j = 0; j < DPL; ++j) craw[j] = rc.raw(j);
                } else {
                    // the R block advances where (t + RSHIFT) & 31 == 0: going left that is t & 31 == 1, i.e. i == 1
                    // of a chunk that starts a 32-pixel block (chunks never straddle blocks)
                    if (DX < 0 && i == (1 % CH) && ((t0 + i + RSHIFT) & 31) == 0) {
                        rcur = rnxt; rnxt = gatherR(((t0 + i + RSHIFT) >> 5) + 1);
                    }
                    const unsigned nw = __shfl_sync(0xffffffffu, rcur, (t0 + i + RSHIFT) & 31);
                    const unsigned lw = __shfl_sync(0xffffffffu, lcur, (t0 + i) & 31);
                    if (DX > 0) {
                        const unsigned up = __shfl_up_sync(0xffffffffu, rw[DPL - 1], 1);
#pragma unroll
                        for (int j = DPL - 1; j > 0; --j) rw[j] = rw[j - 1];
                        rw[0] = lane == 0 ? nw : up;
                    } else {
                        const unsigned dn = __shfl_down_sync(0xffffffffu, rw[0], 1);
#pragma unroll
                        for (int j = 0; j < DPL - 1; ++j) rw[j] = rw[j + 1];
                        rw[DPL - 1] = lane == 31 ? nw : dn;
                    }
#pragma unroll
                    for (int j = 0; j < DPL; ++j) craw[j] = (float)__popc(lw ^ rw[j]);
                }
                const float pix = __shfl_sync(0xffffffffu, icur, (t0 + i) & 31);
                // start pixel: `volH += volC`, lastBestCr = 0 (cu_semi_global_matching.cu:31-35) == a step with P2 = 0
                const float p2 = (i == 0 && first) ? 0.0f : P2;
                const float denom = 1.0f + fabsf(last_c - pix);
                const int x = x0 + DX * i;
                const int lim = MASKED ? min(M, x + 1) - d0 : 0;
                sgm_step<DPL, MASKED, FIRST, IEEE>(hp, lastBest, denom, P1, p2, craw, cscale, hin, lim, lane, hnew, hp, best);
                lastBest = (i == 0 && first) ? 0.0f : best;
                last_c = pix;
                if (EPI != EPI_WTA_ONLY) store_f<DPL>(hst + DX * i * DP, hnew);
                if (EPI != EPI_NONE) {
                    const float out = wta_epilogue<DPL, IEEE>(hp, lane, x, w, M, subpix, wta_scratch);
                    // lane 0 stores; a predicated store, not a branch: the pixels of a chunk stay one basic block
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %0, 0;\n\t@p st.global.f32 [%1], %2;\n\t}" ::"r"(lane), "l"(Drow + x), "f"(out) : "memory");
                }
            }
        }
    };

    const int tmask_lo = DX > 0 ? xf : 0x3fffffff;              // DX > 0: unmasked from t >= xf on
    const int tmask_hi = DX > 0 ? 0x3fffffff : w - 1 - xf;      // DX < 0: unmasked while t <= w-1-xf (x >= xf)
#pragma unroll 1
    for (int k = 0; k < nchunks; ++k) {
        const unsigned s = (unsigned)(k % NST);
        const unsigned st = sbase + s * STAGE_B;
        if (((k * CH) & 31) == 0 && k != 0) {
            icur = inxt; inxt = gather((k * CH >> 5) + 1);
            if (CEN) {
                lcur = lnxt; lnxt = gatherL((k * CH >> 5) + 1);
                if (DX > 0) { rcur = rnxt; rnxt = gatherR((k * CH >> 5) + 1); }
            }
        }
        if (STAGED) mbar_wait(mbar0 + 8 * s, (unsigned)(k / NST) & 1u);
        const int t0 = k * CH;
        const bool unmasked = t0 >= tmask_lo && t0 + CH - 1 <= tmask_hi;
        const bool full = t0 + CH <= w;
        if (full) {
            if (unmasked) chunk(std::false_type{}, std::true_type{}, k, st);
            else chunk(std::true_type{}, std::true_type{}, k, st);
        } else {
            chunk(std::true_type{}, std::false_type{}, k, st);
        }
        if (STAGED) {
            __syncwarp();   // every lane has read the stage: it may be refilled
            if (lane == 0 && k + NST < nchunks) issue(k + NST);
        }
    }
}

template <int DPL, int COST, int EPI, int DX>
static int hsweep_launch4(const SweepArgs& a, cudaStream_t st) {
    constexpr int CH = hs_chunk<DPL, COST>();
    constexpr size_t smem = (size_t)HS_WARPS * HS_NST * (CH * 32 * DPL * (4 + RawCost<DPL, COST>::ELEM)) + HS_WARPS * HS_NST * 8 +
                            (EPI != EPI_NONE ? (size_t)HS_WARPS * 32 * DPL * 4 : 0);
    static_assert(32 % CH == 0 || CH % 32 == 0, "chunks must not straddle a 32-pixel intensity block");
    dim3 grid(cdiv(a.h, HS_WARPS), a.batch);
    const bool ieee = a.ieee != 0;
#define ROO_HS(F, I)                                                                                          \
    do {                                                                                                      \
        auto kern = (EPI != EPI_NONE && a.subpix) ? sgm_hsweep_kernel<DPL, COST, EPI, F, I, DX, EPI != EPI_NONE>    \
                                                  : sgm_hsweep_kernel<DPL, COST, EPI, F, I, DX, false>;       \
        if (smem > 48 * 1024) {                                                                               \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                              \
        }                                                                                                     \
        kern<<<grid, HS_WARPS * 32, smem, st>>>(a);                                                           \
    } while (0)
    if (a.first) { if (ieee) ROO_HS(true, true); else ROO_HS(true, false); }
    else { if (ieee) ROO_HS(false, true); else ROO_HS(false, false); }
#undef ROO_HS
    count_launch();
    return launch_status();
}

template <int DPL, int COST>
static int hsweep_launch2(const SweepArgs& a, cudaStream_t st) {
    const bool fwd = a.dx > 0;
    switch (a.epi) {
        case EPI_NONE: return fwd ? hsweep_launch4<DPL, COST, EPI_NONE, 1>(a, st) : hsweep_launch4<DPL, COST, EPI_NONE, -1>(a, st);
        case EPI_WTA_WRITE: return fwd ? hsweep_launch4<DPL, COST, EPI_WTA_WRITE, 1>(a, st) : hsweep_launch4<DPL, COST, EPI_WTA_WRITE, -1>(a, st);
        default: return fwd ? hsweep_launch4<DPL, COST, EPI_WTA_ONLY, 1>(a, st) : hsweep_launch4<DPL, COST, EPI_WTA_ONLY, -1>(a, st);
    }
}

// a.dy == 0, a.dx == +-1
#ifdef HS_PART
#define ROO_HS_CAT2(a, b) a##b
#define ROO_HS_CAT(a, b) ROO_HS_CAT2(a, b)
int ROO_HS_CAT(launch_hsweep_dpl, HS_PART)(const SweepArgs& a, cudaStream_t st) {
#else
int launch_hsweep(const SweepArgs& a, cudaStream_t st) {
#endif
#define ROO_HS_DP(DPL)                                                          \
    switch (a.cost_kind) {                                                      \
        case COST_F32: return hsweep_launch2<DPL, COST_F32>(a, st);             \
        case COST_U8: return hsweep_launch2<DPL, COST_U8>(a, st);               \
        case COST_CEN32: return hsweep_launch2<DPL, COST_CEN32>(a, st);         \
        default: return ROO_ERR_INVALID_ARGUMENT;                               \
    }
    // The instantiations of one disparity count are one translation unit each (the Makefile compiles this file four
    // times with -DHS_PART=<DPL>): build time, nothing else.
#ifdef HS_PART
    ROO_HS_DP(HS_PART)
#else
    switch (a.DP) {
        case 32: ROO_HS_DP(1)
        case 64: ROO_HS_DP(2)
        case 128: ROO_HS_DP(4)
        case 256: ROO_HS_DP(8)
        case 512: ROO_HS_DP(16)
        default: return ROO_ERR_UNSUPPORTED;
    }
#endif
#undef ROO_HS_DP
}

}  // namespace roo_b200
