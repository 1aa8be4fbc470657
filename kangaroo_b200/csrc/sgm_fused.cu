// Fused vertical path group: the three paths that share a travel direction in y --
//   forward  (0,+1) down, (+1,+1) down-right, (-1,+1) down-left
//   reverse  (0,-1) up,   (-1,-1) up-left,    (+1,-1) up-right
// -- aggregated in ONE pass over the volume instead of three (DESIGN.md "vertical group").
//
// Semantics are exactly three consecutive launches of the reference kernel body
// (src/cu_semi_global_matching.cu:21-63) in that order: path k at pixel p adds its Cr to the aggregate
// left there by path k-1 and uses as "previous row" the aggregate it wrote itself at its previous pixel.
// A pixel therefore needs, besides its own H, only the state of three pixels of the previous image row.
//
// Mapping.  Work in travel coordinates (x', y') (reverse group = image rotated by 180 degrees) and skewed
// columns u = x' - y'.  A skewed column is walked row by row by one warp, lane l holding disparities
// [l*DPL, (l+1)*DPL):
//   * the diagonal path (+1,+1) stays inside the warp: its state never leaves registers;
//   * the vertical path needs the state of column u+1, the anti-diagonal path that of column u+2, both of
//     the previous row -> every dependency points towards HIGHER u.  Inside a CTA (a band of NW columns)
//     the states go through a shared-memory ring guarded by per-warp progress words (no CTA barrier);
//     between CTAs they flow one way only, from band b-1 to band b, through a small L2-resident edge
//     buffer guarded by a monotonic progress flag.  One-way dependencies make the bands a pipeline, not a
//     ping-pong: a band never waits for a band that waits for it, and lower block indices (scheduled
//     first) never wait for higher ones.
//   * a dedicated communication warp per CTA polls the upstream flag, stages the upstream edge rows into
//     shared memory and publishes this band's flag, so the compute warps never touch the global flags.
//   * every compute warp owns NCW = four adjacent skewed columns c0 = u .. and advances
//     all of them by one row per tick: the vertical path of column c continues from column c+1's state of the
//     previous tick and the anti-diagonal path from column c+2's (registers), so only three state rows per tick
//     (c0.vertical, c0.anti-diagonal, c1.anti-diagonal) go through shared memory, and the per-tick bookkeeping
//     (hand-off flags, prefetch issue, addressing) is paid once for NCW pixels.
// HBM traffic of the pass: read H (unless first) + read cost + write H -- the same as ONE single-path sweep.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "sgm_step.cuh"

namespace roo_b200 {

// rows of prefetch per column, staged in shared memory by cp.async (LDGSTS): under load a DRAM access takes
// ~3000 SM cycles on B200, so a band needs ~60-80 KB in flight per SM to stream at HBM speed -- far more
// than a register ring can hold, and without unrolling the row loop.  (227 KB of shared memory per CTA.)
__host__ __device__ constexpr int vg_pfs(int DPL, int CE) {
    return DPL >= 8 ? 2 : (DPL == 4 ? (CE == 4 ? 2 : 4) : (DPL == 2 && CE == 4 ? 4 : 8));
}
// skewed columns per compute warp, and compute warps per band (+ 1 communication warp)
#ifndef VG_NCW
#define VG_NCW 4
#endif
// 256 disparities (229+ registers per thread with four columns): measured on B200, fused pair of passes,
// 1920x1080x256 x4 / 3840x2160x256 x1 (ms): 6 warps x 4 columns 8.45 / 11.97, 8 x 4 8.74 / 11.50, 12 x 2 9.09 / 9.88,
// 8 x 3 (state ring 2 rows deep) 7.98 / 10.47
#ifndef VG_NCW8
#define VG_NCW8 3
#endif
#ifndef VG_NWW8
#define VG_NWW8 8
#endif
// a single pair at 256 disparities (BASELINE config 5) has no second pair to fill the time a band waits for its
// predecessor: 12 warps x 2 columns (same 24-column bands, more warps per SM) instead of 8 x 3 -- geometry class 3
#ifndef VG_NCW8_SOLO
#define VG_NCW8_SOLO 2
#endif
#ifndef VG_NWW8_SOLO
#define VG_NWW8_SOLO 12
#endif
__host__ __device__ constexpr int vg_ncw(int DPL, int cls = 0) { return DPL >= 8 ? (cls == 3 ? VG_NCW8_SOLO : VG_NCW8) : VG_NCW; }
// 12 warps x 4 columns: 13 warps of <= 152 registers fill the register file, and ring + prefetch stages fill
// the 227 KB of shared memory (measured on B200 at 128 disparities: 8 warps 6.6 ms, 10: 6.5 ms, 12: 6.1 ms per 16 pairs)
#ifndef VG_NWW
#define VG_NWW 12
#endif
// a write-only first pass with in-sweep cost prefetches nothing but census strips: shared memory leaves room for more warps
#ifndef VG_NWW_FC
#define VG_NWW_FC VG_NWW
#endif
#ifndef VG_NWW8_FC
#define VG_NWW8_FC VG_NWW8_CEN
#endif
// narrow bands of a first pass with in-sweep cost need little shared memory: several of them share an SM
#ifndef VG_MINB_FC
#define VG_MINB_FC 2
#endif
__host__ __device__ constexpr int vg_min_ctas(int DPL, int cost, bool first, int nww) {
    return (first && cost == COST_CEN32 && nww <= (DPL >= 8 ? 3 : 6)) ? VG_MINB_FC : 1;
}
// with in-sweep cost the prefetch stages hold no cost rows: at 256 disparities that leaves room for more warps per band
#ifndef VG_NWW8_CEN
#define VG_NWW8_CEN VG_NWW8
#endif
// geometry class of a launch: 0 = cost read from a volume, 1 = in-sweep cost, 2 = in-sweep cost and write-only first pass,
// 3 = in-sweep cost, one pair per launch at 256 disparities
__host__ __device__ constexpr int vg_nww(int DPL, int cls = 0) {
    return DPL >= 8 ? (cls == 3 ? VG_NWW8_SOLO : (cls == 2 ? VG_NWW8_FC : (cls == 1 ? VG_NWW8_CEN : VG_NWW8)))
                    : (cls == 2 ? VG_NWW_FC : VG_NWW);
}
inline int vg_cols_of_dp(int DP, int cls = 0) { return vg_ncw(DP / 32, cls) * vg_nww(DP / 32, cls); }

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int DPL>
__device__ __forceinline__ void lds_row(float (&v)[DPL], const float* p) { load_f<DPL>(v, p); }
template <int DPL>
__device__ __forceinline__ void sts_row(float* p, const float (&v)[DPL]) { store_f<DPL>(p, v); }
template <int DPL>
__device__ __forceinline__ void ldcg_row(float (&v)[DPL], const float* p) {
    if constexpr (DPL >= 4) {
#pragma unroll
        for (int q = 0; q < DPL / 4; ++q) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(p) + q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else if constexpr (DPL == 2) {
        const float2 t = __ldcg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = __ldcg(p);
    }
}
template <int DPL>
__device__ __forceinline__ void stcg_row(float* p, const float (&v)[DPL]) {
    if constexpr (DPL >= 4) {
#pragma unroll
        for (int q = 0; q < DPL / 4; ++q)
            __stcg(reinterpret_cast<float4*>(p) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    } else if constexpr (DPL == 2) {
        __stcg(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    } else {
        __stcg(p, v[0]);
    }
}

__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
template <int DPL>
__device__ __forceinline__ void cp_async_row(float* smem_dst, const float* gsrc) {
    static_assert(DPL >= 4, "cp.async.cg moves 16 bytes");
#pragma unroll
    for (int q = 0; q < DPL / 4; ++q) cp_async_16(smem_dst + 4 * q, gsrc + 4 * q);
}

// Ordering of shared-memory accesses between warps of one CTA.  Data and flag both live in shared memory
// and every access is issued through the same in-order LSU pipeline of the SM, so program order (enforced
// for the compiler by the volatile flag accesses and this barrier) is enough; a MEMBAR here would also wait
// for the warp's outstanding GLOBAL prefetch loads and serialise every row on DRAM latency.
__device__ __forceinline__ void smem_order() { asm volatile("" ::: "memory"); }
// VG_RELACQ=1 (default): the in-band progress words are published with st.release.cta and polled with ld.acquire.cta --
// race-free under the PTX memory model.  Measured on B200 (1280x720x128 x16, both fused passes): 5.98 ms against 5.92 ms
// for -DVG_RELACQ=0, the round-1 variant that relied on program order through the SM's in-order shared-memory pipeline
// (volatile accesses + a compiler barrier); a full __threadfence_block() per row had cost 27 %.
#ifndef VG_RELACQ
#define VG_RELACQ 1
#endif
__device__ __forceinline__ void publish_flag(volatile int* flag, int v) {
#if VG_RELACQ
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(const_cast<int*>(flag))), "r"(v) : "memory");
#else
    *flag = v;
#endif
}

#ifndef VG_SPIN_NS
#define VG_SPIN_NS 30
#endif
#ifndef VG_COMM_NS
#define VG_COMM_NS 200   // idle back-off of the communication warp
#endif
// poll a monotonically growing shared-memory flag until it reaches `need`
__device__ __forceinline__ void spin_until(unsigned flag_addr, int need) {
    int v;
#if VG_RELACQ
#define VG_FLAG_LD "ld.acquire.cta.shared.s32 %0, [%1];"
#else
#define VG_FLAG_LD "ld.volatile.shared.s32 %0, [%1];"
#endif
    asm volatile(VG_FLAG_LD : "=r"(v) : "r"(flag_addr) : "memory");
    while (v < need) {
        __nanosleep(VG_SPIN_NS);
        asm volatile(VG_FLAG_LD : "=r"(v) : "r"(flag_addr) : "memory");
    }
#undef VG_FLAG_LD
}

// smem control words
struct VCtl { volatile int halo_ready; volatile int copied; int pad[2]; };


#ifndef VG_R
#define VG_R 4
#endif
__host__ __device__ constexpr int vg_r(int DPL) { return VG_R; }   // max rows per hand-off batch between bands (ring = 2x)
// depth of the in-band state ring: 2 rows measured 1 % faster than 4 (c2: 6.13 -> 6.06 ms for the two fused passes) and
// frees 37 KB of shared memory
#ifndef VG_S_DEPTH
#define VG_S_DEPTH 2
#endif
#ifndef VG_S_DEPTH8
#define VG_S_DEPTH8 2
#endif
// depth (rows) of the in-band state ring in shared memory
__host__ __device__ constexpr int vg_s(int DPL) { return DPL >= 8 ? VG_S_DEPTH8 : VG_S_DEPTH; }

// One pixel of the three paths.  V/D/A = vertical / diagonal / anti-diagonal.  hpV, hpD, hpA: previous pixel's
// state rows on entry, this pixel's on exit.  Handles path starts when EDGE.
template <int DPL, bool MASKED, bool EDGE, bool FIRST, bool IEEE>
__device__ __forceinline__ void vg_pixel(unsigned stg, const float (&cost)[DPL], int lane, int y, int xp, int x, int w, int M,
                                         float P1, float P2, float cscale, float pix, float (&hpV)[DPL], float& lbV, float ppV,
                                         float (&hpD)[DPL], float& lbD, float& pixD,
                                         float (&hpA)[DPL], float& lbA, float ppA, float* hst) {
    const int lim = MASKED ? min(M, x + 1) - lane * DPL : 0;
    float hin[DPL], H3[DPL];
    if (!FIRST) lds_vec<DPL>(hin, stg + lane * DPL * 4);
    float p2V = P2, p2D = P2, p2A = P2;
    bool sV = false, sD = false, sA = false;
    if (EDGE) {
        sV = y == 0; sD = y == 0 || xp == 0; sA = y == 0 || xp == w - 1;
        if (sV) { lbV = 0.0f; ppV = pix; p2V = 0.0f; }
        if (sD) { lbD = 0.0f; p2D = 0.0f; }
        if (sA) { lbA = 0.0f; ppA = pix; p2A = 0.0f; }
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            hpV[j] = sV ? ROO_INF : hpV[j];
            hpD[j] = sD ? ROO_INF : hpD[j];
            hpA[j] = sA ? ROO_INF : hpA[j];
        }
    }
    float bV, bD, bA;
    sgm_step3<DPL, MASKED, FIRST, IEEE>(hpV, lbV, 1.0f + fabsf(ppV - pix), p2V,
                                        hpD, lbD, 1.0f + fabsf(pixD - pix), p2D,
                                        hpA, lbA, 1.0f + fabsf(ppA - pix), p2A,
                                        cost, cscale, hin, P1, lim, lane, H3, bV, bD, bA);
    if (EDGE) { if (sV) bV = 0.0f; if (sD) bD = 0.0f; if (sA) bA = 0.0f; }
    lbV = bV; lbD = bD; lbA = bA;
    pixD = pix;
    store_f<DPL>(hst, H3);
}

template <int DPL, int COST, bool FIRST, bool IEEE, int NWW, int NCW, bool FWD>
__global__ void __launch_bounds__((NWW + 1) * 32, vg_min_ctas(DPL, COST, FIRST, NWW))
sgm_vgroup_kernel(const VGroupArgs a) {
    constexpr int DP = 32 * DPL;
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    constexpr int NC = NCW * NWW;                     // skewed columns per band
    constexpr int PFS = vg_pfs(DPL, CE);
    constexpr int R = vg_r(DPL), RING = 2 * R, S = vg_s(DPL);
    constexpr int HOFF_B = FIRST ? 0 : DP * 4;       // a write-only first pass prefetches no aggregate rows
    constexpr int STAGE_B = HOFF_B + DP * CE;        // one prefetched pixel: aggregate row, cost row
    // COST_CEN32: no cost rows; instead ONE strip of census words per warp and image row, from which every lane
    // recomputes popc(L ^ R) for its NCW x DPL (pixel, disparity) pairs:
    //   seg[0 .. DP+NCW-2]   low words of the right descriptors R(S0 + i), S0 = x(lowest-address column) - (DP-1)
    //   seg[DP+8 .. DP+8+NCW-1]  low words of the NCW left descriptors, ascending x
    // (lane l reads the DPL+NCW-1 words from seg[DPL*(31-l)] on -- 16-byte aligned, conflict-free LDS.128)
    constexpr bool CEN = COST == COST_CEN32;
    constexpr int CEN_W = DP + 16, CEN_B = CEN_W * 4;
    constexpr int WIN = (DPL + NCW - 1 + 3) / 4 * 4;   // words of the lane's window, rounded up to whole LDS.128
    extern __shared__ __align__(16) float smem[];
    // state rows of one warp and one image row, for its two lowest columns c0 and c1 (all that the warp below
    // needs): rec0 = c0.vertical, rec1 = c0.anti-diagonal, rec2 = c1.anti-diagonal;
    // scalars {c0.lastBest(V), c0.lastBest(A), c0.pix, -, -, c1.lastBest(A), c1.pix, -}: the same record layout is
    // used by the band-to-band edge rows, so a warp reads "the three rows of whoever is above me" with one formula
    float* s_hp = smem;                                // [S rows][NWW][3][DP]
    float* s_sc = s_hp + S * NWW * 3 * DP;             // [S rows][NWW][8]
    float* s_halo = s_sc + S * NWW * 8;                // [RING rows][3][DP]  upstream band's lowest warp (row ring)
    float* s_hsc = s_halo + RING * 3 * DP;             // [RING][8]
    float* s_edge = s_hsc + RING * 8;                  // [RING rows][3][DP]  this band's lowest warp, for downstream
    float* s_esc = s_edge + RING * 3 * DP;             // [RING][8]
    VCtl* ctl = reinterpret_cast<VCtl*>(s_esc + RING * 8);
    volatile int* prog = reinterpret_cast<volatile int*>(ctl + 1);   // [NWW] rows < prog[v] of warp v are done
    char* s_pf = reinterpret_cast<char*>(ctl + 1) + ((NWW * 4 + 15) / 16) * 16;   // [NWW][NCW cols][PFS][STAGE_B]
    char* s_cen = s_pf + (size_t)NWW * NCW * PFS * STAGE_B;                        // [NWW][PFS][CEN_B] (COST_CEN32 only)

    // warp index through a shuffle: ptxas then knows it is warp-uniform, and every branch on it is a uniform branch
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // pair fastest: the resident window of CTAs then holds the same few bands of EVERY pair, so the
    // band-to-band pipeline of each pair has only a short ramp
    // The logical CTA index is a ticket drawn when the CTA starts, not blockIdx: band b spins on flags of band b-1 of the
    // same pair, which holds a LOWER ticket and has therefore already been dispatched -- forward progress no longer
    // depends on the hardware handing out CTAs in blockIdx order (which CUDA does not promise).
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int cta = s_ticket;
    const int pair = cta % a.batch, band = cta / a.batch;
    const int w = a.w, h = a.h, M = a.maxDisp;
    constexpr bool fwd = FWD;   // travel direction in y: compile-time, so that the columns' address offsets are immediates
    const float P1 = a.P1, P2 = a.P2, cscale = a.cost_scale;

    const int ulo = w - (band + 1) * NC;                 // lowest skewed column of this band
    const int ymin = max(0, -(ulo + NC - 1));
    const int ymax = min(h - 1, w - 1 - ulo);
    // upstream band (higher u) and the rows of it this band consumes: row y-1 for every own row y >= 1
    const int pulo = ulo + NC;
    const int pymin = max(0, -(pulo + NC - 1));
#ifdef VG_NO_HANDOFF   // timing experiment only (wrong results): bands do not wait for each other
    const int pymax = -1;
#else
    const int pymax = band > 0 ? min(h - 1, w - 1 - pulo) : -1;
#endif
    const int hbeg = max(pymin, ymin - 1), hend = min(pymax, ymax - 1) + 1;   // [hbeg, hend) upstream rows to stage
#ifdef VG_NO_HANDOFF
    const bool downstream = false;
#else
    const bool downstream = band + 1 < a.n_bands;
#endif

    float* e_hp = a.edge_hp + ((size_t)pair * a.n_bands + band) * (size_t)h * 3 * DP;   // this band's outgoing rows
    float* e_sc = a.edge_sc + ((size_t)pair * a.n_bands + band) * (size_t)h * 8;
    int* my_flag = a.progress + (size_t)pair * a.n_bands + band;

    // this warp's skewed columns u0 .. u0+NCW-1 (c0 = lowest u) and their active rows (x' = u + y' in [0, w))
    const int u0 = ulo + NCW * warp;
    // column c is active in row y (>= 0) iff 0 <= u0 + c + y < w and y < h: rows [max(0, -(u0+c)), min(h-1, w-1-u0-c)]
    auto col_active = [&](int c, int y) { return (unsigned)(u0 + c + y) < (unsigned)w && y < h; };
    const int y_in = max(0, -(u0 + NCW - 1)), y_out = min(h - 1, w - 1 - u0);   // the highest column enters first, the lowest leaves last
    const bool any = warp < NWW && y_in <= y_out;
    if (threadIdx.x == 0) { ctl->halo_ready = hbeg; ctl->copied = ymin; }
    if (warp < NWW && lane == 0) prog[warp] = any ? y_in : 0x7fffffff;   // rows before y_in never happen
    __syncthreads();

    if (warp == NWW) {
        // ---------------------------------------------------------------- communication warp
        const float* p_hp = e_hp - (size_t)h * 3 * DP;   // upstream band's rows
        const float* p_sc = e_sc - (size_t)h * 8;
        const int* p_flag = my_flag - 1;
        int seen = 0, hr = hbeg, cp = ymin;
        const int cp_end = downstream ? ymax + 1 : ymin;  // nothing to publish for the last band
        while (hr < hend || cp < cp_end) {
            bool progress = false;
            // ---- stage upstream rows into the halo ring (reader: the highest warp)
            if (hr < hend) {
                // adaptive batch: whatever the upstream band has published and the ring can take (1..R rows) --
                // the hand-off latency is one turn of this loop, not the time to fill a fixed chunk.  The paths
                // that run across the bands (anti-diagonal: a new band every NC/2 rows) are a serial chain of such
                // hand-offs, so this latency, not the copy bandwidth, bounds a single pair's pass.
                const int rdh = prog[NWW - 1];
                if (seen <= hr) {
                    if (lane == 0) seen = ld_acquire_gpu(p_flag);
                    seen = __shfl_sync(0xffffffffu, seen, 0);
                }
                // ring slot of row y was last used by row y-RING, read while computing row y-RING+1
                const int n = min(min(R, hend - hr), min(seen - hr, rdh + RING - 1 - hr));
                if (n > 0) {
                    for (int y = hr; y < hr + n; ++y) {
                        const float* src = p_hp + (size_t)y * 3 * DP + lane * DPL;
                        float* dst = s_halo + (size_t)(y % RING) * 3 * DP + lane * DPL;
                        if constexpr (DPL >= 4) {
#pragma unroll
                            for (int q = 0; q < 3; ++q) cp_async_row<DPL>(dst + q * DP, src + q * DP);
                        } else {   // < 16 B per lane: cp.async would need .ca, and L1 must not cache rows still being written
                            float r0[DPL], r1[DPL], r2[DPL];
                            ldcg_row<DPL>(r0, src); ldcg_row<DPL>(r1, src + DP); ldcg_row<DPL>(r2, src + 2 * DP);
                            sts_row<DPL>(dst, r0); sts_row<DPL>(dst + DP, r1); sts_row<DPL>(dst + 2 * DP, r2);
                        }
                        if (lane < 2) cp_async_16(s_hsc + (y % RING) * 8 + lane * 4, p_sc + (size_t)y * 8 + lane * 4);
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                    hr += n;
                    if (lane == 0) { __threadfence_block(); ctl->halo_ready = hr; }
                    progress = true;
                }
            }
            // ---- publish finished rows of this band's lowest warp
            if (cp < cp_end) {
                const int rd = min((int)prog[0], ymax + 1);
                if (rd > cp) {   // publish every finished row at once (see the latency note above)
                    __threadfence_block();
                    for (int y = cp; y < rd; ++y) {
                        const float* src = s_edge + (size_t)(y % RING) * 3 * DP + lane * DPL;
                        float* dst = e_hp + (size_t)y * 3 * DP + lane * DPL;
                        float r0[DPL], r1[DPL], r2[DPL];
                        lds_row<DPL>(r0, src);
                        lds_row<DPL>(r1, src + DP);
                        lds_row<DPL>(r2, src + 2 * DP);
                        stcg_row<DPL>(dst, r0);
                        stcg_row<DPL>(dst + DP, r1);
                        stcg_row<DPL>(dst + 2 * DP, r2);
                        if (lane < 8) __stcg(e_sc + (size_t)y * 8 + lane, s_esc[(y % RING) * 8 + lane]);
                    }
                    __threadfence();
                    __syncwarp();
                    cp = rd;
                    if (lane == 0) {
                        st_release_gpu(my_flag, rd == ymax + 1 ? 0x7fffffff : rd);
                        ctl->copied = rd;
                    }
                    progress = true;
                }
            }
            if (!progress) __nanosleep(VG_COMM_NS);
        }
        return;
    }
    if (!any) return;

    // -------------------------------------------------------------------- compute warps
    const int xf = (M == DP) ? DP - 1 : 0x3fffffff;               // all lanes in range iff true x >= xf
    const int d0 = lane * DPL;
    const ptrdiff_t estep = (fwd ? (ptrdiff_t)(w + 1) : -(ptrdiff_t)(w + 1)) * DP;   // one row down the travel direction

    // One cursor for all NCW columns: in a given row their pixels are adjacent in x, i.e. DP elements apart, so
    // column c lives at a compile-time offset from the column with the lowest address (c0 forward, the last one
    // backward).  The cursor is a linear function of the row; it may point outside the image while that column is
    // inactive, and is only dereferenced at offsets of active columns.
    float* const Hp = a.H + (size_t)pair * a.h_pair;
    const char* const Cp = (const char*)a.C + (size_t)pair * a.c_pair * CE;
    const float* const Ip = a.img + (size_t)pair * a.img_pair;
    auto coff = [](int c) { return (fwd ? c : NCW - 1 - c) * DP; };   // element offset of column c from the cursor
    const ptrdiff_t e_in = fwd ? ((ptrdiff_t)y_in * w + (u0 + y_in)) * DP + d0
                               : ((ptrdiff_t)(h - 1 - y_in) * w + (w - 1 - (u0 + NCW - 1) - y_in)) * DP + d0;
    float* hst = Hp + e_in;                          // row being computed
    const float* hld = hst;                          // row being prefetched
    const char* cld = Cp + e_in * CE;
    // COST_CEN32 cursors (one u64 descriptor per pixel; the arrays are padded, so strips may stick out of the row):
    // cenR: word i = 32k + lane of the strip;  cenX: lanes 0-7 the strip's words DP + lane, lanes 8-15 the left words
    const int xlow_in = fwd ? u0 + y_in : w - 1 - (u0 + NCW - 1) - y_in;           // x of the lowest-address column in row y_in
    const ptrdiff_t c_in = ((ptrdiff_t)(fwd ? y_in : h - 1 - y_in) * w + xlow_in);
    const unsigned long long* cenR = CEN ? a.cenR + (size_t)pair * a.cen_pair + c_in - (DP - 1) + lane : nullptr;
    const unsigned long long* cenX = CEN ? ((lane & 15) < 8 ? a.cenR + (size_t)pair * a.cen_pair + c_in - (DP - 1) + DP + (lane & 7)
                                                            : a.cenL + (size_t)pair * a.cen_pair + c_in + (lane & 7))
                                         : nullptr;
    const ptrdiff_t cstep = fwd ? (ptrdiff_t)(w + 1) : -(ptrdiff_t)(w + 1);
    // (through a shuffle: kept in a register instead of being rematerialised from the CTA's window base every row)
    const unsigned cen0 = __shfl_sync(0xffffffffu, (unsigned)__cvta_generic_to_shared(s_cen) + warp * PFS * CEN_B, 0);

    // Prefetch: row y+PFS-1 of every column is copied global -> shared (asynchronously, no registers) while row y
    // is computed.  Every lane copies and later reads its own bytes, so no barrier is needed.
    const unsigned pf0 = (unsigned)__cvta_generic_to_shared(s_pf) + (warp * NCW) * PFS * STAGE_B;
    auto issue_px = [&](int c, int yl) {
        const unsigned dst = pf0 + c * PFS * STAGE_B + ((unsigned)yl & (PFS - 1)) * STAGE_B;
        if (!FIRST) cp_async_bytes<DPL * 4>(dst + lane * DPL * 4, hld + coff(c));
        if (CE > 0) cp_async_bytes<DPL * (CE > 0 ? CE : 1)>(dst + HOFF_B + lane * DPL * CE, cld + coff(c) * CE);
    };
    int aLo = max(0, -u0), aHi = min(h - 1, w - 1 - (u0 + NCW - 1));   // rows in which every column is active
    auto issue_row = [&](int yl) {
        if (yl >= aLo && yl <= aHi) {
#pragma unroll
            for (int c = 0; c < NCW; ++c) issue_px(c, yl);
        } else {
#pragma unroll
            for (int c = 0; c < NCW; ++c)
                if (col_active(c, yl)) issue_px(c, yl);
        }
        if (CEN && yl <= y_out) {   // (rows after the last one would leave the image: nothing to fetch)
            const unsigned dst = cen0 + ((unsigned)yl & (PFS - 1)) * CEN_B + lane * 4;
#pragma unroll
            for (int k = 0; k < DPL; ++k) cp_async_bytes<4>(dst + k * 128, reinterpret_cast<const unsigned*>(cenR + 32 * k));
            cp_async_bytes<4>(cen0 + ((unsigned)yl & (PFS - 1)) * CEN_B + (DP + (lane & 15)) * 4, reinterpret_cast<const unsigned*>(cenX));
        }
        hld += estep; cld += estep * CE; cenR += cstep; cenX += cstep;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int k = 0; k < PFS - 1; ++k) issue_row(y_in + k);

    // Intensities: lane k keeps the pixel of row 32*blk + k of every column (one gather per 32 rows, the next
    // block already in flight), and a row takes its value with one shuffle -- nothing per row goes to memory.
    auto gather = [&](int c, int blk) {
        const int r = 32 * blk + lane;
        float v = 0.0f;
        if (col_active(c, r)) {
            const int xp = u0 + c + r;
            const int x = fwd ? xp : w - 1 - xp, yy = fwd ? r : h - 1 - r;
            v = __ldg(Ip + (size_t)yy * w + x);
        }
        return v;
    };
    int iblk = y_in >> 5;
    float icur[NCW], inxt[NCW];
#pragma unroll
    for (int c = 0; c < NCW; ++c) { icur[c] = gather(c, iblk); inxt[c] = gather(c, iblk + 1); }

    // ---- shared-memory addressing, resolved once per warp (32-bit shared-window addresses) ----
    // The three state rows of "the warp above" for image row y-1 come either from the in-band ring (slot
    // (y-1) & (S-1)) or, for the highest warp, from the upstream halo ring (slot (y-1) & (RING-1)):
    // base + ((y-1) & mask) * stride serves both, so the row loop has no role branches.
    const unsigned sh_hp = (unsigned)__cvta_generic_to_shared(s_hp), sh_sc = (unsigned)__cvta_generic_to_shared(s_sc);
    const unsigned sh_halo = (unsigned)__cvta_generic_to_shared(s_halo), sh_hsc = (unsigned)__cvta_generic_to_shared(s_hsc);
    const unsigned sh_edge = (unsigned)__cvta_generic_to_shared(s_edge), sh_esc = (unsigned)__cvta_generic_to_shared(s_esc);
    constexpr unsigned REC_B = DP * 4, SLOT_B = NWW * 3 * REC_B, SLOTSC_B = NWW * 32, HROW_B = 3 * REC_B, HSC_B = 32;
    const bool upIn = warp + 1 < NWW;
    const unsigned upBase = (upIn ? sh_hp + (warp + 1) * 3 * REC_B : sh_halo) + lane * DPL * 4;
    const unsigned upStride = upIn ? SLOT_B : HROW_B, upMask = upIn ? S - 1 : RING - 1;
    const unsigned upScBase = upIn ? sh_sc + (warp + 1) * 32 : sh_hsc, upScStride = upIn ? SLOTSC_B : HSC_B;
    const unsigned myBase = sh_hp + warp * 3 * REC_B + lane * DPL * 4, myScBase = sh_sc + warp * 32;
    const bool edge_out = warp == 0 && downstream;
    const unsigned eBase = sh_edge + lane * DPL * 4;
    // hand-off flags: rows < *flag of the producer are done
    // (32-bit shared-window addresses: the polling loops below are then LDS / ISETP / BRA / NANOSLEEP only)
    unsigned fUp = (unsigned)__cvta_generic_to_shared(upIn ? (const void*)(const_cast<int*>(prog) + warp + 1)
                                                                 : (const void*)const_cast<int*>(&ctl->halo_ready));
    unsigned fDn = warp >= 1 ? (unsigned)__cvta_generic_to_shared(const_cast<int*>(prog) + warp - 1)
                                   : fUp;   // no consumer: alias a flag that is already waited on
    unsigned fCp = (unsigned)__cvta_generic_to_shared(const_cast<int*>(&ctl->copied));
    fUp = __shfl_sync(0xffffffffu, fUp, 0); fDn = __shfl_sync(0xffffffffu, fDn, 0); fCp = __shfl_sync(0xffffffffu, fCp, 0);
    aLo = __shfl_sync(0xffffffffu, aLo, 0); aHi = __shfl_sync(0xffffffffu, aHi, 0);
    const int upCap = upIn ? 0x7fffffff : hend;   // an upstream band only publishes rows < hend
    const int wOff = S - 2;            // the consumer must have finished row y-S+1  <=>  prog >= y-S+2
    const int cOff = RING - 1;         // downstream ring slot free once rows < y-RING+1 were copied out

    // register state of the previous row: the diagonal path of every column (it stays in its column), the
    // vertical path of columns 1.. (input of the column below: c-1) and the anti-diagonal path of columns 2..
    // (input of column c-2); columns 0 and 1 hand theirs to the warp below through shared memory instead.
    float Dr[NCW][DPL], Vr[NCW][DPL], Ar[NCW][DPL];
    float lbD[NCW], pixD[NCW], lbVr[NCW], lbAr[NCW], pxr[NCW];
#pragma unroll
    for (int c = 0; c < NCW; ++c) {
#pragma unroll
        for (int j = 0; j < DPL; ++j) { Dr[c][j] = ROO_INF; Vr[c][j] = ROO_INF; Ar[c][j] = ROO_INF; }
        lbD[c] = 0.0f; pixD[c] = 0.0f; lbVr[c] = 0.0f; lbAr[c] = 0.0f; pxr[c] = 0.0f;
    }

    auto tick = [&](auto masked_tag, auto edge_tag, int y) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        constexpr bool EDGE = decltype(edge_tag)::value;
        const unsigned ym1 = (unsigned)(y - 1);
        const unsigned up = upBase + (ym1 & upMask) * upStride;
        const unsigned upsc = upScBase + (ym1 & upMask) * upScStride;
        const unsigned slot = (unsigned)y & (S - 1);
        const unsigned mine = myBase + slot * SLOT_B;
        const unsigned stg0 = pf0 + ((unsigned)y & (PFS - 1)) * STAGE_B;
        const unsigned er = eBase + ((unsigned)y & (RING - 1)) * HROW_B;
        const float4 scUp0 = lds_f4(upsc);        // upper warp's c0: {lastBest(V), lastBest(A), pix, -}
        const float4 scUp1 = lds_f4(upsc + 16);   // upper warp's c1: {-, lastBest(A), pix, -}
        float4 sc0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), sc1 = sc0;
        unsigned rwin[WIN], lwin[4 * ((NCW + 3) / 4)];
        if (CEN) {
            const unsigned cs = cen0 + ((unsigned)y & (PFS - 1)) * CEN_B;
            if constexpr (DPL >= 4) {
#pragma unroll
                for (int q = 0; q < WIN / 4; ++q)
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rwin[4 * q]), "=r"(rwin[4 * q + 1]), "=r"(rwin[4 * q + 2]), "=r"(rwin[4 * q + 3])
                                 : "r"(cs + (DPL * (31 - lane)) * 4 + 16 * q));
            } else {   // 32 / 64 disparities: the lane's window is not 16-byte aligned
#pragma unroll
                for (int q = 0; q < DPL + NCW - 1; ++q)
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rwin[q]) : "r"(cs + (DPL * (31 - lane) + q) * 4));
            }
#pragma unroll
            for (int q = 0; q < (NCW + 3) / 4; ++q)
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lwin[4 * q]), "=r"(lwin[4 * q + 1]), "=r"(lwin[4 * q + 2]), "=r"(lwin[4 * q + 3])
                             : "r"(cs + (DP + 8) * 4 + 16 * q));
        }
        // ascending c: column c reads the previous-row registers of c+1 and c+2 before those columns overwrite them
#pragma unroll
        for (int c = 0; c < NCW; ++c) {
            const bool act = !EDGE || col_active(c, y);
            if (act) {
                const int xp = u0 + c + y, x = fwd ? xp : w - 1 - xp;
                const float pix = __shfl_sync(0xffffffffu, icur[c], y & 31);
                float hv[DPL], ha[DPL], lbV, lbA, ppV, ppA;
                if (c < NCW - 1) {
#pragma unroll
                    for (int j = 0; j < DPL; ++j) hv[j] = Vr[c + 1][j];
                    lbV = lbVr[c + 1]; ppV = pxr[c + 1];
                } else {
                    lds_vec<DPL>(hv, up);
                    lbV = scUp0.x; ppV = scUp0.z;
                }
                if (c < NCW - 2) {
#pragma unroll
                    for (int j = 0; j < DPL; ++j) ha[j] = Ar[c + 2][j];
                    lbA = lbAr[c + 2]; ppA = pxr[c + 2];
                } else if (c == NCW - 2) {
                    lds_vec<DPL>(ha, up + REC_B);
                    lbA = scUp0.y; ppA = scUp0.z;
                } else {
                    lds_vec<DPL>(ha, up + 2 * REC_B);
                    lbA = scUp1.y; ppA = scUp1.z;
                }
                float cost[DPL];
                if (CEN) {
                    // column c is the (fwd ? c : NCW-1-c)-th pixel of the strip; R(x_c - d) for d = DPL*lane + j
                    const int cc = fwd ? c : NCW - 1 - c;
#pragma unroll
                    for (int j = 0; j < DPL; ++j) cost[j] = (float)__popc(lwin[cc] ^ rwin[DPL - 1 + cc - j]);
                } else {
                    RawCost<DPL, COST> rc;
                    rc.lds(stg0 + c * PFS * STAGE_B + HOFF_B + lane * DPL * CE);
#pragma unroll
                    for (int j = 0; j < DPL; ++j) cost[j] = rc.raw(j);
                }
                vg_pixel<DPL, MASKED, EDGE, FIRST, IEEE>(stg0 + c * PFS * STAGE_B, cost, lane, y, xp, x, w, M, P1, P2, cscale, pix,
                                                               hv, lbV, ppV, Dr[c], lbD[c], pixD[c], ha, lbA, ppA, hst + coff(c));
                if (c == 0) {
                    sts_vec<DPL>(mine, hv);
                    sts_vec<DPL>(mine + REC_B, ha);
                    if (edge_out) { sts_vec<DPL>(er, hv); sts_vec<DPL>(er + REC_B, ha); }
                    sc0 = make_float4(lbV, lbA, pix, 0.0f);
                } else {
                    if (c == 1) {
                        sts_vec<DPL>(mine + 2 * REC_B, ha);
                        if (edge_out) sts_vec<DPL>(er + 2 * REC_B, ha);
                        sc1 = make_float4(0.0f, lbA, pix, 0.0f);
                    } else {
#pragma unroll
                        for (int j = 0; j < DPL; ++j) Ar[c][j] = ha[j];
                        lbAr[c] = lbA;
                    }
#pragma unroll
                    for (int j = 0; j < DPL; ++j) Vr[c][j] = hv[j];
                    lbVr[c] = lbV; pxr[c] = pix;
                }
            }
        }
        if (lane == 0) {
            sts_f4(myScBase + slot * SLOTSC_B, sc0);
            sts_f4(myScBase + slot * SLOTSC_B + 16, sc1);
            if (edge_out) {
                const unsigned esc = sh_esc + ((unsigned)y & (RING - 1)) * HSC_B;
                sts_f4(esc, sc0);
                sts_f4(esc + 16, sc1);
            }
        }
    };

    // Row classes, resolved once per warp.  With x'_c = u0 + c + y (c0 is the leftmost column):
    //   interior rows [eLo, eHi]: y >= 1, x'_0 >= 1, x'_{NCW-1} <= w-2 and every column active -- no path starts or ends;
    //   unmasked rows [mLo, mHi]: the smallest true x of the NCW pixels (x'_0 forward, w-1-x'_{NCW-1} backward) is >= xf.
    int eLo = max(1, 1 - u0), eHi = min(w - 1 - NCW - u0, h - 1);
    int mLo = fwd ? xf - u0 : -0x3fffffff, mHi = fwd ? 0x3fffffff : (xf > w ? -0x3fffffff : w - NCW - u0 - xf);
    // through a shuffle: ptxas cannot rematerialise that, so the bounds stay in registers instead of being
    // recomputed (6-10 instructions each) in every row
    eLo = __shfl_sync(0xffffffffu, eLo, 0); eHi = __shfl_sync(0xffffffffu, eHi, 0);
    mLo = __shfl_sync(0xffffffffu, mLo, 0); mHi = __shfl_sync(0xffffffffu, mHi, 0);

    // No CTA-wide barrier: the warps of a band form a dataflow pipeline through shared memory.  Warp v may start
    // row y once warp v+1 has finished row y-1 (read-after-write) and warp v-1 has finished row y-S+1 (so the
    // ring slot of row y-S is free: write-after-read).
#ifdef VG_TIMING
    long long tcp = 0, tup = 0, tdn = 0, tbody = 0;
#endif
#pragma unroll 1
    for (int y = y_in; y <= y_out; ++y) {
#ifdef VG_TIMING
        long long t0 = clock64();
#endif
        issue_row(y + PFS - 1);
        if ((y & 31) == 0 && y != y_in) {   // next block of 32 rows of intensities: in flight since 32 rows ago
            ++iblk;
#pragma unroll
            for (int c = 0; c < NCW; ++c) { icur[c] = inxt[c]; inxt[c] = gather(c, iblk + 1); }
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(PFS - 1) : "memory");   // row y's stages have landed
#ifdef VG_TIMING
        long long t1 = clock64(); tcp += t1 - t0;
        spin_until(fUp, min(y, upCap));
        long long t2 = clock64(); tup += t2 - t1;
        spin_until(fDn, y - wOff);
        if (edge_out) spin_until(fCp, y - cOff);
        long long t3 = clock64(); tdn += t3 - t2;
#else
        // the flags only grow, so waiting for them one after the other is the same as waiting for all of them
        spin_until(fUp, min(y, upCap));
        spin_until(fDn, y - wOff);
        if (edge_out) spin_until(fCp, y - cOff);
#endif
        smem_order();

        const bool edge = y < eLo || y > eHi;
        const bool unmasked = y >= mLo && y <= mHi;
        if (edge) {
            if (unmasked) tick(std::false_type{}, std::true_type{}, y);
            else tick(std::true_type{}, std::true_type{}, y);
        } else {
            if (unmasked) tick(std::false_type{}, std::false_type{}, y);
            else tick(std::true_type{}, std::false_type{}, y);
        }
        hst += estep;
        __syncwarp();
        smem_order();
        if (lane == 0) publish_flag(prog + warp, (y == y_out) ? 0x7fffffff : y + 1);
#ifdef VG_TIMING
        tbody += clock64() - t3;
#endif
    }
#ifdef VG_TIMING
    if (lane == 0) {
        unsigned long long* dbg = reinterpret_cast<unsigned long long*>(a.ticket + 2);
        const int role = warp == NWW - 1 ? 2 : (warp == 0 ? 0 : 1);   // lowest / interior / highest warp
        atomicAdd(dbg + role * 5 + 0, (unsigned long long)tcp);
        atomicAdd(dbg + role * 5 + 1, (unsigned long long)tup);
        atomicAdd(dbg + role * 5 + 2, (unsigned long long)tdn);
        atomicAdd(dbg + role * 5 + 3, (unsigned long long)tbody);
        atomicAdd(dbg + role * 5 + 4, (unsigned long long)(y_out - y_in + 1));
    }
#endif
}

std::atomic<int> g_solo_geometry{1};
static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
        else n = 148;
    }
    return n;
}

// scratch sizing: the geometry with the narrower bands (more bands)
int vgroup_bands(int w, int h, int DP) {
    int nc = vg_cols_of_dp(DP, 0);
    for (int cls = 1; cls <= 3; ++cls) nc = vg_cols_of_dp(DP, cls) < nc ? vg_cols_of_dp(DP, cls) : nc;
    return cdiv(w + h - 1, nc);
}
size_t vgroup_edge_floats(int w, int h, int DP) { return (size_t)vgroup_bands(w, h, DP) * h * (3 * (size_t)DP + 8); }

// dynamic shared memory of one band (CTA)
template <int DPL, int COST, bool FIRST, int NWW, int NCW>
constexpr size_t vg_smem_bytes() {
    constexpr int DP = 32 * DPL, CE = RawCost<DPL, COST>::ELEM;
    return (size_t)(vg_s(DPL) * NWW * 3 * DP + vg_s(DPL) * NWW * 8 + 2 * (2 * vg_r(DPL) * (3 * DP + 8))) * sizeof(float) +
           sizeof(VCtl) + ((NWW * 4 + 15) / 16) * 16 + (size_t)NWW * NCW * vg_pfs(DPL, CE) * ((FIRST ? 0 : DP * 4) + DP * CE) +
           (COST == COST_CEN32 ? (size_t)NWW * vg_pfs(DPL, CE) * (DP + 16) * 4 : 0);
}

template <int DPL, int COST, bool FIRST, int CLS>
static int vgroup_launch3(VGroupArgs a, cudaStream_t st) {
    constexpr int NWW = vg_nww(DPL, CLS), NCW = vg_ncw(DPL, CLS);
    a.n_bands = cdiv(a.w + a.h - 1, NWW * NCW);
    constexpr size_t smem = vg_smem_bytes<DPL, COST, FIRST, NWW, NCW>();
    static_assert(smem <= 227 * 1024, "vertical-group kernel: shared memory budget of one sm_100 CTA exceeded");
    dim3 grid(a.n_bands * a.batch), block((NWW + 1) * 32);
#define ROO_VG(I)                                                                                             \
    do {                                                                                                      \
        auto kern = a.fwd ? sgm_vgroup_kernel<DPL, COST, FIRST, I, NWW, NCW, true>                            \
                          : sgm_vgroup_kernel<DPL, COST, FIRST, I, NWW, NCW, false>;                          \
        if (smem > 48 * 1024) {                                                                               \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                              \
        }                                                                                                     \
        kern<<<grid, block, smem, st>>>(a);                                                                   \
    } while (0)
    if (a.ieee) ROO_VG(true); else ROO_VG(false);
#undef ROO_VG
    count_launch();
    return launch_status();
}

template <int DPL, int COST>
static int vgroup_launch2(const VGroupArgs& a, bool first, cudaStream_t st) {
    constexpr int CEN_CLS = COST == COST_CEN32 ? 1 : 0;
    if constexpr (DPL >= 8 && COST == COST_CEN32) {
        // measured on B200 (profiles/r2_solo_geometry.json): 3840x2160 (250 bands) 16.95 -> 16.2 ms per pair, but 1920x1080
        // (125 bands, all resident at once) 5.93 -> 6.30 ms: only when the bands do not all fit on the SMs
        if (a.batch == 1 && g_solo_geometry.load(std::memory_order_relaxed) && cdiv(a.w + a.h - 1, vg_cols_of_dp(32 * DPL, 3)) > sm_count())
            return first ? vgroup_launch3<DPL, COST, true, 3>(a, st) : vgroup_launch3<DPL, COST, false, 3>(a, st);
    }
    if (!first) return vgroup_launch3<DPL, COST, false, CEN_CLS>(a, st);
    if constexpr (COST == COST_CEN32 && vg_nww(DPL, 2) != vg_nww(DPL, 1)) return vgroup_launch3<DPL, COST, true, 2>(a, st);
    else return vgroup_launch3<DPL, COST, true, CEN_CLS>(a, st);
}

// scratch: edge buffer of vgroup_edge_floats(w,h,DP) * batch floats and n_bands * batch ints of progress flags
int launch_vgroup(const SweepArgs& s, int fwd, float* edge, int* progress, cudaStream_t st) {
    VGroupArgs a{};
    a.H = s.H; a.h_pair = s.h_pair; a.C = s.C; a.c_pair = s.c_pair; a.img = s.img; a.img_pair = s.img_pair;
    a.cost_scale = s.cost_scale; a.w = s.w; a.h = s.h; a.maxDisp = s.maxDisp; a.batch = s.batch; a.P1 = s.P1; a.P2 = s.P2;
    a.fwd = fwd; a.ieee = s.ieee;
    a.cenL = s.cenL; a.cenR = s.cenR; a.cen_pair = s.cen_pair;
    a.n_bands = vgroup_bands(s.w, s.h, s.DP);
    const size_t hp_floats = (size_t)a.n_bands * s.h * 3 * s.DP;
    a.edge_hp = edge;
    a.edge_sc = edge + hp_floats * s.batch;
    a.progress = progress;
    a.ticket = progress + (size_t)a.n_bands * s.batch;   // one more int after the flags (scratch is sized for it)
    ROO_CUDA_TRY(cudaMemsetAsync(progress, 0, sizeof(int) * ((size_t)a.n_bands * s.batch + 1), st));
    const bool first = s.first != 0;
#define ROO_VG_DP(DPL)                                                              \
    switch (s.cost_kind) {                                                          \
        case COST_F32: return vgroup_launch2<DPL, COST_F32>(a, first, st);          \
        case COST_U8: return vgroup_launch2<DPL, COST_U8>(a, first, st);            \
        case COST_CEN32: return vgroup_launch2<DPL, COST_CEN32>(a, first, st);      \
        default: return ROO_ERR_INVALID_ARGUMENT;                                   \
    }
    switch (s.DP) {
        case 32: ROO_VG_DP(1)
        case 64: ROO_VG_DP(2)
        case 128: ROO_VG_DP(4)
        case 256: ROO_VG_DP(8)
        default: return ROO_ERR_UNSUPPORTED;
    }
#undef ROO_VG_DP
}

}  // namespace roo_b200
