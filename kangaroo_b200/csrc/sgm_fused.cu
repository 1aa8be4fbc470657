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
// columns u = x' - y'.  One warp owns one skewed column and walks it row by row, lane l holding disparities
// [l*DPL, (l+1)*DPL):
//   * the diagonal path (+1,+1) stays inside the warp: its state never leaves registers;
//   * the vertical path needs the state of column u+1, the anti-diagonal path that of column u+2, both of
//     the previous row -> every dependency points towards HIGHER u.  Inside a CTA (a band of NW columns)
//     the states go through shared memory (double-buffered by row parity, one __syncthreads per row);
//     between CTAs they flow one way only, from band b-1 to band b, through a small L2-resident edge
//     buffer guarded by a monotonic progress flag.  One-way dependencies make the bands a pipeline, not a
//     ping-pong: a band never waits for a band that waits for it, and lower block indices (scheduled
//     first) never wait for higher ones.
//   * a dedicated communication warp per CTA polls the upstream flag, stages the upstream edge rows into
//     shared memory and publishes this band's flag, so the compute warps never touch the global flags.
//   * every compute warp owns TWO adjacent skewed columns (B = u, A = u+1) and advances both by one row per
//     tick: the vertical path of B continues from A's state of the previous tick (registers), so only three
//     state rows per tick go through shared memory, and the per-tick bookkeeping (hand-off flags, prefetch
//     issue, addressing) is paid once for two pixels.
// HBM traffic of the pass: read H (unless first) + read cost + write H -- the same as ONE single-path sweep.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "sgm_step.cuh"

namespace roo_b200 {

// rows of prefetch per column, staged in shared memory by cp.async (LDGSTS): under load a DRAM access takes
// ~3000 SM cycles on B200, so a band needs ~60-80 KB in flight per SM to stream at HBM speed -- far more
// than a register ring can hold, and without unrolling the row loop.  (227 KB of shared memory per CTA.)
__host__ __device__ constexpr int vg_pfs(int DPL, int CE) { return DPL >= 8 ? 2 : (DPL == 4 ? (CE == 4 ? 2 : 4) : 8); }
// compute warps per band (each owns two skewed columns) + 1 communication warp
__host__ __device__ constexpr int vg_nww(int DPL) { return DPL >= 8 ? 8 : 16; }
inline int vg_cols_of_dp(int DP) { return 2 * vg_nww(DP / 32); }

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

#ifndef VG_SPIN_NS
#define VG_SPIN_NS 30
#endif
// poll a monotonically growing shared-memory flag until it reaches `need`
__device__ __forceinline__ void spin_until(unsigned flag_addr, int need) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(flag_addr) : "memory");
    while (v < need) {
        __nanosleep(VG_SPIN_NS);
        asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(flag_addr) : "memory");
    }
}

// smem control words
struct VCtl { volatile int halo_ready; volatile int copied; int pad[2]; };


__host__ __device__ constexpr int vg_r(int DPL) { return 4; }   // max rows per hand-off batch between bands (ring = 2x)
constexpr int VG_S = 4;   // depth (rows) of the in-band state ring in shared memory

// One pixel of the three paths.  V/D/A = vertical / diagonal / anti-diagonal.  hpV, hpD, hpA: previous pixel's
// state rows on entry, this pixel's on exit.  Handles path starts when EDGE.
template <int DPL, int COST, bool MASKED, bool EDGE, bool FIRST, bool IEEE>
__device__ __forceinline__ void vg_pixel(unsigned stg, int lane, int y, int xp, int x, int w, int M, float P1, float P2,
                                         float cscale, float (&hpV)[DPL], float& lbV, float ppV,
                                         float (&hpD)[DPL], float& lbD, float& pixD,
                                         float (&hpA)[DPL], float& lbA, float ppA, float& pix_out, float* hst) {
    constexpr int DP = 32 * DPL;
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    const int lim = MASKED ? min(M, x + 1) - lane * DPL : 0;
    float hin[DPL], H3[DPL], cost[DPL];
    if (!FIRST) lds_vec<DPL>(hin, stg + lane * DPL * 4);
    RawCost<DPL, COST> rc;
    rc.lds(stg + DP * 4 + lane * DPL * CE);
    float pix;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pix) : "r"(stg + DP * 4 + DP * CE));
#pragma unroll
    for (int j = 0; j < DPL; ++j) cost[j] = rc.get(j, cscale);
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
                                        cost, hin, P1, lim, lane, H3, bV, bD, bA);
    if (EDGE) { if (sV) bV = 0.0f; if (sD) bD = 0.0f; if (sA) bA = 0.0f; }
    lbV = bV; lbD = bD; lbA = bA;
    pixD = pix;
    pix_out = pix;
    store_f<DPL>(hst, H3);
}

template <int DPL, int COST, bool FIRST, bool IEEE, int NWW>
__global__ void __launch_bounds__((NWW + 1) * 32, 1)
sgm_vgroup_kernel(const VGroupArgs a) {
    constexpr int DP = 32 * DPL;
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    constexpr int NC = 2 * NWW;                       // skewed columns per band
    constexpr int PFS = vg_pfs(DPL, CE);
    constexpr int R = vg_r(DPL), RING = 2 * R, S = VG_S;
    constexpr int STAGE_B = DP * 4 + DP * CE + 16;   // one prefetched pixel: aggregate row, cost row, intensity
    extern __shared__ __align__(16) float smem[];
    // state rows of one warp and one image row: rec0 = B.vertical, rec1 = B.anti-diagonal, rec2 = A.anti-diagonal;
    // scalars {B.lastBest(V), B.lastBest(A), B.pix, -, -, A.lastBest(A), A.pix, -}: the same record layout is
    // used by the band-to-band edge rows, so a warp reads "the three rows of whoever is above me" with one formula
    float* s_hp = smem;                                // [S rows][NWW][3][DP]
    float* s_sc = s_hp + S * NWW * 3 * DP;             // [S rows][NWW][8]
    float* s_halo = s_sc + S * NWW * 8;                // [RING rows][3][DP]  upstream band's lowest warp (row ring)
    float* s_hsc = s_halo + RING * 3 * DP;             // [RING][8]
    float* s_edge = s_hsc + RING * 8;                  // [RING rows][3][DP]  this band's lowest warp, for downstream
    float* s_esc = s_edge + RING * 3 * DP;             // [RING][8]
    VCtl* ctl = reinterpret_cast<VCtl*>(s_esc + RING * 8);
    volatile int* prog = reinterpret_cast<volatile int*>(ctl + 1);   // [NWW] rows < prog[v] of warp v are done
    char* s_pf = reinterpret_cast<char*>(ctl + 1) + ((NWW * 4 + 15) / 16) * 16;   // [NWW][2 cols][PFS][STAGE_B]

    // warp index through a shuffle: ptxas then knows it is warp-uniform, and every branch on it is a uniform branch
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // pair fastest: the resident window of CTAs then holds the same few bands of EVERY pair, so the
    // band-to-band pipeline of each pair has only a short ramp
    const int pair = blockIdx.x % a.batch, band = blockIdx.x / a.batch;
    const int w = a.w, h = a.h, M = a.maxDisp;
    const bool fwd = a.fwd != 0;
    const float P1 = a.P1, P2 = a.P2, cscale = a.cost_scale;

    const int ulo = w - (band + 1) * NC;                 // lowest skewed column of this band
    const int ymin = max(0, -(ulo + NC - 1));
    const int ymax = min(h - 1, w - 1 - ulo);
    // upstream band (higher u) and the rows of it this band consumes: row y-1 for every own row y >= 1
    const int pulo = ulo + NC;
    const int pymin = max(0, -(pulo + NC - 1));
    const int pymax = band > 0 ? min(h - 1, w - 1 - pulo) : -1;
    const int hbeg = max(pymin, ymin - 1), hend = min(pymax, ymax - 1) + 1;   // [hbeg, hend) upstream rows to stage
    const bool downstream = band + 1 < a.n_bands;

    float* e_hp = a.edge_hp + ((size_t)pair * a.n_bands + band) * (size_t)h * 3 * DP;   // this band's outgoing rows
    float* e_sc = a.edge_sc + ((size_t)pair * a.n_bands + band) * (size_t)h * 8;
    int* my_flag = a.progress + (size_t)pair * a.n_bands + band;

    // this warp's two skewed columns and their active rows (x' = u + y' in [0, w))
    const int uB = ulo + 2 * warp, uA = uB + 1;
    const int yinA = max(0, -uA), youtA = min(h - 1, w - 1 - uA);
    const int yinB = max(0, -uB), youtB = min(h - 1, w - 1 - uB);
    const int y_in = min(yinA, yinB), y_out = max(youtA, youtB);   // A enters first, B leaves last
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
            if (!progress) __nanosleep(200);
        }
        return;
    }
    if (!any) return;

    // -------------------------------------------------------------------- compute warps
    const int xf = (M == DP) ? DP - 1 : 0x3fffffff;               // all lanes in range iff true x >= xf
    const int d0 = lane * DPL;
    const ptrdiff_t pstep = fwd ? (ptrdiff_t)(w + 1) : -(ptrdiff_t)(w + 1);   // one row down the travel direction
    const ptrdiff_t estep = pstep * DP;

    // per-column cursors at the column's first active pixel
    auto first_px = [&](int u, int yin, size_t& e0, size_t& p0) {
        const int xp0 = u + yin;
        const int x0 = fwd ? xp0 : w - 1 - xp0, y0 = fwd ? yin : h - 1 - yin;
        p0 = (size_t)y0 * w + x0;
        e0 = p0 * DP + d0;
    };
    size_t e0A = 0, p0A = 0, e0B = 0, p0B = 0;
    if (yinA <= youtA) first_px(uA, yinA, e0A, p0A);
    if (yinB <= youtB) first_px(uB, yinB, e0B, p0B);
    float* const Hp = a.H + (size_t)pair * a.h_pair;
    const char* const Cp = (const char*)a.C + (size_t)pair * a.c_pair * CE;
    const float* const Ip = a.img + (size_t)pair * a.img_pair;
    float* hstA = Hp + e0A; const float* hldA = hstA; const char* cldA = Cp + e0A * CE; const float* ildA = Ip + p0A;
    float* hstB = Hp + e0B; const float* hldB = hstB; const char* cldB = Cp + e0B * CE; const float* ildB = Ip + p0B;

    // Prefetch: row y+PFS-1 of both columns is copied global -> shared (asynchronously, no registers) while row y
    // is computed.  Every lane copies and later reads its own bytes; only the intensity (lane 0) needs a __syncwarp.
    const unsigned pfA = (unsigned)__cvta_generic_to_shared(s_pf) + (warp * 2) * PFS * STAGE_B;
    const unsigned pfB = pfA + PFS * STAGE_B;
    auto issue_px = [&](unsigned base, int yl, const float*& hld, const char*& cld, const float*& ild) {
        const unsigned dst = base + ((unsigned)yl & (PFS - 1)) * STAGE_B;
        if (!FIRST) cp_async_bytes<DPL * 4>(dst + lane * DPL * 4, hld);
        cp_async_bytes<DPL * CE>(dst + DP * 4 + lane * DPL * CE, cld);
        if (lane == 0) cp_async_bytes<4>(dst + DP * 4 + DP * CE, ild);
        hld += estep; cld += estep * CE; ild += pstep;
    };
    auto issue_row = [&](int yl) {
        if (yl >= yinA && yl <= youtA) issue_px(pfA, yl, hldA, cldA, ildA);
        if (yl >= yinB && yl <= youtB) issue_px(pfB, yl, hldB, cldB, ildB);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int k = 0; k < PFS - 1; ++k) issue_row(y_in + k);

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
    const unsigned fUp = (unsigned)__cvta_generic_to_shared(upIn ? (const void*)(const_cast<int*>(prog) + warp + 1)
                                                                 : (const void*)const_cast<int*>(&ctl->halo_ready));
    const unsigned fDn = warp >= 1 ? (unsigned)__cvta_generic_to_shared(const_cast<int*>(prog) + warp - 1)
                                   : fUp;   // no consumer: alias a flag that is already waited on
    const unsigned fCp = (unsigned)__cvta_generic_to_shared(const_cast<int*>(&ctl->copied));
    const int upCap = upIn ? 0x7fffffff : hend;   // an upstream band only publishes rows < hend
    const int wOff = S - 2;            // the consumer must have finished row y-S+1  <=>  prog >= y-S+2
    const int cOff = RING - 1;         // downstream ring slot free once rows < y-RING+1 were copied out

    // register state: diagonal path of both columns; A's vertical state of the previous row (input of B's vertical path)
    float dA[DPL], dB[DPL], vA[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) { dA[j] = ROO_INF; dB[j] = ROO_INF; vA[j] = ROO_INF; }
    float lbDA = 0.0f, pixDA = 0.0f, lbDB = 0.0f, pixDB = 0.0f, lbVA = 0.0f, pixA = 0.0f;

    auto tick = [&](auto masked_tag, auto edge_tag, int y) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        constexpr bool EDGE = decltype(edge_tag)::value;
        const unsigned ym1 = (unsigned)(y - 1);
        const unsigned up = upBase + (ym1 & upMask) * upStride;
        const unsigned upsc = upScBase + (ym1 & upMask) * upScStride;
        const unsigned slot = (unsigned)y & (S - 1);
        const unsigned mine = myBase + slot * SLOT_B;
        const bool actA = !EDGE || (y >= yinA && y <= youtA);
        const bool actB = !EDGE || (y >= yinB && y <= youtB);
        const float4 scUpB = lds_f4(upsc);        // upper warp's B: {lastBest(V), lastBest(A), pix, -}
        const float4 scUpA = lds_f4(upsc + 16);   // upper warp's A: {-, lastBest(A), pix, -}
        float bBV = 0.0f, bBA = 0.0f, pixB = 0.0f, bAA = 0.0f;
        // ---- column B (lower): vertical continues from A's previous row (registers), anti-diagonal from upper B
        if (actB) {
            const int xp = uB + y, x = fwd ? xp : w - 1 - xp;
            float hv[DPL], ha[DPL];
#pragma unroll
            for (int j = 0; j < DPL; ++j) hv[j] = vA[j];
            lds_vec<DPL>(ha, up + REC_B);
            float lbV = lbVA, lbA = scUpB.y;
            vg_pixel<DPL, COST, MASKED, EDGE, FIRST, IEEE>(pfB + ((unsigned)y & (PFS - 1)) * STAGE_B, lane, y, xp, x, w, M, P1, P2,
                                                           cscale, hv, lbV, pixA, dB, lbDB, pixDB, ha, lbA, scUpB.z, pixB, hstB);
            hstB += estep;
            sts_vec<DPL>(mine, hv);
            sts_vec<DPL>(mine + REC_B, ha);
            if (edge_out) {
                const unsigned er = eBase + ((unsigned)y & (RING - 1)) * HROW_B;
                sts_vec<DPL>(er, hv);
                sts_vec<DPL>(er + REC_B, ha);
            }
            bBV = lbV; bBA = lbA;
        }
        // ---- column A (upper): vertical from upper B, anti-diagonal from upper A; its vertical state stays in registers
        if (actA) {
            const int xp = uA + y, x = fwd ? xp : w - 1 - xp;
            float ha[DPL];
            lds_vec<DPL>(vA, up);
            lds_vec<DPL>(ha, up + 2 * REC_B);
            lbVA = scUpB.x;
            float lbA = scUpA.y;
            vg_pixel<DPL, COST, MASKED, EDGE, FIRST, IEEE>(pfA + ((unsigned)y & (PFS - 1)) * STAGE_B, lane, y, xp, x, w, M, P1, P2,
                                                           cscale, vA, lbVA, scUpB.z, dA, lbDA, pixDA, ha, lbA, scUpA.z, pixA, hstA);
            hstA += estep;
            sts_vec<DPL>(mine + 2 * REC_B, ha);
            if (edge_out) sts_vec<DPL>(eBase + ((unsigned)y & (RING - 1)) * HROW_B + 2 * REC_B, ha);
            bAA = lbA;
        }
        if (lane == 0) {
            sts_f4(myScBase + slot * SLOTSC_B, make_float4(bBV, bBA, pixB, 0.0f));
            sts_f4(myScBase + slot * SLOTSC_B + 16, make_float4(0.0f, bAA, pixA, 0.0f));
            if (edge_out) {
                const unsigned esc = sh_esc + ((unsigned)y & (RING - 1)) * HSC_B;
                sts_f4(esc, make_float4(bBV, bBA, pixB, 0.0f));
                sts_f4(esc + 16, make_float4(0.0f, bAA, pixA, 0.0f));
            }
        }
    };

    // Row classes, resolved once per warp.  With x'_B = uB + y (B is the leftmost of the two columns, x'_A = x'_B + 1):
    //   interior rows [eLo, eHi]: y >= 1, x'_B >= 1, x'_A <= w-2 and both columns active -- no path starts or ends;
    //   unmasked rows [mLo, mHi]: the smaller true x of the two pixels (x'_B forward, w-2-x'_B backward) is >= xf.
    int eLo = max(max(1, 1 - uB), max(yinA, yinB)), eHi = min(w - 3 - uB, min(youtA, youtB));
    int mLo = fwd ? xf - uB : -0x3fffffff, mHi = fwd ? 0x3fffffff : (xf > w ? -0x3fffffff : w - 2 - uB - xf);
    asm volatile("" : "+r"(eLo), "+r"(eHi), "+r"(mLo), "+r"(mHi));   // keep them in registers: no per-row rematerialisation

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
        asm volatile("cp.async.wait_group %0;" ::"n"(PFS - 1) : "memory");   // row y's stages have landed
        __syncwarp();
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
        __syncwarp();
        smem_order();
        if (lane == 0) prog[warp] = (y == y_out) ? 0x7fffffff : y + 1;
#ifdef VG_TIMING
        tbody += clock64() - t3;
#endif
    }
#ifdef VG_TIMING
    if (lane == 0) {
        unsigned long long* dbg = reinterpret_cast<unsigned long long*>(a.progress + (size_t)a.n_bands * a.batch);
        const int role = warp == NWW - 1 ? 2 : (warp == 0 ? 0 : 1);   // lowest / interior / highest warp
        atomicAdd(dbg + role * 5 + 0, (unsigned long long)tcp);
        atomicAdd(dbg + role * 5 + 1, (unsigned long long)tup);
        atomicAdd(dbg + role * 5 + 2, (unsigned long long)tdn);
        atomicAdd(dbg + role * 5 + 3, (unsigned long long)tbody);
        atomicAdd(dbg + role * 5 + 4, (unsigned long long)(y_out - y_in + 1));
    }
#endif
}

int vgroup_bands(int w, int h, int DP) { return cdiv(w + h - 1, vg_cols_of_dp(DP)); }
size_t vgroup_edge_floats(int w, int h, int DP) { return (size_t)vgroup_bands(w, h, DP) * h * (3 * (size_t)DP + 8); }

template <int DPL, int COST>
static int vgroup_launch2(const VGroupArgs& a, bool first, cudaStream_t st) {
    constexpr int DP = 32 * DPL;
    constexpr int NWW = vg_nww(DPL);
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    const size_t smem = (size_t)(VG_S * NWW * 3 * DP + VG_S * NWW * 8 + 2 * (2 * vg_r(DPL) * (3 * DP + 8))) * sizeof(float) +
                        sizeof(VCtl) + ((NWW * 4 + 15) / 16) * 16 + (size_t)NWW * 2 * vg_pfs(DPL, CE) * (DP * 4 + DP * CE + 16);
    dim3 grid(a.n_bands * a.batch), block((NWW + 1) * 32);
    const bool ieee = g_ieee_div.load() != 0;
#define ROO_VG(F, I)                                                                                          \
    do {                                                                                                      \
        auto kern = sgm_vgroup_kernel<DPL, COST, F, I, NWW>;                                                  \
        if (smem > 48 * 1024) {                                                                               \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                              \
        }                                                                                                     \
        kern<<<grid, block, smem, st>>>(a);                                                                   \
    } while (0)
    if (first) { if (ieee) ROO_VG(true, true); else ROO_VG(true, false); }
    else { if (ieee) ROO_VG(false, true); else ROO_VG(false, false); }
#undef ROO_VG
    count_launch();
    return launch_status();
}

// scratch: edge buffer of vgroup_edge_floats(w,h,DP) * batch floats and n_bands * batch ints of progress flags
int launch_vgroup(const SweepArgs& s, int fwd, float* edge, int* progress, cudaStream_t st) {
    VGroupArgs a{};
    a.H = s.H; a.h_pair = s.h_pair; a.C = s.C; a.c_pair = s.c_pair; a.img = s.img; a.img_pair = s.img_pair;
    a.cost_scale = s.cost_scale; a.w = s.w; a.h = s.h; a.maxDisp = s.maxDisp; a.batch = s.batch; a.P1 = s.P1; a.P2 = s.P2;
    a.fwd = fwd;
    a.n_bands = vgroup_bands(s.w, s.h, s.DP);
    const size_t hp_floats = (size_t)a.n_bands * s.h * 3 * s.DP;
    a.edge_hp = edge;
    a.edge_sc = edge + hp_floats * s.batch;
    a.progress = progress;
    ROO_CUDA_TRY(cudaMemsetAsync(progress, 0, sizeof(int) * (size_t)a.n_bands * s.batch, st));
    const bool first = s.first != 0;
    const bool f32 = s.cost_kind == COST_F32;
    switch (s.DP) {
        case 32: return f32 ? vgroup_launch2<1, COST_F32>(a, first, st) : vgroup_launch2<1, COST_U8>(a, first, st);
        case 64: return f32 ? vgroup_launch2<2, COST_F32>(a, first, st) : vgroup_launch2<2, COST_U8>(a, first, st);
        case 128: return f32 ? vgroup_launch2<4, COST_F32>(a, first, st) : vgroup_launch2<4, COST_U8>(a, first, st);
        case 256: return f32 ? vgroup_launch2<8, COST_F32>(a, first, st) : vgroup_launch2<8, COST_U8>(a, first, st);
        default: return ROO_ERR_UNSUPPORTED;
    }
}

}  // namespace roo_b200
