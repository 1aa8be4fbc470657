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
//     shared memory and publishes this band's flag, so the NW compute warps never touch the flags.
// HBM traffic of the pass: read H (unless first) + read cost + write H -- the same as ONE single-path sweep.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "sgm_step.cuh"

namespace roo_b200 {

// rows of prefetch, staged in shared memory by cp.async (LDGSTS): under load a DRAM access takes ~3000 SM
// cycles on B200, so a band needs ~60-80 KB in flight per SM to stream at HBM speed -- far more than a
// register ring can hold, and without unrolling the row loop
__host__ __device__ constexpr int vg_pfs(int DPL, int CE) { return DPL >= 8 ? (CE == 4 ? 2 : 4) : (DPL == 4 && CE == 4 ? 4 : 8); }   // (227 KB of shared memory per CTA)
// skewed columns (compute warps) per band: 16 (+1 communication warp) leaves 120 registers per thread, enough
// for DPL <= 4; the 256-disparity variant keeps twice the state per lane and runs 12 + 1 warps
constexpr int vg_nw(int DPL) { return DPL >= 8 ? 12 : 16; }
inline int vg_nw_of_dp(int DP) { return vg_nw(DP / 32); }

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

template <int DPL, int COST>
struct VStage {
    float hin[DPL];
    RawCost<DPL, COST> c;
    float pix;
};

// Ordering of shared-memory accesses between warps of one CTA.  Data and flag both live in shared memory
// and every access is issued through the same in-order LSU pipeline of the SM, so program order (enforced
// for the compiler by the volatile flag accesses and this barrier) is enough; a MEMBAR here would also wait
// for the warp's outstanding GLOBAL prefetch loads and serialise every row on DRAM latency.
__device__ __forceinline__ void smem_order() { asm volatile("" ::: "memory"); }

// smem control words
struct VCtl { volatile int halo_ready; volatile int copied; int pad[2]; };

__host__ __device__ constexpr int vg_r(int DPL) { return DPL >= 8 ? 4 : 8; }   // max rows per hand-off batch between bands (ring = 2x)
#ifndef VG_SPIN_NS
#define VG_SPIN_NS 30
#endif
constexpr int VG_S = 4;   // depth (rows) of the in-band state ring in shared memory

template <int DPL, int COST, bool FIRST, bool IEEE, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
sgm_vgroup_kernel(const VGroupArgs a) {
    constexpr int DP = 32 * DPL;
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    constexpr int PFS = vg_pfs(DPL, CE);
    constexpr int R = vg_r(DPL), RING = 2 * R, S = VG_S;
    constexpr int STAGE_B = DP * 4 + DP * CE + 16;   // one prefetched pixel: aggregate row, cost row, intensity
    extern __shared__ __align__(16) float smem[];
    float* s_hp = smem;                                // [S rows][NW][2 paths][DP]  in-band states (row ring)
    float* s_sc = s_hp + S * NW * 2 * DP;              // [S rows][NW][4]: lastBest(vertical), lastBest(anti-diag), pix, -
    float* s_halo = s_sc + S * NW * 4;                 // [RING rows][3][DP]  upstream band's columns 0,1 (row ring)
    float* s_hsc = s_halo + RING * 3 * DP;             // [RING][8]
    float* s_edge = s_hsc + RING * 8;                  // [RING rows][3][DP]  this band's columns 0,1 for downstream
    float* s_esc = s_edge + RING * 3 * DP;             // [RING][8]
    VCtl* ctl = reinterpret_cast<VCtl*>(s_esc + RING * 8);
    volatile int* prog = reinterpret_cast<volatile int*>(ctl + 1);   // [NW] rows < prog[j] of column j are done
    char* s_pf = reinterpret_cast<char*>(ctl + 1) + ((NW * 4 + 15) / 16) * 16;   // [NW][PFS][STAGE_B] prefetch stages

    // warp index through a shuffle: ptxas then knows it is warp-uniform, and every branch on it (roles, masks,
    // path starts) is a uniform branch without divergence bookkeeping around the shuffles / redux below
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // pair fastest: the resident window of CTAs then holds the same few bands of EVERY pair, so the
    // band-to-band pipeline of each pair has only a short ramp
    const int pair = blockIdx.x % a.batch, band = blockIdx.x / a.batch;
    const int w = a.w, h = a.h, M = a.maxDisp;
    const bool fwd = a.fwd != 0;
    const float P1 = a.P1, P2 = a.P2, cscale = a.cost_scale;

    const int ulo = w - (band + 1) * NW;                 // lowest skewed column of this band
    const int ymin = max(0, -(ulo + NW - 1));
    const int ymax = min(h - 1, w - 1 - ulo);
    // upstream band (higher u) and the rows of it this band consumes: row y-1 for every own row y >= 1
    const int pulo = ulo + NW;
    const int pymin = max(0, -(pulo + NW - 1));
    const int pymax = band > 0 ? min(h - 1, w - 1 - pulo) : -1;
    const int hbeg = max(pymin, ymin - 1), hend = min(pymax, ymax - 1) + 1;   // [hbeg, hend) upstream rows to stage
    const bool downstream = band + 1 < a.n_bands;

    float* e_hp = a.edge_hp + ((size_t)pair * a.n_bands + band) * (size_t)h * 3 * DP;   // this band's outgoing rows
    float* e_sc = a.edge_sc + ((size_t)pair * a.n_bands + band) * (size_t)h * 8;
    int* my_flag = a.progress + (size_t)pair * a.n_bands + band;

    // active rows of skewed column u: x' = u + y' in [0, w)
    const int u = ulo + warp;
    const int y_in = max(0, -u), y_out = min(h - 1, w - 1 - u);
    const bool any = warp < NW && y_in <= y_out;
    if (threadIdx.x == 0) { ctl->halo_ready = hbeg; ctl->copied = ymin; }
    if (warp < NW && lane == 0) prog[warp] = any ? y_in : 0x7fffffff;   // rows before y_in never happen
    __syncthreads();

    if (warp == NW) {
        // ---------------------------------------------------------------- communication warp
        const float* p_hp = e_hp - (size_t)h * 3 * DP;   // upstream band's rows
        const float* p_sc = e_sc - (size_t)h * 8;
        const int* p_flag = my_flag - 1;
        int seen = 0, hr = hbeg, cp = ymin;
        const int cp_end = downstream ? ymax + 1 : ymin;  // nothing to publish for the last band
        while (hr < hend || cp < cp_end) {
            bool progress = false;
            // ---- stage a chunk of upstream rows into the halo ring (readers: the last two columns)
            if (hr < hend) {
                // adaptive batch: whatever the upstream band has published and the ring can take (1..R rows) --
                // the hand-off latency is one turn of this loop, not the time to fill a fixed chunk.  The paths
                // that run across the bands (anti-diagonal: a new band every NW/2 rows) are a serial chain of such
                // hand-offs, so this latency, not the copy bandwidth, bounds a single pair's pass.
                const int rdh = min(prog[NW - 1], prog[NW - 2]);
                if (seen <= hr) {
                    if (lane == 0) seen = ld_acquire_gpu(p_flag);
                    seen = __shfl_sync(0xffffffffu, seen, 0);
                }
                // ring slot of row y was last used by row y-RING, read while computing row y-RING+1
                const int n = min(min(R, hend - hr), min(seen - hr, rdh + RING - 1 - hr));
                if (n > 0) {
                    {
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
            }
            // ---- publish finished rows of this band's columns 0,1 (writers: the first two columns)
            if (cp < cp_end) {
                const int rd = min(min(prog[0], prog[1]), ymax + 1);
                if (rd > cp) {   // publish every finished row at once (see the latency note above)
                    __threadfence_block();
                    const int rd_lim = min(rd, cp + RING);
                    for (int y = cp; y < rd_lim; ++y) {
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
    const int d0 = lane * DPL;
    const int xf = (M == DP) ? DP - 1 : 0x3fffffff;               // all lanes in range iff true x >= xf

    // cursors at the first active pixel; one row down the travel direction = +-(w+1) pixels
    const int xp0 = u + y_in;
    const int x0 = fwd ? xp0 : w - 1 - xp0, y0 = fwd ? y_in : h - 1 - y_in;
    const ptrdiff_t pstep = fwd ? (ptrdiff_t)(w + 1) : -(ptrdiff_t)(w + 1);
    const ptrdiff_t estep = pstep * DP;
    const size_t e0 = ((size_t)y0 * w + x0) * DP + d0;
    float* hst = a.H + (size_t)pair * a.h_pair + e0;
    const float* hld = hst;
    const char* cld = (const char*)a.C + ((size_t)pair * a.c_pair + e0) * CE;
    const float* ild = a.img + (size_t)pair * a.img_pair + (size_t)y0 * w + x0;

    // Prefetch: row y+PFS-1 is copied global -> shared (asynchronously, no registers) while row y is computed.
    // Every lane copies and later reads its own bytes; only the intensity (lane 0) needs a __syncwarp.
    const unsigned pfBase = (unsigned)__cvta_generic_to_shared(s_pf) + warp * PFS * STAGE_B;
    auto issue_row = [&](int yl) {
        if (yl <= y_out) {
            const unsigned dst = pfBase + ((unsigned)yl & (PFS - 1)) * STAGE_B;
            if (!FIRST) cp_async_bytes<DPL * 4>(dst + lane * DPL * 4, hld);
            cp_async_bytes<DPL * CE>(dst + DP * 4 + lane * DPL * CE, cld);
            if (lane == 0) cp_async_bytes<4>(dst + DP * 4 + DP * CE, ild);
            hld += estep; cld += estep * CE; ild += pstep;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int k = 0; k < PFS - 1; ++k) issue_row(y_in + k);

    // ---- shared-memory addressing, resolved once per warp (32-bit shared-window addresses) ----
    // State rows of the previous image row come either from the in-band ring (slot (y-1) & (S-1), stride one
    // slot) or, for the two highest columns, from the upstream halo ring (slot (y-1) & (RING-1), stride one
    // halo row): the same  base + ((y-1) & mask) * stride  serves both, so the row loop has no role branches.
    const unsigned sh_hp = (unsigned)__cvta_generic_to_shared(s_hp), sh_sc = (unsigned)__cvta_generic_to_shared(s_sc);
    const unsigned sh_halo = (unsigned)__cvta_generic_to_shared(s_halo), sh_hsc = (unsigned)__cvta_generic_to_shared(s_hsc);
    const unsigned sh_edge = (unsigned)__cvta_generic_to_shared(s_edge), sh_esc = (unsigned)__cvta_generic_to_shared(s_esc);
    constexpr unsigned SLOT_B = NW * 2 * DP * 4, SLOTSC_B = NW * 16, HROW_B = 3 * DP * 4, HSC_B = 32;
    const bool vIn = warp + 1 < NW, aIn = warp + 2 < NW;
    const unsigned vBase = vIn ? sh_hp + (warp + 1) * 2 * DP * 4 + lane * DPL * 4 : sh_halo + lane * DPL * 4;
    const unsigned vStride = vIn ? SLOT_B : HROW_B, vMask = vIn ? S - 1 : RING - 1;
    const unsigned vScBase = vIn ? sh_sc + (warp + 1) * 16 : sh_hsc, vScStride = vIn ? SLOTSC_B : HSC_B;
    const unsigned aBase = aIn ? sh_hp + ((warp + 2) * 2 + 1) * DP * 4 + lane * DPL * 4
                               : sh_halo + (warp + 2 == NW ? 1 : 2) * DP * 4 + lane * DPL * 4;
    const unsigned aStride = aIn ? SLOT_B : HROW_B, aMask = aIn ? S - 1 : RING - 1;
    const unsigned aScBase = aIn ? sh_sc + (warp + 2) * 16 : sh_hsc + (warp + 2 == NW ? 0 : 16);
    const unsigned aScStride = aIn ? SLOTSC_B : HSC_B;
    const unsigned myBase = sh_hp + warp * 2 * DP * 4 + lane * DPL * 4, myScBase = sh_sc + warp * 16;
    const bool edge_out = warp < 2 && downstream;
    const unsigned eBase = sh_edge + (warp == 0 ? 0 : 2) * DP * 4 + lane * DPL * 4, eScBase = sh_esc + warp * 16;
    // hand-off flags: rows < *flag of the producer are done.  Absent producers / consumers read a constant.
    volatile int* const fV = vIn ? prog + warp + 1 : &ctl->halo_ready;
    volatile int* const fA = aIn ? prog + warp + 2 : &ctl->halo_ready;
    volatile int* const fW1 = warp >= 1 ? prog + warp - 1 : fV;   // no consumer: alias a flag that is already waited on
    volatile int* const fW2 = warp >= 2 ? prog + warp - 2 : fV;
    volatile int* const fC = edge_out ? &ctl->copied : fV;
    const int vCap = vIn ? 0x7fffffff : hend, aCap = aIn ? 0x7fffffff : hend;
    const int wOff = S - 2;            // consumers must have finished row y-S+1  <=>  prog >= y-S+2
    const int cOff = RING - 1;         // downstream ring slot free once rows < y-RING+1 were copied out

    // diagonal path state (registers)
    float hpd[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) hpd[j] = ROO_INF;
    float lbd = 0.0f, pixd = 0.0f;

    // EDGE rows contain a path start (y == 0, x' == 0 or x' == w-1); all other rows take the lean body.
    auto row_body = [&](auto masked_tag, auto edge_tag, int y, int xp, int x) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int lim = MASKED ? min(M, x + 1) - d0 : 0;
        const unsigned stg = pfBase + ((unsigned)y & (PFS - 1)) * STAGE_B;
        float hin[DPL], hpV[DPL], hpA[DPL], H3[DPL], cost[DPL];
        if (!FIRST) lds_vec<DPL>(hin, stg + lane * DPL * 4);
        RawCost<DPL, COST> rc;
        rc.lds(stg + DP * 4 + lane * DPL * CE);
        float pix;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pix) : "r"(stg + DP * 4 + DP * CE));
#pragma unroll
        for (int j = 0; j < DPL; ++j) cost[j] = rc.get(j, cscale);
        const unsigned ym1 = (unsigned)(y - 1);
        lds_vec<DPL>(hpV, vBase + (ym1 & vMask) * vStride);
        lds_vec<DPL>(hpA, aBase + (ym1 & aMask) * aStride);
        const float4 scV = lds_f4(vScBase + (ym1 & vMask) * vScStride);   // {lastBest(vertical), lastBest(anti), pix, -}
        const float4 scA = lds_f4(aScBase + (ym1 & aMask) * aScStride);
        float lbV = scV.x, ppV = scV.z, p2V = P2, lbA = scA.y, ppA = scA.z, p2A = P2, p2D = P2;
        bool sV = false, sD = false, sA = false;
        if (EDGE) {
            sV = y == 0; sD = y == 0 || xp == 0; sA = y == 0 || xp == w - 1;
            if (sV) { lbV = 0.0f; ppV = pix; p2V = 0.0f; }
            if (sD) { lbd = 0.0f; p2D = 0.0f; }
            if (sA) { lbA = 0.0f; ppA = pix; p2A = 0.0f; }
#pragma unroll
            for (int j = 0; j < DPL; ++j) {
                hpV[j] = sV ? ROO_INF : hpV[j];
                hpd[j] = sD ? ROO_INF : hpd[j];
                hpA[j] = sA ? ROO_INF : hpA[j];
            }
        }
        float bV, bD, bA;
        sgm_step3<DPL, MASKED, FIRST, IEEE>(hpV, lbV, 1.0f + fabsf(ppV - pix), p2V,
                                            hpd, lbd, 1.0f + fabsf(pixd - pix), p2D,
                                            hpA, lbA, 1.0f + fabsf(ppA - pix), p2A,
                                            cost, hin, P1, lim, lane, H3, bV, bD, bA);
        if (EDGE) { if (sV) bV = 0.0f; if (sD) bD = 0.0f; if (sA) bA = 0.0f; }
        lbd = bD;
        pixd = pix;

        // ---- publish this pixel's states for the next row, store the aggregate
        const unsigned slot = (unsigned)y & (S - 1);
        sts_vec<DPL>(myBase + slot * SLOT_B, hpV);
        sts_vec<DPL>(myBase + slot * SLOT_B + DP * 4, hpA);
        if (lane == 0) sts_f4(myScBase + slot * SLOTSC_B, make_float4(bV, bA, pix, 0.0f));
        if (edge_out) {
            const unsigned er = (unsigned)y & (RING - 1);
            if (warp == 0) sts_vec<DPL>(eBase + er * HROW_B, hpV);
            sts_vec<DPL>(eBase + er * HROW_B + (warp == 0 ? DP * 4 : 0), hpA);
            if (lane == 0) sts_f4(eScBase + er * HSC_B, make_float4(bV, bA, pix, 0.0f));
        }
        store_f<DPL>(hst, H3);
        hst += estep;
    };

    // No CTA-wide barrier: the columns of a band form a dataflow pipeline through shared memory.  Column j
    // may start row y once columns j+1, j+2 have finished row y-1 (read-after-write) and columns j-1, j-2 have
    // finished row y-S+1 (so the ring slot of row y-S is free: write-after-read).
#pragma unroll 1
    for (int y = y_in; y <= y_out; ++y) {
        issue_row(y + PFS - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(PFS - 1) : "memory");   // row y's stage has landed
        __syncwarp();
        const int xp = u + y;
        // all hand-offs of this row in one polling loop (the flags are read back to back)
        // (an upstream band only publishes rows < hend: rows beyond that need no upstream state)
        while (*fV < min(y, vCap) || *fA < min(y, aCap) || min(*fW1, *fW2) < y - wOff || (edge_out && *fC < y - cOff)) { __nanosleep(VG_SPIN_NS); }
        smem_order();

        const int x = fwd ? xp : w - 1 - xp;
        const bool edge = y == 0 || xp == 0 || xp == w - 1;
        if (edge) {
            if (x >= xf) row_body(std::false_type{}, std::true_type{}, y, xp, x);
            else row_body(std::true_type{}, std::true_type{}, y, xp, x);
        } else {
            if (x >= xf) row_body(std::false_type{}, std::false_type{}, y, xp, x);
            else row_body(std::true_type{}, std::false_type{}, y, xp, x);
        }
        __syncwarp();
        smem_order();
        if (lane == 0) prog[warp] = (y == y_out) ? 0x7fffffff : y + 1;
    }
}

int vgroup_bands(int w, int h, int DP) { return cdiv(w + h - 1, vg_nw_of_dp(DP)); }
size_t vgroup_edge_floats(int w, int h, int DP) { return (size_t)vgroup_bands(w, h, DP) * h * (3 * (size_t)DP + 8); }

template <int DPL, int COST>
static int vgroup_launch2(const VGroupArgs& a, bool first, cudaStream_t st) {
    constexpr int DP = 32 * DPL;
    constexpr int VG_NW = vg_nw(DPL);
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    const size_t smem = (size_t)(VG_S * VG_NW * 2 * DP + VG_S * VG_NW * 4 + 2 * (2 * vg_r(DPL) * (3 * DP + 8))) * sizeof(float) +
                        sizeof(VCtl) + ((VG_NW * 4 + 15) / 16) * 16 + (size_t)VG_NW * vg_pfs(DPL, CE) * (DP * 4 + DP * CE + 16);
    dim3 grid(a.n_bands * a.batch), block((VG_NW + 1) * 32);
    const bool ieee = g_ieee_div.load() != 0;
#define ROO_VG(F, I)                                                                                          \
    do {                                                                                                      \
        auto kern = sgm_vgroup_kernel<DPL, COST, F, I, VG_NW>;                                                \
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
