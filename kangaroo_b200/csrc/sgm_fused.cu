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

constexpr int VG_PF = 3;     // rows of prefetch
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

template <int DPL, int COST>
struct VStage {
    float hin[DPL];
    RawCost<DPL, COST> c;
    float pix;
};

// smem control words
struct VCtl { volatile int halo_ready; volatile int rows_done; volatile int copied; int pad; };

constexpr int VG_R = 8;   // rows per hand-off chunk between bands (flag / fence cost is paid once per chunk)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int DPL, int COST, bool FIRST, bool IEEE, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
sgm_vgroup_kernel(const VGroupArgs a) {
    constexpr int DP = 32 * DPL;
    constexpr int CE = RawCost<DPL, COST>::ELEM;
    constexpr int PF = VG_PF;
    constexpr int R = VG_R, RING = 2 * VG_R;
    extern __shared__ __align__(16) float smem[];
    float* s_hp = smem;                                // [2 parity][NW][2 paths][DP]  in-band states of the previous row
    float* s_sc = s_hp + 2 * NW * 2 * DP;              // [2 parity][NW][4]: lastBest(vertical), lastBest(anti-diag), pix, -
    float* s_halo = s_sc + 2 * NW * 4;                 // [RING rows][3][DP]  upstream band's columns 0,1 (row ring)
    float* s_hsc = s_halo + RING * 3 * DP;             // [RING][8]
    float* s_edge = s_hsc + RING * 8;                  // [RING rows][3][DP]  this band's columns 0,1 for downstream
    float* s_esc = s_edge + RING * 3 * DP;             // [RING][8]
    VCtl* ctl = reinterpret_cast<VCtl*>(s_esc + RING * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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

    if (threadIdx.x == 0) { ctl->halo_ready = hbeg; ctl->rows_done = ymin; ctl->copied = ymin; }
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
            const int rd = ctl->rows_done;
            // ---- stage a chunk of upstream rows into the halo ring
            if (hr < hend) {
                const int n = min(R, hend - hr);
                // ring slot of row y was last used by row y-RING, consumed while computing row y-RING+1
                if (rd >= hr + n - RING + 1) {
                    if (seen < hr + n) {
                        if (lane == 0) seen = ld_acquire_gpu(p_flag);
                        seen = __shfl_sync(0xffffffffu, seen, 0);
                    }
                    if (seen >= hr + n) {
                        for (int y = hr; y < hr + n; ++y) {
                            const float* src = p_hp + (size_t)y * 3 * DP + lane * DPL;
                            float* dst = s_halo + (size_t)(y % RING) * 3 * DP + lane * DPL;
                            float r0[DPL], r1[DPL], r2[DPL];
                            ldcg_row<DPL>(r0, src);
                            ldcg_row<DPL>(r1, src + DP);
                            ldcg_row<DPL>(r2, src + 2 * DP);
                            sts_row<DPL>(dst, r0);
                            sts_row<DPL>(dst + DP, r1);
                            sts_row<DPL>(dst + 2 * DP, r2);
                            if (lane < 8) s_hsc[(y % RING) * 8 + lane] = __ldcg(p_sc + (size_t)y * 8 + lane);
                        }
                        __threadfence_block();
                        __syncwarp();
                        hr += n;
                        if (lane == 0) ctl->halo_ready = hr;
                        progress = true;
                    }
                }
            }
            // ---- publish finished rows of this band's columns 0,1
            if (cp < cp_end && rd > cp && (rd - cp >= R || rd == ymax + 1)) {
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
            if (!progress) __nanosleep(40);
        }
        return;
    }

    // -------------------------------------------------------------------- compute warps
    const int u = ulo + warp;
    const int y_in = max(0, -u), y_out = min(h - 1, w - 1 - u);   // active rows of this skewed column
    const int d0 = lane * DPL;
    const int xf = (M == DP) ? DP - 1 : 0x3fffffff;               // all lanes in range iff true x >= xf

    // cursors at the first active pixel; one row down the travel direction = +-(w+1) pixels
    const int xp0 = u + y_in;
    const int x0 = fwd ? xp0 : w - 1 - xp0, y0 = fwd ? y_in : h - 1 - y_in;
    const ptrdiff_t pstep = fwd ? (ptrdiff_t)(w + 1) : -(ptrdiff_t)(w + 1);
    const ptrdiff_t estep = pstep * DP;
    const bool any = y_in <= y_out;
    const size_t e0 = any ? ((size_t)y0 * w + x0) * DP + d0 : 0;
    float* hst = a.H + (size_t)pair * a.h_pair + e0;
    const float* hld = hst;
    const char* cld = (const char*)a.C + ((size_t)pair * a.c_pair + e0) * CE;
    const float* ild = a.img + (size_t)pair * a.img_pair + (any ? (size_t)y0 * w + x0 : 0);

    VStage<DPL, COST> ring[PF];
    auto load_stage = [&](VStage<DPL, COST>& st) {
        if (!FIRST) load_f<DPL>(st.hin, hld);
        st.c.load(cld);
        st.pix = *ild;
        hld += estep; cld += estep * CE; ild += pstep;
    };
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        const int yl = ymin + k;
        if (yl >= y_in && yl <= y_out) load_stage(ring[k]);
    }

    // diagonal path state (registers)
    float hpd[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) hpd[j] = ROO_INF;
    float lbd = 0.0f, pixd = 0.0f;

    auto row_body = [&](auto masked_tag, VStage<DPL, COST>& st, int y, int xp, int x) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        const int p = y & 1;
        const int lim = MASKED ? min(M, x + 1) - d0 : 0;
        const float* nb_hp = s_hp + (size_t)((p ^ 1) * NW) * 2 * DP + lane * DPL;   // previous row's in-band states
        const float* nb_sc = s_sc + (size_t)((p ^ 1) * NW) * 4;
        const float* ha_hp = s_halo + (size_t)((y - 1 + RING) % RING) * 3 * DP + lane * DPL;   // upstream row y-1
        const float* ha_sc = s_hsc + ((y - 1 + RING) % RING) * 8;
        float H1[DPL], H2[DPL], H3[DPL], hp[DPL], hp1[DPL], hp3[DPL], b1, b2, b3;

        // ---- vertical path: previous pixel (x', y'-1) lives in column u+1
        float lb = 0.0f, pp = st.pix, p2 = 0.0f;
        if (y > 0) {
            if (warp + 1 < NW) {
                lds_row<DPL>(hp, nb_hp + (size_t)(warp + 1) * 2 * DP);
                lb = nb_sc[(warp + 1) * 4 + 0];
                pp = nb_sc[(warp + 1) * 4 + 2];
            } else {                       // upstream column 0: {rec0; sc0 = lastBest(vertical), sc2 = pix}
                lds_row<DPL>(hp, ha_hp);
                lb = ha_sc[0];
                pp = ha_sc[2];
            }
            p2 = P2;
        } else {
#pragma unroll
            for (int j = 0; j < DPL; ++j) hp[j] = ROO_INF;
        }
        sgm_step<DPL, MASKED, FIRST, IEEE>(hp, lb, 1.0f + fabsf(pp - st.pix), P1, p2, st.c, cscale, st.hin, lim, lane, H1, hp1, b1);
        if (y == 0) b1 = 0.0f;

        // ---- diagonal path: previous pixel (x'-1, y'-1) is this column's previous row
        const bool s1 = y == 0 || xp == 0;
        if (s1) {
#pragma unroll
            for (int j = 0; j < DPL; ++j) hpd[j] = ROO_INF;
            lbd = 0.0f;
        }
        sgm_step<DPL, MASKED, false, IEEE>(hpd, lbd, 1.0f + fabsf(pixd - st.pix), P1, s1 ? 0.0f : P2, st.c, cscale, H1, lim, lane, H2, hpd, b2);
        lbd = s1 ? 0.0f : b2;
        pixd = st.pix;

        // ---- anti-diagonal path: previous pixel (x'+1, y'-1) lives in column u+2
        const bool s2 = y == 0 || xp == w - 1;
        lb = 0.0f; pp = st.pix; p2 = 0.0f;
        if (!s2) {
            if (warp + 2 < NW) {
                lds_row<DPL>(hp, nb_hp + ((size_t)(warp + 2) * 2 + 1) * DP);
                lb = nb_sc[(warp + 2) * 4 + 1];
                pp = nb_sc[(warp + 2) * 4 + 2];
            } else if (warp + 2 == NW) {   // upstream column 0: {rec1; sc1 = lastBest(anti), sc2 = pix}
                lds_row<DPL>(hp, ha_hp + DP);
                lb = ha_sc[1];
                pp = ha_sc[2];
            } else {                       // upstream column 1: {rec2; sc3 = lastBest(anti), sc4 = pix}
                lds_row<DPL>(hp, ha_hp + 2 * DP);
                lb = ha_sc[3];
                pp = ha_sc[4];
            }
            p2 = P2;
        } else {
#pragma unroll
            for (int j = 0; j < DPL; ++j) hp[j] = ROO_INF;
        }
        sgm_step<DPL, MASKED, false, IEEE>(hp, lb, 1.0f + fabsf(pp - st.pix), P1, p2, st.c, cscale, H2, lim, lane, H3, hp3, b3);
        if (s2) b3 = 0.0f;

        // ---- publish this pixel's states for the next row, store the aggregate
        float* my_hp = s_hp + (size_t)((p * NW + warp) * 2) * DP + lane * DPL;
        sts_row<DPL>(my_hp, hp1);
        sts_row<DPL>(my_hp + DP, hp3);
        if (lane == 0) {
            float* my_sc = s_sc + (size_t)(p * NW + warp) * 4;
            my_sc[0] = b1; my_sc[1] = b3; my_sc[2] = st.pix;
        }
        if (warp < 2 && downstream) {
            float* dst = s_edge + (size_t)(y % RING) * 3 * DP + lane * DPL;
            float* dsc = s_esc + (y % RING) * 8;
            if (warp == 0) {
                sts_row<DPL>(dst, hp1);
                sts_row<DPL>(dst + DP, hp3);
                if (lane == 0) { dsc[0] = b1; dsc[1] = b3; dsc[2] = st.pix; }
            } else {
                sts_row<DPL>(dst + 2 * DP, hp3);
                if (lane == 0) { dsc[3] = b3; dsc[4] = st.pix; }
            }
        }
        store_f<DPL>(hst, H3);
        hst += estep;
    };

    for (int yb = ymin; yb <= ymax; yb += PF) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int y = yb + k;
            if (y > ymax) break;
            if (y >= y_in && y <= y_out) {
                // the last two columns read the upstream band's row y-1; the first two feed the downstream ring
                if (warp >= NW - 2 && y - 1 >= hbeg && y - 1 < hend) {
                    while (ctl->halo_ready < y) __nanosleep(20);
                    __threadfence_block();
                }
                if (warp < 2 && downstream) {
                    while (ctl->copied < y - RING + 1) __nanosleep(20);
                }
                const int xp = u + y;
                const int x = fwd ? xp : w - 1 - xp;
                if (x >= xf) row_body(std::false_type{}, ring[k], y, xp, x);
                else row_body(std::true_type{}, ring[k], y, xp, x);
            }
            const int yl = y + PF;
            if (yl >= y_in && yl <= y_out) load_stage(ring[k]);
            named_bar_sync(1, NW * 32);                  // compute warps only: row y's states are in shared memory
            if (warp == 0 && lane == 0) { __threadfence_block(); ctl->rows_done = y + 1; }
        }
    }
}

int vgroup_bands(int w, int h, int DP) { return cdiv(w + h - 1, vg_nw_of_dp(DP)); }
size_t vgroup_edge_floats(int w, int h, int DP) { return (size_t)vgroup_bands(w, h, DP) * h * (3 * (size_t)DP + 8); }

template <int DPL, int COST>
static int vgroup_launch2(const VGroupArgs& a, bool first, cudaStream_t st) {
    constexpr int DP = 32 * DPL;
    constexpr int VG_NW = vg_nw(DPL);
    const size_t smem = (size_t)(2 * VG_NW * 2 * DP + 2 * VG_NW * 4 + 2 * (2 * VG_R * (3 * DP + 8))) * sizeof(float) + sizeof(VCtl);
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
