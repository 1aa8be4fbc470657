// One step of the reference's path recurrence (src/cu_semi_global_matching.cu:39-56), written for a
// warp that holds the 32*DPL disparities of one pixel in registers (lane l owns [l*DPL, (l+1)*DPL)).
// Shared by the single-path sweep kernel (sgm.cu) and the fused vertical-group kernel (sgm_fused.cu).
#pragma once
#include "common.cuh"

namespace roo_b200 {

constexpr float SGM_MAX_ERROR = 1E30f;  // cu_semi_global_matching.cu:24
#define ROO_INF __int_as_float(0x7f800000)

template <int DPL>
__device__ __forceinline__ void load_f(float (&v)[DPL], const float* p) {
    if constexpr (DPL >= 8) {
#pragma unroll
        for (int q = 0; q < DPL / 4; ++q) {
            const float4 a = reinterpret_cast<const float4*>(p)[q];
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
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
    if constexpr (DPL >= 8) {
#pragma unroll
        for (int q = 0; q < DPL / 4; ++q) reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else if constexpr (DPL == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (DPL == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        *p = v[0];
    }
}

// shared-memory vector access by 32-bit shared-window address
__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f4(unsigned addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int DPL>
__device__ __forceinline__ void lds_vec(float (&v)[DPL], unsigned addr) {
    if constexpr (DPL >= 4) {
#pragma unroll
        for (int q = 0; q < DPL / 4; ++q) {
            const float4 t = lds_f4(addr + 16 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else if constexpr (DPL == 2) {
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr));
    } else {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(addr));
    }
}
template <int DPL>
__device__ __forceinline__ void sts_vec(unsigned addr, const float (&v)[DPL]) {
    if constexpr (DPL >= 4) {
#pragma unroll
        for (int q = 0; q < DPL / 4; ++q) sts_f4(addr + 16 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    } else if constexpr (DPL == 2) {
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v[0]) : "memory");
    }
}

// BYTES per lane from global to a 32-bit shared address; 16-byte pieces bypass L1 (.cg), smaller ones use .ca
// (only used for data that is read-only during the pass)
template <int BYTES>
__device__ __forceinline__ void cp_async_bytes(unsigned sdst, const void* gsrc) {
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int q = 0; q < BYTES / 16; ++q)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + 16 * q), "l"((const char*)gsrc + 16 * q) : "memory");
    } else if constexpr (BYTES == 8) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sdst), "l"(gsrc) : "memory");
    } else if constexpr (BYTES == 4) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst), "l"(gsrc) : "memory");
    } else {
        static_assert(BYTES == 2 || BYTES == 1, "unsupported cp.async size");
        // 1- and 2-byte cost rows (DPL 1/2 with u8 costs): plain load + shared store
        if constexpr (BYTES == 2) { const unsigned short v = *(const unsigned short*)gsrc; asm volatile("st.shared.u16 [%0], %1;" ::"r"(sdst), "h"(v) : "memory"); }
        else { const unsigned v = *(const unsigned char*)gsrc; asm volatile("st.shared.u8 [%0], %1;" ::"r"(sdst), "r"(v) : "memory"); }
    }
}

// 4-byte copy issued by the lanes with pred != 0 only -- a predicated instruction, not a divergent branch
__device__ __forceinline__ void cp_async_4_if(int pred, unsigned sdst, const void* gsrc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %0, 0;\n\t@p cp.async.ca.shared.global [%1], [%2], 4;\n\t}"
                 ::"r"(pred), "r"(sdst), "l"(gsrc) : "memory");
}

// raw matching cost of one pixel as loaded; converted to float at use
// COST_CEN32: no cost volume at all -- the sweep recomputes popc(L ^ R) from the 32-bit halves of the census
// descriptors the reference's HammingDistance looks at (hamming_distance.h:40-44; 9x7 window, compat popcount)
enum CostKind { COST_F32 = 0, COST_U8 = 1, COST_CEN32 = 2 };
template <int DPL, int COST> struct RawCost;
template <int DPL> struct RawCost<DPL, COST_F32> {
    float v[DPL];
    __device__ __forceinline__ void load(const void* p) { load_f<DPL>(v, (const float*)p); }
    __device__ __forceinline__ void lds(unsigned saddr) {   // from a 32-bit shared-window address
        if constexpr (DPL >= 4) {
#pragma unroll
            for (int q = 0; q < DPL / 4; ++q)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4 * q]), "=f"(v[4 * q + 1]), "=f"(v[4 * q + 2]), "=f"(v[4 * q + 3]) : "r"(saddr + 16 * q));
        } else if constexpr (DPL == 2) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(saddr));
        else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(saddr));
    }
    __device__ __forceinline__ float get(int j, float) const { return v[j]; }
    __device__ __forceinline__ float raw(int j) const { return v[j]; }   // cost = raw * 1
    static constexpr int ELEM = 4;
};
template <int DPL> struct RawCost<DPL, COST_CEN32> {
    static constexpr int ELEM = 0;   // nothing staged: the cost comes out of registers
    __device__ __forceinline__ void lds(unsigned) {}
    __device__ __forceinline__ float raw(int) const { return 0.0f; }
};
template <int DPL> struct RawCost<DPL, COST_U8> {
    unsigned w[(DPL + 3) / 4];
    __device__ __forceinline__ void load(const void* p) {
        if constexpr (DPL == 16) { const uint4 t = *reinterpret_cast<const uint4*>(p); w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w; }
        else if constexpr (DPL == 8) { const uint2 t = *reinterpret_cast<const uint2*>(p); w[0] = t.x; w[1] = t.y; }
        else if constexpr (DPL == 4) w[0] = *reinterpret_cast<const unsigned*>(p);
        else if constexpr (DPL == 2) w[0] = *reinterpret_cast<const unsigned short*>(p);
        else w[0] = *reinterpret_cast<const unsigned char*>(p);
    }
    __device__ __forceinline__ void lds(unsigned saddr) {
        if constexpr (DPL == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(saddr));
        else if constexpr (DPL == 8) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(saddr));
        else if constexpr (DPL == 4) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(saddr));
        else if constexpr (DPL == 2) { unsigned short t; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(t) : "r"(saddr)); w[0] = t; }
        else asm volatile("ld.shared.u8 %0, [%1];" : "=r"(w[0]) : "r"(saddr));
    }
    // Hamming count * (1/bits): exact (power-of-two scale), equals the reference's count / bits
    __device__ __forceinline__ float get(int j, float scale) const {
        return (float)((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) * scale;
    }
    __device__ __forceinline__ float raw(int j) const { return (float)((w[j >> 2] >> (8 * (j & 3))) & 0xFFu); }
    static constexpr int ELEM = 1;
};

// ---- packed fp32 pairs (sm_100a add/fma.rn.f32x2 -> SASS FADD2 / FFMA2) -----------------------------------------
// One instruction, two IEEE-rounded fp32 results (each half is rounded exactly like the scalar instruction, so
// results stay bit-identical); the aggregation kernels are bound by instruction issue, and the recurrence's
// additions on adjacent disparities are independent, so pairing them halves their issue slots.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// The recurrence for one pixel of one path.
//   hp[]      previous pixel's aggregate row on this path, +inf where d >= that pixel's disparity range
//   lastBest  min_d Cr of the previous pixel (0 at a path start)
//   denom     1 + |I(prev) - I(cur)|
//   P2        0 at a path start: together with hp = +inf and lastBest = 0 this makes Cr == cost, i.e. the
//             start pixel's `volH += volC` (cu_semi_global_matching.cu:31-35) is the same code path
//   craw/cs   raw matching cost and its scale: cost = craw * cs, exact (cs is a power of two, or 1 for fp32 costs),
//             so fma(craw, cs, CM) is the reference's CM + cost with its single rounding
//   lim       (MASKED only) number of in-range disparities of this lane: min(maxDisp, x+1) - lane*DPL
// Outputs hnew = hin + Cr (hin where out of range), hp_out = hnew masked with +inf, best = min_d Cr.
template <int DPL, bool MASKED, bool FIRST, bool IEEE>
__device__ __forceinline__ void sgm_step(const float (&hp)[DPL], float lastBest, float denom, float P1, float P2,
                                         const float (&craw)[DPL], float cs, const float (&hin)[DPL], int lim,
                                         int lane, float (&hnew)[DPL], float (&hp_out)[DPL], float& best_out) {
    float hpP[DPL], Cr[DPL], h[DPL];
    if constexpr (DPL >= 2) {
        const f32x2 P1P1 = pk2(P1, P1);
#pragma unroll
        for (int q = 0; q < DPL / 2; ++q) upk2(add2(pk2(hp[2 * q], hp[2 * q + 1]), P1P1), hpP[2 * q], hpP[2 * q + 1]);
    } else {
        hpP[0] = hp[0] + P1;
    }
    float up = __shfl_up_sync(0xffffffffu, hpP[DPL - 1], 1);   // H(prev, d-1) + P1 for j == 0
    float dn = __shfl_down_sync(0xffffffffu, hpP[0], 1);       // H(prev, d+1) + P1 for j == DPL-1
    if (lane == 0) up = ROO_INF;
    if (lane == 31) dn = ROO_INF;
    const float base = sgm_p2_base<IEEE>(lastBest, P2, denom);
    float CM[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
        const float hm = j > 0 ? hpP[j - 1] : up;
        const float hq = j < DPL - 1 ? hpP[j + 1] : dn;
        CM[j] = fminf(fminf(base, hp[j]), fminf(hm, hq));
    }
    // MASKED: an out-of-range disparity gets the cost +inf.  Cr and the new aggregate are then +inf by themselves -- which
    // is exactly the masked state row (hp_out) -- and never win a minimum; only the value written to H needs a select.
    float cm_[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) cm_[j] = (MASKED && !(j < lim)) ? ROO_INF : CM[j];
    if constexpr (DPL >= 2) {
        const f32x2 cs2 = pk2(cs, cs), nlb = pk2(-lastBest, -lastBest);
#pragma unroll
        for (int q = 0; q < DPL / 2; ++q) {
            const f32x2 cr = add2(fma2(pk2(craw[2 * q], craw[2 * q + 1]), cs2, pk2(cm_[2 * q], cm_[2 * q + 1])), nlb);
            upk2(cr, Cr[2 * q], Cr[2 * q + 1]);
            if (FIRST) { h[2 * q] = Cr[2 * q]; h[2 * q + 1] = Cr[2 * q + 1]; }
            else upk2(add2(pk2(hin[2 * q], hin[2 * q + 1]), cr), h[2 * q], h[2 * q + 1]);
        }
    } else {
        Cr[0] = __fmaf_rn(craw[0], cs, cm_[0]) - lastBest;
        h[0] = FIRST ? Cr[0] : hin[0] + Cr[0];
    }
    float best = SGM_MAX_ERROR;
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
        best = fminf(best, Cr[j]);                       // +inf where masked: never the minimum
        hp_out[j] = h[j];                                // +inf where masked
        hnew[j] = (MASKED && !(j < lim)) ? (FIRST ? 0.0f : hin[j]) : h[j];
    }
    best_out = warp_min_f32(best);
}

// Three path steps of one pixel at once -- the vertical, diagonal and anti-diagonal path of a fused
// vertical group (sgm_fused.cu), applied in that order to the same aggregate: H1 = hin + CrV, H2 = H1 + CrD,
// H3 = H2 + CrA.  Numerically identical to three sgm_step() calls; written as ONE straight-line block so
// that the three recurrences -- which only meet in the final additions -- are scheduled interleaved and a
// pixel costs one chain latency (shuffle -> min -> redux) instead of three.  The matching cost is decoded
// once.  hpX are in/out: previous pixel's row on entry, this pixel's masked row on exit.
// All additions run on pairs of adjacent disparities (FADD2 / FFMA2): 12 + 12 + 12 scalar adds and 12 scalar
// fmas per four disparities become 18 packed instructions.
template <int DPL, bool MASKED, bool FIRST, bool IEEE>
__device__ __forceinline__ void sgm_step3(float (&hpV)[DPL], float lbV, float denV, float p2V,
                                          float (&hpD)[DPL], float lbD, float denD, float p2D,
                                          float (&hpA)[DPL], float lbA, float denA, float p2A,
                                          const float (&craw)[DPL], float cs, const float (&hin)[DPL], float P1, int lim,
                                          int lane, float (&H3)[DPL], float& bV, float& bD, float& bA) {
    float pV[DPL], pD[DPL], pA[DPL];
    if constexpr (DPL >= 2) {
        const f32x2 P1P1 = pk2(P1, P1);
#pragma unroll
        for (int q = 0; q < DPL / 2; ++q) {
            upk2(add2(pk2(hpV[2 * q], hpV[2 * q + 1]), P1P1), pV[2 * q], pV[2 * q + 1]);
            upk2(add2(pk2(hpD[2 * q], hpD[2 * q + 1]), P1P1), pD[2 * q], pD[2 * q + 1]);
            upk2(add2(pk2(hpA[2 * q], hpA[2 * q + 1]), P1P1), pA[2 * q], pA[2 * q + 1]);
        }
    } else {
        pV[0] = hpV[0] + P1; pD[0] = hpD[0] + P1; pA[0] = hpA[0] + P1;
    }
    float upV = __shfl_up_sync(0xffffffffu, pV[DPL - 1], 1), dnV = __shfl_down_sync(0xffffffffu, pV[0], 1);
    float upD = __shfl_up_sync(0xffffffffu, pD[DPL - 1], 1), dnD = __shfl_down_sync(0xffffffffu, pD[0], 1);
    float upA = __shfl_up_sync(0xffffffffu, pA[DPL - 1], 1), dnA = __shfl_down_sync(0xffffffffu, pA[0], 1);
    if (lane == 0) { upV = ROO_INF; upD = ROO_INF; upA = ROO_INF; }
    if (lane == 31) { dnV = ROO_INF; dnD = ROO_INF; dnA = ROO_INF; }
    const float baseV = sgm_p2_base<IEEE>(lbV, p2V, denV);
    const float baseD = sgm_p2_base<IEEE>(lbD, p2D, denD);
    const float baseA = sgm_p2_base<IEEE>(lbA, p2A, denA);
    float cmV[DPL], cmD[DPL], cmA[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
        cmV[j] = fminf(fminf(baseV, hpV[j]), fminf(j > 0 ? pV[j - 1] : upV, j < DPL - 1 ? pV[j + 1] : dnV));
        cmD[j] = fminf(fminf(baseD, hpD[j]), fminf(j > 0 ? pD[j - 1] : upD, j < DPL - 1 ? pD[j + 1] : dnD));
        cmA[j] = fminf(fminf(baseA, hpA[j]), fminf(j > 0 ? pA[j - 1] : upA, j < DPL - 1 ? pA[j + 1] : dnA));
    }
    // MASKED: out-of-range disparities get +inf in all three CM rows, so their Cr and aggregates are +inf (the masked
    // state rows) and never win a minimum; only H3 needs a select (see sgm_step)
    if (MASKED) {
#pragma unroll
        for (int j = 0; j < DPL; ++j)
            if (!(j < lim)) { cmV[j] = ROO_INF; cmD[j] = ROO_INF; cmA[j] = ROO_INF; }
    }
    float crV[DPL], crD[DPL], crA[DPL], h1[DPL], h2[DPL], h3[DPL];
    if constexpr (DPL >= 2) {
        const f32x2 cs2 = pk2(cs, cs), nV = pk2(-lbV, -lbV), nD = pk2(-lbD, -lbD), nA = pk2(-lbA, -lbA);
#pragma unroll
        for (int q = 0; q < DPL / 2; ++q) {
            const int a = 2 * q, b = 2 * q + 1;
            const f32x2 c2 = pk2(craw[a], craw[b]);
            const f32x2 rV = add2(fma2(c2, cs2, pk2(cmV[a], cmV[b])), nV);
            const f32x2 rD = add2(fma2(c2, cs2, pk2(cmD[a], cmD[b])), nD);
            const f32x2 rA = add2(fma2(c2, cs2, pk2(cmA[a], cmA[b])), nA);
            const f32x2 s1 = FIRST ? rV : add2(pk2(hin[a], hin[b]), rV);
            const f32x2 s2 = add2(s1, rD);
            const f32x2 s3 = add2(s2, rA);
            upk2(rV, crV[a], crV[b]); upk2(rD, crD[a], crD[b]); upk2(rA, crA[a], crA[b]);
            upk2(s1, h1[a], h1[b]); upk2(s2, h2[a], h2[b]); upk2(s3, h3[a], h3[b]);
        }
    } else {
        crV[0] = __fmaf_rn(craw[0], cs, cmV[0]) - lbV;
        crD[0] = __fmaf_rn(craw[0], cs, cmD[0]) - lbD;
        crA[0] = __fmaf_rn(craw[0], cs, cmA[0]) - lbA;
        h1[0] = FIRST ? crV[0] : hin[0] + crV[0];
        h2[0] = h1[0] + crD[0];
        h3[0] = h2[0] + crA[0];
    }
    float mV = SGM_MAX_ERROR, mD = SGM_MAX_ERROR, mA = SGM_MAX_ERROR;
    float tV = 0.0f, tD = 0.0f, tA = 0.0f;
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
        // pairs first, so that every second min is a three-input FMNMX3 (min is exact: any grouping agrees)
        if (j & 1) {
            mV = fminf(mV, fminf(tV, crV[j])); mD = fminf(mD, fminf(tD, crD[j])); mA = fminf(mA, fminf(tA, crA[j]));
        } else if (j == DPL - 1) {
            mV = fminf(mV, crV[j]); mD = fminf(mD, crD[j]); mA = fminf(mA, crA[j]);
        } else {
            tV = crV[j]; tD = crD[j]; tA = crA[j];
        }
        hpV[j] = h1[j]; hpD[j] = h2[j]; hpA[j] = h3[j];
        H3[j] = (MASKED && !(j < lim)) ? (FIRST ? 0.0f : hin[j]) : h3[j];
    }
    bV = warp_min_f32(mV);
    bD = warp_min_f32(mD);
    bA = warp_min_f32(mA);
}

// Winner-takes-all (+ optional parabola) over the masked row hp[] of pixel x -- CostVolMinimum<float,float>
// (cu_dense_stereo.cu:25-43) or CostVolMinimumSubpix with sd = -1 (cu_dense_stereo.cu:66-109).
// scratch: 32-bit shared-window address of 32*DPL floats private to the warp (0: none).  With it the two parabola taps
// H(bestd-1), H(bestd+1) -- warp-uniform positions -- are read back from the row each lane parks there (2 STS.128 + 2 broadcast
// LDS at 256 disparities) instead of 2*DPL compare-select pairs per lane and two shuffles.
template <int DPL, bool IEEE>
__device__ __forceinline__ float wta_epilogue(const float (&hp)[DPL], int lane, int x, int w, int maxDispVal, int subpix,
                                              unsigned scratch = 0) {
    const int d0 = lane * DPL;
    // lane minimum by a min tree (no index tracking), warp minimum by one redux.sync.min.f32; then the LOWEST
    // disparity that holds it: first j inside the lane, lowest lane through one integer redux.sync.min.s32
    // (all rows masked out are +inf; if every entry is +inf the answer is 0, as with the reference's strict `<`)
    float lc = hp[0];
#pragma unroll
    for (int j = 1; j < DPL; ++j) lc = fminf(lc, hp[j]);
    const float m = warp_min_f32(lc);
    int cand = 0x7fffffff;
#pragma unroll
    for (int j = DPL - 1; j >= 0; --j) cand = (hp[j] == m) ? d0 + j : cand;
    int bestd = __reduce_min_sync(0xffffffffu, cand);
    if (!subpix) return (float)bestd;
    float bestc = m;
    if (!(bestc < 1E10f)) { bestc = 1E10f; bestd = 0; }  // the reference starts from bestc = 1e10
    float out = (float)bestd;
    const int bestxr = x - bestd;
    if (0 < bestxr && bestxr < w - 1 && bestd + 1 < maxDispVal) {  // bestd+1 == vol.d: out of bounds in the reference
        const int dl = max(bestd - 1, 0);      // float -> unsigned saturation in the reference (Q7)
        const int dr = bestd + 1;
        float sl, sr;
        if (scratch != 0) {
            __syncwarp();                       // the previous pixel's taps have been read
            sts_vec<DPL>(scratch + d0 * 4, hp);
            __syncwarp();
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sl) : "r"(scratch + dl * 4));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sr) : "r"(scratch + dr * 4));
        } else {
            float slc = 0.0f, src = 0.0f;
#pragma unroll
            for (int j = 0; j < DPL; ++j) {
                if (d0 + j == dl) slc = hp[j];
                if (d0 + j == dr) src = hp[j];
            }
            sl = __shfl_sync(0xffffffffu, slc, dl / DPL);
            sr = __shfl_sync(0xffffffffu, src, dr / DPL);
        }
        const float sub = parabola_vertex<IEEE>((float)bestd, bestc, sl, sr);
        if ((float)(bestd - 1) < sub && sub < (float)(bestd + 1)) out = sub;
    }
    return out;
}

}  // namespace roo_b200
