/* TEST INFRASTRUCTURE ONLY -- see kangaroo_oracle.h.
 *
 * Scalar CPU restatement of the reference kernel bodies, one function per operator, OpenMP over
 * rows / scanlines.  IEEE fp32 (build with -ffp-contract=off, no -ffast-math) in the reference's
 * operation order.  The reference is compiled with -use_fast_math (CMakeLists.txt:141): its
 * approximate divides differ from the IEEE ones here by <= 2 ulp, inside the
 * 1e-5 / 0.01 px parity bars.  GPU float->unsigned conversions saturate (cvt.rzi); where the
 * reference relies on that (Q7) it is emulated explicitly.
 */
#include "kangaroo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int ko_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ko_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- accessors: Image.h:247-257 (RowPtr/operator()), Volume.h:125-147 ---- */
static inline char* img_at(const ko_image* im, size_t x, size_t y, size_t elem) {
    return (char*)im->ptr + y * im->pitch + x * elem;
}
static inline char* vol_at(const ko_volume* v, size_t x, size_t y, size_t z, size_t elem) {
    return (char*)v->ptr + z * v->img_pitch + y * v->pitch + x * elem;
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Image.h:297-303 GetWithClampedRange, as float for both input types (u8 -> float is exact and
 * order preserving, so `q < p` is the same predicate). */
static inline float px_clamped(const ko_image* im, int in_type, int x, int y) {
    x = clampi(x, 0, (int)im->w - 1);
    y = clampi(y, 0, (int)im->h - 1);
    if (in_type == KO_IMG_U8) return (float)*(const uint8_t*)img_at(im, (size_t)x, (size_t)y, 1);
    return *(const float*)img_at(im, (size_t)x, (size_t)y, 4);
}

/* ---------------------------------------------------------------- census ---- */

/* cu_census.cu:18-46.  bit i = (r+3)*9 + (c+4), set iff img.clamped(x+c,y+r) < img(x,y). */
static uint64_t census9x7(const ko_image* in, int t, int x, int y) {
    const float p = px_clamped(in, t, x, y);
    uint64_t out = 0, bit = 1;
    for (int r = -3; r <= 3; ++r)
        for (int c = -4; c <= 4; ++c) {
            if (px_clamped(in, t, x + c, y + r) < p) out |= bit;
            bit <<= 1;
        }
    return out;
}

/* cu_census.cu:52-110.  x: rows -5..-1 then row 0 c=-5..0 ; y: row 0 c=1..5 then rows 1..5. */
static void census11x11(const ko_image* in, int t, int x, int y, uint64_t o[2]) {
    const float p = px_clamped(in, t, x, y);
    uint64_t bit = 1;
    o[0] = o[1] = 0;
    for (int r = -5; r < 0; ++r)
        for (int c = -5; c <= 5; ++c) {
            if (px_clamped(in, t, x + c, y + r) < p) o[0] |= bit;
            bit <<= 1;
        }
    for (int c = -5; c <= 0; ++c) {
        if (px_clamped(in, t, x + c, y) < p) o[0] |= bit;
        bit <<= 1;
    }
    bit = 1;
    for (int c = 1; c <= 5; ++c) {
        if (px_clamped(in, t, x + c, y) < p) o[1] |= bit;
        bit <<= 1;
    }
    for (int r = 1; r <= 5; ++r)
        for (int c = -5; c <= 5; ++c) {
            if (px_clamped(in, t, x + c, y + r) < p) o[1] |= bit;
            bit <<= 1;
        }
}

/* cu_census.cu:116-177.  "16x16" is c in [-4,3], r in [-8,7]; 4 rows x 8 cols per word. */
static void census16x16(const ko_image* in, int t, int x, int y, uint64_t o[4]) {
    const float p = px_clamped(in, t, x, y);
    for (int k = 0; k < 4; ++k) {
        uint64_t bit = 1, acc = 0;
        for (int r = -8 + 4 * k; r < -4 + 4 * k; ++r)
            for (int c = -4; c < 4; ++c) {
                if (px_clamped(in, t, x + c, y + r) < p) acc |= bit;
                bit <<= 1;
            }
        o[k] = acc;
    }
}

void ko_census(const ko_image* out, const ko_image* in, int window, int in_type) {
    const int w = (int)in->w, h = (int)in->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            if (window == KO_WIN_9x7) {
                *(uint64_t*)img_at(out, (size_t)x, (size_t)y, 8) = census9x7(in, in_type, x, y);
            } else if (window == KO_WIN_11x11) {
                census11x11(in, in_type, x, y, (uint64_t*)img_at(out, (size_t)x, (size_t)y, 16));
            } else {
                census16x16(in, in_type, x, y, (uint64_t*)img_at(out, (size_t)x, (size_t)y, 32));
            }
        }
}

/* ---------------------------------------------------------------- hamming ---- */

/* hamming_distance.h:40-62: __popc (32-bit) applied to a 64-bit XOR => low 32 bits only (Q1). */
unsigned ko_hamming(const uint64_t* p, const uint64_t* q, int words, int popc_mode) {
    unsigned s = 0;
    for (int i = 0; i < words; ++i) {
        const uint64_t v = p[i] ^ q[i];
        if (popc_mode == KO_POPC32_COMPAT) s += (unsigned)__builtin_popcount((uint32_t)v);
        else s += (unsigned)__builtin_popcountll(v);
    }
    return s;
}

/* cu_census.cu:226-259 */
void ko_census_stereo(const ko_image* disp, const ko_image* left, const ko_image* right, int maxDispVal) {
    const int w = (int)disp->w, h = (int)disp->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint64_t p = *(const uint64_t*)img_at(left, (size_t)x, (size_t)y, 8);
            unsigned bestScore = 0xFFFFF;
            int bestDisp = 0; /* InvalidValue<char>::Value(), InvalidValue.h:50-53 */
            int minDisp = maxDispVal < 0 ? maxDispVal : 0;
            int maxDisp = maxDispVal > 0 ? maxDispVal : 0;
            if (minDisp < x - ((int)left->w - 1)) minDisp = x - ((int)left->w - 1);
            if (maxDisp > x) maxDisp = x;
            for (int d = minDisp; d < maxDisp; ++d) {
                const uint64_t q = *(const uint64_t*)img_at(right, (size_t)(x - d), (size_t)y, 8);
                const unsigned score = ko_hamming(&p, &q, 1, KO_POPC32_COMPAT);
                if (score < bestScore) { bestScore = score; bestDisp = d; }
            }
            *(int8_t*)img_at(disp, (size_t)x, (size_t)y, 1) = (int8_t)bestDisp;
        }
}

/* cu_census.cu:272-299.  xd = (int)(x + sd*d) (float, truncation toward zero, Q12);
 * score = Hamming / (float)(8*sizeof(T)) else 0.5; Tvol=unsigned short truncates to 0 (Q2). */
void ko_census_stereo_volume(const ko_volume* vol, const ko_image* left, const ko_image* right, int words,
                             int vol_type, int maxDispVal, float sd, int popc_mode) {
    const int w = (int)left->w, h = (int)left->h;
    const size_t esz = (size_t)words * 8;
    const float bits = (float)(esz * 8);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint64_t* p = (const uint64_t*)img_at(left, (size_t)x, (size_t)y, esz);
            for (int d = 0; d < maxDispVal; ++d) {
                const int xd = (int)((float)x + sd * (float)d);
                float score;
                if (0 <= xd && xd < (int)right->w) {
                    const uint64_t* q = (const uint64_t*)img_at(right, (size_t)xd, (size_t)y, esz);
                    score = (float)ko_hamming(p, q, words, popc_mode) / bits;
                } else {
                    score = 0.5f;
                }
                if (vol_type == KO_VOL_F32) *(float*)vol_at(vol, (size_t)x, (size_t)y, (size_t)d, 4) = score;
                else *(uint16_t*)vol_at(vol, (size_t)x, (size_t)y, (size_t)d, 2) = (uint16_t)score;
            }
        }
}

/* ---------------------------------------------------------------- SGM ---- */

/* CostVolElem.h:12-15 operator float() */
static inline float volc_get(const ko_volume* v, int t, int x, int y, int d) {
    if (t == KO_VOL_F32) return *(const float*)vol_at(v, (size_t)x, (size_t)y, (size_t)d, 4);
    const ko_costvolelem* e = (const ko_costvolelem*)vol_at(v, (size_t)x, (size_t)y, (size_t)d, 8);
    return e->n > 0 ? e->sum / (float)e->n : 1E30f;
}
/* `last_c - c` (cu_semi_global_matching.cu:41): int arithmetic for uchar, float for float; both are
 * exactly representable as the float difference of the converted values. */
static inline float img_get(const ko_image* im, int t, int x, int y) {
    if (t == KO_IMG_U8) return (float)*(const uint8_t*)img_at(im, (size_t)x, (size_t)y, 1);
    return *(const float*)img_at(im, (size_t)x, (size_t)y, 4);
}
static inline int mini(int a, int b) { return a < b ? a : b; }

/* One thread of KernSemiGlobalMatching (cu_semi_global_matching.cu:21-63): one scanline. */
static void sgm_scanline(const ko_volume* H, const ko_volume* C, int ct, const ko_image* left, int it,
                         int maxDispVal, float P1, float P2, int x, int y, int dx, int dy, int pathlen) {
    const float MAX_ERROR = 1E30f;
    float lastBestCr = 0.0f;
    float last_c = img_get(left, it, x, y);
    const int maxDisp0 = mini(maxDispVal, x + 1);
    int lastMaxDisp = maxDisp0;
    for (int d = 0; d < maxDisp0; ++d)
        *(float*)vol_at(H, (size_t)x, (size_t)y, (size_t)d, 4) += volc_get(C, ct, x, y, d);
    x += dx;
    y += dy;
    for (int r = 1; r < pathlen; ++r) {
        const float c = img_get(left, it, x, y);
        const float diff = last_c - c;
        const float _P2 = P2 / (1.0f + fabsf(diff));
        float bestCr = MAX_ERROR;
        const int maxDisp = mini(maxDispVal, x + 1);
        const int px = x - dx, py = y - dy;
        for (int d = 0; d < maxDisp; ++d) {
            float CM = lastBestCr + _P2;
            if (d < lastMaxDisp) CM = fminf(CM, *(const float*)vol_at(H, (size_t)px, (size_t)py, (size_t)d, 4));
            if (d > 0) CM = fminf(CM, *(const float*)vol_at(H, (size_t)px, (size_t)py, (size_t)(d - 1), 4) + P1);
            if (d + 1 < lastMaxDisp)
                CM = fminf(CM, *(const float*)vol_at(H, (size_t)px, (size_t)py, (size_t)(d + 1), 4) + P1);
            const float Cr = CM + volc_get(C, ct, x, y, d) - lastBestCr;
            bestCr = fminf(bestCr, Cr);
            *(float*)vol_at(H, (size_t)x, (size_t)y, (size_t)d, 4) += Cr;
        }
        x += dx;
        y += dy;
        lastBestCr = bestCr;
        last_c = c;
        lastMaxDisp = maxDisp;
    }
}

/* All scanlines of one direction.  Axis-aligned: exactly the launch of cu_semi_global_matching.cu:72-84
 * (one thread per column / row, fixed pathlen).  Diagonal (extension): a scanline starts at every pixel
 * of the entry edges (first row in travel direction, plus the side column the path moves away from)
 * and runs until it leaves the image. */
static void sgm_direction(const ko_volume* H, const ko_volume* C, int ct, const ko_image* left, int it,
                          int maxDisp, float P1, float P2, int dx, int dy) {
    const int w = (int)C->w, h = (int)C->h;
    if (dx == 0) {
        const int y0 = dy > 0 ? 0 : h - 1;
#pragma omp parallel for schedule(dynamic, 8)
        for (int x = 0; x < w; ++x) sgm_scanline(H, C, ct, left, it, maxDisp, P1, P2, x, y0, 0, dy, h);
    } else if (dy == 0) {
        const int x0 = dx > 0 ? 0 : w - 1;
#pragma omp parallel for schedule(dynamic, 8)
        for (int y = 0; y < h; ++y) sgm_scanline(H, C, ct, left, it, maxDisp, P1, P2, x0, y, dx, 0, w);
    } else {
        const int y0 = dy > 0 ? 0 : h - 1;
        const int xs = dx > 0 ? 0 : w - 1;
        const int n = w + h - 1;
#pragma omp parallel for schedule(dynamic, 8)
        for (int s = 0; s < n; ++s) {
            int sx, sy;
            if (s < w) { sx = s; sy = y0; }
            else { sx = xs; sy = dy > 0 ? (s - w + 1) : (h - 1 - (s - w + 1)); }
            const int lenx = dx > 0 ? (w - sx) : (sx + 1);
            const int leny = dy > 0 ? (h - sy) : (sy + 1);
            sgm_scanline(H, C, ct, left, it, maxDisp, P1, P2, sx, sy, dx, dy, mini(lenx, leny));
        }
    }
}

/* cu_semi_global_matching.cu:65-86 */
void ko_sgm(const ko_volume* volH, const ko_volume* volC, int volc_type, const ko_image* left, int img_type,
            int maxDisp, float P1, float P2, int dohoriz, int dovert, int doreverse, int dodiag) {
    /* volH.Memset(0): Volume.h:78-81 clears pitch*h*d bytes */
    memset(volH->ptr, 0, volH->pitch * volH->h * volH->d);
    if (dovert) sgm_direction(volH, volC, volc_type, left, img_type, maxDisp, P1, P2, 0, 1);
    if (dodiag) {
        sgm_direction(volH, volC, volc_type, left, img_type, maxDisp, P1, P2, 1, 1);
        sgm_direction(volH, volC, volc_type, left, img_type, maxDisp, P1, P2, -1, 1);
    }
    if (dovert && doreverse) sgm_direction(volH, volC, volc_type, left, img_type, maxDisp, P1, P2, 0, -1);
    if (dodiag && doreverse) {
        sgm_direction(volH, volC, volc_type, left, img_type, maxDisp, P1, P2, -1, -1);
        sgm_direction(volH, volC, volc_type, left, img_type, maxDisp, P1, P2, 1, -1);
    }
    if (dohoriz) {
        sgm_direction(volH, volC, volc_type, left, img_type, maxDisp, P1, P2, 1, 0);
        if (doreverse) sgm_direction(volH, volC, volc_type, left, img_type, maxDisp, P1, P2, -1, 0);
    }
}

/* ---------------------------------------------------------------- WTA ---- */

static inline double vol_get_num(const ko_volume* v, int t, int x, int y, int d) {
    switch (t) {
        case KO_VOL_F32: return *(const float*)vol_at(v, (size_t)x, (size_t)y, (size_t)d, 4);
        case KO_VOL_I32: return *(const int32_t*)vol_at(v, (size_t)x, (size_t)y, (size_t)d, 4);
        case KO_VOL_U32: return *(const uint32_t*)vol_at(v, (size_t)x, (size_t)y, (size_t)d, 4);
        case KO_VOL_U16: return *(const uint16_t*)vol_at(v, (size_t)x, (size_t)y, (size_t)d, 2);
        default: return *(const uint8_t*)vol_at(v, (size_t)x, (size_t)y, (size_t)d, 1);
    }
}

/* cu_dense_stereo.cu:25-43.  Comparison happens in Tvol; double holds every Tvol value exactly and
 * orders them identically.  bestd is a Tdisp: for char it wraps past 127 (Q: Tdisp=char overflows). */
void ko_costvol_minimum(const ko_image* disp, int disp_type, const ko_volume* vol, int vol_type,
                        unsigned maxDispVal) {
    const int w = (int)disp->w, h = (int)disp->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int bestd = 0;
            double bestc = vol_get_num(vol, vol_type, x, y, 0);
            /* min(unsigned, int): x+1 converts to unsigned; values are small positives */
            const int maxDisp = (int)((maxDispVal < (unsigned)(x + 1)) ? maxDispVal : (unsigned)(x + 1));
            for (int d = 1; d < maxDisp; ++d) {
                const double c = vol_get_num(vol, vol_type, x, y, d);
                if (c < bestc) { bestc = c; bestd = d; }
            }
            if (disp_type == KO_DISP_I8) *(int8_t*)img_at(disp, (size_t)x, (size_t)y, 1) = (int8_t)bestd;
            else *(float*)img_at(disp, (size_t)x, (size_t)y, 4) = (float)bestd;
        }
}

/* cu_dense_stereo.cu:735-755: c = sum / n (no n>0 test here: n==0 gives inf/NaN, never < bestc) */
void ko_costvol_minimum_elem(const ko_image* disp, const ko_volume* vol) {
    const int w = (int)disp->w, h = (int)disp->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float bestd = 0.0f, bestc = 1E30f;
            for (int d = 0; d < (int)vol->d; ++d) {
                const ko_costvolelem* e = (const ko_costvolelem*)vol_at(vol, (size_t)x, (size_t)y, (size_t)d, 8);
                const float c = e->sum / (float)e->n;
                if (c < bestc) { bestc = c; bestd = (float)d; }
            }
            *(float*)img_at(disp, (size_t)x, (size_t)y, 4) = bestd;
        }
}

/* cu_dense_stereo.cu:66-109 */
void ko_costvol_minimum_subpix(const ko_image* disp, const ko_volume* vol, unsigned maxDispVal, float sd,
                               const ko_image* mask) {
    const int w = (int)disp->w, h = (int)disp->h;
    const int have_mask = mask && mask->ptr;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float bestd = 0.0f, bestc = 1E10f;
            for (int d = 0; d < (int)maxDispVal; ++d) {
                const int xr = (int)((float)x + sd * (float)d);
                if (0 <= xr && xr < (int)vol->w) {
                    const float c = *(const float*)vol_at(vol, (size_t)x, (size_t)y, (size_t)d, 4);
                    if (c < bestc) { bestc = c; bestd = (float)d; }
                }
            }
            float out = bestd;
            unsigned char m = 0;
            const int bestxr = (int)((float)x + sd * bestd);
            if (0 < bestxr && bestxr < (int)vol->w - 1) {
                const float dl = bestd - 1.0f, dr = bestd + 1.0f;
                /* vol(x,y,dl): float -> size_t; the GPU's cvt.rzi.u64.f32 saturates -1 to 0 (Q7) */
                const size_t il = dl < 0.0f ? 0 : (size_t)dl;
                const size_t ir = (size_t)dr;
                if (ir >= vol->d) {
                    m = 1; /* reference reads one slice past the volume: undefined, skip the parabola */
                } else {
                    const float sl = *(const float*)vol_at(vol, (size_t)x, (size_t)y, il, 4);
                    const float sr = *(const float*)vol_at(vol, (size_t)x, (size_t)y, ir, 4);
                    const float subpixdisp = bestd - (sr - sl) / (2.0f * (sr - 2.0f * bestc + sl));
                    if (dl < subpixdisp && subpixdisp < dr) out = subpixdisp;
                }
            }
            *(float*)img_at(disp, (size_t)x, (size_t)y, 4) = out;
            if (have_mask) *(uint8_t*)img_at(mask, (size_t)x, (size_t)y, 1) = m;
        }
}

/* cu_dense_stereo.cu:122-174 (KernCostVolMinimumSquarePenaltySubpix<float,float>), IEEE evaluation in source order */
void ko_costvol_minimum_square_penalty_subpix(const ko_image* imga, const ko_volume* vol, const ko_image* imgd,
                                              unsigned maxDispVal, float sd, float lambda, float theta, const ko_image* mask) {
    const int w = (int)imga->w, h = (int)imga->h;
    const int have_mask = mask && mask->ptr;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float lastd = *(const float*)img_at(imgd, (size_t)x, (size_t)y, 4);
            const float inv2theta = 1.0f / (2.0f * theta);
            float bestd = 0.0f;
            float bestc = inv2theta * lastd * lastd + lambda * *(const float*)vol_at(vol, (size_t)x, (size_t)y, 0, 4);
            for (int d = 1; d < (int)maxDispVal; ++d) {
                const int xr = (int)((float)x + sd * (float)d);
                if (0 <= xr && xr < (int)vol->w) {
                    const float ddif = lastd - (float)d;
                    const float c = inv2theta * ddif * ddif + lambda * *(const float*)vol_at(vol, (size_t)x, (size_t)y, (size_t)d, 4);
                    if (c < bestc) { bestc = c; bestd = (float)d; }
                }
            }
            float out = bestd;
            unsigned char m = 0;
            const int bestxr = (int)((float)x + sd * bestd);
            if (0 < bestxr && bestxr < (int)vol->w - 1) {
                const float dl = bestd - 1.0f, dr = bestd + 1.0f;
                const size_t il = dl < 0.0f ? 0 : (size_t)dl;   /* cvt.rzi.u64.f32 saturates -1 to 0 (Q7) */
                const size_t ir = (size_t)dr;
                if (ir >= vol->d) {
                    m = 1; /* the reference reads one slice past the volume */
                } else {
                    const float sl = inv2theta * (lastd - dl) * (lastd - dl) + lambda * *(const float*)vol_at(vol, (size_t)x, (size_t)y, il, 4);
                    const float sr = inv2theta * (lastd - dr) * (lastd - dr) + lambda * *(const float*)vol_at(vol, (size_t)x, (size_t)y, ir, 4);
                    const float subpixdisp = bestd - (sr - sl) / (2.0f * (sr - 2.0f * bestc + sl));
                    if (dl < subpixdisp && subpixdisp < dr) out = subpixdisp;
                }
            }
            *(float*)img_at(imga, (size_t)x, (size_t)y, 4) = out;
            if (have_mask) *(uint8_t*)img_at(mask, (size_t)x, (size_t)y, 1) = m;
        }
}

/* cu_dense_stereo.cu:793-812 (KernFilterDispGrad<float,float>): out = |grad G|^2 < threshold ? in : -1, where G is what the
 * output image held before the call (Image.h:367-379 central differences, (a - b) / 2).  The sum of squares is ONE fused
 * multiply-add in the reference's SASS -- FFMA(dx, dx, dy*dy) -- restated with fmaf().  `grad` must not alias `out`.
 * Border pixels: the reference reads outside the image (undefined); out-of-image neighbours clamp to the edge here. */
void ko_filter_disp_grad(const ko_image* out, const ko_image* grad, const ko_image* in, float threshold) {
    const int w = (int)out->w, h = (int)out->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int xm = x > 0 ? x - 1 : 0, xp = x + 1 < w ? x + 1 : w - 1, ym = y > 0 ? y - 1 : 0, yp = y + 1 < h ? y + 1 : h - 1;
            const float dx = (*(const float*)img_at(grad, (size_t)xp, (size_t)y, 4) - *(const float*)img_at(grad, (size_t)xm, (size_t)y, 4)) / 2.0f;
            const float dy = (*(const float*)img_at(grad, (size_t)x, (size_t)yp, 4) - *(const float*)img_at(grad, (size_t)x, (size_t)ym, 4)) / 2.0f;
            const float m = fmaf(dx, dx, dy * dy);
            *(float*)img_at(out, (size_t)x, (size_t)y, 4) = m < threshold ? *(const float*)img_at(in, (size_t)x, (size_t)y, 4) : -1.0f;
        }
}

/* src/cu_bilateral.cu:110-143 (KernBilateralFilter<float,float,Ti2>), IEEE evaluation in source order with expf():
 * the reference's build uses the approximate ex2/rcp units, so this restatement agrees to ~1e-6 relative, not bit for bit. */
void ko_bilateral_filter_joint(const ko_image* out, const ko_image* in, const ko_image* img, int img_type, float gs, float gr,
                               float gc, int size) {
    const int w = (int)out->w, h = (int)out->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float p = *(const float*)img_at(in, (size_t)x, (size_t)y, 4);
            const float pc = img_type == KO_IMG_U8 ? (float)*(const uint8_t*)img_at(img, (size_t)x, (size_t)y, 1)
                                                   : *(const float*)img_at(img, (size_t)x, (size_t)y, 4);
            float sum = 0.0f, sumw = 0.0f;
            for (int r = -size; r <= size; ++r)
                for (int c = -size; c <= size; ++c) {
                    const int xx = x + c < 0 ? 0 : (x + c > w - 1 ? w - 1 : x + c), yy = y + r < 0 ? 0 : (y + r > h - 1 ? h - 1 : y + r);
                    const float q = *(const float*)img_at(in, (size_t)xx, (size_t)yy, 4);
                    const float qc = img_type == KO_IMG_U8 ? (float)*(const uint8_t*)img_at(img, (size_t)xx, (size_t)yy, 1)
                                                           : *(const float*)img_at(img, (size_t)xx, (size_t)yy, 4);
                    const float rd = p - q, cd = pc - qc, sd2 = (float)(r * r + c * c);
                    const float sw = expf(-(sd2) / (2 * gs * gs)), rw = expf(-(rd * rd) / (2 * gr * gr)), cw = expf(-(cd * cd) / (2 * gc * gc));
                    const float wgt = sw * rw * cw;
                    sumw += wgt;
                    sum += wgt * q;
                }
            *(float*)img_at(out, (size_t)x, (size_t)y, 4) = sumw == 0 ? p : sum / sumw;
        }
}

/* ------------------------------------------------- integral-image box / guided filter ---- */

/* src/cu_operations.cu:91-101,117-127,143-153,169-181 for float images, IEEE evaluation in source order (no contraction):
 * op 0 Multiply   out = s0*(a*b) + s1
 * op 1 Division   out = s2*(a+s0)/(b+s1) + s3
 * op 2 Square     out = (s0*a*a) + s1
 * op 3 MultiplyAdd out = s0*a*b + s1*c + s2 */
void ko_elementwise(int op, const ko_image* out, const ko_image* a, const ko_image* b, const ko_image* c, float s0, float s1,
                    float s2, float s3) {
    const int w = (int)out->w, h = (int)out->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float v1 = *(const float*)img_at(a, (size_t)x, (size_t)y, 4);
            const float v2 = b ? *(const float*)img_at(b, (size_t)x, (size_t)y, 4) : 0.0f;
            const float v3 = c ? *(const float*)img_at(c, (size_t)x, (size_t)y, 4) : 0.0f;
            float r;
            switch (op) {
            case 0: r = s0 * (v1 * v2) + s1; break;
            case 1: r = s2 * (v1 + s0) / (v2 + s1) + s3; break;
            case 2: r = (s0 * v1 * v1) + s1; break;
            default: r = s0 * v1 * v2 + s1 * v3 + s2; break;
            }
            *(float*)img_at(out, (size_t)x, (size_t)y, 4) = r;
        }
}

/* src/cu_integral_image.cu:58-107, the work-efficient exclusive scan of one row as the block executes it: an up-sweep
 * that leaves the sums of aligned power-of-two blocks in place, the root cleared, and a down-sweep that hands every
 * left child its parent's prefix and every right child prefix + left sum.  temp holds n = nextpow2 floats, n = twice
 * the block size PrefixSumRows picks (:118-121: the smallest power of two >= ceil(w/2)). */
static void prefix_sum_tree(float* out, size_t out_stride, const float* in, size_t in_stride, int w, float* temp) {
    int half = 1;
    while (half < (w + 1) / 2) half <<= 1;
    const int n = 2 * half;
    for (int i = 0; i < n; ++i) temp[i] = i < w ? in[(size_t)i * in_stride] : 0.0f;
    int offset = 1;
    for (int d = n >> 1; d > 0; d >>= 1) {
        for (int t = 0; t < d; ++t) temp[offset * (2 * t + 2) - 1] += temp[offset * (2 * t + 1) - 1];
        offset *= 2;
    }
    temp[n - 1] = 0.0f;
    for (int d = 1; d < n; d *= 2) {
        offset >>= 1;
        for (int t = 0; t < d; ++t) {
            const int ai = offset * (2 * t + 1) - 1, bi = offset * (2 * t + 2) - 1;
            const float v = temp[ai];
            temp[ai] = temp[bi];
            temp[bi] += v;
        }
    }
    for (int i = 0; i < w; ++i) out[(size_t)i * out_stride] = temp[i];
}

/* include/kangaroo/cu_integral_image.h:26-38 + src/cu_integral_image.cu:130-157: rows scanned, transposed, scanned
 * again (= the columns of the row sums), then out(x,y) = (C + A - B - D) / area over the EXCLUSIVE sums at the clamped
 * corners -- so the window is [minx, maxx) x [miny, maxy) and area = (maxx-minx)*(maxy-miny), both as the reference
 * has them.  ii: scratch of w*h floats (x-major like the reference's transposed image: ii[x*h + y]). */
static void box_filter_dense(float* out, const float* in, int w, int h, int rad, float* rows, float* ii) {
    const int n = 2 * (w > h ? w : h) + 4;
#pragma omp parallel
    {
        float* temp = (float*)malloc(sizeof(float) * 2 * (size_t)n);
#pragma omp for schedule(static)
        for (int y = 0; y < h; ++y) prefix_sum_tree(rows + (size_t)y * w, 1, in + (size_t)y * w, 1, w, temp);
#pragma omp for schedule(static)
        for (int x = 0; x < w; ++x) prefix_sum_tree(ii + (size_t)x * h, 1, rows + x, (size_t)w, h, temp);
        free(temp);
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int minx = x - rad > 0 ? x - rad : 0, maxx = x + rad < w - 1 ? x + rad : w - 1;
            const int miny = y - rad > 0 ? y - rad : 0, maxy = y + rad < h - 1 ? y + rad : h - 1;
            const int area = (maxx - minx) * (maxy - miny);
            const float A = ii[(size_t)minx * h + miny], B = ii[(size_t)maxx * h + miny];
            const float Cc = ii[(size_t)maxx * h + maxy], D = ii[(size_t)minx * h + maxy];
            const float sum = Cc + A - B - D;
            out[(size_t)y * w + x] = sum / (float)area;
        }
}

void ko_box_filter(const ko_image* out, const ko_image* in, int rad) {
    const int w = (int)out->w, h = (int)out->h;
    float* buf = (float*)calloc(4 * (size_t)w * h, sizeof(float));
    float *din = buf, *dout = buf + (size_t)w * h, *rows = dout + (size_t)w * h, *ii = rows + (size_t)w * h;
    for (int y = 0; y < h; ++y) memcpy(din + (size_t)y * w, img_at(in, 0, (size_t)y, 4), 4 * (size_t)w);
    box_filter_dense(dout, din, w, h, rad, rows, ii);
    for (int y = 0; y < h; ++y) memcpy(img_at(out, 0, (size_t)y, 4), dout + (size_t)y * w, 4 * (size_t)w);
    free(buf);
}

/* applications/stereo2/main.cpp:392-405 on the first maxDisp slices of vol, in place: ComputeMeanVarience once
 * (cu_integral_image.h:42-54), then per slice ComputeCovariance (:56-68) and GuidedFilter (:72-93); every step in the
 * reference's order with IEEE arithmetic. */
void ko_guided_filter_volume(const ko_volume* vol, const ko_image* guide, int rad, float eps, int maxDisp) {
    const int w = (int)vol->w, h = (int)vol->h;
    const size_t n = (size_t)w * h;
    float* buf = (float*)malloc(sizeof(float) * 12 * n);
    float *I = buf, *meanI = I + n, *varI = meanI + n, *P = varI + n, *meanP = P + n, *t = meanP + n, *meanIP = t + n,
          *a = meanIP + n, *b = a + n, *meana = b + n, *rows = meana + n, *ii = rows + n;
    for (int y = 0; y < h; ++y) memcpy(I + (size_t)y * w, img_at(guide, 0, (size_t)y, 4), 4 * (size_t)w);
    box_filter_dense(meanI, I, w, h, rad, rows, ii);
    for (size_t i = 0; i < n; ++i) t[i] = (1.0f * I[i] * I[i]) + 0.0f;                         /* ElementwiseSquare */
    box_filter_dense(varI, t, w, h, rad, rows, ii);                                             /* meanII */
    for (size_t i = 0; i < n; ++i) varI[i] = -1.0f * meanI[i] * meanI[i] + 1.0f * varI[i] + 0.0f; /* var = meanII - meanI^2 */
    for (int d = 0; d < maxDisp; ++d) {
        for (int y = 0; y < h; ++y) memcpy(P + (size_t)y * w, vol_at(vol, 0, (size_t)y, (size_t)d, 4), 4 * (size_t)w);
        box_filter_dense(meanP, P, w, h, rad, rows, ii);
        for (size_t i = 0; i < n; ++i) t[i] = 1.0f * (I[i] * P[i]) + 0.0f;                     /* ElementwiseMultiply */
        box_filter_dense(meanIP, t, w, h, rad, rows, ii);
        for (size_t i = 0; i < n; ++i) {
            const float cov = -1.0f * meanI[i] * meanP[i] + 1.0f * meanIP[i] + 0.0f;
            a[i] = 1.0f * (cov + 0.0f) / (varI[i] + eps) + 0.0f;                               /* Eqn. 5 */
            b[i] = -1.0f * a[i] * meanI[i] + 1.0f * meanP[i] + 0.0f;                           /* Eqn. 6 */
        }
        box_filter_dense(meana, a, w, h, rad, rows, ii);
        box_filter_dense(t, b, w, h, rad, rows, ii);                                            /* meanb */
        for (size_t i = 0; i < n; ++i) P[i] = 1.0f * meana[i] * I[i] + 1.0f * t[i] + 0.0f;     /* Eqn. 8 */
        for (int y = 0; y < h; ++y) memcpy(vol_at(vol, 0, (size_t)y, (size_t)d, 4), P + (size_t)y * w, 4 * (size_t)w);
    }
    free(buf);
}

/* ---------------------------------------------------------------- direct block matcher ---- */

/* patch_score.h:81-100 (rad 0: squared difference of the two pixels) and :257-298 (SANDPatchScore<float,rad,ImgAccessRaw>):
 * raw access -- a column index left of the image addresses the bytes that precede the row, as img(x,y) = row(y)[x] does. */
static float dense_score(const ko_image* i1, int x1, int y1, const ko_image* i2, int x2, int y2, int rad) {
    if (rad == 0) {
        const float diff = (float)((int)*(const uint8_t*)(img_at(i1, 0, (size_t)y1, 1) + x1) - (int)*(const uint8_t*)(img_at(i2, 0, (size_t)y2, 1) + x2));
        return diff * diff;
    }
    const int area = (2 * rad + 1) * (2 * rad + 1);
    float sum1 = 0.0f, sum2 = 0.0f, sad = 0.0f;
    for (int r = -rad; r <= rad; ++r)
        for (int c = -rad; c <= rad; ++c) {
            sum1 += (float)*(const uint8_t*)(img_at(i1, 0, (size_t)(y1 + r), 1) + (x1 + c));
            sum2 += (float)*(const uint8_t*)(img_at(i2, 0, (size_t)(y2 + r), 1) + (x2 + c));
        }
    const float mean1 = sum1 / (float)area, mean2 = sum2 / (float)area;
    for (int r = -rad; r <= rad; ++r)
        for (int c = -rad; c <= rad; ++c) {
            const float a = (float)*(const uint8_t*)(img_at(i1, 0, (size_t)(y1 + r), 1) + (x1 + c));
            const float b = (float)*(const uint8_t*)(img_at(i2, 0, (size_t)(y2 + r), 1) + (x2 + c));
            sad += fabsf((a - mean1) - (b - mean2));
        }
    return sad;
}

/* src/cu_dense_stereo.cu:209-253 (KernDenseStereo<TD, unsigned char, Score>, dispStep 1), TD = unsigned char (is_signed 0) or
 * char.  Border of Score::width = 2 rad + 1 pixels and every rejected pixel get InvalidValue<TD> = 0.  Candidates run from
 * max(min(maxDisp,0), -((w - width) - x)) to min(max(0,maxDisp), x + width); best and second best by strict / non-strict
 * comparison in candidate order; a best whose runner-up is more than one disparity away is dropped when
 * (second - best) / best < acceptThresh. */
void ko_dense_stereo(const ko_image* disp, const ko_image* left, const ko_image* right, int is_signed, int maxDispVal, float acceptThresh,
                     int score_rad) {
    (void)is_signed;   /* both instantiations store the same low byte; unsigned maxDisp is never negative */
    const int w = (int)left->w, h = (int)left->h, width = 2 * score_rad + 1;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int bestDisp = 0;
            if (width <= x && x < w - width && width <= y && y < h - width) {
                float bestScore = 1e36f, sndBestScore = 1e37f;
                int sndBestDisp = 0;
                int minDisp = maxDispVal < 0 ? maxDispVal : 0, maxDisp = maxDispVal > 0 ? maxDispVal : 0;
                if (minDisp < -((w - width) - x)) minDisp = -((w - width) - x);
                if (maxDisp > x + width) maxDisp = x + width;
                for (int c = minDisp; c <= maxDisp; ++c) {
                    const float score = dense_score(left, x, y, right, x - c, y, score_rad);
                    if (score < bestScore) {
                        sndBestDisp = bestDisp; sndBestScore = bestScore;
                        bestDisp = c; bestScore = score;
                    } else if (score <= sndBestScore) {
                        sndBestDisp = c; sndBestScore = score;
                    }
                }
                if (abs(bestDisp - sndBestDisp) > 1) {
                    const float cd = (sndBestScore - bestScore) / bestScore;
                    if (cd < acceptThresh) bestDisp = 0;
                }
            }
            *(int8_t*)img_at(disp, (size_t)x, (size_t)y, 1) = (int8_t)bestDisp;
        }
}

/* ---------------------------------------------------------------- subpixel refine ---- */

/* patch_score.h:257-298, SANDPatchScore<float,2,ImgAccessRaw> on unsigned char images */
static float sand5x5(const ko_image* i1, int x1, int y1, const ko_image* i2, int x2, int y2) {
    float sum1 = 0.0f, sum2 = 0.0f, sad = 0.0f;
    for (int r = -2; r <= 2; ++r)
        for (int c = -2; c <= 2; ++c) {
            sum1 += (float)*(const uint8_t*)img_at(i1, (size_t)(x1 + c), (size_t)(y1 + r), 1);
            sum2 += (float)*(const uint8_t*)img_at(i2, (size_t)(x2 + c), (size_t)(y2 + r), 1);
        }
    const float mean1 = sum1 / 25.0f, mean2 = sum2 / 25.0f;
    for (int r = -2; r <= 2; ++r)
        for (int c = -2; c <= 2; ++c) {
            const float a = (float)*(const uint8_t*)img_at(i1, (size_t)(x1 + c), (size_t)(y1 + r), 1);
            const float b = (float)*(const uint8_t*)img_at(i2, (size_t)(x2 + c), (size_t)(y2 + r), 1);
            sad += fabsf((a - mean1) - (b - mean2));
        }
    return sad;
}

/* cu_dense_stereo.cu:580-619 */
void ko_dense_stereo_subpixel_refine(const ko_image* out, const ko_image* disp, const ko_image* left,
                                     const ko_image* right, const ko_image* mask) {
    const int w = (int)disp->w, h = (int)disp->h;
    const int have_mask = mask && mask->ptr;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int bestDisp = *(const uint8_t*)img_at(disp, (size_t)x, (size_t)y, 1);
            float res = NAN;
            unsigned char m = 0;
            /* windows: left cols x-2..x+2, right cols x-bestDisp-1-2 .. x-bestDisp+1+2, rows y-2..y+2 */
            if (y < 2 || y + 2 >= h || x < 2 || x + 2 >= w || x - bestDisp - 3 < 0 || x - bestDisp + 3 >= w) {
                m = 1; /* reference: unguarded reads (Q8) */
            } else {
                const float d1 = (float)(bestDisp + 1), d2 = (float)bestDisp, d3 = (float)(bestDisp - 1);
                const float s1 = sand5x5(left, x, y, right, x - (bestDisp + 1), y);
                const float s2 = sand5x5(left, x, y, right, x - bestDisp, y);
                const float s3 = sand5x5(left, x, y, right, x - (bestDisp - 1), y);
                const float denom = (d1 - d2) * (d1 - d3) * (d2 - d3);
                const float A = (d3 * (s2 - s1) + d2 * (s1 - s3) + d1 * (s3 - s2)) / denom;
                const float B = (d3 * d3 * (s1 - s2) + d2 * d2 * (s3 - s1) + d1 * d1 * (s2 - s3)) / denom;
                const float newDisp = -B / (2.0f * A);
                if (d3 < newDisp && newDisp < d1) res = newDisp;
            }
            *(float*)img_at(out, (size_t)x, (size_t)y, 4) = res;
            if (have_mask) *(uint8_t*)img_at(mask, (size_t)x, (size_t)y, 1) = m;
        }
}

/* ---------------------------------------------------------------- front end / back end ---- */

/* cu_operations.cu:39-49: v1 = ConvertPixel<float,Tin>(a(x,y)) (plain conversion, pixel_convert.h);
 * b(x,y) = s*v1 + offset -- one FFMA in the reference's SASS (nvcc -fmad=true), fmaf here. */
void ko_elementwise_scale_bias(const ko_image* b, const ko_image* a, int in_type, float s, float offset) {
    const int w = (int)b->w, h = (int)b->h;
    const size_t es = in_type == KO_PIX_U8 ? 1 : (in_type == KO_PIX_U16 ? 2 : 4);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const void* pa = img_at(a, (size_t)x, (size_t)y, es);
            const float v1 = in_type == KO_PIX_U8 ? (float)*(const uint8_t*)pa
                           : in_type == KO_PIX_U16 ? (float)*(const uint16_t*)pa : *(const float*)pa;
            *(float*)img_at(b, (size_t)x, (size_t)y, 4) = fmaf(s, v1, offset);
        }
}

/* cu_resample.cu:53-68.  The sum runs tl, tl+1, bl, bl+1 in that order (it matters for float); the division
 * by 4.0f is exact; ConvertPixel<unsigned char>(float) truncates. */
void ko_box_half(const ko_image* out, const ko_image* in, int pix_type) {
    const int w = (int)out->w, h = (int)out->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            if (pix_type == KO_PIX_U8) {
                const uint8_t* tl = (const uint8_t*)img_at(in, (size_t)(2 * x), (size_t)(2 * y), 1);
                const uint8_t* bl = (const uint8_t*)img_at(in, (size_t)(2 * x), (size_t)(2 * y + 1), 1);
                const unsigned sum = (unsigned)tl[0] + (unsigned)tl[1] + (unsigned)bl[0] + (unsigned)bl[1];
                *(uint8_t*)img_at(out, (size_t)x, (size_t)y, 1) = (uint8_t)((float)sum / 4.0f);
            } else {
                const float* tl = (const float*)img_at(in, (size_t)(2 * x), (size_t)(2 * y), 4);
                const float* bl = (const float*)img_at(in, (size_t)(2 * x), (size_t)(2 * y + 1), 4);
                *(float*)img_at(out, (size_t)x, (size_t)y, 4) = (((tl[0] + tl[1]) + bl[0]) + bl[1]) / 4.0f;
            }
        }
}

/* cu_depth_tools.cu:15-23 (IEEE division here; the reference build uses div.approx) */
void ko_disp2depth(const ko_image* in, const ko_image* out, float fu, float baseline, float minDisp) {
    const int w = (int)out->w, h = (int)out->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float d = *(const float*)img_at(in, (size_t)x, (size_t)y, 4);
            *(float*)img_at(out, (size_t)x, (size_t)y, 4) = d >= minDisp ? fu * baseline / d : NAN;
        }
}

/* disparity.h:9-20 called from cu_dense_stereo.cu:633-639 with minDisp = MinDisparity = 0 */
void ko_disparity_image_to_vbo(const ko_image* vbo, const ko_image* disp, float baseline, float fu, float fv, float u0,
                               float v0) {
    const int w = (int)vbo->w, h = (int)vbo->h;
#pragma omp parallel for schedule(static)
    for (int v = 0; v < h; ++v)
        for (int u = 0; u < w; ++u) {
            const float d = *(const float*)img_at(disp, (size_t)u, (size_t)v, 4);
            float* P = (float*)img_at(vbo, (size_t)u, (size_t)v, 16);
            const float z = d >= 0.0f ? fu * baseline / d : NAN;
            P[0] = z * ((float)u - u0) / fu;
            P[1] = z * ((float)v - v0) / fv;
            P[2] = z;
            P[3] = 1.0f;
        }
}

/* ---------------------------------------------------------------- alternative matching cost (N4) ---- */

/* cu_dense_stereo.cu:820-840; Image.h:367-372 (GetCentralDiffDx) */
void ko_costvol_abs_and_grad(const ko_volume* vol, const ko_image* left, const ko_image* right, float sd, float alpha,
                             float r1, float r2) {
    const int w = (int)vol->w, h = (int)vol->h, dn = (int)vol->d, rw = (int)right->w;
    (void)alpha; (void)r1;   /* overwritten by the reference kernel: alpha = 0, r1 = 1e37 */
#pragma omp parallel for schedule(static) collapse(2)
    for (int d = 0; d < dn; ++d)
        for (int v = 0; v < h; ++v) {
            const float* rl = (const float*)img_at(left, 0, (size_t)v, 4);
            const float* rr = (const float*)img_at(right, 0, (size_t)v, 4);
            for (int u = 0; u < w; ++u) {
                const int r = (int)fmaf((float)d, sd, (float)u);   /* u + sd*d: one FFMA in the reference, truncated */
                float c;
                if (0 <= r && r < rw) {
                    const int rm = r > 0 ? r - 1 : 0, rp = r < rw - 1 ? r + 1 : rw - 1;
                    const int um = u > 0 ? u - 1 : 0, up = u < w - 1 ? u + 1 : w - 1;
                    const float gr = (rr[rp] - rr[rm]) * 0.5f;
                    const float grad = fabsf(fmaf(rl[up] - rl[um], -0.5f, gr));
                    const float absI = fabsf(rr[r] - rl[u]);
                    c = fmaf(0.0f, fminf(grad, r2), fminf(absI, 1e37f));
                } else {
                    c = fmaf(0.0f, r2, 1e37f);
                }
                *(float*)vol_at(vol, (size_t)u, (size_t)v, (size_t)d, 4) = c;
            }
        }
}

/* ---------------------------------------------------------------- rectification warp (N3) ---- */

static inline size_t f2size_floor(float v) {   /* cvt.rmi.u64.f32: floor, negative and NaN -> 0 */
    if (!(v >= 0.0f)) return 0;
    return (size_t)floorf(v);
}

/* cu_lookup_warp.cu:13-30 */
void ko_create_matlab_lookup_table(const ko_image* lookup, float fu, float fv, float u0, float v0, float k1, float k2) {
    const int w = (int)lookup->w, h = (int)lookup->h;
#pragma omp parallel for schedule(static)
    for (int v = 0; v < h; ++v)
        for (int u = 0; u < w; ++u) {
            const float pnu = ((float)u - u0) / fu;
            const float pnv = ((float)v - v0) / fv;
            const float r = sqrtf(pnu * pnu + pnv * pnv);
            const float rr = r * r;
            const float rf = 1 + k1 * rr + k2 * rr * rr;
            float* o = (float*)img_at(lookup, (size_t)u, (size_t)v, 8);
            o[0] = (pnu * rf) * fu + u0;
            o[1] = (pnv * rf) * fv + v0;
        }
}

/* cu_lookup_warp.cu:44-75 */
void ko_create_matlab_lookup_table_h(const ko_image* lookup, float fu, float fv, float u0, float v0, float k1, float k2,
                                     const float* H) {
    const int w = (int)lookup->w, h = (int)lookup->h;
#pragma omp parallel for schedule(static)
    for (int yi = 0; yi < h; ++yi)
        for (int xi = 0; xi < w; ++xi) {
            const float x = (float)xi, y = (float)yi;
            const float hdiv = H[6] * x + H[7] * y + H[8];
            const float u = (H[0] * x + H[1] * y + H[2]) / hdiv;
            const float v = (H[3] * x + H[4] * y + H[5]) / hdiv;
            const float pnu = (u - u0) / fu;
            const float pnv = (v - v0) / fv;
            const float r = sqrtf(pnu * pnu + pnv * pnv);
            const float rr = r * r;
            const float rf = 1 + k1 * rr + k2 * rr * rr;
            float px = (pnu * rf) * fu + u0, py = (pnv * rf) * fv + v0;
            px = fmaxf(px, 1.0f); py = fmaxf(py, 1.0f);
            px = fminf(px, (float)lookup->w - 2.0f); py = fminf(py, (float)lookup->h - 2.0f);
            float* o = (float*)img_at(lookup, (size_t)xi, (size_t)yi, 8);
            o[0] = px; o[1] = py;
        }
}

/* cu_lookup_warp.cu:85-94, Image.h:317-334 */
void ko_warp(const ko_image* out, const ko_image* in, const ko_image* lookup) {
    const int w = (int)out->w, h = (int)out->h;
    const size_t xmax = in->w - 1, ymax = in->h - 1;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float* lu = (const float*)img_at(lookup, (size_t)x, (size_t)y, 8);
            const float u = lu[0], v = lu[1];
            const float ix = floorf(u), iy = floorf(v);
            const float fx = u - ix, fy = v - iy;
            size_t x0 = f2size_floor(u), y0 = f2size_floor(v), y1 = f2size_floor(iy + 1.0f);
            size_t x1 = x0 + 1;
            if (x0 > xmax) x0 = xmax;
            if (x1 > xmax) x1 = xmax;
            if (y0 > ymax) y0 = ymax;
            if (y1 > ymax) y1 = ymax;
            const float b0 = (float)*(const uint8_t*)img_at(in, x0, y0, 1), b1 = (float)*(const uint8_t*)img_at(in, x1, y0, 1);
            const float t0 = (float)*(const uint8_t*)img_at(in, x0, y1, 1), t1 = (float)*(const uint8_t*)img_at(in, x1, y1, 1);
            const float l0 = fmaf(fx, b1 - b0, b0), l1 = fmaf(fx, t1 - t0, t0);
            const float r = fmaf(fy, l1 - l0, l0);
            uint32_t q = !(r > 0.0f) ? 0u : (r >= 4294967296.0f ? 0xffffffffu : (uint32_t)r);   /* cvt.rzi.u32.f32 */
            *(uint8_t*)img_at(out, (size_t)x, (size_t)y, 1) = (uint8_t)(q & 0xffu);
        }
}

/* ---------------------------------------------------------------- median (N1) ---- */

/* CUDA's min/max on floats (FMNMX): a NaN operand is ignored -- the other operand is returned -- and -0 < +0. */
static float fmnmx_min(float a, float b) {
    if (isnan(a)) return b;
    if (isnan(b)) return a;
    if (a == b) return signbit(a) ? a : b;
    return a < b ? a : b;
}
static float fmnmx_max(float a, float b) {
    if (isnan(a)) return b;
    if (isnan(b)) return a;
    if (a == b) return signbit(a) ? b : a;
    return a > b ? a : b;
}

/* The exchange network of cu_median.cu:177-199 / :240-263 / :306-333, generated instead of transcribed.  It is the bitonic
 * sorting network for n inputs -- for every block size k = 2, 4, .. a "flip" stage pairing i with i ^ (k-1), then
 * half-cleaners pairing i with i + j for j = k/4 .. 1 -- with the comparators that would reach past input n-1 dropped, and
 * then every comparator that cannot influence outputs n/2 .. n-1 removed (the reference's "only top half are guaranteed
 * valid": the median index (n + bad)/2 never lies below n/2).  tests/test_oracle_golden.py checks the generated sequence
 * against the reference file when /root/reference is present: 155 / 439 / 968 comparators for 25 / 49 / 81, same order. */
typedef struct { unsigned char a, b; } ko_cmp;
static int median_network(int n, ko_cmp* net) {
    static ko_cmp all[4096];
    static int stage_end[64];
    int N = 1, total = 0, nstage = 0;
    while (N < n) N <<= 1;
    for (int k = 2; k <= N; k <<= 1) {
        for (int i = 0; i < N; ++i) {
            const int l = i ^ (k - 1);
            if (l > i && l < n) { all[total].a = (unsigned char)i; all[total].b = (unsigned char)l; ++total; }
        }
        stage_end[nstage++] = total;
        for (int j = k >> 2; j >= 1; j >>= 1) {
            for (int i = 0; i < N; ++i)
                if (!(i & j) && i + j < n) { all[total].a = (unsigned char)i; all[total].b = (unsigned char)(i + j); ++total; }
            stage_end[nstage++] = total;
        }
    }
    static unsigned char keep[4096];
    unsigned char need[128] = {0};
    for (int i = n / 2; i < n; ++i) need[i] = 1;
    for (int s = nstage - 1; s >= 0; --s) {
        const int lo = s ? stage_end[s - 1] : 0, hi = stage_end[s];
        for (int c = lo; c < hi; ++c) keep[c] = need[all[c].a] || need[all[c].b];
        for (int c = lo; c < hi; ++c)
            if (keep[c]) { need[all[c].a] = 1; need[all[c].b] = 1; }
    }
    int m = 0;
    for (int c = 0; c < total; ++c)
        if (keep[c]) net[m++] = all[c];
    return m;
}
int ko_median_network(int size, unsigned char* pairs) {   /* for the tests: 2 bytes per comparator, returns the count */
    static ko_cmp net[4096];
    const int m = median_network(size * size, net);
    for (int c = 0; c < m; ++c) { pairs[2 * c] = net[c].a; pairs[2 * c + 1] = net[c].b; }
    return m;
}

/* cu_median.cu:160-207 (5x5), :217-273 (7x7), :283-342 (9x9): gather with GetWithClampedRange in the reference's order
 * (v[(dX + r) * size + (dY + r)], column-major), count the invalid samples (InvalidValue<float>::IsValid = isfinite), run the
 * exchange network s2(a,b): a = min(a,b), b = max(a_old,b) on the raw samples -- min/max ignore NaNs, so a NaN is
 * overwritten by a copy of its partner -- and return v[(size^2 + bad)/2].  Without invalid samples that is the exact
 * median; with them it is whatever the network leaves at that index, reproduced here exactly. */
void ko_median_filter_reject_negative(const ko_image* out, const ko_image* in, int size, int maxbad) {
    const int w = (int)out->w, h = (int)out->h, krad = size / 2, kpix = size * size;
    static ko_cmp nets[3][1024];
    static int nnet[3];
    const int slot = size == 5 ? 0 : (size == 7 ? 1 : 2);
#pragma omp critical(ko_median_net)
    if (!nnet[slot]) nnet[slot] = median_network(kpix, nets[slot]);
    const ko_cmp* net = nets[slot];
    const int m = nnet[slot];
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float v[81];
            int bad = 0;
            for (int dx = -krad; dx <= krad; ++dx)
                for (int dy = -krad; dy <= krad; ++dy) {
                    const int xx = x + dx < 0 ? 0 : (x + dx > w - 1 ? w - 1 : x + dx);
                    const int yy = y + dy < 0 ? 0 : (y + dy > h - 1 ? h - 1 : y + dy);
                    const float s = *(const float*)img_at(in, (size_t)xx, (size_t)yy, 4);
                    v[(dx + krad) * size + (dy + krad)] = s;
                    if (!isfinite(s)) ++bad;
                }
            float r = NAN;
            if (bad < maxbad && bad < kpix) {
                for (int c = 0; c < m; ++c) {
                    const float a = v[net[c].a], b = v[net[c].b];
                    v[net[c].a] = fmnmx_min(a, b);
                    v[net[c].b] = fmnmx_max(a, b);
                }
                r = v[(kpix + bad) / 2];
            }
            *(float*)img_at(out, (size_t)x, (size_t)y, 4) = r;
        }
}

/* ---------------------------------------------------------------- left-right check ---- */

/* cu_dense_stereo.cu:512-532 with TD=float; InvalidValue<float>: NaN / isfinite (InvalidValue.h:18-47) */
void ko_left_right_check_f32(const ko_image* dispL, const ko_image* dispR, float sd, float maxDiff) {
    const int w = (int)dispL->w, h = (int)dispL->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float* pl = (float*)img_at(dispL, (size_t)x, (size_t)y, 4);
            const float dl = *pl;
            const float xr = (float)x + sd * dl;
            if (0.0f <= xr && xr < (float)dispR->w) {
                const float dr = *(const float*)img_at(dispR, (size_t)xr, (size_t)y, 4);
                if (!isfinite(dr) || fabsf(dl - dr) > maxDiff) *pl = NAN;
            } else {
                *pl = NAN;
            }
        }
}

/* float -> char on the GPU: the sm_100a PTX of the reference is cvt.rzi.ftz.s32.f32 followed by
 * cvt.s64.s8, i.e. truncate to int32 (saturating) and keep the low 8 bits sign-extended (wraps). */
static inline int8_t wrap_i8(float v) {
    int32_t i;
    if (!(v == v)) i = 0;
    else if (v <= -2147483648.0f) i = INT32_MIN;
    else if (v >= 2147483648.0f) i = INT32_MAX;
    else i = (int32_t)v;
    return (int8_t)(uint8_t)((uint32_t)i & 0xFFu);
}

/* cu_dense_stereo.cu:512-539 with TD=char: xr is computed in char (wraps for x > 127); InvalidValue<char>::IsValid(v) = !v
 * while Value() = 0 (InvalidValue.h:50-59, Q10). */
void ko_left_right_check_i8(const ko_image* dispL, const ko_image* dispR, int sdi, int maxDiffi) {
    const int w = (int)dispL->w, h = (int)dispL->h;
    const float sd = (float)sdi, maxDiff = (float)maxDiffi;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int8_t* pl = (int8_t*)img_at(dispL, (size_t)x, (size_t)y, 1);
            const int8_t dl = *pl;
            const int8_t xr = wrap_i8((float)x + sd * (float)dl);
            if (0 <= xr && (size_t)xr < dispR->w) {
                const int8_t dr = *(const int8_t*)img_at(dispR, (size_t)xr, (size_t)y, 1);
                const int valid = !(uint8_t)dr;
                if (!valid || (float)abs((int)dl - (int)dr) > maxDiff) *pl = 0;
            } else {
                *pl = 0;
            }
        }
}

/* ---------------------------------------------------------------- whole path ---- */

/* applications/stereo2/main.cpp:375-454 with use_census, no cost-volume filters, no median:
 *   img = u8/255 (ElementwiseScaleBias, :376) ; Census (:380) ; CensusStereoVolume sd=-1 (:384) and,
 *   for the LR check, sd=+1 (:385) ; SemiGlobalMatching<float,float,float> on vol[0] only (:423-427) ;
 *   CostVolMinimumSubpix / CostVolMinimum<float,float> on both (:430-436) ;
 *   LeftRightCheck(disp[1],disp[0],+1) then LeftRightCheck(disp[0],disp[1],-1) (:451-454). */
int ko_pipeline_u8(const uint8_t* left, const uint8_t* right, int w, int h, int maxDisp, int window,
                   int popc_mode, float P1, float P2, int dohoriz, int dovert, int doreverse, int dodiag,
                   int subpix, int lrcheck, float lr_maxdiff, float* disp_out, float* volH_out) {
    const int words = window == KO_WIN_9x7 ? 1 : (window == KO_WIN_11x11 ? 2 : 4);
    const size_t npx = (size_t)w * (size_t)h;
    float* imgf[2] = {(float*)malloc(npx * 4), (float*)malloc(npx * 4)};
    uint64_t* cen[2] = {(uint64_t*)malloc(npx * 8 * (size_t)words), (uint64_t*)malloc(npx * 8 * (size_t)words)};
    float* volC = (float*)malloc(npx * (size_t)maxDisp * 4);
    float* volH = volH_out ? volH_out : (float*)malloc(npx * (size_t)maxDisp * 4);
    float* dispR = lrcheck ? (float*)malloc(npx * 4) : NULL;
    if (!imgf[0] || !imgf[1] || !cen[0] || !cen[1] || !volC || !volH || (lrcheck && !dispR)) return -1;
    const uint8_t* src[2] = {left, right};
    ko_image fi[2], ci[2];
    for (int i = 0; i < 2; ++i) {
        for (size_t k = 0; k < npx; ++k) imgf[i][k] = (float)src[i][k] * (1.0f / 255.0f);
        fi[i] = (ko_image){(size_t)w * 4, imgf[i], (size_t)w, (size_t)h};
        ci[i] = (ko_image){(size_t)w * 8 * (size_t)words, cen[i], (size_t)w, (size_t)h};
        ko_census(&ci[i], &fi[i], window, KO_IMG_F32);
    }
    ko_volume vC = {(size_t)w * 4, volC, (size_t)w, (size_t)h, npx * 4, (size_t)maxDisp};
    ko_volume vH = {(size_t)w * 4, volH, (size_t)w, (size_t)h, npx * 4, (size_t)maxDisp};
    ko_image dL = {(size_t)w * 4, disp_out, (size_t)w, (size_t)h};
    ko_image dR = {(size_t)w * 4, dispR, (size_t)w, (size_t)h};
    if (lrcheck) { /* right-reference volume is NOT aggregated (main.cpp:424 loops i<1) */
        ko_census_stereo_volume(&vC, &ci[1], &ci[0], words, KO_VOL_F32, maxDisp, +1.0f, popc_mode);
        if (subpix) ko_costvol_minimum_subpix(&dR, &vC, (unsigned)maxDisp, +1.0f, NULL);
        else ko_costvol_minimum(&dR, KO_DISP_F32, &vC, KO_VOL_F32, (unsigned)maxDisp);
    }
    ko_census_stereo_volume(&vC, &ci[0], &ci[1], words, KO_VOL_F32, maxDisp, -1.0f, popc_mode);
    const ko_volume* vfinal = &vC;
    if (dohoriz || dovert || dodiag) {
        ko_sgm(&vH, &vC, KO_VOL_F32, &fi[0], KO_IMG_F32, maxDisp, P1, P2, dohoriz, dovert, doreverse, dodiag);
        vfinal = &vH; /* the app copies vol[2] back into vol[0] (:426) */
    }
    if (subpix) ko_costvol_minimum_subpix(&dL, vfinal, (unsigned)maxDisp, -1.0f, NULL);
    else ko_costvol_minimum(&dL, KO_DISP_F32, vfinal, KO_VOL_F32, (unsigned)maxDisp);
    if (lrcheck) {
        ko_left_right_check_f32(&dR, &dL, +1.0f, lr_maxdiff);
        ko_left_right_check_f32(&dL, &dR, -1.0f, lr_maxdiff);
    }
    free(imgf[0]); free(imgf[1]); free(cen[0]); free(cen[1]); free(volC);
    if (!volH_out) free(volH);
    if (dispR) free(dispR);
    return 0;
}
