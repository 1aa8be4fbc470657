/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the census / semi-global-matching
 * path of arpg/Kangaroo.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path (kangaroo_b200/, libroo_b200.so)
 * never links, imports or calls it.
 *
 * Parity pinning: the reference ships no tests or golden vectors.  This
 * restatement is pinned against outputs of the UNMODIFIED reference kernels (oracle/_ref, built by
 * oracle/Makefile from /root/reference/src) run on a B200; those outputs are committed as
 * tests/golden/ (.npz files) together with the generating script tests/golden/make_golden.py, and
 * tests/test_oracle_golden.py checks every function below against them.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Images / volumes use the reference's own pitched layout (include/kangaroo/Image.h:617-620,
 * Volume.h:363-369): element (x,y,z) at (char*)ptr + z*img_pitch + y*pitch + x*sizeof(T).
 */
#ifndef KANGAROO_ORACLE_H
#define KANGAROO_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { size_t pitch; void* ptr; size_t w; size_t h; } ko_image;
typedef struct { size_t pitch; void* ptr; size_t w; size_t h; size_t img_pitch; size_t d; } ko_volume;

/* include/kangaroo/CostVolElem.h:10-19 */
typedef struct { int32_t n; float sum; } ko_costvolelem;

enum { KO_WIN_9x7 = 0, KO_WIN_11x11 = 1, KO_WIN_16x16 = 2 };
enum { KO_IMG_U8 = 0, KO_IMG_F32 = 1 };
/* popcount mode: 0 = the reference's 32-bit __popc on 64-bit words (hamming_distance.h:40-62,
 * low 32 bits of every word only), 1 = full 64-bit popcount (extension) */
enum { KO_POPC32_COMPAT = 0, KO_POPC64 = 1 };
enum { KO_VOL_U16 = 0, KO_VOL_F32 = 1, KO_VOL_I32 = 2, KO_VOL_U32 = 3, KO_VOL_U8 = 4, KO_VOL_ELEM = 5 };
enum { KO_DISP_I8 = 0, KO_DISP_F32 = 1 };

int ko_num_threads(void);
/* torchrun exports OMP_NUM_THREADS=1 for multi-rank launches: the CPU arm sets the thread count explicitly */
void ko_set_num_threads(int n);

/* src/cu_census.cu:18-46 (9x7), :52-110 (11x11), :116-177 (16x16); out holds 1/2/4 uint64 per px */
void ko_census(const ko_image* out, const ko_image* in, int window, int in_type);

/* include/kangaroo/hamming_distance.h:40-62 */
unsigned ko_hamming(const uint64_t* p, const uint64_t* q, int words, int popc_mode);

/* src/cu_census.cu:226-266 */
void ko_census_stereo(const ko_image* disp_i8, const ko_image* left, const ko_image* right, int maxDisp);

/* src/cu_census.cu:272-314 ; vol_type in {KO_VOL_U16, KO_VOL_F32} */
void ko_census_stereo_volume(const ko_volume* vol, const ko_image* left, const ko_image* right, int words,
                             int vol_type, int maxDisp, float sd, int popc_mode);

/* src/cu_semi_global_matching.cu:21-89.  volc_type in {KO_VOL_F32, KO_VOL_ELEM}; img_type in
 * {KO_IMG_U8, KO_IMG_F32}.  Path order: down, [down-right, down-left], up, [up-left, up-right],
 * right, left; the bracketed diagonal sweeps (dodiag) are an extension with the same per-path
 * kernel body -- with dodiag == 0 this is exactly the reference. */
void ko_sgm(const ko_volume* volH, const ko_volume* volC, int volc_type, const ko_image* left, int img_type,
            int maxDisp, float P1, float P2, int dohoriz, int dovert, int doreverse, int dodiag);

/* src/cu_dense_stereo.cu:25-60 (guarded: the reference kernel has no bounds test) */
void ko_costvol_minimum(const ko_image* disp, int disp_type, const ko_volume* vol, int vol_type, unsigned maxDisp);

/* src/cu_dense_stereo.cu:735-763 */
void ko_costvol_minimum_elem(const ko_image* disp_f32, const ko_volume* vol_elem);

/* src/cu_dense_stereo.cu:66-116.  mask_u8 (optional, may be NULL / ptr==NULL) receives 1 where the
 * reference reads out of bounds (bestd+1 == vol.d) and this restatement skipped the parabola. */
void ko_costvol_minimum_subpix(const ko_image* disp_f32, const ko_volume* vol_f32, unsigned maxDisp, float sd,
                               const ko_image* mask_u8);

/* src/cu_dense_stereo.cu:580-627 + include/kangaroo/patch_score.h:257-298.  Pixels whose 5x5 windows
 * leave the images (reference: unguarded reads) get NaN and mask 1. */
void ko_dense_stereo_subpixel_refine(const ko_image* out_f32, const ko_image* disp_u8, const ko_image* left_u8,
                                     const ko_image* right_u8, const ko_image* mask_u8);

/* ---- callers either side of the path (N3 front end, N2 back end) ---- */

enum { KO_PIX_U8 = 0, KO_PIX_F32 = 1, KO_PIX_U16 = 2 };

/* src/cu_operations.cu:39-57,260-262: b = s*a + offset, Tout = Tup = float, Tin = in_type.  The
 * reference build (nvcc default -fmad=true) contracts s*a+offset into one fused multiply-add; so does this. */
void ko_elementwise_scale_bias(const ko_image* b_f32, const ko_image* a, int in_type, float s, float offset);

/* src/cu_resample.cu:53-83: out(x,y) = (in(2x,2y) + in(2x+1,2y) + in(2x,2y+1) + in(2x+1,2y+1)) / 4.0f, summed
 * in unsigned int (u8, result truncated back to u8) or float (f32); pix_type in {KO_PIX_U8, KO_PIX_F32} */
void ko_box_half(const ko_image* out, const ko_image* in, int pix_type);

/* src/cu_depth_tools.cu:15-30: out = in >= minDisp ? fu*baseline / in : NaN */
void ko_disp2depth(const ko_image* in_f32, const ko_image* out_f32, float fu, float baseline, float minDisp);

/* src/cu_dense_stereo.cu:633-646 + include/kangaroo/disparity.h:9-20 (MinDisparity = 0, cu_dense_stereo.cu:15);
 * vbo holds float4 {x, y, z, 1} per pixel */
void ko_disparity_image_to_vbo(const ko_image* vbo_f32x4, const ko_image* disp_f32, float baseline, float fu, float fv,
                               float u0, float v0);

/* src/cu_dense_stereo.cu:820-848 (N4, the non-census matching cost of both applications).  The kernel overrides its
 * arguments: alpha = 0, r1 = 1e37 (:829-830), so for r = (int)(u + sd*d) inside the right image
 * vol(u,v,d) = fma(0, min(|grad difference|, r2), min(|R(r,v) - L(u,v)|, 1e37)) -- the absolute difference unless the
 * gradient term is not finite -- and fma(0, r2, 1e37) outside.  The gradient taps row[x-1], row[x+1] are unguarded in
 * the reference; here they are clamped into the row (they only matter through 0 * non-finite). */
void ko_costvol_abs_and_grad(const ko_volume* vol_f32, const ko_image* left_f32, const ko_image* right_f32, float sd,
                             float alpha, float r1, float r2);

/* src/cu_lookup_warp.cu:13-38 (the variant without homography): radial distortion lookup for roo::Warp,
 * lookup(u,v) = (pnu*rf*fu + u0, pnv*rf*fv + v0), pn = ((u-u0)/fu, (v-v0)/fv), r = sqrt(pn.pn), rf = 1 + k1 r^2 + k2 r^4.
 * IEEE here; the reference build uses reciprocal / square-root approximations and contractions. */
void ko_create_matlab_lookup_table(const ko_image* lookup_f32x2, float fu, float fv, float u0, float v0, float k1, float k2);

/* src/cu_lookup_warp.cu:44-83: the same table after the homography H_on (row-major 3x3, new image -> original), clamped
 * to [1, w-2] x [1, h-2] (:69-73). */
void ko_create_matlab_lookup_table_h(const ko_image* lookup_f32x2, float fu, float fv, float u0, float v0, float k1, float k2,
                                     const float* H_on);

/* src/cu_lookup_warp.cu:85-106 + Image.h:317-334 (GetBilinear): out(x,y) = (unsigned char) bilinear sample of `in` at
 * lookup(x,y) = (u, v).  lerp(a,b,t) = a + t*(b-a), one fused multiply-add each as in the reference build; the float
 * result is truncated to unsigned 32 bit and its low byte stored.  Row / column indices come from float -> size_t
 * conversions (negative saturates to 0); the reference reads unguarded, this clamps the taps into the image. */
void ko_warp(const ko_image* out_u8, const ko_image* in_u8, const ko_image* lookup_f32x2);

/* src/cu_median.cu:160-350 (MedianFilterRejectNegative5x5 / 7x7 / 9x9), OUT OF PLACE.  size in {5,7,9}.
 * Window = clamp-to-edge neighbourhood (Image.h:298-303); bad = number of non-finite samples
 * (InvalidValue<float>::IsValid = isfinite, InvalidValue.h:18-47); out = NaN unless bad < maxbad && bad < size^2.
 * The window is gathered in the reference's order, run through its exchange network (the bitonic network for size^2 inputs
 * with dead comparators removed, generated in kangaroo_oracle.c, on min/max that ignore NaNs) and v[(size^2+bad)/2] is
 * returned: the exact median for windows without invalid samples, and for windows with them exactly the sample the
 * reference's partially sorted array holds there.  tests/golden/median.npz pins both, bit for bit. */
void ko_median_filter_reject_negative(const ko_image* out_f32, const ko_image* in_f32, int size, int maxbad);
/* the generated comparator sequence (2 bytes per comparator into pairs, capacity >= 2048 bytes); returns the count */
int ko_median_network(int size, unsigned char* pairs);

/* src/cu_dense_stereo.cu:512-546 */
/* src/cu_dense_stereo.cu:122-174; mask (optional, u8): 1 where the reference reads slice vol.d (Q7) */
void ko_costvol_minimum_square_penalty_subpix(const ko_image* imga_f32, const ko_volume* vol_f32, const ko_image* imgd_f32,
                                              unsigned maxDisp, float sd, float lambda, float theta, const ko_image* mask);
/* src/cu_dense_stereo.cu:793-812; grad = the image whose gradient gates the output (the reference uses the output image's
 * own previous contents), must not alias out */
void ko_filter_disp_grad(const ko_image* out_f32, const ko_image* grad_f32, const ko_image* in_f32, float threshold);
/* src/cu_bilateral.cu:110-143, float in/out, guide image of KO_IMG_U8 or KO_IMG_F32; out must not alias in */
void ko_bilateral_filter_joint(const ko_image* out_f32, const ko_image* in_f32, const ko_image* img, int img_type, float gs,
                               float gr, float gc, int size);
/* src/cu_operations.cu:91-181, float images: op 0 Multiply s0*(a*b)+s1, 1 Division s2*(a+s0)/(b+s1)+s3, 2 Square (s0*a*a)+s1,
 * 3 MultiplyAdd s0*a*b + s1*c + s2 (b, c may be NULL where unused) */
void ko_elementwise(int op, const ko_image* out_f32, const ko_image* a, const ko_image* b, const ko_image* c, float s0, float s1,
                    float s2, float s3);
/* include/kangaroo/cu_integral_image.h:26-38, src/cu_integral_image.cu:58-157: box mean through two tree-ordered exclusive
 * prefix sums; the window is [x-rad, x+rad) x [y-rad, y+rad) clamped, as the reference has it */
void ko_box_filter(const ko_image* out_f32, const ko_image* in_f32, int rad);
/* applications/stereo2/main.cpp:392-405 (cu_integral_image.h:42-93): guided filtering of the first maxDisp slices, in place */
void ko_guided_filter_volume(const ko_volume* vol_f32, const ko_image* guide_f32, int rad, float eps, int maxDisp);
/* src/cu_dense_stereo.cu:209-253,376-406: DenseStereo<{unsigned char, char}, unsigned char>, score_rad 0..7 */
void ko_dense_stereo(const ko_image* disp_8, const ko_image* left_u8, const ko_image* right_u8, int is_signed, int maxDisp,
                     float acceptThresh, int score_rad);
void ko_left_right_check_f32(const ko_image* dispL, const ko_image* dispR, float sd, float maxDiff);
void ko_left_right_check_i8(const ko_image* dispL, const ko_image* dispR, int sd, int maxDiff);

/* Whole path as applications/stereo2/main.cpp:380-454 runs it (census -> volume(s) -> SGM -> WTA/subpix
 * -> LR check), on tightly packed host arrays; used as the CPU baseline.  scratch volumes are
 * allocated inside.  Returns 0 on success. */
int ko_pipeline_u8(const uint8_t* left, const uint8_t* right, int w, int h, int maxDisp, int window,
                   int popc_mode, float P1, float P2, int dohoriz, int dovert, int doreverse, int dodiag,
                   int subpix, int lrcheck, float lr_maxdiff, float* disp_out, float* volH_out /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif
