/* roo_b200.h -- C ABI of the B200-native census / semi-global-matching engine.
 *
 * This is the drop-in boundary for the one hot path of arpg/Kangaroo this repository rebuilds
 * (SURVEY.md section 8b).  The reference exposes that path as C++ free functions in namespace roo
 * taking roo::Image / roo::Volume by value (include/kangaroo/cu_census.h:12-38,
 * cu_semi_global_matching.h:10-12, cu_dense_stereo.h:13-47,81-85); those types have user-provided
 * destructors, so a C ABI cannot take them directly.  Each entry point below names the reference
 * function it replaces (paths relative to /root/reference); include/kangaroo_b200/roo.hpp
 * re-creates the exact roo:: overloads on top of these, and INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - roo_image_t / roo_volume_t are binary-identical to roo::Image<T> / roo::Volume<T>
 *    (Image.h:617-620, Volume.h:363-369): element (x,y,z) lives at
 *    (char*)ptr + z*img_pitch + y*pitch + x*sizeof(T); any pitch is honoured (sub-views work).
 *  - Every pointer is DEVICE memory on the current device unless a name says `host`.
 *  - `stream` is a cudaStream_t passed as void*; NULL (the legacy default stream) reproduces the
 *    reference's ordering.  All calls are asynchronous like the reference's launchers.
 *  - Operators never allocate or free user-visible memory (reference: same).  roo_sgm() and the
 *    engine use stream-ordered scratch from the CUDA memory pool.
 *  - Return value: ROO_OK (0), a positive cudaError_t, or a negative roo_status code.  The reference
 *    launchers return void and never check errors; roo.hpp ignores the code unless
 *    ROO_B200_THROW is defined.
 *  - There is no CPU fallback anywhere behind this header.
 */
#ifndef ROO_B200_H
#define ROO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct roo_image_t { size_t pitch; void* ptr; size_t w; size_t h; } roo_image_t;
typedef struct roo_volume_t { size_t pitch; void* ptr; size_t w; size_t h; size_t img_pitch; size_t d; } roo_volume_t;
/* include/kangaroo/CostVolElem.h:10-19 */
typedef struct roo_costvolelem_t { int32_t n; float sum; } roo_costvolelem_t;

enum roo_status {
    ROO_OK = 0,
    ROO_ERR_INVALID_ARGUMENT = -1,
    ROO_ERR_UNSUPPORTED = -2,     /* e.g. maxDisp > 512 or a non-integer disparity step sd */
    ROO_ERR_OUT_OF_MEMORY = -3,
    ROO_ERR_NO_DEVICE = -4
};

enum roo_window { ROO_WIN_9x7 = 0, ROO_WIN_11x11 = 1, ROO_WIN_16x16 = 2 };   /* -> 1 / 2 / 4 uint64 per pixel */
enum roo_img_type { ROO_IMG_U8 = 0, ROO_IMG_F32 = 1 };
/* ROO_POPC32_COMPAT reproduces the reference's 32-bit __popc on 64-bit words
 * (hamming_distance.h:40-62: only the low 32 bits of every word are compared). */
enum roo_popc_mode { ROO_POPC32_COMPAT = 0, ROO_POPC64 = 1 };
enum roo_vol_type { ROO_VOL_U16 = 0, ROO_VOL_F32 = 1, ROO_VOL_I32 = 2, ROO_VOL_U32 = 3, ROO_VOL_U8 = 4, ROO_VOL_ELEM = 5 };
enum roo_disp_type { ROO_DISP_I8 = 0, ROO_DISP_F32 = 1 };

const char* roo_b200_version(void);
const char* roo_status_string(int status);
/* Number of kernels this library has launched in this process (all threads), for bench.py. */
unsigned long long roo_launch_count(void);
/* 0 (default): divisions as the reference's -use_fast_math build (div.approx.ftz) -> results bit-identical to
 * the reference kernels; 1: IEEE division -> bit-identical to the CPU oracle.  Process-wide DEFAULT: the granular
 * operators (whose signatures are the reference's and carry no mode) read it at call time; an engine takes its own
 * mode from roo_pipeline_params_t.fp_mode when it is created, so engines in different modes can run side by side. */
void roo_set_ieee_division(int on);
/* Development knobs for A/B measurements (never needed for correct results).  ROO_TUNE_HSWEEP: 1 (default) runs the
 * horizontal aggregation paths through the bulk-copy kernel (sgm_hsweep.cu), 0 through the generic sweep kernel. */
enum roo_tuning_knob { ROO_TUNE_HSWEEP = 0,
                       /* 1 (default): passes that can recompute the matching cost from the census words do so and do
                        * not read the u8 cost volume; 0: always through the materialised volume */
                       ROO_TUNE_INSWEEP_COST = 1,
                       /* roo_split_engine: CTAs per SM of a sweep that crosses strips (0 = as many as fit; default 3).  A
                        * small number makes the grid run in waves, so a strip hands its first scanlines on early */
                       ROO_TUNE_STRIP_CTAS_PER_SM = 2,
                       /* roo_guided_filter_volume: scratch budget in MiB (default 2048); the slices go through in chunks
                        * of budget / (4 fp32 planes) -- a small value exercises the chunk loop on small volumes */
                       ROO_TUNE_GUIDED_SCRATCH_MIB = 3,
                       /* 1 (default): a launch of ONE pair at 256 disparities runs its fused vertical passes with 12 warps x 2
                        * columns per band instead of 8 x 3 (more warps per SM while a band waits for its predecessor) */
                       ROO_TUNE_SOLO_GEOMETRY = 4 };
int roo_set_tuning(int knob, int value);

/* ---- granular operators: one per reference launcher -------------------------------------- */

/* roo::Census x6 (cu_census.h:13-23; cu_census.cu:180-220).  census holds 1/2/4 uint64 per pixel. */
int roo_census(const roo_image_t* census, const roo_image_t* img, int window, int in_type, void* stream);

/* roo::CensusStereo (cu_census.h:33; cu_census.cu:226-266).  unsigned long descriptors -> char disparity. */
int roo_census_stereo(const roo_image_t* disp_i8, const roo_image_t* left, const roo_image_t* right, int maxDisp,
                      void* stream);

/* roo::CensusStereoVolume<Tvol,T> (cu_census.h:36-38; cu_census.cu:272-314).
 * words in {1,2,4}; vol_type in {ROO_VOL_U16, ROO_VOL_F32}; sd must be -1 or +1. */
int roo_census_stereo_volume(const roo_volume_t* vol, const roo_image_t* left, const roo_image_t* right, int words,
                             int vol_type, int maxDisp, float sd, int popc_mode, void* stream);

/* roo::SemiGlobalMatching<TH,TC,Timg> (cu_semi_global_matching.h:10-12; .cu:21-89).
 * volH float; volc_type in {ROO_VOL_F32, ROO_VOL_ELEM}; img_type in {ROO_IMG_U8, ROO_IMG_F32};
 * 1 <= maxDisp <= 512 (above 256 one pass per path: the fused vertical groups hold up to 256 disparities).  dodiag = 0 is the reference (paths down, up, right, left in that order);
 * dodiag = 1 adds the four diagonal paths (extension). */
int roo_sgm(const roo_volume_t* volH, const roo_volume_t* volC, int volc_type, const roo_image_t* left, int img_type,
            int maxDisp, float P1, float P2, int dohoriz, int dovert, int doreverse, int dodiag, void* stream);

/* roo::CostVolMinimum<Tdisp,Tvol> (cu_dense_stereo.h:13-15; .cu:25-60); bounds-guarded. */
int roo_costvol_minimum(const roo_image_t* disp, int disp_type, const roo_volume_t* vol, int vol_type,
                        unsigned maxDisp, void* stream);

/* roo::CostVolMinimum(Image<float>, Volume<CostVolElem>) (cu_dense_stereo.h:81-82; .cu:735-763). */
int roo_costvol_minimum_elem(const roo_image_t* disp_f32, const roo_volume_t* vol_elem, void* stream);

/* roo::CostVolMinimumSubpix (cu_dense_stereo.h:84-85; .cu:66-116); sd must be -1 or +1.
 * Where the reference reads slice bestd+1 == vol.d (out of bounds) the integer disparity is kept. */
int roo_costvol_minimum_subpix(const roo_image_t* disp_f32, const roo_volume_t* vol_f32, unsigned maxDisp, float sd,
                               void* stream);

/* roo::DenseStereoSubpixelRefine (cu_dense_stereo.h:45-47; .cu:580-627).  Pixels whose 5x5 windows
 * leave the images (undefined in the reference) get NaN. */
int roo_dense_stereo_subpixel_refine(const roo_image_t* out_f32, const roo_image_t* disp_u8,
                                     const roo_image_t* left_u8, const roo_image_t* right_u8, void* stream);

/* roo::LeftRightCheck (cu_dense_stereo.h:37-41; .cu:512-546); in place on dispL. */
int roo_left_right_check_f32(const roo_image_t* dispL, const roo_image_t* dispR, float sd, float maxDiff, void* stream);
int roo_left_right_check_i8(const roo_image_t* dispL, const roo_image_t* dispR, int sd, int maxDiff, void* stream);

/* ---- callers either side of the path (N3 front end, N2 back end) ------------------------------- */

enum roo_pix_type { ROO_PIX_U8 = 0, ROO_PIX_F32 = 1, ROO_PIX_U16 = 2 };

/* roo::ElementwiseScaleBias<float, {unsigned char, unsigned short, float}, float> (cu_operations.h:14-15;
 * cu_operations.cu:39-57,260-262): b = s*a + offset (one fused multiply-add, as in the reference build).
 * applications/stereo2/main.cpp:376 calls it with s = 1/255 to feed Census / SemiGlobalMatching. */
int roo_elementwise_scale_bias(const roo_image_t* b_f32, const roo_image_t* a, int in_type, float s, float offset,
                               void* stream);

/* roo::BoxHalf<unsigned char,unsigned int,unsigned char> / <float,float,float> (reduce.h:7-8;
 * cu_resample.cu:53-83): out(x,y) = mean of in(2x..2x+1, 2y..2y+1); one level of BoxReduce (reduce.h:35-46).
 * `in` must cover 2*out.w x 2*out.h (the reference reads it unguarded). */
int roo_box_half(const roo_image_t* out, const roo_image_t* in, int pix_type, void* stream);

/* roo::CreateMatlabLookupTable(lookup, fu, fv, u0, v0, k1, k2) (cu_lookup_warp.cu:13-38): the radial-distortion table
 * roo::Warp consumes; one-time setup.  The second form is the overload that first applies the homography H_on
 * (Mat<float,9>, row-major 3x3, a HOST pointer here) and clamps the positions to [1, w-2] x [1, h-2] (:44-83). */
int roo_create_matlab_lookup_table(const roo_image_t* lookup_f32x2, float fu, float fv, float u0, float v0, float k1,
                                   float k2, void* stream);
int roo_create_matlab_lookup_table_homography(const roo_image_t* lookup_f32x2, float fu, float fv, float u0, float v0,
                                              float k1, float k2, const float* H_on, void* stream);

/* roo::Warp (cu_lookup_warp.h; cu_lookup_warp.cu:85-106): out(x,y) = bilinear sample of `in` at lookup(x,y) (float2
 * pixel coordinates), the rectification step of applications/stereo2/main.cpp:362-365.  Taps outside `in` are clamped
 * (the reference reads them unguarded). */
int roo_warp(const roo_image_t* out_u8, const roo_image_t* in_u8, const roo_image_t* lookup_f32x2, void* stream);

/* roo::Disp2Depth (cu_depth_tools.h:11; cu_depth_tools.cu:15-30): out = in >= minDisp ? fu*baseline/in : NaN. */
int roo_disp2depth(const roo_image_t* in_f32, const roo_image_t* out_f32, float fu, float baseline, float minDisp,
                   void* stream);

/* roo::DisparityImageToVbo (cu_dense_stereo.h; cu_dense_stereo.cu:633-646; disparity.h:9-20): vbo = float4
 * {z*(u-u0)/fu, z*(v-v0)/fv, z, 1} with z = disp >= 0 ? fu*baseline/disp : NaN.  vbo rows must be 16-byte aligned. */
int roo_disparity_image_to_vbo(const roo_image_t* vbo_f32x4, const roo_image_t* disp_f32, float baseline, float fu,
                               float fv, float u0, float v0, void* stream);

/* roo::CostVolumeFromStereoTruncatedAbsAndGrad (cu_dense_stereo.h:66; cu_dense_stereo.cu:820-848), the non-census
 * matching cost of both applications (stereo2/main.cpp:387-388): a float volume for roo_sgm / roo_costvol_minimum*.
 * As in the reference, alpha and r1 are ignored (its kernel overwrites them with 0 and 1e37): the cost is the absolute
 * intensity difference, 1e37 where the right pixel is outside the image.  Bounds-guarded (the reference is not). */
int roo_costvol_from_stereo_truncated_abs_and_grad(const roo_volume_t* vol_f32, const roo_image_t* left_f32,
                                                   const roo_image_t* right_f32, float sd, float alpha, float r1, float r2,
                                                   void* stream);

/* roo::MedianFilterRejectNegative5x5 / 7x7 / 9x9 (cu_median.h:19-32; cu_median.cu:160-350), size in {5,7,9}: NaN unless
 * fewer than maxbad (and not all) samples of the clamp-to-edge window are non-finite, else element (size^2 + bad)/2 of the
 * window after the reference's exchange network: the exact median for windows without invalid samples, and with them the
 * same comparator-order-dependent near-median the reference returns -- bit-identical to the reference
 * kernels for every input.  `out` may alias `in` (both reference applications call it in place, where the reference itself races):
 * an overlapping call filters into a stream-ordered temporary and copies back, i.e. gives the out-of-place result. */
int roo_median_filter_reject_negative(const roo_image_t* out_f32, const roo_image_t* in_f32, int size, int maxbad,
                                      void* stream);

/* roo::FilterDispGrad (cu_dense_stereo.h:101-103; cu_dense_stereo.cu:793-812; stereo2/main.cpp:457):
 * out(x,y) = dx^2 + dy^2 < threshold ? in(x,y) : -1 with dx, dy the central differences of what `out` holds when the call
 * is made.  The reference reads the image it is writing (meaningful only in place, where it races with itself); here
 * the gradient source is a snapshot of `out`, so in-place calls (out == in, as in the applications) are deterministic.
 * On the border pixels the reference reads outside the image (undefined); here those neighbours clamp to the edge. */
int roo_filter_disp_grad(const roo_image_t* out_f32, const roo_image_t* in_f32, float threshold, void* stream);

/* roo::CostVolMinimumSquarePenaltySubpix (cu_dense_stereo.h:87-89; cu_dense_stereo.cu:122-174): argmin over d of
 * (imgd(x,y) - d)^2 / (2 theta) + lambda * vol(x,y,d) with the parabola refinement of CostVolMinimumSubpix on the
 * penalised costs; same guards as roo_costvol_minimum_subpix (Q7), sd must be +-1 (Q12). */
int roo_costvol_minimum_square_penalty_subpix(const roo_image_t* imga_f32, const roo_volume_t* vol_f32,
                                              const roo_image_t* imgd_f32, unsigned maxDisp, float sd, float lambda,
                                              float theta, void* stream);

/* roo::BilateralFilter<float,float,Timg>(dOut, dIn, dImg, gs, gr, gc, size) -- joint bilateral filter with spatial, range and
 * guide-image weights (cu_bilateral.h:18-22; cu_bilateral.cu:110-155), Timg = unsigned char or float.  Clamp-to-edge window
 * of (2 size + 1)^2 taps; arithmetic = the reference's fast-math SASS (bit-identical results).  out must not overlap in. */
int roo_bilateral_filter_joint(const roo_image_t* out_f32, const roo_image_t* in_f32, const roo_image_t* img, int img_type,
                               float gs, float gr, float gc, unsigned size, void* stream);
/* The applications filter a cost volume with it slice by slice (stereo2/main.cpp:407-421: per disparity a device copy of
 * the slice and one launch).  This is the same result for the first maxDisp slices in ONE launch, out of place. */
int roo_bilateral_filter_volume(const roo_volume_t* out_f32, const roo_volume_t* in_f32, const roo_image_t* img, int img_type,
                                float gs, float gr, float gc, unsigned size, int maxDisp, void* stream);

/* roo::DenseStereo<TDisp, unsigned char>(dDisp, dCamLeft, dCamRight, maxDisp, acceptThresh, score_rad) -- the direct block
 * matcher (cu_dense_stereo.h:24-28; cu_dense_stereo.cu:209-253,376-406).  disp_signed: 0 = TDisp unsigned char, 1 = char
 * (negative maxDisp searches the other way).  score_rad 0 = squared pixel difference, 1..7 = SANDPatchScore<float,rad>.
 * Pixels within 2 rad + 1 of the border and rejected matches get 0.  maxDisp 255 (127 for char) hangs the reference (its
 * candidate counter wraps) and returns ROO_ERR_UNSUPPORTED here; any width (the reference: w <= 1024).  Candidates that
 * reach left of the image read the bytes preceding the row, as the reference's raw access does. */
int roo_dense_stereo(const roo_image_t* disp_8, int disp_signed, const roo_image_t* left_u8, const roo_image_t* right_u8, int maxDisp,
                     float acceptThresh, int score_rad, void* stream);

/* ---- integral-image box filter and guided filter (gfilter.cu) ------------------------------ */

/* roo::ElementwiseMultiply / Division / Square / MultiplyAdd for float images (cu_operations.h:22-35; cu_operations.cu:85-190):
 *   c = scalar*(a*b) + offset;  c = scalar*(a+sa)/(b+sb) + offset;  b = scalar*a*a + offset;  d = sab*a*b + sc*c + offset.
 * Default fp mode = the reference's fast-math SASS forms (bit-identical to its kernels), IEEE mode = source order. */
int roo_elementwise_multiply(const roo_image_t* c_f32, const roo_image_t* a_f32, const roo_image_t* b_f32, float scalar, float offset,
                             void* stream);
int roo_elementwise_division(const roo_image_t* c_f32, const roo_image_t* a_f32, const roo_image_t* b_f32, float sa, float sb,
                             float scalar, float offset, void* stream);
int roo_elementwise_square(const roo_image_t* b_f32, const roo_image_t* a_f32, float scalar, float offset, void* stream);
int roo_elementwise_multiply_add(const roo_image_t* d_f32, const roo_image_t* a_f32, const roo_image_t* b_f32, const roo_image_t* c_f32,
                                 float sab, float sc, float offset, void* stream);

/* roo::BoxFilter<float,float,float>(out, in, scratch, rad) (cu_integral_image.h:26-38): box mean through two exclusive prefix
 * sums in the reference's tree order and its four-corner lookup (window [x-rad, x+rad) x [y-rad, y+rad), clamped; divisor =
 * that window's area) -- bit-identical to the reference kernels.  No scratch image (stream-ordered internal scratch), any
 * w <= 16384 and h <= 65536 (the reference: w, h <= 2048); out may be in. */
int roo_box_filter(const roo_image_t* out_f32, const roo_image_t* in_f32, int rad, void* stream);

/* The applications' guided filtering of a cost volume (stereo2/main.cpp:392-405): ComputeMeanVarience(I) once, then per slice
 * ComputeCovariance + GuidedFilter (cu_integral_image.h:42-93), in place on the first maxDisp slices -- here 3 launches for
 * the guide image + 5 per chunk of slices (one chunk unless 4 fp32 copies of the chunk exceed 2 GiB) instead of 37 per
 * slice.  Same results as that sequence of reference calls, bit for bit. */
int roo_guided_filter_volume(const roo_volume_t* vol_f32, const roo_image_t* guide_f32, int rad, float eps, int maxDisp, void* stream);
/* Both take their scratch stream-ordered from a memory pool of this library that keeps it between calls (4 fp32 copies of
 * the chunk of slices); this returns the current device's unused scratch to the driver. */
int roo_release_scratch(void);

/* ---- fused engine: the whole per-frame path of applications/stereo2/main.cpp:375-454 ------- */

typedef struct roo_engine roo_engine_t;

/* Divisions of the path (adaptive P2, subpixel parabola): ROO_FP_REFERENCE reproduces the SASS of the reference's
 * -use_fast_math build (results bit-identical to its kernels), ROO_FP_IEEE uses IEEE division (bit-identical to the
 * CPU oracle).  ROO_FP_DEFAULT = whatever roo_set_ieee_division() says when the engine is created. */
enum roo_fp_mode { ROO_FP_DEFAULT = 0, ROO_FP_REFERENCE = 1, ROO_FP_IEEE = 2 };

typedef struct roo_pipeline_params_t {
    int w, h;             /* image size (any; not limited to 1024 like the reference, Q4) */
    int max_disp;         /* 1..512 (257..512: one pass per path, no fused vertical groups) */
    int window;           /* enum roo_window */
    int popc_mode;        /* enum roo_popc_mode */
    float P1, P2;         /* stereo2/main.cpp:246-247 defaults: 0.01, 0.02 */
    float img_scale;      /* adaptive-P2 intensity = u8 * img_scale; 1/255 matches main.cpp:376, 1 matches the uchar instantiation */
    int dohoriz, dovert, doreverse, dodiag;
    int subpix;           /* 0: CostVolMinimum<float,float>; 1: CostVolMinimumSubpix */
    int lrcheck;          /* 1: right-reference WTA on the un-aggregated volume + both LeftRightChecks (main.cpp:385,432,451-454) */
    float lr_maxdiff;
    int max_batch;        /* stereo pairs in flight per call (scratch is sized for this many) */
    int keep_volume;      /* 1: the last sweep also writes the aggregate so roo_engine_export_volume() works */
    int fuse_vertical;    /* a vertical path and its two diagonals aggregated in ONE pass: 1 always, -1 never, 0 (default) unless the group is a large frame in too few pairs to fill the GPU (one or two pairs at 1280x720x128 run faster with one pass per path) */
    int median_size;      /* 0 (none), 5, 7 or 9: MedianFilterRejectNegativeNxN on the disparities between WTA and the */
    int median_maxbad;    /*   left-right check, median_iters times, on both disparity images when lrcheck is set      */
    int median_iters;     /*   (main.cpp:438-444; out of place into engine scratch, so without the reference's race)   */
    int fp_mode;          /* enum roo_fp_mode: floating-point mode of THIS engine (two engines may differ) */
    float filtgrad_threshold; /* > 0: FilterDispGrad(disp, disp, threshold) as the last stage (main.cpp:456-458); 0: off */
} roo_pipeline_params_t;

/* The engine allocates its scratch on the CURRENT device; later calls must come with that device current (else
 * ROO_ERR_INVALID_ARGUMENT).  An engine runs one group at a time on its scratch: use it from one host thread, and do
 * not mix roo_engine_run_device with groups still in flight from roo_engine_submit_host. */
int roo_engine_create(roo_engine_t** out, const roo_pipeline_params_t* params);
int roo_engine_destroy(roo_engine_t* e);
size_t roo_engine_scratch_bytes(const roo_engine_t* e);

/* Optional front end, the steps before Census in applications/stereo2/main.cpp:360-375: the engine then takes RAW
 * frames of (w << level) x (h << level) pixels, rectifies them through the two float2 lookup tables (roo::Warp; both
 * NULL = already rectified) and reduces them `level` times with BoxHalf<uchar,uint,uchar> (BoxReduce) to its working
 * size w x h.  The tables stay caller-owned device memory and must outlive the engine.  Call once, before the first
 * run; every later run_device / run_host / submit_host call passes frames of the raw size.  (Not available through
 * roo_multi_engine_*.) */
int roo_engine_set_front_end(roo_engine_t* e, int level, const roo_image_t* lookup_left_f32x2,
                             const roo_image_t* lookup_right_f32x2);

/* n_pairs tightly packed (h x w) uint8 images each side, device memory; disp: n_pairs x h x w float.
 * Processes the pairs in groups of max_batch on `stream`. */
int roo_engine_run_device(roo_engine_t* e, const uint8_t* left, const uint8_t* right, float* disp, int n_pairs,
                          void* stream);
/* Same with HOST buffers (pinned memory recommended): uploads, runs, downloads, and synchronises. */
int roo_engine_run_host(roo_engine_t* e, const uint8_t* left_host, const uint8_t* right_host, float* disp_host,
                        int n_pairs);
/* Asynchronous form of roo_engine_run_host for callers that stream frames: enqueue one group (n_pairs <= max_batch)
 * and return; roo_engine_wait(ticket) blocks until that group's disparities are in disp_host.  Two groups may be in
 * flight: the upload of group t+1 and the download of group t-1 overlap the computation of group t (a third submit
 * first waits, on the host, for the group two tickets back).  The host buffers must stay valid (and should be
 * pinned) until the wait returns.  Not thread-safe per engine, like the rest of the engine API. */
int roo_engine_submit_host(roo_engine_t* e, const uint8_t* left_host, const uint8_t* right_host, float* disp_host,
                           int n_pairs, long long* ticket);
int roo_engine_wait(roo_engine_t* e, long long ticket);
/* Copies the aggregated volume of batch slot `slot` from the last run into a roo::Volume<float>
 * (d >= max_disp slices are left untouched); for parity tests. */
int roo_engine_export_volume(roo_engine_t* e, int slot, const roo_volume_t* volH, void* stream);
/* Same for the census descriptors of the last run: side 0 = left, 1 = right. */
int roo_engine_export_census(roo_engine_t* e, int slot, int side, const roo_image_t* census, void* stream);

/* ---- multi-GPU: the batch is sharded across the GPUs of the box by pair index (one engine + one host thread per
 * device, no collective: pairs are independent).  devices == NULL / n_devices <= 0: all visible devices. */
typedef struct roo_multi_engine roo_multi_engine_t;
int roo_multi_engine_create(roo_multi_engine_t** out, const roo_pipeline_params_t* params, const int* devices, int n_devices);
int roo_multi_engine_destroy(roo_multi_engine_t* m);
int roo_multi_engine_device_count(const roo_multi_engine_t* m);
int roo_multi_engine_run_host(roo_multi_engine_t* m, const uint8_t* left_host, const uint8_t* right_host, float* disp_host,
                              int n_pairs);

/* ---- multi-GPU, ONE pair: row-strip split with halo hand-off over NVLink (BASELINE config 5).
 * GPU k owns a strip of image rows (its part of the aggregate never moves); the six paths that travel in y hand their
 * state from strip to strip through peer memory (st.release.sys / ld.acquire.sys on the exported record), horizontal paths,
 * winner-takes-all and the left-right check are strip-local.  No collective.  Results are bit-identical to roo_engine_*.
 * devices == NULL / n_devices <= 0: all visible devices, one strip each.  A device may be listed more than once (strips that
 * share a device are ordered by CUDA events instead of in-kernel polling).  median_size and filtgrad_threshold must be 0. */
typedef struct roo_split_engine roo_split_engine_t;
int roo_split_engine_create(roo_split_engine_t** out, const roo_pipeline_params_t* params, const int* devices, int n_devices);
int roo_split_engine_destroy(roo_split_engine_t* e);
int roo_split_engine_strip_count(const roo_split_engine_t* e);
/* one pair: whole (h x w) uint8 frames in host memory (pinned recommended) -> (h x w) float disparities in host memory */
int roo_split_engine_run_host(roo_split_engine_t* e, const uint8_t* left_host, const uint8_t* right_host, float* disp_host);
/* device time of the last frame (census .. left-right check, max over strips, CUDA events) and bytes handed between strips */
int roo_split_engine_last_stats(const roo_split_engine_t* e, float* device_ms, unsigned long long* exchanged_bytes);

/* Per-kernel device timing for bench.py: with profiling on, the engine records a CUDA event after
 * every launch on the launching stream; roo_engine_get_profile (call after synchronising) returns the
 * accumulated milliseconds and launch counts per kernel kind since profiling was switched on. */
enum roo_prof_kind { ROO_PROF_CENSUS = 0, ROO_PROF_COST = 1, ROO_PROF_SWEEP = 2, ROO_PROF_WTA = 3, ROO_PROF_LRCHECK = 4,
                     ROO_PROF_VGROUP = 5,
                     ROO_PROF_PASS0 = 6,   /* ROO_PROF_PASS0 + i: the i-th aggregation pass of the plan (also counted
                                              under SWEEP / VGROUP), i < 8 */
                     ROO_PROF_KINDS = 14 };
int roo_engine_set_profiling(roo_engine_t* e, int on);
int roo_engine_get_profile(roo_engine_t* e, double* ms_by_kind, long long* launches_by_kind);
/* development aid: in-kernel cycle counters of a -DVG_TIMING build (zeros in a normal build) */
int roo_engine_debug_counters(roo_engine_t* e, unsigned long long* out, int n, int reset);

#ifdef __cplusplus
}
#endif
#endif /* ROO_B200_H */
