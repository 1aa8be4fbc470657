// Internal launch API shared between the translation units of libroo_b200.
#pragma once
#include "common.cuh"

namespace roo_b200 {

// ---- census.cu ----
int launch_census(char* out, size_t out_pitch, size_t out_batch, const char* in, size_t in_pitch, size_t in_batch,
                  int w, int h, int batch, int window, int in_type, cudaStream_t st);
int launch_cost_u8(unsigned char* c8, const void* cl, const void* cr, int w, int h, int batch, int DP, int maxDisp,
                   int words, int popc_mode, cudaStream_t st);
int launch_census_wta(float* disp, const void* cself, const void* cother, int w, int h, int batch, int maxDisp,
                      int words, int popc_mode, int subpix, int sdi, int ieee, cudaStream_t st);

// ---- sgm.cu ----
enum EpiKind { EPI_NONE = 0, EPI_WTA_WRITE = 1, EPI_WTA_ONLY = 2 };

// One aggregation sweep (one path direction) over `batch` pairs on the internal layout
// H[pair][y][x][DP] (fp32, disparity innermost, DP = 32 * ceil(maxDisp/32) rounded to 32/64/128/256/512).
struct SweepArgs {
    float* H;            // in/out aggregate
    size_t h_pair;       // elements between pairs
    const void* C;       // cost, same indexing as H: float or unsigned char
    size_t c_pair;       // elements between pairs
    const float* img;    // adaptive-P2 intensity image, tightly packed fp32 [pair][y][x] (launch_image_to_f32)
    size_t img_pair;     // elements between pairs
    float cost_scale;    // U8 cost: cost = count * cost_scale
    int w, h, DP, maxDisp, batch;
    float P1, P2;
    int dx, dy;
    int first;           // 1: H holds nothing yet (treated as 0, not read)
    int cost_kind;       // CostKind (sgm_step.cuh): 0 fp32, 1 u8 Hamming counts
    int epi;             // EpiKind
    int subpix;          // epilogue: 0 CostVolMinimum<float,float>, 1 CostVolMinimumSubpix(sd=-1)
    float* disp;         // epilogue output [pair][y][x]
    size_t disp_pair;    // elements
    int ieee;            // 0: the reference's fast-math divisions (bit-identical to its kernels), 1: IEEE (== CPU oracle)
    // cost_kind == COST_CEN32: one-word census descriptors [pair][y][x] (u64, low half used) instead of C
    const unsigned long long* cenL;
    const unsigned long long* cenR;
    size_t cen_pair;     // elements between pairs
    // Row-strip split of ONE pair across GPUs (roo_split_engine): this launch covers a strip of image rows;
    // a path that enters the strip through its first row (in travel direction) continues from the state the upstream
    // strip exported for the pixel it comes from, and a path that leaves through the last row exports its state for the
    // downstream strip.  Records of strip_rec_floats(DP) floats, indexed by the x of the boundary-row pixel; the last
    // word is a sequence number published with st.release.sys (the record lives in the CONSUMER's memory, written
    // over NVLink) and polled with ld.acquire.sys.  nullptr: the strip edge is the image edge.
    const float* strip_import;   // local memory, written by the upstream GPU
    float* strip_export;         // peer memory of the downstream GPU
    int strip_seq;               // value that marks the records of THIS sweep
};
inline size_t strip_rec_floats(int DP) { return (size_t)DP + 4; }   // [DP aggregate row | lastBest | pixel | - | seq]
int launch_sweep(const SweepArgs& a, cudaStream_t st);
// sgm_hsweep.cu: the horizontal paths (a.dy == 0), bulk-copy prefetch
int launch_hsweep(const SweepArgs& a, cudaStream_t st);
// development knob (roo_set_tuning): 0 = horizontal paths through the generic sweep kernel
extern std::atomic<int> g_use_hsweep;
// development knob: 0 = always materialise the u8 cost volume (no in-sweep cost from census words)
extern std::atomic<int> g_insweep_cost;
// split engine: CTAs per SM of a strip-crossing sweep (0 = as many as fit); fewer = waves = earlier hand-offs
extern std::atomic<int> g_strip_ctas_per_sm;
// sgm_fused.cu: 1 = a single pair at 256 disparities runs 12 warps x 2 columns per band (development knob)
extern std::atomic<int> g_solo_geometry;
// gfilter.cu: scratch budget of roo_guided_filter_volume in MiB
extern std::atomic<int> g_guided_scratch_mib;
int launch_image_to_f32(float* dst, const void* src, size_t pitch, size_t src_pair, int img_type, int w, int h,
                        int batch, float scale, cudaStream_t st);
// 512: single-path sweeps only (the fused vertical groups hold three state rows per column in registers: up to 256)
inline int disp_padded(int maxDisp) { return maxDisp <= 32 ? 32 : (maxDisp <= 64 ? 64 : (maxDisp <= 128 ? 128 : (maxDisp <= 256 ? 256 : 512))); }
constexpr int ROO_MAX_DISP = 512, ROO_MAX_DISP_FUSED = 256;

// ---- sgm_fused.cu: the three paths sharing a y travel direction in one pass ----
struct VGroupArgs {
    float* H; size_t h_pair;
    const void* C; size_t c_pair;
    const float* img; size_t img_pair;
    float cost_scale;
    int w, h, maxDisp, batch;
    float P1, P2;
    int fwd;            // 1: (0,+1),(+1,+1),(-1,+1) ; 0: (0,-1),(-1,-1),(+1,-1)
    int ieee;           // fp mode, as SweepArgs::ieee
    const unsigned long long* cenL;   // COST_CEN32: census descriptors (padded arrays), as SweepArgs
    const unsigned long long* cenR;
    size_t cen_pair;
    float* edge_hp;     // [pair][band][h][3][DP]  states handed from band b to band b+1
    float* edge_sc;     // [pair][band][h][8]
    int* progress;      // [pair][band] rows published
    int* ticket;        // CTA start counter of this launch (zeroed with the flags)
    int n_bands;
};
size_t vgroup_edge_floats(int w, int h, int DP);   // per pair
int vgroup_bands(int w, int h, int DP);
int launch_vgroup(const SweepArgs& s, int fwd, float* edge, int* progress, cudaStream_t st);

// Execution plan of one SemiGlobalMatching call: passes in the reference's order (down, up, right, left;
// cu_semi_global_matching.cu:71-85), the diagonal extension inserted after the vertical path of the same
// travel direction.  A vertical path and its two diagonals become ONE fused pass when `fuse` is set.
struct SgmPass { int fused; int dx, dy; };   // fused: dy = travel direction, dx unused
struct SgmPlan { int n; SgmPass pass[8]; };
SgmPlan sgm_plan(int dohoriz, int dovert, int doreverse, int dodiag, int fuse);
// runs pass i of the plan; a.first / a.epi are set by the caller
int launch_pass(SweepArgs a, const SgmPass& pass, float* edge, int* progress, cudaStream_t st);

// direction list in execution order (reference order down, up, right, left; diagonals are an extension)
int launch_internal_to_vol(const roo_volume_t* dst, const float* src, int DP, int maxDisp, cudaStream_t st);
int sgm_directions(int dohoriz, int dovert, int doreverse, int dodiag, int dxs[8], int dys[8]);

// ---- wta.cu ----
// ---- frontback.cu: batched front-end stages of the engine (tightly packed u8 images)
int launch_warp_u8(unsigned char* out, const unsigned char* in, int w, int h, int batch, const roo_image_t& lookup,
                   cudaStream_t st);
int launch_box_half_u8(unsigned char* out, const unsigned char* in, int w_out, int h_out, int w_in, int h_in, int batch,
                       cudaStream_t st);
// MedianFilterRejectNegative{5,7,9} over a batch of images (out must not overlap in)
int launch_median(float* out, size_t out_pitch, size_t out_batch, const float* in, size_t in_pitch, size_t in_batch, int w,
                  int h, int batch, int size, int maxbad, cudaStream_t st);
// FilterDispGrad over `batch` images (byte strides between them); grad must not overlap out
int launch_filter_disp_grad(const roo_image_t& out, const roo_image_t& grad, const roo_image_t& in, float threshold, cudaStream_t st,
                            int batch = 1, size_t out_pair = 0, size_t grad_pair = 0, size_t in_pair = 0);
int launch_lr_check_f32(float* dispL, size_t pitchL, const float* dispR, size_t pitchR, int w, int h, int batch,
                        size_t batchL, size_t batchR, float sd, float maxDiff, cudaStream_t st);

}  // namespace roo_b200
