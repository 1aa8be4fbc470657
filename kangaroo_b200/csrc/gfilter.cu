// Integral-image box filter and the guided cost-volume filter of the applications.
//
// Reference: roo::BoxFilter<float,float,float> (include/kangaroo/cu_integral_image.h:26-38) = PrefixSumRows -> Transpose ->
// PrefixSumRows -> BoxFilterIntegralImage (src/cu_integral_image.cu:15-157); ComputeMeanVarience / ComputeCovariance /
// GuidedFilter (cu_integral_image.h:42-93) compose it with the float elementwise operators of src/cu_operations.cu:85-190.
// The applications run that composition once per disparity slice of a cost volume (applications/stereo2/main.cpp:392-405:
// per slice 4 box filters = 16 launches + 5 elementwise launches, one w/2-thread block per image row, w and h <= 2048).
//
// What has to be kept: a box sum is the difference of four large fp32 prefix sums, so the result depends on the ORDER of the
// reference's scan at the 1e-4 relative level.  Its scan is the work-efficient tree: the sums of aligned power-of-two blocks
// (balanced pairwise, up-sweep), then prefix(i) = ((0 + S_top) + ...) + S_low over the aligned blocks named by the set bits
// of i from the most significant down (down-sweep).  That order does not need the tree's storage: a pairwise-summation
// stack (one partial sum per level, a binary counter) holds exactly those block sums when element i arrives.  So here
//   * scan_rows_kernel: one warp per image row, 256 elements per step (8 per lane in registers, five shuffle levels),
//     the levels above 256 on the per-warp stack -- any width, coalesced; P and I*P (I and I*I) come from one read, and a, b
//     are computed on the fly from the integral images of P and I*P, never written;
//   * scan_cols_kernel: one thread per column streaming down the rows, eight rows per step, in place -- this replaces
//     Transpose + the second PrefixSumRows (the transposed image is never written);
//   * box_epilogue_kernel: the four-corner lookup fused with the elementwise algebra that follows it in the guided filter;
// and the whole volume goes through 5 launches per chunk of slices (all slices at c2's size) instead of 37 per slice.
// Results are bit-identical to the reference kernels (tests/golden/guided.npz) in the default fp mode -- every operation
// is the reference's -use_fast_math SASS form (FADD/FMUL/FFMA.FTZ, MUFU.RCP) -- and to the CPU oracle in IEEE mode.
#include "common.cuh"
#include "ftz.cuh"
#include "kernels.cuh"

#include <mutex>

namespace roo_b200 {

// ---- arithmetic in the two fp modes ---------------------------------------------------------------------------------
template <bool IEEE> __device__ __forceinline__ float gadd(float a, float b) { return IEEE ? __fadd_rn(a, b) : fadd_ftz(a, b); }
template <bool IEEE> __device__ __forceinline__ float gmul(float a, float b) { return IEEE ? __fmul_rn(a, b) : fmul_ftz(a, b); }

// cu_operations.cu:91-101.  SASS: FMUL.FTZ t = a*b; FFMA.FTZ(t, scalar, offset)
template <bool IEEE> __device__ __forceinline__ float ew_multiply(float a, float b, float scalar, float offset) {
    return IEEE ? __fadd_rn(__fmul_rn(scalar, __fmul_rn(a, b)), offset) : ffma_ftz(fmul_ftz(a, b), scalar, offset);
}
// cu_operations.cu:117-127.  SASS: FADD.FTZ den = b+sb; MUFU.RCP; FADD.FTZ n = a+sa; FMUL.FTZ n *= scalar; FFMA.FTZ(n, rcp, offset)
template <bool IEEE> __device__ __forceinline__ float ew_division(float a, float b, float sa, float sb, float scalar, float offset) {
    if (IEEE) return __fadd_rn(__fdiv_rn(__fmul_rn(scalar, __fadd_rn(a, sa)), __fadd_rn(b, sb)), offset);
    const float r = rcp_approx_ftz(fadd_ftz(b, sb));
    return ffma_ftz(fmul_ftz(fadd_ftz(a, sa), scalar), r, offset);
}
// cu_operations.cu:143-153.  SASS: FMUL.FTZ s = a*scalar; FFMA.FTZ(a, s, offset)
template <bool IEEE> __device__ __forceinline__ float ew_square(float a, float scalar, float offset) {
    return IEEE ? __fadd_rn(__fmul_rn(__fmul_rn(scalar, a), a), offset) : ffma_ftz(a, fmul_ftz(a, scalar), offset);
}
// cu_operations.cu:169-181.  SASS: FMUL.FTZ t = a*sab; FMUL.FTZ u = c*sc; FFMA.FTZ r = b*t + u; FADD.FTZ r + offset
template <bool IEEE> __device__ __forceinline__ float ew_multiply_add(float a, float b, float c, float sab, float sc, float offset) {
    if (IEEE) return __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(sab, a), b), __fmul_rn(sc, c)), offset);
    return fadd_ftz(ffma_ftz(b, fmul_ftz(a, sab), fmul_ftz(c, sc)), offset);
}

// ---- standalone elementwise operators -------------------------------------------------------------------------------
enum EwOp { EW_MULTIPLY = 0, EW_DIVISION = 1, EW_SQUARE = 2, EW_MULTIPLY_ADD = 3 };

template <int OP, bool IEEE>
__global__ void __launch_bounds__(256)
elementwise_kernel(Img<float> out, Img<float> a, Img<float> b, Img<float> c, float s0, float s1, float s2, float s3) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= out.w) return;
    float r;
    if (OP == EW_MULTIPLY) r = ew_multiply<IEEE>(a(x, y), b(x, y), s0, s1);
    else if (OP == EW_DIVISION) r = ew_division<IEEE>(a(x, y), b(x, y), s0, s1, s2, s3);
    else if (OP == EW_SQUARE) r = ew_square<IEEE>(a(x, y), s0, s1);
    else r = ew_multiply_add<IEEE>(a(x, y), b(x, y), c(x, y), s0, s1, s2);
    out(x, y) = r;
}

template <int OP>
static int launch_elementwise(const roo_image_t& out, const roo_image_t& a, const roo_image_t& b, const roo_image_t& c, float s0,
                              float s1, float s2, float s3, cudaStream_t st) {
    dim3 grid(cdiv((int)out.w, 256), (unsigned)out.h);
    if (g_ieee_div.load())
        elementwise_kernel<OP, true><<<grid, 256, 0, st>>>(Img<float>(out), Img<float>(a), Img<float>(b), Img<float>(c), s0, s1, s2, s3);
    else
        elementwise_kernel<OP, false><<<grid, 256, 0, st>>>(Img<float>(out), Img<float>(a), Img<float>(b), Img<float>(c), s0, s1, s2, s3);
    count_launch();
    return launch_status();
}

// ---- the reference's scan order on eight consecutive elements -------------------------------------------------------
// Eight elements of one aligned block: the block sums of the three levels inside it (pairwise, as the up-sweep forms them),
// the eight exclusive prefixes continued from `base` (the prefix of the block's first element), and the block's sum.
template <bool IEEE>
__device__ __forceinline__ float tree8_sum(const float (&v)[8], float (&t1)[4], float (&t2)[2]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) t1[k] = gadd<IEEE>(v[2 * k + 1], v[2 * k]);
    t2[0] = gadd<IEEE>(t1[1], t1[0]);
    t2[1] = gadd<IEEE>(t1[3], t1[2]);
    return gadd<IEEE>(t2[1], t2[0]);
}
template <bool IEEE>
__device__ __forceinline__ void tree8_prefix(float base, const float (&v)[8], const float (&t1)[4], const float (&t2)[2], float (&o)[8]) {
    const float p2 = gadd<IEEE>(base, t1[0]), p4 = gadd<IEEE>(base, t2[0]), p6 = gadd<IEEE>(p4, t1[2]);
    o[0] = base;
    o[1] = gadd<IEEE>(base, v[0]);
    o[2] = p2;
    o[3] = gadd<IEEE>(p2, v[2]);
    o[4] = p4;
    o[5] = gadd<IEEE>(p4, v[4]);
    o[6] = p6;
    o[7] = gadd<IEEE>(p6, v[6]);
}
// The levels above a step: a pairwise-summation stack.  After `n` blocks have been pushed, level b holds the sum of the
// aligned group of 2^b blocks that ends at n iff bit b of n is set -- the left siblings the down-sweep adds, top level first.
template <bool IEEE, int LEVELS>
__device__ __forceinline__ float stack_prefix(const float (&st)[LEVELS], unsigned n) {
    float acc = 0.0f;
#pragma unroll
    for (int b = LEVELS - 1; b >= 0; --b)
        if ((n >> b) & 1u) acc = gadd<IEEE>(acc, st[b]);
    return acc;
}
template <bool IEEE, int LEVELS>
__device__ __forceinline__ void stack_push(float (&st)[LEVELS], unsigned n, float s) {
    bool done = false;
#pragma unroll
    for (int b = 0; b < LEVELS; ++b) {
        if (!done) {
            if ((n >> b) & 1u) s = gadd<IEEE>(s, st[b]);
            else { st[b] = s; done = true; }
        }
    }
}

// ---- pass 1: exclusive prefix sums along the rows ---------------------------------------------------------------------
struct Planes {       // dense fp32 scratch [plane][y][x], rows padded to a multiple of 8 elements (16-byte aligned)
    float* ptr;
    size_t pitch, plane;   // elements
    __device__ __forceinline__ float* row(int y, int z) const { return ptr + (size_t)z * plane + (size_t)y * pitch; }
};

// ---- pass 3: four-corner lookup + the algebra that follows it ----------------------------------------------------------
// cu_integral_image.cu:130-157.  SASS: I2FP area; MUFU.RCP; FADD.FTZ (C + A); FADD.FTZ -B; FADD.FTZ -D; FMUL.FTZ sum * rcp.
// The sums are exclusive, so the window is [minx, maxx) x [miny, maxy) and area = (maxx - minx) * (maxy - miny).
struct BoxWin { int o_a, o_b, o_c, o_d; float area; };   // element offsets inside a plane
__device__ __forceinline__ BoxWin box_window(int x, int y, int w, int h, int rad, size_t pitch) {
    const int minx = max(0, x - rad), maxx = min(w - 1, x + rad), miny = max(0, y - rad), maxy = min(h - 1, y + rad);
    BoxWin b;
    b.o_a = (int)(miny * pitch) + minx;
    b.o_b = (int)(miny * pitch) + maxx;
    b.o_c = (int)(maxy * pitch) + maxx;
    b.o_d = (int)(maxy * pitch) + minx;
    b.area = (float)((maxx - minx) * (maxy - miny));
    return b;
}
template <bool IEEE>
__device__ __forceinline__ float box_mean(const float* __restrict__ ii, const BoxWin& b, float rcp_area) {
    const float sum = gadd<IEEE>(gadd<IEEE>(gadd<IEEE>(ii[b.o_c], ii[b.o_a]), -ii[b.o_b]), -ii[b.o_d]);
    return IEEE ? __fdiv_rn(sum, b.area) : fmul_ftz(sum, rcp_area);
}

// What the rows a warp scans are made of.  The pair modes read their input once and scan two planes (z and S + z):
//   ROWS_PLANE   one plane of p                                         (BoxFilter)
//   ROWS_I_II    I and I*I of the guide image                           (ComputeMeanVarience, cu_integral_image.h:45-50)
//   ROWS_P_IP    P and I*P of slice z0 + z                              (ComputeCovariance, :59-64)
//   ROWS_A_B     a = cov_Ip / (var_I + eps) and b = mean_p - a mean_I, computed on the fly from the integral images of P and
//                I*P (the four-corner lookups of ComputeCovariance and the first half of GuidedFilter, :59-86): a and b are
//                never written, only their row sums
enum RowMode { ROWS_PLANE = 0, ROWS_I_II = 1, ROWS_P_IP = 2, ROWS_A_B = 3 };
struct RowArgs {
    Vol<float> p;
    Img<float> g, meanI, varI;
    Planes ii;            // ROWS_A_B: integral images of P (plane z) and I*P (plane S + z)
    Planes dst;
    int w, h, nz, z0, S, rad;
    float eps;
};

constexpr int SCAN_ROW_WARPS = 8, SCAN_ROW_LEVELS = 6;    // 256 << 6 = 16384 elements per row at most

template <bool IEEE, int MODE>
__global__ void __launch_bounds__(SCAN_ROW_WARPS * 32, 4)
scan_rows_kernel(RowArgs a) {
    constexpr int NV = MODE == ROWS_PLANE ? 1 : 2;
    const int lane = threadIdx.x & 31;
    const long long rid = (long long)blockIdx.x * SCAN_ROW_WARPS + (threadIdx.x >> 5);
    if (rid >= (long long)a.h * a.nz) return;
    const int y = (int)(rid % a.h), z = (int)(rid / a.h), w = a.w;
    const float* prow = MODE == ROWS_I_II ? a.g.row(y) : (MODE == ROWS_A_B ? nullptr : a.p.row(y, a.z0 + z));
    const float* grow = a.g.row(y);
    const int miny = max(0, y - a.rad), maxy = min(a.h - 1, y + a.rad);
    float st[NV][SCAN_ROW_LEVELS];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int b = 0; b < SCAN_ROW_LEVELS; ++b) st[k][b] = 0.0f;
    const int nseg = (w + 255) >> 8;
    for (int seg = 0; seg < nseg; ++seg) {
        const int x0 = (seg << 8) + lane * 8;
        float v[NV][8];
        if (MODE == ROWS_A_B) {
            // lookups with the lanes on consecutive pixels (coalesced corners), then through shared memory to 8 per lane
            __shared__ __align__(16) float sh[SCAN_ROW_WARPS][2][256];
            float(*mine)[256] = sh[threadIdx.x >> 5];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int x = (seg << 8) + j * 32 + lane;
                float e0 = 0.0f, e1 = 0.0f;
                if (x < w) {
                    const int minx = max(0, x - a.rad), maxx = min(w - 1, x + a.rad);
                    BoxWin b;
                    b.o_a = (int)(miny * a.ii.pitch) + minx; b.o_b = (int)(miny * a.ii.pitch) + maxx;
                    b.o_c = (int)(maxy * a.ii.pitch) + maxx; b.o_d = (int)(maxy * a.ii.pitch) + minx;
                    b.area = (float)((maxx - minx) * (maxy - miny));
                    const float rcp_area = IEEE ? 0.0f : rcp_approx_ftz(b.area);
                    const float* ii0 = a.ii.ptr + (size_t)z * a.ii.plane;
                    const float mP = box_mean<IEEE>(ii0, b, rcp_area), mIP = box_mean<IEEE>(ii0 + (size_t)a.S * a.ii.plane, b, rcp_area);
                    const float mI = a.meanI(x, y);
                    const float cov = ew_multiply_add<IEEE>(mI, mP, mIP, -1.0f, 1.0f, 0.0f);
                    e0 = ew_division<IEEE>(cov, a.varI(x, y), 0.0f, a.eps, 1.0f, 0.0f);          // Eqn. 5
                    e1 = ew_multiply_add<IEEE>(e0, mI, mP, -1.0f, 1.0f, 0.0f);                   // Eqn. 6
                }
                mine[0][j * 32 + lane] = e0;
                mine[1][j * 32 + lane] = e1;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float4 lo = *reinterpret_cast<const float4*>(&mine[k][lane * 8]), hi = *reinterpret_cast<const float4*>(&mine[k][lane * 8 + 4]);
                v[k ? NV - 1 : 0][0] = lo.x; v[k ? NV - 1 : 0][1] = lo.y; v[k ? NV - 1 : 0][2] = lo.z; v[k ? NV - 1 : 0][3] = lo.w;
                v[k ? NV - 1 : 0][4] = hi.x; v[k ? NV - 1 : 0][5] = hi.y; v[k ? NV - 1 : 0][6] = hi.z; v[k ? NV - 1 : 0][7] = hi.w;
            }
            __syncwarp();
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int x = x0 + j;
                float e0 = 0.0f, e1 = 0.0f;
                if (x < w) {
                    e0 = prow[x];
                    // ElementwiseSquare(II, I) / ElementwiseMultiply(IP, I, P) with scalar 1, offset 0
                    if (MODE == ROWS_I_II) e1 = ew_square<IEEE>(e0, 1.0f, 0.0f);
                    if (MODE == ROWS_P_IP) e1 = ew_multiply<IEEE>(grow[x], e0, 1.0f, 0.0f);
                }
                v[0][j] = e0;
                if (NV == 2) v[NV - 1][j] = e1;
            }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float t1[4], t2[2], sib[5];
            float cur = tree8_sum<IEEE>(v[k], t1, t2);
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                sib[b] = __shfl_xor_sync(0xffffffffu, cur, 1 << b);
                cur = gadd<IEEE>(cur, sib[b]);      // a+b == b+a bit for bit: both lanes of a pair hold the same block sum
            }
            float base = stack_prefix<IEEE, SCAN_ROW_LEVELS>(st[k], (unsigned)seg);
#pragma unroll
            for (int b = 4; b >= 0; --b)
                if ((lane >> b) & 1) base = gadd<IEEE>(base, sib[b]);
            float o[8];
            tree8_prefix<IEEE>(base, v[k], t1, t2, o);
            if (x0 < (int)a.dst.pitch) {           // rows are padded to 8: whole vectors, the tail past w is never read
                float* drow = a.dst.row(y, k ? a.S + z : z);
                *reinterpret_cast<float4*>(drow + x0) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(drow + x0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
            stack_push<IEEE, SCAN_ROW_LEVELS>(st[k], (unsigned)seg, cur);
        }
    }
}

// ---- pass 2: exclusive prefix sums down the columns, in place ----------------------------------------------------------
constexpr int SCAN_COL_LEVELS = 13;   // 8 << 13 rows at most

template <bool IEEE>
__global__ void __launch_bounds__(128)
scan_cols_kernel(Planes io, int w, int h) {
    const int x = blockIdx.x * 128 + threadIdx.x, z = blockIdx.y;
    if (x >= w) return;
    float* col = io.row(0, z) + x;
    float st[SCAN_COL_LEVELS];
#pragma unroll
    for (int b = 0; b < SCAN_COL_LEVELS; ++b) st[b] = 0.0f;
    const int nblk = (h + 7) >> 3;
    for (int blk = 0; blk < nblk; ++blk) {
        const int y0 = blk << 3;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = y0 + j < h ? col[(size_t)(y0 + j) * io.pitch] : 0.0f;
        float t1[4], t2[2], o[8];
        const float sum = tree8_sum<IEEE>(v, t1, t2);
        tree8_prefix<IEEE>(stack_prefix<IEEE, SCAN_COL_LEVELS>(st, (unsigned)blk), v, t1, t2, o);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (y0 + j < h) col[(size_t)(y0 + j) * io.pitch] = o[j];
        stack_push<IEEE, SCAN_COL_LEVELS>(st, (unsigned)blk, sum);
    }
}

enum BoxEpi { BEPI_MEAN = 0, BEPI_MEANVAR = 1, BEPI_Q = 2 };
struct BoxEpiArgs {
    Planes ii;            // integral images of this pass
    Vol<float> vol;       // BEPI_MEAN: destination planes from z0; BEPI_Q: the cost volume (slice z0 + z)
    Img<float> guide, meanI, varI;
    int w, h, rad, z0, S;
    float eps;
};

constexpr int BEPI_ZPT = 4;   // slices per thread: one window, 8 (16) independent corner loads per slice in flight

template <int EPI, bool IEEE>
__global__ void __launch_bounds__(128)
box_epilogue_kernel(BoxEpiArgs a, int nz) {
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
    if (x >= a.w) return;
    const BoxWin b = box_window(x, y, a.w, a.h, a.rad, a.ii.pitch);
    const float rcp_area = IEEE ? 0.0f : rcp_approx_ftz(b.area);
    if (EPI == BEPI_MEANVAR) {
        // ComputeMeanVarience (cu_integral_image.h:42-54): var_I = mean_II - mean_I * mean_I
        const float mI = box_mean<IEEE>(a.ii.ptr, b, rcp_area), mII = box_mean<IEEE>(a.ii.ptr + a.ii.plane, b, rcp_area);
        a.meanI(x, y) = mI;
        a.varI(x, y) = ew_multiply_add<IEEE>(mI, mI, mII, -1.0f, 1.0f, 0.0f);
        return;
    }
    const float g = EPI == BEPI_Q ? a.guide(x, y) : 0.0f;
    float q[BEPI_ZPT];
#pragma unroll
    for (int k = 0; k < BEPI_ZPT; ++k) {
        const int z = blockIdx.z * BEPI_ZPT + k;
        if (z < nz) {
            const float* ii0 = a.ii.ptr + (size_t)z * a.ii.plane;
            if (EPI == BEPI_MEAN) q[k] = box_mean<IEEE>(ii0, b, rcp_area);
            else {
                // GuidedFilter (:88-92): q = mean_a * I + mean_b                                  Eqn. 8
                const float ma = box_mean<IEEE>(ii0, b, rcp_area), mb = box_mean<IEEE>(ii0 + (size_t)a.S * a.ii.plane, b, rcp_area);
                q[k] = ew_multiply_add<IEEE>(ma, g, mb, 1.0f, 1.0f, 0.0f);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < BEPI_ZPT; ++k) {
        const int z = blockIdx.z * BEPI_ZPT + k;
        if (z < nz) a.vol(x, y, a.z0 + z) = q[k];
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
// Stream-ordered scratch from a pool of this library's own that keeps what it has been given (the device's default pool
// returns everything to the driver at the next synchronisation: 47 ms per call for the 1.5 GB of a c2-sized volume).
// roo_release_scratch() hands the memory back.
static std::mutex g_pool_mu;
static cudaMemPool_t g_pool[64] = {};
static cudaMemPool_t scratch_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pool[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&g_pool[dev], &props) != cudaSuccess) { cudaGetLastError(); g_pool[dev] = nullptr; return nullptr; }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return g_pool[dev];
}

struct AsyncBuf {
    void* p = nullptr;
    cudaStream_t st;
    explicit AsyncBuf(cudaStream_t s) : st(s) {}
    int alloc(size_t bytes) {
        cudaMemPool_t pool = scratch_pool();
        return (int)(pool ? cudaMallocFromPoolAsync(&p, bytes, pool, st) : cudaMallocAsync(&p, bytes, st));
    }
    ~AsyncBuf() { if (p) cudaFreeAsync(p, st); }
    AsyncBuf(const AsyncBuf&) = delete;
    AsyncBuf& operator=(const AsyncBuf&) = delete;
};

static size_t plane_pitch(size_t w) { return (w + 7) / 8 * 8; }

// rows (pass 1), then columns in place (pass 2): `planes` integral images in a.dst
template <bool IEEE, int MODE>
static int scan_planes(const RowArgs& a, int planes, cudaStream_t st) {
    scan_rows_kernel<IEEE, MODE><<<cdiv((long long)a.h * a.nz, SCAN_ROW_WARPS), SCAN_ROW_WARPS * 32, 0, st>>>(a);
    scan_cols_kernel<IEEE><<<dim3(cdiv(a.w, 128), planes), 128, 0, st>>>(a.dst, a.w, a.h);
    count_launch(2);
    return launch_status();
}

template <int EPI, bool IEEE>
static int box_epilogue(const BoxEpiArgs& a, int nz, cudaStream_t st) {
    box_epilogue_kernel<EPI, IEEE><<<dim3(cdiv(a.w, 128), a.h, cdiv(nz, BEPI_ZPT)), 128, 0, st>>>(a, nz);
    count_launch();
    return launch_status();
}

static roo_volume_t as_volume(const roo_image_t& i) { return roo_volume_t{i.pitch, i.ptr, i.w, i.h, i.pitch * i.h, 1}; }

template <bool IEEE>
static int box_filter_impl(const roo_image_t& out, const roo_image_t& in, int rad, cudaStream_t st) {
    const int w = (int)in.w, h = (int)in.h;
    AsyncBuf buf(st);
    const size_t pitch = plane_pitch(w);
    if (buf.alloc(pitch * h * sizeof(float))) { cudaGetLastError(); return ROO_ERR_OUT_OF_MEMORY; }
    RowArgs r{};
    r.p = Vol<float>(as_volume(in)); r.g = Img<float>(in); r.dst = Planes{(float*)buf.p, pitch, pitch * h};
    r.w = w; r.h = h; r.nz = 1; r.z0 = 0; r.S = 1; r.rad = rad;
    int rc = scan_planes<IEEE, ROWS_PLANE>(r, 1, st);
    if (rc) return rc;
    BoxEpiArgs a{};
    a.ii = r.dst; a.vol = Vol<float>(as_volume(out)); a.w = w; a.h = h; a.rad = rad; a.z0 = 0; a.S = 1;
    return box_epilogue<BEPI_MEAN, IEEE>(a, 1, st);
}

// Scratch budget of the volume filter (MiB; ROO_TUNE_GUIDED_SCRATCH_MIB): 4 planes per slice in flight.
std::atomic<int> g_guided_scratch_mib{2048};

template <bool IEEE>
static int guided_filter_impl(const roo_volume_t& vol, const roo_image_t& guide, int rad, float eps, int nd, cudaStream_t st) {
    const int w = (int)vol.w, h = (int)vol.h;
    const size_t pitch = plane_pitch(w), plane = pitch * h;
    int S = (int)(((size_t)g_guided_scratch_mib.load() << 20) / (4 * plane * sizeof(float)));
    S = S < 1 ? 1 : (S > nd ? nd : S);
    AsyncBuf buf(st);
    // [2S planes: integral images of P, I*P | 2S planes: integral images of a, b | meanI | varI]
    if (buf.alloc((4 * (size_t)S + 2) * plane * sizeof(float))) { cudaGetLastError(); return ROO_ERR_OUT_OF_MEMORY; }
    float* base = (float*)buf.p;
    const Planes A{base, pitch, plane}, B{base + 2 * (size_t)S * plane, pitch, plane};
    const roo_image_t meanI{pitch * sizeof(float), base + 4 * (size_t)S * plane, (size_t)w, (size_t)h};
    const roo_image_t varI{pitch * sizeof(float), base + (4 * (size_t)S + 1) * plane, (size_t)w, (size_t)h};

    RowArgs r{};
    r.p = Vol<float>(vol); r.g = Img<float>(guide); r.meanI = Img<float>(meanI); r.varI = Img<float>(varI);
    r.w = w; r.h = h; r.rad = rad; r.eps = eps;
    BoxEpiArgs a{};
    a.vol = Vol<float>(vol); a.guide = Img<float>(guide); a.meanI = Img<float>(meanI); a.varI = Img<float>(varI);
    a.w = w; a.h = h; a.rad = rad; a.eps = eps;
    int rc;
    // guide statistics: box(I), box(I*I) -> mean_I, var_I
    r.dst = A; r.nz = 1; r.z0 = 0; r.S = 1;
    if ((rc = scan_planes<IEEE, ROWS_I_II>(r, 2, st))) return rc;
    a.ii = A; a.S = 1;
    if ((rc = box_epilogue<BEPI_MEANVAR, IEEE>(a, 1, st))) return rc;
    for (int z0 = 0; z0 < nd; z0 += S) {
        const int s = nd - z0 < S ? nd - z0 : S;
        r.dst = A; r.nz = s; r.z0 = z0; r.S = s;
        if ((rc = scan_planes<IEEE, ROWS_P_IP>(r, 2 * s, st))) return rc;
        r.ii = A; r.dst = B;
        if ((rc = scan_planes<IEEE, ROWS_A_B>(r, 2 * s, st))) return rc;
        a.ii = B; a.z0 = z0; a.S = s;
        if ((rc = box_epilogue<BEPI_Q, IEEE>(a, s, st))) return rc;
    }
    return ROO_OK;
}

}  // namespace roo_b200

using namespace roo_b200;

static bool same_size(const roo_image_t* a, const roo_image_t* b) { return a->w == b->w && a->h == b->h; }
static bool ranges_overlap(const void* a, size_t na, const void* b, size_t nb) {
    const char *pa = (const char*)a, *pb = (const char*)b;
    return pa < pb + nb && pb < pa + na;
}

extern "C" int roo_elementwise_multiply(const roo_image_t* c, const roo_image_t* a, const roo_image_t* b, float scalar, float offset,
                                        void* stream) {
    if (!valid_image(c, 4) || !valid_image(a, 4) || !valid_image(b, 4) || !same_size(c, a) || !same_size(c, b)) return ROO_ERR_INVALID_ARGUMENT;
    return launch_elementwise<EW_MULTIPLY>(*c, *a, *b, *a, scalar, offset, 0.0f, 0.0f, as_stream(stream));
}

extern "C" int roo_elementwise_division(const roo_image_t* c, const roo_image_t* a, const roo_image_t* b, float sa, float sb,
                                        float scalar, float offset, void* stream) {
    if (!valid_image(c, 4) || !valid_image(a, 4) || !valid_image(b, 4) || !same_size(c, a) || !same_size(c, b)) return ROO_ERR_INVALID_ARGUMENT;
    return launch_elementwise<EW_DIVISION>(*c, *a, *b, *a, sa, sb, scalar, offset, as_stream(stream));
}

extern "C" int roo_elementwise_square(const roo_image_t* b, const roo_image_t* a, float scalar, float offset, void* stream) {
    if (!valid_image(b, 4) || !valid_image(a, 4) || !same_size(b, a)) return ROO_ERR_INVALID_ARGUMENT;
    return launch_elementwise<EW_SQUARE>(*b, *a, *a, *a, scalar, offset, 0.0f, 0.0f, as_stream(stream));
}

extern "C" int roo_elementwise_multiply_add(const roo_image_t* d, const roo_image_t* a, const roo_image_t* b, const roo_image_t* c,
                                            float sab, float sc, float offset, void* stream) {
    if (!valid_image(d, 4) || !valid_image(a, 4) || !valid_image(b, 4) || !valid_image(c, 4)) return ROO_ERR_INVALID_ARGUMENT;
    if (!same_size(d, a) || !same_size(d, b) || !same_size(d, c)) return ROO_ERR_INVALID_ARGUMENT;
    return launch_elementwise<EW_MULTIPLY_ADD>(*d, *a, *b, *c, sab, sc, offset, 0.0f, as_stream(stream));
}

extern "C" int roo_release_scratch(void) {
    int dev = 0;
    ROO_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (dev >= 0 && dev < 64 && g_pool[dev]) ROO_CUDA_TRY(cudaMemPoolTrimTo(g_pool[dev], 0));
    return ROO_OK;
}

extern "C" int roo_box_filter(const roo_image_t* out, const roo_image_t* in, int rad, void* stream) {
    if (!valid_image(out, 4) || !valid_image(in, 4) || !same_size(out, in) || rad < 0) return ROO_ERR_INVALID_ARGUMENT;
    if (in->w > ((size_t)256 << SCAN_ROW_LEVELS) || in->h > ((size_t)8 << SCAN_COL_LEVELS)) return ROO_ERR_UNSUPPORTED;
    if ((in->w + 8) * in->h >= ((size_t)1 << 31)) return ROO_ERR_UNSUPPORTED;   // plane offsets are 32-bit
    return g_ieee_div.load() ? box_filter_impl<true>(*out, *in, rad, as_stream(stream))
                             : box_filter_impl<false>(*out, *in, rad, as_stream(stream));
}

extern "C" int roo_guided_filter_volume(const roo_volume_t* vol, const roo_image_t* guide, int rad, float eps, int maxDisp,
                                        void* stream) {
    if (!valid_volume(vol, 4) || !valid_image(guide, 4) || guide->w != vol->w || guide->h != vol->h || rad < 0) return ROO_ERR_INVALID_ARGUMENT;
    if (maxDisp <= 0) return ROO_OK;
    if ((size_t)maxDisp > vol->d) return ROO_ERR_INVALID_ARGUMENT;
    if (ranges_overlap(vol->ptr, vol->img_pitch * vol->d, guide->ptr, guide->pitch * guide->h)) return ROO_ERR_INVALID_ARGUMENT;
    if (vol->w > ((size_t)256 << SCAN_ROW_LEVELS) || vol->h > ((size_t)8 << SCAN_COL_LEVELS)) return ROO_ERR_UNSUPPORTED;
    if ((vol->w + 8) * vol->h >= ((size_t)1 << 31)) return ROO_ERR_UNSUPPORTED;   // plane offsets are 32-bit
    return g_ieee_div.load() ? guided_filter_impl<true>(*vol, *guide, rad, eps, maxDisp, as_stream(stream))
                             : guided_filter_impl<false>(*vol, *guide, rad, eps, maxDisp, as_stream(stream));
}
