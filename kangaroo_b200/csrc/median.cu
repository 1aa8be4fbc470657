// MedianFilterRejectNegative{5x5,7x7,9x9} (src/cu_median.cu:160-350) -- the step between winner-takes-all and the
// left-right check in both applications (stereo2/main.cpp:438-444).  SURVEY.md section 8f, N1.
//
// Semantics = the reference's, for every input: window = clamp-to-edge neighbourhood gathered column-major
// (v[(dX + r) * size + (dY + r)]), bad = number of non-finite samples, output NaN unless bad < maxbad && bad < size^2,
// else v[(size^2 + bad)/2] after the reference's exchange network s2(a,b): a = min(a,b), b = max(a_old,b) has run over the
// raw samples.  min/max ignore a NaN operand, so an invalid sample is overwritten by a copy of its partner: for windows
// without invalid samples the result is the exact median, with them it is the sample the partially sorted array holds at
// that index -- which depends on the comparator sequence.  That sequence is not transcribed but GENERATED at compile
// time: it is the bitonic sorting network for n = size^2 inputs (for every block size k = 2, 4, ..: a flip stage pairing
// i with i ^ (k-1), then half-cleaners pairing i with i + j for j = k/4 .. 1), comparators that would reach past input
// n-1 dropped, and every comparator that cannot influence outputs n/2 .. n-1 removed (the median index never lies below
// n/2) -- 155 / 439 / 968 compare-exchanges for 25 / 49 / 81, the reference's sequence comparator for comparator
// (tests/test_oracle_golden.py checks that against the reference file when it is present; tests/golden/median.npz pins the
// outputs, windows with invalid samples included).  Calling the filter in place (as the applications do) races in the
// reference; here an overlapping call goes through a temporary.
//
// Kernel: a 32x8 tile (+ apron) staged in shared memory with clamp-to-edge; every thread keeps its size^2 samples in
// registers -- every register index of the network is a template constant -- and exchanges with min.ftz / max.ftz (the
// reference build's FMNMX.FTZ); the wanted index (size^2 + bad)/2 is then picked with a select chain.
#include <utility>

#include "common.cuh"
#include "kernels.cuh"

namespace roo_b200 {

constexpr int MED_TX = 32, MED_TY = 8;

// The reference's exchange network for exactly K inputs (see the header), built at compile time.
template <int K> struct SortNet { unsigned char a[1024]; unsigned char b[1024]; int n; };
template <int K> __host__ __device__ constexpr SortNet<K> make_sort_net() {
    unsigned char ca[2048] = {}, cb[2048] = {};
    bool keep[2048] = {};
    int stage_end[64] = {};
    int N = 1, total = 0, nstage = 0;
    while (N < K) N <<= 1;
    for (int k = 2; k <= N; k <<= 1) {
        for (int i = 0; i < N; ++i) {
            const int l = i ^ (k - 1);
            if (l > i && l < K) { ca[total] = (unsigned char)i; cb[total] = (unsigned char)l; ++total; }
        }
        stage_end[nstage++] = total;
        for (int j = k >> 2; j >= 1; j >>= 1) {
            for (int i = 0; i < N; ++i)
                if (!(i & j) && i + j < K) { ca[total] = (unsigned char)i; cb[total] = (unsigned char)(i + j); ++total; }
            stage_end[nstage++] = total;
        }
    }
    bool need[128] = {};
    for (int i = K / 2; i < K; ++i) need[i] = true;
    for (int s = nstage - 1; s >= 0; --s) {             // dead-comparator elimination, last stage first
        const int lo = s ? stage_end[s - 1] : 0, hi = stage_end[s];
        for (int c = lo; c < hi; ++c) keep[c] = need[ca[c]] || need[cb[c]];
        for (int c = lo; c < hi; ++c)
            if (keep[c]) { need[ca[c]] = true; need[cb[c]] = true; }
    }
    SortNet<K> r{};
    int n = 0;
    for (int c = 0; c < total; ++c)
        if (keep[c]) { r.a[n] = ca[c]; r.b[n] = cb[c]; ++n; }
    r.n = n;
    return r;
}
template <int K> struct SortNetOf { static constexpr SortNet<K> net = make_sort_net<K>(); };
static_assert(SortNetOf<25>::net.n == 155 && SortNetOf<49>::net.n == 439 && SortNetOf<81>::net.n == 968, "comparator counts of cu_median.cu");

// s2(a, b) of cu_median.cu:15-16 on FMNMX.FTZ: a NaN operand is ignored, -0 < +0
template <int A, int B, int K>
__device__ __forceinline__ void compare_exchange(float (&v)[K]) {
    const float x = v[A], y = v[B];
    asm("min.ftz.f32 %0, %1, %2;" : "=f"(v[A]) : "f"(x), "f"(y));
    asm("max.ftz.f32 %0, %1, %2;" : "=f"(v[B]) : "f"(x), "f"(y));
}
// every register index is a template constant, so the samples never leave registers
template <int K, int... I>
__device__ __forceinline__ void run_network(float (&v)[K], std::integer_sequence<int, I...>) {
    (compare_exchange<SortNetOf<K>::net.a[I], SortNetOf<K>::net.b[I], K>(v), ...);
}

template <int SIZE>
__global__ void __launch_bounds__(MED_TX* MED_TY)
median_reject_kernel(Img<float> out, Img<float> in, int maxbad, size_t out_batch, size_t in_batch) {
    constexpr int R = SIZE / 2, K = SIZE * SIZE, TW = MED_TX + 2 * R, TH = MED_TY + 2 * R;
    __shared__ float tile[TH][TW + 1];
    out.ptr += (size_t)blockIdx.z * out_batch;   // image blockIdx.z of a batch (clamp-to-edge stays per image)
    in.ptr += (size_t)blockIdx.z * in_batch;
    const int x0 = blockIdx.x * MED_TX, y0 = blockIdx.y * MED_TY;
    for (int i = threadIdx.y * MED_TX + threadIdx.x; i < TW * TH; i += MED_TX * MED_TY) {
        const int ty = i / TW, tx = i - ty * TW;
        tile[ty][tx] = in(clampi(x0 + tx - R, 0, in.w - 1), clampi(y0 + ty - R, 0, in.h - 1));
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= out.w || y >= out.h) return;
    float v[K];
    int bad = 0;
#pragma unroll
    for (int dx = 0; dx < SIZE; ++dx)
#pragma unroll
        for (int dy = 0; dy < SIZE; ++dy) {
            const float s = tile[threadIdx.y + dy][threadIdx.x + dx];
            bad += isfinite(s) ? 0 : 1;             // InvalidValue<float>::IsValid (InvalidValue.h:18-47)
            v[dx * SIZE + dy] = s;                  // the reference's column-major gather order
        }
    float r = __int_as_float(0x7fffffff);
    if (bad < maxbad && bad < K) {
        run_network<K>(v, std::make_integer_sequence<int, SortNetOf<K>::net.n>{});
        const int idx = (K + bad) / 2;
        r = v[K / 2];
#pragma unroll
        for (int i = K / 2 + 1; i < K; ++i) r = (i == idx) ? v[i] : r;
    }
    out(x, y) = r;
}

int launch_median(float* out, size_t out_pitch, size_t out_batch, const float* in, size_t in_pitch, size_t in_batch, int w,
                  int h, int batch, int size, int maxbad, cudaStream_t st) {
    Img<float> o, i;
    o.ptr = (char*)out; o.pitch = out_pitch; o.w = w; o.h = h;
    i.ptr = (char*)in; i.pitch = in_pitch; i.w = w; i.h = h;
    const dim3 grid(cdiv(w, MED_TX), cdiv(h, MED_TY), batch), block(MED_TX, MED_TY);
    if (size == 5) median_reject_kernel<5><<<grid, block, 0, st>>>(o, i, maxbad, out_batch, in_batch);
    else if (size == 7) median_reject_kernel<7><<<grid, block, 0, st>>>(o, i, maxbad, out_batch, in_batch);
    else if (size == 9) median_reject_kernel<9><<<grid, block, 0, st>>>(o, i, maxbad, out_batch, in_batch);
    else return ROO_ERR_UNSUPPORTED;
    count_launch();
    return launch_status();
}

}  // namespace roo_b200

using namespace roo_b200;

extern "C" int roo_median_filter_reject_negative(const roo_image_t* out, const roo_image_t* in, int size, int maxbad,
                                                 void* stream) {
    if (size != 5 && size != 7 && size != 9) return ROO_ERR_UNSUPPORTED;
    if (!valid_image(out, 4) || !valid_image(in, 4) || in->w != out->w || in->h != out->h) return ROO_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    // Both reference applications call the filter IN PLACE (stereo2/main.cpp:440-442), which races in the reference
    // (a block reads neighbours another block has already overwritten).  Here an overlapping call filters into a
    // stream-ordered temporary and copies back: the result is what the out-of-place call gives, never a silent no-op.
    const char *ob = (const char*)out->ptr, *ib = (const char*)in->ptr;
    if (ob < ib + in->pitch * in->h && ib < ob + out->pitch * out->h) {
        const size_t tp = out->w * sizeof(float);
        float* tmp = nullptr;
        ROO_CUDA_TRY(cudaMallocAsync((void**)&tmp, tp * out->h, st));
        int rc = launch_median(tmp, tp, 0, (const float*)in->ptr, in->pitch, 0, (int)out->w, (int)out->h, 1, size, maxbad, st);
        if (rc == 0) rc = (int)cudaMemcpy2DAsync(out->ptr, out->pitch, tmp, tp, tp, out->h, cudaMemcpyDeviceToDevice, st);
        const cudaError_t fe = cudaFreeAsync(tmp, st);
        return rc != 0 ? rc : (int)fe;
    }
    return launch_median((float*)out->ptr, out->pitch, 0, (const float*)in->ptr, in->pitch, 0, (int)out->w, (int)out->h, 1,
                         size, maxbad, st);
}
