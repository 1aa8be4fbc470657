// MedianFilterRejectNegative{5x5,7x7,9x9} (src/cu_median.cu:160-350) -- the step between winner-takes-all and the
// left-right check in both applications (stereo2/main.cpp:438-444).  SURVEY.md section 8f, N1.
//
// Semantics (out of place): window = clamp-to-edge neighbourhood, bad = number of non-finite samples, output NaN
// unless bad < maxbad && bad < size^2, else the median of the valid samples: sorted valid samples, element
// (size^2 + bad)/2 - bad.  For windows without invalid samples this is bit-identical to the reference (its exchange
// network then returns the exact median); with invalid samples the reference's result depends on its comparator
// order (fminf/fmaxf overwrite NaNs with copies of their partners) and is NOT reproduced -- see oracle header and
// tests/golden/median.npz.  Calling it in place (as the applications do) races in the reference; here in == out is
// refused.
//
// Kernel: a 32x8 tile (+ apron) staged in shared memory with clamp-to-edge; every thread keeps its size^2 samples in
// registers as order-preserving integer keys (invalid samples = the largest key) and sorts them with Batcher's
// odd-even merge sort, generated at compile time for exactly size^2 inputs (140 / 394 / 864 compare-exchanges of two
// integer min/max instructions each) -- a network of this file's own making, exact for any input; the wanted rank
// (size^2 + bad)/2 - bad is then picked with a select chain.
#include <utility>

#include "common.cuh"
#include "kernels.cuh"

namespace roo_b200 {

constexpr int MED_TX = 32, MED_TY = 8;

__device__ __forceinline__ unsigned float_key(float f) {   // monotonic: a < b  <=>  key(a) < key(b)
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Batcher's odd-even merge sort for exactly K inputs, built at compile time: the comparators that would touch the
// virtual +inf padding up to the next power of two are dropped (140 / 394 / 864 compare-exchanges for 25 / 49 / 81).
template <int K> struct SortNet { int a[1024]; int b[1024]; int n; };
template <int K> __host__ __device__ constexpr SortNet<K> make_sort_net() {
    SortNet<K> r{};
    int n = 0;
    for (int p = 1; p < K; p <<= 1)
        for (int k = p; k >= 1; k >>= 1)
            for (int j = k % p; j + k < K; j += 2 * k)
                for (int i = 0; i < k; ++i)
                    if (i + j + k < K && (i + j) / (2 * p) == (i + j + k) / (2 * p)) { r.a[n] = i + j; r.b[n] = i + j + k; ++n; }
    r.n = n;
    return r;
}
template <int K> struct SortNetOf { static constexpr SortNet<K> net = make_sort_net<K>(); };

template <int A, int B, int K>
__device__ __forceinline__ void compare_exchange(unsigned (&key)[K]) {
    const unsigned x = key[A], y = key[B];
    key[A] = min(x, y);
    key[B] = max(x, y);
}
// every register index is a template constant, so the keys never leave registers
template <int K, int... I>
__device__ __forceinline__ void sort_keys(unsigned (&key)[K], std::integer_sequence<int, I...>) {
    (compare_exchange<SortNetOf<K>::net.a[I], SortNetOf<K>::net.b[I], K>(key), ...);
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
    unsigned key[K];
    int bad = 0;
#pragma unroll
    for (int dy = 0; dy < SIZE; ++dy)
#pragma unroll
        for (int dx = 0; dx < SIZE; ++dx) {
            const float v = tile[threadIdx.y + dy][threadIdx.x + dx];
            const bool ok = isfinite(v);            // InvalidValue<float>::IsValid (InvalidValue.h:18-47)
            bad += ok ? 0 : 1;
            key[dy * SIZE + dx] = ok ? float_key(v) : 0xffffffffu;   // invalid samples sort last
        }
    float r = __int_as_float(0x7fffffff);
    if (bad < maxbad && bad < K) {
        sort_keys<K>(key, std::make_integer_sequence<int, SortNetOf<K>::net.n>{});
        const int rank = (K + bad) / 2 - bad;        // valid samples sorted ascending, invalid ones last
        unsigned sel = key[K / 2];
#pragma unroll
        for (int i = 0; i < K / 2; ++i) sel = (i == rank) ? key[i] : sel;
        r = key_float(sel);
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
