// How long does __nanosleep(N) really take on this GPU?  (one warp, clock64 around the call; development aid)
#include <cstdio>
#include <cuda_runtime.h>
template <int N> __device__ long long probe() {
    long long best = 1 << 30, sum = 0;
    for (int i = 0; i < 64; ++i) {
        const long long t0 = clock64();
        __nanosleep(N);
        const long long dt = clock64() - t0;
        best = dt < best ? dt : best;
        sum += dt;
    }
    return (best << 32) | (sum / 64);
}
__global__ void k(long long* out) {
    if (threadIdx.x != 0) return;
    out[0] = probe<0>(); out[1] = probe<30>(); out[2] = probe<60>(); out[3] = probe<100>(); out[4] = probe<200>();
    out[5] = probe<400>(); out[6] = probe<800>(); out[7] = probe<1600>();
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    k<<<1, 32>>>(d);
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    const int n[8] = {0, 30, 60, 100, 200, 400, 800, 1600};
    for (int i = 0; i < 8; ++i) printf("nanosleep(%4d): min %lld cycles, mean %lld cycles\n", n[i], h[i] >> 32, h[i] & 0xffffffff);
    return 0;
}
