// Host-side C++ test of the drop-in shim: the reference's call sequence (applications/stereo2/main.cpp:380-454)
// written against roo.hpp, checked for self-consistency with the fused engine.  Built and run by
// tests/test_gpu_cpp_shim.py on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define ROO_B200_THROW
#include "kangaroo_b200/roo.hpp"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

template <typename T> roo::Image<T> alloc_image(size_t w, size_t h) {
    T* p; size_t pitch; cudaMallocPitch((void**)&p, &pitch, w * sizeof(T), h); return roo::Image<T>(p, w, h, pitch);
}
template <typename T> roo::Volume<T> alloc_volume(size_t w, size_t h, size_t d) {
    T* p; size_t pitch; cudaMallocPitch((void**)&p, &pitch, w * sizeof(T), h * d); return roo::Volume<T>(p, w, h, d, pitch);
}

int main() {
    const int w = 200, h = 96, D = 64;
    std::vector<unsigned char> L(w * h), R(w * h);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 24; };
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) L[y * w + x] = (unsigned char)((rnd() + 3 * (x / 7) + 5 * (y / 5)) & 0xff);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) R[y * w + x] = L[y * w + std::min(w - 1, x + 9)];   // constant disparity 9

    auto imgL = alloc_image<unsigned char>(w, h), imgR = alloc_image<unsigned char>(w, h);
    CK(cudaMemcpy2D(imgL.ptr, imgL.pitch, L.data(), w, w, h, cudaMemcpyHostToDevice));
    CK(cudaMemcpy2D(imgR.ptr, imgR.pitch, R.data(), w, w, h, cudaMemcpyHostToDevice));
    auto cenL = alloc_image<unsigned long>(w, h), cenR = alloc_image<unsigned long>(w, h);
    auto volC = alloc_volume<float>(w, h, D), volH = alloc_volume<float>(w, h, D);
    auto disp = alloc_image<float>(w, h), dispR = alloc_image<float>(w, h);

    // the reference's sequence
    roo::Census(cenL, imgL);
    roo::Census(cenR, imgR);
    roo::CensusStereoVolume<float, unsigned long>(volC, cenL, cenR, D, -1);
    roo::SemiGlobalMatching<float, float, unsigned char>(volH, volC, imgL, D, 0.01f * 255, 0.02f * 255, true, true, true);
    roo::CostVolMinimum<float, float>(disp, volH, D);
    roo::CensusStereoVolume<float, unsigned long>(volC, cenR, cenL, D, +1);
    roo::CostVolMinimum<float, float>(dispR, volC, D);
    roo::LeftRightCheck(dispR, disp, +1.0f, 1.0f);
    roo::LeftRightCheck(disp, dispR, -1.0f, 1.0f);
    CK(cudaDeviceSynchronize());

    std::vector<float> out(w * h);
    CK(cudaMemcpy2D(out.data(), w * 4, disp.ptr, disp.pitch, w * 4, h, cudaMemcpyDeviceToHost));
    int good = 0, total = 0;
    for (int y = 8; y < h - 8; ++y)
        for (int x = D; x < w - 16; ++x) { ++total; if (std::fabs(out[y * w + x] - 9.0f) < 0.5f) ++good; }
    std::printf("shim: %d / %d interior pixels at the true disparity\n", good, total);
    if (good < total * 0.95) return 1;

    // invalid arguments raise under ROO_B200_THROW
    bool threw = false;
    try { roo::CensusStereoVolume<float, unsigned long>(volC, cenL, cenR, D, 0.5f); } catch (const std::exception&) { threw = true; }
    if (!threw) { std::printf("expected an exception for sd = 0.5\n"); return 1; }
    std::printf("OK\n");
    return 0;
}
