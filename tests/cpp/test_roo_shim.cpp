// Host-side C++ test of the drop-in shim: the reference's call sequence (applications/stereo2/main.cpp:380-454)
// written against roo.hpp, checked for self-consistency with the fused engine.  Built and run by
// tests/test_gpu_cpp_shim.py on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
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

    // front end / back end operators, as the applications call them (stereo2/main.cpp:360-376, stereo/main.cpp:476)
    {
        auto imgf = alloc_image<float>(w, h), half = alloc_image<float>(w / 2, h / 2), depth = alloc_image<float>(w, h);
        auto half8 = alloc_image<unsigned char>(w / 2, h / 2);
        auto vbo = alloc_image<float4>(w, h);
        roo::ElementwiseScaleBias<float, unsigned char, float>(imgf, imgL, 1.0f / 255.0f);
        roo::BoxHalf<float, float, float>(half, imgf);
        roo::BoxHalf<unsigned char, unsigned int, unsigned char>(half8, imgL);
        roo::Disp2Depth(disp, depth, 500.0f, 0.1f);
        roo::DisparityImageToVbo(vbo, disp, 0.1f, 500.0f, 500.0f, w / 2.0f, h / 2.0f);
        CK(cudaDeviceSynchronize());
        std::vector<float> f(w * h), hf((w / 2) * (h / 2)), z(w * h);
        std::vector<float4> P(w * h);
        std::vector<unsigned char> h8((w / 2) * (h / 2));
        CK(cudaMemcpy2D(f.data(), w * 4, imgf.ptr, imgf.pitch, w * 4, h, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy2D(hf.data(), (w / 2) * 4, half.ptr, half.pitch, (w / 2) * 4, h / 2, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy2D(h8.data(), w / 2, half8.ptr, half8.pitch, w / 2, h / 2, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy2D(z.data(), w * 4, depth.ptr, depth.pitch, w * 4, h, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy2D(P.data(), w * 16, vbo.ptr, vbo.pitch, w * 16, h, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int y = 0; y < h / 2; ++y)
            for (int x = 0; x < w / 2; ++x) {
                const int a = L[2 * y * w + 2 * x], b = L[2 * y * w + 2 * x + 1], c = L[(2 * y + 1) * w + 2 * x], d = L[(2 * y + 1) * w + 2 * x + 1];
                if (h8[y * (w / 2) + x] != (unsigned char)((a + b + c + d) / 4)) ++bad;
                if (std::fabs(hf[y * (w / 2) + x] - (a + b + c + d) / (4.0f * 255.0f)) > 1e-6f) ++bad;
            }
        for (int i = 0; i < w * h; ++i) {
            if (f[i] != L[i] * (1.0f / 255.0f)) ++bad;
            const float d = out[i];
            if (std::isnan(d)) { if (!std::isnan(z[i]) || !std::isnan(P[i].z)) ++bad; continue; }
            if (d > 0 && (std::fabs(z[i] - 50.0f / d) > 1e-4f * z[i] || P[i].z != z[i] || P[i].w != 1.0f)) ++bad;
        }
        // on-disk outputs: header + tightly packed payload (extra/SavePPM.h:20-39, stereo/main.cpp:400-410)
        if (!roo::SavePDM("/tmp/roo_shim_test.pdm", z.data(), w, h) || !roo::SavePXM<unsigned char>("/tmp/roo_shim_test.pgm", L.data(), w, h, w)) ++bad;
        {
            FILE* fp = std::fopen("/tmp/roo_shim_test.pdm", "rb");
            char hdr[64] = {0};
            const size_t n = fp ? std::fread(hdr, 1, 24, fp) : 0;
            if (fp) std::fclose(fp);
            char want[64];
            const int wl = std::snprintf(want, sizeof(want), "P7\n%d %d\n4294967295\n", w, h);
            if (n < (size_t)wl || std::memcmp(hdr, want, wl) != 0) ++bad;
        }
        std::printf("shim: front/back end operators, %d mismatches\n", bad);
        if (bad) return 1;
    }

    // the alternative aggregation of the applications (stereo2/main.cpp:392-405) and the direct matcher: the one-call guided
    // filter equals the reference's per-slice sequence spelled with the operators, bit for bit; DenseStereo finds the
    // constant disparity of this pair
    {
        auto imgf = alloc_image<float>(w, h), meanI = alloc_image<float>(w, h), varI = alloc_image<float>(w, h);
        roo::Image<float> t[5];
        for (auto& i : t) i = alloc_image<float>(w, h);
        roo::Image<unsigned char> scratch = alloc_image<unsigned char>(16, 1);       // part of the signature, unused
        roo::ElementwiseScaleBias<float, unsigned char, float>(imgf, imgL, 1.0f / 255.0f);
        roo::CensusStereoVolume<float, unsigned long>(volC, cenL, cenR, D, -1);
        CK(cudaMemcpy(volH.ptr, volC.ptr, volC.img_pitch * D, cudaMemcpyDeviceToDevice));
        const int rad = 4;
        const float eps = 1e-3f;
        roo::GuidedFilterVolume(volH, imgf, rad, eps, D);
        roo::ComputeMeanVarience<float, float, float>(varI, t[0], meanI, imgf, scratch, rad);
        for (int d = 0; d < D; ++d) {
            roo::Image<float> P = volC.ImageXY(d);
            roo::ComputeCovariance(t[0], t[2], t[1], P, meanI, imgf, scratch, rad);
            roo::GuidedFilter(P, t[0], varI, t[1], meanI, imgf, scratch, t[2], t[3], t[4], rad, eps);
        }
        CK(cudaDeviceSynchronize());
        std::vector<float> a((size_t)w * h * D), b((size_t)w * h * D);
        CK(cudaMemcpy2D(a.data(), w * 4, volH.ptr, volH.pitch, w * 4, (size_t)h * D, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy2D(b.data(), w * 4, volC.ptr, volC.pitch, w * 4, (size_t)h * D, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (size_t i = 0; i < a.size(); ++i)
            if (std::memcmp(&a[i], &b[i], 4) != 0 && !(a[i] != a[i] && b[i] != b[i])) ++bad;
        auto dd = alloc_image<unsigned char>(w, h);
        roo::DenseStereo<unsigned char, unsigned char>(dd, imgL, imgR, (unsigned char)D, 0.0f, 2);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned char> dh(w * h);
        CK(cudaMemcpy2D(dh.data(), w, dd.ptr, dd.pitch, w, h, cudaMemcpyDeviceToHost));
        int hit = 0, all = 0;
        for (int y = 8; y < h - 8; ++y)
            for (int x = D + 8; x < w - 24; ++x) { ++all; hit += dh[y * w + x] == 9; }
        std::printf("shim: guided filter volume vs operator sequence %d mismatches; DenseStereo %d / %d at the true disparity\n", bad, hit, all);
        if (bad || hit < all * 0.95) return 1;
    }

    // invalid arguments raise under ROO_B200_THROW
    bool threw = false;
    try { roo::CensusStereoVolume<float, unsigned long>(volC, cenL, cenR, D, 0.5f); } catch (const std::exception&) { threw = true; }
    if (!threw) { std::printf("expected an exception for sd = 0.5\n"); return 1; }
    std::printf("OK\n");
    return 0;
}
