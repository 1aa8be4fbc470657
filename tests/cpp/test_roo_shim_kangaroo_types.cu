// Compiles the roo:: shim against the REFERENCE's own data model (kangaroo/Image.h, Volume.h, CostVolElem.h with
// Manage-owning images and volumes) and spells out the per-frame call sequence of applications/stereo2/main.cpp:375-458
// plus every overload / explicit instantiation of the hot-path operators.  Built by tests/test_capi_symbols.py when
// /root/reference is present (compile + link only; tests/test_gpu_cpp_shim.py runs it on a GPU box that has it).
#define ROO_B200_USE_KANGAROO_TYPES
#include <kangaroo_b200/roo.hpp>

#include <cstdio>

typedef ulong4 census_t;   // stereo2/main.cpp:169

int main() {
    const int w = 320, h = 96, maxdisp = 64;
    // owners exactly as the application declares them (main.cpp:160-200)
    roo::Image<unsigned char, roo::TargetDevice, roo::Manage> upload(w, h), disp_c(w, h);
    roo::Image<float, roo::TargetDevice, roo::Manage> img[] = {{(size_t)w, (size_t)h}, {(size_t)w, (size_t)h}};
    roo::Image<census_t, roo::TargetDevice, roo::Manage> census[] = {{(size_t)w, (size_t)h}, {(size_t)w, (size_t)h}};
    roo::Image<unsigned long, roo::TargetDevice, roo::Manage> census1(w, h), census1r(w, h);
    roo::Image<ulong2, roo::TargetDevice, roo::Manage> census2(w, h), census2r(w, h);
    roo::Volume<float, roo::TargetDevice, roo::Manage> vol[] = {{(size_t)w, (size_t)h, (size_t)maxdisp}, {(size_t)w, (size_t)h, (size_t)maxdisp}, {(size_t)w, (size_t)h, (size_t)maxdisp}};
    roo::Volume<unsigned short, roo::TargetDevice, roo::Manage> volu(w, h, maxdisp);
    roo::Volume<roo::CostVolElem, roo::TargetDevice, roo::Manage> vole(w, h, maxdisp);
    roo::Image<float, roo::TargetDevice, roo::Manage> disp[] = {{(size_t)w, (size_t)h}, {(size_t)w, (size_t)h}}, dispf(w, h), depth(w, h);
    roo::Image<char, roo::TargetDevice, roo::Manage> dispi8(w, h), dispi8r(w, h);
    roo::Image<float4, roo::TargetDevice, roo::Manage> vbo(w, h);
    upload.Memset(0); vole.Memset(0);

    // ---- the frame loop of stereo2/main.cpp
    roo::ElementwiseScaleBias<float, unsigned char, float>(img[0], upload, 1.0f / 255.0f);      // :376
    roo::ElementwiseScaleBias<float, unsigned char, float>(img[1], upload, 1.0f / 255.0f);
    roo::Census(census[0], img[0]);                                                             // :380
    roo::Census(census[1], img[1]);
    roo::CensusStereoVolume<float, census_t>(vol[0], census[0], census[1], maxdisp, -1);       // :384
    roo::CensusStereoVolume<float, census_t>(vol[1], census[1], census[0], maxdisp, +1);       // :385
    roo::SemiGlobalMatching<float, float, float>(vol[2], vol[0], img[0], maxdisp, 0.01f, 0.02f, true, true, true);   // :425
    vol[0].CopyFrom(vol[2]);                                                                    // :426
    roo::CostVolMinimumSubpix(disp[0], vol[0], maxdisp, -1);                                    // :431
    roo::CostVolMinimumSubpix(disp[1], vol[1], maxdisp, +1);                                    // :432
    roo::CostVolMinimum<float, float>(disp[0], vol[0], maxdisp);                                // :434
    roo::MedianFilterRejectNegative9x9(disp[0], disp[0], 50);                                   // :440 (in place, like the application)
    roo::MedianFilterRejectNegative7x7(disp[0], disp[0], 50);                                   // :441
    roo::MedianFilterRejectNegative5x5(disp[0], disp[0], 50);                                   // :442
    roo::LeftRightCheck(disp[1], disp[0], +1, 1.0f);                                            // :452
    roo::LeftRightCheck(disp[0], disp[1], -1, 1.0f);                                            // :453
    roo::FilterDispGrad(dispf, disp[0], 0.5f);                                                  // :457
    roo::Disp2Depth(disp[0], depth, 500.0f, 0.1f);
    roo::DisparityImageToVbo(vbo, disp[0], 0.1f, 500.0f, 500.0f, 160.0f, 48.0f);               // :500

    // ---- the rest of the instantiation set (cu_census.cu:180-220,309-314; cu_semi_global_matching.cu:88-89;
    //      cu_dense_stereo.cu:54-60,735-763,512-546,580-627)
    roo::Census(census1, upload); roo::Census(census2, upload); roo::Census(census[0], upload);
    roo::Census(census1, img[0]); roo::Census(census2, img[0]);
    roo::CensusStereo(dispi8, census1, census1r, maxdisp);
    roo::CensusStereoVolume<float, unsigned long>(vol[0], census1, census1r, maxdisp, -1);
    roo::CensusStereoVolume<float, ulong2>(vol[0], census2, census2r, maxdisp, -1);
    roo::CensusStereoVolume<unsigned short, unsigned long>(volu, census1, census1r, maxdisp, -1);
    roo::CensusStereoVolume<unsigned short, ulong2>(volu, census2, census2r, maxdisp, -1);
    roo::CensusStereoVolume<unsigned short, ulong4>(volu, census[0], census[1], maxdisp, -1);
    roo::SemiGlobalMatching<float, roo::CostVolElem, unsigned char>(vol[2], vole, upload, maxdisp, 0.01f, 0.02f, true, true, true);
    roo::CostVolMinimum<char, float>(dispi8, vol[0], maxdisp);
    roo::CostVolMinimum<char, unsigned short>(dispi8, volu, maxdisp);
    roo::CostVolMinimum<float, unsigned short>(disp[0], volu, maxdisp);
    roo::CostVolMinimum(disp[0], vole);
    roo::DenseStereoSubpixelRefine(dispf, disp_c, upload, upload);
    roo::CostVolMinimumSquarePenaltySubpix(disp[0], vol[0], dispf, maxdisp, -1, 1.0f, 0.5f);      // stereo/main.cpp:376
    roo::BilateralFilter<float, float, float>(dispf, depth, img[0], 2.0f, 0.2f, 0.1f, 3);         // stereo2/main.cpp:417 (one slice)
    roo::BilateralFilter<float, float, unsigned char>(dispf, depth, upload, 2.0f, 0.2f, 10.0f, 3);
    roo::BilateralFilterVolume<float>(vol[1], vol[0], img[0], 2.0f, 0.2f, 0.1f, 2, maxdisp);             // the loop at :407-421 in one launch
    {   // the guided-filter branch of the frame loop, spelled as stereo2/main.cpp:392-405 spells it, and its one-call form
        roo::Image<unsigned char, roo::TargetDevice, roo::Manage> Scratch(w * sizeof(float), h);          // :186
        roo::Image<float, roo::TargetDevice, roo::Manage> meanI(w, h), varI(w, h);
        roo::Image<float, roo::TargetDevice, roo::Manage> temp[] = {{(size_t)w, (size_t)h}, {(size_t)w, (size_t)h}, {(size_t)w, (size_t)h}, {(size_t)w, (size_t)h}, {(size_t)w, (size_t)h}};
        const int rad = 9;
        const float eps = 0.01f;
        roo::Image<float, roo::TargetDevice, roo::Manage>& I = img[0];
        roo::ComputeMeanVarience<float, float, float>(varI, temp[0], meanI, I, Scratch, rad);       // :396
        for (int d = 0; d < 2; ++d) {
            roo::Image<float> P = vol[0].ImageXY(d);
            roo::ComputeCovariance(temp[0], temp[2], temp[1], P, meanI, I, Scratch, rad);            // :401
            roo::GuidedFilter(P, temp[0], varI, temp[1], meanI, I, Scratch, temp[2], temp[3], temp[4], rad, eps);   // :402
        }
        roo::GuidedFilterVolume(vol[1], I, rad, eps, maxdisp);
        roo::BoxFilter<float, float, float>(temp[0], img[1], Scratch, 5);                           // :377 (commented out there)
    }
    roo::DenseStereo<unsigned char, unsigned char>(disp_c, upload, upload, (unsigned char)maxdisp, 0.05f, 2);   // cu_dense_stereo.cu:405
    roo::DenseStereo<char, unsigned char>(dispi8, upload, upload, (char)-40, 0.05f, 0);                          // :406
    roo::LeftRightCheck(dispi8, dispi8r, -1, 0);
    const cudaError_t err = cudaDeviceSynchronize();
    std::printf("%s\n", err == cudaSuccess ? "OK" : cudaGetErrorString(err));
    return err == cudaSuccess ? 0 : 1;
}
