// Dispatch of the horizontal sweep to the translation unit that holds the instantiations of its disparity count
// (sgm_hsweep.cu is compiled once per DPL with -DHS_PART=<DPL>; see the Makefile).
#include "common.cuh"
#include "kernels.cuh"

namespace roo_b200 {
int launch_hsweep_dpl1(const SweepArgs& a, cudaStream_t st);
int launch_hsweep_dpl2(const SweepArgs& a, cudaStream_t st);
int launch_hsweep_dpl4(const SweepArgs& a, cudaStream_t st);
int launch_hsweep_dpl8(const SweepArgs& a, cudaStream_t st);
int launch_hsweep_dpl16(const SweepArgs& a, cudaStream_t st);

int launch_hsweep(const SweepArgs& a, cudaStream_t st) {
    switch (a.DP) {
        case 32: return launch_hsweep_dpl1(a, st);
        case 64: return launch_hsweep_dpl2(a, st);
        case 128: return launch_hsweep_dpl4(a, st);
        case 256: return launch_hsweep_dpl8(a, st);
        case 512: return launch_hsweep_dpl16(a, st);
        default: return ROO_ERR_UNSUPPORTED;
    }
}
}  // namespace roo_b200
