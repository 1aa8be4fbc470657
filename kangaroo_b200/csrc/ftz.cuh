// FMUL / FADD / FFMA.FTZ as the reference's -use_fast_math build issues them (denormal inputs and results flush to zero);
// used by the operators that reproduce that build's SASS operation for operation (volfilter.cu, gfilter.cu).
#pragma once

namespace roo_b200 {

__device__ __forceinline__ float fmul_ftz(float a, float b) { float r; asm("mul.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fadd_ftz(float a, float b) { float r; asm("add.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ffma_ftz(float a, float b, float c) { float r; asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

}  // namespace roo_b200
