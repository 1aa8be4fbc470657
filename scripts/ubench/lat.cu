// latency micro-benchmarks for the ops on the SGM recurrence's critical path (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float redux_min(float v){ float m; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v)); return m; }
__device__ __forceinline__ float shfl_min(float v){
  #pragma unroll
  for (int o=16;o>0;o>>=1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v; }
__device__ __forceinline__ int redux_min_s32(int v){ return __reduce_min_sync(0xffffffffu, v); }
template<int OP> __global__ void k(float* out, long long* cyc, int iters){
  float v = out[threadIdx.x];
  long long t0 = clock64();
  for (int i=0;i<iters;++i){
    if (OP==0) v = redux_min(v) + 1.0f;
    else if (OP==1) v = shfl_min(v) + 1.0f;
    else if (OP==2) v = __shfl_up_sync(0xffffffffu, v, 1) + 1.0f;
    else if (OP==3) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); v = r + 1.0f; }
    else if (OP==4) v = (float)((__float_as_uint(v) >> 8) & 0xffu) + 1.0f;
    else if (OP==5) v = __int_as_float(redux_min_s32(__float_as_int(v))) + 1.0f;
    else if (OP==6) v = fminf(v, 3.0f) + 1.0f;
  }
  long long t1 = clock64();
  out[threadIdx.x] = v; if (threadIdx.x==0) cyc[0] = t1-t0;
}
int main(){
  float* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8); cudaMemset(d, 0, 4096);
  const char* names[] = {"redux.min.f32+fadd","shfl-tree min+fadd","shfl.up+fadd","rcp.approx+fadd","shift+and+i2f+fadd","redux.min.s32+fadd","fmnmx+fadd"};
  const int iters = 4096;
  for (int op=0; op<7; ++op){
    for (int rep=0; rep<2; ++rep){
      switch(op){case 0:k<0><<<1,32>>>(d,c,iters);break;case 1:k<1><<<1,32>>>(d,c,iters);break;case 2:k<2><<<1,32>>>(d,c,iters);break;case 3:k<3><<<1,32>>>(d,c,iters);break;case 4:k<4><<<1,32>>>(d,c,iters);break;case 5:k<5><<<1,32>>>(d,c,iters);break;case 6:k<6><<<1,32>>>(d,c,iters);break;}
      cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h,c,8,cudaMemcpyDeviceToHost);
    printf("%-24s %.1f cycles/iter\n", names[op], (double)h/iters);
  }
  return 0;
}
