// How fast can a warp stream warp-uniform weights from constant memory into uniform registers (LDCU.128) next
// to FFMA2 work?  Kernel R issues R FFMA2 (uniform .F32 broadcast operand) per LDCU.128, all static addresses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 ldcu_rate.cu -o ldcu_rate && ./ldcu_rate
#include <cstdio>
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>
extern "C" { __constant__ __align__(16) float cw[4096]; }
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi){ f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c){ f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template<int OFF> __device__ __forceinline__ void ldc4(float* v){ asm volatile("ld.const.v4.f32 {%0,%1,%2,%3}, [cw+%4];" : "=f"(v[0]),"=f"(v[1]),"=f"(v[2]),"=f"(v[3]) : "n"(OFF*4)); }
template <int I, int N, class F> __device__ __forceinline__ void static_for(F&& f){ if constexpr (I<N){ f(std::integral_constant<int,I>{}); static_for<I+1,N>(f);} }
constexpr int NLOAD = 96;   // LDCU.128 per loop iteration
// R = FFMA2 per load (uses up to 4 weights of the load; R>4 re-uses weights with different h)
template<int R>
__global__ void k(float* out, int iters, float a0){
  f2 acc[10], h[4];
  for (int i=0;i<10;++i) acc[i]=pk(threadIdx.x*1e-3f+i, i);
  for (int i=0;i<4;++i) h[i]=pk(a0+i*1e-3f+threadIdx.x*1e-6f, a0-i*1e-3f);
#pragma unroll 1
  for (int it=0; it<iters; ++it){
    static_for<0,NLOAD>([&](auto Q){ constexpr int q=decltype(Q)::value;
      float w[4]; ldc4<4*q>(w);
#pragma unroll
      for (int r=0;r<R;++r) acc[(q*R+r)%10] = fma2(h[(r/4)%4], pk(w[r%4],w[r%4]), acc[(q*R+r)%10]);
    });
  }
  f2 s=0; for(int i=0;i<10;++i) s^=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=__uint_as_float((uint32_t)s)+__uint_as_float((uint32_t)(s>>32));
}
template<int R> void run(int sms, int occ){
  float* out; cudaMalloc(&out, sizeof(float)*sms*occ*128*4);
  int iters=2000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<R><<<sms*occ,128>>>(out,10,1.f); cudaDeviceSynchronize();
  float best=1e9; for(int rep=0;rep<5;++rep){ cudaEventRecord(e0); k<R><<<sms*occ,128>>>(out,iters,1.f); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); best=ms<best?ms:best; }
  double clk=1.965e9; double cyc = best*1e-3*clk;           // cycles elapsed
  double warps_per_smsp = occ*4/4.0;                        // 4 warps per block, 4 SMSPs
  double loads_per_smsp = (double)iters*NLOAD*warps_per_smsp;
  printf("{\"ffma2_per_ldcu128\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"cycles_per_ldcu128_per_smsp\": %.2f, \"fma_per_clk_per_sm\": %.1f}\n", R, occ, best, cyc/loads_per_smsp, (double)iters*NLOAD*R*2*32*4*occ/cyc);
  cudaFree(out);
}
int main(){ cudaDeviceProp p; cudaGetDeviceProperties(&p,0); int sms=p.multiProcessorCount;
  for (int occ : {2,4}) { run<0>(sms,occ); run<1>(sms,occ); run<2>(sms,occ); run<3>(sms,occ); run<4>(sms,occ); run<6>(sms,occ); run<8>(sms,occ); }
  return 0; }
