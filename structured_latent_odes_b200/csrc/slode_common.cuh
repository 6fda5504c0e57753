// Shared device helpers and host-side error plumbing for the slode_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "slode_b200.h"

namespace slode {

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SLODE_CUDA_TRY(expr)                                        \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return ::slode::cuda_fail(_e, #expr);    \
  } while (0)

// Per-device serialisation of the packed-weight constant buffer (see slode_pack.cu).
struct PackGuard {
  explicit PackGuard(cudaStream_t s);
  ~PackGuard();
  int status;  // 0 ok
  cudaStream_t stream;
  float* staging;  // device scratch the pack kernel writes, then copied into constant memory
  int sms;         // SM count of the current device
};

// Per-device scratch of the reverse sweep (flip records, see slode_mlp_kernels.cuh); grown on demand, owned by
// the library, valid while the caller holds a PackGuard (which serialises the kernels that use it).
float* flip_workspace(size_t bytes);

// launch counters reported by slode_query: kernels launched by the last forward / backward entry-point call of
// the process (autograd runs backward calls on its own thread, so these are process-wide), and the running total
struct LaunchCount {
  std::atomic<int> last{0};
  LaunchCount& operator=(int v);
  operator int() const { return last.load(); }
};
extern LaunchCount g_fwd_launches;
extern LaunchCount g_bwd_launches;
extern std::atomic<long long> g_total_launches;

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
constexpr float kNegLog2e = -1.4426950408889634f;
constexpr float kOneThird = (float)(1.0 / 3.0);   // torch casts the python double to fp32
constexpr float kTwoThirds = (float)(2.0 / 3.0);

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(u) given v = -log2(e) * u  (the scale is folded into the packed weights)
__device__ __forceinline__ float sigmoid_from_scaled(float v) { return rcp_approx(1.0f + ex2_approx(v)); }

__device__ __forceinline__ float ld_stream(const float* p) {  // read-once data: bypass L1 allocation
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

}  // namespace slode
