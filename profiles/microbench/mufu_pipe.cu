// MUFU (XU pipe) throughput on sm_100a: ex2-only, rcp-only and the sigmoid mix, as MUFU lane-operations per clock
// and SM.  The roofline's XU denominator (bench.py XU_PEAK_TOPS = 148 SM x 16 lanes x 1.965 GHz) is checked here.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mufu_pipe profiles/microbench/mufu_pipe.cu
#include <cuda_runtime.h>
#include <stdio.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int NACC = 16;
constexpr int ITERS = 2048;

template <int KIND>  // 0: ex2 only, 1: rcp only, 2: ex2 + rcp (sigmoid chain, one FADD between), 3: ex2 with an independent FFMA per MUFU
__global__ void k(float* out, float a0) {
  float v[NACC], w[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { v[i] = a0 + i * 1e-2f + threadIdx.x * 1e-5f; w[i] = 1.0f + i; }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      float e;
      if (KIND == 0) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[i])); v[i] = e; }
      if (KIND == 1) { asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[i])); v[i] = e; }
      if (KIND == 2) {
        float r;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[i]));
        e += 1.0f;
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e));
        v[i] = r - 0.75f;
      }
      if (KIND == 3) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[i])); v[i] = e; w[i] = fmaf(w[i], 0.999f, 1e-3f); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += v[i] + w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
int run(const char* name, double mufu_per_thread, int blocks, int threads, int sms, double ghz, float* out) {
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) k<KIND><<<blocks, threads>>>(out, 0.1f);
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CHECK(cudaEventRecord(e0));
    k<KIND><<<blocks, threads>>>(out, 0.1f);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double ops = mufu_per_thread * (double)blocks * threads;
  printf("{\"variant\": \"%s\", \"ms\": %.4f, \"mufu_Tops\": %.3f, \"mufu_lane_ops_per_clk_per_sm_at_%.3fGHz\": %.2f, \"blocks\": %d, \"threads\": %d}\n",
         name, best, ops / (best * 1e-3) / 1e12, ghz, ops / (best * 1e-3) / (ghz * 1e9) / sms, blocks, threads);
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CHECK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double ghz = clk_khz * 1e-6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", prop.name, sms, ghz);
  const int threads = 256;
  for (int occ : {2, 4, 8}) {
    const int blocks = sms * occ;
    float* out;
    CHECK(cudaMalloc(&out, sizeof(float) * blocks * threads));
    char n[64];
    snprintf(n, sizeof n, "ex2_only_occ%d", occ);
    if (run<0>(n, (double)ITERS * NACC, blocks, threads, sms, ghz, out)) return 1;
    snprintf(n, sizeof n, "rcp_only_occ%d", occ);
    if (run<1>(n, (double)ITERS * NACC, blocks, threads, sms, ghz, out)) return 1;
    snprintf(n, sizeof n, "sigmoid_ex2_rcp_occ%d", occ);
    if (run<2>(n, (double)ITERS * NACC * 2, blocks, threads, sms, ghz, out)) return 1;
    snprintf(n, sizeof n, "ex2_plus_ffma_occ%d", occ);
    if (run<3>(n, (double)ITERS * NACC, blocks, threads, sms, ghz, out)) return 1;
    CHECK(cudaFree(out));
  }
  return 0;
}
