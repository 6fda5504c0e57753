// FP32-pipe microbenchmarks for the roofline denominator and the kernel design (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp32_pipes profiles/microbench/fp32_pipes.cu
// Prints one JSON line per variant: achieved FMA/clk/SM and TFLOP/s (2 flop per FMA).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int NACC = 16;
constexpr int ITERS = 4096;

__constant__ float c_w[1024];

// (1) three distinct register operands
__global__ void k_ffma_rrr(float* out, float a0, float b0) {
  float acc[NACC], a[NACC], b[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] = threadIdx.x * 1e-3f + i; a[i] = a0 + i * 1e-3f + threadIdx.x * 1e-6f; b[i] = b0 + i * 2e-3f; }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = fmaf(a[(i + r) % NACC], b[(i + 2 * r + 1) % NACC], acc[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (2) one operand warp-uniform, hoisted (few constants -> uniform registers, no loads in the loop)
__global__ void k_ffma_rur(float* out, float a0) {
  float acc[NACC], a[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] = threadIdx.x * 1e-3f + i; a[i] = a0 + i * 1e-3f + threadIdx.x * 1e-6f; }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = fmaf(a[(i + r) % NACC], c_w[(i + 4 * r) % 16], acc[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (3) streaming a 512-weight set from constant memory: one uniform load per 4 FFMA per trajectory
template <int NTRAJ>
__global__ void k_ffma_stream(float* out, float a0) {
  float acc[NTRAJ][8], h[NTRAJ][8];
#pragma unroll
  for (int q = 0; q < NTRAJ; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[q][i] = threadIdx.x * 1e-3f + i + q; h[q][i] = a0 + i * 1e-3f + threadIdx.x * 1e-6f + q; }
#pragma unroll 1
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int w = 0; w < 512; ++w)
#pragma unroll
      for (int q = 0; q < NTRAJ; ++q) acc[q][w % 8] = fmaf(h[q][(w / 8) % 8], c_w[w], acc[q][w % 8]);
  }
  float s = 0;
#pragma unroll
  for (int q = 0; q < NTRAJ; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[q][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (4) packed fp32x2 FMA (Blackwell FFMA2)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__global__ void k_ffma2_rrr(float* out, float a0, float b0) {
  uint64_t acc[NACC], a[NACC], b[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    acc[i] = pack2(threadIdx.x * 1e-3f + i, i);
    a[i] = pack2(a0 + i * 1e-3f + threadIdx.x * 1e-6f, a0 - i * 1e-3f);
    b[i] = pack2(b0 + i * 2e-3f, b0 - i * 2e-3f);
  }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = ffma2(a[(i + r) % NACC], b[(i + 2 * r + 1) % NACC], acc[i]);
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((uint32_t)s) + __uint_as_float((uint32_t)(s >> 32));
}

// (5) FFMA2 with a warp-uniform weight pair streamed from constant memory, 2 trajectories x output pairs
__global__ void k_ffma2_stream(float* out, float a0) {
  // acc[q][p]: trajectory q, output pair p (4 pairs = 8 outputs); h[q][j] duplicated into both halves
  uint64_t acc[2][4], h[2][8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[q][p] = pack2(threadIdx.x * 1e-3f + p + q, p);
#pragma unroll
    for (int j = 0; j < 8; ++j) { float v = a0 + j * 1e-3f + threadIdx.x * 1e-6f + q; h[q][j] = pack2(v, v); }
  }
  const float2* w2 = reinterpret_cast<const float2*>(c_w);
#pragma unroll 1
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int w = 0; w < 256; ++w) {  // 256 weight pairs = 512 weights
      const float2 ww = w2[w];
      const uint64_t wp = pack2(ww.x, ww.y);
#pragma unroll
      for (int q = 0; q < 2; ++q) acc[q][w % 4] = ffma2(h[q][(w / 4) % 8], wp, acc[q][w % 4]);
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int p = 0; p < 4; ++p) s ^= acc[q][p];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((uint32_t)s) + __uint_as_float((uint32_t)(s >> 32));
}

// (5b) FFMA2 with two trajectories packed in the register pair and ONE warp-uniform weight broadcast
//      to both halves: acc(traj0,traj1)[k] += W_kj * h(traj0,traj1)[j].  One uniform load per 4 FFMA2.
template <int NPAIR>
__global__ void k_ffma2_bcastw(float* out, float a0) {
  uint64_t acc[NPAIR][8], h[NPAIR][8];
#pragma unroll
  for (int q = 0; q < NPAIR; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[q][i] = pack2(threadIdx.x * 1e-3f + i + q, i);
      h[q][i] = pack2(a0 + i * 1e-3f + threadIdx.x * 1e-6f + q, a0 - i * 1e-3f);
    }
#pragma unroll 1
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int w = 0; w < 512; ++w) {
      const float ww = c_w[w];
      const uint64_t wp = pack2(ww, ww);
#pragma unroll
      for (int q = 0; q < NPAIR; ++q) acc[q][w % 8] = ffma2(h[q][(w / 8) % 8], wp, acc[q][w % 8]);
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int q = 0; q < NPAIR; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= acc[q][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((uint32_t)s) + __uint_as_float((uint32_t)(s >> 32));
}

// (6) sigmoid pipe: ex2 + add + rcp
__global__ void k_mufu(float* out, float a0) {
  float v[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) v[i] = a0 + i * 1e-2f + threadIdx.x * 1e-5f;
#pragma unroll 1
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      float e, r;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[i]));
      e += 1.0f;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e));
      v[i] = r - 0.75f;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static int run(const char* name, F launch, double fma_per_thread, int blocks, int threads, int sms, double clk_ghz_nominal) {
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CHECK(cudaEventRecord(e0));
    launch();
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CHECK(cudaGetLastError());
  const double fma = fma_per_thread * (double)blocks * threads;
  const double tflops = 2.0 * fma / (best * 1e-3) / 1e12;
  const double per_clk_sm = fma / (best * 1e-3) / (clk_ghz_nominal * 1e9) / sms;
  printf("{\"variant\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f, \"fma_per_clk_per_sm_at_%.3fGHz\": %.1f, \"blocks\": %d, \"threads\": %d}\n",
         name, best, tflops, clk_ghz_nominal, per_clk_sm, blocks, threads);
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
  std::vector<float> w(1024);
  for (int i = 0; i < 1024; ++i) w[i] = 1e-3f * ((i * 37) % 101 - 50);
  CHECK(cudaMemcpyToSymbol(c_w, w.data(), sizeof(float) * 1024));
  float* out;
  const int threads = 256;
  for (int occ : {2, 4, 8}) {
    const int blocks = sms * occ;
    CHECK(cudaMalloc(&out, sizeof(float) * blocks * threads));
    char name[64];
    snprintf(name, sizeof name, "ffma_rrr_occ%d", occ);
    if (run(name, [&] { k_ffma_rrr<<<blocks, threads>>>(out, 1.0001f, 0.9999f); }, (double)ITERS * 4 * NACC, blocks, threads, sms, ghz)) return 1;
    snprintf(name, sizeof name, "ffma_rur_occ%d", occ);
    if (run(name, [&] { k_ffma_rur<<<blocks, threads>>>(out, 1.0001f); }, (double)ITERS * 4 * NACC, blocks, threads, sms, ghz)) return 1;
    snprintf(name, sizeof name, "ffma_stream_1traj_occ%d", occ);
    if (run(name, [&] { k_ffma_stream<1><<<blocks, threads>>>(out, 1.0001f); }, (double)(ITERS / 8) * 512, blocks, threads, sms, ghz)) return 1;
    snprintf(name, sizeof name, "ffma_stream_2traj_occ%d", occ);
    if (run(name, [&] { k_ffma_stream<2><<<blocks, threads>>>(out, 1.0001f); }, (double)(ITERS / 8) * 512 * 2, blocks, threads, sms, ghz)) return 1;
    snprintf(name, sizeof name, "ffma_stream_4traj_occ%d", occ);
    if (run(name, [&] { k_ffma_stream<4><<<blocks, threads>>>(out, 1.0001f); }, (double)(ITERS / 8) * 512 * 4, blocks, threads, sms, ghz)) return 1;
    snprintf(name, sizeof name, "ffma2_rrr_occ%d", occ);
    if (run(name, [&] { k_ffma2_rrr<<<blocks, threads>>>(out, 1.0001f, 0.9999f); }, (double)ITERS * 4 * NACC * 2, blocks, threads, sms, ghz)) return 1;
    snprintf(name, sizeof name, "ffma2_stream_occ%d", occ);
    if (run(name, [&] { k_ffma2_stream<<<blocks, threads>>>(out, 1.0001f); }, (double)(ITERS / 8) * 256 * 2 * 2, blocks, threads, sms, ghz)) return 1;
    snprintf(name, sizeof name, "ffma2_bcastw_1pair_occ%d", occ);
    if (run(name, [&] { k_ffma2_bcastw<1><<<blocks, threads>>>(out, 1.0001f); }, (double)(ITERS / 8) * 512 * 2, blocks, threads, sms, ghz)) return 1;
    snprintf(name, sizeof name, "ffma2_bcastw_2pair_occ%d", occ);
    if (run(name, [&] { k_ffma2_bcastw<2><<<blocks, threads>>>(out, 1.0001f); }, (double)(ITERS / 8) * 512 * 4, blocks, threads, sms, ghz)) return 1;
    snprintf(name, sizeof name, "mufu_sigmoid_occ%d(count=sigmoids)", occ);
    if (run(name, [&] { k_mufu<<<blocks, threads>>>(out, 0.1f); }, (double)(ITERS / 4) * NACC, blocks, threads, sms, ghz)) return 1;
    CHECK(cudaFree(out));
  }
  return 0;
}
