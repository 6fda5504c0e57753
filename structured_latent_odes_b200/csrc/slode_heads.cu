// Decoder quantile / mean heads on the latent trajectories (sm_100a).
//
// Reference: Decoder.forward / GaussianDecoder.forward (models/decoders.py:42-54, 84-91):
//     mu_q = Linear_q(solution).permute(0, 2, 1)        Linear_q: ode_state_dim -> obs_dim, no bias
// for q in {q50, q75, q25} (Decoder) or the single mean head (GaussianDecoder); solution is (B,T,S), mu_q (B,O,T).
// In the reference this is one tiny-K matmul + a permuted view per head (and, backwards, two more matmuls per
// head); here one pass over the trajectories produces all heads in their final (B,O,T) layout, and one pass
// backwards produces dL/dsolution and the head-weight gradients.  HBM-bound: 4*S bytes read and 4*NQ*O bytes
// written per (trajectory, time) forward; 4*(S + NQ*O) read and 4*S written backward.
//
// One thread = one (trajectory, time) point, time fastest: with (B,T,S)-contiguous storage of the solution
// (layout="bts") both the S-float read and the per-(q,o) writes along T are coalesced.
#include <algorithm>

#include "slode_common.cuh"

namespace slode {
namespace heads {

constexpr int kBlock = 256;
constexpr int kMaxW = 3 * 8 * 8;  // NQ * O * S

// Backward: QM = compile-time bound on the number of (head, output) pairs NQ*O of a launch (3, 9 or 24): the loops over
// them are unrolled and predicated, and the head-weight gradients accumulate in registers over the thread's whole
// grid-stride loop.  (The first version did one warp reduction + shared-memory atomic per pair and POINT: 850
// instructions per point; this one ~300.  Both run at 1.75 TB/s -- the kernel is bound by memory-level parallelism,
// 14 scalar loads per point -- so the gain is issue slots, not time.)
// Forward: run-time loops over (q, o) -- measured FASTER than the unrolled / predicated form the backward uses
// (1.53 vs 2.28 ms at 2^20 x 100, three heads x obs_dim 3, profiles/r02): the kernel is bound by its 9 scattered
// output streams, not by instructions, and the rolled loop spreads its stores out in time.
template <int S>
__global__ void __launch_bounds__(kBlock)
heads_fwd_kernel(int64_t B, int T, int O, int NQ, const float* __restrict__ sol, int64_t st, int64_t sb,
                 const float* __restrict__ W, float* __restrict__ mu) {
  __shared__ float sW[kMaxW];
  for (int i = threadIdx.x; i < NQ * O * S; i += kBlock) sW[i] = W[i];
  __syncthreads();
  const int64_t n = B * (int64_t)T;
  for (int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * kBlock) {
    const int64_t b = idx / T;
    const int t = (int)(idx - b * T);
    const float* x = sol + b * sb + (int64_t)t * st;
    float xs[S];
#pragma unroll
    for (int s = 0; s < S; ++s) xs[s] = __ldg(x + s);
    for (int q = 0; q < NQ; ++q) {
      for (int o = 0; o < O; ++o) {
        const float* w = sW + (q * O + o) * S;
        float acc = 0.0f;
#pragma unroll
        for (int s = 0; s < S; ++s) acc = fmaf(w[s], xs[s], acc);
        mu[(((int64_t)q * B + b) * O + o) * T + t] = acc;
      }
    }
  }
}

template <int S, int QM>
__global__ void __launch_bounds__(kBlock)
heads_bwd_kernel(int64_t B, int T, int O, int NQ, const float* __restrict__ sol, int64_t st, int64_t sb,
                 const float* __restrict__ W, const float* __restrict__ g0, const float* __restrict__ g1,
                 const float* __restrict__ g2, float* __restrict__ gsol, int64_t gst, int64_t gsb,
                 float* __restrict__ gW) {
  // g0, g1, g2: dL/dmu of head 0, 1, 2, each (B,O,T) contiguous, or null (that head did not enter the loss)
  __shared__ float sW[QM * S];
  __shared__ float sG[QM * S];
  __shared__ const float* sPtr[QM];
  const int nqo = NQ * O;
  for (int i = threadIdx.x; i < QM * S; i += kBlock) {
    sW[i] = i < nqo * S ? W[i] : 0.0f;
    sG[i] = 0.0f;
  }
  for (int i = threadIdx.x; i < QM; i += kBlock) {
    const int q = i / O, o = i - q * O;
    const float* base = q == 0 ? g0 : (q == 1 ? g1 : g2);
    sPtr[i] = (i < nqo && base) ? base + (int64_t)o * T : nullptr;
  }
  __syncthreads();
  float acc[QM][S];   // this thread's share of dL/dW[q][o][s] = sum over its points of grad_mu * x_s
#pragma unroll
  for (int qo = 0; qo < QM; ++qo) {
#pragma unroll
    for (int s = 0; s < S; ++s) acc[qo][s] = 0.0f;
  }
  const int64_t n = B * (int64_t)T, OT = (int64_t)O * T;
  for (int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * kBlock) {
    const int64_t b = idx / T;
    const int t = (int)(idx - b * T);
    const float* x = sol + b * sb + (int64_t)t * st;
    const int64_t off = b * OT + t;
    float xs[S], gs[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      xs[s] = __ldg(x + s);
      gs[s] = 0.0f;
    }
#pragma unroll
    for (int qo = 0; qo < QM; ++qo) {
      if (qo < nqo) {
        const float* gp = sPtr[qo];
        const float g = gp ? __ldg(gp + off) : 0.0f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          gs[s] = fmaf(g, sW[qo * S + s], gs[s]);
          acc[qo][s] = fmaf(g, xs[s], acc[qo][s]);
        }
      }
    }
    float* gx = gsol + b * gsb + (int64_t)t * gst;
#pragma unroll
    for (int s = 0; s < S; ++s) gx[s] = gs[s];
  }
  // once per thread: warp sums -> block sums in shared memory -> one global atomic per element and block
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int qo = 0; qo < QM; ++qo) {
    if (qo < nqo) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float v = acc[qo][s];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) atomicAdd(&sG[qo * S + s], v);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nqo * S; i += kBlock) atomicAdd(gW + i, sG[i]);
}

// Vectorised variants: one thread = FOUR consecutive time points of one trajectory.  With (B,T,S)-contiguous storage
// and T % 4 == 0 the four state rows are 4*S contiguous floats (S 16-byte loads), every head output / head gradient row
// contributes one 16-byte access, and a thread has 14 (NQ*O = 9, S = 5) 16-byte loads in flight instead of 14 4-byte
// ones: the scalar kernels are bound by memory-level parallelism (1.75 TB/s backward), not by bytes.
template <int S, int QM>
__global__ void __launch_bounds__(kBlock)
heads_fwd_vec_kernel(int64_t B, int T, int O, int NQ, const float* __restrict__ sol, int64_t sb,
                     const float* __restrict__ W, float* __restrict__ mu) {
  __shared__ float sW[QM * S];
  __shared__ int64_t sOff[QM];
  const int nqo = NQ * O;
  for (int i = threadIdx.x; i < QM * S; i += kBlock) sW[i] = i < nqo * S ? W[i] : 0.0f;
  for (int i = threadIdx.x; i < QM; i += kBlock) {
    const int q = i / O, o = i - q * O;
    sOff[i] = ((int64_t)q * B * O + o) * T;
  }
  __syncthreads();
  const int T4 = T / 4;
  const int64_t n = B * (int64_t)T4, OT = (int64_t)O * T;
  for (int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * kBlock) {
    const int64_t b = idx / T4;
    const int t = 4 * (int)(idx - b * T4);
    const float4* x4 = reinterpret_cast<const float4*>(sol + b * sb + (int64_t)t * S);
    float x[4 * S];
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const float4 v = __ldg(x4 + k);
      x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
    }
    float* out = mu + b * OT + t;
#pragma unroll
    for (int qo = 0; qo < QM; ++qo) {
      if (qo < nqo) {
        float a[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float w = sW[qo * S + s];
#pragma unroll
          for (int tt = 0; tt < 4; ++tt) a[tt] = fmaf(w, x[tt * S + s], a[tt]);
        }
        *reinterpret_cast<float4*>(out + sOff[qo]) = make_float4(a[0], a[1], a[2], a[3]);
      }
    }
  }
}

template <int S, int QM>
__global__ void __launch_bounds__(kBlock)
heads_bwd_vec_kernel(int64_t B, int T, int O, int NQ, const float* __restrict__ sol, int64_t sb,
                     const float* __restrict__ W, const float* __restrict__ g0, const float* __restrict__ g1,
                     const float* __restrict__ g2, float* __restrict__ gsol, int64_t gsb, float* __restrict__ gW) {
  __shared__ float sW[QM * S];
  __shared__ float sG[QM * S];
  __shared__ const float* sPtr[QM];
  const int nqo = NQ * O;
  for (int i = threadIdx.x; i < QM * S; i += kBlock) {
    sW[i] = i < nqo * S ? W[i] : 0.0f;
    sG[i] = 0.0f;
  }
  for (int i = threadIdx.x; i < QM; i += kBlock) {
    const int q = i / O, o = i - q * O;
    const float* base = q == 0 ? g0 : (q == 1 ? g1 : g2);
    sPtr[i] = (i < nqo && base) ? base + (int64_t)o * T : nullptr;
  }
  __syncthreads();
  float acc[QM][S];
#pragma unroll
  for (int qo = 0; qo < QM; ++qo) {
#pragma unroll
    for (int s = 0; s < S; ++s) acc[qo][s] = 0.0f;
  }
  const int T4 = T / 4;
  const int64_t n = B * (int64_t)T4, OT = (int64_t)O * T;
  for (int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * kBlock) {
    const int64_t b = idx / T4;
    const int t = 4 * (int)(idx - b * T4);
    const float4* x4 = reinterpret_cast<const float4*>(sol + b * sb + (int64_t)t * S);
    float x[4 * S], gs[4 * S];
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const float4 v = __ldg(x4 + k);
      x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < 4 * S; ++k) gs[k] = 0.0f;
    const int64_t off = b * OT + t;
#pragma unroll
    for (int qo = 0; qo < QM; ++qo) {
      if (qo < nqo) {
        const float* gp = sPtr[qo];
        const float4 g4 = gp ? __ldg(reinterpret_cast<const float4*>(gp + off)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float w = sW[qo * S + s];
#pragma unroll
          for (int tt = 0; tt < 4; ++tt) {
            gs[tt * S + s] = fmaf(g[tt], w, gs[tt * S + s]);
            acc[qo][s] = fmaf(g[tt], x[tt * S + s], acc[qo][s]);
          }
        }
      }
    }
    float4* o4 = reinterpret_cast<float4*>(gsol + b * gsb + (int64_t)t * S);
#pragma unroll
    for (int k = 0; k < S; ++k) o4[k] = make_float4(gs[4 * k], gs[4 * k + 1], gs[4 * k + 2], gs[4 * k + 3]);
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int qo = 0; qo < QM; ++qo) {
    if (qo < nqo) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float v = acc[qo][s];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) atomicAdd(&sG[qo * S + s], v);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nqo * S; i += kBlock) atomicAdd(gW + i, sG[i]);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// resident blocks per SM of a kernel (cached per kernel)
template <class K>
static int blocks_per_sm(K kern) {
  static int cached = 0;
  if (cached == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kBlock, 0) != cudaSuccess || n < 1) n = 1;
    cached = n;
  }
  return cached;
}

static int device_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

static int check(const char* who, int64_t B, int T, int S, int O, int NQ) {
  if (B < 0 || T < 1 || O < 1 || O > 8 || NQ < 1 || NQ > 3) {
    set_error("%s: bad sizes B=%lld T=%d O=%d NQ=%d (O <= 8, NQ <= 3)", who, (long long)B, T, O, NQ);
    return SLODE_EINVAL;
  }
  if (S != 4 && S != 5 && S != 8) {
    set_error("%s: ode_state_dim=%d is not compiled in (4, 5, 8)", who, S);
    return SLODE_EUNSUPPORTED;
  }
  return SLODE_OK;
}

}  // namespace heads
}  // namespace slode

using namespace slode;

extern "C" int slode_heads_fwd(int64_t B, int T, int S, int O, int NQ, const float* sol, int64_t sol_stride_t,
                               int64_t sol_stride_b, const float* W, float* mu, void* stream_) {
  int rc = heads::check("slode_heads_fwd", B, T, S, O, NQ);
  if (rc) return rc;
  if (!W || (B > 0 && (!sol || !mu))) {
    set_error("slode_heads_fwd: null pointer");
    return SLODE_EINVAL;
  }
  if (B == 0) return SLODE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nqo = NQ * O;
  if (T % 4 == 0 && sol_stride_t == S && (sol_stride_b % 4) == 0 && nqo <= 12 && heads::aligned16(sol) &&
      heads::aligned16(mu)) {
    const int64_t blocks = (B * (int64_t)(T / 4) + heads::kBlock - 1) / heads::kBlock;
    const int sms = heads::device_sms();
#define GOV(SS, QM)                                                                                                \
  do {                                                                                                             \
    auto kern = heads::heads_fwd_vec_kernel<SS, QM>;                                                               \
    const int grid = (int)std::min<int64_t>(blocks, (int64_t)sms * heads::blocks_per_sm(kern));                    \
    kern<<<grid, heads::kBlock, 0, stream>>>(B, T, O, NQ, sol, sol_stride_b, W, mu);                               \
  } while (0)
#define GOVQ(SS)                                                                                                   \
  do {                                                                                                             \
    if (nqo <= 3) GOV(SS, 3); else if (nqo <= 9) GOV(SS, 9); else GOV(SS, 12);                                     \
  } while (0)
    if (S == 4) GOVQ(4); else if (S == 5) GOVQ(5); else GOVQ(8);
#undef GOVQ
#undef GOV
  } else {
    const int64_t blocks = (B * (int64_t)T + heads::kBlock - 1) / heads::kBlock;
    const int grid = (int)std::min<int64_t>(blocks, (int64_t)heads::device_sms() * 32);
#define GO(SS) heads::heads_fwd_kernel<SS><<<grid, heads::kBlock, 0, stream>>>(B, T, O, NQ, sol, sol_stride_t, sol_stride_b, W, mu)
    if (S == 4) GO(4); else if (S == 5) GO(5); else GO(8);
#undef GO
  }
  SLODE_CUDA_TRY(cudaGetLastError());
  g_fwd_launches = 1;
  return SLODE_OK;
}

extern "C" int slode_heads_bwd(int64_t B, int T, int S, int O, int NQ, const float* sol, int64_t sol_stride_t,
                               int64_t sol_stride_b, const float* W, const float* grad_mu, float* grad_sol,
                               int64_t gsol_stride_t, int64_t gsol_stride_b, float* grad_W, void* stream_) {
  if (B > 0 && !grad_mu) {
    set_error("slode_heads_bwd: null pointer");
    return SLODE_EINVAL;
  }
  const int64_t hs = B * (int64_t)(O > 0 ? O : 0) * T;
  return slode_heads_bwd_split(B, T, S, O, NQ, sol, sol_stride_t, sol_stride_b, W, grad_mu,
                               NQ > 1 ? grad_mu + hs : nullptr, NQ > 2 ? grad_mu + 2 * hs : nullptr, grad_sol,
                               gsol_stride_t, gsol_stride_b, grad_W, stream_);
}

extern "C" int slode_heads_bwd_split(int64_t B, int T, int S, int O, int NQ, const float* sol, int64_t sol_stride_t,
                                     int64_t sol_stride_b, const float* W, const float* g0, const float* g1,
                                     const float* g2, float* grad_sol, int64_t gsol_stride_t, int64_t gsol_stride_b,
                                     float* grad_W, void* stream_) {
  int rc = heads::check("slode_heads_bwd", B, T, S, O, NQ);
  if (rc) return rc;
  if (!W || !grad_W || (B > 0 && (!sol || !grad_sol))) {
    set_error("slode_heads_bwd: null pointer");
    return SLODE_EINVAL;
  }
  if (NQ < 2) g1 = nullptr;
  if (NQ < 3) g2 = nullptr;
  if (B == 0) return SLODE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int sms = heads::device_sms(), nqo = NQ * O;
  if (T % 4 == 0 && sol_stride_t == S && gsol_stride_t == S && (sol_stride_b % 4) == 0 && (gsol_stride_b % 4) == 0 &&
      nqo <= 12 && heads::aligned16(sol) && heads::aligned16(grad_sol) && heads::aligned16(g0) && heads::aligned16(g1) &&
      heads::aligned16(g2)) {
    const int64_t blocks = (B * (int64_t)(T / 4) + heads::kBlock - 1) / heads::kBlock;
#define GOV(SS, QM)                                                                                                \
  do {                                                                                                             \
    auto kern = heads::heads_bwd_vec_kernel<SS, QM>;                                                               \
    const int grid = (int)std::min<int64_t>(blocks, (int64_t)sms * heads::blocks_per_sm(kern));                    \
    kern<<<grid, heads::kBlock, 0, stream>>>(B, T, O, NQ, sol, sol_stride_b, W, g0, g1, g2, grad_sol,              \
                                             gsol_stride_b, grad_W);                                               \
  } while (0)
#define GOVQ(SS)                                                                                                   \
  do {                                                                                                             \
    if (nqo <= 3) GOV(SS, 3); else if (nqo <= 9) GOV(SS, 9); else GOV(SS, 12);                                     \
  } while (0)
    if (S == 4) GOVQ(4); else if (S == 5) GOVQ(5); else GOVQ(8);
#undef GOVQ
#undef GOV
  } else {
  const int64_t blocks = (B * (int64_t)T + heads::kBlock - 1) / heads::kBlock;
    const int sms = heads::device_sms(), nqo = NQ * O;
  #define GO(SS, QM)                                                                                                 \
    do {                                                                                                             \
      auto kern = heads::heads_bwd_kernel<SS, QM>;                                                                   \
      const int grid = (int)std::min<int64_t>(blocks, (int64_t)sms * heads::blocks_per_sm(kern));                    \
      kern<<<grid, heads::kBlock, 0, stream>>>(B, T, O, NQ, sol, sol_stride_t, sol_stride_b, W, g0, g1, g2, grad_sol, \
                                               gsol_stride_t, gsol_stride_b, grad_W);                                \
    } while (0)
  #define GOQ(SS)                                                                                                    \
    do {                                                                                                             \
      if (nqo <= 3) GO(SS, 3); else if (nqo <= 9) GO(SS, 9); else GO(SS, 24);                                        \
    } while (0)
    if (S == 4) GOQ(4); else if (S == 5) GOQ(5); else GOQ(8);
  #undef GOQ
  #undef GO
  }
  SLODE_CUDA_TRY(cudaGetLastError());
  g_bwd_launches = 1;
  return SLODE_OK;
}
