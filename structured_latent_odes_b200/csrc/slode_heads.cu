// Decoder quantile / mean heads on the latent trajectories (sm_100a).
//
// Reference: Decoder.forward / GaussianDecoder.forward (models/decoders.py:42-54, 84-91):
//     mu_q = Linear_q(solution).permute(0, 2, 1)        Linear_q: ode_state_dim -> obs_dim, no bias
// for q in {q50, q75, q25} (Decoder) or the single mean head (GaussianDecoder); solution is (B,T,S), mu_q (B,O,T).
// In the reference this is one tiny-K matmul + a permuted view per head (and, backwards, two more matmuls per
// head); here one pass over the trajectories produces all heads in their final (B,O,T) layout, and one pass
// backwards produces dL/dsolution and the head-weight gradients.  HBM-bound: 4*S bytes read and 4*NQ*O bytes
// written per (trajectory, time).
//
// One thread = one (trajectory, time) point, time fastest: with (B,T,S)-contiguous storage of the solution
// (layout="bts") both the S-float read and the per-(q,o) writes along T are coalesced.
#include <algorithm>

#include "slode_common.cuh"

namespace slode {
namespace heads {

constexpr int kBlock = 256;
constexpr int kMaxW = 3 * 8 * 8;  // NQ * O * S

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

template <int K>
__device__ __forceinline__ float warp_scatter(float (&v)[K], int lane, int& slot) {
  int base = 0, n = K;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    if (n > 1) {
      const int hn = n / 2;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int k = 0; k < K / 2; ++k) {
        if (k < hn) {
          const float mine = upper ? v[k + hn] : v[k];
          const float give = upper ? v[k] : v[k + hn];
          v[k] = mine + __shfl_xor_sync(0xffffffffu, give, off);
        }
      }
      if (upper) base += hn;
      n = hn;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
  slot = base;
  return v[0];
}

template <int S>
__global__ void __launch_bounds__(kBlock)
heads_bwd_kernel(int64_t B, int T, int O, int NQ, const float* __restrict__ sol, int64_t st, int64_t sb,
                 const float* __restrict__ W, const float* __restrict__ gmu, float* __restrict__ gsol, int64_t gst,
                 int64_t gsb, float* __restrict__ gW) {
  constexpr int KS = 8;  // S padded to a power of two for the warp reduction
  static_assert(S <= KS, "state dimension");
  __shared__ float sW[kMaxW];
  __shared__ float sG[kMaxW];
  for (int i = threadIdx.x; i < NQ * O * S; i += kBlock) {
    sW[i] = W[i];
    sG[i] = 0.0f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t n = B * (int64_t)T;
  // all lanes of a warp run the same number of iterations (the warp reductions need every lane)
  const int64_t n_pad = (n + 31) / 32 * 32;
  for (int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x; idx < n_pad; idx += (int64_t)gridDim.x * kBlock) {
    const bool ok = idx < n;
    const int64_t ii = ok ? idx : n - 1;
    const int64_t b = ii / T;
    const int t = (int)(ii - b * T);
    const float* x = sol + b * sb + (int64_t)t * st;
    float xs[S], gs[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      xs[s] = ok ? __ldg(x + s) : 0.0f;
      gs[s] = 0.0f;
    }
    for (int q = 0; q < NQ; ++q) {
      for (int o = 0; o < O; ++o) {
        const float g = ok ? __ldg(gmu + (((int64_t)q * B + b) * O + o) * T + t) : 0.0f;
        const float* w = sW + (q * O + o) * S;
        float v[KS];
#pragma unroll
        for (int s = 0; s < KS; ++s) v[s] = 0.0f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          gs[s] = fmaf(g, w[s], gs[s]);
          v[s] = g * xs[s];
        }
        int slot;
        const float tot = warp_scatter<KS>(v, lane, slot);
        if ((lane & (32 / KS - 1)) == 0 && slot < S) atomicAdd(&sG[(q * O + o) * S + slot], tot);
      }
    }
    if (ok) {
      float* gx = gsol + b * gsb + (int64_t)t * gst;
#pragma unroll
      for (int s = 0; s < S; ++s) gx[s] = gs[s];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * O * S; i += kBlock) atomicAdd(gW + i, sG[i]);
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
  const int64_t blocks = (B * (int64_t)T + heads::kBlock - 1) / heads::kBlock;
  const int grid = (int)std::min<int64_t>(blocks, (int64_t)heads::device_sms() * 32);
#define GO(SS) heads::heads_fwd_kernel<SS><<<grid, heads::kBlock, 0, stream>>>(B, T, O, NQ, sol, sol_stride_t, sol_stride_b, W, mu)
  if (S == 4) GO(4); else if (S == 5) GO(5); else GO(8);
#undef GO
  SLODE_CUDA_TRY(cudaGetLastError());
  g_fwd_launches = 1;
  return SLODE_OK;
}

extern "C" int slode_heads_bwd(int64_t B, int T, int S, int O, int NQ, const float* sol, int64_t sol_stride_t,
                               int64_t sol_stride_b, const float* W, const float* grad_mu, float* grad_sol,
                               int64_t gsol_stride_t, int64_t gsol_stride_b, float* grad_W, void* stream_) {
  int rc = heads::check("slode_heads_bwd", B, T, S, O, NQ);
  if (rc) return rc;
  if (!W || !grad_W || (B > 0 && (!sol || !grad_mu || !grad_sol))) {
    set_error("slode_heads_bwd: null pointer");
    return SLODE_EINVAL;
  }
  if (B == 0) return SLODE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t blocks = (B * (int64_t)T + heads::kBlock - 1) / heads::kBlock;
  const int grid = (int)std::min<int64_t>(blocks, (int64_t)heads::device_sms() * 8);
#define GO(SS)                                                                                                      \
  heads::heads_bwd_kernel<SS><<<grid, heads::kBlock, 0, stream>>>(B, T, O, NQ, sol, sol_stride_t, sol_stride_b, W, \
                                                                  grad_mu, grad_sol, gsol_stride_t, gsol_stride_b, grad_W)
  if (S == 4) GO(4); else if (S == 5) GO(5); else GO(8);
#undef GO
  SLODE_CUDA_TRY(cudaGetLastError());
  g_bwd_launches = 1;
  return SLODE_OK;
}
