// Fixed-grid solve of the SLODE blackbox latent ODE and its reverse sweep, hand-written for sm_100a.
//
// Right-hand side (reference: Dynamics.forward, models/blackbox_ode.py:97-109):
//     h_j(t)  = relu(w1t_j * t + c_j)                       c = z W1[:,1:]^T + b1  (per trajectory)
//     A_k(t)  = sigmoid(bg_k + sum_j Wg_kj h_j(t))          "growth"
//     D_k(t)  = sigmoid(bd_k + sum_j Wd_kj h_j(t))          "degradation"
//     f(t,x)  = A(t) - D(t) * x                              (affine in the state, elementwise)
//
// Mapping: one thread = TWO trajectories for the whole time loop, packed in the two halves of a 64-bit
// register pair; every arithmetic instruction of the kernel is a packed fp32x2 op (FFMA2 / FADD2 / FMUL2) over
// that pair.  Weights are warp-uniform scalars streamed from constant memory into uniform registers (one
// LDCU.128 = four weights) and enter the FMA as the broadcast operand:
//     acc_o(traj0,traj1) += W_jo (UR, .F32 broadcast) * h_j(traj0,traj1)
// Measured on B200 (profiles/r01/fp32_pipes_microbench.jsonl): this form sustains 127 of the 128 FMA/clk/SM
// with one LDCU.128 per four FFMA2, where a 3-register FFMA reaches 84.  Per trajectory and MLP evaluation
// (H=25, S=5): 137.5 FFMA2 + 39 LDCU.128 + 25 FMNMX + 20 MUFU + ~10 others  ->  FMA-pipe bound.
//
// rk4 (3/8 rule) re-uses the evaluation at t1 as the next step's evaluation at t0 (same float).
//
// Backward: reverse sweep over the stored grid states sol[i]; stages are recomputed.  Because the hidden layer
// sees only (t, z), the cotangents delta_o(e) of the head pre-activations at the evaluation times t_e determine
// every hidden-layer gradient through prefix sums
//     P_o = sum_e delta_o(e),   Q_o = sum_e delta_o(e) t_e
// taken over the evaluations where unit j is active.  The sweep keeps running P,Q and, whenever a unit's relu
// gate flips between consecutive evaluations (once per unit for monotone t; the summation by parts is valid for
// any number of flips), adds +-snapshot contributions
//     dc_j   += s * sum_o W_oj P_o              (per trajectory -> grad_c)
//     dw1t_j += s * sum_o W_oj Q_o              (block accumulator)
//     dW_oj  += s * (w1t_j Q_o + c_j P_o)       (block accumulator, = sum_e delta_o h_j)
// with s=+1 when the unit turns off, -1 when it turns on, and +1 for every unit still active when the sweep
// ends.  This replaces the two dense 2S*H products per evaluation of a textbook backward by O(S) work.
#pragma once

#include <algorithm>
#include <type_traits>

#include "slode_common.cuh"
#include "slode_mlp_api.h"

#ifndef SLODE_PACK_SYM
#error "define SLODE_PACK_SYM (the per-translation-unit constant symbol) before including slode_mlp_kernels.cuh"
#endif
// build-time switches (kernel A/B measurements build variants of the library with -D...)
#ifndef SLODE_FWD_CSMEM
#define SLODE_FWD_CSMEM 0
#endif
#ifndef SLODE_PL
#define SLODE_PL 2  // fixed-grid kernels: 2 = piecewise-linear heads with the sorted walk over relu crossings,
                   // 1 = piecewise-linear heads with all H gates tested per evaluation, 0 = dense products
#endif
#ifndef SLODE_FWD_MINB
#define SLODE_FWD_MINB 3
#endif
#ifndef SLODE_BWD_MINB
#define SLODE_BWD_MINB 2
#endif
#define SLODE_STR2(x) #x
#define SLODE_STR(x) SLODE_STR2(x)

namespace slode {

// ---------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic: lo = first trajectory of the thread, hi = second
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f2;

__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 bc(float v) { return pk(v, v); }
__device__ __forceinline__ void unpk(f2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
template <int HALF> __device__ __forceinline__ float half_of(f2 v) {
  float a, b;
  unpk(v, a, b);
  return HALF ? b : a;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 neg2(f2 a) { return a ^ 0x8000000080000000ull; }

// S per-state values of the two trajectories
template <int S>
struct Vec {
  f2 v[S];
};
#define SLODE_FOR_S for (int s = 0; s < S; ++s)

template <int S> __device__ __forceinline__ Vec<S> vadd(const Vec<S>& a, const Vec<S>& b) {
  Vec<S> r;
#pragma unroll
  SLODE_FOR_S r.v[s] = add2(a.v[s], b.v[s]);
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vsub(const Vec<S>& a, const Vec<S>& b) {
  Vec<S> r;
#pragma unroll
  SLODE_FOR_S r.v[s] = sub2(a.v[s], b.v[s]);
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vmul(const Vec<S>& a, const Vec<S>& b) {
  Vec<S> r;
#pragma unroll
  SLODE_FOR_S r.v[s] = mul2(a.v[s], b.v[s]);
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vscale(const Vec<S>& a, float c) {
  Vec<S> r;
  const f2 cc = bc(c);
#pragma unroll
  SLODE_FOR_S r.v[s] = mul2(a.v[s], cc);
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vscale2(const Vec<S>& a, f2 cc) {
  Vec<S> r;
#pragma unroll
  SLODE_FOR_S r.v[s] = mul2(a.v[s], cc);
  return r;
}
// c*a + b  (scalar c)
template <int S> __device__ __forceinline__ Vec<S> vaxpy(float c, const Vec<S>& a, const Vec<S>& b) {
  Vec<S> r;
  const f2 cc = bc(c);
#pragma unroll
  SLODE_FOR_S r.v[s] = fma2(cc, a.v[s], b.v[s]);
  return r;
}
// a*b + c
template <int S> __device__ __forceinline__ Vec<S> vfma(const Vec<S>& a, const Vec<S>& b, const Vec<S>& c) {
  Vec<S> r;
#pragma unroll
  SLODE_FOR_S r.v[s] = fma2(a.v[s], b.v[s], c.v[s]);
  return r;
}
// c - a*b
template <int S> __device__ __forceinline__ Vec<S> vnfma(const Vec<S>& a, const Vec<S>& b, const Vec<S>& c) {
  Vec<S> r;
#pragma unroll
  SLODE_FOR_S r.v[s] = fma2(neg2(a.v[s]), b.v[s], c.v[s]);
  return r;
}
// -(a*b)
template <int S> __device__ __forceinline__ Vec<S> vnmul(const Vec<S>& a, const Vec<S>& b) {
  Vec<S> r;
#pragma unroll
  SLODE_FOR_S r.v[s] = mul2(neg2(a.v[s]), b.v[s]);
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vload2(const float* p0, const float* p1) {
  Vec<S> r;
#pragma unroll
  SLODE_FOR_S r.v[s] = pk(__ldg(p0 + s), __ldg(p1 + s));
  return r;
}
// a load the compiler may not sink to its use
__device__ __forceinline__ float ld_early(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// pull the S floats at p (20-32 bytes, may straddle two sectors) towards L1 for a later vload2
template <int S> __device__ __forceinline__ void vprefetch(const float* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p + S - 1));
}
template <int S> __device__ __forceinline__ void vstore2(float* p0, bool ok0, float* p1, bool ok1, const Vec<S>& a) {
#pragma unroll
  SLODE_FOR_S {
    float lo, hi;
    unpk(a.v[s], lo, hi);
    if (ok0) p0[s] = lo;
    if (ok1) p1[s] = hi;
  }
}

// ---------------------------------------------------------------------------------------------
// packed weights in constant memory
// ---------------------------------------------------------------------------------------------
constexpr int kPackMax = 8192;   // floats (32 KB of the 64 KB constant bank), one buffer per compiled shape
constexpr int kSlots = 2;        // copies of the packed weights, one per evaluation site of a kernel
constexpr int kBlock = 128;      // threads per block = 256 trajectories per tile
}  // namespace slode

// C linkage: the loads below name the symbol from inline PTX.  One symbol per translation unit (= per compiled
// (H,S) shape), named by SLODE_PACK_SYM.
extern "C" {
__constant__ __align__(16) float SLODE_PACK_SYM[slode::kPackMax];
}
namespace slode {

// Weight loads are `asm volatile` with a STATIC address (symbol + immediate): NVVM may not move or merge them
// and ptxas keeps them inside the evaluation as 16-byte uniform-register loads  LDCU.128 UR, c[3][imm].
// (Left alone, both compilers hoist the loop-invariant loads out of the time loop into registers and spill; a
// register-offset address c[3][UR+imm] makes ptxas split every 16-byte load into two LDCU.64.)  ptxas would
// still merge loads of the SAME address issued by different evaluations of one time step into ordinary
// registers, so every evaluation site of a kernel reads its own copy ("slot") of the packed weights.
template <int OFF_FLOATS>
__device__ __forceinline__ void ldc4(float* v) {
  static_assert(OFF_FLOATS % 4 == 0, "16-byte aligned");
  asm volatile("ld.const.v4.f32 {%0, %1, %2, %3}, [" SLODE_STR(SLODE_PACK_SYM) "+%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
               : "n"(OFF_FLOATS * 4));
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// One slot: [ head biases (2S, padded to 4) | H unit records ], record j = [ w1t_j | W_j,0 .. W_j,2S-1 ] padded
// to a multiple of 4 floats.  Head outputs o < S are growth, o >= S degradation; biases and head weights are
// pre-scaled by -log2(e) so that sigmoid(u) = rcp(1 + ex2(v)).
template <int H, int S>
struct Pack {
  static constexpr int K2 = 2 * S;
  static constexpr int KP = (K2 + 3) / 4 * 4;
  static constexpr int UNIT = (1 + K2 + 3) / 4 * 4;
  static constexpr int N = KP + H * UNIT;
};

template <int H, int S>
__global__ void pack_kernel(const float* __restrict__ w1t, const float* __restrict__ Wg, const float* __restrict__ bg,
                            const float* __restrict__ Wd, const float* __restrict__ bd, float* __restrict__ out) {
  using P = Pack<H, S>;
  for (int i = threadIdx.x; i < P::N; i += blockDim.x) {
    float v = 0.0f;
    if (i < P::KP) {
      if (i < P::K2) v = kNegLog2e * ((i < S) ? bg[i] : bd[i - S]);
    } else {
      const int j = (i - P::KP) / P::UNIT, r = (i - P::KP) % P::UNIT;
      if (r == 0) {
        v = w1t[j];
      } else if (r <= P::K2) {
        const int o = r - 1;
        v = kNegLog2e * ((o < S) ? Wg[o * H + j] : Wd[(o - S) * H + j]);
      }
    }
    for (int slot = 0; slot < kSlots; ++slot) out[slot * P::N + i] = v;
  }
}

template <int H, int S>
int upload_pack(const PackSrc& w, float* staging, cudaStream_t stream) {
  using P = Pack<H, S>;
  static_assert(kSlots * P::N <= kPackMax, "packed weights exceed the constant buffer");
  pack_kernel<H, S><<<1, 256, 0, stream>>>(w.w1t, w.Wg, w.bg, w.Wd, w.bd, staging);
  SLODE_CUDA_TRY(cudaGetLastError());
  SLODE_CUDA_TRY(cudaMemcpyToSymbolAsync(SLODE_PACK_SYM, staging, sizeof(float) * kSlots * P::N, 0,
                                         cudaMemcpyDeviceToDevice, stream));
  return SLODE_OK;
}

template <int H>
struct Gate {
  static constexpr int NW = (H + 31) / 32;
  uint32_t w[2][NW];  // [half][word]; unit j sits at bit (n_w - 1 - (j - 32 word)) of its word
};

// NE RHS evaluations (times t[0..NE)) for both trajectories in one pass over the weights.  The MLP sees only
// (t, z), never the state, so all evaluations of one solver step can be taken together: every weight streamed
// from constant memory then feeds NE FFMA2.  This matters because LDCU.128 sustains only one load per ~8 cycles
// per SM sub-partition (profiles/microbench/ldcu_rate.cu): at one FFMA2 (2 cycles) per weight the weight stream
// and the FMA pipe are exactly balanced and neither can be saturated; at NE >= 2 the kernel is FMA-bound.
// cj(j) returns the pair (c_j of traj0, c_j of traj1).
// Outputs: A = sigmoid(growth heads) and ND = MINUS sigmoid(degradation heads).  The sign costs nothing
// (-D = rcp(-1 - e) instead of rcp(1 + e)) and removes every negation downstream: f = A + ND*x, the stage
// adjoint dL/dY = gk*ND, D^2 - D = ND^2 + ND.  (fma.rn.f32x2 has no operand-negate form; a packed negation is
// two LOP3 per pair.)
template <int H, int S, int NE, bool MASK, int SLOT, class CLoad>
__device__ __forceinline__ void mlp_eval(const float (&t)[NE], CLoad cj, Vec<S> (&A)[NE], Vec<S> (&ND)[NE],
                                         Gate<H> (&gate)[NE]) {
  using P = Pack<H, S>;
  constexpr int K2 = 2 * S;
  constexpr int BASE = SLOT * P::N;
  constexpr int NW = Gate<H>::NW;
  f2 acc[NE][P::KP];
  static_for<0, P::KP / 4>([&](auto I) {
    constexpr int q = decltype(I)::value;
    float b[4];
    ldc4<BASE + 4 * q>(b);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[e][4 * q + r] = bc(b[r]);
    }
  });
  uint32_t neg[NE][2][NW];
  if (MASK) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
#pragma unroll
      for (int w = 0; w < NW; ++w) neg[e][0][w] = neg[e][1][w] = 0u;
    }
  }
  f2 tt[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) tt[e] = bc(t[e]);
  static_for<0, H>([&](auto J) {
    constexpr int j = decltype(J)::value;
    float r[P::UNIT];
    static_for<0, P::UNIT / 4>([&](auto Q) {
      constexpr int q = decltype(Q)::value;
      ldc4<BASE + P::KP + j * P::UNIT + 4 * q>(r + 4 * q);
    });
    const f2 c = cj(j);
    const f2 w1 = bc(r[0]);
    f2 h[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      float p0, p1;
      unpk(fma2(w1, tt[e], c), p0, p1);
      if (MASK) {
        neg[e][0][j / 32] = __funnelshift_l(__float_as_uint(p0), neg[e][0][j / 32], 1);
        neg[e][1][j / 32] = __funnelshift_l(__float_as_uint(p1), neg[e][1][j / 32], 1);
      }
      h[e] = pk(fmaxf(p0, 0.0f), fmaxf(p1, 0.0f));
    }
#pragma unroll
    for (int o = 0; o < K2; ++o) {
      const f2 w = bc(r[1 + o]);
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[e][o] = fma2(h[e], w, acc[e][o]);
    }
  });
  if (MASK) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const int nw = (w == NW - 1) ? (H - 32 * w) : 32;
        const uint32_t low = (nw == 32) ? 0xffffffffu : ((1u << nw) - 1u);
        gate[e].w[0][w] = (~neg[e][0][w]) & low;
        gate[e].w[1][w] = (~neg[e][1][w]) & low;
      }
    }
  }
  const f2 one = bc(1.0f), minus_one = bc(-1.0f);
#pragma unroll
  for (int e = 0; e < NE; ++e) {
#pragma unroll
    for (int o = 0; o < K2; ++o) {
      float v0, v1;
      unpk(acc[e][o], v0, v1);
      const f2 ex = pk(ex2_approx(v0), ex2_approx(v1));
      unpk(o < S ? add2(ex, one) : sub2(minus_one, ex), v0, v1);
      const f2 sg = pk(rcp_approx(v0), rcp_approx(v1));
      if (o < S) A[e].v[o] = sg; else ND[e].v[o - S] = sg;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// evaluation checkpoints: the (A, -D) of every MLP evaluation of a solve, in the threads' own pair layout
//   element (tile, warp w, evaluation e, head output k, lane) = one f2 at
//       ckpt[(((tile * 4 + w) * n_evals + e) * 2S + k) * 32 + lane]
// Warp-major: the 64 trajectories of a warp own one contiguous stream (n_evals x 2S x 256 B) that the forward
// fills front to back and the reverse sweep consumes back to front -- one interval's evaluations are ONE contiguous
// run (7.5 KB for rk4 at S = 5), which the sweep fetches with a single bulk copy per warp and interval.
// Evaluation order: euler e = i (time t_i); midpoint e = 2i (t_i), 2i+1 (t_i + dt/2); rk4 e = 0 (t_0) and
// 3i+1, 3i+2, 3i+3 = (t_i + dt/3, t_i + 2dt/3, t_{i+1}).  120 B per trajectory and rk4 step at S = 5: writing
// them costs the forward ~5 clk/SM per trajectory-step of HBM time, re-computing them costs the reverse sweep
// >= 6.4 clk/SM of FMA time at PEAK (13 measured) -- on 180 GB of HBM3e the checkpoint is the better trade.
// ---------------------------------------------------------------------------------------------
template <int METHOD>
__host__ __device__ constexpr int64_t ckpt_evals(int T) {
  return METHOD == SLODE_METHOD_RK4 ? 3 * (int64_t)(T - 1) + 1 : (METHOD == SLODE_METHOD_MIDPOINT ? 2 * (int64_t)(T - 1) : T - 1);
}
// start of the calling warp's checkpoint stream
template <int S>
__device__ __forceinline__ int64_t ckpt_warp_offset(int64_t tile, int64_t n_evals) {
  return ((tile * (kBlock / 32) + (threadIdx.x >> 5)) * n_evals) * (2 * S) * 32;
}
template <int S>
__device__ __forceinline__ void ckpt_store(f2* __restrict__ warp_base, int64_t e, const Vec<S>& A, const Vec<S>& ND) {
  f2* p = warp_base + e * (2 * S) * 32 + (threadIdx.x & 31);
#pragma unroll
  SLODE_FOR_S {
    p[s * 32] = A.v[s];
    p[(S + s) * 32] = ND.v[s];
  }
}
template <int S>
__device__ __forceinline__ void ckpt_load(const f2* __restrict__ warp_base, int64_t e, Vec<S>& A, Vec<S>& ND) {
  const f2* p = warp_base + e * (2 * S) * 32 + (threadIdx.x & 31);
#pragma unroll
  SLODE_FOR_S {
    A.v[s] = __ldg(p + s * 32);
    ND.v[s] = __ldg(p + (S + s) * 32);
  }
}

// pull one evaluation's checkpoint (2S pairs, one cache line per (output, warp)) towards L1 an iteration ahead
template <int S>
__device__ __forceinline__ void ckpt_prefetch(const f2* __restrict__ warp_base, int64_t e) {
  const f2* p = warp_base + e * (2 * S) * 32 + (threadIdx.x & 31);
#pragma unroll
  for (int k = 0; k < 2 * S; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + k * 32));
}

// relu gates of NE evaluation times from the hidden layer alone (the reverse sweep with checkpoints needs the
// gates for the prefix-sum bookkeeping but not the heads)
template <int H, int NE, class CLoad>
__device__ __forceinline__ void gates_only(const float* __restrict__ w1t_smem, const float (&t)[NE], CLoad cj,
                                           Gate<H> (&gate)[NE], int w1t_stride = 1) {
  constexpr int NW = Gate<H>::NW;
  uint32_t neg[NE][2][NW];
#pragma unroll
  for (int e = 0; e < NE; ++e) {
#pragma unroll
    for (int w = 0; w < NW; ++w) neg[e][0][w] = neg[e][1][w] = 0u;
  }
  f2 tt[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) tt[e] = bc(t[e]);
#pragma unroll
  for (int j = 0; j < H; ++j) {
    const f2 c = cj(j);
    const f2 w1 = bc(w1t_smem[j * w1t_stride]);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      float p0, p1;
      unpk(fma2(w1, tt[e], c), p0, p1);
      neg[e][0][j / 32] = __funnelshift_l(__float_as_uint(p0), neg[e][0][j / 32], 1);
      neg[e][1][j / 32] = __funnelshift_l(__float_as_uint(p1), neg[e][1][j / 32], 1);
    }
  }
#pragma unroll
  for (int e = 0; e < NE; ++e) {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int nw = (w == NW - 1) ? (H - 32 * w) : 32;
      const uint32_t low = (nw == 32) ? 0xffffffffu : ((1u << nw) - 1u);
      gate[e].w[0][w] = (~neg[e][0][w]) & low;
      gate[e].w[1][w] = (~neg[e][1][w]) & low;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-warp bulk-copy (TMA, 1-D) staging of the checkpoints: lane 0 streams the evaluations of the interval two
// ahead into the warp's own shared-memory stage while the warp works; completion arrives on the warp's mbarrier.
// No block-wide barrier is involved (a first version that shared the stages across the block lost more to the
// barrier than the copies hid).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  int spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1 << 26)) __trap();  // a lost copy must not hang the GPU
  }
}
template <int S, int NE>
__device__ __forceinline__ void stage_read(const f2* __restrict__ st, Vec<S> (&A)[NE], Vec<S> (&ND)[NE]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int e = 0; e < NE; ++e) {
#pragma unroll
    SLODE_FOR_S {
      A[e].v[s] = st[(e * 2 * S + s) * 32 + lane];
      ND[e].v[s] = st[(e * 2 * S + S + s) * 32 + lane];
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Piecewise-linear evaluation of the heads.  The hidden layer sees only (t, z): along one trajectory the head
// pre-activations are
//     o_k(t) = b_k + sum_j W_kj relu(w1t_j t + c_j) = alpha_k t + beta_k,
//     alpha_k = sum_{j active} W_kj w1t_j,   beta_k = b_k + sum_{j active} W_kj c_j,
// with coefficients that change only when a relu gate flips -- at most once per unit over a monotone sweep of t.
// The evaluator keeps (alpha, beta) of the thread's two trajectories in registers; one evaluation is the H gate
// tests (one FFMA2 + two funnel shifts per unit), 2S FFMA2 and the 2S sigmoids instead of the dense
// H x (2S + 1) FFMA2, and every flip costs one rank-one update read from shared memory (lanes flip different
// units at different times: the update loop is divergent but short, <= H trips per trajectory and solve).
// tb is a shared-memory copy of one slot of the packed weights (already scaled by -log2 e), cf the block's
// c table [H][kBlock] of f2 viewed as floats.
// ---------------------------------------------------------------------------------------------
template <int H, int S>
struct PlEval {
  using P = Pack<H, S>;
  static constexpr int K2 = 2 * S;
  static constexpr int NW = Gate<H>::NW;
  f2 al[K2], be[K2];
  uint32_t cur[2][NW];

  __device__ __forceinline__ void init(const float* __restrict__ tb, const float* __restrict__ cf, const Gate<H>& g) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int k = 0; k < K2; ++k) {
      al[k] = 0ull;
      be[k] = bc(tb[k]);
    }
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int nw = (w == NW - 1) ? (H - 32 * w) : 32;
      const uint32_t g0 = g.w[0][w], g1 = g.w[1][w];
#pragma unroll 1
      for (int q = 0; q < nw; ++q) {
        const int j = 32 * w + (nw - 1 - q);
        const float* r = tb + P::KP + j * P::UNIT;
        const f2 m = pk(((g0 >> q) & 1u) ? 1.0f : 0.0f, ((g1 >> q) & 1u) ? 1.0f : 0.0f);
        const f2 u = mul2(bc(r[0]), m);
        const f2 v = mul2(reinterpret_cast<const f2*>(cf)[j * kBlock + tid], m);
#pragma unroll
        for (int k = 0; k < K2; ++k) {
          const f2 wk = bc(r[1 + k]);
          al[k] = fma2(wk, u, al[k]);
          be[k] = fma2(wk, v, be[k]);
        }
      }
      cur[0][w] = g0;
      cur[1][w] = g1;
    }
  }

  // bring (alpha, beta) to the gate pattern g: one trip per flipped unit, both trajectories served per trip
  __device__ __forceinline__ void update(const float* __restrict__ tb, const float* __restrict__ cf, const Gate<H>& g) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int nw = (w == NW - 1) ? (H - 32 * w) : 32;
      const uint32_t g0 = g.w[0][w], g1 = g.w[1][w];
      uint32_t d0 = g0 ^ cur[0][w], d1 = g1 ^ cur[1][w];
      while (d0 | d1) {
        const int q0 = d0 ? __ffs(d0) - 1 : 0, q1 = d1 ? __ffs(d1) - 1 : 0;
        const float s0 = d0 ? (((g0 >> q0) & 1u) ? 1.0f : -1.0f) : 0.0f;
        const float s1 = d1 ? (((g1 >> q1) & 1u) ? 1.0f : -1.0f) : 0.0f;
        d0 &= d0 - 1;
        d1 &= d1 - 1;
        const int j0 = 32 * w + (nw - 1 - q0), j1 = 32 * w + (nw - 1 - q1);
        const float* r0 = tb + P::KP + j0 * P::UNIT;
        const float* r1 = tb + P::KP + j1 * P::UNIT;
        float a0[P::UNIT], a1[P::UNIT];
#pragma unroll
        for (int k = 0; k < P::UNIT / 4; ++k) {
          const float4 x0 = reinterpret_cast<const float4*>(r0)[k];
          const float4 x1 = reinterpret_cast<const float4*>(r1)[k];
          a0[4 * k] = x0.x; a0[4 * k + 1] = x0.y; a0[4 * k + 2] = x0.z; a0[4 * k + 3] = x0.w;
          a1[4 * k] = x1.x; a1[4 * k + 1] = x1.y; a1[4 * k + 2] = x1.z; a1[4 * k + 3] = x1.w;
        }
        const float u0 = s0 * a0[0], u1 = s1 * a1[0];
        const float v0 = s0 * cf[(j0 * kBlock + tid) * 2], v1 = s1 * cf[(j1 * kBlock + tid) * 2 + 1];
#pragma unroll
        for (int k = 0; k < K2; ++k) {
          float lo, hi;
          unpk(al[k], lo, hi);
          al[k] = pk(fmaf(a0[1 + k], u0, lo), fmaf(a1[1 + k], u1, hi));
          unpk(be[k], lo, hi);
          be[k] = pk(fmaf(a0[1 + k], v0, lo), fmaf(a1[1 + k], v1, hi));
        }
      }
      cur[0][w] = g0;
      cur[1][w] = g1;
    }
  }

  // ---- sorted walk (SLODE_PL == 2): instead of testing all H gates at every evaluation, the trajectory's
  // flips are visited in the order the sweep meets them.  p_j(t) = w1t_j t + c_j crosses zero at t*_j = -c_j/w1t_j;
  // in the sweep coordinate u = dirsign (t - t_start) >= 0 the pending crossings are sorted once per trajectory
  // (bitonic network in registers, keys = float bits of u*_j with the unit index in the low mantissa bits, biased
  // early) and kept in shared memory.  An evaluation compares its u with the next key (one FSETP per trajectory);
  // a due candidate is confirmed with the exact fp32 gate test the dense evaluation would make, so the gate
  // patterns are those of gates_only whenever the key is not late, and a key cannot be late by more than the
  // rounding of t*_j.
  static constexpr int IB = H <= 32 ? 5 : (H <= 64 ? 6 : 7);  // index bits inside a key
  static constexpr uint32_t IMASK = (1u << IB) - 1u;
  static constexpr int N2 = H <= 32 ? 32 : (H <= 64 ? 64 : 128);
  static constexpr uint32_t kNever = 0x7f800000u;  // +inf: never due
  float nk[2];
  int pos[2];

  __device__ __forceinline__ void build(uint32_t* __restrict__ ks, const float* __restrict__ tb,
                                        const float* __restrict__ cf, float t_start, float dirsign) {
    const int tid = threadIdx.x;
    const float* rinv = tb + P::N;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t k[N2];
#pragma unroll
      for (int j = 0; j < N2; ++j) {
        k[j] = kNever;
        if (j < H) {
          const float ts = cf[(j * kBlock + tid) * 2 + h] * rinv[j];
          const float slack = 4e-7f * (fabsf(ts) + fabsf(t_start));
          const float u = dirsign * (ts - t_start);
          const float ub = fmaxf(fmaf(u, 0.99998474f, -slack), 0.0f);
          // pending iff the crossing is not behind the start (rinv = 0 marks w1t_j = 0: never flips; NaN fails the test)
          if (rinv[j] != 0.0f && u > -(64.0f * slack + 1e-30f) && ub < 3.0e38f)
            k[j] = (__float_as_uint(ub) & ~IMASK) | (uint32_t)j;
        }
      }
#pragma unroll
      for (int size = 2; size <= N2; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
          for (int i = 0; i < N2; ++i) {
            const int l = i ^ stride;
            if (l > i) {
              const uint32_t a = k[i], b = k[l];
              const bool up = (i & size) == 0;
              k[i] = up ? min(a, b) : max(a, b);
              k[l] = up ? max(a, b) : min(a, b);
            }
          }
        }
      }
      uint32_t* col = ks + (size_t)h * (H + 1) * kBlock + tid;
#pragma unroll
      for (int j = 0; j < H; ++j) col[j * kBlock] = k[j];
      col[H * kBlock] = kNever;
      pos[h] = 0;
      nk[h] = __uint_as_float(k[0]);
    }
  }

  // move the gate pattern and (alpha, beta) to evaluation time te.  One loop per trajectory of the thread: a trip
  // usually serves a single lane of the warp, so serving both halves in one trip would double its cost.
  template <int HALF>
  __device__ __forceinline__ void advance_half(const uint32_t* __restrict__ ks, const float* __restrict__ tb,
                                               const float* __restrict__ cf, float te, float uq, float dirsign) {
    const int tid = threadIdx.x;
    while (uq >= nk[HALF]) {
      const int j = (int)(__float_as_uint(nk[HALF]) & IMASK);
      const float* r = tb + P::KP + j * P::UNIT;
      float a[P::UNIT];
#pragma unroll
      for (int k = 0; k < P::UNIT / 4; ++k) {
        const float4 x = reinterpret_cast<const float4*>(r)[k];
        a[4 * k] = x.x; a[4 * k + 1] = x.y; a[4 * k + 2] = x.z; a[4 * k + 3] = x.w;
      }
      const float c = cf[(j * kBlock + tid) * 2 + HALF];
      // state after the crossing, and the gate the dense test gives at te (on <=> sign bit clear)
      const bool post = dirsign * a[0] > 0.0f;
      const bool now = (__float_as_uint(fmaf(a[0], te, c)) >> 31) == 0u;
      const int w = j >> 5;
      const int b = ((w == NW - 1) ? (H - 32 * w) : 32) - 1 - (j & 31);
      uint32_t cw = cur[HALF][0];
#pragma unroll
      for (int ww = 1; ww < NW; ++ww) {
        if (w == ww) cw = cur[HALF][ww];
      }
      const bool is = (cw >> b) & 1u;
      if (is != post) {        // not yet in the state behind the crossing
        if (now != post) break;  // and not there at te either: the candidate stays pending
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) {
          if (w == ww) cur[HALF][ww] ^= 1u << b;
        }
        const float sgn = post ? 1.0f : -1.0f;
        const float u = sgn * a[0], v = sgn * c;
#pragma unroll
        for (int k = 0; k < K2; ++k) {
          float lo, hi;
          unpk(al[k], lo, hi);
          al[k] = HALF ? pk(lo, fmaf(a[1 + k], u, hi)) : pk(fmaf(a[1 + k], u, lo), hi);
          unpk(be[k], lo, hi);
          be[k] = HALF ? pk(lo, fmaf(a[1 + k], v, hi)) : pk(fmaf(a[1 + k], v, lo), hi);
        }
      }
      ++pos[HALF];
      nk[HALF] = __uint_as_float(ks[((size_t)HALF * (H + 1) + pos[HALF]) * kBlock + tid]);
    }
  }
  __device__ __forceinline__ void advance(const uint32_t* __restrict__ ks, const float* __restrict__ tb,
                                          const float* __restrict__ cf, float te, float t_start, float dirsign) {
    const float uq = dirsign * (te - t_start);
    advance_half<0>(ks, tb, cf, te, uq, dirsign);
    advance_half<1>(ks, tb, cf, te, uq, dirsign);
  }

  // A = sigmoid(growth heads), ND = -sigmoid(degradation heads) at time te (same conventions as mlp_eval).
  // The XU pipe (16 MUFU lanes per SM) is what bounds this kernel once the dense products are gone, so the 2S
  // reciprocals are taken two denominators at a time: 1/a = b * rcp(ab), 1/b = a * rcp(ab) -- one MUFU.RCP and
  // three packed multiplies instead of two MUFU.RCP.  Exponents are clamped at 2^60 so that ab stays finite
  // (sigmoid floor 1e-18).
  __device__ __forceinline__ void eval(float te, Vec<S>& A, Vec<S>& ND) const {
    const f2 tt = bc(te);
    const f2 one = bc(1.0f), minus_one = bc(-1.0f);
    f2 d[K2];
#pragma unroll
    for (int o = 0; o < K2; ++o) {
      float v0, v1;
      unpk(fma2(al[o], tt, be[o]), v0, v1);
      const f2 ex = pk(ex2_approx(fminf(v0, 60.0f)), ex2_approx(fminf(v1, 60.0f)));
      d[o] = o < S ? add2(ex, one) : sub2(minus_one, ex);
    }
    f2 r[K2];
#pragma unroll
    for (int o = 0; o + 1 < K2; o += 2) {
      float m0, m1;
      unpk(mul2(d[o], d[o + 1]), m0, m1);
      const f2 rm = pk(rcp_approx(m0), rcp_approx(m1));
      r[o] = mul2(rm, d[o + 1]);
      r[o + 1] = mul2(rm, d[o]);
    }
    if (K2 & 1) {
      float m0, m1;
      unpk(d[K2 - 1], m0, m1);
      r[K2 - 1] = pk(rcp_approx(m0), rcp_approx(m1));
    }
#pragma unroll
    for (int o = 0; o < K2; ++o) {
      if (o < S) A.v[o] = r[o]; else ND.v[o - S] = r[o];
    }
  }
};

// shared-memory block of the piecewise-linear evaluator:
//   tb   [Pack::N]        copy of one slot of the packed weights
//   rinv [H, padded]      -1 / w1t_j (0 where w1t_j = 0)
//   keys [2][H+1][kBlock] sorted crossing keys of the thread's two trajectories (SLODE_PL == 2)
// The sorted walk serves hidden layers of up to SLODE_WALK_MAXH = 32 units (one gate word, 32-key sorting network
// in registers); wider layers test all H gates per evaluation.  Built with the walk (-DSLODE_WALK_MAXH=64) the
// (64,5) midpoint reverse sweep fails parity at the default ptxas -O3 and passes at -Xptxas -O1, with the same
// shared-memory layout passing without the walk: a code-generation problem under ~1.4 KB of spills, see DESIGN.md.
template <int H>
#ifndef SLODE_WALK_MAXH
#define SLODE_WALK_MAXH 32
#endif
__host__ __device__ constexpr bool pl_walk() { return SLODE_PL == 2 && H <= SLODE_WALK_MAXH; }
template <int H, int S>
__host__ __device__ constexpr int pl_smem_floats() {
  return Pack<H, S>::N + (H + 3) / 4 * 4 + (pl_walk<H>() ? 2 * (H + 1) * kBlock : 0);
}
struct PlCtx {
  const float* tb;
  const float* cf;
  uint32_t* ks;
  float dirsign;  // +1: this kernel visits increasing times, -1: decreasing
};
template <int H, int S>
__device__ __forceinline__ PlCtx pl_ctx(float* base, const float* cf, float dirsign) {
  return PlCtx{base, cf, reinterpret_cast<uint32_t*>(base + Pack<H, S>::N + (H + 3) / 4 * 4), dirsign};
}

// NE evaluations through the piecewise-linear evaluator, visited in sweep order (REV: last index first) so that
// the gate pattern moves monotonically; the first evaluation of a trajectory initialises the coefficients.
template <int H, int S, int NE, bool REV, class CLoad>
__device__ __forceinline__ void pl_evals(PlEval<H, S>& pl, bool& inited, float& t_start, const PlCtx& pc,
                                         const float (&t)[NE], CLoad cj, Vec<S> (&A)[NE], Vec<S> (&ND)[NE],
                                         Gate<H> (&gate)[NE]) {
  using P = Pack<H, S>;
  constexpr bool WALK = pl_walk<H>();
  if (!WALK) gates_only<H, NE>(pc.tb + P::KP, t, cj, gate, P::UNIT);
#pragma unroll
  for (int n = 0; n < NE; ++n) {
    const int e = REV ? NE - 1 - n : n;
    if (!inited) {
      if (WALK) {
        Gate<H> g1[1];
        const float t1[1] = {t[e]};
        gates_only<H, 1>(pc.tb + P::KP, t1, cj, g1, P::UNIT);
        pl.init(pc.tb, pc.cf, g1[0]);
        t_start = t[e];
        pl.build(pc.ks, pc.tb, pc.cf, t_start, pc.dirsign);
      } else {
        pl.init(pc.tb, pc.cf, gate[e]);
      }
      inited = true;
    } else if (WALK) {
      pl.advance(pc.ks, pc.tb, pc.cf, t[e], t_start, pc.dirsign);
    } else {
      pl.update(pc.tb, pc.cf, gate[e]);
    }
    if (WALK) {
#pragma unroll
      for (int w = 0; w < Gate<H>::NW; ++w) {
        gate[e].w[0][w] = pl.cur[0][w];
        gate[e].w[1][w] = pl.cur[1][w];
      }
    }
    pl.eval(t[e], A[e], ND[e]);
  }
}

// copy slot 0 of the packed weights from constant to shared memory (dynamic unit indices need shared memory:
// divergent constant-cache reads serialise) and tabulate -1/w1t_j
template <int H, int S>
__device__ __forceinline__ void pl_stage_tables(float* __restrict__ tb) {
  using P = Pack<H, S>;
  for (int i = threadIdx.x; i < P::N; i += kBlock) tb[i] = SLODE_PACK_SYM[i];
  for (int j = threadIdx.x; j < H; j += kBlock) {
    const float w = SLODE_PACK_SYM[P::KP + j * P::UNIT];
    tb[P::N + j] = (w == 0.0f) ? 0.0f : -1.0f / w;
  }
}

// f = A - D*x = A + ND*x
template <int S>
__device__ __forceinline__ Vec<S> rhs(const Vec<S>& A, const Vec<S>& ND, const Vec<S>& x) { return vfma<S>(ND, x, A); }


// trajectory pair of a thread
struct PairIdx {
  int64_t b0, b1;
  bool ok0, ok1;
};
__device__ __forceinline__ PairIdx pair_index(int64_t tile, int64_t B) {
  const int64_t r = 2 * (tile * kBlock + threadIdx.x);
  PairIdx p;
  p.ok0 = r < B;
  p.ok1 = r + 1 < B;
  p.b0 = p.ok0 ? r : B - 1;      // tail threads redo trajectory B-1 with stores / cotangents masked off
  p.b1 = p.ok1 ? r + 1 : B - 1;
  return p;
}

// Sum K values (K a power of two <= 32) over the 32 lanes of the warp with ~K (not 5K) shuffles: in every round
// each lane keeps half of its values and hands the other half to its partner, so after log2(K) rounds a lane
// holds ONE value (index `slot`) summed over the lanes met so far; the remaining rounds are plain butterflies.
// On return the lanes with (lane & (32/K - 1)) == 0 hold the warp totals of K distinct slots.
template <int K>
__device__ __forceinline__ float warp_sum_scatter(float (&v)[K], int lane, int& slot) {
  int base = 0;
  int n = K;
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
// warp total of every slot added to dst(slot) (a shared-memory accumulator, or null to drop the slot)
template <int K, class Dst>
__device__ __forceinline__ void warp_reduce_to(float (&v)[K], int lane, Dst dst) {
  int slot;
  const float tot = warp_sum_scatter<K>(v, lane, slot);
  if ((lane & (32 / K - 1)) == 0) {
    float* p = dst(slot);
    if (p) atomicAdd(p, tot);
  }
}

// ---------------------------------------------------------------------------------------------
// fused prologue: c = z W1[:,1:]^T + b1 and x0 = latent_to_ode_net(z), weights staged in shared memory
// ---------------------------------------------------------------------------------------------
struct LatSmem {
  float *Wz, *Wa, *b1, *ba, *Wb, *bb;  // Wz, Wa: [L][H] (transposed);  Wb: [H][S] (transposed)
};
__host__ __device__ inline int lat_floats(int L, int H, int S) { return 2 * L * H + 2 * H + H * S + S; }

template <int H, int S>
__device__ __forceinline__ LatSmem lat_stage(float* base, const LatentSrc& lat) {  // caller syncs afterwards
  const int L = lat.L;
  LatSmem ls;
  ls.Wz = base;
  ls.Wa = ls.Wz + L * H;
  ls.b1 = ls.Wa + L * H;
  ls.ba = ls.b1 + H;
  ls.Wb = ls.ba + H;
  ls.bb = ls.Wb + H * S;
  const bool fx0 = lat.Wa != nullptr;
  for (int i = threadIdx.x; i < L * H; i += kBlock) {
    const int l = i / H, j = i % H;
    ls.Wz[i] = lat.W1[j * (L + 1) + 1 + l];
    ls.Wa[i] = fx0 ? lat.Wa[j * L + l] : 0.0f;
  }
  for (int i = threadIdx.x; i < H; i += kBlock) {
    ls.b1[i] = lat.b1[i];
    ls.ba[i] = fx0 ? lat.ba[i] : 0.0f;
  }
  for (int i = threadIdx.x; i < H * S; i += kBlock) ls.Wb[i] = fx0 ? lat.Wb[(i % S) * H + i / S] : 0.0f;
  for (int i = threadIdx.x; i < S; i += kBlock) ls.bb[i] = fx0 ? lat.bb[i] : 0.0f;
  return ls;
}

// hidden pre-activations of both small nets for the thread's two trajectories: c (dynamics) and ha (x0 net)
template <int H, bool WANT_HA>
__device__ __forceinline__ void lat_hidden(const LatSmem& ls, const LatentSrc& lat, const PairIdx& pi, f2 (&c2)[H],
                                           f2 (&ha)[H]) {
#pragma unroll
  for (int j = 0; j < H; ++j) {
    c2[j] = bc(ls.b1[j]);
    ha[j] = bc(ls.ba[j]);
  }
  const float* z0 = lat.z + pi.b0 * lat.L;
  const float* z1 = lat.z + pi.b1 * lat.L;
  // the latent row is read one element ahead of its use (each element feeds 2H dependent-free FFMA2, enough to
  // cover an L1 hit but not a miss taken at the point of use)
  f2 znext = lat.L > 0 ? pk(__ldg(z0), __ldg(z1)) : 0ull;
#pragma unroll 1
  for (int l = 0; l < lat.L; ++l) {
    const f2 zl = znext;
    if (l + 1 < lat.L) znext = pk(ld_early(z0 + l + 1), ld_early(z1 + l + 1));
    const float* wz = ls.Wz + l * H;
    const float* wa = ls.Wa + l * H;
#pragma unroll
    for (int j = 0; j < H; ++j) {
      c2[j] = fma2(bc(wz[j]), zl, c2[j]);
      if (WANT_HA) ha[j] = fma2(bc(wa[j]), zl, ha[j]);
    }
  }
}

template <int H, int S>
__device__ __forceinline__ Vec<S> lat_x0(const LatSmem& ls, const f2 (&ha)[H]) {
  Vec<S> x;
  f2 acc[S];
#pragma unroll
  SLODE_FOR_S acc[s] = bc(ls.bb[s]);
#pragma unroll
  for (int j = 0; j < H; ++j) {
    float p0, p1;
    unpk(ha[j], p0, p1);
    const f2 h = pk(fmaxf(p0, 0.0f), fmaxf(p1, 0.0f));
#pragma unroll
    SLODE_FOR_S acc[s] = fma2(bc(ls.Wb[j * S + s]), h, acc[s]);
  }
#pragma unroll
  SLODE_FOR_S {
    float v0, v1;
    unpk(mul2(acc[s], bc(kNegLog2e)), v0, v1);
    x.v[s] = pk(sigmoid_from_scaled(v0), sigmoid_from_scaled(v1));
  }
  return x;
}

// Output stage of the forward kernel.  With (T,B,S) storage a warp's 64 trajectories are adjacent in memory and
// plain stores coalesce.  With (B,T,S) storage (layout="bts", what the decoder reads) a trajectory's rows are
// contiguous in TIME instead: four consecutive output times are staged in shared memory and written as one
// run of 4*S floats per trajectory with 16-byte stores.
constexpr int kStageT = 4;
template <int S>
struct OutStage {
  float buf[kBlock][2][kStageT * S];
};
template <int S>
__device__ __forceinline__ void out_put(OutStage<S>& os, bool time_major_rows, int k, int T, float* row0, bool ok0,
                                        float* row1, bool ok1, const Vec<S>& x) {
  // row0/row1: start of the trajectory's storage (element (k,s) at row + k*st + s)
  if (!time_major_rows) return;
  const int tid = threadIdx.x, slot = k & (kStageT - 1);
#pragma unroll
  SLODE_FOR_S {
    float lo, hi;
    unpk(x.v[s], lo, hi);
    os.buf[tid][0][slot * S + s] = lo;
    os.buf[tid][1][slot * S + s] = hi;
  }
  if (slot == kStageT - 1 || k == T - 1) {
    const int n = (slot + 1) * S;            // floats staged
    const int k0 = k - slot;                 // first staged output time
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* dst = (h ? row1 : row0) + (int64_t)k0 * S;
      if (!(h ? ok1 : ok0)) continue;
      const float* src = os.buf[tid][h];
      if (n == kStageT * S && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (kStageT * S) % 4 == 0) {
#pragma unroll
        for (int q = 0; q < kStageT * S / 4; ++q)
          reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(src)[q];
      } else {
        for (int q = 0; q < n; ++q) dst[q] = src[q];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int H, int S, int METHOD>
__global__ void __launch_bounds__(kBlock, ((S >= 8 && METHOD == SLODE_METHOD_RK4) ? 2 : SLODE_FWD_MINB))  // S=8 rk4: 48 accumulator pairs
mlp_fixed_fwd_kernel(int64_t B, int T, const float* __restrict__ tgrid, const float* __restrict__ cin,
                     const float* __restrict__ y0, float* __restrict__ sol, int64_t st, int64_t sb, LatentSrc lat,
                     f2* __restrict__ eval_ckpt) {
  extern __shared__ __align__(16) float fwd_dyn[];
  const bool rows_in_time = (st == S);  // (B,T,S)-contiguous storage
  // wide hidden layers: the per-trajectory c_j do not fit in registers next to the accumulators -> shared memory
  constexpr bool PL = SLODE_PL != 0;  // piecewise-linear evaluation of the heads (PlEval)
  constexpr bool C_IN_SMEM = PL || H > 32 || SLODE_FWD_CSMEM;
  f2 (*csm)[kBlock] = reinterpret_cast<f2 (*)[kBlock]>(fwd_dyn);
  float* const tb = fwd_dyn + (C_IN_SMEM ? 2 * H * kBlock : 0);  // PL: packed weights, 1/w1t, crossing keys
  float* lat_base = tb + (PL ? pl_smem_floats<H, S>() : 0);
  const float* const cf = fwd_dyn;
  // the output stage of the (B,T,S) layout sits behind the staged latent nets (launch_fwd sizes the block)
  OutStage<S>& ostage = *reinterpret_cast<OutStage<S>*>(lat_base + (lat.z ? (lat_floats(lat.L, H, S) + 3) / 4 * 4 : 0));
  const PlCtx plc = pl_ctx<H, S>(tb, cf, (T < 2 || __ldg(tgrid + T - 1) >= __ldg(tgrid)) ? 1.0f : -1.0f);
  LatSmem ls{};
  if (PL) pl_stage_tables<H, S>(tb);
  if (lat.z) ls = lat_stage<H, S>(lat_base, lat);
  if (PL || lat.z) __syncthreads();
  const int64_t ntiles = ((B + 1) / 2 + kBlock - 1) / kBlock;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const PairIdx pi = pair_index(tile, B);
    // every thread stores (tail threads their duplicate of trajectory B-1): the tile's region holds finite numbers
    const bool save = eval_ckpt != nullptr;
    f2* const ck = eval_ckpt + ckpt_warp_offset<S>(tile, ckpt_evals<METHOD>(T));
    f2 c2[C_IN_SMEM ? 1 : H];
    Vec<S> x;
    {
      f2 ctmp[H];
      if (lat.z) {
        f2 ha[H];
        if (lat.Wa) {
          lat_hidden<H, true>(ls, lat, pi, ctmp, ha);
          x = lat_x0<H, S>(ls, ha);
        } else {
          lat_hidden<H, false>(ls, lat, pi, ctmp, ha);
          x = vload2<S>(y0 + pi.b0 * S, y0 + pi.b1 * S);
        }
      } else {
#pragma unroll
        for (int j = 0; j < H; ++j) ctmp[j] = pk(ld_stream(cin + pi.b0 * H + j), ld_stream(cin + pi.b1 * H + j));
        x = vload2<S>(y0 + pi.b0 * S, y0 + pi.b1 * S);
      }
#pragma unroll
      for (int j = 0; j < H; ++j) {
        if (C_IN_SMEM) csm[j][threadIdx.x] = ctmp[j]; else c2[C_IN_SMEM ? 0 : j] = ctmp[j];
      }
    }
    auto cj = [&](int j) { return C_IN_SMEM ? csm[j][threadIdx.x] : c2[C_IN_SMEM ? 0 : j]; };
    float* out0 = sol + pi.b0 * sb;
    float* out1 = sol + pi.b1 * sb;
    float* const row0 = out0;
    float* const row1 = out1;
    if (rows_in_time) out_put<S>(ostage, true, 0, T, row0, pi.ok0, row1, pi.ok1, x);
    else vstore2<S>(out0, pi.ok0, out1, pi.ok1, x);
    float t0 = __ldg(tgrid);
    Vec<S> k1;
    PlEval<H, S> pl;
    bool pl_on = false;
    float pl_t0 = 0.0f;
    if (METHOD == SLODE_METHOD_RK4) {  // k1 of the first step; afterwards carried over from the step before
      Vec<S> A[1], D[1];
      Gate<H> ng[1];
      const float te[1] = {t0};
      if (PL) pl_evals<H, S, 1, false>(pl, pl_on, pl_t0, plc, te, cj, A, D, ng);
      else mlp_eval<H, S, 1, false, 1>(te, cj, A, D, ng);
      if (save) ckpt_store<S>(ck, 0, A[0], D[0]);
      k1 = rhs<S>(A[0], D[0], x);
    }

    float t_ahead = T > 1 ? __ldg(tgrid + 1) : t0;  // the grid is read one step ahead of its use
#pragma unroll 1
    for (int i = 0; i + 1 < T; ++i) {
      const float t1 = t_ahead;
      if (i + 2 < T) t_ahead = ld_early(tgrid + i + 2);
      const float dt = t1 - t0;
      if (METHOD == SLODE_METHOD_EULER) {
        Vec<S> A[1], D[1];
        Gate<H> ng[1];
        const float te[1] = {t0};
        if (PL) pl_evals<H, S, 1, false>(pl, pl_on, pl_t0, plc, te, cj, A, D, ng);
        else mlp_eval<H, S, 1, false, 0>(te, cj, A, D, ng);
        if (save) ckpt_store<S>(ck, i, A[0], D[0]);
        x = vaxpy<S>(dt, rhs<S>(A[0], D[0], x), x);
      } else if (METHOD == SLODE_METHOD_MIDPOINT) {
        const float half_dt = 0.5f * dt;
        Vec<S> A[2], D[2];
        Gate<H> ng[2];
        const float te[2] = {t0, t0 + half_dt};
        if (PL) pl_evals<H, S, 2, false>(pl, pl_on, pl_t0, plc, te, cj, A, D, ng);
        else mlp_eval<H, S, 2, false, 0>(te, cj, A, D, ng);
        if (save) {
          ckpt_store<S>(ck, 2 * (int64_t)i, A[0], D[0]);
          ckpt_store<S>(ck, 2 * (int64_t)i + 1, A[1], D[1]);
        }
        const Vec<S> ym = vaxpy<S>(half_dt, rhs<S>(A[0], D[0], x), x);
        x = vaxpy<S>(dt, rhs<S>(A[1], D[1], ym), x);
      } else {  // rk4, 3/8 rule (torchdiffeq rk4_alt_step_func); the three new evaluations are taken together
        Vec<S> A[3], D[3];
        Gate<H> ng[3];
        const float te[3] = {t0 + dt * kOneThird, t0 + dt * kTwoThirds, t1};
        if (PL) pl_evals<H, S, 3, false>(pl, pl_on, pl_t0, plc, te, cj, A, D, ng);
        else mlp_eval<H, S, 3, false, 0>(te, cj, A, D, ng);
        if (save) {
#pragma unroll
          for (int e = 0; e < 3; ++e) ckpt_store<S>(ck, 3 * (int64_t)i + 1 + e, A[e], D[e]);
        }
        Vec<S> y = vaxpy<S>(dt * kOneThird, k1, x);
        const Vec<S> k2 = rhs<S>(A[0], D[0], y);
        y = vaxpy<S>(dt, vaxpy<S>(-kOneThird, k1, k2), x);
        const Vec<S> k3 = rhs<S>(A[1], D[1], y);
        y = vaxpy<S>(dt, vadd<S>(vsub<S>(k1, k2), k3), x);
        const Vec<S> k4 = rhs<S>(A[2], D[2], y);
        x = vaxpy<S>(dt * 0.125f, vadd<S>(vaxpy<S>(3.0f, vadd<S>(k2, k3), k1), k4), x);
        k1 = rhs<S>(A[2], D[2], x);
      }
      out0 += st;
      out1 += st;
      if (rows_in_time) out_put<S>(ostage, true, i + 1, T, row0, pi.ok0, row1, pi.ok1, x);
      else vstore2<S>(out0, pi.ok0, out1, pi.ok1, x);
      t0 = t1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int H, int S>
struct BwdSmem {
  static constexpr int K2 = 2 * S;
  f2 c[H][kBlock];           // c_j of the thread's two trajectories
  float W[H][K2];            // original (unscaled) head weights, [unit][output]
  float w1t[H];
  float G[K2][H];            // block accumulators of dW, [output][unit]
  float gw1t[H];
  float gb[K2];
};

template <int H, int S>
struct Sweep {
  static constexpr int NW = Gate<H>::NW;
  static constexpr int K2 = 2 * S;
  static constexpr int REC = 2 * K2;              // floats per record: P[0..K2) then Q[0..K2) of one trajectory
  static constexpr int REC_PER_THREAD = H * 2 * REC;  // [unit][half][REC]
  f2 P[K2], Q[K2];
  uint32_t prev[2][NW], first[2][NW];

  __device__ __forceinline__ void init(const Gate<H>& g) {
#pragma unroll
    for (int o = 0; o < K2; ++o) P[o] = Q[o] = 0ull;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      prev[0][w] = first[0][w] = g.w[0][w];
      prev[1][w] = first[1][w] = g.w[1][w];
    }
  }

  // A relu gate of one of the thread's trajectories flipped between the previous contributing evaluation and
  // this one: record that trajectory's prefix sums as they stand (before this evaluation is added) in the
  // unit's slot of the thread's scratch.  p_j(t) = w1t_j t + c_j is monotone in t, so each unit flips at most
  // once per sweep.  One loop serves both halves (the half is selected at run time) to keep the code small.
  __device__ __forceinline__ void events(float* __restrict__ rec, const Gate<H>& g) {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      uint32_t d0 = g.w[0][w] ^ prev[0][w];
      uint32_t d1 = g.w[1][w] ^ prev[1][w];
      const int nw = (w == NW - 1) ? (H - 32 * w) : 32;
      while (d0 | d1) {
        const bool hi = d0 == 0u;
        const uint32_t d = hi ? d1 : d0;
        const int q = __ffs(d) - 1;
        if (hi) d1 &= d1 - 1; else d0 &= d0 - 1;
        const int j = 32 * w + (nw - 1 - q);
        float4* dst = reinterpret_cast<float4*>(rec + (j * 2 + (hi ? 1 : 0)) * REC);
        float v[REC];
#pragma unroll
        for (int o = 0; o < K2; ++o) {
          float a, b;
          unpk(P[o], a, b);
          v[o] = hi ? b : a;
          unpk(Q[o], a, b);
          v[K2 + o] = hi ? b : a;
        }
#pragma unroll
        for (int k = 0; k < REC / 4; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      }
      prev[0][w] = g.w[0][w];
      prev[1][w] = g.w[1][w];
    }
  }

  // add the cotangents of the head pre-activations of one evaluation at time te:
  //   f = A - D*y with upstream gf:  d(pre_A) = gf*A(1-A),  d(pre_D) = -gf*y*D(1-D)
  __device__ __forceinline__ void add(float te, const Vec<S>& gf, const Vec<S>& y, const Vec<S>& A, const Vec<S>& ND) {
    const f2 tt = bc(te);
#pragma unroll
    SLODE_FOR_S {
      const f2 dg = mul2(gf.v[s], fma2(neg2(A.v[s]), A.v[s], A.v[s]));                 // gf * (A - A^2)
      const f2 dd = mul2(mul2(gf.v[s], y.v[s]), fma2(ND.v[s], ND.v[s], ND.v[s]));      // gf*y * (D^2 - D), ND = -D
      P[s] = add2(P[s], dg);
      Q[s] = fma2(dg, tt, Q[s]);
      P[S + s] = add2(P[S + s], dd);
      Q[S + s] = fma2(dd, tt, Q[S + s]);
    }
  }

  // End of the sweep: for every hidden unit (uniform loop, all lanes busy) combine the recorded and the final
  // prefix sums into the sums over the evaluations where the unit was active,
  //     active throughout: final      turned off: record      turned on: final - record      never: 0
  // and turn them into dc_j (per trajectory), dw1t_j and dW_oj (warp reduction, one shared atomic per warp and
  // value).  The next unit's records are loaded while the current one is processed.
  __device__ __forceinline__ void finish(BwdSmem<H, S>& sm, const float* __restrict__ rec, float* gc0, float* gc1,
                                         bool gc_to_smem) {
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    constexpr int KR = (K2 + 1 <= 16) ? 16 : 32;  // values reduced per unit, padded to a power of two

    auto coeffs = [&](int j, float (&arec)[2], float (&afin)[2]) {
      const int w = j >> 5;
      const int nw = (w == NW - 1) ? (H - 32 * w) : 32;
      const int bit = nw - 1 - (j - 32 * w);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t fw = first[h][0], lw = prev[h][0];
#pragma unroll
        for (int ww = 1; ww < NW; ++ww) {
          if (w == ww) { fw = first[h][ww]; lw = prev[h][ww]; }
        }
        const bool f = (fw >> bit) & 1u, l = (lw >> bit) & 1u;
        arec[h] = (f == l) ? 0.0f : (f ? 1.0f : -1.0f);
        afin[h] = l ? 1.0f : 0.0f;
      }
    };
    auto load = [&](int j, const float (&arec)[2], float4 (&a)[REC / 4], float4 (&b)[REC / 4]) {
      const float4* r0 = reinterpret_cast<const float4*>(rec + (j * 2 + 0) * REC);
      const float4* r1 = reinterpret_cast<const float4*>(rec + (j * 2 + 1) * REC);
#pragma unroll
      for (int k = 0; k < REC / 4; ++k) {
        a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        b[k] = a[k];
        if (arec[0] != 0.0f) a[k] = r0[k];
        if (arec[1] != 0.0f) b[k] = r1[k];
      }
    };

    float arec[2], afin[2];
    float4 ra[REC / 4], rb[REC / 4];
    coeffs(0, arec, afin);
    load(0, arec, ra, rb);
#pragma unroll 1
    for (int j = 0; j < H; ++j) {
      float v0[REC], v1[REC];
#pragma unroll
      for (int k = 0; k < REC / 4; ++k) {
        v0[4 * k] = ra[k].x; v0[4 * k + 1] = ra[k].y; v0[4 * k + 2] = ra[k].z; v0[4 * k + 3] = ra[k].w;
        v1[4 * k] = rb[k].x; v1[4 * k + 1] = rb[k].y; v1[4 * k + 2] = rb[k].z; v1[4 * k + 3] = rb[k].w;
      }
      const f2 ar = pk(arec[0], arec[1]), af = pk(afin[0], afin[1]);
      if (j + 1 < H) {  // next unit's records: in flight during this unit's arithmetic
        coeffs(j + 1, arec, afin);
        load(j + 1, arec, ra, rb);
      }
      const f2 wj = bc(sm.w1t[j]);
      const f2 cj = sm.c[j][tid];
      f2 s1 = 0ull, s2 = 0ull;
      float red[KR];
#pragma unroll
      for (int k = 0; k < KR; ++k) red[k] = 0.0f;
#pragma unroll
      for (int o = 0; o < K2; ++o) {
        const f2 pe = fma2(ar, pk(v0[o], v1[o]), mul2(af, P[o]));
        const f2 qe = fma2(ar, pk(v0[K2 + o], v1[K2 + o]), mul2(af, Q[o]));
        const f2 wo = bc(sm.W[j][o]);
        s1 = fma2(wo, pe, s1);
        s2 = fma2(wo, qe, s2);
        float lo, hi;
        unpk(fma2(wj, qe, mul2(cj, pe)), lo, hi);
        red[o] = lo + hi;
      }
      {
        float lo, hi;
        unpk(s1, lo, hi);
        if (gc_to_smem) {
          sm.c[j][tid] = s1;  // c_j is not needed any more: the fused epilogue picks dL/dc_j up from here
        } else {
          if (gc0) gc0[j] = lo;
          if (gc1) gc1[j] = hi;
        }
        unpk(s2, lo, hi);
        red[K2] = lo + hi;
      }
      warp_reduce_to<KR>(red, lane, [&](int slot) -> float* {
        return slot < K2 ? &sm.G[slot][j] : (slot == K2 ? &sm.gw1t[j] : nullptr);
      });
    }
    // head biases: total of the cotangents over all evaluations
    {
      float red[KR];
#pragma unroll
      for (int k = 0; k < KR; ++k) red[k] = 0.0f;
#pragma unroll
      for (int o = 0; o < K2; ++o) {
        float lo, hi;
        unpk(P[o], lo, hi);
        red[o] = lo + hi;
      }
      warp_reduce_to<KR>(red, lane, [&](int slot) -> float* { return slot < K2 ? &sm.gb[slot] : nullptr; });
    }
  }
};

// Fused epilogue of the reverse sweep (after Sweep::finish left dL/dc_j in sm.c): gradients of the two small
// nets in front of the solve.
//     dz_l     = sum_j dc_j W1z_jl (discrete mode only: odeint_adjoint gives z no gradient through the dynamics)
//              + sum_j da_j Wa_jl,   da_j = [ha_j > 0] sum_s Wb_sj db_s,   db = dL/dx0 * x0 (1 - x0)
//     dW1z_jl += dc_j z_l,  db1_j += dc_j,  dWa_jl += da_j z_l,  dba_j += da_j,  dWb_sj += db_s relu(ha_j),  dbb_s += db_s
// The sums over trajectories go lane -> warp (scatter reduction) -> block accumulators in shared memory.
// reduce vals(k), k in [0,H), over the warp's lanes and both trajectories into dst(k) (shared accumulators)
template <int H, class Val, class Dst>
__device__ __forceinline__ void reduce_units(int lane, Val val, Dst dst) {
  constexpr int KR = (H <= 16) ? 16 : 32;
#pragma unroll
  for (int base = 0; base < H; base += KR) {
    float v[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) {
      float lo = 0.0f, hi = 0.0f;
      if (base + k < H) unpk(val(base + k), lo, hi);
      v[k] = lo + hi;
    }
    warp_reduce_to<KR>(v, lane, [&](int slot) -> float* { return base + slot < H ? dst(base + slot) : nullptr; });
  }
}

template <int H, int S>
__device__ __forceinline__ void lat_epilogue(BwdSmem<H, S>& sm, const LatSmem& ls, float* acc, const LatentSrc& lat,
                                             const PairIdx& pi, bool discrete, const Vec<S>& lam, const Vec<S>& x0,
                                             float* __restrict__ gz) {
  constexpr int KS = 16;
  static_assert(S <= KS, "state dimension");
  const int tid = threadIdx.x, lane = tid & 31;
  const int L = lat.L;
  const bool fx0 = lat.Wa != nullptr;
  float* gWz = acc;
  float* gb1 = gWz + H * L;
  float* gWa = gb1 + H;
  float* gba = gWa + H * L;
  float* gWb = gba + H;
  float* gbb = gWb + S * H;

  f2 gcr[H], da[H];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    gcr[j] = sm.c[j][tid];
    da[j] = 0ull;
  }
  if (fx0) {
    f2 cdummy[H], ha[H];
    lat_hidden<H, true>(ls, lat, pi, cdummy, ha);
    Vec<S> db;
#pragma unroll
    SLODE_FOR_S db.v[s] = mul2(lam.v[s], fma2(neg2(x0.v[s]), x0.v[s], x0.v[s]));
    f2 hr[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      float p0, p1;
      unpk(ha[j], p0, p1);
      hr[j] = pk(fmaxf(p0, 0.0f), fmaxf(p1, 0.0f));
      f2 a = 0ull;
#pragma unroll
      SLODE_FOR_S a = fma2(bc(ls.Wb[j * S + s]), db.v[s], a);
      float a0, a1;
      unpk(a, a0, a1);
      da[j] = pk(p0 > 0.0f ? a0 : 0.0f, p1 > 0.0f ? a1 : 0.0f);
    }
#pragma unroll
    SLODE_FOR_S {
      reduce_units<H>(lane, [&](int k) { return mul2(db.v[s], hr[k]); }, [&](int k) { return gWb + s * H + k; });
    }
    {
      float v[KS];
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        float lo = 0.0f, hi = 0.0f;
        if (k < S) unpk(db.v[k < S ? k : 0], lo, hi);
        v[k] = lo + hi;
      }
      warp_reduce_to<KS>(v, lane, [&](int slot) -> float* { return slot < S ? gbb + slot : nullptr; });
    }
    reduce_units<H>(lane, [&](int k) { return da[k]; }, [&](int k) { return gba + k; });
  }
  reduce_units<H>(lane, [&](int k) { return gcr[k]; }, [&](int k) { return gb1 + k; });
  const float* z0 = lat.z + pi.b0 * L;
  const float* z1 = lat.z + pi.b1 * L;
  f2 znext = L > 0 ? pk(__ldg(z0), __ldg(z1)) : 0ull;
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const f2 zl = znext;
    if (l + 1 < L) znext = pk(ld_early(z0 + l + 1), ld_early(z1 + l + 1));
    const float* wz = ls.Wz + l * H;
    const float* wa = ls.Wa + l * H;
    f2 dz = 0ull;
    reduce_units<H>(lane, [&](int k) { return mul2(gcr[k], zl); }, [&](int k) { return gWz + k * L + l; });
    if (discrete) {
#pragma unroll
      for (int j = 0; j < H; ++j) dz = fma2(bc(wz[j]), gcr[j], dz);
    }
    if (fx0) {
      reduce_units<H>(lane, [&](int k) { return mul2(da[k], zl); }, [&](int k) { return gWa + k * L + l; });
#pragma unroll
      for (int j = 0; j < H; ++j) dz = fma2(bc(wa[j]), da[j], dz);
    }
    float d0, d1;
    unpk(dz, d0, d1);
    if (pi.ok0) gz[pi.b0 * L + l] = d0;
    if (pi.ok1) gz[pi.b1 * L + l] = d1;
  }
}

template <int S, bool CKPT>
__host__ __device__ constexpr bool bwd_pl() { return SLODE_PL != 0 && !CKPT && S <= 5; }

template <int H, int S, int METHOD, int MODE, bool CKPT>
__global__ void __launch_bounds__(kBlock, SLODE_BWD_MINB)
mlp_fixed_bwd_kernel(int64_t B, int T, const float* __restrict__ tgrid, const float* __restrict__ cin,
                     const float* __restrict__ w1t, const float* __restrict__ Wg, const float* __restrict__ Wd,
                     const float* __restrict__ sol, int64_t st, int64_t sb,
                     const float* __restrict__ gsol, int64_t gst, int64_t gsb,
                     float* __restrict__ grad_y0, float* __restrict__ grad_c, float* __restrict__ grad_w,
                     float* __restrict__ flip_ws, LatentSrc lat, float* __restrict__ grad_z,
                     const f2* __restrict__ eval_ckpt) {
  static_assert(!CKPT || MODE == SLODE_BWD_DISCRETE, "evaluation checkpoints serve the discrete sweep");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem<H, S>& sm = *reinterpret_cast<BwdSmem<H, S>*>(smem_raw);
  constexpr int K2 = 2 * S;
  const int tid = threadIdx.x;
  // checkpointed sweep (S <= 5): per-warp double-buffered stages filled by bulk copies
  constexpr bool STAGED = CKPT && S <= 5;
  constexpr int per_step = (METHOD == SLODE_METHOD_RK4) ? 3 : (METHOD == SLODE_METHOD_MIDPOINT ? 2 : 1);
  constexpr int kStageF2 = per_step * K2 * 32;                           // f2 per warp and interval
  constexpr int kWarps = kBlock / 32;
  constexpr size_t kStageBytes = STAGED ? (size_t)kWarps * 2 * kStageF2 * sizeof(f2) + kWarps * 2 * sizeof(uint64_t) : 0;
  // heads re-evaluated piecewise-linearly (PlEval).  Not for S > 5: (alpha, beta) next to the prefix sums P, Q are
  // 8S more live registers, and at S = 8 the sweep spills (measured 20 ms against 12.7 ms for the proc shape)
  constexpr bool PL = bwd_pl<S, CKPT>();
  float* const tb = reinterpret_cast<float*>(smem_raw + (sizeof(BwdSmem<H, S>) + 15) / 16 * 16);
  const float* const cf = reinterpret_cast<const float*>(&sm.c[0][0]);
  unsigned char* stage_raw = reinterpret_cast<unsigned char*>(tb + (PL ? pl_smem_floats<H, S>() : 0));
  // the reverse sweep visits the grid from its last time to its first
  const PlCtx plc = pl_ctx<H, S>(tb, cf, (T < 2 || __ldg(tgrid + T - 1) >= __ldg(tgrid)) ? -1.0f : 1.0f);
  const int warp = tid >> 5, lane = tid & 31;
  f2* const wstage = reinterpret_cast<f2*>(stage_raw) + (size_t)warp * 2 * kStageF2;   // this warp's two stages
  uint64_t* const wbars = reinterpret_cast<uint64_t*>(stage_raw + (size_t)kWarps * 2 * kStageF2 * sizeof(f2)) + warp * 2;
  uint32_t uses0 = 0u, uses1 = 0u;  // completed fills of each stage (its mbarrier's phase parity)
  if (STAGED && lane == 0) {
    mbar_init(&wbars[0], 1);
    mbar_init(&wbars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // fused mode: staged weights of the two small nets, then the block accumulators of their gradients
  float* ext = reinterpret_cast<float*>(stage_raw + (kStageBytes + 15) / 16 * 16);
  LatSmem ls{};
  float* lat_acc = nullptr;
  int n_lat_acc = 0;
  if (lat.z) {
    ls = lat_stage<H, S>(ext, lat);
    lat_acc = ext + lat_floats(lat.L, H, S);
    n_lat_acc = lat.Wa ? lat_floats(lat.L, H, S) : (H * lat.L + H);
    for (int i = tid; i < lat_floats(lat.L, H, S); i += kBlock) lat_acc[i] = 0.0f;
  }

  for (int i = tid; i < H * K2; i += kBlock) {
    const int j = i / K2, o = i % K2;
    sm.W[j][o] = (o < S) ? Wg[o * H + j] : Wd[(o - S) * H + j];
  }
  for (int i = tid; i < H; i += kBlock) {
    sm.w1t[i] = w1t[i];
    sm.gw1t[i] = 0.0f;
  }
  for (int i = tid; i < K2 * H; i += kBlock) (&sm.G[0][0])[i] = 0.0f;
  if (tid < K2) sm.gb[tid] = 0.0f;
  if (PL) pl_stage_tables<H, S>(tb);
  __syncthreads();

  const int64_t ntiles = ((B + 1) / 2 + kBlock - 1) / kBlock;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const PairIdx pi = pair_index(tile, B);
    const f2* const ck = CKPT ? eval_ckpt + ckpt_warp_offset<S>(tile, ckpt_evals<METHOD>(T)) : nullptr;
    // a masked-off half aliases trajectory B-1 (owned by another half): it must never touch grad_c
    float* gc0 = pi.ok0 ? grad_c + pi.b0 * H : nullptr;
    float* gc1 = pi.ok1 ? grad_c + pi.b1 * H : nullptr;
    // this thread's flip-record slots (re-used tile after tile)
    float* rec = flip_ws + ((size_t)blockIdx.x * kBlock + tid) * Sweep<H, S>::REC_PER_THREAD;
    if (lat.z) {
      f2 c2[H], ha[H];
      lat_hidden<H, false>(ls, lat, pi, c2, ha);
#pragma unroll
      for (int j = 0; j < H; ++j) sm.c[j][tid] = c2[j];
    } else {
#pragma unroll
      for (int j = 0; j < H; ++j)
        sm.c[j][tid] = pk(ld_stream(cin + pi.b0 * H + j), ld_stream(cin + pi.b1 * H + j));
    }
    auto cj = [&](int j) { return sm.c[j][tid]; };
    const float* xs0 = sol + pi.b0 * sb;
    const float* xs1 = sol + pi.b1 * sb;
    const float* gs0 = gsol + pi.b0 * gsb;
    const float* gs1 = gsol + pi.b1 * gsb;
    const f2 live = pk(pi.ok0 ? 1.0f : 0.0f, pi.ok1 ? 1.0f : 0.0f);
    Vec<S> lam = vscale2<S>(vload2<S>(gs0 + (int64_t)(T - 1) * gst, gs1 + (int64_t)(T - 1) * gst), live);

    // lane 0 streams interval n (counted from the end: i = T-2-n) of this warp into its stage n & 1
    auto stage_issue = [&](int n) {
      const int st_ = n & 1;
      mbar_expect_tx(&wbars[st_], (uint32_t)(kStageF2 * sizeof(f2)));
      bulk_g2s(wstage + st_ * kStageF2, ck + (int64_t)per_step * (T - 2 - n) * K2 * 32, (uint32_t)(kStageF2 * sizeof(f2)),
               &wbars[st_]);
    };
    if (STAGED && lane == 0) {
      if (T >= 2) stage_issue(0);
      if (T >= 3) stage_issue(1);
    }
    Sweep<H, S> sw;
    PlEval<H, S> pl;
    bool pl_on = false;
    float pl_t0 = 0.0f;
    float t1 = __ldg(tgrid + T - 1);
    bool started = false;
    Vec<S> Ac, Dc;  // rk4: the evaluation at t1, carried over from the interval processed before
    if (METHOD == SLODE_METHOD_RK4) {
      Vec<S> A[1], D[1];
      Gate<H> g[1];
      const float te[1] = {t1};
      if (CKPT) {
        ckpt_load<S>(ck, 3 * (int64_t)(T - 1), A[0], D[0]);
        gates_only<H, 1>(sm.w1t, te, cj, g);
      } else {
        if (PL) pl_evals<H, S, 1, true>(pl, pl_on, pl_t0, plc, te, cj, A, D, g);
        else mlp_eval<H, S, 1, true, 1>(te, cj, A, D, g);
      }
      Ac = A[0];
      Dc = D[0];
      sw.init(g[0]);
      started = true;
    }

#pragma unroll 1
    for (int i = T - 2; i >= 0; --i) {
      const float t0 = __ldg(tgrid + i);
      const Vec<S> x = vload2<S>(xs0 + (int64_t)i * st, xs1 + (int64_t)i * st);
      // the cotangent row of grid point i is consumed at the END of this interval: issued here (volatile asm keeps
      // it here) its latency hides behind the whole interval -- read at the point of use it cost 12 % of the sweep
      Vec<S> gi;
#pragma unroll
      SLODE_FOR_S gi.v[s] = pk(ld_early(gs0 + (int64_t)i * gst + s), ld_early(gs1 + (int64_t)i * gst + s));
      if (i > 0) {  // next interval's state and cotangent rows: in L1 by the time they are read
        vprefetch<S>(xs0 + (int64_t)(i - 1) * st);
        vprefetch<S>(xs1 + (int64_t)(i - 1) * st);
        vprefetch<S>(gs0 + (int64_t)(i - 1) * gst);
        vprefetch<S>(gs1 + (int64_t)(i - 1) * gst);
        if (CKPT && !STAGED) {
#pragma unroll
          for (int e = 0; e < per_step; ++e) ckpt_prefetch<S>(ck, (int64_t)per_step * (i - 1) + e);
        }
      }

      if (MODE == SLODE_BWD_DISCRETE) {
        const float dt = t1 - t0;
        if (METHOD == SLODE_METHOD_EULER) {
          Vec<S> A[1], D[1];
          Gate<H> g[1];
          const float te[1] = {t0};
          if (STAGED) {
              const int n_ = T - 2 - i, st_ = n_ & 1;
              mbar_wait(&wbars[st_], (st_ ? uses1 : uses0) & 1u);
              if (st_) ++uses1; else ++uses0;
              stage_read<S, 1>(wstage + st_ * kStageF2, A, D);
              __syncwarp();  // every lane has its copy: the warp's stage can be refilled
              if (lane == 0 && n_ + 2 <= T - 2) stage_issue(n_ + 2);
            }
          if (CKPT) {
            if (!STAGED) ckpt_load<S>(ck, i, A[0], D[0]);
            gates_only<H, 1>(sm.w1t, te, cj, g);
          } else if (PL) {
            pl_evals<H, S, 1, true>(pl, pl_on, pl_t0, plc, te, cj, A, D, g);
          } else {
            mlp_eval<H, S, 1, true, 0>(te, cj, A, D, g);
          }
          const Vec<S> gk = vscale<S>(lam, dt);
          if (!started) { sw.init(g[0]); started = true; } else sw.events(rec, g[0]);
          sw.add(t0, gk, x, A[0], D[0]);
          lam = vfma<S>(gk, D[0], lam);
        } else if (METHOD == SLODE_METHOD_MIDPOINT) {
          const float half_dt = 0.5f * dt;
          Vec<S> A[2], D[2];
          Gate<H> g[2];
          const float te[2] = {t0, t0 + half_dt};
          if (STAGED) {
              const int n_ = T - 2 - i, st_ = n_ & 1;
              mbar_wait(&wbars[st_], (st_ ? uses1 : uses0) & 1u);
              if (st_) ++uses1; else ++uses0;
              stage_read<S, 2>(wstage + st_ * kStageF2, A, D);
              __syncwarp();  // every lane has its copy: the warp's stage can be refilled
              if (lane == 0 && n_ + 2 <= T - 2) stage_issue(n_ + 2);
            }
          if (CKPT) {
            if (!STAGED) {
              ckpt_load<S>(ck, 2 * (int64_t)i, A[0], D[0]);
              ckpt_load<S>(ck, 2 * (int64_t)i + 1, A[1], D[1]);
            }
            gates_only<H, 2>(sm.w1t, te, cj, g);
          } else if (PL) {
            pl_evals<H, S, 2, true>(pl, pl_on, pl_t0, plc, te, cj, A, D, g);
          } else {
            mlp_eval<H, S, 2, true, 0>(te, cj, A, D, g);
          }
          const Vec<S> ym = vaxpy<S>(half_dt, rhs<S>(A[0], D[0], x), x);
          Vec<S> gk = vscale<S>(lam, dt);  // dL/dk2
          if (!started) { sw.init(g[1]); started = true; } else sw.events(rec, g[1]);
          sw.add(te[1], gk, ym, A[1], D[1]);
          const Vec<S> gy = vmul<S>(gk, D[1]);  // dL/dy_mid (D holds -sigmoid)
          lam = vadd<S>(lam, gy);
          gk = vscale<S>(gy, half_dt);  // dL/dk1
          sw.events(rec, g[0]);
          sw.add(t0, gk, x, A[0], D[0]);
          lam = vfma<S>(gk, D[0], lam);
        } else {  // rk4 3/8: the three new evaluations (t0, ta, tb) are taken together
          const float dt3 = dt * kOneThird;
          Vec<S> A[3], D[3];
          Gate<H> g[3];
          const float te[3] = {t0, t0 + dt * kOneThird, t0 + dt * kTwoThirds};
          if (STAGED) {
              const int n_ = T - 2 - i, st_ = n_ & 1;
              mbar_wait(&wbars[st_], (st_ ? uses1 : uses0) & 1u);
              if (st_) ++uses1; else ++uses0;
              stage_read<S, 3>(wstage + st_ * kStageF2, A, D);
              __syncwarp();  // every lane has its copy: the warp's stage can be refilled
              if (lane == 0 && n_ + 2 <= T - 2) stage_issue(n_ + 2);
            }
          if (CKPT) {
            if (!STAGED) {
#pragma unroll
              for (int e = 0; e < 3; ++e) ckpt_load<S>(ck, 3 * (int64_t)i + e, A[e], D[e]);
            }
            gates_only<H, 3>(sm.w1t, te, cj, g);
          } else if (PL) {
            pl_evals<H, S, 3, true>(pl, pl_on, pl_t0, plc, te, cj, A, D, g);
          } else {
            mlp_eval<H, S, 3, true, 0>(te, cj, A, D, g);
          }
          Vec<S> Y2, Y3, Y4;
          {
            const Vec<S> k1 = rhs<S>(A[0], D[0], x);
            Y2 = vaxpy<S>(dt3, k1, x);
            const Vec<S> k2 = rhs<S>(A[1], D[1], Y2);
            Y3 = vaxpy<S>(dt, vaxpy<S>(-kOneThird, k1, k2), x);
            const Vec<S> k3 = rhs<S>(A[2], D[2], Y3);
            Y4 = vaxpy<S>(dt, vadd<S>(vsub<S>(k1, k2), k3), x);
          }
          const Vec<S> w = vscale<S>(lam, 0.125f * dt);
          // stage 4 (time t1, carried evaluation; its gates are the sweep's current ones): gk4 = w
          sw.add(t1, w, Y4, Ac, Dc);
          Vec<S> gy = vmul<S>(w, Dc);
          lam = vadd<S>(lam, gy);
          Vec<S> gk1 = vaxpy<S>(dt, gy, w);
          Vec<S> gk2 = vaxpy<S>(-dt, gy, vscale<S>(w, 3.0f));
          const Vec<S> gk3 = vaxpy<S>(dt, gy, vscale<S>(w, 3.0f));
          // stage 3
          sw.events(rec, g[2]);
          sw.add(te[2], gk3, Y3, A[2], D[2]);
          gy = vmul<S>(gk3, D[2]);
          lam = vadd<S>(lam, gy);
          gk2 = vaxpy<S>(dt, gy, gk2);
          gk1 = vaxpy<S>(-dt3, gy, gk1);
          // stage 2
          sw.events(rec, g[1]);
          sw.add(te[1], gk2, Y2, A[1], D[1]);
          gy = vmul<S>(gk2, D[1]);
          lam = vadd<S>(lam, gy);
          gk1 = vaxpy<S>(dt3, gy, gk1);
          // stage 1 (time t0; its evaluation is the carried one of the next interval)
          sw.events(rec, g[0]);
          sw.add(t0, gk1, x, A[0], D[0]);
          lam = vfma<S>(gk1, D[0], lam);
          Ac = A[0];
          Dc = D[0];
        }
      } else {
        // torchdiffeq.odeint_adjoint emulation: one step of the same method on the augmented system
        // [y, a, a_theta] from t1 down to t0, y restarted from the stored sol[i+1].  In reversed time s=-t the
        // step is ds = t1 - t0 > 0 with  Ky = D*y - A,  Ka = -a*D,  a_theta += w_m * a_m^T df/dtheta(t_m, y_m).
        const float ds = t1 - t0;
        const Vec<S> y = vload2<S>(xs0 + (int64_t)(i + 1) * st, xs1 + (int64_t)(i + 1) * st);
        if (METHOD == SLODE_METHOD_EULER) {
          Vec<S> A[1], D[1];
          Gate<H> g[1];
          const float te[1] = {t1};
          if (PL) pl_evals<H, S, 1, false>(pl, pl_on, pl_t0, plc, te, cj, A, D, g);
          else mlp_eval<H, S, 1, true, 0>(te, cj, A, D, g);
          const Vec<S> v = vscale<S>(lam, ds);
          if (!started) { sw.init(g[0]); started = true; } else sw.events(rec, g[0]);
          sw.add(t1, v, y, A[0], D[0]);
          lam = vfma<S>(v, D[0], lam);
        } else if (METHOD == SLODE_METHOD_MIDPOINT) {
          const float half = 0.5f * ds;
          Vec<S> A[2], D[2];
          Gate<H> g[2];
          const float te[2] = {t1, t1 - half};
          if (PL) pl_evals<H, S, 2, false>(pl, pl_on, pl_t0, plc, te, cj, A, D, g);  // the stage at t1 has weight 0: its gates are not used
          else mlp_eval<H, S, 2, true, 0>(te, cj, A, D, g);
          const Vec<S> ym = vaxpy<S>(-half, rhs<S>(A[0], D[0], y), y);  // y + half*(D1*y - A1)
          const Vec<S> am = vaxpy<S>(half, vmul<S>(lam, D[0]), lam);    // a + half*(-a*D1), D holds -sigmoid
          const Vec<S> v = vscale<S>(am, ds);
          if (!started) { sw.init(g[1]); started = true; } else sw.events(rec, g[1]);
          sw.add(te[1], v, ym, A[1], D[1]);
          lam = vfma<S>(v, D[1], lam);
        } else {  // rk4 3/8 on the augmented system; Ky = -f, Ka = -a*D; new evaluations at ta, tb, t0 together
          const float w8 = 0.125f * ds;
          Vec<S> A[3], D[3];
          Gate<H> g[3];
          const float te[3] = {t1 - ds * kOneThird, t1 - ds * kTwoThirds, t0};
          if (PL) pl_evals<H, S, 3, false>(pl, pl_on, pl_t0, plc, te, cj, A, D, g);
          else mlp_eval<H, S, 3, true, 0>(te, cj, A, D, g);
          // (A, D with D = -sigmoid) => f = A + D*y; the augmented step uses Ky = -f, Ka = -a*sigmoid = a*D
          // stage 1 at t1 (carried evaluation)
          const Vec<S> f1 = rhs<S>(Ac, Dc, y);
          const Vec<S> ka1 = vmul<S>(lam, Dc);
          sw.add(t1, vscale<S>(lam, w8), y, Ac, Dc);
          // stage 2
          Vec<S> ym = vaxpy<S>(-ds * kOneThird, f1, y);
          Vec<S> am = vaxpy<S>(ds * kOneThird, ka1, lam);
          const Vec<S> f2_ = rhs<S>(A[0], D[0], ym);
          const Vec<S> ka2 = vmul<S>(am, D[0]);
          sw.events(rec, g[0]);
          sw.add(te[0], vscale<S>(am, 3.0f * w8), ym, A[0], D[0]);
          // stage 3
          ym = vaxpy<S>(-ds, vaxpy<S>(-kOneThird, f1, f2_), y);
          am = vaxpy<S>(ds, vaxpy<S>(-kOneThird, ka1, ka2), lam);
          const Vec<S> f3 = rhs<S>(A[1], D[1], ym);
          const Vec<S> ka3 = vmul<S>(am, D[1]);
          sw.events(rec, g[1]);
          sw.add(te[1], vscale<S>(am, 3.0f * w8), ym, A[1], D[1]);
          // stage 4 at t0 (becomes the carried evaluation)
          ym = vaxpy<S>(-ds, vadd<S>(vsub<S>(f1, f2_), f3), y);
          am = vaxpy<S>(ds, vadd<S>(vsub<S>(ka1, ka2), ka3), lam);
          const Vec<S> ka4 = vmul<S>(am, D[2]);
          sw.events(rec, g[2]);
          sw.add(t0, vscale<S>(am, w8), ym, A[2], D[2]);
          const Vec<S> asum = vadd<S>(vaxpy<S>(3.0f, vadd<S>(ka2, ka3), ka1), ka4);
          lam = vaxpy<S>(w8, asum, lam);
          Ac = A[2];
          Dc = D[2];
        }
      }
      lam = vadd<S>(lam, vscale2<S>(gi, live));
      t1 = t0;
    }

    const bool fused = lat.z != nullptr;
    if (started) {
      sw.finish(sm, rec, gc0, gc1, fused);  // a masked-off half carries zero cotangents: it only adds zeros
    } else {  // T == 1: no evaluation at all
#pragma unroll
      for (int j = 0; j < H; ++j) {
        if (fused) {
          sm.c[j][tid] = 0ull;
        } else {
          if (gc0) gc0[j] = 0.0f;
          if (gc1) gc1[j] = 0.0f;
        }
      }
    }
    if (fused) {
      const Vec<S> x0 = vload2<S>(xs0, xs1);
      lat_epilogue<H, S>(sm, ls, lat_acc, lat, pi, MODE == SLODE_BWD_DISCRETE, lam, x0, grad_z);
    }
    if (!fused || !lat.Wa) vstore2<S>(grad_y0 + pi.b0 * S, pi.ok0, grad_y0 + pi.b1 * S, pi.ok1, lam);
  }

  __syncthreads();
  // flush block accumulators: grad_w = [ dw1t (H) | dWg (S*H) | dbg (S) | dWd (S*H) | dbd (S) ]
  for (int i = tid; i < H; i += kBlock) atomicAdd(grad_w + i, sm.gw1t[i]);
  for (int i = tid; i < K2 * H; i += kBlock) {
    const int o = i / H, j = i % H;
    const int base = (o < S) ? (H + o * H) : (H + S * H + S + (o - S) * H);
    atomicAdd(grad_w + base + j, sm.G[o][j]);
  }
  if (tid < K2) {
    const int base = (tid < S) ? (H + S * H + tid) : (H + S * H + S + S * H + (tid - S));
    atomicAdd(grad_w + base, sm.gb[tid]);
  }
  // fused mode: [ dW1z (H*L) | db1 (H) | dWa (H*L) | dba (H) | dWb (S*H) | dbb (S) ] follow the five segments above
  for (int i = tid; i < n_lat_acc; i += kBlock) atomicAdd(grad_w + (H + 2 * (S * H + S)) + i, lat_acc[i]);
}

// ---------------------------------------------------------------------------------------------
// per-shape launchers (instantiated in slode_mlp_inst_*.cu, dispatched from slode_mlp.cu)
// ---------------------------------------------------------------------------------------------
template <int H, int S, int METHOD>
int launch_fwd(const FwdArgs& a) {
  auto kern = mlp_fixed_fwd_kernel<H, S, METHOD>;
  const size_t smem = (a.lat.z ? sizeof(float) * ((lat_floats(a.lat.L, H, S) + 3) / 4 * 4) : 0) +
                      ((SLODE_PL || H > 32 || SLODE_FWD_CSMEM) ? sizeof(f2) * H * kBlock : 0) +
                      (SLODE_PL ? sizeof(float) * pl_smem_floats<H, S>() : 0) +
                      (a.st == S ? sizeof(OutStage<S>) + 16 : 0);
  if (smem > 16 * 1024)  // static (output stage) + dynamic may pass the 48 KB default
    SLODE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks_per_sm = 0;
  SLODE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kBlock, smem));
  blocks_per_sm = std::max(blocks_per_sm, 1);
  const int64_t tiles = ((a.B + 1) / 2 + kBlock - 1) / kBlock;
  // whole waves of resident blocks; tiles are handed out grid-stride
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)a.sms * blocks_per_sm);
  kern<<<grid, kBlock, smem, a.stream>>>(a.B, a.T, a.t, a.c, a.y0, a.sol, a.st, a.sb, a.lat,
                                         reinterpret_cast<f2*>(a.eval_ckpt));
  SLODE_CUDA_TRY(cudaGetLastError());
  return SLODE_OK;
}

template <int H, int S, int METHOD, int MODE, bool CKPT>
int launch_bwd(const BwdArgs& a) {
  auto kern = mlp_fixed_bwd_kernel<H, S, METHOD, MODE, CKPT>;
  constexpr int per_step = (METHOD == SLODE_METHOD_RK4) ? 3 : (METHOD == SLODE_METHOD_MIDPOINT ? 2 : 1);
  const size_t stage_bytes =
      (CKPT && S <= 5) ? ((size_t)(kBlock / 32) * 2 * per_step * 2 * S * 32 * sizeof(f2) + (kBlock / 32) * 2 * 8 + 15) / 16 * 16 : 0;
  const size_t smem = (sizeof(BwdSmem<H, S>) + 15) / 16 * 16 + stage_bytes +
                      (bwd_pl<S, CKPT>() ? sizeof(float) * pl_smem_floats<H, S>() : 0) +
                      (a.lat.z ? 2 * sizeof(float) * lat_floats(a.lat.L, H, S) : 0);
  SLODE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks_per_sm = 0;
  SLODE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kBlock, smem));
  blocks_per_sm = std::max(blocks_per_sm, 1);
  const int64_t tiles = ((a.B + 1) / 2 + kBlock - 1) / kBlock;
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)a.sms * blocks_per_sm);
  float* ws = flip_workspace(sizeof(float) * (size_t)grid * kBlock * Sweep<H, S>::REC_PER_THREAD);
  if (!ws) return SLODE_ECUDA;
  kern<<<grid, kBlock, smem, a.stream>>>(a.B, a.T, a.t, a.c, a.w1t, a.Wg, a.Wd, a.sol, a.st, a.sb, a.gsol, a.gst,
                                         a.gsb, a.gy0, a.gc, a.gw, ws, a.lat, a.gz,
                                         reinterpret_cast<const f2*>(a.eval_ckpt));
  SLODE_CUDA_TRY(cudaGetLastError());
  return SLODE_OK;
}

template <int H, int S>
int fwd_shape(const FwdArgs& a) {
  switch (a.method) {
    case SLODE_METHOD_EULER: return launch_fwd<H, S, SLODE_METHOD_EULER>(a);
    case SLODE_METHOD_MIDPOINT: return launch_fwd<H, S, SLODE_METHOD_MIDPOINT>(a);
    case SLODE_METHOD_RK4: return launch_fwd<H, S, SLODE_METHOD_RK4>(a);
  }
  set_error("fixed-grid forward: unknown method %d", a.method);
  return SLODE_EINVAL;
}

template <int H, int S>
int bwd_shape(const BwdArgs& a) {
#define SLODE_BWD_CASE(M)                                                                            \
  case M:                                                                                            \
    if (a.mode == SLODE_BWD_DISCRETE)                                                                \
      return a.eval_ckpt ? launch_bwd<H, S, M, SLODE_BWD_DISCRETE, true>(a)                          \
                         : launch_bwd<H, S, M, SLODE_BWD_DISCRETE, false>(a);                        \
    return launch_bwd<H, S, M, SLODE_BWD_TDE_ADJOINT, false>(a);
  switch (a.method) {
    SLODE_BWD_CASE(SLODE_METHOD_EULER)
    SLODE_BWD_CASE(SLODE_METHOD_MIDPOINT)
    SLODE_BWD_CASE(SLODE_METHOD_RK4)
  }
#undef SLODE_BWD_CASE
  set_error("fixed-grid backward: unknown method %d", a.method);
  return SLODE_EINVAL;
}

}  // namespace slode

// Defines the two entry functions of one compiled shape (looked up by slode_mlp.cu through slode_mlp_api.h).
#define SLODE_DEFINE_SHAPE(H, S)                                                                         \
  namespace slode {                                                                                      \
  int mlp_fwd_##H##_##S(const FwdArgs& a, const PackSrc& w, float* staging) {                            \
    const int rc = upload_pack<H, S>(w, staging, a.stream);                                              \
    return rc ? rc : fwd_shape<H, S>(a);                                                                 \
  }                                                                                                      \
  int mlp_bwd_##H##_##S(const BwdArgs& a, const PackSrc& w, float* staging) {                            \
    const int rc = upload_pack<H, S>(w, staging, a.stream);                                              \
    return rc ? rc : bwd_shape<H, S>(a);                                                                 \
  }                                                                                                      \
  }
