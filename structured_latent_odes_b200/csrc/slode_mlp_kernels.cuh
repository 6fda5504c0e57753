// Round-1 building blocks that the dopri5 kernels (slode_dopri5_kernels.cuh) still run on: packed fp32x2
// arithmetic over TWO trajectories per thread, the dynamics weights packed into constant memory and streamed through
// uniform registers, the dense evaluation of the heads (mlp_eval), and the prefix-sum / flip-record bookkeeping of
// the hidden-layer gradients (Sweep).  The fixed-grid kernels that were built on them in round 1 were replaced by
// slode_fixed.cuh (one trajectory per thread, shared-memory weights, piecewise-linear heads) and are gone.
//
// Right-hand side (reference: Dynamics.forward, models/blackbox_ode.py:97-109):
//     h_j(t)  = relu(w1t_j * t + c_j)                       c = z W1[:,1:]^T + b1  (per trajectory)
//     A_k(t)  = sigmoid(bg_k + sum_j Wg_kj h_j(t))          "growth"
//     D_k(t)  = sigmoid(bd_k + sum_j Wd_kj h_j(t))          "degradation"
//     f(t,x)  = A(t) - D(t) * x                              (affine in the state, elementwise)
//
// Weights are warp-uniform scalars streamed from constant memory into uniform registers (one LDCU.128 = four
// weights) and enter the FMA as the broadcast operand:
//     acc_o(traj0,traj1) += W_jo (UR, .F32 broadcast) * h_j(traj0,traj1)
// Measured on B200 (profiles/r01/fp32_pipes_microbench.jsonl): this form sustains 127 of the 128 FMA/clk/SM
// with one LDCU.128 per four FFMA2, where a 3-register FFMA reaches 84.
//
// Backward bookkeeping: because the hidden layer sees only (t, z), the cotangents delta_o(e) of the head
// pre-activations at the evaluation times t_e determine every hidden-layer gradient through prefix sums
//     P_o = sum_e delta_o(e),   Q_o = sum_e delta_o(e) t_e
// taken over the evaluations where unit j is active.  The sweep keeps running P,Q and, whenever a unit's relu
// gate flips between consecutive evaluations, adds +-snapshot contributions
//     dc_j   += s * sum_o W_oj P_o              (per trajectory -> grad_c)
//     dw1t_j += s * sum_o W_oj Q_o              (block accumulator)
//     dW_oj  += s * (w1t_j Q_o + c_j P_o)       (block accumulator, = sum_e delta_o h_j)
// with s=+1 when the unit turns off, -1 when it turns on, and +1 for every unit still active when the sweep ends.
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

}  // namespace slode
