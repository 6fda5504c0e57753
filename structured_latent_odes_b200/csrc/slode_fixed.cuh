// Fixed-grid solve of the SLODE blackbox latent ODE and its reverse sweep, hand-written for sm_100a (round 2).
//
// Right-hand side (reference: Dynamics.forward, models/blackbox_ode.py:97-109):
//     h_j(t)  = relu(w1t_j * t + c_j)                       c = z W1[:,1:]^T + b1  (per trajectory)
//     A_s(t)  = sigmoid(bg_s + sum_j Wg_sj h_j(t))          "growth"
//     D_s(t)  = sigmoid(bd_s + sum_j Wd_sj h_j(t))          "degradation"
//     f(t,x)  = A(t) - D(t) * x                              (affine in the state, elementwise)
// integrated by torchdiffeq's fixed-grid euler / midpoint / rk4 (= 3/8 rule) with grid == output times
// (models/blackbox_ode.py:40-45), reverse sweep = exact discrete adjoint (odeint + autograd) or the
// odeint_adjoint re-discretisation.
//
// Formulation (unchanged from round 1): along one trajectory the head pre-activations are
// piecewise linear in t,  o_k(t) = alpha_k t + beta_k,  with coefficients that change only where a relu gate
// flips; the crossings are sorted once per trajectory and walked in sweep order.  The weight / hidden-layer
// gradients come from prefix sums P = sum delta, Q = sum delta*t of the head cotangents, snapshotted when a gate
// flips ("flip records") and combined per unit at the end of the sweep.
//
// Mapping (new in round 2).  ONE thread = ONE trajectory.  Round 1 packed two trajectories per thread in fp32x2
// halves, which doubled the live registers (254, 2 warps per scheduler) and left the reverse sweep bound by
// instruction issue latency at 0.44 IPC.  Here the fp32x2 halves hold two STATES of the same trajectory
// (s = 2p, 2p+1): every S-vector operation is ceil(S/2) packed instructions, the evaluator's merged reciprocal
// pairs the two sigmoids of a register pair, and the register footprint roughly halves, so that 4-5 blocks of 128
// threads are resident per SM (16-20 warps) and the schedulers always have an eligible warp.
// Weights are staged per block into shared memory straight from the torch tensors (no constant-memory symbol,
// no pack kernel, no cross-call state: the entry points are re-entrant and CUDA-graph capturable).
#pragma once

#include <algorithm>
#include <type_traits>

#include "slode_common.cuh"
#include "slode_mlp_api.h"

#ifndef SLODE_FX_ROW_AHEAD
#define SLODE_FX_ROW_AHEAD 2   // reverse sweep: intervals the state / cotangent rows are fetched ahead (1 or 2)
#endif
#ifndef SLODE_FX_FWD_MINB
#define SLODE_FX_FWD_MINB 5
#endif
#ifndef SLODE_FX_BWD_MINB
#define SLODE_FX_BWD_MINB 4      // resident blocks per SM the reverse sweep is compiled for (euler, midpoint: 128 registers)
#endif
#ifndef SLODE_FX_BWD_MINB_WIDE
#define SLODE_FX_BWD_MINB_WIDE 3 // state dimensions above 5 (proc: S = 8, four register pairs per vector): euler and
                                 // midpoint fit 168 registers, rk4 needs the full 255 (2 blocks)
#endif
#ifndef SLODE_FX_BWD_UNROLL
#define SLODE_FX_BWD_UNROLL 1    // time-loop unrolling of the reverse sweep
#endif
#ifndef SLODE_FX_BWD_MINB_RK4
#define SLODE_FX_BWD_MINB_RK4 3  // rk4 holds three evaluations at once: 168 registers (at 128 it spills in the time loop)
#endif

namespace slode {
namespace fx {

// ---------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (lo = state 2p, hi = state 2p+1 of one trajectory)
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
__device__ __forceinline__ float lo_of(f2 v) {
  float a, b;
  unpk(v, a, b);
  return a;
}
__device__ __forceinline__ float hi_of(f2 v) {
  float a, b;
  unpk(v, a, b);
  return b;
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
// packed negation: the two scalar negations fold into the consumer's operand modifier (FFMA2 R, R, -R, R)
__device__ __forceinline__ f2 neg2(f2 v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(-a), "f"(-b));
  return r;
}
// min that propagates NaN (torch's relu / sigmoid would; fminf launders it)
__device__ __forceinline__ float min_nan(float a, float b) {
  float d;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float relu_nan(float a) {
  float d;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(d) : "f"(a));
  return d;
}
__device__ __forceinline__ float ld_early(const float* p) {  // a load the compiler may not sink to its use
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l1(const float* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const float* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// asynchronous 4-byte copies global -> shared (LDGSTS): no destination register, completion tracked per group and
// not through the scoreboards the evaluator's own loads wait on
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(float* smem_dst, const float* gsrc, bool live) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(live ? 4 : 0)
               : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int kThreads = 128;
constexpr int kBwdUnroll = SLODE_FX_BWD_UNROLL;
constexpr int kWarps = kThreads / 32;
constexpr float kNegLn2 = -0.6931471805599453f;  // unscaled weight = packed weight * kNegLn2

template <int N>
struct V {  // N register pairs = 2N states (the last half is padding when S is odd: it stays exactly zero)
  f2 v[N];
};
#define FX_FOR(N) _Pragma("unroll") for (int p = 0; p < N; ++p)

template <int N> __device__ __forceinline__ V<N> vzero() {
  V<N> r;
  FX_FOR(N) r.v[p] = 0ull;
  return r;
}
template <int N> __device__ __forceinline__ V<N> vadd(const V<N>& a, const V<N>& b) {
  V<N> r;
  FX_FOR(N) r.v[p] = add2(a.v[p], b.v[p]);
  return r;
}
template <int N> __device__ __forceinline__ V<N> vsub(const V<N>& a, const V<N>& b) {
  V<N> r;
  FX_FOR(N) r.v[p] = sub2(a.v[p], b.v[p]);
  return r;
}
template <int N> __device__ __forceinline__ V<N> vmul(const V<N>& a, const V<N>& b) {
  V<N> r;
  FX_FOR(N) r.v[p] = mul2(a.v[p], b.v[p]);
  return r;
}
template <int N> __device__ __forceinline__ V<N> vscale(const V<N>& a, float c) {
  V<N> r;
  const f2 cc = bc(c);
  FX_FOR(N) r.v[p] = mul2(a.v[p], cc);
  return r;
}
template <int N> __device__ __forceinline__ V<N> vaxpy(float c, const V<N>& a, const V<N>& b) {  // c*a + b
  V<N> r;
  const f2 cc = bc(c);
  FX_FOR(N) r.v[p] = fma2(cc, a.v[p], b.v[p]);
  return r;
}
template <int N> __device__ __forceinline__ V<N> vfma(const V<N>& a, const V<N>& b, const V<N>& c) {  // a*b + c
  V<N> r;
  FX_FOR(N) r.v[p] = fma2(a.v[p], b.v[p], c.v[p]);
  return r;
}
// f = A - D*x = G + ND*x  (ND holds MINUS the degradation sigmoid)
template <int N> __device__ __forceinline__ V<N> rhs(const V<N>& G, const V<N>& ND, const V<N>& x) {
  return vfma<N>(ND, x, G);
}

// S floats at p (row of sol / grad_sol / y0) -> pairs; the padding half of an odd S is zero
template <int S> __device__ __forceinline__ V<(S + 1) / 2> vload(const float* p) {
  V<(S + 1) / 2> r;
#pragma unroll
  for (int q = 0; q < (S + 1) / 2; ++q) r.v[q] = pk(__ldg(p + 2 * q), (2 * q + 1 < S) ? __ldg(p + 2 * q + 1) : 0.0f);
  return r;
}
template <int S> __device__ __forceinline__ V<(S + 1) / 2> vload_early(const float* p) {
  V<(S + 1) / 2> r;
#pragma unroll
  for (int q = 0; q < (S + 1) / 2; ++q) r.v[q] = pk(ld_early(p + 2 * q), (2 * q + 1 < S) ? ld_early(p + 2 * q + 1) : 0.0f);
  return r;
}
template <int S> __device__ __forceinline__ void vstore(float* p, bool ok, const V<(S + 1) / 2>& a) {
  if (!ok) return;
#pragma unroll
  for (int q = 0; q < (S + 1) / 2; ++q) {
    float lo, hi;
    unpk(a.v[q], lo, hi);
    p[2 * q] = lo;
    if (2 * q + 1 < S) p[2 * q + 1] = hi;
  }
}
template <int S> __device__ __forceinline__ void vprefetch(const float* p) {
  prefetch_l1(p);
  prefetch_l1(p + S - 1);
}
// a row of S floats of this thread: k-th float at col[k * kThreads] (conflict-free column of a [S][kThreads] block);
// ZFILL: a thread with !live gets zeros (cp.async with a source size of 0 bytes writes zeros and reads nothing)
template <int S, bool ZFILL = false>
__device__ __forceinline__ void row_fetch(float* col, const float* grow, bool live = true) {
#pragma unroll
  for (int k = 0; k < S; ++k) {
    if (ZFILL) cp_async4_zfill(col + k * 128, grow + k, live);
    else cp_async4(col + k * 128, grow + k);
  }
}
template <int S> __device__ __forceinline__ V<(S + 1) / 2> row_read(const float* col) {
  V<(S + 1) / 2> r;
#pragma unroll
  for (int q = 0; q < (S + 1) / 2; ++q) r.v[q] = pk(col[2 * q * 128], (2 * q + 1 < S) ? col[(2 * q + 1) * 128] : 0.0f);
  return r;
}


// ---------------------------------------------------------------------------------------------
// compile-time shape parameters
// ---------------------------------------------------------------------------------------------
template <int H, int S>
struct Shape {
  static constexpr int NP = (S + 1) / 2;               // state pairs
  static constexpr int NQ = 2 * NP;                    // head pairs: growth pairs [0,NP), degradation pairs [NP,NQ)
  static constexpr int UNIT = (2 * NQ + 2 + 3) / 4 * 4;  // floats per unit record: NQ pairs, w1t, -1/w1t, padding
  static constexpr int BIAS = (2 * NQ + 3) / 4 * 4;
  static constexpr int WT = BIAS + H * UNIT;           // floats of the staged weight table
  static constexpr int IB = H <= 32 ? 5 : (H <= 64 ? 6 : (H <= 128 ? 7 : (H <= 256 ? 8 : 9)));  // unit-index bits of a key
  static constexpr uint32_t IMASK = (1u << IB) - 1u;
  static constexpr bool BIG = H > 64;                  // per-trajectory tables in global memory, rolled unit loops
  static constexpr int N2 = H <= 32 ? 32 : 64;         // register sorting network size (!BIG)
  static constexpr int NWR = BIG ? 1 : (H + 31) / 32;  // gate words kept in registers (!BIG)
  static constexpr int JC = H <= 32 ? H : 32;          // unit chunk of the small-net prologue / epilogue
  static constexpr int JS = (JC % 2) ? JC : JC + 1;    // odd row stride of the warp's transposition buffer
  // the warp's key-table region doubles, once the walk is over, as the exchange buffer [32][PQS] of the final
  // prefix sums (4 NQ floats per trajectory) and as the x0 net's transposition buffer [32][JS]; PQS/4 odd keeps
  // the 16-byte row accesses conflict-free
  static constexpr int PQ0 = (4 * NQ > JC ? 4 * NQ : JC);
  static constexpr int PQ1 = (PQ0 + 3) / 4;
  static constexpr int PQS = 4 * ((PQ1 % 2) ? PQ1 : PQ1 + 1);
  static constexpr int HP = (H + 1 > PQS) ? H + 1 : PQS;  // rows of the region (>= H + 1 keys)
  static constexpr int CS = 33;                        // lane stride of the c table rows (odd: element (j, b) is
                                                       // reached conflict-free with lanes over b AND over j)
  static constexpr int HQ = (H + 3) / 4 * 4;           // padded row length of the staged small-net weights
  static_assert(H <= 512, "key layout holds 9 index bits");
  static_assert(H <= 32 || H % 32 == 0, "hidden widths above 32 must be multiples of 32");
};

constexpr uint32_t kNever = 0x7f800000u;  // +inf: a key that is never due

// Stage the dynamics weights from the torch tensors into the block's table (see Shape):
//   [ bias pairs (NQ) | H unit records ],  record j = [ NQ head-weight pairs | w1t_j | -1/w1t_j | pad ]
// biases and head weights pre-scaled by -log2(e) so that sigmoid(u) = rcp(1 + ex2(v)).
template <int H, int S>
__device__ __forceinline__ void stage_weights(float* __restrict__ wt, const float* __restrict__ w1t, int w1t_stride,
                                              const float* __restrict__ Wg, const float* __restrict__ bg,
                                              const float* __restrict__ Wd, const float* __restrict__ bd) {
  using SH = Shape<H, S>;
  for (int i = threadIdx.x; i < SH::WT; i += kThreads) {
    float v = 0.0f;
    const bool bias = i < SH::BIAS;
    const int j = bias ? 0 : (i - SH::BIAS) / SH::UNIT;
    const int r = bias ? i : (i - SH::BIAS) % SH::UNIT;
    if (r < 2 * SH::NQ) {
      const int q = r >> 1, h = r & 1;
      const bool growth = q < SH::NP;
      const int s = 2 * (growth ? q : q - SH::NP) + h;
      if (s < S) {
        if (bias) v = kNegLog2e * (growth ? bg[s] : bd[s]);
        else v = kNegLog2e * (growth ? Wg[s * H + j] : Wd[s * H + j]);
      }
    } else if (!bias && r == 2 * SH::NQ) {
      v = w1t[(size_t)j * w1t_stride];
    } else if (!bias && r == 2 * SH::NQ + 1) {
      const float w = w1t[(size_t)j * w1t_stride];
      v = (w == 0.0f) ? 0.0f : -1.0f / w;
    }
    wt[i] = v;
  }
}

// per-trajectory tables of the thread: c_j and the sorted crossing keys.  Shared memory, warp-major
// ([warp][row][32 lanes]: conflict-free for any row index) -- or, for wide hidden layers, global scratch
// ([row][all resident threads]: coalesced).
struct Tab {
  float* c;       // element j at c[j * cs]
  uint32_t* k;    // key p at k[p * ks]
  uint8_t* fs;    // BIG only: per-unit status bytes (bit0: gate at the first evaluation, bit1: flipped), stride ks
  int cs, ks;
};

// ---------------------------------------------------------------------------------------------
// piecewise-linear evaluator of one trajectory
// ---------------------------------------------------------------------------------------------
template <int H, int S>
struct Pl {
  using SH = Shape<H, S>;
  static constexpr int NP = SH::NP, NQ = SH::NQ;
  f2 al[NQ], be[NQ];
  float nk;       // next pending key (as float: keys are non-negative floats with the unit index in the low bits)
  int pos;        // its position in the sorted table
  float t_start, dirsign;

  __device__ __forceinline__ static f2 wpair(const float* rec, int q) { return reinterpret_cast<const f2*>(rec)[q]; }

  // one unit at the first evaluation time: dense contribution to (alpha, beta), its key, its gate
  __device__ __forceinline__ uint32_t init_unit(const float* __restrict__ rec, int j, float c, float ts, bool& on) {
    const float w1 = rec[2 * NQ], rinv = rec[2 * NQ + 1];
    on = (__float_as_uint(fmaf(w1, ts, c)) >> 31) == 0u;  // the gate test of the dense evaluation
    const float u = on ? w1 : 0.0f, v = on ? c : 0.0f;
    const f2 uu = bc(u), vv = bc(v);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const f2 w = wpair(rec, q);
      al[q] = fma2(w, uu, al[q]);
      be[q] = fma2(w, vv, be[q]);
    }
    // p_j(t) = w1t_j t + c_j is monotone in t (also in fp32: fma is correctly rounded), so a unit changes state at
    // most once along the sweep, and only if it is not already in the state it has beyond its crossing.
    const bool post = dirsign * w1 > 0.0f;
    const float tx = c * rinv;                               // crossing time -c_j / w1t_j
    const float slack = 4e-7f * (fabsf(tx) + fabsf(ts));
    const float uq = dirsign * (tx - ts);                    // sweep coordinate of the crossing
    const float ub = fmaxf(fmaf(uq, 0.99998474f, -slack), 0.0f);  // biased early: a key is never late
    const bool pending = (rinv != 0.0f) && (on != post) && (ub < 3.0e38f);
    return pending ? ((__float_as_uint(ub) & ~SH::IMASK) | (uint32_t)j) : kNever;
  }

  // first evaluation of the trajectory at time ts; first[] receives the gate pattern (bit j of word j/32)
  __device__ __forceinline__ void init(const float* __restrict__ wt, const Tab& tab, float ts, float dir,
                                       uint32_t (&first)[SH::NWR]) {
    t_start = ts;
    dirsign = dir;
    const f2* bp = reinterpret_cast<const f2*>(wt);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      al[q] = 0ull;
      be[q] = bp[q];
    }
    const float* recs = wt + SH::BIAS;
    if constexpr (!SH::BIG) {
      uint32_t k[SH::N2];
#pragma unroll
      for (int w = 0; w < SH::NWR; ++w) first[w] = 0u;
#pragma unroll
      for (int j = 0; j < SH::N2; ++j) {
        k[j] = kNever;
        if (j < H) {
          bool on;
          k[j] = init_unit(recs + j * SH::UNIT, j, tab.c[j * tab.cs], ts, on);
          first[j >> 5] |= on ? (1u << (j & 31)) : 0u;
        }
      }
      // bitonic sorting network in registers
#pragma unroll
      for (int size = 2; size <= SH::N2; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
          for (int i = 0; i < SH::N2; ++i) {
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
#pragma unroll
      for (int j = 0; j < H; ++j) tab.k[j * tab.ks] = k[j];
      tab.k[H * tab.ks] = kNever;
      nk = __uint_as_float(k[0]);
    } else {
      first[0] = 0u;
      constexpr int NS = H <= 128 ? 128 : (H <= 256 ? 256 : 512);  // power of two >= H
#pragma unroll 2
      for (int j = 0; j < H; ++j) {
        bool on;
        const uint32_t key = init_unit(recs + j * SH::UNIT, j, tab.c[(size_t)j * tab.cs], ts, on);
        tab.k[(size_t)j * tab.ks] = key;
        tab.fs[(size_t)j * tab.ks] = on ? 1 : 0;
      }
      for (int j = H; j <= NS; ++j) tab.k[(size_t)j * tab.ks] = kNever;  // padding + the end sentinel
      // bitonic sort of the thread's own column of the global key table (the table holds NS + 1 rows)
      for (int size = 2; size <= NS; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll 4
          for (int m = 0; m < NS / 2; ++m) {
            const int i = ((m & ~(stride - 1)) << 1) | (m & (stride - 1));  // index with bit `stride` clear
            const int l = i | stride;
            uint32_t* pa = tab.k + (size_t)i * tab.ks;
            uint32_t* pb = tab.k + (size_t)l * tab.ks;
            const uint32_t a = *pa, b = *pb;
            const bool up = (i & size) == 0;
            *pa = up ? min(a, b) : max(a, b);
            *pb = up ? max(a, b) : min(a, b);
          }
        }
      }
      nk = __uint_as_float(tab.k[0]);
    }
    pos = 0;
  }

  // move (alpha, beta) to evaluation time te: every pending crossing whose key is due is confirmed with the exact
  // gate test of the dense evaluation and applied as one rank-one update.  Divergent but short: a unit flips at
  // most once per trajectory and sweep.
  // on_flip(j) is called for every confirmed flip, inside the same divergent trip (reverse sweep: the flip record, where
  // the prefix sums do not change between this seek and the evaluation's accumulation).
  struct NoFlipAction {
    __device__ __forceinline__ void operator()(int) const {}
  };
  __device__ __forceinline__ void seek(const float* __restrict__ wt, const Tab& tab, float te) {
    seek(wt, tab, te, NoFlipAction());
  }
  template <class OnFlip>
  __device__ __forceinline__ void seek(const float* __restrict__ wt, const Tab& tab, float te, OnFlip on_flip) {
    const float uq = dirsign * (te - t_start);
    while (uq >= nk) {
      const int j = (int)(__float_as_uint(nk) & SH::IMASK);
      const float* rec = wt + SH::BIAS + j * SH::UNIT;
      const float w1 = rec[2 * NQ];
      const float c = tab.c[(size_t)j * tab.cs];
      const bool post = dirsign * w1 > 0.0f;
      const bool now = (__float_as_uint(fmaf(w1, te, c)) >> 31) == 0u;
      if (now != post) break;  // not across yet at te (the key is early by construction): stays pending
      const float sgn = post ? 1.0f : -1.0f;
      const f2 uu = bc(sgn * w1), vv = bc(sgn * c);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const f2 w = wpair(rec, q);
        al[q] = fma2(w, uu, al[q]);
        be[q] = fma2(w, vv, be[q]);
      }
      on_flip(j);
      ++pos;
      nk = __uint_as_float(tab.k[(size_t)pos * tab.ks]);
    }
  }

  // G = sigmoid(growth heads), ND = -sigmoid(degradation heads) at time te.
  // MERGE (forward kernel): the XU pipe (16 MUFU lanes per SM) is what the forward comes closest to, so the two
  //   reciprocals of a register pair share one MUFU.RCP:  1/a = b * rcp(ab), 1/b = a * rcp(ab);  the exponentials are
  //   written into swapped halves so that the final packed multiply lands each quotient in its own half.  Exponents
  //   are clamped at 2^60 so ab stays finite (sigmoid floor 1e-18); the clamp propagates NaN.
  // !MERGE (reverse sweep): one MUFU.RCP per sigmoid (rcp(inf) = 0: no clamp needed).  The sweep is bound by
  //   instruction issue with the XU pipe a quarter busy: 6 instructions per pair instead of 10 for 1.33x the MUFU work.
  template <bool MERGE>
  __device__ __forceinline__ void eval(float te, V<NP>& G, V<NP>& ND) const {
    const f2 tt = bc(te);
    const f2 one = bc(1.0f), minus_one = bc(-1.0f);
    float tail_g = 0.0f, tail_d = 0.0f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const bool growth = q < NP;
      const int p = growth ? q : q - NP;
      float v0, v1;
      unpk(fma2(al[q], tt, be[q]), v0, v1);
      if (MERGE) {
        const float e0 = ex2_approx(min_nan(v0, 60.0f));
        if (2 * p + 1 < S) {
          const float e1 = ex2_approx(min_nan(v1, 60.0f));
          const f2 esw = pk(e1, e0);
          const f2 dsw = growth ? add2(esw, one) : sub2(minus_one, esw);
          float d0, d1;
          unpk(dsw, d0, d1);
          const f2 r = bc(rcp_approx(d0 * d1));
          if (growth) G.v[p] = mul2(dsw, r); else ND.v[p] = mul2(dsw, r);
        } else {
          if (growth) tail_g = 1.0f + e0; else tail_d = -1.0f - e0;
        }
      } else {
        const bool full = 2 * p + 1 < S;
        const f2 e = pk(ex2_approx(v0), full ? ex2_approx(v1) : 0.0f);
        float d0, d1;
        unpk(growth ? add2(e, one) : sub2(minus_one, e), d0, d1);
        const f2 r = pk(rcp_approx(d0), full ? rcp_approx(d1) : 0.0f);
        if (growth) G.v[p] = r; else ND.v[p] = r;
      }
    }
    if constexpr (MERGE && (S & 1) != 0) {
      const float r = rcp_approx(tail_g * tail_d);
      G.v[NP - 1] = pk(tail_d * r, 0.0f);
      ND.v[NP - 1] = pk(tail_g * r, 0.0f);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// warp reductions (as in round 1): sum K values over the 32 lanes with ~K shuffles
// ---------------------------------------------------------------------------------------------
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
// the two small nets in front of the solve, weights staged in shared memory:
//   Wz [L][HQ] = W1[:,1:]^T,  Wa [L][HQ] = latent_to_ode_net[0].weight^T,  b1 [HQ], ba [HQ],  Wb [S][HQ], bb [S]
// ---------------------------------------------------------------------------------------------
template <int H, int S>
struct LatSmem {
  float *Wz, *Wa, *b1, *ba, *Wb, *bb;
  static constexpr int HQ = Shape<H, S>::HQ;
  __host__ __device__ static int floats(int L) { return 2 * L * HQ + 2 * HQ + S * HQ + (S + 3) / 4 * 4; }
  __device__ __forceinline__ void stage(float* base, const LatentSrc& lat) {  // caller syncs afterwards
    const int L = lat.L;
    Wz = base;
    Wa = Wz + L * HQ;
    b1 = Wa + L * HQ;
    ba = b1 + HQ;
    Wb = ba + HQ;
    bb = Wb + S * HQ;
    const bool fx0 = lat.Wa != nullptr;
    for (int i = threadIdx.x; i < L * HQ; i += kThreads) {
      const int l = i / HQ, j = i % HQ;
      Wz[i] = j < H ? lat.W1[j * (L + 1) + 1 + l] : 0.0f;
      Wa[i] = (fx0 && j < H) ? lat.Wa[j * L + l] : 0.0f;
    }
    for (int i = threadIdx.x; i < HQ; i += kThreads) {
      b1[i] = i < H ? lat.b1[i] : 0.0f;
      ba[i] = (fx0 && i < H) ? lat.ba[i] : 0.0f;
    }
    for (int i = threadIdx.x; i < S * HQ; i += kThreads) {
      const int s = i / HQ, j = i % HQ;
      Wb[i] = (fx0 && j < H) ? lat.Wb[s * H + j] : 0.0f;
    }
    for (int i = threadIdx.x; i < S; i += kThreads) bb[i] = fx0 ? lat.bb[i] : 0.0f;
  }
};

// The warp's 32 latent rows are ONE contiguous run of 32*L floats of z: they are copied into shared memory with
// coalesced loads once per tile (zT[row][LQ], LQ = L rounded up to four, padding zero).  Read row by row straight
// from global memory (first version) every load of the L-step loop was an uncoalesced L2 round trip one short
// iteration ahead of its use: 7 % of the forward's and 5 % of the reverse sweep's stall samples.
__device__ __forceinline__ void stage_z_rows(float* __restrict__ zT, int LQ, const float* __restrict__ z, int L,
                                             int64_t row0, int64_t B, int lane) {
  __syncwarp();   // every lane is done with the rows of the tile before
  const int n = 32 * L;
  for (int idx = lane; idx < n; idx += 32) {
    const int r = idx / L, col = idx - r * L;
    const int64_t gr = min(row0 + r, B - 1);   // tail rows repeat trajectory B-1 (their stores are masked off)
    zT[r * LQ + col] = __ldg(z + gr * L + col);
  }
  for (int col = L; col < LQ; ++col) zT[lane * LQ + col] = 0.0f;
  __syncwarp();
}

// acc[j0 .. j0+JC) = bias + sum_l W[l][j] z_l  for one chunk of units; zrow = this thread's staged latent row (16-byte
// aligned, zero-padded to a multiple of four).  Two units per instruction: the staged weight rows are zero-padded to
// a multiple of four floats and 16-byte aligned, so a row is read as 16-byte vectors of two weight pairs each.
// ZS = false: zrow is the row in global memory (the widest layer's forward has no shared memory left for the rows).
template <int JC, bool ZS = true>
__device__ __forceinline__ void lat_chunk(const float* __restrict__ W, const float* __restrict__ bias, int HQ, int j0,
                                          const float* __restrict__ zrow, int L, float (&acc)[JC]) {
  constexpr int J4 = (JC + 3) / 4;  // 16-byte vectors per row
  f2 a2[2 * J4];
  const ulonglong2* b4 = reinterpret_cast<const ulonglong2*>(bias + j0);
#pragma unroll
  for (int q = 0; q < J4; ++q) {
    const ulonglong2 v = b4[q];
    a2[2 * q] = v.x;
    a2[2 * q + 1] = v.y;
  }
#pragma unroll 1
  for (int l0 = 0; l0 < L; l0 += 4) {
    float zz[4];
    if (ZS) {
      const float4 z4 = *reinterpret_cast<const float4*>(zrow + l0);
      zz[0] = z4.x; zz[1] = z4.y; zz[2] = z4.z; zz[3] = z4.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) zz[k] = l0 + k < L ? __ldg(zrow + l0 + k) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (l0 + k < L) {
        const f2 zl = bc(zz[k]);
        const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(W + (l0 + k) * HQ + j0);
#pragma unroll
        for (int q = 0; q < J4; ++q) {
          const ulonglong2 v = w4[q];
          a2[2 * q] = fma2(v.x, zl, a2[2 * q]);
          a2[2 * q + 1] = fma2(v.y, zl, a2[2 * q + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < JC; ++j) acc[j] = (j & 1) ? hi_of(a2[j >> 1]) : lo_of(a2[j >> 1]);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int kStageT = 4;  // output times staged per trajectory for (B,T,S)-contiguous storage
constexpr int kHeadRing = 8;  // fused decoder heads: states kept per trajectory = floats of one 32-byte sector of a mu row
constexpr int kHeadRows = 24; // at most NQ * O = 3 * 8 (head, output) rows
constexpr int kHeadTab = 2 * kHeadRows + 8;  // floats: row offsets (int64 each) + due masks by sector phase

// one 32-byte store (STG.256): a whole, aligned sector of a head-output row in ONE request
__device__ __forceinline__ void st_sector(float* dst, const float (&a)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(a[0]), "f"(a[1]), "f"(a[2]),
               "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7])
               : "memory");
}

// shared-memory floats of the per-trajectory tables of a block (narrow layers): c [warp][H][CS], keys [warp][HP][32]
template <int H, int S>
__host__ __device__ constexpr size_t table_floats() {
  return (size_t)kWarps * H * Shape<H, S>::CS + (size_t)kThreads * Shape<H, S>::HP;
}

// the forward keeps the warps' latent rows in shared memory unless the staged weights of the widest layer leave no
// room for them next to a second resident block
template <int H> __host__ __device__ constexpr bool fwd_stages_z() { return H <= 256; }

// fused decoder heads: floats per thread of the state ring's region (also holds the thread's staged z row before the
// time loop starts), and the floats of the time grid staged per block (grids above kHeadGridMax stay in global memory)
constexpr int kHeadGridMax = 1024;
template <int H, int S> __host__ __device__ constexpr int head_ring_floats(int L, bool lat) {
  const int LQ = (L + 3) / 4 * 4;
  return (lat && fwd_stages_z<H>() && LQ > kHeadRing * S) ? LQ : kHeadRing * S;
}
__host__ __device__ constexpr int head_grid_floats(int T) { return T <= kHeadGridMax ? (T + 3) / 4 * 4 : 0; }

template <int H, int S>
__host__ __device__ constexpr size_t fwd_smem_bytes(int L, bool lat, bool rows_in_time, bool heads = false, int T = 0) {
  using SH = Shape<H, S>;
  size_t n = SH::WT;
  if (!SH::BIG) n += table_floats<H, S>();
  if (lat) n += LatSmem<H, S>::floats(L);                                  // staged nets
  if (lat && fwd_stages_z<H>() && !heads) n += (size_t)kThreads * ((L + 3) / 4 * 4);   // + the warps' z rows
  if (rows_in_time && !heads) n += (size_t)kThreads * kStageT * S;                    // staged output rows
  if (heads) {
    // state ring (its region doubles as the warp's z rows during the prologue) + head weights + row tables + time grid
    n += (size_t)kThreads * head_ring_floats<H, S>(L, lat) + kMaxHeadW + kHeadTab + head_grid_floats(T);
  }
  return n * sizeof(float);
}

// bytes of global scratch per resident thread (wide hidden layers only)
template <int H, int S>
__host__ __device__ constexpr size_t big_tab_bytes() {
  constexpr int NS = H <= 128 ? 128 : (H <= 256 ? 256 : 512);
  return Shape<H, S>::BIG ? (size_t)H * 4 + (size_t)(NS + 1) * 4 + (size_t)H : 0;
}

template <int H, int S>
__device__ __forceinline__ Tab make_tab(float* smem_tables, unsigned char* ws, int warp, int lane) {
  using SH = Shape<H, S>;
  Tab tab;
  if constexpr (!SH::BIG) {
    tab.c = smem_tables + (size_t)warp * H * SH::CS + lane;
    tab.k = reinterpret_cast<uint32_t*>(smem_tables + (size_t)kWarps * H * SH::CS) + (size_t)warp * SH::HP * 32 + lane;
    tab.fs = nullptr;
    tab.cs = SH::CS;
    tab.ks = 32;
  } else {
    constexpr int NS = H <= 128 ? 128 : (H <= 256 ? 256 : 512);
    const size_t nt = (size_t)gridDim.x * kThreads, gt = (size_t)blockIdx.x * kThreads + threadIdx.x;
    tab.c = reinterpret_cast<float*>(ws) + gt;
    tab.k = reinterpret_cast<uint32_t*>(ws + nt * H * 4) + gt;
    tab.fs = ws + nt * H * 4 + nt * (NS + 1) * 4 + gt;
    tab.cs = tab.ks = (int)nt;
  }
  return tab;
}

// c_j into the thread's table (from z through the staged W1[:,1:], or from the given (B,H) array); returns x0
template <int H, int S, bool WANT_X0, bool ZS = true>
__device__ __forceinline__ void prologue(const LatSmem<H, S>& ls, const LatentSrc& lat, const float* __restrict__ cin,
                                         const float* __restrict__ zrow, int64_t b, const Tab& tab,
                                         V<(S + 1) / 2>& x0) {
  using SH = Shape<H, S>;
  constexpr int JC = SH::JC;
  if (lat.z) {
    float xa[S];
    if (WANT_X0) {
#pragma unroll
      for (int s = 0; s < S; ++s) xa[s] = ls.bb[s];
    }
#pragma unroll 1
    for (int j0 = 0; j0 < H; j0 += JC) {
      float acc[JC];
      lat_chunk<JC, ZS>(ls.Wz, ls.b1, SH::HQ, j0, zrow, lat.L, acc);
#pragma unroll
      for (int j = 0; j < JC; ++j) tab.c[(size_t)(j0 + j) * tab.cs] = acc[j];
      if (WANT_X0) {
        lat_chunk<JC, ZS>(ls.Wa, ls.ba, SH::HQ, j0, zrow, lat.L, acc);
#pragma unroll
        for (int j = 0; j < JC; ++j) {
          const float h = relu_nan(acc[j]);
#pragma unroll
          for (int s = 0; s < S; ++s) xa[s] = fmaf(ls.Wb[s * SH::HQ + j0 + j], h, xa[s]);
        }
      }
    }
    if (WANT_X0) {
#pragma unroll
      for (int p = 0; p < SH::NP; ++p) {
        const float a = sigmoid_from_scaled(kNegLog2e * xa[2 * p]);
        const float bq = (2 * p + 1 < S) ? sigmoid_from_scaled(kNegLog2e * xa[(2 * p + 1 < S) ? 2 * p + 1 : 0]) : 0.0f;
        x0.v[p] = pk(a, bq);
      }
    }
  } else {
    const float* crow = cin + b * H;
#pragma unroll 4
    for (int j = 0; j < H; ++j) tab.c[(size_t)j * tab.cs] = ld_stream(crow + j);
  }
}

// HEADS: the decoder heads are applied to the states while the thread still holds them and written in the reference's
// (B, obs_dim, T) layout (hd.mu, (NQ,B,O,T)); sol is then optional (null: never written).  A mu row runs along TIME, a
// thread produces one time point per step, and a row's pitch (T floats) puts its 32-byte sectors at a phase that
// depends on (trajectory, head, output).  Writing less than a whole sector per request is what must not happen: the
// first version stored 16 bytes per row every four steps and the L2 filled every sector from DRAM for the partial
// write and wrote most of them back twice (ncu: 3.5 GB read + 6.7 GB written for 3.8 GB of output, 6.5 ms).  So the
// thread keeps its last kHeadRing = 8 states in a shared-memory ring (warp-major, conflict-free) and, at the step
// where a row's current sector becomes complete, computes the sector's eight outputs and stores them with ONE 32-byte
// store; only a row's first and last sectors can be partial.  The dot products run in the order of heads_fwd_kernel
// (slode_heads.cu), so fused and two-kernel results are bit-equal.
template <int H, int S, int METHOD, bool HEADS = false>
__global__ void __launch_bounds__(kThreads, (Shape<H, S>::BIG || S > 5) ? 3 : (HEADS ? 4 : SLODE_FX_FWD_MINB))
fixed_fwd_kernel(int64_t B, int T, const float* __restrict__ tgrid, const float* __restrict__ cin,
                 const float* __restrict__ y0, float* __restrict__ sol, int64_t st, int64_t sb, PackSrc w,
                 int w1t_stride, LatentSrc lat, unsigned char* __restrict__ ws, HeadsSrc hd) {
  using SH = Shape<H, S>;
  constexpr int NP = SH::NP;
  extern __shared__ __align__(16) float fx_smem[];
  float* const wt = fx_smem;
  float* const tables = wt + SH::WT;
  float* lat_base = tables + (SH::BIG ? 0 : table_floats<H, S>());
  const bool rows_in_time = !HEADS && (st == S);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  LatSmem<H, S> ls{};
  stage_weights<H, S>(wt, w.w1t, w1t_stride, w.Wg, w.bg, w.Wd, w.bd);
  if (lat.z) ls.stage(lat_base, lat);
  const int LQ = (lat.L + 3) / 4 * 4;
  // this warp's z rows (lat.z only); HEADS: inside the warp's state-ring region, which is idle until the first output
  const int ring_w = 32 * head_ring_floats<H, S>(lat.L, lat.z != nullptr);   // floats of a warp's ring region
  float* const zT = lat_base + LatSmem<H, S>::floats(lat.L) + (size_t)warp * (HEADS ? ring_w : 32 * LQ);
  constexpr bool kZS = fwd_stages_z<H>();
  float* const ostage0 =
      lat_base + (lat.z ? LatSmem<H, S>::floats(lat.L) + ((kZS && !HEADS) ? (size_t)kThreads * LQ : 0) : 0);
  float* const ostage = ostage0 + (size_t)tid * kStageT * S;
  // HEADS: ring[(slot * S + s) * 32] of this lane, then the stacked head weights (NQ,O,S), then two small tables:
  //   roff[r]  float offset of row r = (q,o) from row (0,0) of the same trajectory: (q B O + o) P, P = hd.pitch >= T the
  //            row pitch (with P % 8 == 0 every row of the launch has the same sector phase: all lanes flush all
  //            their rows at the same steps, k % 8 == 7, and nothing diverges)
  //   due[c]   bit r set: row r completes a 32-byte sector at a step with ((e_b + k) & 7) == c
  float* const ring = ostage0 + (size_t)warp * ring_w + lane;
  float* const hw = ostage0 + (size_t)kWarps * ring_w;
  int64_t* const roff = reinterpret_cast<int64_t*>(hw + kMaxHeadW);   // 8-byte aligned: every region is a multiple of 4 floats
  uint32_t* const due = reinterpret_cast<uint32_t*>(roff + kHeadRows);
  // the time grid: the scattered sector stores queue in front of every global load (a grid element fetched one step
  // ahead cost 10 % of the fused kernel's stall samples), so the loop reads it from shared memory
  float* const sgrid = reinterpret_cast<float*>(due + 8);
  const float* const tg = (HEADS && T <= kHeadGridMax) ? sgrid : tgrid;
  if (HEADS && T <= kHeadGridMax) {
    for (int i = tid; i < T; i += kThreads) sgrid[i] = __ldg(tgrid + i);
  }
  const int nqo = HEADS ? hd.NQ * hd.O : 0;
  if (HEADS) {
    for (int i = tid; i < nqo * S; i += kThreads) hw[i] = hd.W[i];
    if (tid < nqo) roff[tid] = ((int64_t)(tid / hd.O) * B * hd.O + (tid % hd.O)) * hd.pitch;
    if (tid >= 32 && tid < 40) {
      const int c = tid - 32;
      uint32_t m = 0;
      for (int r = 0; r < nqo; ++r) {
        const int64_t e = ((int64_t)(r / hd.O) * B * hd.O + (r % hd.O)) * hd.pitch;
        if (((c + (int)(e & 7)) & 7) == 7) m |= 1u << r;
      }
      due[c] = m;
    }
  }
  __syncthreads();
  const Tab tab = make_tab<H, S>(tables, ws, warp, lane);
  const float dir = (T < 2 || __ldg(tgrid + T - 1) >= __ldg(tgrid)) ? 1.0f : -1.0f;
  const int64_t ntiles = (B + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t br = tile * kThreads + tid;
    const bool ok = br < B;
    const int64_t b = ok ? br : B - 1;  // tail threads redo trajectory B-1 with their stores masked off
    V<NP> x;
    if (lat.z && kZS) stage_z_rows(zT, LQ, lat.z, lat.L, tile * kThreads + warp * 32, B, lane);
    const float* const zrow = kZS ? zT + lane * LQ : lat.z + b * lat.L;
    if (lat.z && lat.Wa) {
      prologue<H, S, true, kZS>(ls, lat, cin, zrow, b, tab, x);
    } else {
      prologue<H, S, false, kZS>(ls, lat, cin, zrow, b, tab, x);
      x = vload<S>(y0 + b * S);
    }
    float* out = sol + b * sb;
    float* const row = out;
    int hd_eb = 0;
    if (HEADS)
      hd_eb = (int)((((b & 7) * (hd.O & 7) * (hd.pitch & 7)) + ((reinterpret_cast<uintptr_t>(hd.mu) >> 2) & 7)) & 7);

    auto put = [&](int k, const V<NP>& xv) {
      if (HEADS) {
        if (sol) vstore<S>(out, ok, xv);
        const int slot = k & (kHeadRing - 1);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          float lo, hi;
          unpk(xv.v[p], lo, hi);
          ring[(slot * S + 2 * p) * 32] = lo;
          if (2 * p + 1 < S) ring[(slot * S + 2 * p + 1) * 32] = hi;
        }
        const bool last = (k == T - 1);
        const int base = hd_eb + k;   // sector phase of time k in row (0,0) of this trajectory
        uint32_t todo = last ? (nqo >= 32 ? 0xffffffffu : (1u << nqo) - 1u) : due[base & 7];
        if (!ok) todo = 0;
        if (todo) {
          // the last eight states, oldest first, as four pairs of consecutive times: state at time k - 7 + j is
          // half j & 1 of xs[j >> 1] (garbage before time 0: never stored)
          V<S> xs[kHeadRing / 2];   // V<S>: S register pairs
#pragma unroll
          for (int m = 0; m < kHeadRing / 2; ++m) {
            const int sa = (k + 1 + 2 * m) & (kHeadRing - 1), sb2 = (k + 2 + 2 * m) & (kHeadRing - 1);
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) xs[m].v[s2] = pk(ring[(sa * S + s2) * 32], ring[(sb2 * S + s2) * 32]);
          }
          float* const mrow = hd.mu + b * (hd.O * hd.pitch) + k;   // time k of row (0,0)
          do {   // this lane's due rows; lanes with fewer of them idle in the last trips
            const int r = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t ro = roff[r];
            const int ph = (base + (int)(ro & 7)) & 7;   // position of time k inside its sector
            const float* wr = hw + r * S;
            f2 acc[kHeadRing / 2];
#pragma unroll
            for (int m = 0; m < kHeadRing / 2; ++m) acc[m] = bc(0.0f);
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) {   // same order as heads_fwd_kernel: acc = fma(w_s, x_s, acc), s ascending
              const f2 wv = bc(wr[s2]);
#pragma unroll
              for (int m = 0; m < kHeadRing / 2; ++m) acc[m] = fma2(wv, xs[m].v[s2], acc[m]);
            }
            float a[kHeadRing];
#pragma unroll
            for (int m = 0; m < kHeadRing / 2; ++m) unpk(acc[m], a[2 * m], a[2 * m + 1]);
            float* const dk = mrow + ro;   // &mu[q][b][o][k]
            if (ph == 7 && k >= 7) {
              st_sector(dk - 7, a);
            } else {
              const int n = min(ph, k) + 1;   // times k - n + 1 .. k are pending
#pragma unroll
              for (int j = 0; j < kHeadRing; ++j)
                if (j >= kHeadRing - n) dk[j - (kHeadRing - 1)] = a[j];
            }
          } while (todo);
        }
        return;
      }
      if (!rows_in_time) {
        vstore<S>(out, ok, xv);
        return;
      }
      // (B,T,S)-contiguous storage: a trajectory's rows are contiguous in time -> stage kStageT of them and write
      // one run of kStageT*S floats with 16-byte stores
      const int slot = k & (kStageT - 1);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        float lo, hi;
        unpk(xv.v[p], lo, hi);
        ostage[slot * S + 2 * p] = lo;
        if (2 * p + 1 < S) ostage[slot * S + 2 * p + 1] = hi;
      }
      if ((slot == kStageT - 1 || k == T - 1) && ok) {
        const int n = (slot + 1) * S;
        float* dst = row + (int64_t)(k - slot) * S;
        if (n == kStageT * S && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (kStageT * S) % 4 == 0) {
#pragma unroll
          for (int q = 0; q < kStageT * S / 4; ++q)
            reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(ostage)[q];
        } else {
          for (int q = 0; q < n; ++q) dst[q] = ostage[q];
        }
      }
    };

    if (HEADS) __syncwarp();   // every lane of the warp has read its z row: the region becomes the state ring
    put(0, x);
    float t0 = HEADS ? tg[0] : __ldg(tgrid);
    Pl<H, S> pl;
    uint32_t first[SH::NWR];
    if (T > 1) pl.init(wt, tab, t0, dir, first);
    V<NP> k1;
    if (METHOD == SLODE_METHOD_RK4 && T > 1) {  // k1 of the first step; afterwards carried over from the step before
      V<NP> G, D;
      pl.template eval<true>(t0, G, D);
      k1 = rhs<NP>(G, D, x);
    }
    float t_ahead = T > 1 ? (HEADS ? tg[1] : __ldg(tgrid + 1)) : t0;  // the grid is read one step ahead of its use
#pragma unroll 1
    for (int i = 0; i + 1 < T; ++i) {
      const float t1 = t_ahead;
      if (i + 2 < T) t_ahead = HEADS ? tg[i + 2] : ld_early(tgrid + i + 2);
      const float dt = t1 - t0;
      if (METHOD == SLODE_METHOD_EULER) {
        V<NP> G, D;
        pl.seek(wt, tab, t0);
        pl.template eval<true>(t0, G, D);
        x = vaxpy<NP>(dt, rhs<NP>(G, D, x), x);
      } else if (METHOD == SLODE_METHOD_MIDPOINT) {
        const float half_dt = 0.5f * dt;
        V<NP> G, D;
        pl.seek(wt, tab, t0);
        pl.template eval<true>(t0, G, D);
        const V<NP> ym = vaxpy<NP>(half_dt, rhs<NP>(G, D, x), x);
        const float tm = t0 + half_dt;
        pl.seek(wt, tab, tm);
        pl.template eval<true>(tm, G, D);
        x = vaxpy<NP>(dt, rhs<NP>(G, D, ym), x);
      } else {  // rk4, 3/8 rule (torchdiffeq rk4_alt_step_func); the evaluation at t1 is the next step's k1
        V<NP> G, D;
        const float ta = t0 + dt * kOneThird, tb = t0 + dt * kTwoThirds;
        V<NP> y = vaxpy<NP>(dt * kOneThird, k1, x);
        pl.seek(wt, tab, ta);
        pl.template eval<true>(ta, G, D);
        const V<NP> k2 = rhs<NP>(G, D, y);
        y = vaxpy<NP>(dt, vaxpy<NP>(-kOneThird, k1, k2), x);
        pl.seek(wt, tab, tb);
        pl.template eval<true>(tb, G, D);
        const V<NP> k3 = rhs<NP>(G, D, y);
        y = vaxpy<NP>(dt, vadd<NP>(vsub<NP>(k1, k2), k3), x);
        pl.seek(wt, tab, t1);
        pl.template eval<true>(t1, G, D);
        const V<NP> k4 = rhs<NP>(G, D, y);
        x = vaxpy<NP>(dt * 0.125f, vadd<NP>(vaxpy<NP>(3.0f, vadd<NP>(k2, k3), k1), k4), x);
        k1 = rhs<NP>(G, D, x);
      }
      out += st;
      put(i + 1, x);
      t0 = t1;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// Flat layout of the block's gradient accumulators == layout of grad_params (slode_b200.h):
//   [ dw1t (H) | dWg (S*H) | dbg (S) | dWd (S*H) | dbd (S) | dW1z (H*L) | db1 (H) | dWa (H*L) | dba (H) | dWb (S*H) | dbb (S) ]
template <int H, int S>
struct GradLayout {
  static constexpr int w1t = 0, Wg = H, bg = H + S * H, Wd = H + S * H + S, bd = H + 2 * S * H + S;
  static constexpr int base = H + 2 * (S * H + S);
  __host__ __device__ static int W1z(int) { return base; }
  __host__ __device__ static int b1(int L) { return base + H * L; }
  __host__ __device__ static int Wa(int L) { return base + H * L + H; }
  __host__ __device__ static int ba(int L) { return base + 2 * H * L + H; }
  __host__ __device__ static int Wb(int L) { return base + 2 * H * L + 2 * H; }
  __host__ __device__ static int bb(int L) { return base + 2 * H * L + 2 * H + S * H; }
  __host__ __device__ static int total(int L, bool lat, bool fx0) {
    return base + (lat ? H * L + H : 0) + (fx0 ? H * L + H + S * H + S : 0);
  }
};

// flip records: per resident thread and hidden unit one snapshot of (P, Q) = 2 * NQ register pairs, stored
// [thread][unit][q] -> (P[q], Q[q]): a record is one contiguous run written with immediate offsets from a single
// address (the write sits in a divergent trip that usually serves one lane); the end-of-sweep pass reads a unit's
// records of all lanes through L1
template <int H, int S>
__host__ __device__ constexpr size_t rec_bytes_per_thread() {
  return (size_t)H * Shape<H, S>::NQ * 16;
}

template <int H, int S>
struct Sweep {
  using SH = Shape<H, S>;
  static constexpr int NP = SH::NP, NQ = SH::NQ;
  f2 P[NQ], Q[NQ];
  uint32_t flipped[SH::NWR];

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int q = 0; q < NQ; ++q) P[q] = Q[q] = 0ull;
#pragma unroll
    for (int w = 0; w < SH::NWR; ++w) flipped[w] = 0u;
  }

  // the units whose keys sit at positions [p0, p1) of the sorted table flipped between the previous contributing
  // evaluation and the next one: snapshot the prefix sums as they stand into each unit's record
  __device__ __forceinline__ void events(f2* __restrict__ rec, const Tab& tab, int p0, int p1) {
    for (int p = p0; p < p1; ++p) record(rec, tab, (int)(tab.k[(size_t)p * tab.ks] & SH::IMASK));
  }

  // snapshot of the prefix sums as they stand into unit j's record
  __device__ __forceinline__ void record(f2* __restrict__ rec, const Tab& tab, int j) {
    // record: [q] -> (P[q], Q[q]).  Two 8-byte stores per q: one 16-byte store would tie P[q] and Q[q] to an aligned
    // register quad for the whole sweep and the time loop would end in ~24 register copies per interval
    f2* dst = rec + j * (2 * NQ);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      dst[2 * q] = P[q];
      dst[2 * q + 1] = Q[q];
    }
    if constexpr (SH::BIG) {
      tab.fs[(size_t)j * tab.ks] |= 2;
    } else {
#pragma unroll
      for (int w = 0; w < SH::NWR; ++w) {
        if (w == (j >> 5)) flipped[w] |= 1u << (j & 31);
      }
    }
  }

  // add the cotangents of the head pre-activations of one evaluation at time te:
  //   f = G + ND*y with upstream gf and gy = gf*ND (= dL/dy through this evaluation, which the caller needs anyway):
  //   d(pre_G) = gf*(G - G^2),   d(pre_D) = gf*y*(ND^2 + ND) = gy*(y + ND*y)          (ND = -D)
  // 8 packed instructions per state pair (the first version spent 10)
  __device__ __forceinline__ void add(float te, const V<NP>& gf, const V<NP>& gy, const V<NP>& y, const V<NP>& G,
                                      const V<NP>& ND) {
    const f2 tt = bc(te);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const f2 dg = mul2(gf.v[p], fma2(G.v[p], neg2(G.v[p]), G.v[p]));
      const f2 dd = mul2(gy.v[p], fma2(ND.v[p], y.v[p], y.v[p]));
      P[p] = add2(P[p], dg);
      Q[p] = fma2(dg, tt, Q[p]);
      P[NP + p] = add2(P[NP + p], dd);
      Q[NP + p] = fma2(dd, tt, Q[NP + p]);
    }
  }
};

template <int H, int S>
__host__ __device__ constexpr size_t bwd_smem_bytes(int L, bool lat) {
  using SH = Shape<H, S>;
  size_t n = SH::WT;
  if (!SH::BIG) n += table_floats<H, S>();
  else n += (size_t)kThreads * SH::PQS;              // the warps' exchange buffers
  if (!SH::BIG) n += (size_t)(GradLayout<H, S>::total(L, lat, lat) + 3) / 4 * 4;  // block accumulators
  if (lat) n += LatSmem<H, S>::floats(L) + (size_t)kThreads * ((L + 3) / 4 * 4);  // staged nets + the warps' z rows
  n += (size_t)2 * SLODE_FX_ROW_AHEAD * S * kThreads;  // the threads' state / cotangent rows of the intervals in flight (cp.async targets)
  return n * sizeof(float);
}

// resident blocks per SM the reverse sweep is compiled for (each value measured on the B200 at 2^20 x 100): wide
// layers 2; S > 5 (four register pairs per vector): rk4 2, else 3; S <= 5: rk4 and the DISCRETE midpoint sweep (two
// evaluations and two cotangent accumulations live at once: 3.67 ms at 128 registers, 2.70 at 168) 3, the rest 4
template <int H, int S, int METHOD, int MODE>
__host__ __device__ constexpr int bwd_min_blocks() {
  if (Shape<H, S>::BIG) return 2;
  if (S > 5) return METHOD == SLODE_METHOD_RK4 ? 2 : SLODE_FX_BWD_MINB_WIDE;
  if (METHOD == SLODE_METHOD_RK4) return SLODE_FX_BWD_MINB_RK4;
  if (METHOD == SLODE_METHOD_MIDPOINT && MODE == SLODE_BWD_DISCRETE) return SLODE_FX_BWD_MINB_RK4;
  return SLODE_FX_BWD_MINB;
}

template <int H, int S, int METHOD, int MODE>
__global__ void __launch_bounds__(kThreads, bwd_min_blocks<H, S, METHOD, MODE>())
fixed_bwd_kernel(int64_t B, int T, const float* __restrict__ tgrid, const float* __restrict__ cin,
                 const float* __restrict__ sol, int64_t st, int64_t sb, const float* __restrict__ gsol, int64_t gst,
                 int64_t gsb, float* __restrict__ grad_y0, float* __restrict__ grad_c, float* __restrict__ grad_w,
                 PackSrc w, int w1t_stride, LatentSrc lat, float* __restrict__ grad_z,
                 unsigned char* __restrict__ ws) {
  using SH = Shape<H, S>;
  using GL = GradLayout<H, S>;
  constexpr int NP = SH::NP, NQ = SH::NQ;
  extern __shared__ __align__(16) float fx_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = lat.L;
  const bool fused = lat.z != nullptr, fx0 = fused && lat.Wa != nullptr;
  float* const wt = fx_smem;
  float* const tables = wt + SH::WT;
  float* p_ = tables + (SH::BIG ? (size_t)kThreads * SH::PQS : table_floats<H, S>());
  // parameter-gradient sums: block accumulators in shared memory, flushed once at the end -- except for wide layers,
  // whose 25k-100k sums do not fit next to the staged weights and are added to grad_params directly
  float* const acc = SH::BIG ? grad_w : p_;
  const int n_acc = SH::BIG ? 0 : GL::total(L, fused, fx0);
  if (!SH::BIG) p_ += (GL::total(L, fused, fused) + 3) / 4 * 4;  // keeps the float4 rows behind it 16-byte aligned
  LatSmem<H, S> ls{};
  float* zT = nullptr;  // this warp's z rows, [32][LQ]
  const int LQ = (L + 3) / 4 * 4;
  if (fused) {
    ls.stage(p_, lat);
    zT = p_ + LatSmem<H, S>::floats(L) + (size_t)warp * 32 * LQ;
    p_ += LatSmem<H, S>::floats(L) + (size_t)kThreads * LQ;
  }
  constexpr int kAhead = SLODE_FX_ROW_AHEAD;
  static_assert(kAhead == 1 || kAhead == 2, "row slots are addressed with i & (kAhead - 1)");
  float* const rowx = p_ + tid;                          // this thread's state rows of the intervals in flight: slot i % kAhead
  float* const rowg = p_ + kAhead * S * kThreads + tid;  // ... and its upstream-gradient rows
  stage_weights<H, S>(wt, w.w1t, w1t_stride, w.Wg, w.bg, w.Wd, w.bd);
  for (int i = tid; i < n_acc; i += kThreads) acc[i] = 0.0f;
  __syncthreads();
  const size_t nthreads = (size_t)gridDim.x * kThreads, gthread = (size_t)blockIdx.x * kThreads + tid;
  const Tab tab = make_tab<H, S>(tables, ws, warp, lane);
  f2* const rec = reinterpret_cast<f2*>(ws + big_tab_bytes<H, S>() * nthreads) + gthread * (size_t)(H * 2 * NQ);
  // the warp's exchange buffer [32][PQS]: the key table's rows once the walk is over (narrow layers), else its own
  float* const pqw = SH::BIG ? tables + (size_t)warp * 32 * SH::PQS
                             : tables + (size_t)kWarps * H * SH::CS + (size_t)warp * SH::HP * 32;
  // the reverse sweep visits the grid from its last time to its first
  const float dir = (T < 2 || __ldg(tgrid + T - 1) >= __ldg(tgrid)) ? -1.0f : 1.0f;
  const int64_t ntiles = (B + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t br = tile * kThreads + tid;
    const bool ok = br < B;
    const int64_t b = ok ? br : B - 1;
    if (fused) stage_z_rows(zT, LQ, lat.z, L, tile * kThreads + warp * 32, B, lane);
    {
      V<NP> dummy;
      prologue<H, S, false>(ls, lat, cin, zT + lane * LQ, b, tab, dummy);
    }
    const float* xs = sol + b * sb;
    const float* gs = gsol + b * gsb;
    if (fused && tile + gridDim.x < ntiles) {  // the next tile's latent row: in L2 by the time its prologue reads it
      const int64_t bn = min(br + (int64_t)gridDim.x * kThreads, B - 1);
      prefetch_l2(lat.z + bn * L);
      prefetch_l2(lat.z + bn * L + L - 1);
    }
    const float live = ok ? 1.0f : 0.0f;   // a masked-off thread carries zero cotangents: it only adds zeros (its
                                           // upstream-gradient rows are zero-filled by the row fetches)
    V<NP> lam = vscale<NP>(vload<S>(gs + (int64_t)(T - 1) * gst), live);

    Sweep<H, S> sw;
    sw.init();
    Pl<H, S> pl;
    uint32_t first[SH::NWR];
#pragma unroll
    for (int ww = 0; ww < SH::NWR; ++ww) first[ww] = 0u;
    float t1 = __ldg(tgrid + T - 1);
    V<NP> Gc, Dc;  // rk4: the evaluation at t1, carried over from the interval processed before
    int pdone = 0;  // keys [0, pdone) have had their records written
    if (T > 1) {
      // first evaluation time of the sweep
      float tfirst = t1;
      if (MODE == SLODE_BWD_DISCRETE) {
        const float tp = __ldg(tgrid + T - 2);
        if (METHOD == SLODE_METHOD_EULER) tfirst = tp;
        if (METHOD == SLODE_METHOD_MIDPOINT) tfirst = tp + 0.5f * (t1 - tp);
      }
      pl.init(wt, tab, tfirst, dir, first);
      if (METHOD == SLODE_METHOD_RK4) pl.template eval<false>(t1, Gc, Dc);
    }

    float t_ahead = T > 1 ? __ldg(tgrid + T - 2) : t1;  // the grid is read one interval ahead of its use
    // The state row the interval needs (sol[i]; sol[i+1] for the odeint_adjoint restart) and the cotangent row of
    // grid point i are fetched ONE INTERVAL AHEAD with cp.async into the thread's own shared-memory column: read at
    // the point of use they cost a full L2 round trip per interval (12 % of the sweep's stall samples;
    // prefetch.global.L1 does not help the read-only path), forced early into registers they occupy 12 registers
    // for a whole interval and make the evaluator's shared-memory waits wait for them too (shared scoreboards).
    // One interval ahead still left waits behind where an interval is shorter than an HBM round trip, so the rows are
    // fetched TWO intervals ahead into two slots (i & 1): euler 2.19 -> 2.04 ms, midpoint odeint_adjoint 2.35 -> 2.28,
    // midpoint discrete 2.69 -> 2.64, rk4 unchanged (3.61) at 2^20 x 100 (SLODE_FX_ROW_AHEAD=1 builds the old schedule).
    // Groups are committed in the order state(i), cotangent(i), state(i-1), cotangent(i-1), ...: each read waits for
    // all but the three newest groups.
    constexpr int kStateAhead = (MODE == SLODE_BWD_DISCRETE) ? 0 : 1;
    const float* px = xs + (int64_t)(T - 2 + kStateAhead) * st;   // running pointers: no 64-bit multiplies in the loop
    const float* pg = gs + (int64_t)(T - 2) * gst;
    constexpr int kSlot = S * kThreads;   // floats between the two slots of a row
    if (T > 1) {
#pragma unroll
      for (int a = 0; a < kAhead; ++a) {   // intervals T-2, (T-3): groups state, cotangent, (state, cotangent)
        if (T - 2 - a >= 0) row_fetch<S>(rowx + ((T - 2 - a) & (kAhead - 1)) * kSlot, px);
        cp_commit();
        if (T - 2 - a >= 0) row_fetch<S, true>(rowg + ((T - 2 - a) & (kAhead - 1)) * kSlot, pg, ok);
        cp_commit();
        px -= st;   // px / pg end up at the rows of the interval kAhead behind the one in hand
        pg -= gst;
      }
    }
    V<NP> G0, D0, G1, D1, G2, D2;   // rk4: the three evaluations of the interval in hand
    int pa = 0, pb = 0;             // rk4: walk positions after the first two seeks of the interval
#pragma unroll kBwdUnroll
    for (int i = T - 2; i >= 0; --i) {
      const float t0 = t_ahead;
      if (i > 0) t_ahead = ld_early(tgrid + i - 1);
      // euler (both modes) and the midpoint odeint_adjoint step accumulate ONE evaluation per interval, after all of the
      // interval's seeks: the prefix sums do not move between a seek and that accumulation, so the flip records are
      // written inside the seek trips themselves and the separate pass over the flipped keys (another divergent trip
      // per flip) is not needed there
      auto rec_now = [&](int j) { sw.record(rec, tab, j); };
      auto state_row = [&]() {  // the interval's state row; its slot is refilled for interval i - 2 at once
        float* const slot = rowx + (i & (kAhead - 1)) * kSlot;
        cp_wait<2 * kAhead - 1>();
        const V<NP> r = row_read<S>(slot);
        if (i >= kAhead) row_fetch<S>(slot, px);
        px -= st;
        cp_commit();
        return r;
      };
      if (MODE == SLODE_BWD_DISCRETE) {
        const float dt = t1 - t0;
        if (METHOD == SLODE_METHOD_EULER) {
          V<NP> G, D;
          pl.seek(wt, tab, t0, rec_now);
          pl.template eval<false>(t0, G, D);
          const V<NP> x = state_row();
          const V<NP> gk = vscale<NP>(lam, dt);
          const V<NP> gy = vmul<NP>(gk, D);   // D holds -sigmoid
          sw.add(t0, gk, gy, x, G, D);
          lam = vadd<NP>(lam, gy);
        } else if (METHOD == SLODE_METHOD_MIDPOINT) {
          const float half_dt = 0.5f * dt;
          const float tm = t0 + half_dt;
          V<NP> Gm, Dm, Ga, Da;
          pl.seek(wt, tab, tm);
          const int pm = pl.pos;
          pl.template eval<false>(tm, Gm, Dm);
          pl.seek(wt, tab, t0);
          pl.template eval<false>(t0, Ga, Da);
          const V<NP> x = state_row();
          const V<NP> ym = vaxpy<NP>(half_dt, rhs<NP>(Ga, Da, x), x);
          V<NP> gk = vscale<NP>(lam, dt);  // dL/dk2
          V<NP> gy = vmul<NP>(gk, Dm);     // dL/dy_mid
          sw.events(rec, tab, pdone, pm);
          sw.add(tm, gk, gy, ym, Gm, Dm);
          lam = vadd<NP>(lam, gy);
          gk = vscale<NP>(gy, half_dt);  // dL/dk1
          gy = vmul<NP>(gk, Da);
          sw.events(rec, tab, pm, pl.pos);
          pdone = pl.pos;
          sw.add(t0, gk, gy, x, Ga, Da);
          lam = vadd<NP>(lam, gy);
        } else {  // rk4 3/8
          const float dt3 = dt * kOneThird;
          const float ta = t0 + dt * kOneThird, tb = t0 + dt * kTwoThirds;
          pl.seek(wt, tab, tb);
          pb = pl.pos;
          pl.template eval<false>(tb, G2, D2);
          pl.seek(wt, tab, ta);
          pa = pl.pos;
          pl.template eval<false>(ta, G1, D1);
          pl.seek(wt, tab, t0);
          pl.template eval<false>(t0, G0, D0);
          const V<NP> x = state_row();
          V<NP> Y2, Y3, Y4;
          {
            const V<NP> k1 = rhs<NP>(G0, D0, x);
            Y2 = vaxpy<NP>(dt3, k1, x);
            const V<NP> k2 = rhs<NP>(G1, D1, Y2);
            Y3 = vaxpy<NP>(dt, vaxpy<NP>(-kOneThird, k1, k2), x);
            const V<NP> k3 = rhs<NP>(G2, D2, Y3);
            Y4 = vaxpy<NP>(dt, vadd<NP>(vsub<NP>(k1, k2), k3), x);
          }
          const V<NP> wv = vscale<NP>(lam, 0.125f * dt);
          // stage 4 (time t1, carried evaluation): gk4 = w
          V<NP> gy = vmul<NP>(wv, Dc);
          sw.add(t1, wv, gy, Y4, Gc, Dc);
          lam = vadd<NP>(lam, gy);
          V<NP> gk1 = vaxpy<NP>(dt, gy, wv);
          V<NP> gk2 = vaxpy<NP>(-dt, gy, vscale<NP>(wv, 3.0f));
          const V<NP> gk3 = vaxpy<NP>(dt, gy, vscale<NP>(wv, 3.0f));
          // stage 3
          sw.events(rec, tab, pdone, pb);
          gy = vmul<NP>(gk3, D2);
          sw.add(tb, gk3, gy, Y3, G2, D2);
          lam = vadd<NP>(lam, gy);
          gk2 = vaxpy<NP>(dt, gy, gk2);
          gk1 = vaxpy<NP>(-dt3, gy, gk1);
          // stage 2
          sw.events(rec, tab, pb, pa);
          gy = vmul<NP>(gk2, D1);
          sw.add(ta, gk2, gy, Y2, G1, D1);
          lam = vadd<NP>(lam, gy);
          gk1 = vaxpy<NP>(dt3, gy, gk1);
          // stage 1 (time t0; its evaluation is the carried one of the next interval)
          sw.events(rec, tab, pa, pl.pos);
          pdone = pl.pos;
          gy = vmul<NP>(gk1, D0);
          sw.add(t0, gk1, gy, x, G0, D0);
          lam = vadd<NP>(lam, gy);
          Gc = G0;
          Dc = D0;
        }
      } else {
        // torchdiffeq.odeint_adjoint emulation: one step of the same method on the augmented system
        // [y, a, a_theta] from t1 down to t0, y restarted from the stored sol[i+1].  In reversed time s=-t the
        // step is ds = t1 - t0 > 0 with  Ky = D*y - A,  Ka = -a*D,  a_theta += w_m * a_m^T df/dtheta(t_m, y_m).
        const float ds = t1 - t0;
        const V<NP> y = state_row();
        if (METHOD == SLODE_METHOD_EULER) {
          V<NP> G, D;
          pl.seek(wt, tab, t1, rec_now);
          pl.template eval<false>(t1, G, D);
          const V<NP> v = vscale<NP>(lam, ds);
          const V<NP> gy = vmul<NP>(v, D);
          sw.add(t1, v, gy, y, G, D);
          lam = vadd<NP>(lam, gy);
        } else if (METHOD == SLODE_METHOD_MIDPOINT) {
          const float half = 0.5f * ds;
          const float tm = t1 - half;
          V<NP> Ga, Da, Gm, Dm;
          pl.seek(wt, tab, t1, rec_now);   // (the stage at t1 has weight 0 in a_theta: nothing is accumulated between
          pl.template eval<false>(t1, Ga, Da);   //  its flips and those found on the way to tm)
          pl.seek(wt, tab, tm, rec_now);
          pl.template eval<false>(tm, Gm, Dm);
          const V<NP> ym = vaxpy<NP>(-half, rhs<NP>(Ga, Da, y), y);  // y + half*(D1*y - A1)
          const V<NP> am = vaxpy<NP>(half, vmul<NP>(lam, Da), lam);  // a + half*(-a*D1), D holds -sigmoid
          const V<NP> v = vscale<NP>(am, ds);
          const V<NP> gy = vmul<NP>(v, Dm);
          sw.add(tm, v, gy, ym, Gm, Dm);
          lam = vadd<NP>(lam, gy);
        } else {  // rk4 3/8 on the augmented system; Ky = -f, Ka = -a*D; new evaluations at ta, tb, t0
          const float w8 = 0.125f * ds;
          const float ta = t1 - ds * kOneThird, tb = t1 - ds * kTwoThirds;
          pl.seek(wt, tab, ta);
          pa = pl.pos;
          pl.template eval<false>(ta, G0, D0);
          pl.seek(wt, tab, tb);
          pb = pl.pos;
          pl.template eval<false>(tb, G1, D1);
          pl.seek(wt, tab, t0);
          pl.template eval<false>(t0, G2, D2);
          // stage 1 at t1 (carried evaluation)
          const V<NP> f1 = rhs<NP>(Gc, Dc, y);
          const V<NP> ka1 = vmul<NP>(lam, Dc);
          sw.add(t1, vscale<NP>(lam, w8), vscale<NP>(ka1, w8), y, Gc, Dc);
          // stage 2
          V<NP> ym = vaxpy<NP>(-ds * kOneThird, f1, y);
          V<NP> am = vaxpy<NP>(ds * kOneThird, ka1, lam);
          const V<NP> f2_ = rhs<NP>(G0, D0, ym);
          const V<NP> ka2 = vmul<NP>(am, D0);
          sw.events(rec, tab, pdone, pa);
          sw.add(ta, vscale<NP>(am, 3.0f * w8), vscale<NP>(ka2, 3.0f * w8), ym, G0, D0);
          // stage 3
          ym = vaxpy<NP>(-ds, vaxpy<NP>(-kOneThird, f1, f2_), y);
          am = vaxpy<NP>(ds, vaxpy<NP>(-kOneThird, ka1, ka2), lam);
          const V<NP> f3 = rhs<NP>(G1, D1, ym);
          const V<NP> ka3 = vmul<NP>(am, D1);
          sw.events(rec, tab, pa, pb);
          sw.add(tb, vscale<NP>(am, 3.0f * w8), vscale<NP>(ka3, 3.0f * w8), ym, G1, D1);
          // stage 4 at t0 (becomes the carried evaluation)
          ym = vaxpy<NP>(-ds, vadd<NP>(vsub<NP>(f1, f2_), f3), y);
          am = vaxpy<NP>(ds, vadd<NP>(vsub<NP>(ka1, ka2), ka3), lam);
          const V<NP> ka4 = vmul<NP>(am, D2);
          sw.events(rec, tab, pb, pl.pos);
          pdone = pl.pos;
          sw.add(t0, vscale<NP>(am, w8), vscale<NP>(ka4, w8), ym, G2, D2);
          const V<NP> asum = vadd<NP>(vaxpy<NP>(3.0f, vadd<NP>(ka2, ka3), ka1), ka4);
          lam = vaxpy<NP>(w8, asum, lam);
          Gc = G2;
          Dc = D2;
        }
      }
      {
        float* const slot = rowg + (i & (kAhead - 1)) * kSlot;
        cp_wait<2 * kAhead - 1>();
        lam = vadd<NP>(lam, row_read<S>(slot));   // a masked-off thread's row was zero-filled
        if (i >= kAhead) row_fetch<S, true>(slot, pg, ok);
        pg -= gst;
        cp_commit();
      }
      t1 = t0;
    }

    // ---- end of the sweep.  For every (trajectory, hidden unit) combine the recorded and the final prefix sums
    // into the sums over the evaluations where the unit was active,
    //     active throughout: final      turned off: record      turned on: final - record      never: 0
    // and turn them into dc_j (per trajectory), dw1t_j and dW_oj (sums over trajectories).  The warp switches
    // from "lane = trajectory" to "lane = hidden unit": it walks its 32 trajectories one after the other with
    // every lane holding ITS unit's sums in registers, so no cross-lane reduction is needed at all (round 1 and the
    // first version of this kernel spent ~10 % of their instructions in shuffle reductions here).  What the lanes
    // exchange goes through shared memory: each trajectory's final (P, Q) in the warp's key-table region (the walk
    // is over), dc_j written IN PLACE over c_j (element (j, trajectory) of the c table, stride odd: conflict-free
    // both ways), z rows in the warp's zT.
    constexpr int JC = SH::JC, JS = SH::JS, PQS = SH::PQS;
    constexpr int KR = (2 * NQ + 1 <= 16) ? 16 : 32;
    static_assert(2 * NQ + 1 <= 32, "state dimension");
    // head biases: total of the cotangents over all evaluations (P is still in registers)
    {
      float red[KR];
#pragma unroll
      for (int k = 0; k < KR; ++k) red[k] = 0.0f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) unpk(sw.P[q], red[2 * q], red[2 * q + 1]);
      warp_reduce_to<KR>(red, lane, [&](int slot) -> float* {
        if (slot >= 2 * NQ) return nullptr;
        const int q = slot >> 1, h = slot & 1;
        const bool growth = q < NP;
        const int s = 2 * (growth ? q : q - NP) + h;
        if (s >= S) return nullptr;
        return acc + (growth ? GL::bg : GL::bd) + s;
      });
    }
    __syncwarp();  // every lane is done with the key table: it becomes the (P, Q) exchange buffer
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float p0, p1, q0, q1;
      unpk(sw.P[q], p0, p1);
      unpk(sw.Q[q], q0, q1);
      reinterpret_cast<float4*>(pqw + lane * PQS)[q] = make_float4(p0, p1, q0, q1);
    }
    __syncwarp();   // (the warp's z rows have been in zT since the tile's prologue)

    const float* recs = wt + SH::BIAS;
    float* const cw = tab.c - lane;                 // element (j, trajectory bb of this warp) at cw[j * stride + bb]
    const f2* const rec0 = rec - (size_t)lane * (H * 2 * NQ);  // thread bb's records start at rec0 + bb * H * 2NQ
    // weight-gradient outer products of one chunk of units with the z rows: lane <-> unit j0 + lane,
    //   gW[j][l] (row stride L) += sum_bb val(bb) * z[bb][l],   gb[j] += sum_bb val(bb)
    auto outer = [&](int j0, auto val, float* gW, float* gb) {
      for (int l0 = 0; l0 < L; l0 += 16) {
        float a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = 0.0f;
        float sb_ = 0.0f;
#pragma unroll 2
        for (int bb_ = 0; bb_ < 32; ++bb_) {
          const float d = val(bb_);
          sb_ += d;
          const float4* zr = reinterpret_cast<const float4*>(zT + bb_ * LQ + l0);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            if (l0 + 4 * k4 < L) {
              const float4 zz = zr[k4];
              a[4 * k4] = fmaf(d, zz.x, a[4 * k4]);
              a[4 * k4 + 1] = fmaf(d, zz.y, a[4 * k4 + 1]);
              a[4 * k4 + 2] = fmaf(d, zz.z, a[4 * k4 + 2]);
              a[4 * k4 + 3] = fmaf(d, zz.w, a[4 * k4 + 3]);
            }
          }
        }
        if (lane < JC) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (l0 + k < L) atomicAdd(gW + (j0 + lane) * L + l0 + k, a[k]);
          }
          if (l0 == 0) atomicAdd(gb + j0 + lane, sb_);
        }
      }
    };
    // per trajectory (lane <-> trajectory): grad_z[b][l] (+)= sum_jj Wl[l][j0 + jj] * val(jj); the weight rows are
    // read as float4 over the units (broadcast); partial sums of later chunks / the second net go through grad_z
    // itself (the row stays in L1)
    auto dz_from = [&](int j0, auto val, const float* Wl, bool accumulate) {
      constexpr int JC4 = (JC + 3) / 4 * 4;
      float d[JC4];
#pragma unroll
      for (int jj = 0; jj < JC4; ++jj) d[jj] = jj < JC ? val(jj) : 0.0f;
      float* gzrow = grad_z + b * L;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const float4* wr = reinterpret_cast<const float4*>(Wl + l * SH::HQ + j0);
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int q4 = 0; q4 < JC4 / 4; ++q4) {
          const float4 ww = wr[q4];
          a0 = fmaf(ww.x, d[4 * q4], a0);
          a1 = fmaf(ww.y, d[4 * q4 + 1], a1);
          a0 = fmaf(ww.z, d[4 * q4 + 2], a0);
          a1 = fmaf(ww.w, d[4 * q4 + 3], a1);
        }
        if (ok) gzrow[l] = (accumulate ? gzrow[l] : 0.0f) + (a0 + a1);
      }
    };

#pragma unroll 1
    for (int j0 = 0; j0 < H; j0 += JC) {
      // ---- dynamics part of the chunk, lane <-> unit
      const bool act = lane < JC;
      const int j = act ? j0 + lane : j0;   // idle lanes shadow unit j0 with zero weights
      const float* wrec = recs + j * SH::UNIT;
      f2 wq[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) wq[q] = act ? reinterpret_cast<const f2*>(wrec)[q] : 0ull;
      const f2 wj = bc(act ? wrec[2 * NQ] : 0.0f);
      f2 racc[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) racc[q] = 0ull;
      float rw1t = 0.0f;
      // status of unit j along trajectory bb: gate at the first evaluation, flipped since
      auto flags = [&](int bb, bool& f, bool& fl) {
        if constexpr (SH::BIG) {
          const uint8_t sbits = T > 1 ? (tab.fs - lane)[(size_t)j * tab.ks + bb] : 0;
          f = sbits & 1;
          fl = sbits & 2;
        } else {
          uint32_t fw = __shfl_sync(0xffffffffu, first[0], bb), lw = __shfl_sync(0xffffffffu, sw.flipped[0], bb);
#pragma unroll
          for (int ww = 1; ww < SH::NWR; ++ww) {
            const uint32_t fw2 = __shfl_sync(0xffffffffu, first[ww], bb);
            const uint32_t lw2 = __shfl_sync(0xffffffffu, sw.flipped[ww], bb);
            if ((j >> 5) == ww) { fw = fw2; lw = lw2; }
          }
          f = (fw >> (j & 31)) & 1u;
          fl = (lw >> (j & 31)) & 1u;
        }
      };
      // the record of unit j written by trajectory bb (zeros if the unit never flipped there)
      auto fetch = [&](int bb, bool fl, ulonglong2 (&r)[NQ]) {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(rec0 + ((size_t)bb * H + j) * (2 * NQ));
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          r[q] = make_ulonglong2(0ull, 0ull);
          if (fl) r[q] = src[q];
        }
      };
      bool f_next, fl_next;
      ulonglong2 r_next[NQ];
      flags(0, f_next, fl_next);
      fetch(0, fl_next, r_next);
#pragma unroll 1
      for (int bb = 0; bb < 32; ++bb) {
        const bool f = f_next, fl = fl_next;
        ulonglong2 r_cur[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) r_cur[q] = r_next[q];
        if (bb + 1 < 32) {  // the next trajectory's record is in flight while this one is processed
          flags(bb + 1, f_next, fl_next);
          fetch(bb + 1, fl_next, r_next);
        }
        const bool last = f != fl;
        const f2 ar = bc(fl ? (f ? 1.0f : -1.0f) : 0.0f), af = bc(last ? 1.0f : 0.0f);
        const f2 cj = bc(cw[(size_t)j * tab.cs + bb]);
        const float4* pq4 = reinterpret_cast<const float4*>(pqw + bb * PQS);
        f2 s1 = 0ull, s2 = 0ull;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const float4 v = pq4[q];
          const f2 pe = fma2(ar, r_cur[q].x, mul2(af, pk(v.x, v.y)));
          const f2 qe = fma2(ar, r_cur[q].y, mul2(af, pk(v.z, v.w)));
          s1 = fma2(wq[q], pe, s1);
          s2 = fma2(wq[q], qe, s2);
          racc[q] = fma2(wj, qe, fma2(cj, pe, racc[q]));
        }
        rw1t = fmaf(kNegLn2, lo_of(s2) + hi_of(s2), rw1t);  // packed head weights are scaled by -log2(e)
        if (act) cw[(size_t)j * tab.cs + bb] = kNegLn2 * (lo_of(s1) + hi_of(s1));  // dc_j of trajectory bb, in place
      }
      if (act) {
        atomicAdd(acc + GL::w1t + j, rw1t);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const bool growth = q < NP;
          const int s0 = 2 * (growth ? q : q - NP);
          float r0, r1;
          unpk(racc[q], r0, r1);
          atomicAdd(acc + (growth ? GL::Wg : GL::Wd) + s0 * H + j, r0);
          if (s0 + 1 < S) atomicAdd(acc + (growth ? GL::Wg : GL::Wd) + (s0 + 1) * H + j, r1);
        }
      }
      __syncwarp();
      if (fused) {
        // dW1z[j][l] += sum_bb dc[bb][j] z[bb][l],  db1[j] += sum_bb dc[bb][j]
        outer(j0, [&](int bb_) { return lane < JC ? cw[(size_t)(j0 + lane) * tab.cs + bb_] : 0.0f; },
              acc + GL::W1z(L), acc + GL::b1(L));
        // dz through c (discrete mode only: odeint_adjoint gives z no gradient through the dynamics)
        if (MODE == SLODE_BWD_DISCRETE) {
          dz_from(j0, [&](int jj) { return cw[(size_t)(j0 + jj) * tab.cs + lane]; }, ls.Wz, j0 != 0);
        } else if (j0 == 0 && ok) {
          for (int l = 0; l < L; ++l) grad_z[b * L + l] = 0.0f;
        }
      } else if (ok) {
#pragma unroll 4
        for (int jj = 0; jj < JC; ++jj) grad_c[b * H + j0 + jj] = cw[(size_t)(j0 + jj) * tab.cs + lane];
      }
    }
    __syncwarp();  // the (P, Q) exchange buffer is free: it serves as the x0 net's transposition buffer [32][JS]
    if (fx0) {
      // ---- x0 net: da_j = [ha_j > 0] sum_s Wb[s][j] db_s,  db = dL/dx0 * x0 (1 - x0);  hr_j = relu(ha_j)
      float* const dcT = pqw;
      const V<NP> x0 = vload<S>(xs);
      float db[S];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        float l0_, l1_, a0, a1;
        unpk(lam.v[p], l0_, l1_);
        unpk(x0.v[p], a0, a1);
        db[2 * p] = l0_ * (a0 - a0 * a0);
        if (2 * p + 1 < S) db[2 * p + 1] = l1_ * (a1 - a1 * a1);
      }
#pragma unroll 1
      for (int j0 = 0; j0 < H; j0 += JC) {
        float ha[JC];
        lat_chunk<JC>(ls.Wa, ls.ba, SH::HQ, j0, zT + lane * LQ, L, ha);
        // da -> buffer
#pragma unroll
        for (int jj = 0; jj < JC; ++jj) {
          float a = 0.0f;
#pragma unroll
          for (int s = 0; s < S; ++s) a = fmaf(ls.Wb[s * SH::HQ + j0 + jj], db[s], a);
          dcT[lane * JS + jj] = ha[jj] > 0.0f ? a * live : 0.0f;
        }
        __syncwarp();
        outer(j0, [&](int bb_) { return lane < JC ? dcT[bb_ * JS + lane] : 0.0f; }, acc + GL::Wa(L), acc + GL::ba(L));
        dz_from(j0, [&](int jj) { return dcT[lane * JS + jj]; }, ls.Wa, true);
        __syncwarp();
        // hr -> buffer;  dWb[s][j] += sum_b db[b][s] hr[b][j]
#pragma unroll
        for (int jj = 0; jj < JC; ++jj) dcT[lane * JS + jj] = fmaxf(ha[jj], 0.0f) * live;
        __syncwarp();
        {
          float a[S];
#pragma unroll
          for (int s = 0; s < S; ++s) a[s] = 0.0f;
#pragma unroll 2
          for (int bb_ = 0; bb_ < 32; ++bb_) {
            const float h = lane < JC ? dcT[bb_ * JS + lane] : 0.0f;
#pragma unroll
            for (int s = 0; s < S; ++s) a[s] = fmaf(__shfl_sync(0xffffffffu, db[s], bb_), h, a[s]);
          }
          if (lane < JC) {
#pragma unroll
            for (int s = 0; s < S; ++s) atomicAdd(acc + GL::Wb(L) + s * H + j0 + lane, a[s]);
          }
        }
        __syncwarp();
      }
      {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = k < S ? db[k < S ? k : 0] * live : 0.0f;
        warp_reduce_to<16>(v, lane, [&](int slot) -> float* { return slot < S ? acc + GL::bb(L) + slot : nullptr; });
      }
    }
    if (!fx0) vstore<S>(grad_y0 + b * S, ok, lam);
    __syncwarp();  // the transposition buffer doubles as the next tile's key table
  }

  __syncthreads();
  for (int i = tid; i < n_acc; i += kThreads) atomicAdd(grad_w + i, acc[i]);
}

// ---------------------------------------------------------------------------------------------
// launch plans and launchers (one instantiation set per compiled shape, see slode_fixed_<H>_<S>.cu)
// ---------------------------------------------------------------------------------------------
struct Plan {
  int grid;
  size_t smem;
  size_t ws_bytes;   // global scratch the caller provides: wide-layer tables (+ flip records, backward)
};

// Resident blocks per SM of a kernel at a dynamic shared-memory size, looked up once per (kernel, size, device):
// the hot path then makes no CUDA runtime call except the launch itself (cheap at the reference's batch sizes, and
// nothing that could be refused while a stream is being captured into a graph).
int cached_blocks_per_sm(const void* kern, size_t smem, int* blocks_per_sm);  // slode_fixed_api.cu

template <class K>
int plan_grid(K kern, size_t smem, int64_t B, int sms, int* grid) {
  int blocks_per_sm = 0;
  const int rc = cached_blocks_per_sm(reinterpret_cast<const void*>(kern), smem, &blocks_per_sm);
  if (rc) return rc;
  const int64_t tiles = (B + kThreads - 1) / kThreads;
  // whole waves of resident blocks; tiles are handed out grid-stride
  *grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)sms * blocks_per_sm));
  return SLODE_OK;
}

template <int H, int S, int METHOD, bool HEADS>
int plan_fwd(const FwdArgs& a, Plan* p) {
  p->smem = fwd_smem_bytes<H, S>(a.lat.L, a.lat.z != nullptr, a.st == S, HEADS, a.T);
  const int rc = plan_grid(fixed_fwd_kernel<H, S, METHOD, HEADS>, p->smem, a.B, a.sms, &p->grid);
  if (rc) return rc;
  p->ws_bytes = big_tab_bytes<H, S>() * (size_t)p->grid * kThreads;
  return SLODE_OK;
}

template <int H, int S, int METHOD, int MODE>
int plan_bwd(const BwdArgs& a, Plan* p) {
  p->smem = bwd_smem_bytes<H, S>(a.lat.L, a.lat.z != nullptr);
  const int rc = plan_grid(fixed_bwd_kernel<H, S, METHOD, MODE>, p->smem, a.B, a.sms, &p->grid);
  if (rc) return rc;
  p->ws_bytes = (big_tab_bytes<H, S>() + rec_bytes_per_thread<H, S>()) * (size_t)p->grid * kThreads;
  return SLODE_OK;
}

template <int H, int S, int METHOD, bool HEADS>
int launch_fwd(const FwdArgs& a, const PackSrc& w, int w1t_stride, bool plan_only, size_t* ws_need) {
  Plan p;
  const int rc = plan_fwd<H, S, METHOD, HEADS>(a, &p);
  if (rc) return rc;
  *ws_need = p.ws_bytes;
  if (plan_only) return SLODE_OK;
  if (p.ws_bytes > a.ws_bytes || (p.ws_bytes && !a.ws)) {
    set_error("fixed-grid forward: workspace of %zu bytes given, %zu needed (slode_fixed_workspace_bytes)", a.ws_bytes,
              p.ws_bytes);
    return SLODE_EINVAL;
  }
  if (p.ws_bytes && (reinterpret_cast<uintptr_t>(a.ws) & 15)) {
    set_error("fixed-grid forward: workspace must be 16-byte aligned");
    return SLODE_EINVAL;
  }
  fixed_fwd_kernel<H, S, METHOD, HEADS><<<p.grid, kThreads, p.smem, a.stream>>>(
      a.B, a.T, a.t, a.c, a.y0, a.sol, a.st, a.sb, w, w1t_stride, a.lat, static_cast<unsigned char*>(a.ws), a.heads);
  SLODE_CUDA_TRY(cudaGetLastError());
  return SLODE_OK;
}

template <int H, int S, int METHOD, int MODE>
int launch_bwd(const BwdArgs& a, const PackSrc& w, int w1t_stride, bool plan_only, size_t* ws_need) {
  Plan p;
  const int rc = plan_bwd<H, S, METHOD, MODE>(a, &p);
  if (rc) return rc;
  *ws_need = p.ws_bytes;
  if (plan_only) return SLODE_OK;
  if (p.ws_bytes > a.ws_bytes || !a.ws) {
    set_error("fixed-grid backward: workspace of %zu bytes given, %zu needed (slode_fixed_workspace_bytes)", a.ws_bytes,
              p.ws_bytes);
    return SLODE_EINVAL;
  }
  if (reinterpret_cast<uintptr_t>(a.ws) & 15) {
    set_error("fixed-grid backward: workspace must be 16-byte aligned");
    return SLODE_EINVAL;
  }
  fixed_bwd_kernel<H, S, METHOD, MODE><<<p.grid, kThreads, p.smem, a.stream>>>(
      a.B, a.T, a.t, a.c, a.sol, a.st, a.sb, a.gsol, a.gst, a.gsb, a.gy0, a.gc, a.gw, w, w1t_stride, a.lat, a.gz,
      static_cast<unsigned char*>(a.ws));
  SLODE_CUDA_TRY(cudaGetLastError());
  return SLODE_OK;
}

template <int H, int S>
int fwd_shape(const FwdArgs& a, const PackSrc& w, int w1t_stride, bool plan_only, size_t* ws_need) {
  if (a.heads.W) {
    switch (a.method) {
      case SLODE_METHOD_EULER: return launch_fwd<H, S, SLODE_METHOD_EULER, true>(a, w, w1t_stride, plan_only, ws_need);
      case SLODE_METHOD_MIDPOINT:
        return launch_fwd<H, S, SLODE_METHOD_MIDPOINT, true>(a, w, w1t_stride, plan_only, ws_need);
      case SLODE_METHOD_RK4: return launch_fwd<H, S, SLODE_METHOD_RK4, true>(a, w, w1t_stride, plan_only, ws_need);
    }
  }
  switch (a.method) {
    case SLODE_METHOD_EULER: return launch_fwd<H, S, SLODE_METHOD_EULER, false>(a, w, w1t_stride, plan_only, ws_need);
    case SLODE_METHOD_MIDPOINT:
      return launch_fwd<H, S, SLODE_METHOD_MIDPOINT, false>(a, w, w1t_stride, plan_only, ws_need);
    case SLODE_METHOD_RK4: return launch_fwd<H, S, SLODE_METHOD_RK4, false>(a, w, w1t_stride, plan_only, ws_need);
  }
  set_error("fixed-grid forward: unknown method %d", a.method);
  return SLODE_EINVAL;
}

template <int H, int S>
int bwd_shape(const BwdArgs& a, const PackSrc& w, int w1t_stride, bool plan_only, size_t* ws_need) {
#define SLODE_FX_BWD_CASE(M)                                                                              \
  case M:                                                                                                 \
    if (a.mode == SLODE_BWD_DISCRETE)                                                                     \
      return launch_bwd<H, S, M, SLODE_BWD_DISCRETE>(a, w, w1t_stride, plan_only, ws_need);               \
    return launch_bwd<H, S, M, SLODE_BWD_TDE_ADJOINT>(a, w, w1t_stride, plan_only, ws_need);
  switch (a.method) {
    SLODE_FX_BWD_CASE(SLODE_METHOD_EULER)
    SLODE_FX_BWD_CASE(SLODE_METHOD_MIDPOINT)
    SLODE_FX_BWD_CASE(SLODE_METHOD_RK4)
  }
#undef SLODE_FX_BWD_CASE
  set_error("fixed-grid backward: unknown method %d", a.method);
  return SLODE_EINVAL;
}

}  // namespace fx
}  // namespace slode

// Defines the two entry functions of one compiled shape (looked up by slode_mlp.cu through slode_mlp_api.h).
#define SLODE_DEFINE_FIXED_SHAPE(H, S)                                                                             \
  namespace slode {                                                                                                \
  int fixed_fwd_##H##_##S(const FwdArgs& a, const PackSrc& w, int w1t_stride, bool plan_only, size_t* ws_need) {   \
    return fx::fwd_shape<H, S>(a, w, w1t_stride, plan_only, ws_need);                                              \
  }                                                                                                                \
  int fixed_bwd_##H##_##S(const BwdArgs& a, const PackSrc& w, int w1t_stride, bool plan_only, size_t* ws_need) {   \
    return fx::bwd_shape<H, S>(a, w, w1t_stride, plan_only, ws_need);                                              \
  }                                                                                                                \
  }
