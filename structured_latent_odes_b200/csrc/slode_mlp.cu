// Fixed-grid solve of the SLODE blackbox latent ODE and its reverse sweep, hand-written for sm_100a.
//
// Right-hand side (reference: Dynamics.forward, models/blackbox_ode.py:97-109):
//     h_j(t)  = relu(w1t_j * t + c_j)                       c = z W1[:,1:]^T + b1  (per trajectory)
//     A_k(t)  = sigmoid(bg_k + sum_j Wg_kj h_j(t))          "growth"
//     D_k(t)  = sigmoid(bd_k + sum_j Wd_kj h_j(t))          "degradation"
//     f(t,x)  = A(t) - D(t) * x                              (linear in the state, elementwise)
//
// Mapping: one thread = one trajectory for the whole time loop; state, Butcher stages and the
// time-invariant hidden pre-activations c[H] stay in registers.
//
// FP32 pipe (measured on B200, profiles/r01/fp32_pipes_microbench.jsonl): a 3-register FFMA
// sustains 84 FMA/clk/SM, FFMA with a uniform-register operand 118, and the packed FFMA2
// (fma.rn.f32x2) with a uniform-register operand 127 = the full 128-lane peak, even with one
// LDCU.128 per four FFMA2.  sm_100a has no constant-bank operand form: warp-uniform weights reach
// the FMA pipe through uniform registers (LDCU).  The kernels are therefore built on FFMA2 with
//     pair of OUTPUTS (o, o+1)  +=  (W[j][o], W[j][o+1]) (uniform pair)  *  h_j (broadcast .F32)
// so one trajectory per thread still fills both halves of every FMA.  Head outputs are ordered so
// that pairs line up with pairs of state components:
//     o = 4q+{0,1}: growth of states 2q,2q+1   o = 4q+{2,3}: degradation of states 2q,2q+1
//     S odd: the last pair is (growth_{S-1}, degradation_{S-1})
// and all per-state arithmetic (stages, adjoints) runs on the same pair layout (Vec<S>).
// rk4 (3/8 rule) re-uses the evaluation at t1 as the next step's evaluation at t0 (same float).
//
// Backward: reverse sweep over the stored grid states sol[i]; stages are recomputed.  Because the
// hidden layer sees only (t, z), the cotangents of the head pre-activations delta_o(e) at the
// evaluation times t_e determine every hidden-layer gradient through prefix sums
//     P_o = sum_e delta_o(e),   Q_o = sum_e delta_o(e) t_e
// taken over the evaluations where unit j is active.  The sweep keeps running P,Q (2*2S registers)
// and, whenever a unit's relu gate flips between consecutive evaluations (at most once per unit
// for monotone t, but the summation-by-parts below is valid for any number of flips), adds
// +-snapshot contributions:
//     dc_j   += s * sum_o W_oj P_o              (per trajectory -> grad_c)
//     dw1t_j += s * sum_o W_oj Q_o              (block accumulator)
//     dW_oj  += s * (w1t_j Q_o + c_j P_o)       (block accumulator, = sum_e delta_o h_j)
// with s=+1 when the unit turns off, -1 when it turns on, and +1 for every unit still active when
// the sweep ends.  This replaces the two dense 2S*H products per evaluation of a textbook backward.
#include <algorithm>
#include <type_traits>

#include "slode_common.cuh"

namespace slode {

// ---------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic
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
__device__ __forceinline__ float lo_of(f2 v) { float a, b; unpk(v, a, b); return a; }
__device__ __forceinline__ float hi_of(f2 v) { float a, b; unpk(v, a, b); return b; }
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

// S per-state values stored as S/2 pairs (+ one scalar when S is odd)
template <int S>
struct Vec {
  static constexpr int NP = S / 2;
  static constexpr bool TAIL = (S & 1) != 0;
  f2 p[NP > 0 ? NP : 1];
  float t;
};

#define VEC_FOR_PAIRS for (int q = 0; q < Vec<S>::NP; ++q)

template <int S> __device__ __forceinline__ Vec<S> vbc(float s) {
  Vec<S> r;
#pragma unroll
  VEC_FOR_PAIRS r.p[q] = bc(s);
  r.t = s;
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vadd(const Vec<S>& a, const Vec<S>& b) {
  Vec<S> r;
#pragma unroll
  VEC_FOR_PAIRS r.p[q] = add2(a.p[q], b.p[q]);
  r.t = Vec<S>::TAIL ? a.t + b.t : 0.0f;
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vsub(const Vec<S>& a, const Vec<S>& b) {
  Vec<S> r;
#pragma unroll
  VEC_FOR_PAIRS r.p[q] = sub2(a.p[q], b.p[q]);
  r.t = Vec<S>::TAIL ? a.t - b.t : 0.0f;
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vmul(const Vec<S>& a, const Vec<S>& b) {
  Vec<S> r;
#pragma unroll
  VEC_FOR_PAIRS r.p[q] = mul2(a.p[q], b.p[q]);
  r.t = Vec<S>::TAIL ? a.t * b.t : 0.0f;
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vscale(const Vec<S>& a, float s) {
  Vec<S> r;
  const f2 ss = bc(s);
#pragma unroll
  VEC_FOR_PAIRS r.p[q] = mul2(a.p[q], ss);
  r.t = Vec<S>::TAIL ? a.t * s : 0.0f;
  return r;
}
// s*a + c  (scalar s)
template <int S> __device__ __forceinline__ Vec<S> vaxpy(float s, const Vec<S>& a, const Vec<S>& c) {
  Vec<S> r;
  const f2 ss = bc(s);
#pragma unroll
  VEC_FOR_PAIRS r.p[q] = fma2(ss, a.p[q], c.p[q]);
  r.t = Vec<S>::TAIL ? fmaf(s, a.t, c.t) : 0.0f;
  return r;
}
// c - a*b
template <int S> __device__ __forceinline__ Vec<S> vnfma(const Vec<S>& a, const Vec<S>& b, const Vec<S>& c) {
  Vec<S> r;
  const f2 m1 = bc(-1.0f);
#pragma unroll
  VEC_FOR_PAIRS r.p[q] = fma2(mul2(a.p[q], m1), b.p[q], c.p[q]);
  r.t = Vec<S>::TAIL ? fmaf(-a.t, b.t, c.t) : 0.0f;
  return r;
}
// -(a*b)
template <int S> __device__ __forceinline__ Vec<S> vnmul(const Vec<S>& a, const Vec<S>& b) {
  Vec<S> r;
  const f2 m1 = bc(-1.0f);
#pragma unroll
  VEC_FOR_PAIRS r.p[q] = mul2(mul2(a.p[q], m1), b.p[q]);
  r.t = Vec<S>::TAIL ? -a.t * b.t : 0.0f;
  return r;
}
template <int S> __device__ __forceinline__ Vec<S> vload(const float* p) {
  Vec<S> r;
  float v[S];
#pragma unroll
  for (int s = 0; s < S; ++s) v[s] = ld_stream(p + s);
#pragma unroll
  VEC_FOR_PAIRS r.p[q] = pk(v[2 * q], v[2 * q + 1]);
  r.t = Vec<S>::TAIL ? v[S - 1] : 0.0f;
  return r;
}
template <int S> __device__ __forceinline__ void vstore(float* p, const Vec<S>& a) {
#pragma unroll
  VEC_FOR_PAIRS {
    float lo, hi;
    unpk(a.p[q], lo, hi);
    p[2 * q] = lo;
    p[2 * q + 1] = hi;
  }
  if (Vec<S>::TAIL) p[S - 1] = a.t;
}

// ---------------------------------------------------------------------------------------------
// packed weights in constant memory
// ---------------------------------------------------------------------------------------------
constexpr int kPackMax = 8192;
}  // namespace slode
// C linkage: the loads below name the symbol from inline PTX.
extern "C" {
__constant__ __align__(16) float slode_c_pack[slode::kPackMax];
}
namespace slode {

// Weight loads.  Left alone, both NVVM and ptxas hoist the (loop-invariant) constant loads out of the time loop
// into ~300 registers and spill them.  The loads are therefore `asm volatile` with a STATIC address
// (symbol + immediate): NVVM may not move or merge them, and ptxas keeps them inside the evaluation as 16-byte
// uniform-register loads  LDCU.128 UR, c[3][imm]  -- one load per two FFMA2.  (A register-offset address
// c[3][UR+imm] makes ptxas split every 16-byte load into two LDCU.64, one per FFMA2, and the kernel becomes
// issue-bound: measured in profiles/r01.)  ptxas would still merge loads of the SAME address issued by different
// evaluations of one time step into ordinary registers, so every evaluation site of a kernel reads its own copy
// ("slot") of the packed weights: kSlots copies sit back to back in constant memory.
template <int OFF_FLOATS>
__device__ __forceinline__ void ldc_pair2(f2& a, f2& b) {
  static_assert(OFF_FLOATS % 4 == 0, "16-byte aligned");
  float x, y, z, w;
  asm volatile("ld.const.v4.f32 {%0, %1, %2, %3}, [slode_c_pack+%4];"
               : "=f"(x), "=f"(y), "=f"(z), "=f"(w)
               : "n"(OFF_FLOATS * 4));
  a = pk(x, y);
  b = pk(z, w);
}
template <int OFF_FLOATS>
__device__ __forceinline__ void ldc_pair1(f2& a) {
  static_assert(OFF_FLOATS % 2 == 0, "8-byte aligned");
  float x, y;
  asm volatile("ld.const.v2.f32 {%0, %1}, [slode_c_pack+%2];" : "=f"(x), "=f"(y) : "n"(OFF_FLOATS * 4));
  a = pk(x, y);
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// head-output order used everywhere on the device (see file comment)
__host__ __device__ constexpr int out_is_degr(int o, int S) {
  return (o >= 4 * (S / 2)) ? (o & 1) : ((o >> 1) & 1);
}
__host__ __device__ constexpr int out_state(int o, int S) {
  return (o >= 4 * (S / 2)) ? (S - 1) : (2 * (o >> 2) + (o & 1));
}

constexpr int kSlots = 5;
template <int H, int S>
struct Pack {
  static constexpr int K2 = 2 * S;
  static constexpr int HP = (H + 3) / 4 * 4;   // every region starts 16-byte aligned
  static constexpr int KP = (K2 + 3) / 4 * 4;
  static constexpr int W1T = 0;                // [j]     time column of the hidden layer
  static constexpr int BH = HP;                // [o]     head biases, pre-scaled by -log2(e)
  static constexpr int WH = HP + KP;           // [j][o]  head weights, pre-scaled by -log2(e)
  static constexpr int N = (WH + H * K2 + 3) / 4 * 4;  // one slot, 16-byte multiple
};

__global__ void pack_kernel(int H, int S, const float* __restrict__ w1t, const float* __restrict__ Wg,
                            const float* __restrict__ bg, const float* __restrict__ Wd,
                            const float* __restrict__ bd, float* __restrict__ out) {
  const int K2 = 2 * S, HP = (H + 3) / 4 * 4, KP = (K2 + 3) / 4 * 4, BH = HP, WH = HP + KP,
            N = (WH + H * K2 + 3) / 4 * 4;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float v = 0.0f;
    if (i >= WH + H * K2) {
    } else if (i < HP) {
      if (i < H) v = w1t[i];
    } else if (i < WH) {
      const int o = i - BH;
      if (o < K2) {
        const int s = out_state(o, S);
        v = kNegLog2e * (out_is_degr(o, S) ? bd[s] : bg[s]);
      }
    } else {
      const int j = (i - WH) / K2, o = (i - WH) % K2, s = out_state(o, S);
      v = kNegLog2e * (out_is_degr(o, S) ? Wd[s * H + j] : Wg[s * H + j]);
    }
    for (int slot = 0; slot < kSlots; ++slot) out[slot * N + i] = v;
  }
}

template <int H>
struct MaskWords {
  static constexpr int NW = (H + 31) / 32;
};

// Result of one RHS evaluation in pair layout: sig[2q] = (A_2q, A_2q+1), sig[2q+1] = (D_2q, D_2q+1),
// S odd: sig[S-1] = (A_{S-1}, D_{S-1}).
template <int S>
struct Sig {
  f2 v[S];
  __device__ __forceinline__ Vec<S> A() const {
    Vec<S> r;
#pragma unroll
    VEC_FOR_PAIRS r.p[q] = v[2 * q];
    r.t = Vec<S>::TAIL ? lo_of(v[S - 1]) : 0.0f;
    return r;
  }
  __device__ __forceinline__ Vec<S> D() const {
    Vec<S> r;
#pragma unroll
    VEC_FOR_PAIRS r.p[q] = v[2 * q + 1];
    r.t = Vec<S>::TAIL ? hi_of(v[S - 1]) : 0.0f;
    return r;
  }
};

// One RHS evaluation at time t.  c2[jp] = (c_2jp, c_2jp+1).  Gate word w covers units
// [32w, 32w+n_w); unit j sits at bit (n_w - 1 - (j - 32w)).
template <int H, int S, bool MASK, int SLOT>
__device__ __forceinline__ void mlp_eval(float t, const f2 (&c2)[(H + 1) / 2], Sig<S>& out,
                                         uint32_t (&gate)[MaskWords<H>::NW]) {
  using P = Pack<H, S>;
  constexpr int NW = MaskWords<H>::NW;
  constexpr int BASE = SLOT * P::N;
  float h[(H + 3) / 4 * 4];
  uint32_t neg[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) neg[w] = 0u;
  const f2 tt = bc(t);
  // hidden layer: two unit pairs per 16-byte uniform load
  static_for<0, (H + 3) / 4>([&](auto I) {
    constexpr int qq = decltype(I)::value;
    f2 w0, w1;
    ldc_pair2<BASE + P::W1T + 4 * qq>(w0, w1);
    float p[4];
    unpk(fma2(w0, tt, c2[2 * qq]), p[0], p[1]);
    if constexpr (2 * qq + 1 < (H + 1) / 2) unpk(fma2(w1, tt, c2[2 * qq + 1]), p[2], p[3]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = 4 * qq + r;
      if (j < H) {
        if (MASK) neg[j / 32] = __funnelshift_l(__float_as_uint(p[r]), neg[j / 32], 1);
        h[j] = fmaxf(p[r], 0.0f);
      }
    }
  });
  if (MASK) {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int nw = (w == NW - 1) ? (H - 32 * w) : 32;
      const uint32_t low = (nw == 32) ? 0xffffffffu : ((1u << nw) - 1u);
      gate[w] = (~neg[w]) & low;
    }
  }
  f2 acc[(S + 1) / 2 * 2];
  static_for<0, (S + 1) / 2>([&](auto I) {
    constexpr int q = decltype(I)::value;
    ldc_pair2<BASE + P::BH + 4 * q>(acc[2 * q], acc[2 * q + 1]);
  });
  // heads: flat pair index pi = j*S + op; two pairs per 16-byte uniform load
  constexpr int NPAIR = H * S;
  static_for<0, NPAIR / 2>([&](auto I) {
    constexpr int q = decltype(I)::value;
    constexpr int pa = 2 * q, pb = 2 * q + 1;
    f2 wa, wb2;
    ldc_pair2<BASE + P::WH + 4 * q>(wa, wb2);
    acc[pa % S] = fma2(bc(h[pa / S]), wa, acc[pa % S]);
    acc[pb % S] = fma2(bc(h[pb / S]), wb2, acc[pb % S]);
  });
  if constexpr (NPAIR % 2 == 1) {
    constexpr int pa = NPAIR - 1;
    f2 wa;
    ldc_pair1<BASE + P::WH + 2 * pa>(wa);
    acc[pa % S] = fma2(bc(h[pa / S]), wa, acc[pa % S]);
  }
  const f2 one = bc(1.0f);
#pragma unroll
  for (int op = 0; op < S; ++op) {
    float v0, v1;
    unpk(acc[op], v0, v1);
    const f2 e = add2(pk(ex2_approx(v0), ex2_approx(v1)), one);
    unpk(e, v0, v1);
    out.v[op] = pk(rcp_approx(v0), rcp_approx(v1));
  }
}

// Scheduling fence: the RHS evaluations of one step do not depend on each other (the MLP sees only
// t), so ptxas would interleave all of them and blow the register budget.  Making the next
// evaluation's time nominally depend on the previous evaluation's outputs serialises them.
template <int S>
__device__ __forceinline__ float after(float t, const Sig<S>& dep) {
  asm volatile("" : "+f"(t) : "l"(dep.v[S - 1]), "l"(dep.v[0]));
  return t;
}

// f = A - D*x
template <int S>
__device__ __forceinline__ Vec<S> rhs(const Sig<S>& e, const Vec<S>& x) { return vnfma<S>(e.D(), x, e.A()); }

#ifndef SLODE_FWD_MINB
#define SLODE_FWD_MINB 4
#endif
constexpr int kBlock = 128;

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int H, int S, int METHOD>
__global__ void __launch_bounds__(kBlock, SLODE_FWD_MINB)
mlp_fixed_fwd_kernel(int64_t B, int T, const float* __restrict__ tgrid, const float* __restrict__ cin,
                     const float* __restrict__ y0, float* __restrict__ sol, int64_t st, int64_t sb,
                     unsigned wzero) {
  uint32_t nogate[MaskWords<H>::NW];
  // Uniform control flow: every thread of the block runs the same tile and time loops (tail
  // threads redo trajectory B-1 with their stores predicated off), so the loop counters and the
  // weight base stay in uniform registers.
  const int64_t ntiles = (B + kBlock - 1) / kBlock;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t b_raw = tile * kBlock + threadIdx.x;
    const bool valid = b_raw < B;
    const int64_t b = valid ? b_raw : B - 1;
    f2 c2[(H + 1) / 2];
#pragma unroll
    for (int jp = 0; jp < (H + 1) / 2; ++jp) {
      const float c0 = ld_stream(cin + b * H + 2 * jp);
      const float c1 = (2 * jp + 1 < H) ? ld_stream(cin + b * H + 2 * jp + 1) : 0.0f;
      c2[jp] = pk(c0, c1);
    }
    float* out = sol + b * sb;
    Vec<S> x = vload<S>(y0 + b * S);
    if (valid) vstore<S>(out, x);
    float t0 = __ldg(tgrid);
    Sig<S> e0;  // rk4: evaluation at the current grid time, carried over from the previous step
    if (METHOD == SLODE_METHOD_RK4) mlp_eval<H, S, false, 4>(t0, c2, e0, nogate);

#pragma unroll 1
    for (int i = 0; i + 1 < T; ++i) {
      const float t1 = __ldg(tgrid + i + 1);
      const float dt = t1 - t0;
      if (METHOD == SLODE_METHOD_EULER) {
        Sig<S> e;
        mlp_eval<H, S, false, 0>(t0, c2, e, nogate);
        x = vaxpy<S>(dt, rhs<S>(e, x), x);
      } else if (METHOD == SLODE_METHOD_MIDPOINT) {
        const float half_dt = 0.5f * dt;
        Sig<S> e, em;
        mlp_eval<H, S, false, 1>(t0, c2, e, nogate);
        const Vec<S> ym = vaxpy<S>(half_dt, rhs<S>(e, x), x);
        mlp_eval<H, S, false, 2>(after<S>(t0 + half_dt, e), c2, em, nogate);
        x = vaxpy<S>(dt, rhs<S>(em, ym), x);
      } else {  // rk4, 3/8 rule (torchdiffeq rk4_alt_step_func)
        Sig<S> e;
        const Vec<S> k1 = rhs<S>(e0, x);
        Vec<S> y = vaxpy<S>(dt * kOneThird, k1, x);
        mlp_eval<H, S, false, 3>(after<S>(t0 + dt * kOneThird, e0), c2, e, nogate);
        const Vec<S> k2 = rhs<S>(e, y);
        y = vaxpy<S>(dt, vaxpy<S>(-kOneThird, k1, k2), x);
        mlp_eval<H, S, false, 0>(after<S>(t0 + dt * kTwoThirds, e), c2, e, nogate);
        const Vec<S> k3 = rhs<S>(e, y);
        y = vaxpy<S>(dt, vadd<S>(vsub<S>(k1, k2), k3), x);
        mlp_eval<H, S, false, 1>(after<S>(t1, e), c2, e0, nogate);
        const Vec<S> k4 = rhs<S>(e0, y);
        const Vec<S> sum = vadd<S>(vaxpy<S>(3.0f, vadd<S>(k2, k3), k1), k4);
        x = vaxpy<S>(dt * 0.125f, sum, x);
      }
      out += st;
      if (valid) vstore<S>(out, x);
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
  float c[H][kBlock];     // per-thread copy of c for dynamic unit index
  float gc[H][kBlock];    // per-thread dL/dc accumulators
  float2 W[H][S];         // original (unscaled) head weights, [unit][output pair] in device output order
  float w1t[H];
  float G[K2][H];         // block accumulators, [output (device order)][unit]
  float gw1t[H];
  float gb[K2];
};

template <int H, int S>
struct Sweep {
  static constexpr int NW = MaskWords<H>::NW;
  f2 P[S], Q[S];  // pair layout == Sig layout
  uint32_t prev[NW];

  __device__ __forceinline__ void init(const uint32_t (&gate)[NW]) {
#pragma unroll
    for (int k = 0; k < S; ++k) P[k] = Q[k] = 0ull;
#pragma unroll
    for (int w = 0; w < NW; ++w) prev[w] = gate[w];
  }

  __device__ __forceinline__ void snapshot(BwdSmem<H, S>& sm, int j, float sign) {
    const int tid = threadIdx.x;
    const f2 wj = bc(sign * sm.w1t[j]);
    const f2 cj = bc(sign * sm.c[j][tid]);
    f2 s1 = 0ull, s2 = 0ull;
#pragma unroll
    for (int op = 0; op < S; ++op) {
      const float2 w = sm.W[j][op];
      const f2 W = pk(w.x, w.y);
      s1 = fma2(W, P[op], s1);
      s2 = fma2(W, Q[op], s2);
      float g0, g1;
      unpk(fma2(wj, Q[op], mul2(cj, P[op])), g0, g1);
      atomicAdd(&sm.G[2 * op][j], g0);
      atomicAdd(&sm.G[2 * op + 1][j], g1);
    }
    sm.gc[j][tid] += sign * (lo_of(s1) + hi_of(s1));
    atomicAdd(&sm.gw1t[j], sign * (lo_of(s2) + hi_of(s2)));
  }

  // gate flips between the previous contributing evaluation and this one
  __device__ __forceinline__ void events(BwdSmem<H, S>& sm, const uint32_t (&gate)[NW]) {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      uint32_t diff = gate[w] ^ prev[w];
      const int nw = (w == NW - 1) ? (H - 32 * w) : 32;
      while (diff) {
        const int q = __ffs(diff) - 1;
        diff &= diff - 1;
        const float sign = ((prev[w] >> q) & 1u) ? 1.0f : -1.0f;
        snapshot(sm, 32 * w + (nw - 1 - q), sign);
      }
      prev[w] = gate[w];
    }
  }

  // add the cotangents of the head pre-activations of one evaluation at time te:
  //   f = A - D*y with upstream gf:  d(pre_A) = gf*A(1-A),  d(pre_D) = -gf*y*D(1-D)
  __device__ __forceinline__ void add(float te, const Vec<S>& gf, const Vec<S>& y, const Sig<S>& e) {
    const Vec<S> one = vbc<S>(1.0f);
    const Vec<S> A = e.A(), D = e.D();
    const Vec<S> dg = vmul<S>(gf, vmul<S>(A, vsub<S>(one, A)));
    const Vec<S> dd = vmul<S>(vmul<S>(gf, y), vmul<S>(D, vsub<S>(D, one)));
    const f2 tt = bc(te);
#pragma unroll
    VEC_FOR_PAIRS {
      P[2 * q] = add2(P[2 * q], dg.p[q]);
      Q[2 * q] = fma2(dg.p[q], tt, Q[2 * q]);
      P[2 * q + 1] = add2(P[2 * q + 1], dd.p[q]);
      Q[2 * q + 1] = fma2(dd.p[q], tt, Q[2 * q + 1]);
    }
    if (Vec<S>::TAIL) {
      const f2 d = pk(dg.t, dd.t);
      P[S - 1] = add2(P[S - 1], d);
      Q[S - 1] = fma2(d, tt, Q[S - 1]);
    }
  }

  __device__ __forceinline__ void finish(BwdSmem<H, S>& sm) {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      uint32_t act = prev[w];
      const int nw = (w == NW - 1) ? (H - 32 * w) : 32;
      while (act) {
        const int q = __ffs(act) - 1;
        act &= act - 1;
        snapshot(sm, 32 * w + (nw - 1 - q), 1.0f);
      }
    }
#pragma unroll
    for (int op = 0; op < S; ++op) {
      float p0, p1;
      unpk(P[op], p0, p1);
      atomicAdd(&sm.gb[2 * op], p0);
      atomicAdd(&sm.gb[2 * op + 1], p1);
    }
  }
};

template <int H, int S>
__device__ __forceinline__ void load_c2(const BwdSmem<H, S>& sm, f2 (&c2)[(H + 1) / 2]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int jp = 0; jp < (H + 1) / 2; ++jp)
    c2[jp] = pk(sm.c[2 * jp][tid], (2 * jp + 1 < H) ? sm.c[2 * jp + 1][tid] : 0.0f);
}

template <int H, int S, int METHOD, int MODE>
__global__ void __launch_bounds__(kBlock, 3)
mlp_fixed_bwd_kernel(int64_t B, int T, const float* __restrict__ tgrid, const float* __restrict__ cin,
                     const float* __restrict__ w1t, const float* __restrict__ Wg, const float* __restrict__ Wd,
                     const float* __restrict__ sol, int64_t st, int64_t sb,
                     const float* __restrict__ gsol, int64_t gst, int64_t gsb,
                     float* __restrict__ grad_y0, float* __restrict__ grad_c, float* __restrict__ grad_w,
                     unsigned wzero) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem<H, S>& sm = *reinterpret_cast<BwdSmem<H, S>*>(smem_raw);
  constexpr int K2 = 2 * S;
  constexpr int NW = MaskWords<H>::NW;
  const int tid = threadIdx.x;

  for (int i = tid; i < H * K2; i += kBlock) {
    const int j = i / K2, o = i % K2, s = out_state(o, S);
    (&sm.W[0][0].x)[i] = out_is_degr(o, S) ? Wd[s * H + j] : Wg[s * H + j];
  }
  for (int i = tid; i < H; i += kBlock) {
    sm.w1t[i] = w1t[i];
    sm.gw1t[i] = 0.0f;
  }
  for (int i = tid; i < K2 * H; i += kBlock) (&sm.G[0][0])[i] = 0.0f;
  if (tid < K2) sm.gb[tid] = 0.0f;
  __syncthreads();

  const int64_t ntiles = (B + kBlock - 1) / kBlock;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t b_raw = tile * kBlock + tid;
    const bool valid = b_raw < B;  // tail threads redo trajectory B-1 with zero cotangents
    const int64_t b = valid ? b_raw : B - 1;
    {
#pragma unroll
      for (int j = 0; j < H; ++j) {
        sm.c[j][tid] = ld_stream(cin + b * H + j);
        sm.gc[j][tid] = 0.0f;
      }
      f2 c2[(H + 1) / 2];
      load_c2<H, S>(sm, c2);
      const float* xs = sol + b * sb;
      const float* gs = gsol + b * gsb;
      const float live = valid ? 1.0f : 0.0f;
      Vec<S> lam = vscale<S>(vload<S>(gs + (int64_t)(T - 1) * gst), live);

      Sweep<H, S> sw;
      float t1 = __ldg(tgrid + T - 1);
      Sig<S> ec;  // evaluation carried across intervals (rk4: at the shared grid time)
      uint32_t g0[NW], g1[NW], g2[NW], g3[NW];
      bool started = false;
      if (METHOD == SLODE_METHOD_RK4) {
        mlp_eval<H, S, true, 4>(t1, c2, ec, g0);
        sw.init(g0);
        started = true;
      }

#pragma unroll 1
      for (int i = T - 2; i >= 0; --i) {
        const float t0 = __ldg(tgrid + i);
        const Vec<S> x = vload<S>(xs + (int64_t)i * st);
        const Vec<S> gnext = vscale<S>(vload<S>(gs + (int64_t)i * gst), live);

        if (MODE == SLODE_BWD_DISCRETE) {
          const float dt = t1 - t0;
          if (METHOD == SLODE_METHOD_EULER) {
            Sig<S> e;
            mlp_eval<H, S, true, 2>(t0, c2, e, g1);
            const Vec<S> gk = vscale<S>(lam, dt);
            if (!started) { sw.init(g1); started = true; } else sw.events(sm, g1);
            sw.add(t0, gk, x, e);
            lam = vadd<S>(vnfma<S>(gk, e.D(), lam), gnext);
          } else if (METHOD == SLODE_METHOD_MIDPOINT) {
            const float half_dt = 0.5f * dt;
            const float tm = t0 + half_dt;
            Sig<S> e1, e2;
            mlp_eval<H, S, true, 3>(t0, c2, e1, g1);
            mlp_eval<H, S, true, 0>(after<S>(tm, e1), c2, e2, g2);
            const Vec<S> ym = vaxpy<S>(half_dt, rhs<S>(e1, x), x);
            Vec<S> gk = vscale<S>(lam, dt);  // dL/dk2
            if (!started) { sw.init(g2); started = true; } else sw.events(sm, g2);
            sw.add(tm, gk, ym, e2);
            const Vec<S> gy = vnmul<S>(gk, e2.D());  // dL/dy_mid
            lam = vadd<S>(lam, gy);
            gk = vscale<S>(gy, half_dt);  // dL/dk1
            sw.events(sm, g1);
            sw.add(t0, gk, x, e1);
            lam = vadd<S>(vnfma<S>(gk, e1.D(), lam), gnext);
          } else {  // rk4 3/8
            const float ta = t0 + dt * kOneThird;
            const float tb = t0 + dt * kTwoThirds;
            const float dt3 = dt * kOneThird;
            Sig<S> e1, e2, e3;
            mlp_eval<H, S, true, 1>(t0, c2, e1, g1);
            mlp_eval<H, S, true, 2>(after<S>(ta, e1), c2, e2, g2);
            mlp_eval<H, S, true, 3>(after<S>(tb, e2), c2, e3, g3);
            const Vec<S> k1 = rhs<S>(e1, x);
            const Vec<S> y2 = vaxpy<S>(dt3, k1, x);
            const Vec<S> k2 = rhs<S>(e2, y2);
            const Vec<S> y3 = vaxpy<S>(dt, vaxpy<S>(-kOneThird, k1, k2), x);
            const Vec<S> k3 = rhs<S>(e3, y3);
            const Vec<S> y4 = vaxpy<S>(dt, vadd<S>(vsub<S>(k1, k2), k3), x);
            const Vec<S> w = vscale<S>(lam, 0.125f * dt);
            Vec<S> gk1 = w, gk2 = vscale<S>(w, 3.0f), gk3 = gk2;
            // stage 4 (time t1, carried evaluation): gk4 = w
            sw.add(t1, w, y4, ec);
            Vec<S> gy = vnmul<S>(w, ec.D());
            lam = vadd<S>(lam, gy);
            gk1 = vaxpy<S>(dt, gy, gk1);
            gk2 = vaxpy<S>(-dt, gy, gk2);
            gk3 = vaxpy<S>(dt, gy, gk3);
            // stage 3
            sw.events(sm, g3);
            sw.add(tb, gk3, y3, e3);
            gy = vnmul<S>(gk3, e3.D());
            lam = vadd<S>(lam, gy);
            gk2 = vaxpy<S>(dt, gy, gk2);
            gk1 = vaxpy<S>(-dt3, gy, gk1);
            // stage 2
            sw.events(sm, g2);
            sw.add(ta, gk2, y2, e2);
            gy = vnmul<S>(gk2, e2.D());
            lam = vadd<S>(lam, gy);
            gk1 = vaxpy<S>(dt3, gy, gk1);
            // stage 1 (time t0; becomes the carried evaluation of the next interval)
            sw.events(sm, g1);
            sw.add(t0, gk1, x, e1);
            lam = vadd<S>(vnfma<S>(gk1, e1.D(), lam), gnext);
            ec = e1;
          }
        } else {
          // torchdiffeq.odeint_adjoint emulation: one step of the same method on the augmented
          // system [y, a, a_theta] from t1 down to t0, y restarted from the stored sol[i+1].
          // In reversed time s=-t the step is ds = t1 - t0 > 0 with
          //   Ky = D*y - A,  Ka = -a*D,  a_theta += w_m * a_m^T df/dtheta(t_m, y_m).
          const float ds = t1 - t0;
          const Vec<S> y = vload<S>(xs + (int64_t)(i + 1) * st);
          const Vec<S> zero = vbc<S>(0.0f);
          if (METHOD == SLODE_METHOD_EULER) {
            Sig<S> e;
            mlp_eval<H, S, true, 0>(t1, c2, e, g1);
            const Vec<S> v = vscale<S>(lam, ds);
            if (!started) { sw.init(g1); started = true; } else sw.events(sm, g1);
            sw.add(t1, v, y, e);
            lam = vadd<S>(vnfma<S>(v, e.D(), lam), gnext);
          } else if (METHOD == SLODE_METHOD_MIDPOINT) {
            const float half = 0.5f * ds;
            const float tm = t1 - half;
            Sig<S> e1, e2;
            mlp_eval<H, S, false, 1>(t1, c2, e1, g1);
            mlp_eval<H, S, true, 2>(after<S>(tm, e1), c2, e2, g2);
            const Vec<S> ky1 = vsub<S>(zero, rhs<S>(e1, y));  // D1*y - A1
            const Vec<S> ka1 = vnmul<S>(lam, e1.D());         // -a*D1
            const Vec<S> ym = vaxpy<S>(half, ky1, y);
            const Vec<S> am = vaxpy<S>(half, ka1, lam);
            const Vec<S> v = vscale<S>(am, ds);
            if (!started) { sw.init(g2); started = true; } else sw.events(sm, g2);
            sw.add(tm, v, ym, e2);
            lam = vadd<S>(vnfma<S>(v, e2.D(), lam), gnext);
          } else {  // rk4 3/8 on the augmented system
            const float ta = t1 - ds * kOneThird;
            const float tb = t1 - ds * kTwoThirds;
            const float w8 = 0.125f * ds;
            Sig<S> e;
            // stage 1 at t1 (carried evaluation)
            const Vec<S> ky1 = vsub<S>(zero, rhs<S>(ec, y));
            const Vec<S> ka1 = vnmul<S>(lam, ec.D());
            sw.add(t1, vscale<S>(lam, w8), y, ec);
            // stage 2
            mlp_eval<H, S, true, 3>(after<S>(ta, ec), c2, e, g1);
            Vec<S> ym = vaxpy<S>(ds * kOneThird, ky1, y);
            Vec<S> am = vaxpy<S>(ds * kOneThird, ka1, lam);
            const Vec<S> ky2 = vsub<S>(zero, rhs<S>(e, ym));
            const Vec<S> ka2 = vnmul<S>(am, e.D());
            sw.events(sm, g1);
            sw.add(ta, vscale<S>(am, 3.0f * w8), ym, e);
            // stage 3
            mlp_eval<H, S, true, 0>(after<S>(tb, e), c2, e, g1);
            ym = vaxpy<S>(ds, vaxpy<S>(-kOneThird, ky1, ky2), y);
            am = vaxpy<S>(ds, vaxpy<S>(-kOneThird, ka1, ka2), lam);
            const Vec<S> ky3 = vsub<S>(zero, rhs<S>(e, ym));
            const Vec<S> ka3 = vnmul<S>(am, e.D());
            sw.events(sm, g1);
            sw.add(tb, vscale<S>(am, 3.0f * w8), ym, e);
            // stage 4 at t0 (becomes the carried evaluation)
            mlp_eval<H, S, true, 1>(after<S>(t0, e), c2, ec, g1);
            ym = vaxpy<S>(ds, vadd<S>(vsub<S>(ky1, ky2), ky3), y);
            am = vaxpy<S>(ds, vadd<S>(vsub<S>(ka1, ka2), ka3), lam);
            const Vec<S> ka4 = vnmul<S>(am, ec.D());
            sw.events(sm, g1);
            sw.add(t0, vscale<S>(am, w8), ym, ec);
            const Vec<S> asum = vadd<S>(vaxpy<S>(3.0f, vadd<S>(ka2, ka3), ka1), ka4);
            lam = vadd<S>(vaxpy<S>(w8, asum, lam), gnext);
          }
        }
        t1 = t0;
      }

      if (started) sw.finish(sm);
      if (valid) {
        vstore<S>(grad_y0 + b * S, lam);
#pragma unroll
        for (int j = 0; j < H; ++j) grad_c[b * H + j] = sm.gc[j][tid];
      }
    }
  }

  __syncthreads();
  // flush block accumulators: grad_w = [ dw1t (H) | dWg (S*H) | dbg (S) | dWd (S*H) | dbd (S) ]
  for (int i = tid; i < H; i += kBlock) atomicAdd(grad_w + i, sm.gw1t[i]);
  for (int i = tid; i < K2 * H; i += kBlock) {
    const int o = i / H, j = i % H, s = out_state(o, S);
    const int base = out_is_degr(o, S) ? (H + S * H + S) : H;
    atomicAdd(grad_w + base + s * H + j, sm.G[o][j]);
  }
  if (tid < K2) {
    const int s = out_state(tid, S);
    const int base = out_is_degr(tid, S) ? (H + S * H + S + S * H) : (H + S * H);
    atomicAdd(grad_w + base + s, sm.gb[tid]);
  }
}

// ---------------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------------
struct Shape {
  int H, S;
};
// (25,5): CVS / challenge configs; (25,8): proc config; the rest serve tests and the width sweep.
#ifdef SLODE_ONLY_25_5
#define SLODE_SHAPES(X) X(25, 5)
#else
#define SLODE_SHAPES(X) X(25, 5) X(25, 8) X(16, 4) X(32, 5)
#endif


static const Shape kShapes[] = {
#define X(h, s) {h, s},
    SLODE_SHAPES(X)
#undef X
};
constexpr int kNumShapes = sizeof(kShapes) / sizeof(kShapes[0]);

static int upload_pack(PackGuard& g, int H, int S, const float* w1t, const float* Wg, const float* bg,
                       const float* Wd, const float* bd) {
  const int n = kSlots * (((H + 3) / 4 * 4 + (2 * S + 3) / 4 * 4 + H * 2 * S + 3) / 4 * 4);
  if (n > kPackMax) {
    set_error("packed weights (%d floats) exceed the constant buffer", n);
    return SLODE_EUNSUPPORTED;
  }
  pack_kernel<<<1, 256, 0, g.stream>>>(H, S, w1t, Wg, bg, Wd, bd, g.staging);
  SLODE_CUDA_TRY(cudaGetLastError());
  SLODE_CUDA_TRY(cudaMemcpyToSymbolAsync(slode_c_pack, g.staging, sizeof(float) * n, 0, cudaMemcpyDeviceToDevice, g.stream));
  return SLODE_OK;
}

template <int H, int S, int METHOD>
static int launch_fwd(int64_t B, int T, const float* t, const float* c, const float* y0, float* sol, int64_t st,
                      int64_t sb, cudaStream_t stream, int sms) {
  const int64_t tiles = (B + kBlock - 1) / kBlock;
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)sms * 64);
  mlp_fixed_fwd_kernel<H, S, METHOD><<<grid, kBlock, 0, stream>>>(B, T, t, c, y0, sol, st, sb, 0u);
  SLODE_CUDA_TRY(cudaGetLastError());
  return SLODE_OK;
}

template <int H, int S, int METHOD, int MODE>
static int launch_bwd(int64_t B, int T, const float* t, const float* c, const float* w1t, const float* Wg,
                      const float* Wd, const float* sol, int64_t st, int64_t sb, const float* gsol, int64_t gst,
                      int64_t gsb, float* gy0, float* gc, float* gw, cudaStream_t stream, int sms) {
  auto kern = mlp_fixed_bwd_kernel<H, S, METHOD, MODE>;
  const size_t smem = sizeof(BwdSmem<H, S>);
  static int blocks_per_sm = 0;  // per instantiation
  if (blocks_per_sm == 0) {
    SLODE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int n = 0;
    SLODE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kBlock, smem));
    blocks_per_sm = std::max(n, 1);
  }
  const int64_t tiles = (B + kBlock - 1) / kBlock;
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)sms * blocks_per_sm);
  kern<<<grid, kBlock, smem, stream>>>(B, T, t, c, w1t, Wg, Wd, sol, st, sb, gsol, gst, gsb, gy0, gc, gw, 0u);
  SLODE_CUDA_TRY(cudaGetLastError());
  return SLODE_OK;
}

template <int H, int S>
static int fwd_shape(int method, int64_t B, int T, const float* t, const float* c, const float* y0, float* sol,
                     int64_t st, int64_t sb, cudaStream_t stream, int sms) {
  switch (method) {
    case SLODE_METHOD_EULER: return launch_fwd<H, S, SLODE_METHOD_EULER>(B, T, t, c, y0, sol, st, sb, stream, sms);
    case SLODE_METHOD_MIDPOINT: return launch_fwd<H, S, SLODE_METHOD_MIDPOINT>(B, T, t, c, y0, sol, st, sb, stream, sms);
    case SLODE_METHOD_RK4: return launch_fwd<H, S, SLODE_METHOD_RK4>(B, T, t, c, y0, sol, st, sb, stream, sms);
  }
  set_error("fixed-grid forward: unknown method %d", method);
  return SLODE_EINVAL;
}

template <int H, int S, int MODE>
static int bwd_mode(int method, int64_t B, int T, const float* t, const float* c, const float* w1t, const float* Wg,
                    const float* Wd, const float* sol, int64_t st, int64_t sb, const float* gsol, int64_t gst,
                    int64_t gsb, float* gy0, float* gc, float* gw, cudaStream_t stream, int sms) {
#define ARGS B, T, t, c, w1t, Wg, Wd, sol, st, sb, gsol, gst, gsb, gy0, gc, gw, stream, sms
  switch (method) {
    case SLODE_METHOD_EULER: return launch_bwd<H, S, SLODE_METHOD_EULER, MODE>(ARGS);
    case SLODE_METHOD_MIDPOINT: return launch_bwd<H, S, SLODE_METHOD_MIDPOINT, MODE>(ARGS);
    case SLODE_METHOD_RK4: return launch_bwd<H, S, SLODE_METHOD_RK4, MODE>(ARGS);
  }
#undef ARGS
  set_error("fixed-grid backward: unknown method %d", method);
  return SLODE_EINVAL;
}

}  // namespace slode

using namespace slode;

extern "C" int slode_mlp_supported(int H, int S) {
  for (int i = 0; i < kNumShapes; ++i)
    if (kShapes[i].H == H && kShapes[i].S == S) return 1;
  return 0;
}

extern "C" int slode_query(int what) {
  switch (what) {
    case SLODE_Q_VERSION: return 1;
    case SLODE_Q_SM_ARCH: return 100;
    case SLODE_Q_MAX_HIDDEN: {
      int m = 0;
      for (int i = 0; i < kNumShapes; ++i) m = std::max(m, kShapes[i].H);
      return m;
    }
    case SLODE_Q_MAX_STATE: {
      int m = 0;
      for (int i = 0; i < kNumShapes; ++i) m = std::max(m, kShapes[i].S);
      return m;
    }
    case SLODE_Q_N_SHAPES: return kNumShapes;
    case SLODE_Q_FWD_LAUNCHES: return g_fwd_launches;
    case SLODE_Q_BWD_LAUNCHES: return g_bwd_launches;
  }
  if (what >= SLODE_Q_SHAPE_BASE && what < SLODE_Q_SHAPE_BASE + 2 * kNumShapes) {
    const int i = (what - SLODE_Q_SHAPE_BASE) / 2;
    return ((what - SLODE_Q_SHAPE_BASE) & 1) ? kShapes[i].S : kShapes[i].H;
  }
  return -1;
}

static int check_common(const char* who, int64_t B, int T, int H, int S) {
  if (B < 0 || T < 1 || H < 1 || S < 1) {
    set_error("%s: bad sizes B=%lld T=%d H=%d S=%d", who, (long long)B, T, H, S);
    return SLODE_EINVAL;
  }
  if (!slode_mlp_supported(H, S)) {
    set_error("%s: (hidden=%d, state=%d) is not compiled in; there is no generic fallback", who, H, S);
    return SLODE_EUNSUPPORTED;
  }
  return SLODE_OK;
}

extern "C" int slode_mlp_fixed_fwd(int method, int64_t B, int T, int H, int S, const float* t, const float* c,
                                   const float* y0, const float* w1t, const float* Wg, const float* bg,
                                   const float* Wd, const float* bd, float* sol, int64_t sol_stride_t,
                                   int64_t sol_stride_b, void* stream_) {
  int rc = check_common("slode_mlp_fixed_fwd", B, T, H, S);
  if (rc) return rc;
  if (!t || !w1t || !Wg || !bg || !Wd || !bd || (B > 0 && (!c || !y0 || !sol))) {
    set_error("slode_mlp_fixed_fwd: null pointer");
    return SLODE_EINVAL;
  }
  g_fwd_launches = 0;
  if (B == 0) return SLODE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  PackGuard guard(stream);
  if (guard.status) return guard.status;
  rc = upload_pack(guard, H, S, w1t, Wg, bg, Wd, bd);
  if (rc) return rc;
#define X(h, s) \
  if (H == h && S == s) rc = fwd_shape<h, s>(method, B, T, t, c, y0, sol, sol_stride_t, sol_stride_b, stream, guard.sms);
  SLODE_SHAPES(X)
#undef X
  if (rc == SLODE_OK) g_fwd_launches = 2;
  return rc;
}

extern "C" int slode_mlp_fixed_bwd(int method, int mode, int64_t B, int T, int H, int S, const float* t,
                                   const float* c, const float* w1t, const float* Wg, const float* bg,
                                   const float* Wd, const float* bd, const float* sol, int64_t sol_stride_t,
                                   int64_t sol_stride_b, const float* grad_sol, int64_t gsol_stride_t,
                                   int64_t gsol_stride_b, float* grad_y0, float* grad_c, float* grad_w,
                                   void* stream_) {
  int rc = check_common("slode_mlp_fixed_bwd", B, T, H, S);
  if (rc) return rc;
  if (mode != SLODE_BWD_DISCRETE && mode != SLODE_BWD_TDE_ADJOINT) {
    set_error("slode_mlp_fixed_bwd: unknown mode %d", mode);
    return SLODE_EINVAL;
  }
  if (!t || !w1t || !Wg || !bg || !Wd || !bd || !grad_w || (B > 0 && (!c || !sol || !grad_sol || !grad_y0 || !grad_c))) {
    set_error("slode_mlp_fixed_bwd: null pointer");
    return SLODE_EINVAL;
  }
  g_bwd_launches = 0;
  if (B == 0) return SLODE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  PackGuard guard(stream);
  if (guard.status) return guard.status;
  rc = upload_pack(guard, H, S, w1t, Wg, bg, Wd, bd);
  if (rc) return rc;
#define ARGS method, B, T, t, c, w1t, Wg, Wd, sol, sol_stride_t, sol_stride_b, grad_sol, gsol_stride_t, \
             gsol_stride_b, grad_y0, grad_c, grad_w, stream, guard.sms
#define X(h, s)                                                                      \
  if (H == h && S == s)                                                              \
    rc = (mode == SLODE_BWD_DISCRETE) ? bwd_mode<h, s, SLODE_BWD_DISCRETE>(ARGS)     \
                                      : bwd_mode<h, s, SLODE_BWD_TDE_ADJOINT>(ARGS);
  SLODE_SHAPES(X)
#undef X
#undef ARGS
  if (rc == SLODE_OK) g_bwd_launches = 2;
  return rc;
}
