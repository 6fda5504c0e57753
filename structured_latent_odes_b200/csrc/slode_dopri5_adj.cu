// torchdiffeq.odeint_adjoint(func=OdeFunc, y0, t, method="dopri5") -- the BACKWARD pass -- for the SLODE blackbox
// latent ODE, as one persistent cooperative kernel for sm_100a.
//
// Reference call: models/blackbox_ode.py:40-42 (adjoint_solver=True is the shipped default of every config, so
// this is the only way a reference user reaches dopri5 by editing config.solver); algorithm restated in
// oracle/torchdiffeq_oracle.py::_OdeintAdjoint:
//   for i = T-1 .. 1:  one FRESH adaptive dopri5 solve of the augmented system [y, a, a_theta] from t[i] down to
//   t[i-1] (torchdiffeq flips time: s = -t, RHS negated), Hairer initial step per interval, ONE step size for the
//   whole augmented state, error norm = torchdiffeq's mixed norm = the largest per-tensor RMS of err / (atol +
//   rtol max(|y0|,|y1|)) among y, a and each parameter tensor's adjoint, float64 controller time, FSAL inside an
//   interval, the interval's end value taken from the 4th-order interpolant of the step that passes t[i-1]; then
//   y <- the stored forward value sol[i-1], a += grad_sol[i-1].
// In reversed time the augmented right-hand side is
//   dy/ds = -f(t,y),   da/ds = a * df/dy = -a D,   da_theta/ds = sum_b a_b^T df_b/dtheta,      f = G - D y,
// G, D = sigmoid(heads(relu(w1t t + c))), c = z W1[:,1:]^T + b1.  a_theta covers func.parameters() only
// (dynamics_hidden, dyanamics_growth, dyanmics_degradation): OdeFunc.constants gets no gradient.
//
// Mapping.  One thread = one trajectory for the state part.  The parameter adjoint is a reduction over the whole
// batch of six stage derivatives per attempt; only four linear combinations of them are ever needed (the step's
// increment, its error estimate, its dense-output midpoint and the last stage for FSAL), so each warp switches to
// "lane = hidden unit", walks its 32 trajectories with the unit's sums in registers (no cross-lane reductions, no
// atomics: the step sequence is deterministic), and blocks meet at two grid barriers per pass: one before the
// per-element reduction over blocks (each parameter element has one owner thread), one before the norms are
// combined and every thread takes the same accept / reject decision.
#include <algorithm>

#include "slode_common.cuh"
#include "slode_mlp_api.h"

namespace slode {
namespace adj {

constexpr int kT = 128;      // threads per block
// TILE = trajectories per block and pass: 128 (one per thread) for batches that fill the device, 32 for small ones.
// A pass is as long as ONE warp's instruction stream (one warp per scheduler, every pass ends in a grid barrier), and
// 70 % of that stream is the "lane = hidden unit" walk over the warp's trajectories: with TILE = 32 the four warps
// of a block share one warp's 32 trajectories, eight each, instead of walking 32 each (the state part then runs on
// warp 0 alone).  Row stride of the per-trajectory c table: TILE + 1 (conflict-free for lanes over trajectories).
constexpr int kNT = 8;       // tensors of the mixed norm: y, a, W1, b1, Wg, bg, Wd, bd

// Dormand-Prince tableau (torchdiffeq _DORMAND_PRINCE_SHAMPINE_TABLEAU)
__device__ constexpr float kAlpha[6] = {(float)(1.0 / 5), (float)(3.0 / 10), (float)(4.0 / 5), (float)(8.0 / 9), 1.0f, 1.0f};
__device__ constexpr float kBeta[6][6] = {
    {(float)(1.0 / 5), 0, 0, 0, 0, 0},
    {(float)(3.0 / 40), (float)(9.0 / 40), 0, 0, 0, 0},
    {(float)(44.0 / 45), (float)(-56.0 / 15), (float)(32.0 / 9), 0, 0, 0},
    {(float)(19372.0 / 6561), (float)(-25360.0 / 2187), (float)(64448.0 / 6561), (float)(-212.0 / 729), 0, 0},
    {(float)(9017.0 / 3168), (float)(-355.0 / 33), (float)(46732.0 / 5247), (float)(49.0 / 176), (float)(-5103.0 / 18656), 0},
    {(float)(35.0 / 384), 0, (float)(500.0 / 1113), (float)(125.0 / 192), (float)(-2187.0 / 6784), (float)(11.0 / 84)},
};
__device__ constexpr float kCErr[7] = {
    (float)(35.0 / 384 - 1951.0 / 21600), 0.0f, (float)(500.0 / 1113 - 22642.0 / 50085),
    (float)(125.0 / 192 - 451.0 / 720), (float)(-2187.0 / 6784 - -12231.0 / 42400),
    (float)(11.0 / 84 - 649.0 / 6300), (float)(-1.0 / 60.0)};
__device__ constexpr float kCMid[7] = {
    (float)(6025192743.0 / 30085553152.0 / 2), 0.0f, (float)(51252292925.0 / 65400821598.0 / 2),
    (float)(-2691868925.0 / 45128329728.0 / 2), (float)(187940372067.0 / 1594534317056.0 / 2),
    (float)(-1776094331.0 / 19743644256.0 / 2), (float)(11237099.0 / 235043384.0 / 2)};

enum { kStatusOk = 0, kStatusUnderflow = 1, kStatusMaxSteps = 2, kStatusReplayShort = 4 };
enum { kPassF0 = 0, kPassProbe = 1, kPassAttempt = 2 };
enum { kS = 0, kE = 1, kM = 2, kK = 3 };  // combinations: step increment, error, midpoint, last stage alone

struct Args {
  int64_t B;
  int T, L, H;
  const float *t, *z, *c, *W1, *Wg, *bg, *Wd, *bd;
  const float* sol;
  int64_t st, sb;
  const float* gsol;
  int64_t gst, gsb;
  float rtol, atol;
  int64_t max_attempts;
  const double* replay;  // optional (n_replay, 4) rows in the step_log format: take these step sizes and decisions
  int64_t n_replay;
  float *gy0, *gparams;
  double* step_log;  // (log_cap, 4): interval index i, s0 = -t at the start of the attempt, ds, accepted
  int64_t log_cap;
  int64_t* stats;    // accepted, rejected, RHS evaluations per trajectory, status
  // scratch (carved out of the caller's workspace)
  float *y, *a, *k1y, *k1a, *y1, *a1, *k7y, *k7a, *am;  // (B,S)
  float *g, *k1g, *g1, *gm, *k7g;                       // (P)
  float* gpart;                                         // [grid][4][P]
  double* npart;                                        // [2][2 * kNT][grid]
  unsigned long long* barrier;
};

__host__ __device__ inline int n_params(int L, int H, int S) { return H * (L + 1) + H + 2 * (S * H + S); }
// flat offsets of the parameter tensors, in func.parameters() order
struct Layout {
  int W1, b1, Wg, bg, Wd, bd, P;
  __host__ __device__ Layout(int L, int H, int S) {
    W1 = 0;
    b1 = H * (L + 1);
    Wg = b1 + H;
    bg = Wg + S * H;
    Wd = bg + S;
    bd = Wd + S * H;
    P = bd + S;
  }
  __device__ int tensor_of(int e) const { return 2 + (e >= b1) + (e >= Wg) + (e >= bg) + (e >= Wd) + (e >= bd); }
};

struct Smem {
  float *wgd, *w1t, *bgd;  // [H][2S] head weights (growth 0..S-1, degradation S..2S-1), [H], [2S]
  float* ct;               // [H][TILE+1] c_j of the block's trajectories
  float* dl;               // [6][2S][TILE] head cotangents of the pass's stages
  float* zt;               // [TILE][LZ]  latent rows
  float* bacc;             // [4][P]     block sums of the four combinations
  float* coef;             // [4][6]     combination coefficient of every stage slot
  float* ts;               // [6]        evaluation time (t, not s) of every stage slot
  double* red;             // [64]
  int LZ;
};

__host__ __device__ inline int lz_of(int L) { return L | 1; }
__host__ __device__ inline size_t smem_floats(int L, int H, int S, int TILE) {
  return (size_t)H * 2 * S + H + 2 * S + (size_t)H * (TILE + 1) + (size_t)6 * 2 * S * TILE + (size_t)TILE * lz_of(L) +
         (size_t)4 * n_params(L, H, S) + 24 + 8;
}
__host__ __device__ inline size_t smem_bytes(int L, int H, int S, int TILE) {
  return (smem_floats(L, H, S, TILE) * 4 + 15) / 16 * 16 + 64 * sizeof(double);
}

__device__ __forceinline__ void grid_barrier(unsigned long long* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(counter, 1ull);
    const unsigned long long target = (ticket / gridDim.x + 1ull) * gridDim.x;
    while (*reinterpret_cast<volatile unsigned long long*>(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// block sums of NV doubles -> npart[(buf * 2kNT + v) * grid + block]
template <int NV>
__device__ __forceinline__ void block_partials(double (&v)[NV], double* npart, int buf, double* red) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) red[warp * NV + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < kT / 32; ++w) s += red[w * NV + threadIdx.x];
    npart[((size_t)buf * 2 * kNT + threadIdx.x) * gridDim.x + blockIdx.x] = s;
  }
}

// after the grid barrier: every block adds the per-block sums in the same fixed order
template <int NV>
__device__ __forceinline__ void grid_totals(double (&tot)[NV], const double* npart, int buf, double* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  for (int k = warp; k < NV; k += kT / 32) {
    double s = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32)
      s += *reinterpret_cast<const volatile double*>(&npart[((size_t)buf * 2 * kNT + k) * gridDim.x + b]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) red[32 + k] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) tot[k] = red[32 + k];
  __syncthreads();
}

__device__ __forceinline__ float sigmoidf_(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

// one evaluation of the augmented right-hand side (reversed time) for this thread's trajectory at time te;
// the head cotangents go to the stage slot's rows of sm.dl
template <int S, int TILE>
__device__ __forceinline__ void eval_stage(const Smem& sm, int H, int ti, float te, const float (&y)[S],
                                           const float (&a)[S], float (&ky)[S], float (&ka)[S], int slot) {
  float pg[S], pd[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    pg[s] = sm.bgd[s];
    pd[s] = sm.bgd[S + s];
  }
  const float* crow = sm.ct + ti;
  for (int j = 0; j < H; ++j) {
    const float h = fmaxf(fmaf(sm.w1t[j], te, crow[j * (TILE + 1)]), 0.0f);
    const float* w = sm.wgd + j * 2 * S;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      pg[s] = fmaf(w[s], h, pg[s]);
      pd[s] = fmaf(w[S + s], h, pd[s]);
    }
  }
  float* d = sm.dl + (size_t)slot * 2 * S * TILE + ti;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const float G = sigmoidf_(pg[s]), D = sigmoidf_(pd[s]);
    ky[s] = D * y[s] - G;          // -f
    ka[s] = -a[s] * D;             // a * df/dy
    d[s * TILE] = a[s] * (G - G * G);                // a * df/d(pre_G)
    d[(S + s) * TILE] = -a[s] * y[s] * (D - D * D);  // a * df/d(pre_D)
  }
}

template <int S, int TILE>
__global__ void __launch_bounds__(kT, 1) dopri5_adjoint_kernel(Args p) {
  extern __shared__ __align__(16) unsigned char adj_smem[];
  const int H = p.H, L = p.L, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kCT = TILE + 1;
  const int ti = TILE == kT ? tid : lane;          // this thread's trajectory slot in the tile (state part)
  const bool worker = TILE == kT || warp == 0;     // ... which only the first TILE threads work on
  const Layout lay(L, H, S);
  const int P = lay.P;
  Smem sm;
  {
    float* f = reinterpret_cast<float*>(adj_smem);
    sm.wgd = f; f += (size_t)H * 2 * S;
    sm.w1t = f; f += H;
    sm.bgd = f; f += 2 * S;
    sm.ct = f; f += (size_t)H * kCT;
    sm.dl = f; f += (size_t)6 * 2 * S * TILE;
    sm.LZ = lz_of(L);
    sm.zt = f; f += (size_t)TILE * sm.LZ;
    sm.bacc = f; f += (size_t)4 * P;
    sm.coef = f; f += 24;
    sm.ts = f; f += 8;
    sm.red = reinterpret_cast<double*>(adj_smem + (smem_floats(L, H, S, TILE) * 4 + 15) / 16 * 16);
  }
  for (int i = tid; i < H * 2 * S; i += kT) {
    const int j = i / (2 * S), o = i % (2 * S);
    sm.wgd[i] = o < S ? p.Wg[o * H + j] : p.Wd[(o - S) * H + j];
  }
  for (int i = tid; i < H; i += kT) sm.w1t[i] = p.W1[(size_t)i * (L + 1)];
  for (int i = tid; i < 2 * S; i += kT) sm.bgd[i] = i < S ? p.bg[i] : p.bd[i - S];
  __syncthreads();

  const int64_t ntiles = (p.B + TILE - 1) / TILE;
  const bool resident = ntiles <= (int64_t)gridDim.x;  // every block owns at most one tile: its tables stay loaded
  const size_t nthreads = (size_t)gridDim.x * kT, gthread = (size_t)blockIdx.x * kT + tid;
  const double n_of[kNT] = {(double)p.B * S, (double)p.B * S, (double)H * (L + 1), (double)H,
                            (double)S * H,   (double)S,       (double)S * H,       (double)S};

  auto load_tables = [&](int64_t tile) {   // all threads: consecutive threads read consecutive floats of the tile's rows
    for (int i = tid; i < TILE * H; i += kT) {
      const int r = i / H, j = i - r * H;
      sm.ct[j * kCT + r] = __ldg(p.c + min(tile * TILE + r, p.B - 1) * H + j);
    }
    for (int i = tid; i < TILE * L; i += kT) {
      const int r = i / L, l = i - r * L;
      sm.zt[r * sm.LZ + l] = __ldg(p.z + min(tile * TILE + r, p.B - 1) * L + l);
    }
  };
  if (resident && blockIdx.x < ntiles) load_tables(blockIdx.x);
  for (int e = (int)gthread; e < P; e += (int)nthreads) p.g[e] = 0.0f;
  __syncthreads();

  // ---- "lane = hidden unit" pass of one tile: the warp's 32 trajectories x the pass's stage slots -> the four
  // combinations of every parameter element of the lane's units, added to the block sums in warp order -------
  auto unit_pass = [&](int nst) {
    constexpr int LC = S <= 5 ? 16 : 8;  // latent columns per trip (registers)
    // the trajectories this warp walks: its own 32 (TILE = 128) or its quarter of the tile's 32 (TILE = 32)
    const int tb = TILE == kT ? warp * 32 : 0;
    const int bb0 = TILE == kT ? 0 : warp * (32 / (kT / 32)), bb1 = TILE == kT ? 32 : bb0 + 32 / (kT / 32);
    for (int j0 = 0; j0 < H; j0 += 32) {
      const int j = j0 + lane;
      const bool act = j < H;
      float w[2 * S];
#pragma unroll
      for (int o = 0; o < 2 * S; ++o) w[o] = act ? sm.wgd[j * 2 * S + o] : 0.0f;
      const float w1 = act ? sm.w1t[j] : 0.0f;
      for (int l0 = 0; l0 < L || l0 == 0; l0 += LC) {
        const bool head = l0 == 0;
        float aW[4][2 * S], aw1t[4], ab1[4], aZ[4][LC];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          aw1t[k] = ab1[k] = 0.0f;
#pragma unroll
          for (int o = 0; o < 2 * S; ++o) aW[k][o] = 0.0f;
#pragma unroll
          for (int l = 0; l < LC; ++l) aZ[k][l] = 0.0f;
        }
        for (int bb = bb0; bb < bb1; ++bb) {
          const float cj = act ? sm.ct[j * kCT + tb + bb] : 0.0f;
          float dc[4] = {0.0f, 0.0f, 0.0f, 0.0f}, dct[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          for (int i = 0; i < nst; ++i) {
            const float te = sm.ts[i];
            const float pre = fmaf(w1, te, cj);
            const float h = fmaxf(pre, 0.0f);
            const float* d = sm.dl + (size_t)i * 2 * S * TILE + tb + bb;
            float v = 0.0f;
            const float c0 = sm.coef[kS * 6 + i], c1 = sm.coef[kE * 6 + i], c2 = sm.coef[kM * 6 + i],
                        c3 = sm.coef[kK * 6 + i];
#pragma unroll
            for (int o = 0; o < 2 * S; ++o) {
              const float dd = d[o * TILE];
              v = fmaf(w[o], dd, v);
              if (head) {
                const float u = dd * h;
                aW[0][o] = fmaf(c0, u, aW[0][o]);
                aW[1][o] = fmaf(c1, u, aW[1][o]);
                aW[2][o] = fmaf(c2, u, aW[2][o]);
                aW[3][o] = fmaf(c3, u, aW[3][o]);
              }
            }
            const float dci = pre > 0.0f ? v : 0.0f;  // relu'(0) = 0 as in torch
            dc[0] = fmaf(c0, dci, dc[0]);
            dc[1] = fmaf(c1, dci, dc[1]);
            dc[2] = fmaf(c2, dci, dc[2]);
            dc[3] = fmaf(c3, dci, dc[3]);
            const float dt_ = dci * te;
            dct[0] = fmaf(c0, dt_, dct[0]);
            dct[1] = fmaf(c1, dt_, dct[1]);
            dct[2] = fmaf(c2, dt_, dct[2]);
            dct[3] = fmaf(c3, dt_, dct[3]);
          }
          if (head) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ab1[k] += dc[k];
              aw1t[k] += dct[k];
            }
          }
          const float* zr = sm.zt + (size_t)(tb + bb) * sm.LZ + l0;
#pragma unroll
          for (int l = 0; l < LC; ++l) {
            if (l0 + l < L) {
              const float zz = zr[l];
#pragma unroll
              for (int k = 0; k < 4; ++k) aZ[k][l] = fmaf(dc[k], zz, aZ[k][l]);
            }
          }
        }
        // block sums, warps in turn (fixed order: deterministic)
        for (int ww = 0; ww < kT / 32; ++ww) {
          if (ww == warp && act) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float* acc = sm.bacc + (size_t)k * P;
              if (head) {
                acc[lay.W1 + j * (L + 1)] += aw1t[k];
                acc[lay.b1 + j] += ab1[k];
#pragma unroll
                for (int o = 0; o < S; ++o) {
                  acc[lay.Wg + o * H + j] += aW[k][o];
                  acc[lay.Wd + o * H + j] += aW[k][S + o];
                }
              }
#pragma unroll
              for (int l = 0; l < LC; ++l)
                if (l0 + l < L) acc[lay.W1 + j * (L + 1) + 1 + l0 + l] += aZ[k][l];
            }
          }
          __syncthreads();
        }
      }
    }
    // head biases: lane o < 2S sums its cotangent over the warp's trajectories
    {
      float ab[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      if (lane < 2 * S) {
        for (int i = 0; i < nst; ++i) {
          const float* d = sm.dl + ((size_t)i * 2 * S + lane) * TILE + tb;
          float s = 0.0f;
          for (int bb = bb0; bb < bb1; ++bb) s += d[bb];
#pragma unroll
          for (int k = 0; k < 4; ++k) ab[k] = fmaf(sm.coef[k * 6 + i], s, ab[k]);
        }
      }
      for (int ww = 0; ww < kT / 32; ++ww) {
        if (ww == warp && lane < 2 * S) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            sm.bacc[(size_t)k * P + (lane < S ? lay.bg + lane : lay.bd + lane - S)] += ab[k];
        }
        __syncthreads();
      }
    }
  };

  // ---- controller state (identical in every thread) ------------------------------------------------------
  int64_t attempt = 0, n_acc = 0, n_rej = 0, n_rhs = 0;
  int status = kStatusOk;
  int nbuf = 0;

  // one pass over the trajectories + the reduction over blocks.  kind: F0 (right-hand side at the interval's
  // start), Probe (Euler probe of the initial-step selection), Attempt (six new stages).  Returns the norms.
  auto run_pass = [&](int kind, int iv, float s0f, float dsf, float s1f, double (&norms)[2 * kNT]) {
    const int nst = kind == kPassAttempt ? 6 : 1;
    if (tid < 24) sm.coef[tid] = 0.0f;
    __syncthreads();
    if (tid == 0) {
      if (kind == kPassAttempt) {
        for (int i = 0; i < 6; ++i) {
          sm.ts[i] = -(i < 4 ? __fadd_rn(s0f, __fmul_rn(kAlpha[i], dsf)) : s1f);
          sm.coef[kS * 6 + i] = i < 5 ? __fmul_rn(kBeta[5][i + 1], dsf) : 0.0f;
          sm.coef[kE * 6 + i] = __fmul_rn(kCErr[i + 1], dsf);
          sm.coef[kM * 6 + i] = __fmul_rn(kCMid[i + 1], dsf);
        }
        sm.coef[kK * 6 + 5] = 1.0f;
      } else {
        sm.ts[0] = kind == kPassF0 ? -s0f : -__fadd_rn(s0f, dsf);
        sm.coef[kK * 6] = 1.0f;
      }
    }
    for (int e = tid; e < 4 * P; e += kT) sm.bacc[e] = 0.0f;
    __syncthreads();
    double part[2 * kNT];
#pragma unroll
    for (int k = 0; k < 2 * kNT; ++k) part[k] = 0.0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      if (!resident) {
        __syncthreads();
        load_tables(tile);
        __syncthreads();
      }
      const int64_t br = tile * TILE + ti;
      const bool ok = worker && br < p.B;
      const int64_t b = br < p.B ? br : p.B - 1;
      float y0[S], a0[S];
      if (!worker) {
        // (TILE = 32: warps 1..3 only take part in the unit pass below)
      } else if (kind == kPassF0) {
        // interval start: y <- the stored forward value, a <- the carried adjoint + the output cotangent
#pragma unroll
        for (int s = 0; s < S; ++s) {
          y0[s] = __ldg(p.sol + (int64_t)iv * p.st + b * p.sb + s);
          const float gi = __ldg(p.gsol + (int64_t)iv * p.gst + b * p.gsb + s);
          a0[s] = ok ? (iv == p.T - 1 ? gi : p.a[b * S + s] + gi) : 0.0f;
        }
        float ky[S], ka[S];
        eval_stage<S, TILE>(sm, H, ti, sm.ts[0], y0, a0, ky, ka, 0);
        if (ok) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            p.y[b * S + s] = y0[s];
            p.a[b * S + s] = a0[s];
            p.k1y[b * S + s] = ky[s];
            p.k1a[b * S + s] = ka[s];
            const float sy = fmaf(fabsf(y0[s]), p.rtol, p.atol), sa = fmaf(fabsf(a0[s]), p.rtol, p.atol);
            const float r0 = __fdiv_rn(y0[s], sy), r1 = __fdiv_rn(a0[s], sa);
            const float q0 = __fdiv_rn(ky[s], sy), q1 = __fdiv_rn(ka[s], sa);
            part[0] += (double)(r0 * r0);
            part[1] += (double)(r1 * r1);
            part[kNT + 0] += (double)(q0 * q0);
            part[kNT + 1] += (double)(q1 * q1);
          }
        }
      } else if (kind == kPassProbe) {
        float f0y[S], f0a[S], y1[S], a1[S], ky[S], ka[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          y0[s] = p.y[b * S + s];
          a0[s] = ok ? p.a[b * S + s] : 0.0f;
          f0y[s] = p.k1y[b * S + s];
          f0a[s] = ok ? p.k1a[b * S + s] : 0.0f;
          y1[s] = fmaf(dsf, f0y[s], y0[s]);
          a1[s] = fmaf(dsf, f0a[s], a0[s]);
        }
        eval_stage<S, TILE>(sm, H, ti, sm.ts[0], y1, a1, ky, ka, 0);
        if (ok) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float sy = fmaf(fabsf(y0[s]), p.rtol, p.atol), sa = fmaf(fabsf(a0[s]), p.rtol, p.atol);
            const float q0 = __fdiv_rn(ky[s] - f0y[s], sy), q1 = __fdiv_rn(ka[s] - f0a[s], sa);
            part[0] += (double)(q0 * q0);
            part[1] += (double)(q1 * q1);
          }
        }
      } else {
        float ky[7][S], ka[7][S], yi[S], ai[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          y0[s] = p.y[b * S + s];
          a0[s] = ok ? p.a[b * S + s] : 0.0f;
          ky[0][s] = p.k1y[b * S + s];
          ka[0][s] = ok ? p.k1a[b * S + s] : 0.0f;
        }
#pragma unroll
        for (int i = 1; i <= 6; ++i) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float sy = 0.0f, sa = 0.0f;
#pragma unroll
            for (int j = 0; j < i; ++j) {
              if (kBeta[i - 1][j] != 0.0f) {
                const float cj = __fmul_rn(kBeta[i - 1][j], dsf);
                sy = fmaf(ky[j][s], cj, sy);
                sa = fmaf(ka[j][s], cj, sa);
              }
            }
            yi[s] = y0[s] + sy;
            ai[s] = a0[s] + sa;
          }
          eval_stage<S, TILE>(sm, H, ti, sm.ts[i - 1], yi, ai, ky[i], ka[i], i - 1);
        }
        if (ok) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float ey = 0.0f, ea = 0.0f, ma = 0.0f;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
              if (kCErr[j] != 0.0f) {
                const float cj = __fmul_rn(kCErr[j], dsf);
                ey = fmaf(ky[j][s], cj, ey);
                ea = fmaf(ka[j][s], cj, ea);
              }
              if (kCMid[j] != 0.0f) ma = fmaf(ka[j][s], __fmul_rn(kCMid[j], dsf), ma);
            }
            const float ty = fmaf(fmaxf(fabsf(y0[s]), fabsf(yi[s])), p.rtol, p.atol);
            const float ta = fmaf(fmaxf(fabsf(a0[s]), fabsf(ai[s])), p.rtol, p.atol);
            const float r0 = __fdiv_rn(ey, ty), r1 = __fdiv_rn(ea, ta);
            part[0] += (double)(r0 * r0);
            part[1] += (double)(r1 * r1);
            p.y1[b * S + s] = yi[s];
            p.a1[b * S + s] = ai[s];
            p.k7y[b * S + s] = ky[6][s];
            p.k7a[b * S + s] = ka[6][s];
            p.am[b * S + s] = a0[s] + ma;
          }
        }
      }
      __syncthreads();  // the stage cotangents of the whole tile are in shared memory
      unit_pass(nst);
    }
    // block sums -> global, then every parameter element is reduced over the blocks by its owner thread
    for (int e = tid; e < 4 * P; e += kT) p.gpart[(size_t)blockIdx.x * 4 * P + e] = sm.bacc[e];
    grid_barrier(p.barrier);
    for (int e = (int)gthread; e < P; e += (int)nthreads) {
      float tot[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int blk = 0; blk < (int)gridDim.x; ++blk) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (kind == kPassAttempt || k == kK)
            tot[k] += *reinterpret_cast<const volatile float*>(&p.gpart[((size_t)blk * 4 + k) * P + e]);
        }
      }
      const int tn = lay.tensor_of(e);
      const float g0 = p.g[e];
      if (kind == kPassF0) {
        p.k1g[e] = tot[kK];
        const float sc = fmaf(fabsf(g0), p.rtol, p.atol);
        const float r0 = __fdiv_rn(g0, sc), q0 = __fdiv_rn(tot[kK], sc);
        part[tn] += (double)(r0 * r0);
        part[kNT + tn] += (double)(q0 * q0);
      } else if (kind == kPassProbe) {
        const float sc = fmaf(fabsf(g0), p.rtol, p.atol);
        const float q0 = __fdiv_rn(tot[kK] - p.k1g[e], sc);
        part[tn] += (double)(q0 * q0);
      } else {
        const float k1 = p.k1g[e];
        const float g1 = g0 + fmaf(k1, __fmul_rn(kBeta[5][0], dsf), tot[kS]);
        const float er = fmaf(k1, __fmul_rn(kCErr[0], dsf), tot[kE]);
        const float gm = g0 + fmaf(k1, __fmul_rn(kCMid[0], dsf), tot[kM]);
        p.g1[e] = g1;
        p.gm[e] = gm;
        p.k7g[e] = tot[kK];
        const float tl = fmaf(fmaxf(fabsf(g0), fabsf(g1)), p.rtol, p.atol);
        const float r0 = __fdiv_rn(er, tl);
        part[tn] += (double)(r0 * r0);
      }
    }
    block_partials<2 * kNT>(part, p.npart, nbuf, sm.red);
    grid_barrier(p.barrier);
    double tot[2 * kNT];
    grid_totals<2 * kNT>(tot, p.npart, nbuf, sm.red);
    nbuf ^= 1;
#pragma unroll
    for (int k = 0; k < 2 * kNT; ++k) norms[k] = tot[k];
    n_rhs += nst;
  };
  // torchdiffeq's mixed norm of one set of per-tensor sums of squares
  auto mixed = [&](const double* sums) {
    float m = 0.0f;
    for (int k = 0; k < kNT; ++k) m = fmaxf(m, sqrtf((float)(sums[k] / n_of[k])));
    return m;
  };
  // 4th-order dense output (torchdiffeq _interp_fit / _interp_evaluate) at relative position x
  auto interp = [](float v0, float v1, float vm, float f0, float f1, float ds, float x) {
    const float ca = 2.0f * ds * (f1 - f0) - 8.0f * (v1 + v0) + 16.0f * vm;
    const float cb = ds * (5.0f * f0 - 3.0f * f1) + 18.0f * v0 + 14.0f * v1 - 32.0f * vm;
    const float cc = ds * (f1 - 4.0f * f0) - 11.0f * v0 - 5.0f * v1 + 16.0f * vm;
    const float cd = ds * f0;
    float tot = v0 + x * cd;
    float xp = x * x;
    tot = tot + xp * cc;
    xp = xp * x;
    tot = tot + xp * cb;
    xp = xp * x;
    tot = tot + xp * ca;
    return tot;
  };

  for (int iv = p.T - 1; iv >= 1 && status == kStatusOk; --iv) {
    const double s_start = -(double)__ldg(p.t + iv), s_end = -(double)__ldg(p.t + iv - 1);
    double norms[2 * kNT];
    // ---- f0 and Hairer's initial step (torchdiffeq _select_initial_step, order 4) -------------------------
    run_pass(kPassF0, iv, (float)s_start, 0.0f, 0.0f, norms);
    const float d0 = mixed(norms), d1 = mixed(norms + kNT);
    const float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : __fdiv_rn(__fmul_rn(0.01f, d0), d1);
    run_pass(kPassProbe, iv, (float)s_start, h0, 0.0f, norms);
    const float d2 = __fdiv_rn(mixed(norms), h0);
    float h1;
    if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, __fmul_rn(h0, 1e-3f));
    else h1 = powf(__fdiv_rn(0.01f, fmaxf(d1, d2)), 1.0f / 5.0f);
    double ds = (double)fminf(__fmul_rn(100.0f, h0), h1);
    double s_cur = s_start;
    // ---- adaptive steps until one passes the interval's end ------------------------------------------------
    while (s_end > s_cur) {
      if (attempt >= p.max_attempts) { status = kStatusMaxSteps; break; }
      if (p.replay) {  // prescribed step sequence (parity tests against a reference step log)
        if (attempt >= p.n_replay) { status = kStatusReplayShort; break; }
        ds = p.replay[attempt * 4 + 2];
      }
      const double a_s0 = s_cur, a_ds = ds, a_s1 = a_s0 + a_ds;
      if (!(a_s1 > a_s0)) { status = kStatusUnderflow; break; }
      const float s0f = (float)a_s0, dsf = (float)a_ds, s1f = (float)a_s1;
      run_pass(kPassAttempt, iv, s0f, dsf, s1f, norms);
      const float ratio = mixed(norms);
      const bool accept = p.replay ? (p.replay[attempt * 4 + 3] != 0.0) : (ratio <= 1.0f);
      if (blockIdx.x == 0 && tid == 0 && p.step_log && attempt < p.log_cap) {
        p.step_log[attempt * 4 + 0] = (double)iv;
        p.step_log[attempt * 4 + 1] = a_s0;
        p.step_log[attempt * 4 + 2] = a_ds;
        p.step_log[attempt * 4 + 3] = accept ? 1.0 : 0.0;
      }
      ++attempt;
      if (accept) {
        ++n_acc;
        s_cur = a_s1;
        const bool done = !(s_end > s_cur);
        const float x = __fdiv_rn(__fsub_rn((float)s_end, s0f), __fsub_rn(s1f, s0f));
        // commit: every trajectory / parameter element by its owner thread
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
          const int64_t b = tile * TILE + ti;
          if (worker && b < p.B) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const int64_t o = b * S + s;
              if (done) {
                p.a[o] = interp(p.a[o], p.a1[o], p.am[o], p.k1a[o], p.k7a[o], dsf, x);
              } else {
                p.y[o] = p.y1[o];
                p.a[o] = p.a1[o];
                p.k1y[o] = p.k7y[o];
                p.k1a[o] = p.k7a[o];
              }
            }
          }
        }
        for (int e = (int)gthread; e < P; e += (int)nthreads) {
          if (done) {
            p.g[e] = interp(p.g[e], p.g1[e], p.gm[e], p.k1g[e], p.k7g[e], dsf, x);
          } else {
            p.g[e] = p.g1[e];
            p.k1g[e] = p.k7g[e];
          }
        }
      } else {
        ++n_rej;
      }
      // torchdiffeq _optimal_step_size (float64)
      double factor;
      if (ratio == 0.0f) {
        factor = 10.0;
      } else {
        const double dfac = (ratio < 1.0f) ? 1.0 : 0.2;
        factor = fmin(10.0, fmax(0.9 / pow((double)ratio, 0.2), dfac));
      }
      ds = a_ds * factor;
    }
  }
  // ---- outputs: dL/dy0 = a + grad_sol[0], dL/dtheta = a_theta -------------------------------------------------
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t b = tile * TILE + ti;
    if (worker && b < p.B) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float gi = __ldg(p.gsol + b * p.gsb + s);
        p.gy0[b * S + s] = p.T > 1 ? p.a[b * S + s] + gi : gi;
      }
    }
  }
  for (int e = (int)gthread; e < P; e += (int)nthreads) p.gparams[e] = p.g[e];
  if (blockIdx.x == 0 && tid == 0) {
    p.stats[0] = n_acc;
    p.stats[1] = n_rej;
    p.stats[2] = n_rhs;
    p.stats[3] = status;
  }
}

template <int S, int TILE>
static int plan(int L, int H, int sms, int64_t B, int* grid, size_t* smem) {
  *smem = smem_bytes(L, H, S, TILE);
  auto kern = dopri5_adjoint_kernel<S, TILE>;
  int dev = 0, max_optin = 0;
  SLODE_CUDA_TRY(cudaGetDevice(&dev));
  SLODE_CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (*smem > (size_t)max_optin) {
    set_error("dopri5 adjoint: (L=%d, H=%d, S=%d) needs %zu bytes of shared memory per block, the device offers %d", L, H,
              S, *smem, max_optin);
    return SLODE_EUNSUPPORTED;
  }
  SLODE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
  int per_sm = 0;
  SLODE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kT, *smem));
  if (per_sm < 1) {
    set_error("dopri5 adjoint: kernel does not fit an SM");
    return SLODE_ECUDA;
  }
  const int64_t tiles = (B + TILE - 1) / TILE;
  *grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)sms * per_sm));
  return SLODE_OK;
}

static size_t align256(size_t n) { return (n + 255) / 256 * 256; }

static size_t workspace_bytes(int64_t B, int L, int H, int S, int grid) {
  const size_t P = n_params(L, H, S);
  return align256(sizeof(float) * 9 * (size_t)B * S) + align256(sizeof(float) * 5 * P) +
         align256(sizeof(float) * 4 * P * (size_t)grid) + align256(sizeof(double) * 2 * 2 * kNT * (size_t)grid) + 256;
}

template <int S, int TILE>
static int launch_tile(Args a, int sms, void* ws, size_t ws_bytes, bool plan_only, size_t* need, cudaStream_t stream) {
  int grid = 0;
  size_t smem = 0;
  const int rc = plan<S, TILE>(a.L, a.H, sms, a.B, &grid, &smem);
  if (rc) return rc;
  *need = workspace_bytes(a.B, a.L, a.H, S, grid);
  if (plan_only) return SLODE_OK;
  if (!ws || ws_bytes < *need || (reinterpret_cast<uintptr_t>(ws) & 255)) {
    set_error("dopri5 adjoint: workspace of %zu bytes given (256-byte aligned?), %zu needed "
              "(slode_mlp_dopri5_adjoint_workspace_bytes)", ws_bytes, *need);
    return SLODE_EINVAL;
  }
  const size_t P = n_params(a.L, a.H, S), nstate = (size_t)a.B * S;
  char* w = static_cast<char*>(ws);
  float* f = reinterpret_cast<float*>(w);
  a.y = f; a.a = f + nstate; a.k1y = f + 2 * nstate; a.k1a = f + 3 * nstate; a.y1 = f + 4 * nstate;
  a.a1 = f + 5 * nstate; a.k7y = f + 6 * nstate; a.k7a = f + 7 * nstate; a.am = f + 8 * nstate;
  w += align256(sizeof(float) * 9 * nstate);
  f = reinterpret_cast<float*>(w);
  a.g = f; a.k1g = f + P; a.g1 = f + 2 * P; a.gm = f + 3 * P; a.k7g = f + 4 * P;
  w += align256(sizeof(float) * 5 * P);
  a.gpart = reinterpret_cast<float*>(w);
  w += align256(sizeof(float) * 4 * P * (size_t)grid);
  a.npart = reinterpret_cast<double*>(w);
  w += align256(sizeof(double) * 2 * 2 * kNT * (size_t)grid);
  a.barrier = reinterpret_cast<unsigned long long*>(w);
  SLODE_CUDA_TRY(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned long long), stream));
  void* args[] = {&a};
  SLODE_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)dopri5_adjoint_kernel<S, TILE>, dim3(grid), dim3(kT), args,
                                             smem, stream));
  return SLODE_OK;
}

template <int S>
static int launch(const Args& a, int sms, void* ws, size_t ws_bytes, bool plan_only, size_t* need, cudaStream_t stream) {
  // small batches: 32 trajectories per block (four warps share them); batches that give every SM a 128-trajectory
  // block keep one trajectory per thread
  if ((a.B + kT - 1) / kT < sms) return launch_tile<S, 32>(a, sms, ws, ws_bytes, plan_only, need, stream);
  return launch_tile<S, kT>(a, sms, ws, ws_bytes, plan_only, need, stream);
}

static int dispatch(const Args& a, int S, int sms, void* ws, size_t ws_bytes, bool plan_only, size_t* need,
                    cudaStream_t stream) {
  switch (S) {
    case 4: return launch<4>(a, sms, ws, ws_bytes, plan_only, need, stream);
    case 5: return launch<5>(a, sms, ws, ws_bytes, plan_only, need, stream);
    case 8: return launch<8>(a, sms, ws, ws_bytes, plan_only, need, stream);
  }
  set_error("dopri5 adjoint: ode_state_dim=%d is not compiled in (4, 5, 8); there is no generic fallback", S);
  return SLODE_EUNSUPPORTED;
}

}  // namespace adj
}  // namespace slode

using namespace slode;

extern "C" int64_t slode_mlp_dopri5_adjoint_workspace_bytes(int64_t B, int L, int H, int S) {
  if (B < 0 || L < 1 || H < 1 || S < 1) {
    set_error("slode_mlp_dopri5_adjoint_workspace_bytes: bad sizes");
    return -1;
  }
  if (B == 0) return 0;
  int sms = 0;
  if (device_sms(&sms)) return -1;
  adj::Args a{};
  a.B = B; a.L = L; a.H = H;
  size_t need = 0;
  if (adj::dispatch(a, S, sms, nullptr, 0, true, &need, nullptr)) return -1;
  return (int64_t)need;
}

extern "C" int slode_mlp_dopri5_adjoint_bwd(int64_t B, int T, int L, int H, int S, const float* t, const float* z,
                                            const float* c, const float* W1, const float* Wg, const float* bg,
                                            const float* Wd, const float* bd, const float* sol, int64_t sol_stride_t,
                                            int64_t sol_stride_b, const float* grad_sol, int64_t gsol_stride_t,
                                            int64_t gsol_stride_b, double rtol, double atol, int64_t max_attempts,
                                            const double* replay_steps, int64_t n_replay, float* grad_y0, float* grad_params, double* step_log, int64_t log_capacity,
                                            int64_t* stats, void* workspace, int64_t workspace_bytes, void* stream_) {
  if (B < 0 || T < 1 || L < 1 || H < 1 || S < 1) {
    set_error("slode_mlp_dopri5_adjoint_bwd: bad sizes B=%lld T=%d L=%d H=%d S=%d", (long long)B, T, L, H, S);
    return SLODE_EINVAL;
  }
  if (!(rtol >= 0.0) || !(atol >= 0.0) || (rtol == 0.0 && atol == 0.0) || max_attempts < 1 || log_capacity < 0) {
    set_error("slode_mlp_dopri5_adjoint_bwd: bad tolerances / capacities (rtol=%g atol=%g max_attempts=%lld)", rtol, atol,
              (long long)max_attempts);
    return SLODE_EINVAL;
  }
  if (!t || !W1 || !Wg || !bg || !Wd || !bd || !stats || !grad_params ||
      (B > 0 && (!z || !c || !sol || !grad_sol || !grad_y0))) {
    set_error("slode_mlp_dopri5_adjoint_bwd: null pointer");
    return SLODE_EINVAL;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  g_bwd_launches = 0;
  if (B == 0) {
    SLODE_CUDA_TRY(cudaMemsetAsync(stats, 0, 4 * sizeof(int64_t), stream));
    SLODE_CUDA_TRY(cudaMemsetAsync(grad_params, 0, sizeof(float) * adj::n_params(L, H, S), stream));
    return SLODE_OK;
  }
  int sms = 0;
  int rc = device_sms(&sms);
  if (rc) return rc;
  adj::Args a{};
  a.B = B; a.T = T; a.L = L; a.H = H;
  a.t = t; a.z = z; a.c = c; a.W1 = W1; a.Wg = Wg; a.bg = bg; a.Wd = Wd; a.bd = bd;
  a.sol = sol; a.st = sol_stride_t; a.sb = sol_stride_b;
  a.gsol = grad_sol; a.gst = gsol_stride_t; a.gsb = gsol_stride_b;
  a.rtol = (float)rtol; a.atol = (float)atol; a.max_attempts = max_attempts;
  a.replay = replay_steps; a.n_replay = replay_steps ? n_replay : 0;
  a.gy0 = grad_y0; a.gparams = grad_params; a.step_log = step_log; a.log_cap = step_log ? log_capacity : 0;
  a.stats = stats;
  size_t need = 0;
  rc = adj::dispatch(a, S, sms, workspace, (size_t)workspace_bytes, false, &need, stream);
  if (rc == SLODE_OK) g_bwd_launches = 1;
  return rc;
}
