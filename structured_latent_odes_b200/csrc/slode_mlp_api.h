// Host-side interface between the C-ABI dispatchers (slode_mlp.cu, slode_fixed_api.cu) and the per-shape translation
// units: slode_fixed_<H>_<S>.cu (fixed-grid kernels) and slode_mlp_<H>_<S>.cu (dopri5 kernels, each of which owns its
// packed-weight buffer in constant memory).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace slode {

struct PackSrc {  // device pointers, torch Linear layouts
  const float *w1t, *Wg, *bg, *Wd, *bd;
};

// Optional fused prologue / epilogue.  With z given, the kernels compute
// c = z W1[:,1:]^T + b1 themselves (and, backwards, dz, dW1[:,1:], db1); with the latent_to_ode_net weights given
// as well, also x0 = sigmoid(Wb relu(Wa z + ba) + bb) (models/blackbox_ode.py:19-22,32-34) and its gradients.
struct LatentSrc {
  const float* z;                  // (B,L) row-major, or null: c / y0 are given instead
  int L;
  const float *W1, *b1;            // dynamics_hidden.weight (H, L+1) / bias (H)
  const float *Wa, *ba, *Wb, *bb;  // latent_to_ode_net Linear(L,H) / Linear(H,S); null: y0 is given
};

// Optional fused decoder-head epilogue of the forward solve: with W given the kernel
// writes mu_q = Linear_q(solution).permute(0,2,1) (models/decoders.py:45-47, :86) for all heads straight from the
// states it holds in registers; sol itself then becomes optional (null: the trajectories never reach HBM).
struct HeadsSrc {
  const float* W;   // (NQ,O,S) stacked bias-free Linear(S -> O) weights, or null: no fused heads
  float* mu;        // (NQ,B,O,pitch): element (q,b,o,t) at ((q B + b) O + o) pitch + t
  int NQ, O;
  int64_t pitch;    // row pitch in floats, >= T (a multiple of 8 keeps every row's 32-byte sectors in phase)
};
constexpr int kMaxHeadW = 3 * 8 * 8;  // NQ * O * S floats staged per block

struct FwdArgs {
  int method;
  int64_t B;
  int T;
  const float *t, *c, *y0;
  float* sol;
  int64_t st, sb;
  cudaStream_t stream;
  int sms;
  LatentSrc lat;
  void* ws;          // caller-provided scratch (slode_fixed_workspace_bytes), may be null when 0 bytes are needed
  size_t ws_bytes;
  HeadsSrc heads;    // fused decoder heads (W null: off); with them sol may be null
};

struct BwdArgs {
  int method, mode;
  int64_t B;
  int T;
  const float *t, *c, *w1t, *Wg, *Wd, *sol;
  int64_t st, sb;
  const float* gsol;
  int64_t gst, gsb;
  float *gy0, *gc, *gw;
  cudaStream_t stream;
  int sms;
  LatentSrc lat;
  float* gz;  // (B,L), fused mode only
  void* ws;                // caller-provided scratch: flip records (+ wide-layer tables)
  size_t ws_bytes;
};

// Controller state of a dopri5 solve that is advanced one pass per launch (stepper mode: trajectory-sharded solves
// whose batch-wide sums are combined across shards by the host between launches)
struct Dopri5Ctrl {
  int phase, status, out_idx, prev_acc, emit_lo, emit_hi, buf, pad_;
  long long attempt, n_acc, n_rej, n_rhs;
  double t_cur, dt, pv_t0, pv_t1, pv_dt, a_t0, a_dt;
  float d1, h0;
};
enum { kPhStart = 0, kPhProbe = 1, kPhFirstDt = 2, kPhAttempt = 3, kPhDecide = 4, kPhDone = 5 };

// dopri5 forward (slode_dopri5_kernels.cuh); scratch pointers are filled in by the launcher
struct Dopri5Args {
  int64_t B;
  int T;
  const float *t, *c, *y0;
  float* sol;
  int64_t st, sb;
  float rtol, atol;
  double first_step;      // <= 0: Hairer's selection
  const double* replay;   // optional (n_replay,3) prescribed (t0, dt, accepted) sequence
  int64_t n_replay;
  int64_t max_attempts;
  // scratch (library-owned)
  float *ys, *fs, *cy1, *cf1, *cym;    // (B,S) each
  double* partial;                      // [2][3][grid]
  unsigned long long* barrier;          // monotonic arrival counter (zeroed before the launch)
  // optional outputs
  float* ckpt_y;        // (ckpt_cap, B, S): state at the start of every accepted step (for the reverse sweep)
  int64_t ckpt_cap;
  double* step_log;     // (log_cap, 3): t0, dt, accepted(1/0) of every attempted step
  int64_t log_cap;
  int64_t* stats;       // [0] accepted, [1] rejected, [2] RHS evaluations per trajectory, [3] status (0 ok)
                        // (stepper mode: [4] phase after this launch, kPhDone when the solve is over)
  // stepper mode (ctrl != null): caller-owned workspace, one pass per launch
  Dopri5Ctrl* ctrl;
  int restart;             // 1: first launch of a solve (ctrl is initialised by the kernel)
  int64_t n_global;        // trajectories over ALL shards (the norms' element count); 0: B
  const double* ext_sums;  // [2] sums over all shards of the previous launch's out_sums
  double* out_sums;        // [2] this shard's sums of the pass this launch ran
  void* step_ws;           // caller workspace holding the state arrays, partial sums, barrier and ctrl
  size_t step_ws_bytes;
};

// dopri5 reverse sweep; flip_ws is filled in by the launcher
struct Dopri5BwdArgs {
  int64_t B;
  int T;
  const float *t, *c, *w1t, *Wg, *Wd;
  int64_t n_acc;
  const double* acc_steps;   // (n_acc, 2): t0, dt of every accepted step
  const int* emit;           // (n_acc + 1): outputs [emit[n], emit[n+1]) were interpolated inside accepted step n
  const float* ckpt_y;       // (n_acc, B, S)
  const float* gsol;
  int64_t gst, gsb;
  float *gy0, *gc, *gw;
  float* flip_ws;
};

// round-2 fixed-grid kernels (slode_fixed.cuh): w1t may be strided (W1[:,0] of dynamics_hidden.weight); with
// plan_only nothing is launched and *ws_need receives the scratch bytes a launch with these arguments needs
typedef int (*fixed_fwd_fn)(const FwdArgs&, const PackSrc&, int w1t_stride, bool plan_only, size_t* ws_need);
typedef int (*fixed_bwd_fn)(const BwdArgs&, const PackSrc&, int w1t_stride, bool plan_only, size_t* ws_need);
typedef int (*dopri5_fwd_fn)(const Dopri5Args&, const PackSrc&, float* staging, cudaStream_t stream, int sms);
typedef int (*dopri5_bwd_fn)(const Dopri5BwdArgs&, const PackSrc&, float* staging, cudaStream_t stream, int sms);

// (25,5): CVS / challenge configs; (25,8): proc config; the rest serve tests and the width sweep.
// SLODE_SHAPES: shapes of the dopri5 translation units (slode_mlp_<H>_<S>.cu)
#define SLODE_SHAPES(X) X(25, 5) X(25, 8) X(16, 4) X(32, 5) X(64, 5)
// SLODE_FIXED_SHAPES: shapes of the fixed-grid kernels (slode_fixed.cuh, one unit slode_fixed_<H>_<S>.cu each)
#ifndef SLODE_FIXED_SHAPES
#define SLODE_FIXED_SHAPES(X) X(25, 5) X(25, 8) X(16, 4) X(32, 5) X(64, 5) X(128, 5) X(256, 5) X(512, 5)
#endif

#define SLODE_DECLARE_SHAPE(H, S)                                         \
  int dopri5_fwd_##H##_##S(const Dopri5Args&, const PackSrc&, float* staging, cudaStream_t stream, int sms); \
  int dopri5_bwd_##H##_##S(const Dopri5BwdArgs&, const PackSrc&, float* staging, cudaStream_t stream, int sms);
SLODE_SHAPES(SLODE_DECLARE_SHAPE)
#undef SLODE_DECLARE_SHAPE
#define SLODE_DECLARE_FIXED_SHAPE(H, S)                                                                       \
  int fixed_fwd_##H##_##S(const FwdArgs&, const PackSrc&, int w1t_stride, bool plan_only, size_t* ws_need);   \
  int fixed_bwd_##H##_##S(const BwdArgs&, const PackSrc&, int w1t_stride, bool plan_only, size_t* ws_need);
SLODE_FIXED_SHAPES(SLODE_DECLARE_FIXED_SHAPE)
#undef SLODE_DECLARE_FIXED_SHAPE

struct ShapeEntry {
  int H, S;
  dopri5_fwd_fn dopri5_fwd;
  dopri5_bwd_fn dopri5_bwd;
};
struct FixedShapeEntry {
  int H, S;
  fixed_fwd_fn fwd;
  fixed_bwd_fn bwd;
};
const ShapeEntry* find_shape(int H, int S);             // dopri5 units
const FixedShapeEntry* find_fixed_shape(int H, int S);  // fixed-grid kernels
int device_sms(int* sms);                               // SM count of the current device (cached)

}  // namespace slode
