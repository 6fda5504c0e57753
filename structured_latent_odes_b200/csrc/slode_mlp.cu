// C-ABI entry points: library queries and the dopri5 solve (argument checks and dispatch to the compiled (H,S)
// shapes; kernels in slode_dopri5_kernels.cuh, one translation unit per shape slode_mlp_<H>_<S>.cu).  The
// fixed-grid entry points live in slode_fixed_api.cu.
#include <algorithm>

#include "slode_common.cuh"
#include "slode_mlp_api.h"

namespace slode {

static const ShapeEntry kShapes[] = {
#define X(h, s) {h, s, dopri5_fwd_##h##_##s, dopri5_bwd_##h##_##s},
    SLODE_SHAPES(X)
#undef X
};
constexpr int kNumShapes = sizeof(kShapes) / sizeof(kShapes[0]);

const ShapeEntry* find_shape(int H, int S) {
  for (int i = 0; i < kNumShapes; ++i)
    if (kShapes[i].H == H && kShapes[i].S == S) return &kShapes[i];
  return nullptr;
}

static const FixedShapeEntry kFixedShapes[] = {
#define X(h, s) {h, s, fixed_fwd_##h##_##s, fixed_bwd_##h##_##s},
    SLODE_FIXED_SHAPES(X)
#undef X
};
constexpr int kNumFixedShapes = sizeof(kFixedShapes) / sizeof(kFixedShapes[0]);

const FixedShapeEntry* find_fixed_shape(int H, int S) {
  for (int i = 0; i < kNumFixedShapes; ++i)
    if (kFixedShapes[i].H == H && kFixedShapes[i].S == S) return &kFixedShapes[i];
  return nullptr;
}

// mirrors dopri5_scratch_bytes of slode_dopri5_kernels.cuh (that header is only included by the shape units)
static size_t dopri5_scratch_bytes_host(int64_t B, int S, int sms) {
  const size_t nstate = (size_t)B * S;
  return (sizeof(float) * 5 * nstate + 255) / 256 * 256 + (sizeof(double) * 6 * (size_t)sms * 4 + 255) / 256 * 256 + 256 +
         (sizeof(Dopri5Ctrl) + 255) / 256 * 256;
}

static int check_common(const char* who, int64_t B, int T, int H, int S) {
  if (B < 0 || T < 1 || H < 1 || S < 1) {
    set_error("%s: bad sizes B=%lld T=%d H=%d S=%d", who, (long long)B, T, H, S);
    return SLODE_EINVAL;
  }
  if (!find_shape(H, S)) {
    set_error("%s: (hidden=%d, state=%d) is not compiled in; there is no generic fallback", who, H, S);
    return SLODE_EUNSUPPORTED;
  }
  return SLODE_OK;
}

}  // namespace slode

using namespace slode;

extern "C" int slode_mlp_supported(int H, int S) { return find_fixed_shape(H, S) ? 1 : 0; }
extern "C" int slode_dopri5_supported(int H, int S) { return find_shape(H, S) ? 1 : 0; }

extern "C" int slode_query(int what) {
  switch (what) {
    case SLODE_Q_VERSION: return 4;
    case SLODE_Q_SM_ARCH: return 100;
    case SLODE_Q_MAX_HIDDEN: {
      int m = 0;
      for (int i = 0; i < kNumFixedShapes; ++i) m = std::max(m, kFixedShapes[i].H);
      return m;
    }
    case SLODE_Q_MAX_STATE: {
      int m = 0;
      for (int i = 0; i < kNumFixedShapes; ++i) m = std::max(m, kFixedShapes[i].S);
      return m;
    }
    case SLODE_Q_N_SHAPES: return kNumFixedShapes;
    case SLODE_Q_FWD_LAUNCHES: return g_fwd_launches;
    case SLODE_Q_BWD_LAUNCHES: return g_bwd_launches;
    case SLODE_Q_TOTAL_LAUNCHES: return (int)(g_total_launches.load() & 0x7fffffff);
#ifdef SLODE_SOURCE_HASH
    case SLODE_Q_SOURCE_HASH: return (int)(SLODE_SOURCE_HASH);
#else
    case SLODE_Q_SOURCE_HASH: return 0;
#endif
  }
  if (what >= SLODE_Q_SHAPE_BASE && what < SLODE_Q_SHAPE_BASE + 2 * kNumFixedShapes) {
    const int i = (what - SLODE_Q_SHAPE_BASE) / 2;
    return ((what - SLODE_Q_SHAPE_BASE) & 1) ? kFixedShapes[i].S : kFixedShapes[i].H;
  }
  return -1;
}

static int dopri5_fwd_common(const char* who, int64_t B, int T, int H, int S, const float* t, const float* c,
                             const float* y0, const float* w1t, const float* Wg, const float* bg, const float* Wd,
                             const float* bd, double rtol, double atol, double first_step, int64_t max_attempts,
                             const double* replay_steps, int64_t n_replay, float* sol, int64_t sol_stride_t,
                             int64_t sol_stride_b, float* ckpt_y, int64_t ckpt_capacity, double* step_log,
                             int64_t log_capacity, int64_t* stats, Dopri5Args step, cudaStream_t stream) {
  int rc = check_common(who, B, T, H, S);
  if (rc) return rc;
  if (!(rtol >= 0.0) || !(atol >= 0.0) || (rtol == 0.0 && atol == 0.0) || max_attempts < 1 || ckpt_capacity < 0 ||
      log_capacity < 0) {
    set_error("%s: bad tolerances / capacities (rtol=%g atol=%g max_attempts=%lld)", who, rtol, atol,
              (long long)max_attempts);
    return SLODE_EINVAL;
  }
  if (!t || !w1t || !Wg || !bg || !Wd || !bd || !stats || (B > 0 && (!c || !y0 || !sol))) {
    set_error("%s: null pointer", who);
    return SLODE_EINVAL;
  }
  g_fwd_launches = 0;
  PackGuard guard(stream);
  if (guard.status) return guard.status;
  Dopri5Args a = step;
  a.B = B; a.T = T; a.t = t; a.c = c; a.y0 = y0; a.sol = sol; a.st = sol_stride_t; a.sb = sol_stride_b;
  a.rtol = (float)rtol; a.atol = (float)atol; a.first_step = first_step; a.max_attempts = max_attempts;
  a.replay = replay_steps; a.n_replay = replay_steps ? n_replay : 0;
  a.ckpt_y = ckpt_y; a.ckpt_cap = ckpt_capacity; a.step_log = step_log; a.log_cap = log_capacity; a.stats = stats;
  const PackSrc w{w1t, Wg, bg, Wd, bd};
  rc = find_shape(H, S)->dopri5_fwd(a, w, guard.staging, stream, guard.sms);
  if (rc == SLODE_OK) g_fwd_launches = 2;
  return rc;
}

extern "C" int slode_mlp_dopri5_fwd(int64_t B, int T, int H, int S, const float* t, const float* c, const float* y0,
                                    const float* w1t, const float* Wg, const float* bg, const float* Wd,
                                    const float* bd, double rtol, double atol, double first_step,
                                    int64_t max_attempts, const double* replay_steps, int64_t n_replay, float* sol, int64_t sol_stride_t, int64_t sol_stride_b,
                                    float* ckpt_y, int64_t ckpt_capacity, double* step_log, int64_t log_capacity,
                                    int64_t* stats, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (B == 0) {
    if (!stats) {
      set_error("slode_mlp_dopri5_fwd: null pointer");
      return SLODE_EINVAL;
    }
    if (cudaMemsetAsync(stats, 0, 4 * sizeof(int64_t), stream) != cudaSuccess) return cuda_fail(cudaGetLastError(), "memset");
    return SLODE_OK;
  }
  return dopri5_fwd_common("slode_mlp_dopri5_fwd", B, T, H, S, t, c, y0, w1t, Wg, bg, Wd, bd, rtol, atol, first_step,
                           max_attempts, replay_steps, n_replay, sol, sol_stride_t, sol_stride_b, ckpt_y, ckpt_capacity,
                           step_log, log_capacity, stats, Dopri5Args{}, stream);
}

extern "C" int64_t slode_mlp_dopri5_step_workspace_bytes(int64_t B, int S) {
  if (B < 1 || S < 1) {
    set_error("slode_mlp_dopri5_step_workspace_bytes: bad sizes");
    return -1;
  }
  int sms = 0;
  if (device_sms(&sms)) return -1;
  return (int64_t)dopri5_scratch_bytes_host(B, S, sms);
}

extern "C" int slode_mlp_dopri5_fwd_step(int64_t B, int T, int H, int S, const float* t, const float* c,
                                         const float* y0, const float* w1t, const float* Wg, const float* bg,
                                         const float* Wd, const float* bd, double rtol, double atol,
                                         double first_step, int64_t max_attempts, int64_t n_global, int restart,
                                         const double* ext_sums, double* out_sums, float* sol, int64_t sol_stride_t,
                                         int64_t sol_stride_b, float* ckpt_y, int64_t ckpt_capacity, double* step_log,
                                         int64_t log_capacity, int64_t* stats, void* workspace, int64_t workspace_bytes,
                                         void* stream_) {
  if (B < 1 || n_global < B || !ext_sums || !out_sums || !workspace) {
    set_error("slode_mlp_dopri5_fwd_step: needs a non-empty shard (B=%lld of n_global=%lld), ext_sums, out_sums and a "
              "workspace", (long long)B, (long long)n_global);
    return SLODE_EINVAL;
  }
  Dopri5Args step{};
  step.restart = restart ? 1 : 0;
  step.n_global = n_global;
  step.ext_sums = ext_sums;
  step.out_sums = out_sums;
  step.step_ws = workspace;
  step.step_ws_bytes = (size_t)workspace_bytes;
  return dopri5_fwd_common("slode_mlp_dopri5_fwd_step", B, T, H, S, t, c, y0, w1t, Wg, bg, Wd, bd, rtol, atol,
                           first_step, max_attempts, nullptr, 0, sol, sol_stride_t, sol_stride_b, ckpt_y, ckpt_capacity,
                           step_log, log_capacity, stats, step, (cudaStream_t)stream_);
}

extern "C" int slode_mlp_dopri5_bwd(int64_t B, int T, int H, int S, const float* t, const float* c, const float* w1t,
                                    const float* Wg, const float* bg, const float* Wd, const float* bd,
                                    int64_t n_accepted, const double* accepted_steps, const int* emit_ranges,
                                    const float* ckpt_y, const float* grad_sol, int64_t gsol_stride_t,
                                    int64_t gsol_stride_b, float* grad_y0, float* grad_c, float* grad_w,
                                    void* stream_) {
  int rc = check_common("slode_mlp_dopri5_bwd", B, T, H, S);
  if (rc) return rc;
  if (n_accepted < 0 || (T > 1 && n_accepted < 1)) {
    set_error("slode_mlp_dopri5_bwd: n_accepted=%lld for T=%d", (long long)n_accepted, T);
    return SLODE_EINVAL;
  }
  if (!t || !w1t || !Wg || !bg || !Wd || !bd || !grad_w || (n_accepted > 0 && (!accepted_steps || !emit_ranges)) ||
      (B > 0 && (!c || !grad_sol || !grad_y0 || !grad_c || (n_accepted > 0 && !ckpt_y)))) {
    set_error("slode_mlp_dopri5_bwd: null pointer");
    return SLODE_EINVAL;
  }
  g_bwd_launches = 0;
  if (B == 0) return SLODE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  PackGuard guard(stream);
  if (guard.status) return guard.status;
  Dopri5BwdArgs a{};
  a.B = B; a.T = T; a.t = t; a.c = c; a.w1t = w1t; a.Wg = Wg; a.Wd = Wd; a.n_acc = n_accepted;
  a.acc_steps = accepted_steps; a.emit = emit_ranges; a.ckpt_y = ckpt_y; a.gsol = grad_sol; a.gst = gsol_stride_t;
  a.gsb = gsol_stride_b; a.gy0 = grad_y0; a.gc = grad_c; a.gw = grad_w;
  const PackSrc w{w1t, Wg, bg, Wd, bd};
  rc = find_shape(H, S)->dopri5_bwd(a, w, guard.staging, stream, guard.sms);
  if (rc == SLODE_OK) g_bwd_launches = 2;
  return rc;
}

