// C-ABI entry points of the fixed-grid solve (euler / midpoint / rk4-3/8) and its reverse sweep: argument checks
// and dispatch to the compiled (H,S) shapes (kernels: slode_fixed.cuh, one translation unit per shape).
//
// These entry points keep NO state between calls: the kernels stage the weights from the caller's tensors, scratch
// memory is the caller's (slode_fixed_workspace_bytes), nothing is locked, recorded or allocated here.  They are
// re-entrant, can run concurrently on different streams and can be captured into a CUDA graph.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

#include "slode_common.cuh"
#include "slode_mlp_api.h"

namespace slode {

int device_sms(int* sms) {
  static int cached[64];
  static std::once_flag once[64];
  int dev = 0;
  SLODE_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    set_error("device index %d out of range", dev);
    return SLODE_EINVAL;
  }
  std::call_once(once[dev], [dev] {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
    cached[dev] = n;
  });
  if (cached[dev] <= 0) return cuda_fail(cudaGetLastError(), "cudaDeviceGetAttribute(multiProcessorCount)");
  *sms = cached[dev];
  return SLODE_OK;
}

namespace fx {
int cached_blocks_per_sm(const void* kern, size_t smem, int* blocks_per_sm) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, size_t, int>, int> cache;   // (kernel, dynamic smem, device) -> blocks per SM
  static std::map<std::pair<const void*, int>, bool> opted_in;        // kernels whose smem limit has been raised
  int dev = 0;
  SLODE_CUDA_TRY(cudaGetDevice(&dev));
  const auto key = std::make_tuple(kern, smem, dev);
  std::lock_guard<std::mutex> lock(mu);
  const auto it = cache.find(key);
  if (it != cache.end()) {
    *blocks_per_sm = it->second;
    return SLODE_OK;
  }
  if (!opted_in[{kern, dev}]) {
    // once per kernel and device: allow the largest dynamic shared memory the device offers, so that launches with
    // different sizes (different L, layouts, fused or not) never depend on the order in which they were first seen
    int optin = 0;
    SLODE_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SLODE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    opted_in[{kern, dev}] = true;
  }
  int n = 0;
  SLODE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 128, smem));
  if (n < 1) {
    set_error("fixed-grid kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    return SLODE_EUNSUPPORTED;
  }
  cache[key] = n;
  *blocks_per_sm = n;
  return SLODE_OK;
}
}  // namespace fx

static int check_sizes(const char* who, int method, int mode, int64_t B, int T, int H, int S) {
  if (B < 0 || T < 1 || H < 1 || S < 1) {
    set_error("%s: bad sizes B=%lld T=%d H=%d S=%d", who, (long long)B, T, H, S);
    return SLODE_EINVAL;
  }
  if (method != SLODE_METHOD_EULER && method != SLODE_METHOD_MIDPOINT && method != SLODE_METHOD_RK4) {
    set_error("%s: unknown method %d", who, method);
    return SLODE_EINVAL;
  }
  if (mode != SLODE_BWD_DISCRETE && mode != SLODE_BWD_TDE_ADJOINT) {
    set_error("%s: unknown mode %d", who, mode);
    return SLODE_EINVAL;
  }
  if (!find_fixed_shape(H, S)) {
    set_error("%s: (hidden=%d, state=%d) is not compiled in; there is no generic fallback", who, H, S);
    return SLODE_EUNSUPPORTED;
  }
  return SLODE_OK;
}

static int check_latent(const char* who, int L, const float* z, const float* W1, const float* b1, const float* Wa,
                        const float* ba, const float* Wb, const float* bb, const float* y0or) {
  if (L < 1 || !z || !W1 || !b1) {
    set_error("%s: latent inputs missing (L=%d)", who, L);
    return SLODE_EINVAL;
  }
  const int n = (Wa != nullptr) + (ba != nullptr) + (Wb != nullptr) + (bb != nullptr);
  if (n != 0 && n != 4) {
    set_error("%s: latent_to_ode_net weights must be given all together or not at all", who);
    return SLODE_EINVAL;
  }
  if (n == 0 && !y0or) {
    set_error("%s: neither latent_to_ode_net weights nor y0 / grad_y0 given", who);
    return SLODE_EINVAL;
  }
  return SLODE_OK;
}


}  // namespace slode

using namespace slode;

extern "C" int64_t slode_fixed_workspace_bytes(int backward, int method, int mode, int64_t B, int T, int L, int H,
                                               int S, int fused, int rows_in_time) {
  if (check_sizes("slode_fixed_workspace_bytes", method, mode, B, T, H, S)) return -1;
  if (fused < 0 || fused > 2 || (fused && L < 1)) return -1;
  if (B == 0) return 0;
  int sms = 0;
  if (device_sms(&sms)) return -1;
  // only the sizes and the null-ness of the latent pointers enter the launch plan
  static const float dummy = 0.0f;
  LatentSrc lat{};
  if (fused) {
    lat.z = &dummy;
    lat.L = L;
    if (fused == 2) lat.Wa = &dummy;
  }
  const PackSrc w{};
  size_t need = 0;
  int rc;
  if (backward) {
    BwdArgs a{};
    a.method = method; a.mode = mode; a.B = B; a.T = T; a.sms = sms; a.lat = lat;
    rc = find_fixed_shape(H, S)->bwd(a, w, 1, true, &need);
  } else {
    FwdArgs a{};
    a.method = method; a.B = B; a.T = T; a.sms = sms; a.lat = lat;
    a.st = rows_in_time ? S : B * (int64_t)S;
    rc = find_fixed_shape(H, S)->fwd(a, w, 1, true, &need);
  }
  return rc ? -1 : (int64_t)need;
}

static int run_fwd(const char* who, int H, int S, FwdArgs& a, const PackSrc& w, int w1t_stride) {
  g_fwd_launches = 0;
  if (a.B == 0) return SLODE_OK;
  int rc;
  rc = device_sms(&a.sms);
  if (rc) return rc;
  size_t need = 0;
  rc = find_fixed_shape(H, S)->fwd(a, w, w1t_stride, false, &need);
  if (rc == SLODE_OK) g_fwd_launches = 1;
  return rc;
}

static int run_bwd(const char* who, int H, int S, BwdArgs& a, const PackSrc& w, int w1t_stride) {
  g_bwd_launches = 0;
  if (a.B == 0) return SLODE_OK;
  int rc;
  rc = device_sms(&a.sms);
  if (rc) return rc;
  size_t need = 0;
  rc = find_fixed_shape(H, S)->bwd(a, w, w1t_stride, false, &need);
  if (rc == SLODE_OK) g_bwd_launches = 1;
  return rc;
}

extern "C" int slode_mlp_fixed_fwd(int method, int64_t B, int T, int H, int S, const float* t, const float* c,
                                   const float* y0, const float* w1t, const float* Wg, const float* bg,
                                   const float* Wd, const float* bd, float* sol, int64_t sol_stride_t,
                                   int64_t sol_stride_b, void* workspace, int64_t workspace_bytes, void* stream_) {
  int rc = check_sizes("slode_mlp_fixed_fwd", method, SLODE_BWD_DISCRETE, B, T, H, S);
  if (rc) return rc;
  if (!t || !w1t || !Wg || !bg || !Wd || !bd || (B > 0 && (!c || !y0 || !sol)) || workspace_bytes < 0) {
    set_error("slode_mlp_fixed_fwd: null pointer");
    return SLODE_EINVAL;
  }
  FwdArgs a{};
  a.method = method; a.B = B; a.T = T; a.t = t; a.c = c; a.y0 = y0; a.sol = sol; a.st = sol_stride_t;
  a.sb = sol_stride_b; a.stream = (cudaStream_t)stream_; a.ws = workspace; a.ws_bytes = (size_t)workspace_bytes;
  const PackSrc w{w1t, Wg, bg, Wd, bd};
  return run_fwd("slode_mlp_fixed_fwd", H, S, a, w, 1);
}

extern "C" int slode_mlp_fixed_bwd(int method, int mode, int64_t B, int T, int H, int S, const float* t,
                                   const float* c, const float* w1t, const float* Wg, const float* bg,
                                   const float* Wd, const float* bd, const float* sol, int64_t sol_stride_t,
                                   int64_t sol_stride_b, const float* grad_sol, int64_t gsol_stride_t,
                                   int64_t gsol_stride_b, float* grad_y0, float* grad_c, float* grad_w,
                                   void* workspace, int64_t workspace_bytes, void* stream_) {
  int rc = check_sizes("slode_mlp_fixed_bwd", method, mode, B, T, H, S);
  if (rc) return rc;
  if (!t || !w1t || !Wg || !bg || !Wd || !bd || !grad_w || workspace_bytes < 0 ||
      (B > 0 && (!c || !sol || !grad_sol || !grad_y0 || !grad_c))) {
    set_error("slode_mlp_fixed_bwd: null pointer");
    return SLODE_EINVAL;
  }
  BwdArgs a{};
  a.method = method; a.mode = mode; a.B = B; a.T = T; a.t = t; a.c = c; a.w1t = w1t; a.Wg = Wg; a.Wd = Wd;
  a.sol = sol; a.st = sol_stride_t; a.sb = sol_stride_b; a.gsol = grad_sol; a.gst = gsol_stride_t;
  a.gsb = gsol_stride_b; a.gy0 = grad_y0; a.gc = grad_c; a.gw = grad_w; a.stream = (cudaStream_t)stream_;
  a.ws = workspace; a.ws_bytes = (size_t)workspace_bytes;
  const PackSrc w{w1t, Wg, bg, Wd, bd};
  return run_bwd("slode_mlp_fixed_bwd", H, S, a, w, 1);
}

extern "C" int slode_latent_fixed_fwd(int method, int64_t B, int T, int L, int H, int S, const float* t, const float* z,
                                      const float* W1, const float* b1, const float* Wg, const float* bg,
                                      const float* Wd, const float* bd, const float* Wa, const float* ba,
                                      const float* Wb, const float* bb, const float* y0, float* sol,
                                      int64_t sol_stride_t, int64_t sol_stride_b, void* workspace,
                                      int64_t workspace_bytes, void* stream_) {
  int rc = check_sizes("slode_latent_fixed_fwd", method, SLODE_BWD_DISCRETE, B, T, H, S);
  if (rc) return rc;
  if (B > 0) {
    rc = check_latent("slode_latent_fixed_fwd", L, z, W1, b1, Wa, ba, Wb, bb, y0);
    if (rc) return rc;
  }
  if (!t || !Wg || !bg || !Wd || !bd || (B > 0 && !sol) || workspace_bytes < 0) {
    set_error("slode_latent_fixed_fwd: null pointer");
    return SLODE_EINVAL;
  }
  FwdArgs a{};
  a.method = method; a.B = B; a.T = T; a.t = t; a.y0 = y0; a.sol = sol; a.st = sol_stride_t; a.sb = sol_stride_b;
  a.stream = (cudaStream_t)stream_; a.lat = LatentSrc{z, L, W1, b1, Wa, ba, Wb, bb};
  a.ws = workspace; a.ws_bytes = (size_t)workspace_bytes;
  const PackSrc w{W1, Wg, bg, Wd, bd};  // w1t = W1[:,0], stride L + 1
  return run_fwd("slode_latent_fixed_fwd", H, S, a, w, L + 1);
}

extern "C" int slode_latent_fixed_heads_fwd(int method, int64_t B, int T, int L, int H, int S, const float* t,
                                            const float* z, const float* W1, const float* b1, const float* Wg,
                                            const float* bg, const float* Wd, const float* bd, const float* Wa,
                                            const float* ba, const float* Wb, const float* bb, const float* y0,
                                            int O, int NQ, const float* head_W, float* mu, int64_t mu_row_pitch,
                                            float* sol,
                                            int64_t sol_stride_t, int64_t sol_stride_b, void* workspace,
                                            int64_t workspace_bytes, void* stream_) {
  int rc = check_sizes("slode_latent_fixed_heads_fwd", method, SLODE_BWD_DISCRETE, B, T, H, S);
  if (rc) return rc;
  if (B > 0) {
    rc = check_latent("slode_latent_fixed_heads_fwd", L, z, W1, b1, Wa, ba, Wb, bb, y0);
    if (rc) return rc;
  }
  if (O < 1 || O > 8 || NQ < 1 || NQ > 3 || NQ * O * S > kMaxHeadW) {
    set_error("slode_latent_fixed_heads_fwd: obs_dim=%d (1..8), heads=%d (1..3) out of range", O, NQ);
    return SLODE_EINVAL;
  }
  if (!t || !Wg || !bg || !Wd || !bd || !head_W || (B > 0 && !mu) || workspace_bytes < 0) {
    set_error("slode_latent_fixed_heads_fwd: null pointer");
    return SLODE_EINVAL;
  }
  if (mu_row_pitch < T || (reinterpret_cast<uintptr_t>(mu) & 3)) {
    set_error("slode_latent_fixed_heads_fwd: mu_row_pitch=%lld must be >= T=%d and mu 4-byte aligned",
              (long long)mu_row_pitch, T);
    return SLODE_EINVAL;
  }
  FwdArgs a{};
  a.method = method; a.B = B; a.T = T; a.t = t; a.y0 = y0; a.sol = sol;
  // without sol the strides only steer the staging layout: use the (T,B,S) ones
  a.st = sol ? sol_stride_t : B * (int64_t)S; a.sb = sol ? sol_stride_b : S;
  a.stream = (cudaStream_t)stream_; a.lat = LatentSrc{z, L, W1, b1, Wa, ba, Wb, bb};
  a.ws = workspace; a.ws_bytes = (size_t)workspace_bytes;
  a.heads = HeadsSrc{head_W, mu, NQ, O, mu_row_pitch};
  const PackSrc w{W1, Wg, bg, Wd, bd};
  return run_fwd("slode_latent_fixed_heads_fwd", H, S, a, w, L + 1);
}

extern "C" int slode_latent_fixed_bwd(int method, int mode, int64_t B, int T, int L, int H, int S, const float* t,
                                      const float* z, const float* W1, const float* b1, const float* Wg,
                                      const float* bg, const float* Wd, const float* bd, const float* Wa,
                                      const float* ba, const float* Wb, const float* bb, const float* sol,
                                      int64_t sol_stride_t, int64_t sol_stride_b, const float* grad_sol,
                                      int64_t gsol_stride_t, int64_t gsol_stride_b, float* grad_z, float* grad_y0,
                                      float* grad_params, void* workspace, int64_t workspace_bytes, void* stream_) {
  int rc = check_sizes("slode_latent_fixed_bwd", method, mode, B, T, H, S);
  if (rc) return rc;
  if (B > 0) {
    rc = check_latent("slode_latent_fixed_bwd", L, z, W1, b1, Wa, ba, Wb, bb, grad_y0);
    if (rc) return rc;
  }
  if (!t || !Wg || !bg || !Wd || !bd || !grad_params || workspace_bytes < 0 ||
      (B > 0 && (!sol || !grad_sol || !grad_z))) {
    set_error("slode_latent_fixed_bwd: null pointer");
    return SLODE_EINVAL;
  }
  BwdArgs a{};
  a.method = method; a.mode = mode; a.B = B; a.T = T; a.t = t; a.Wg = Wg; a.Wd = Wd;
  a.sol = sol; a.st = sol_stride_t; a.sb = sol_stride_b; a.gsol = grad_sol; a.gst = gsol_stride_t;
  a.gsb = gsol_stride_b; a.gy0 = grad_y0; a.gw = grad_params; a.stream = (cudaStream_t)stream_;
  a.lat = LatentSrc{z, L, W1, b1, Wa, ba, Wb, bb}; a.gz = grad_z;
  a.ws = workspace; a.ws_bytes = (size_t)workspace_bytes;
  const PackSrc w{W1, Wg, bg, Wd, bd};
  return run_bwd("slode_latent_fixed_bwd", H, S, a, w, L + 1);
}
