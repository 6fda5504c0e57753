// Host-side interface between the C-ABI dispatcher (slode_mlp.cu) and the per-shape translation units
// (slode_mlp_<H>_<S>.cu), each of which owns its own packed-weight buffer in constant memory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace slode {

struct PackSrc {  // device pointers, torch Linear layouts
  const float *w1t, *Wg, *bg, *Wd, *bd;
};

struct FwdArgs {
  int method;
  int64_t B;
  int T;
  const float *t, *c, *y0;
  float* sol;
  int64_t st, sb;
  cudaStream_t stream;
  int sms;
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
};

typedef int (*mlp_fwd_fn)(const FwdArgs&, const PackSrc&, float* staging);
typedef int (*mlp_bwd_fn)(const BwdArgs&, const PackSrc&, float* staging);

// (25,5): CVS / challenge configs; (25,8): proc config; the rest serve tests and the width sweep.
#define SLODE_SHAPES(X) X(25, 5) X(25, 8) X(16, 4) X(32, 5)

#define SLODE_DECLARE_SHAPE(H, S)                                         \
  int mlp_fwd_##H##_##S(const FwdArgs&, const PackSrc&, float* staging);  \
  int mlp_bwd_##H##_##S(const BwdArgs&, const PackSrc&, float* staging);
SLODE_SHAPES(SLODE_DECLARE_SHAPE)
#undef SLODE_DECLARE_SHAPE

}  // namespace slode
