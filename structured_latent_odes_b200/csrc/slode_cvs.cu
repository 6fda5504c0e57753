// Fixed-grid solve of the CVS mechanistic ODE and its reverse sweep (sm_100a).
//
// Right-hand side (reference: dx_dt, data/cvs/cvs_data.py:52-91), states x = (Pa/100, Pv/10, S, SV/100):
//     f_hr  = S (f_max - f_min) + f_min
//     r_tpr = S (r_max - r_min) + r_min - r_tpr_mod
//     dva   = -(Pa - Pv) / r_tpr + SV f_hr
//     dPa   = dva / (100 ca)            dPv = (-dva + i_ext) / (10 cv)
//     dS    = (1 - 1/(1 + exp(-k (Pa - p_set))) - S) / tau
//     dSV   = i_ext sv_mod
// with per-trajectory treatments i_ext, r_tpr_mod (:24-26) and ten shared constants theta (:28-49) in the order
//     theta = [f_hr_max, f_hr_min, r_tpr_max, r_tpr_min, sv_mod, ca, cv, k_width, p_aset, tau].
// The RHS is autonomous, ~30 flop + one exp + two divisions per evaluation against 16 bytes stored per output
// time: the forward is bound by the HBM store of sol, the backward by the reads of sol and grad_sol.
//
// One thread = one trajectory; an output interval may be split into `substeps` equal solver steps (the
// generator's accuracy knob; substeps == 1 is torchdiffeq's "grid == t").  Templated on the real type: float for
// the odeint drop-in, double for the data generator (the reference generates in float64).
#include <algorithm>

#include "slode_common.cuh"

namespace slode {
namespace cvs {

constexpr int kBlock = 256;
constexpr int kMaxSub = 16;
constexpr int NTH = 10;

template <class R> struct Theta {
  R fmax, fmin, rmax, rmin, svm, ca, cv, k, pset, tau;
  R dF, dR, ica, icv, itau, rca, rcv;  // derived once per thread: f_max - f_min, r_max - r_min, 1/(100 ca),
                                       // 1/(10 cv), 1/tau, 1/ca, 1/cv
};

template <class R> __device__ __forceinline__ R rexp(R x);
template <> __device__ __forceinline__ float rexp<float>(float x) { return __expf(x); }
template <> __device__ __forceinline__ double rexp<double>(double x) { return exp(x); }
// 1/x: float32 = MUFU.RCP + one Newton step (about 1 ulp; the IEEE division is a ~10-instruction sequence with a
// subroutine call for the slow path, twice per RHS evaluation, and these kernels are bound by exactly that latency
// chain); float64 (the data generator) keeps the exact division
template <class R> __device__ __forceinline__ R rrcp(R x);
template <> __device__ __forceinline__ float rrcp<float>(float x) {
  const float r = rcp_approx(x);
  return fmaf(fmaf(-x, r, 1.0f), r, r);
}
template <> __device__ __forceinline__ double rrcp<double>(double x) { return 1.0 / x; }

template <class R> struct V4 { R v[4]; };

template <class R> __device__ __forceinline__ V4<R> axpy(R a, const V4<R>& x, const V4<R>& y) {
  V4<R> r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = fma(a, x.v[i], y.v[i]);
  return r;
}

template <class R>
__device__ __forceinline__ V4<R> rhs(const V4<R>& x, R ie, R rm, const Theta<R>& th) {
  const R pa = R(100) * x.v[0], pv = R(10) * x.v[1], s = x.v[2], sv = R(100) * x.v[3];
  const R fhr = fma(s, th.dF, th.fmin);
  const R r = fma(s, th.dR, th.rmin) - rm;
  const R dva = sv * fhr - (pa - pv) * rrcp<R>(r);
  const R e = rexp<R>(-th.k * (pa - th.pset));
  V4<R> f;
  f.v[0] = dva * th.ica;
  f.v[1] = (ie - dva) * th.icv;
  f.v[2] = (R(1) - rrcp<R>(R(1) + e) - s) * th.itau;
  f.v[3] = ie * th.svm;
  return f;
}

// vector-Jacobian product g^T df/d(x, ie, rm, theta) at x; returns g^T df/dx, accumulates the rest scaled by w
template <class R, bool PARAMS>
__device__ __forceinline__ V4<R> vjp(const V4<R>& x, R ie, R rm, const Theta<R>& th, const V4<R>& g, R w, R& gie,
                                     R& grm, R (&gth)[NTH]) {
  const R pa = R(100) * x.v[0], pv = R(10) * x.v[1], s = x.v[2], sv = R(100) * x.v[3];
  const R dF = th.dF, dR = th.dR;
  const R fhr = fma(s, dF, th.fmin);
  const R r = fma(s, dR, th.rmin) - rm;
  const R rinv = rrcp<R>(r);
  const R q = (pa - pv) * rinv;
  const R dva = sv * fhr - q;
  const R e = rexp<R>(-th.k * (pa - th.pset));
  const R sg = rrcp<R>(R(1) + e);
  const R ica = th.ica, icv = th.icv, itau = th.itau;
  const R g_dva = g.v[0] * ica - g.v[1] * icv;
  const R g_r = g_dva * q * rinv;          // d(-q)/dr = q/r
  const R g_fhr = g_dva * sv;
  const R g_sg = -g.v[2] * itau;
  const R sgp = sg * (R(1) - sg);          // = sg^2 e
  const R g_u = -g_sg * sgp;               // u = -k (pa - pset), dsg/du = -sg(1-sg) ... see below
  // sg = 1/(1+e^u): dsg/du = -e^u/(1+e^u)^2 = -sg(1-sg)  => g_u = g_sg * (-sgp)
  V4<R> gx;
  gx.v[0] = R(100) * (-g_dva * rinv + g_u * (-th.k));
  gx.v[1] = R(10) * (g_dva * rinv);
  gx.v[2] = g_fhr * dF + g_r * dR - g.v[2] * itau;
  gx.v[3] = R(100) * g_dva * fhr;
  if (PARAMS) {
    gie = fma(w, g.v[1] * icv + g.v[3] * th.svm, gie);
    grm = fma(w, -g_r, grm);
    gth[0] = fma(w, g_fhr * s, gth[0]);
    gth[1] = fma(w, g_fhr * (R(1) - s), gth[1]);
    gth[2] = fma(w, g_r * s, gth[2]);
    gth[3] = fma(w, g_r * (R(1) - s), gth[3]);
    gth[4] = fma(w, g.v[3] * ie, gth[4]);
    gth[5] = fma(w, -g.v[0] * dva * ica * th.rca, gth[5]);
    gth[6] = fma(w, -g.v[1] * (ie - dva) * icv * th.rcv, gth[6]);
    gth[7] = fma(w, g_u * (-(pa - th.pset)), gth[7]);
    gth[8] = fma(w, g_u * th.k, gth[8]);
    gth[9] = fma(w, -g.v[2] * (R(1) - sg - s) * itau * itau, gth[9]);
  }
  return gx;
}

// one solver step y -> y1; stage points Y[] are returned for the reverse sweep
template <class R, int METHOD>
__device__ __forceinline__ V4<R> step(const V4<R>& y, R dt, R ie, R rm, const Theta<R>& th, V4<R> (&Y)[4]) {
  Y[0] = y;
  const V4<R> k1 = rhs<R>(y, ie, rm, th);
  if (METHOD == SLODE_METHOD_EULER) return axpy<R>(dt, k1, y);
  if (METHOD == SLODE_METHOD_MIDPOINT) {
    Y[1] = axpy<R>(R(0.5) * dt, k1, y);
    return axpy<R>(dt, rhs<R>(Y[1], ie, rm, th), y);
  }
  const R third = R(1.0 / 3.0);
  Y[1] = axpy<R>(dt * third, k1, y);
  const V4<R> k2 = rhs<R>(Y[1], ie, rm, th);
  V4<R> tmp;
#pragma unroll
  for (int i = 0; i < 4; ++i) tmp.v[i] = k2.v[i] - k1.v[i] * third;
  Y[2] = axpy<R>(dt, tmp, y);
  const V4<R> k3 = rhs<R>(Y[2], ie, rm, th);
#pragma unroll
  for (int i = 0; i < 4; ++i) tmp.v[i] = k1.v[i] - k2.v[i] + k3.v[i];
  Y[3] = axpy<R>(dt, tmp, y);
  const V4<R> k4 = rhs<R>(Y[3], ie, rm, th);
#pragma unroll
  for (int i = 0; i < 4; ++i) tmp.v[i] = (k1.v[i] + R(3) * (k2.v[i] + k3.v[i]) + k4.v[i]);
  return axpy<R>(dt * R(0.125), tmp, y);
}

// exact adjoint of step(): lam = dL/dy1 -> dL/dy; parameter cotangents accumulated
template <class R, int METHOD>
__device__ __forceinline__ V4<R> step_adjoint(const V4<R> (&Y)[4], R dt, R ie, R rm, const Theta<R>& th,
                                              const V4<R>& lam, R& gie, R& grm, R (&gth)[NTH]) {
  V4<R> out = lam;
  if (METHOD == SLODE_METHOD_EULER) {
    V4<R> gk;
#pragma unroll
    for (int i = 0; i < 4; ++i) gk.v[i] = dt * lam.v[i];
    const V4<R> gy = vjp<R, true>(Y[0], ie, rm, th, gk, R(1), gie, grm, gth);
#pragma unroll
    for (int i = 0; i < 4; ++i) out.v[i] += gy.v[i];
    return out;
  }
  if (METHOD == SLODE_METHOD_MIDPOINT) {
    V4<R> gk;
#pragma unroll
    for (int i = 0; i < 4; ++i) gk.v[i] = dt * lam.v[i];
    const V4<R> gy2 = vjp<R, true>(Y[1], ie, rm, th, gk, R(1), gie, grm, gth);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      out.v[i] += gy2.v[i];
      gk.v[i] = R(0.5) * dt * gy2.v[i];
    }
    const V4<R> gy1 = vjp<R, true>(Y[0], ie, rm, th, gk, R(1), gie, grm, gth);
#pragma unroll
    for (int i = 0; i < 4; ++i) out.v[i] += gy1.v[i];
    return out;
  }
  const R third = R(1.0 / 3.0);
  V4<R> gk1, gk2, gk3, gk4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const R w = R(0.125) * dt * lam.v[i];
    gk1.v[i] = w; gk2.v[i] = R(3) * w; gk3.v[i] = R(3) * w; gk4.v[i] = w;
  }
  V4<R> gy = vjp<R, true>(Y[3], ie, rm, th, gk4, R(1), gie, grm, gth);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out.v[i] += gy.v[i];
    gk1.v[i] += dt * gy.v[i]; gk2.v[i] -= dt * gy.v[i]; gk3.v[i] += dt * gy.v[i];
  }
  gy = vjp<R, true>(Y[2], ie, rm, th, gk3, R(1), gie, grm, gth);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out.v[i] += gy.v[i];
    gk2.v[i] += dt * gy.v[i]; gk1.v[i] -= dt * third * gy.v[i];
  }
  gy = vjp<R, true>(Y[1], ie, rm, th, gk2, R(1), gie, grm, gth);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out.v[i] += gy.v[i];
    gk1.v[i] += dt * third * gy.v[i];
  }
  gy = vjp<R, true>(Y[0], ie, rm, th, gk1, R(1), gie, grm, gth);
#pragma unroll
  for (int i = 0; i < 4; ++i) out.v[i] += gy.v[i];
  return out;
}

// torchdiffeq.odeint_adjoint emulation: one solver step of the augmented system [y, a, a_params] in reversed
// time (ds > 0):  dy/ds = -f,  da/ds = +a^T df/dy,  da_p/ds = +a^T df/dp.
template <class R, int METHOD>
__device__ __forceinline__ void aug_step(V4<R>& y, V4<R>& a, R ds, R ie, R rm, const Theta<R>& th, R& gie, R& grm,
                                         R (&gth)[NTH]) {
  auto K = [&](const V4<R>& yy, const V4<R>& aa, R w, V4<R>& ky, V4<R>& ka) {
    const V4<R> f = rhs<R>(yy, ie, rm, th);
#pragma unroll
    for (int i = 0; i < 4; ++i) ky.v[i] = -f.v[i];
    ka = vjp<R, true>(yy, ie, rm, th, aa, w, gie, grm, gth);
  };
  V4<R> ky1, ka1;
  if (METHOD == SLODE_METHOD_EULER) {
    K(y, a, ds, ky1, ka1);
    y = axpy<R>(ds, ky1, y);
    a = axpy<R>(ds, ka1, a);
    return;
  }
  if (METHOD == SLODE_METHOD_MIDPOINT) {
    K(y, a, R(0), ky1, ka1);
    const V4<R> ym = axpy<R>(R(0.5) * ds, ky1, y), am = axpy<R>(R(0.5) * ds, ka1, a);
    V4<R> ky2, ka2;
    K(ym, am, ds, ky2, ka2);
    y = axpy<R>(ds, ky2, y);
    a = axpy<R>(ds, ka2, a);
    return;
  }
  const R third = R(1.0 / 3.0), w8 = R(0.125) * ds;
  V4<R> ky2, ka2, ky3, ka3, ky4, ka4, ty, ta;
  K(y, a, w8, ky1, ka1);
  K(axpy<R>(ds * third, ky1, y), axpy<R>(ds * third, ka1, a), R(3) * w8, ky2, ka2);
#pragma unroll
  for (int i = 0; i < 4; ++i) { ty.v[i] = ky2.v[i] - ky1.v[i] * third; ta.v[i] = ka2.v[i] - ka1.v[i] * third; }
  K(axpy<R>(ds, ty, y), axpy<R>(ds, ta, a), R(3) * w8, ky3, ka3);
#pragma unroll
  for (int i = 0; i < 4; ++i) { ty.v[i] = ky1.v[i] - ky2.v[i] + ky3.v[i]; ta.v[i] = ka1.v[i] - ka2.v[i] + ka3.v[i]; }
  K(axpy<R>(ds, ty, y), axpy<R>(ds, ta, a), w8, ky4, ka4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ty.v[i] = ky1.v[i] + R(3) * (ky2.v[i] + ky3.v[i]) + ky4.v[i];
    ta.v[i] = ka1.v[i] + R(3) * (ka2.v[i] + ka3.v[i]) + ka4.v[i];
  }
  y = axpy<R>(w8, ty, y);
  a = axpy<R>(w8, ta, a);
}

template <class R> __device__ __forceinline__ Theta<R> load_theta(const R* th) {
  Theta<R> t;
  t.fmax = th[0]; t.fmin = th[1]; t.rmax = th[2]; t.rmin = th[3]; t.svm = th[4];
  t.ca = th[5]; t.cv = th[6]; t.k = th[7]; t.pset = th[8]; t.tau = th[9];
  t.dF = t.fmax - t.fmin; t.dR = t.rmax - t.rmin;
  t.ica = R(1) / (t.ca * R(100)); t.icv = R(1) / (t.cv * R(10)); t.itau = R(1) / t.tau;
  t.rca = R(1) / t.ca; t.rcv = R(1) / t.cv;
  return t;
}

template <class R> __device__ __forceinline__ V4<R> load4(const R* p) {
  V4<R> r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = p[i];
  return r;
}
template <> __device__ __forceinline__ V4<float> load4<float>(const float* p) {
  V4<float> r;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    r.v[0] = v.x; r.v[1] = v.y; r.v[2] = v.z; r.v[3] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = p[i];
  }
  return r;
}
template <class R> __device__ __forceinline__ void store4(R* p, const V4<R>& a) {
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = a.v[i];
}
template <> __device__ __forceinline__ void store4<float>(float* p, const V4<float>& a) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = a.v[i];
  }
}

template <class R, int METHOD>
__global__ void __launch_bounds__(kBlock)
cvs_fwd_kernel(int64_t B, int T, int nsub, const R* __restrict__ tgrid, const R* __restrict__ y0,
               const R* __restrict__ iext, const R* __restrict__ rmod, const R* __restrict__ theta,
               R* __restrict__ sol, int64_t st, int64_t sb) {
  const Theta<R> th = load_theta<R>(theta);
  for (int64_t b = (int64_t)blockIdx.x * kBlock + threadIdx.x; b < B; b += (int64_t)gridDim.x * kBlock) {
    const R ie = iext[b], rm = rmod[b];
    V4<R> y = load4<R>(y0 + b * 4);
    R* out = sol + b * sb;
    store4<R>(out, y);
    R t0 = tgrid[0];
    V4<R> Y[4];
    for (int i = 0; i + 1 < T; ++i) {
      const R t1 = tgrid[i + 1];
      const R h = nsub == 1 ? (t1 - t0) : (t1 - t0) / R(nsub);   // the common case pays no division
      for (int k = 0; k < nsub; ++k) y = step<R, METHOD>(y, h, ie, rm, th, Y);
      out += st;
      store4<R>(out, y);
      t0 = t1;
    }
  }
}

template <class R, int METHOD, int MODE>
__global__ void __launch_bounds__(kBlock)
cvs_bwd_kernel(int64_t B, int T, int nsub, const R* __restrict__ tgrid, const R* __restrict__ iext,
               const R* __restrict__ rmod, const R* __restrict__ theta, const R* __restrict__ sol, int64_t st,
               int64_t sb, const R* __restrict__ gsol, int64_t gst, int64_t gsb, R* __restrict__ grad_y0,
               R* __restrict__ grad_iext, R* __restrict__ grad_rmod, R* __restrict__ grad_theta) {
  __shared__ R red[NTH];
  if (threadIdx.x < NTH) red[threadIdx.x] = R(0);
  __syncthreads();
  const Theta<R> th = load_theta<R>(theta);
  R gth[NTH];
#pragma unroll
  for (int p = 0; p < NTH; ++p) gth[p] = R(0);
  for (int64_t b = (int64_t)blockIdx.x * kBlock + threadIdx.x; b < B; b += (int64_t)gridDim.x * kBlock) {
    const R ie = iext[b], rm = rmod[b];
    const R* xs = sol + b * sb;
    const R* gs = gsol + b * gsb;
    R gie = R(0), grm = R(0);
    V4<R> lam = load4<R>(gs + (int64_t)(T - 1) * gst);
    R t1 = tgrid[T - 1];
    // the rows of grid point i are loaded one interval ahead of their use (loop-carried, so the compiler cannot
    // sink the loads to the point of use, where their latency would be exposed once per interval)
    V4<R> g_ahead = lam, x_ahead = lam;
    if (T >= 2) {
      g_ahead = load4<R>(gs + (int64_t)(T - 2) * gst);
      x_ahead = load4<R>(xs + (int64_t)(MODE == SLODE_BWD_DISCRETE ? T - 2 : T - 1) * st);
    }
    for (int i = T - 2; i >= 0; --i) {
      const R t0 = tgrid[i];
      const R h = nsub == 1 ? (t1 - t0) : (t1 - t0) / R(nsub);   // the common case pays no division
      const V4<R> g = g_ahead, xrow = x_ahead;
      if (i > 0) {
        g_ahead = load4<R>(gs + (int64_t)(i - 1) * gst);
        x_ahead = load4<R>(xs + (int64_t)(MODE == SLODE_BWD_DISCRETE ? i - 1 : i) * st);
      }
      if (MODE == SLODE_BWD_DISCRETE) {
        V4<R> ys[kMaxSub];
        V4<R> Y[4];
        ys[0] = xrow;
        for (int k = 0; k + 1 < nsub; ++k) ys[k + 1] = step<R, METHOD>(ys[k], h, ie, rm, th, Y);
        for (int k = nsub - 1; k >= 0; --k) {
          step<R, METHOD>(ys[k], h, ie, rm, th, Y);
          lam = step_adjoint<R, METHOD>(Y, h, ie, rm, th, lam, gie, grm, gth);
        }
      } else {
        V4<R> y = xrow;
        for (int k = 0; k < nsub; ++k) aug_step<R, METHOD>(y, lam, h, ie, rm, th, gie, grm, gth);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) lam.v[c] += g.v[c];
      t1 = t0;
    }
    store4<R>(grad_y0 + b * 4, lam);
    grad_iext[b] = gie;
    grad_rmod[b] = grm;
  }
  // block reduction of the shared-constant gradients: warp shuffle, then one shared atomic per warp
#pragma unroll
  for (int p = 0; p < NTH; ++p) {
    R v = gth[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[p], v);
  }
  __syncthreads();
  if (threadIdx.x < NTH) atomicAdd(grad_theta + threadIdx.x, red[threadIdx.x]);
}

static int device_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <class R>
static int fwd(int method, int64_t B, int T, int nsub, const void* t, const void* y0, const void* ie, const void* rm,
               const void* th, void* sol, int64_t st, int64_t sb, cudaStream_t s) {
  const int64_t tiles = (B + kBlock - 1) / kBlock;
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)device_sms() * 16);
#define GO(M) cvs_fwd_kernel<R, M><<<grid, kBlock, 0, s>>>(B, T, nsub, (const R*)t, (const R*)y0, (const R*)ie, \
                                                            (const R*)rm, (const R*)th, (R*)sol, st, sb)
  switch (method) {
    case SLODE_METHOD_EULER: GO(SLODE_METHOD_EULER); break;
    case SLODE_METHOD_MIDPOINT: GO(SLODE_METHOD_MIDPOINT); break;
    case SLODE_METHOD_RK4: GO(SLODE_METHOD_RK4); break;
    default: set_error("slode_cvs_fixed_fwd: unknown method %d", method); return SLODE_EINVAL;
  }
#undef GO
  SLODE_CUDA_TRY(cudaGetLastError());
  return SLODE_OK;
}

template <class R, int MODE>
static int bwd(int method, int64_t B, int T, int nsub, const void* t, const void* ie, const void* rm, const void* th,
               const void* sol, int64_t st, int64_t sb, const void* gsol, int64_t gst, int64_t gsb, void* gy0,
               void* gie, void* grm, void* gth, cudaStream_t s) {
  const int64_t tiles = (B + kBlock - 1) / kBlock;
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)device_sms() * 8);
#define GO(M)                                                                                                      \
  cvs_bwd_kernel<R, M, MODE><<<grid, kBlock, 0, s>>>(B, T, nsub, (const R*)t, (const R*)ie, (const R*)rm,          \
                                                     (const R*)th, (const R*)sol, st, sb, (const R*)gsol, gst, gsb, \
                                                     (R*)gy0, (R*)gie, (R*)grm, (R*)gth)
  switch (method) {
    case SLODE_METHOD_EULER: GO(SLODE_METHOD_EULER); break;
    case SLODE_METHOD_MIDPOINT: GO(SLODE_METHOD_MIDPOINT); break;
    case SLODE_METHOD_RK4: GO(SLODE_METHOD_RK4); break;
    default: set_error("slode_cvs_fixed_bwd: unknown method %d", method); return SLODE_EINVAL;
  }
#undef GO
  SLODE_CUDA_TRY(cudaGetLastError());
  return SLODE_OK;
}

static int check(const char* who, int dtype, int64_t B, int T, int nsub) {
  if (dtype != SLODE_F32 && dtype != SLODE_F64) {
    set_error("%s: unknown dtype %d", who, dtype);
    return SLODE_EINVAL;
  }
  if (B < 0 || T < 1 || nsub < 1 || nsub > kMaxSub) {
    set_error("%s: bad sizes B=%lld T=%d substeps=%d (1..%d)", who, (long long)B, T, nsub, kMaxSub);
    return SLODE_EINVAL;
  }
  return SLODE_OK;
}

}  // namespace cvs
}  // namespace slode

using namespace slode;

extern "C" int slode_cvs_fixed_fwd(int method, int dtype, int64_t B, int T, int substeps, const void* t,
                                   const void* y0, const void* i_ext, const void* r_tpr_mod, const void* theta,
                                   void* sol, int64_t sol_stride_t, int64_t sol_stride_b, void* stream) {
  int rc = cvs::check("slode_cvs_fixed_fwd", dtype, B, T, substeps);
  if (rc) return rc;
  if (!t || !theta || (B > 0 && (!y0 || !i_ext || !r_tpr_mod || !sol))) {
    set_error("slode_cvs_fixed_fwd: null pointer");
    return SLODE_EINVAL;
  }
  g_fwd_launches = 0;
  if (B == 0) return SLODE_OK;
  rc = (dtype == SLODE_F32)
           ? cvs::fwd<float>(method, B, T, substeps, t, y0, i_ext, r_tpr_mod, theta, sol, sol_stride_t, sol_stride_b,
                             (cudaStream_t)stream)
           : cvs::fwd<double>(method, B, T, substeps, t, y0, i_ext, r_tpr_mod, theta, sol, sol_stride_t,
                              sol_stride_b, (cudaStream_t)stream);
  if (rc == SLODE_OK) g_fwd_launches = 1;
  return rc;
}

extern "C" int slode_cvs_fixed_bwd(int method, int mode, int dtype, int64_t B, int T, int substeps, const void* t,
                                   const void* i_ext, const void* r_tpr_mod, const void* theta, const void* sol,
                                   int64_t sol_stride_t, int64_t sol_stride_b, const void* grad_sol,
                                   int64_t gsol_stride_t, int64_t gsol_stride_b, void* grad_y0, void* grad_i_ext,
                                   void* grad_r_tpr_mod, void* grad_theta, void* stream) {
  int rc = cvs::check("slode_cvs_fixed_bwd", dtype, B, T, substeps);
  if (rc) return rc;
  if (mode != SLODE_BWD_DISCRETE && mode != SLODE_BWD_TDE_ADJOINT) {
    set_error("slode_cvs_fixed_bwd: unknown mode %d", mode);
    return SLODE_EINVAL;
  }
  if (!t || !theta || !grad_theta ||
      (B > 0 && (!i_ext || !r_tpr_mod || !sol || !grad_sol || !grad_y0 || !grad_i_ext || !grad_r_tpr_mod))) {
    set_error("slode_cvs_fixed_bwd: null pointer");
    return SLODE_EINVAL;
  }
  g_bwd_launches = 0;
  if (B == 0) return SLODE_OK;
  cudaStream_t s = (cudaStream_t)stream;
#define ARGS method, B, T, substeps, t, i_ext, r_tpr_mod, theta, sol, sol_stride_t, sol_stride_b, grad_sol, \
             gsol_stride_t, gsol_stride_b, grad_y0, grad_i_ext, grad_r_tpr_mod, grad_theta, s
  if (dtype == SLODE_F32)
    rc = (mode == SLODE_BWD_DISCRETE) ? cvs::bwd<float, SLODE_BWD_DISCRETE>(ARGS) : cvs::bwd<float, SLODE_BWD_TDE_ADJOINT>(ARGS);
  else
    rc = (mode == SLODE_BWD_DISCRETE) ? cvs::bwd<double, SLODE_BWD_DISCRETE>(ARGS) : cvs::bwd<double, SLODE_BWD_TDE_ADJOINT>(ARGS);
#undef ARGS
  if (rc == SLODE_OK) g_bwd_launches = 1;
  return rc;
}
