// Host-side plumbing shared by all slode_b200 entry points: error strings, device queries and the
// per-device guard that orders users of the packed-weight constant buffer.
#include <stdarg.h>

#include <mutex>

#include "slode_common.cuh"

namespace slode {

static thread_local char g_err[512] = "";
LaunchCount g_fwd_launches;
LaunchCount g_bwd_launches;
std::atomic<long long> g_total_launches{0};
LaunchCount& LaunchCount::operator=(int v) {
  last.store(v);
  if (v > 0) g_total_launches.fetch_add(v);
  return *this;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return SLODE_ECUDA;
}

namespace {
constexpr int kMaxDevices = 64;
struct DeviceState {
  bool init = false;
  int sms = 0;
  cudaEvent_t pack_done = nullptr;
  float* staging = nullptr;
  float* flip_ws = nullptr;
  size_t flip_bytes = 0;
};
DeviceState g_dev[kMaxDevices];
std::mutex g_mu;

DeviceState* current_device_state() {  // g_mu held
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  DeviceState& d = g_dev[dev];
  if (!d.init) {
    if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&d.pack_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaMalloc(&d.staging, sizeof(float) * 16384) != cudaSuccess) return nullptr;
    d.init = true;
  }
  return &d;
}
}  // namespace

float* flip_workspace(size_t bytes) {  // g_mu is held by the caller's PackGuard
  DeviceState* d = current_device_state();
  if (!d) {
    cuda_fail(cudaGetLastError(), "device state");
    return nullptr;
  }
  if (bytes > d->flip_bytes) {
    // kernels of earlier calls may still be using the old buffer: all of them are ordered before pack_done
    if (d->flip_ws) {
      if (cudaEventSynchronize(d->pack_done) != cudaSuccess || cudaFree(d->flip_ws) != cudaSuccess) {
        cuda_fail(cudaGetLastError(), "cudaFree(flip workspace)");
        return nullptr;
      }
      d->flip_ws = nullptr;
      d->flip_bytes = 0;
    }
    const cudaError_t e = cudaMalloc(&d->flip_ws, bytes);
    if (e != cudaSuccess) {
      cuda_fail(e, "cudaMalloc(flip workspace)");
      return nullptr;
    }
    d->flip_bytes = bytes;
  }
  return d->flip_ws;
}

PackGuard::PackGuard(cudaStream_t s) : status(0), stream(s), staging(nullptr), sms(148) {
  g_mu.lock();
  DeviceState* d = current_device_state();
  if (!d) {
    status = cuda_fail(cudaGetLastError(), "device state init");
    return;
  }
  staging = d->staging;
  sms = d->sms;
  // The constant buffer is shared by every call on this device: wait for the previous user.
  cudaError_t e = cudaStreamWaitEvent(stream, d->pack_done, 0);
  if (e != cudaSuccess) status = cuda_fail(e, "cudaStreamWaitEvent(pack_done)");
}

PackGuard::~PackGuard() {
  DeviceState* d = current_device_state();
  if (d && status == 0) cudaEventRecord(d->pack_done, stream);
  g_mu.unlock();
}

}  // namespace slode

extern "C" const char* slode_last_error(void) { return slode::g_err; }
