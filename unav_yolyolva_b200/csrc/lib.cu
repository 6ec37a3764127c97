// lib.cu — library-level entry points: version, per-thread error string, device check, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace unav {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<int> g_pdl_override{-1};      // unav_set_pdl: -1 = follow UNAV_PDL, 0 / 1 = forced (e.g. around a graph capture)

bool pdl_enabled() {
  const int o = g_pdl_override.load(std::memory_order_relaxed);
  if (o >= 0) return o == 1;
  static int v = -1;
  if (v < 0) {
    // measured on B200 (scripts/gemm_probe.py): -0.8 us per back-to-back tiny GEMM, no change on the whole forward,
    // so the attribute is opt-in (UNAV_PDL=1)
    const char* e = getenv("UNAV_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

long long* g_phase_buf = nullptr;    // unav_set_phase_trace: per-CTA clock stamps of the tcgen05 kernels (diagnostics)
int g_phase_cap = 0;

static std::mutex g_smem_attr_mutex;
void smem_attr_lock() { g_smem_attr_mutex.lock(); }
void smem_attr_unlock() { g_smem_attr_mutex.unlock(); }

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int finish_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace unav

extern "C" const char* unav_version(void) {
  return unav::kHalfF16 ? "unav_b200 0.1 (sm_100a, FP16 halves)" : "unav_b200 0.1 (sm_100a, BF16 halves)";
}
extern "C" const char* unav_last_error(void) { return unav::g_err; }
extern "C" long long unav_launch_count(void) { return unav::g_launches.load(); }
extern "C" int unav_set_pdl(int mode) {
  unav::g_pdl_override.store(mode < 0 ? -1 : (mode ? 1 : 0), std::memory_order_relaxed);
  return 0;
}
extern "C" int unav_set_phase_trace(long long* device_buf, int capacity_ctas) {
  unav::g_phase_buf = device_buf;
  unav::g_phase_cap = device_buf ? capacity_ctas : 0;
  return 0;
}

extern "C" int unav_check_device(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || dev >= n) {
    unav::set_error("no usable CUDA device (%s)", cudaGetErrorString(e));
    return UNAV_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { unav::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return (int)e; }
  if (prop.major != 10) {
    unav::set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
    return UNAV_ERR_UNSUPPORTED;
  }
  return 0;
}
