// lib.cu — library-level entry points: error string, ABI version, device info,
// launch counter.
#include "common.cuh"
#include <atomic>
#include <string.h>

namespace gdr {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace gdr

extern "C" {

int gdr_abi_version(void) { return GDR_ABI_VERSION; }

const char* gdr_last_error(void) { return gdr::g_err; }

int64_t gdr_launch_count(void) { return gdr::g_launches.load(std::memory_order_relaxed); }

int gdr_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  GDR_CUDA(cudaGetDevice(&dev));
  int v = 0;
  if (sm_count) {
    GDR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    *sm_count = v;
  }
  if (cc_major) {
    GDR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    *cc_major = v;
  }
  if (cc_minor) {
    GDR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    *cc_minor = v;
  }
  return GDR_OK;
}

}  // extern "C"
