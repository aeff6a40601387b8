// lib.cu — library-level entry points: error string, ABI version, device info,
// launch counter.
#include "common.cuh"
#include <atomic>
#include <mutex>
#include <string.h>
#include <utility>
#include <vector>

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

// ---- optional per-kernel timing (bench.py's roofline leg) ---------------------------
// When a kind is enabled every launch site of that kind is bracketed by a CUDA event pair
// recorded on the launching stream; gdr_profile_collect() sums the elapsed times.
static int g_prof_kind = 0;
static std::mutex g_prof_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;

ProfileScope::ProfileScope(int kind, cudaStream_t s) : stop_(nullptr), stream_(s) {
  if (kind != g_prof_kind || kind == 0) return;
  cudaEvent_t a, b;
  if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
  cudaEventRecord(a, s);
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_events.emplace_back(a, b);
  }
  stop_ = b;
}
bool profiling_enabled() { return g_prof_kind != 0; }

ProfileScope::~ProfileScope() {
  if (stop_) cudaEventRecord((cudaEvent_t)stop_, stream_);
}

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

int gdr_profile_enable(int kind) {
  std::lock_guard<std::mutex> lk(gdr::g_prof_mu);
  for (auto& e : gdr::g_prof_events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  gdr::g_prof_events.clear();
  gdr::g_prof_kind = kind;
  return GDR_OK;
}

int gdr_profile_collect(double* total_ms_host, int64_t* launches_host) {
  std::lock_guard<std::mutex> lk(gdr::g_prof_mu);
  double tot = 0.0;
  for (auto& e : gdr::g_prof_events) {
    GDR_CUDA(cudaEventSynchronize(e.second));
    float ms = 0.f;
    GDR_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
    tot += ms;
  }
  if (total_ms_host) *total_ms_host = tot;
  if (launches_host) *launches_host = (int64_t)gdr::g_prof_events.size();
  return GDR_OK;
}

}  // extern "C"
