// kmeans_tc.cu — tensor-core (tcgen05 + TMA, 3xTF32) E-step.  Placeholder until
// the exact-fp32 path is parity-green on the GPU; precision_mode 1 reports
// GDR_EUNSUPPORTED so that nothing silently falls back.
#include "common.cuh"

namespace gdr {

int64_t kmeans_assign_tc_ws_bytes(int64_t, int64_t, int64_t) { return 256; }

int kmeans_assign_tc(int64_t, int64_t, int64_t, const float*, int64_t, const float*, int64_t, int32_t*,
                     const int32_t*, int32_t*, float*, void*, int64_t, cudaStream_t) {
  set_error("kmeans_assign: precision_mode 1 (tcgen05) is not built into this library yet");
  return GDR_EUNSUPPORTED;
}

}  // namespace gdr
