// lloyd.cu — the Lloyd loop of one k-means run as a single C-ABI call.
//
// Control flow of sklearn's _kmeans_single_lloyd (sklearn/cluster/_kmeans.py:630-758),
// the routine behind the reference's KMeans(...).fit() call sites
// (clustgdd_agent_transduct.py:105, clustgdd_agent_induct.py:134, distill_recsys.py:178):
//   repeat: E-step -> per-cluster sums -> relocate empties -> average -> centre shift
//   stop on identical labels (strict) or sum shift^2 <= tol; if not strict, one more E-step;
//   inertia of the final assignment.
// One iteration is ~25 short kernels; its body is captured ONCE per buffer parity into a CUDA
// graph (on an internal stream ordered after the caller's stream by events) and replayed, so
// the host issues one graph launch plus ONE 24-byte status read-back per iteration (the
// convergence test is a host decision in the reference as well).
#include "common.cuh"
#include <stdio.h>

using namespace gdr;

namespace gdr {
int64_t kmeans_assign_tc_ws_bytes(int64_t N, int64_t K, int64_t D);
int64_t kmeans_tc_xsplit_bytes(int64_t N, int64_t D);
int g_lloyd_graph = 1;  // gdr_debug_set("lloyd_graph", 0) disables graph replay
bool profiling_enabled();
}  // namespace gdr

namespace {

struct LloydBuffers {
  int32_t* labels[2];
  float* centers[2];
  float* sums;
  int32_t* counts;
  int32_t* n_changed;
  double* stats;  // [2 + K]
  void* xsplit;   // tensor-core operand cache (mode 1)
  void* ws_assign;
  int64_t ws_assign_b;
  void* ws_seg;
  int64_t ws_seg_b;
  void* ws_misc;
  int64_t ws_misc_b;
};

int64_t carve(LloydBuffers* B, void* ws, int64_t N, int64_t K, int64_t D, int mode) {
  const int64_t ldc = align_up(D, 4);
  Workspace W(ws, INT64_MAX);
  auto take_bytes = [&](int64_t bytes) { return (void*)W.take<char>(bytes); };
  LloydBuffers tmp;
  LloydBuffers& b = B ? *B : tmp;
  b.labels[0] = W.take<int32_t>(N);
  b.labels[1] = W.take<int32_t>(N);
  b.centers[0] = W.take<float>(K * ldc);
  b.centers[1] = W.take<float>(K * ldc);
  b.sums = W.take<float>(K * ldc);
  b.counts = W.take<int32_t>(K);
  b.n_changed = W.take<int32_t>(1);
  b.stats = W.take<double>(2 + K);
  b.xsplit = mode == 1 ? take_bytes(kmeans_tc_xsplit_bytes(N, D)) : nullptr;
  b.ws_assign_b = mode == 1 ? kmeans_assign_tc_ws_bytes(N, K, D) : gdr_kmeans_assign_ws_bytes(N, K, D, 0);
  b.ws_assign = take_bytes(b.ws_assign_b);
  b.ws_seg_b = gdr_segment_sum_ws_bytes(N, K, D);
  b.ws_seg = take_bytes(b.ws_seg_b);
  b.ws_misc_b = std::max(gdr_inertia_ws_bytes(N, D), gdr_kmeans_relocate_ws_bytes(N, K, D));
  b.ws_misc = take_bytes(b.ws_misc_b);
  return W.off + 256;
}

struct HostStatus {
  double stats[2];
  int32_t n_changed;
  int32_t pad;
};

HostStatus* pinned_status() {
  static thread_local HostStatus* p = nullptr;
  if (!p && cudaHostAlloc((void**)&p, sizeof(HostStatus), cudaHostAllocDefault) != cudaSuccess) p = nullptr;
  return p;
}

// RAII for the internal stream / graphs so that every early return cleans up
struct GraphCtx {
  cudaStream_t gs = nullptr;
  cudaEvent_t ev = nullptr;
  cudaGraphExec_t exec[2] = {nullptr, nullptr};
  ~GraphCtx() {
    for (auto& e : exec)
      if (e) cudaGraphExecDestroy(e);
    if (ev) cudaEventDestroy(ev);
    if (gs) cudaStreamDestroy(gs);
  }
};

}  // namespace

extern "C" {

int64_t gdr_kmeans_lloyd_ws_bytes(int64_t N, int64_t K, int64_t D, int precision_mode) {
  return carve(nullptr, nullptr, N, K, D, precision_mode);
}

int gdr_kmeans_lloyd(int64_t N, int64_t K, int64_t D, const float* Xc, int64_t ldx, float* C_inout,
                     int64_t ldc, int32_t* labels_out, int max_iter, double tol_abs, int precision_mode,
                     double* inertia_out_host, int32_t* n_iter_out_host, int32_t* info_out_host, int verbose,
                     void* ws, int64_t ws_bytes, gdr_stream_t stream) {
  GDR_CHECK_ARG(N > 0 && K > 0 && D > 0 && Xc && C_inout && labels_out && ws, "kmeans_lloyd: bad arguments");
  GDR_CHECK_ARG(N >= K, "kmeans_lloyd: n_samples=%lld should be >= n_clusters=%lld", (long long)N, (long long)K);
  GDR_CHECK_ARG(precision_mode == 0 || precision_mode == 1, "kmeans_lloyd: precision_mode");
  GDR_CHECK_ARG(max_iter >= 0, "kmeans_lloyd: max_iter");
  if (ws_bytes < gdr_kmeans_lloyd_ws_bytes(N, K, D, precision_mode)) {
    set_error("kmeans_lloyd: workspace too small");
    return GDR_EWORKSPACE;
  }
  cudaStream_t caller = (cudaStream_t)stream;
  HostStatus* hs = pinned_status();
  if (!hs) {
    set_error("kmeans_lloyd: cannot allocate the pinned status block");
    return GDR_ECUDA;
  }
  // Graph replay needs a capturable stream: run the whole loop on an internal stream that is
  // ordered after the caller's stream (and the caller's stream after it at the end).
  GraphCtx G;
  bool use_graph = g_lloyd_graph != 0 && !profiling_enabled() && max_iter >= 3;
  cudaStream_t s = caller;
  if (use_graph) {
    if (cudaStreamCreateWithFlags(&G.gs, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&G.ev, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      use_graph = false;
    } else {
      GDR_CUDA(cudaEventRecord(G.ev, caller));
      GDR_CUDA(cudaStreamWaitEvent(G.gs, G.ev, 0));
      s = G.gs;
    }
  }
  gdr_stream_t st = (gdr_stream_t)s;

  LloydBuffers B;
  carve(&B, ws, N, K, D, precision_mode);
  const int64_t ldw = align_up(D, 4);
  int rc;
  GDR_CUDA(cudaMemsetAsync(B.centers[0], 0, K * ldw * 4, s));
  GDR_CUDA(cudaMemsetAsync(B.centers[1], 0, K * ldw * 4, s));
  GDR_CUDA(cudaMemcpy2DAsync(B.centers[0], ldw * 4, C_inout, ldc * 4, D * 4, K, cudaMemcpyDeviceToDevice, s));
  GDR_CUDA(cudaMemsetAsync(B.labels[1], 0xff, N * 4, s));  // labels_old = -1
  if (precision_mode == 1) {
    rc = gdr_kmeans_tc_prepare(N, D, Xc, ldx, B.xsplit, kmeans_tc_xsplit_bytes(N, D), st);
    if (rc) return rc;
  }
  auto assign = [&](const float* C, int32_t* lab, const int32_t* prev, int32_t* nchg) -> int {
    if (precision_mode == 1)
      return gdr_kmeans_assign_tc(N, K, D, Xc, ldx, B.xsplit, C, ldw, lab, prev, nchg, nullptr, nullptr,
                                  B.ws_assign, B.ws_assign_b, st);
    return gdr_kmeans_assign(N, K, D, Xc, ldx, C, ldw, lab, prev, nchg, nullptr, 0, B.ws_assign, B.ws_assign_b, st);
  };
  // one iteration for buffer parity p (centres[p] -> centres[1-p], labels[p] new, labels[1-p] old),
  // ending with the status copy into pinned host memory
  auto enqueue_iteration = [&](int p) -> int {
    const int cur = p, nxt = 1 - p, lab_new = p, lab_old = 1 - p;
    GDR_CUDA(cudaMemsetAsync(B.n_changed, 0, 4, s));
    int r;
    if ((r = assign(B.centers[cur], B.labels[lab_new], B.labels[lab_old], B.n_changed))) return r;
    if ((r = gdr_segment_sum(N, K, D, Xc, ldx, B.labels[lab_new], B.sums, ldw, B.counts, B.ws_seg, B.ws_seg_b, st)))
      return r;
    if ((r = gdr_kmeans_finalize(K, D, B.sums, ldw, B.counts, B.centers[cur], ldw, B.centers[nxt], ldw, B.stats, 0, st)))
      return r;
    GDR_CUDA(cudaMemcpyAsync(hs->stats, B.stats, 16, cudaMemcpyDeviceToHost, s));
    GDR_CUDA(cudaMemcpyAsync(&hs->n_changed, B.n_changed, 4, cudaMemcpyDeviceToHost, s));
    return GDR_OK;
  };
  auto run_iteration = [&](int p) -> int {
    if (use_graph) {
      if (!G.exec[p]) {
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          int r = enqueue_iteration(p);
          cudaError_t e = cudaStreamEndCapture(s, &graph);
          if (r == GDR_OK && e == cudaSuccess && graph &&
              cudaGraphInstantiate(&G.exec[p], graph, 0) == cudaSuccess) {
            cudaGraphDestroy(graph);
          } else {
            if (graph) cudaGraphDestroy(graph);
            G.exec[p] = nullptr;
            cudaGetLastError();
            use_graph = false;  // capture not possible here: issue the launches directly
            if (r != GDR_OK) return r;
          }
        } else {
          cudaGetLastError();
          use_graph = false;
        }
      }
      if (use_graph && G.exec[p]) {
        GDR_CUDA(cudaGraphLaunch(G.exec[p], s));
        count_launch(1);
        return GDR_OK;
      }
    }
    return enqueue_iteration(p);
  };

  int p = 0;
  bool strict = false;
  int n_iter = 0, relocations = 0;
  for (int i = 0; i < max_iter; ++i) {
    n_iter = i + 1;
    if ((rc = run_iteration(p))) return rc;
    GDR_CUDA(cudaStreamSynchronize(s));
    double shift_tot = hs->stats[0];
    const int n_empty = (int)hs->stats[1];
    const int n_changed = hs->n_changed;
    const int cur = p, nxt = 1 - p;
    if (n_empty > 0) {
      // _relocate_empty_clusters_dense (sklearn/_k_means_common.pyx:167-211), then re-average
      ++relocations;
      if ((rc = gdr_kmeans_relocate(N, K, D, Xc, ldx, B.centers[cur], ldw, B.labels[p], B.sums, ldw, B.counts,
                                    B.ws_misc, B.ws_misc_b, st)))
        return rc;
      if ((rc = gdr_kmeans_finalize(K, D, B.sums, ldw, B.counts, B.centers[cur], ldw, B.centers[nxt], ldw, B.stats, 0,
                                    st)))
        return rc;
      GDR_CUDA(cudaMemcpyAsync(hs->stats, B.stats, 16, cudaMemcpyDeviceToHost, s));
      GDR_CUDA(cudaStreamSynchronize(s));
      shift_tot = hs->stats[0];
    }
    if (verbose) printf("Iteration %d, center shift %.6g, labels changed %d.\n", i, shift_tot, n_changed);
    if (n_changed == 0) {  // np.array_equal(labels, labels_old)  (:723-729)
      strict = true;
      p = 1 - p;  // centres[nxt] are current; labels[cur parity] stay the newest
      break;
    }
    p = 1 - p;
    if (shift_tot <= tol_abs) break;  // (:731-738)
  }
  // after the loop: centres[p] are the current centres, labels[1 - p] the newest labels
  // (max_iter == 0: centres[0], labels undefined -> the E-step below fills them)
  int32_t* labels = n_iter == 0 ? B.labels[0] : B.labels[1 - p];
  if (!strict) {
    // rerun the E-step so that the labels match the final centres (:742-754)
    if ((rc = assign(B.centers[p], labels, nullptr, nullptr))) return rc;
  }
  double* inertia_dev = B.stats;
  if ((rc = gdr_inertia(N, D, Xc, ldx, B.centers[p], ldw, labels, inertia_dev, B.ws_misc, B.ws_misc_b, st))) return rc;
  GDR_CUDA(cudaMemcpyAsync(hs->stats, inertia_dev, 8, cudaMemcpyDeviceToHost, s));
  GDR_CUDA(cudaMemcpyAsync(labels_out, labels, N * 4, cudaMemcpyDeviceToDevice, s));
  GDR_CUDA(cudaMemcpy2DAsync(C_inout, ldc * 4, B.centers[p], ldw * 4, D * 4, K, cudaMemcpyDeviceToDevice, s));
  GDR_CUDA(cudaStreamSynchronize(s));
  if (s != caller) {
    // later work on the caller's stream must see our results (already complete: we synchronised)
    GDR_CUDA(cudaEventRecord(G.ev, s));
    GDR_CUDA(cudaStreamWaitEvent(caller, G.ev, 0));
  }
  if (inertia_out_host) *inertia_out_host = hs->stats[0];
  if (n_iter_out_host) *n_iter_out_host = n_iter;
  if (info_out_host) {
    info_out_host[0] = strict ? 1 : 0;
    info_out_host[1] = relocations;
  }
  return GDR_OK;
}

}  // extern "C"
