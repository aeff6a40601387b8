// kmeans_init.cu — greedy k-means++ seeding on the device.
//
// Algorithm of sklearn's _kmeans_plusplus (sklearn/cluster/_kmeans.py:180-278), the default
// initialisation behind the reference's `KMeans(n_clusters=n).fit(...)`
// (clustgdd_agent_transduct.py:105, distill_recsys.py:178):
//   first centre = a uniformly drawn sample; then for every further centre draw
//   n_local_trials = 2 + int(log K) candidates with probability proportional to the squared
//   distance to the closest centre so far (searchsorted on the cumulative sum), keep the
//   candidate that lowers the total potential most.
// The random numbers are drawn by the HOST from numpy's RandomState exactly as sklearn
// consumes them (one choice(), then uniform(size=n_local_trials) per round) and handed in;
// every arithmetic step runs here, and the K rounds need no host synchronisation.
// Potentials / cumulative sums are accumulated in fp64 in a fixed order (deterministic); sklearn
// uses an fp32 cumsum, so a draw that lands within rounding of a boundary may pick a neighbour.
#include "common.cuh"

namespace gdr {

constexpr int PP_THREADS = 256;
constexpr int PP_ROWS_PER_BLOCK = 1024;   // rows summarised by one block-sum entry
constexpr int PP_MAX_TRIALS = 16;   // 2 + int(log K) <= 16 up to K ~ 1.2e6

struct PpState {
  int64_t cur_id;      // centre chosen in the previous round
  double pot;          // current potential
  int64_t cand[PP_MAX_TRIALS];
};

// closest[i] = min(closest[i], |x_i - x_cur|^2)  (first round: plain assignment);
// bsum[b] = sum of closest over the block's rows (fp64, fixed order)
__global__ void __launch_bounds__(PP_THREADS) k_pp_update(int64_t N, int D, const float* __restrict__ X, int64_t ldx,
                                                          const PpState* __restrict__ st, int first,
                                                          float* __restrict__ closest, double* __restrict__ bsum) {
  extern __shared__ float s_c[];   // D floats: the current centre
  __shared__ double s_w[PP_THREADS / 32];
  const int64_t cur = st->cur_id;
  for (int k = threadIdx.x; k < D; k += PP_THREADS) s_c[k] = X[cur * ldx + k];
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = (int64_t)blockIdx.x * PP_ROWS_PER_BLOCK;
  const int64_t r1 = min(N, r0 + PP_ROWS_PER_BLOCK);
  double acc = 0.0;
  for (int64_t r = r0 + w; r < r1; r += PP_THREADS / 32) {
    float d = 0.f;
    for (int k = lane; k < D; k += 32) {
      float t = __fsub_rn(X[r * ldx + k], s_c[k]);
      d = fmaf(t, t, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    float c = first ? d : fminf(closest[r], d);
    if (lane == 0) closest[r] = c;
    acc += (double)c;
  }
  if (lane == 0) s_w[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PP_THREADS / 32; ++i) s += s_w[i];
    bsum[blockIdx.x] = s;
  }
}

// single CTA: prefix over the block sums, potential, and the searchsorted of the L draws
__global__ void __launch_bounds__(1024) k_pp_pick(int64_t N, int nblocks, const double* __restrict__ bsum,
                                                  double* __restrict__ bprefix /*[nblocks + 1]*/,
                                                  const float* __restrict__ closest,
                                                  const double* __restrict__ rand_vals, int L,
                                                  PpState* __restrict__ st) {
  __shared__ double s_part[1024];
  const int t = threadIdx.x;
  const int per = (nblocks + 1023) / 1024;
  const int b0 = t * per, b1 = min(nblocks, b0 + per);
  double s = 0.0;
  for (int b = b0; b < b1; ++b) s += bsum[b];
  s_part[t] = s;
  __syncthreads();
  // exclusive scan of the 1024 partials (Hillis-Steele on shared memory, fixed order)
  for (int o = 1; o < 1024; o <<= 1) {
    double v = t >= o ? s_part[t - o] : 0.0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  double run = t == 0 ? 0.0 : s_part[t - 1];
  for (int b = b0; b < b1; ++b) {
    bprefix[b] = run;
    run += bsum[b];
  }
  if (t == 1023) bprefix[nblocks] = s_part[1023];
  __syncthreads();
  const double pot = bprefix[nblocks];
  if (t == 0) st->pot = pot;
  // warp l resolves draw l: first index i with cumsum[i] >= rand * pot  (np.searchsorted, side='left')
  const int w = t >> 5, lane = t & 31;
  if (w < L) {
    const double thr = rand_vals[w] * pot;
    // first block whose inclusive prefix reaches thr
    int lo = 0, hi = nblocks;  // search in [0, nblocks)
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (bprefix[mid + 1] >= thr) hi = mid;
      else lo = mid + 1;
    }
    int64_t found = N - 1;  // np.clip(candidate_ids, None, n - 1)
    if (lo < nblocks) {
      const int64_t r0 = (int64_t)lo * PP_ROWS_PER_BLOCK, r1 = min(N, r0 + PP_ROWS_PER_BLOCK);
      double base = bprefix[lo];
      bool done = false;
      for (int64_t r = r0; r < r1 && !done; r += 32) {
        int64_t i = r + lane;
        double v = i < r1 ? (double)closest[i] : 0.0;
        double incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          double u = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += u;
        }
        unsigned hit = __ballot_sync(0xffffffffu, i < r1 && base + incl >= thr);
        if (hit) {
          found = r + (__ffs(hit) - 1);
          done = true;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    if (lane == 0) st->cand[w] = found;
  }
}

// per block: for every candidate l the sum over the block's rows of min(closest_i, |x_i - x_cand_l|^2)
__global__ void __launch_bounds__(PP_THREADS) k_pp_score(int64_t N, int D, const float* __restrict__ X, int64_t ldx,
                                                         const PpState* __restrict__ st, int L,
                                                         const float* __restrict__ closest,
                                                         double* __restrict__ bpot /*[nblocks][L]*/) {
  extern __shared__ float s_cand[];   // L * D floats
  __shared__ double s_w[PP_THREADS / 32][PP_MAX_TRIALS];
  for (int i = threadIdx.x; i < L * D; i += PP_THREADS) {
    int l = i / D, k = i - l * D;
    s_cand[i] = X[st->cand[l] * ldx + k];
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = (int64_t)blockIdx.x * PP_ROWS_PER_BLOCK;
  const int64_t r1 = min(N, r0 + PP_ROWS_PER_BLOCK);
  double acc[PP_MAX_TRIALS];
#pragma unroll
  for (int l = 0; l < PP_MAX_TRIALS; ++l) acc[l] = 0.0;
  for (int64_t r = r0 + w; r < r1; r += PP_THREADS / 32) {
    const float cl = closest[r];
#pragma unroll
    for (int l = 0; l < PP_MAX_TRIALS; ++l) {
      if (l >= L) break;
      float d = 0.f;
      for (int k = lane; k < D; k += 32) {
        float t = __fsub_rn(X[r * ldx + k], s_cand[l * D + k]);
        d = fmaf(t, t, d);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      acc[l] += (double)fminf(cl, d);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < PP_MAX_TRIALS; ++l)
      if (l < L) s_w[w][l] = acc[l];
  }
  __syncthreads();
  if (threadIdx.x < L) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PP_THREADS / 32; ++i) s += s_w[i][threadIdx.x];
    bpot[(int64_t)blockIdx.x * L + threadIdx.x] = s;
  }
}

// single CTA: total potential per candidate (fixed order), first argmin, commit the winner
__global__ void __launch_bounds__(1024) k_pp_choose(int nblocks, int L, int D, const double* __restrict__ bpot,
                                                    const float* __restrict__ X, int64_t ldx,
                                                    PpState* __restrict__ st, float* __restrict__ centers, int64_t ldc,
                                                    int64_t* __restrict__ indices, int c) {
  __shared__ double s_pot[PP_MAX_TRIALS];
  __shared__ int s_best;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w < L) {
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += bpot[(int64_t)b * L + w];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_pot[w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = 0;
    for (int l = 1; l < L; ++l)
      if (s_pot[l] < s_pot[best]) best = l;   // np.argmin: first minimum
    s_best = best;
    st->pot = s_pot[best];
    st->cur_id = st->cand[best];
    if (indices) indices[c] = st->cand[best];
  }
  __syncthreads();
  const int64_t id = st->cand[s_best];
  for (int k = threadIdx.x; k < D; k += 1024) centers[(int64_t)c * ldc + k] = X[id * ldx + k];
}

__global__ void k_pp_first(int64_t first, int D, const float* __restrict__ X, int64_t ldx, PpState* __restrict__ st,
                           float* __restrict__ centers, int64_t* __restrict__ indices) {
  if (threadIdx.x == 0) {
    st->cur_id = first;
    st->pot = 0.0;
    if (indices) indices[0] = first;
  }
  for (int k = threadIdx.x; k < D; k += blockDim.x) centers[k] = X[first * ldx + k];
}

}  // namespace gdr

using namespace gdr;

extern "C" {

int64_t gdr_kmeans_plusplus_ws_bytes(int64_t N, int64_t K, int64_t D, int n_trials) {
  (void)D;
  int64_t nb = cdiv(N > 0 ? N : 1, PP_ROWS_PER_BLOCK);
  return ws_need(N, 4) + 2 * ws_need(nb + 1, 8) + ws_need(nb * n_trials, 8) + ws_need((K > 1 ? K - 1 : 1) * n_trials, 8) +
         ws_need(1, sizeof(PpState)) + 256;
}

int gdr_kmeans_plusplus(int64_t N, int64_t K, int64_t D, const float* X, int64_t ldx, int64_t first_center,
                        const double* rand_vals_host, int n_trials, float* centers_out, int64_t ldc,
                        int64_t* indices_out_dev, void* ws, int64_t ws_bytes, gdr_stream_t stream) {
  GDR_CHECK_ARG(N > 0 && K > 0 && D > 0 && X && centers_out && ws, "kmeans_plusplus: bad arguments");
  GDR_CHECK_ARG(K <= N, "kmeans_plusplus: n_samples=%lld should be >= n_clusters=%lld", (long long)N, (long long)K);
  GDR_CHECK_ARG(first_center >= 0 && first_center < N, "kmeans_plusplus: first centre out of range");
  GDR_CHECK_ARG(n_trials >= 1 && n_trials <= PP_MAX_TRIALS, "kmeans_plusplus: n_trials must be in [1, 16]");
  GDR_CHECK_ARG(K == 1 || rand_vals_host, "kmeans_plusplus: null random numbers");
  GDR_CHECK_ARG(ldx >= D && ldc >= D, "kmeans_plusplus: leading dimension");
  if (ws_bytes < gdr_kmeans_plusplus_ws_bytes(N, K, D, n_trials)) {
    set_error("kmeans_plusplus: workspace too small");
    return GDR_EWORKSPACE;
  }
  const size_t smem_score = (size_t)n_trials * D * 4, smem_upd = (size_t)D * 4;
  if (smem_score > 200 * 1024 || smem_upd > 48 * 1024) {
    set_error("kmeans_plusplus: n_trials * D = %lld floats exceeds the shared-memory plan", (long long)n_trials * D);
    return GDR_EUNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = (int)cdiv(N, PP_ROWS_PER_BLOCK);
  Workspace W(ws, ws_bytes);
  float* closest = W.take<float>(N);
  double* bsum = W.take<double>(nb + 1);
  double* bprefix = W.take<double>(nb + 1);
  double* bpot = W.take<double>((int64_t)nb * n_trials);
  double* rand_dev = W.take<double>((K > 1 ? K - 1 : 1) * n_trials);
  PpState* st = W.take<PpState>(1);
  if (K > 1)
    GDR_CUDA(cudaMemcpyAsync(rand_dev, rand_vals_host, (K - 1) * n_trials * 8, cudaMemcpyHostToDevice, s));
  if (smem_score > 48 * 1024)
    GDR_CUDA(cudaFuncSetAttribute(k_pp_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_score));
  k_pp_first<<<1, 256, 0, s>>>(first_center, (int)D, X, ldx, st, centers_out, indices_out_dev);
  GDR_LAUNCHED();
  for (int64_t c = 1; c < K; ++c) {
    k_pp_update<<<nb, PP_THREADS, smem_upd, s>>>(N, (int)D, X, ldx, st, c == 1, closest, bsum);
    GDR_LAUNCHED();
    k_pp_pick<<<1, 1024, 0, s>>>(N, nb, bsum, bprefix, closest, rand_dev + (c - 1) * n_trials, n_trials, st);
    GDR_LAUNCHED();
    k_pp_score<<<nb, PP_THREADS, smem_score, s>>>(N, (int)D, X, ldx, st, n_trials, closest, bpot);
    GDR_LAUNCHED();
    k_pp_choose<<<1, 1024, 0, s>>>(nb, n_trials, (int)D, bpot, X, ldx, st, centers_out, ldc, indices_out_dev, (int)c);
    GDR_LAUNCHED();
  }
  // the pageable host buffer must stay valid until the copy above has executed
  if (K > 1) GDR_CUDA(cudaStreamSynchronize(s));
  return GDR_OK;
}

}  // extern "C"
