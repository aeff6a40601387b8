/*
 * oracle.c — CPU restatement of the floating-point inner loops of the distillation
 * core.  TEST INFRASTRUCTURE ONLY: nothing under oracle/ is imported by the product
 * package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker or the timed baseline.
 *
 * Each function restates the algorithm of the reference call site it names
 * (paths relative to /root/reference/ClustGDD; "sklearn/" = scikit-learn 1.9.0's
 * sklearn/cluster, the third-party package that owns the k-means arithmetic — the
 * reference pins 1.3.2 in README.md:14, same Lloyd code).  Plain sequential
 * arithmetic, one fp32 (or fp64) chain per output, compiled with
 * -ffp-contract=off so no FMA is introduced: the result is a function of the
 * inputs only, not of thread count or vector width.  OpenMP splits ROWS only.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC oracle.c -o _build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------
 * Stage 2.  prop = (alpha * A) @ X ; target += beta * prop
 *   clustgdd_agent_transduct.py:64-65  (`alpha*adj_norm @ prop_feat` — Python
 *   precedence scales the sparse VALUES first: fp32(v * fp32(alpha)), then SpMM)
 * ---------------------------------------------------------------------------- */
void oracle_spmm_prop(int64_t rows, int64_t F, const int32_t* rowptr, const int32_t* colidx,
                      const float* vals, float alpha, const float* X, int64_t ldx, float* Y,
                      int64_t ldy, float* T, int64_t ldt, float beta) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < rows; ++r) {
    float* y = Y + r * ldy;
    for (int64_t c = 0; c < F; ++c) y[c] = 0.0f;
    for (int32_t j = rowptr[r]; j < rowptr[r + 1]; ++j) {
      float av = vals ? vals[j] * alpha : alpha;
      const float* x = X + (int64_t)colidx[j] * ldx;
      for (int64_t c = 0; c < F; ++c) y[c] = y[c] + av * x[c];
    }
    if (T) {
      float* t = T + r * ldt;
      for (int64_t c = 0; c < F; ++c) t[c] = t[c] + beta * y[c];
    }
  }
}

/* fp64 evaluation of the same hop (used to measure how far BOTH fp32 paths sit from the
 * exact value; the 1e-5 contract of BASELINE.json is relative to max|X|). */
void oracle_spmm_prop_f64(int64_t rows, int64_t F, const int32_t* rowptr, const int32_t* colidx,
                          const float* vals, float alpha, const double* X, int64_t ldx, double* Y,
                          int64_t ldy) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < rows; ++r) {
    double* y = Y + r * ldy;
    for (int64_t c = 0; c < F; ++c) y[c] = 0.0;
    for (int32_t j = rowptr[r]; j < rowptr[r + 1]; ++j) {
      double av = (double)(vals ? vals[j] * alpha : alpha);
      const double* x = X + (int64_t)colidx[j] * ldx;
      for (int64_t c = 0; c < F; ++c) y[c] += av * x[c];
    }
  }
}

/* ------------------------------------------------------------------------------
 * Stage 3, E-step.   sklearn/_k_means_lloyd.pyx:196-213
 *   pairwise = |c_j|^2 (row_norms, fp32) ; pairwise += -2 * X.C^T (sgemm) ;
 *   label = first j with the strictly smallest value.
 * fp32 version = the reference's arithmetic up to BLAS summation order.
 * ---------------------------------------------------------------------------- */
void oracle_kmeans_assign_f32(int64_t N, int64_t K, int64_t D, const float* X, int64_t ldx,
                              const float* C, int64_t ldc, int32_t* labels, float* best_out) {
  float* cn = (float*)malloc(sizeof(float) * (size_t)K);
  for (int64_t j = 0; j < K; ++j) {
    float s = 0.0f;
    for (int64_t k = 0; k < D; ++k) s = s + C[j * ldc + k] * C[j * ldc + k];
    cn[j] = s;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i) {
    const float* x = X + i * ldx;
    float best = 0.0f;
    int32_t bj = 0;
    for (int64_t j = 0; j < K; ++j) {
      const float* c = C + j * ldc;
      float dot = 0.0f;
      for (int64_t k = 0; k < D; ++k) dot = dot + x[k] * c[k];
      float d = cn[j] + (-2.0f) * dot;
      if (j == 0 || d < best) {
        best = d;
        bj = (int32_t)j;
      }
    }
    labels[i] = bj;
    if (best_out) best_out[i] = best;
  }
  free(cn);
}

/* fp64 "truth": nearest and second-nearest squared distance |x-c|^2 in double.
 * margin_out[i] = (d2 - d1) / max(d2, 1e-30) — the relative distance margin of
 * BASELINE.json's contract ("bit-exact wherever the margin exceeds 1e-6 relative");
 * second_out[i] = index of the runner-up (inside the band either answer is accepted). */
void oracle_kmeans_assign_f64(int64_t N, int64_t K, int64_t D, const float* X, int64_t ldx,
                              const float* C, int64_t ldc, int32_t* labels, int32_t* second_out,
                              double* margin_out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i) {
    const float* x = X + i * ldx;
    double d1 = INFINITY, d2 = INFINITY;
    int32_t j1 = 0, j2 = 0;
    for (int64_t j = 0; j < K; ++j) {
      const float* c = C + j * ldc;
      double s = 0.0;
      for (int64_t k = 0; k < D; ++k) {
        double t = (double)x[k] - (double)c[k];
        s += t * t;
      }
      if (s < d1) {
        d2 = d1;
        j2 = j1;
        d1 = s;
        j1 = (int32_t)j;
      } else if (s < d2) {
        d2 = s;
        j2 = (int32_t)j;
      }
    }
    labels[i] = j1;
    if (second_out) second_out[i] = K > 1 ? j2 : j1;
    if (margin_out) margin_out[i] = K > 1 ? (d2 - d1) / (d2 > 1e-30 ? d2 : 1e-30) : 1.0;
  }
}

/* ------------------------------------------------------------------------------
 * Stage 3, M-step.   sklearn/_k_means_lloyd.pyx:215-218 with n_threads = 1:
 *   weight[label] += 1 ; centers_new[label, :] += X[i, :]   in sample order, fp32.
 * (Also the index_add_ + bincount pooling of distill_recsys.py:628-636.)
 * ---------------------------------------------------------------------------- */
void oracle_segment_sum(int64_t N, int64_t K, int64_t D, const float* X, int64_t ldx,
                        const int32_t* labels, float* sums, int64_t lds, int32_t* counts) {
  for (int64_t j = 0; j < K; ++j) {
    counts[j] = 0;
    for (int64_t k = 0; k < D; ++k) sums[j * lds + k] = 0.0f;
  }
  for (int64_t i = 0; i < N; ++i) {
    int32_t l = labels[i];
    counts[l] += 1;
    float* s = sums + (int64_t)l * lds;
    const float* x = X + i * ldx;
    for (int64_t k = 0; k < D; ++k) s[k] = s[k] + x[k];
  }
}

/* _average_centers + _center_shift   sklearn/_k_means_common.pyx:274-311
 * alpha = 1.0f / weight ; centers *= alpha ; empty -> row of the first largest cluster, copied IN PLACE while j
 * walks upwards as sklearn does: an empty j < argmax sees that row before it is scaled (raw sums), j > argmax after.
 * Returns sum_j |new_j - old_j|^2 (accumulated in double). */
double oracle_kmeans_finalize(int64_t K, int64_t D, const float* sums, int64_t lds,
                              const int32_t* counts, const float* C_old, int64_t ldo, float* C_new,
                              int64_t ldn) {
  int64_t amax = 0;
  for (int64_t j = 1; j < K; ++j)
    if (counts[j] > counts[amax]) amax = j;
  for (int64_t j = 0; j < K; ++j) {
    if (counts[j] > 0) {
      float alpha = 1.0f / (float)counts[j];
      for (int64_t k = 0; k < D; ++k) C_new[j * ldn + k] = sums[j * lds + k] * alpha;
    }
  }
  for (int64_t j = 0; j < K; ++j) {
    if (counts[j] <= 0)
      for (int64_t k = 0; k < D; ++k)
        C_new[j * ldn + k] = j < amax ? sums[amax * lds + k] : C_new[amax * ldn + k];
  }
  double tot = 0.0;
  for (int64_t j = 0; j < K; ++j)
    for (int64_t k = 0; k < D; ++k) {
      double d = (double)C_new[j * ldn + k] - (double)C_old[j * ldo + k];
      tot += d * d;
    }
  return tot;
}

/* _inertia_dense   sklearn/_k_means_common.pyx:94-124 — per-sample squared distance in
 * fp32 (sklearn's 4-way unrolled expression), summed here in double. */
double oracle_inertia(int64_t N, int64_t D, const float* X, int64_t ldx, const float* C, int64_t ldc,
                      const int32_t* labels) {
  double tot = 0.0;
  for (int64_t i = 0; i < N; ++i) {
    const float* a = X + i * ldx;
    const float* b = C + (int64_t)labels[i] * ldc;
    float r = 0.0f;
    int64_t n4 = D / 4, k = 0;
    for (int64_t q = 0; q < n4; ++q, k += 4)
      r = r + ((a[k] - b[k]) * (a[k] - b[k]) + (a[k + 1] - b[k + 1]) * (a[k + 1] - b[k + 1]) +
               (a[k + 2] - b[k + 2]) * (a[k + 2] - b[k + 2]) + (a[k + 3] - b[k + 3]) * (a[k + 3] - b[k + 3]));
    for (; k < D; ++k) r = r + (a[k] - b[k]) * (a[k] - b[k]);
    tot += (double)r;
  }
  return tot;
}

/* exact WCSS in double (the 1e-4 relative contract is checked against this as well) */
double oracle_inertia_f64(int64_t N, int64_t D, const float* X, int64_t ldx, const float* C,
                          int64_t ldc, const int32_t* labels) {
  double tot = 0.0;
  for (int64_t i = 0; i < N; ++i) {
    const float* a = X + i * ldx;
    const float* b = C + (int64_t)labels[i] * ldc;
    for (int64_t k = 0; k < D; ++k) {
      double d = (double)a[k] - (double)b[k];
      tot += d * d;
    }
  }
  return tot;
}
