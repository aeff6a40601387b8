// spmm.cu — stage 2: one propagation hop  Y = (alpha*A) @ X ; T += beta*Y  as a
// CSR SpMM for sm_100a.
//
// Replaces the reference expression `alpha*adj_norm @ prop_feat` + the axpy on
// the next line (clustgdd_agent_transduct.py:64-65, clustgdd_agent_induct.py
// :77-94) which the reference runs as ATen mul(sparse,scalar) + coalesce +
// cuSPARSE/MKL SpMM, and the index_add_ message passing of
// distill_recsys.py:340-345.
//
// Mapping (HBM/L2-gather bound, no tensor cores):
//   * a CTA owns a block of SPMM_ROWS_PER_CTA consecutive rows; their rowptr
//     slice is staged in shared memory once, and the CTA's warps pull rows from
//     it through a shared counter (dynamic balance inside the block);
//   * a row is processed by a group of LPR lanes (8/16/32, chosen from F) so a
//     gathered X row is one fully coalesced run of 128-bit loads;
//   * the row's (colidx, val) pairs are read coalesced, LPR at a time, and
//     broadcast with warp shuffles; 4 gathers are in flight per lane before
//     the first FMA consumes one (MLP), accumulation order stays CSR order so
//     the result is deterministic;
//   * rows longer than HEAVY_NNZ are split across the CTA's warps and reduced
//     through shared memory in fixed order.
#include "common.cuh"

namespace gdr {

constexpr int SPMM_THREADS = 256;
constexpr int SPMM_WARPS = SPMM_THREADS / 32;
constexpr int SPMM_ROWS_PER_CTA = 64;

template <int LPR, int NCH>
__device__ __forceinline__ void spmm_row(const int32_t* __restrict__ colidx,
                                         const float* __restrict__ vals, float alpha,
                                         const float* __restrict__ X, int64_t ldx, int start, int end,
                                         int F4, int gl /*lane in group*/, unsigned gmask,
                                         int col4_base, float4 (&acc)[NCH]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  // every lane of the WARP must execute the same number of shuffle rounds:
  // the caller passes a warp-uniform trip count through start/end of the
  // longest row in the warp when LPR < 32 (see below); here start/end are
  // per-group and inactive iterations are predicated.
  for (int k = start; k < end; k += LPR) {
    int my = k + gl;
    int c_l = 0;
    float v_l = 0.f;
    if (my < end) {
      c_l = __ldg(colidx + my);
      v_l = vals ? __fmul_rn(__ldg(vals + my), alpha) : alpha;
    }
    int cnt = min(LPR, end - k);
#pragma unroll 1
    for (int j = 0; j < cnt; j += 4) {
      int cj[4];
      float vj[4];
      float4 xv[4][NCH];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cj[u] = __shfl_sync(gmask, c_l, (j + u) & (LPR - 1), LPR);
        vj[u] = __shfl_sync(gmask, v_l, (j + u) & (LPR - 1), LPR);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u < cnt) {
          const float* xr = X + (int64_t)cj[u] * ldx;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            int c4 = col4_base + c * LPR + gl;
            if (c4 < F4) xv[u][c] = ldg_f4(xr + 4 * c4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u < cnt) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            int c4 = col4_base + c * LPR + gl;
            if (c4 < F4) {
              acc[c].x = fmaf(vj[u], xv[u][c].x, acc[c].x);
              acc[c].y = fmaf(vj[u], xv[u][c].y, acc[c].y);
              acc[c].z = fmaf(vj[u], xv[u][c].z, acc[c].z);
              acc[c].w = fmaf(vj[u], xv[u][c].w, acc[c].w);
            }
          }
        }
      }
    }
  }
}

template <int LPR, int NCH>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm(int64_t rows, int F4, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
       const float* __restrict__ vals, float alpha, const float* __restrict__ X, int64_t ldx,
       float* __restrict__ Y, int64_t ldy, float* __restrict__ T, int64_t ldt, float beta,
       int col4_base) {
  __shared__ int s_rowptr[SPMM_ROWS_PER_CTA + 1];
  __shared__ int s_next;
  constexpr int GROUPS = 32 / LPR;  // rows processed concurrently by one warp
  const int64_t row0 = (int64_t)blockIdx.x * SPMM_ROWS_PER_CTA;
  const int nrows = (int)min((int64_t)SPMM_ROWS_PER_CTA, rows - row0);
  for (int i = threadIdx.x; i <= nrows; i += SPMM_THREADS) s_rowptr[i] = rowptr[row0 + i];
  if (threadIdx.x == 0) s_next = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int g = lane / LPR, gl = lane % LPR;
  const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << (g * LPR));
  while (true) {
    int r_base = 0;
    if (lane == 0) r_base = atomicAdd(&s_next, GROUPS);
    r_base = __shfl_sync(0xffffffffu, r_base, 0);
    if (r_base >= nrows) break;
    int r = r_base + g;
    int start = 0, end = 0;
    if (r < nrows) {
      start = s_rowptr[r];
      end = s_rowptr[r + 1];
    }
    float4 acc[NCH];
    spmm_row<LPR, NCH>(colidx, vals, alpha, X, ldx, start, end, F4, gl, gmask, col4_base, acc);
    if (r < nrows) {
      int64_t row = row0 + r;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        int c4 = col4_base + c * LPR + gl;
        if (c4 < F4) {
          *reinterpret_cast<float4*>(Y + row * ldy + 4 * c4) = acc[c];
          if (T) {
            float4* tp = reinterpret_cast<float4*>(T + row * ldt + 4 * c4);
            float4 t = *tp;
            t.x = __fadd_rn(t.x, __fmul_rn(beta, acc[c].x));
            t.y = __fadd_rn(t.y, __fmul_rn(beta, acc[c].y));
            t.z = __fadd_rn(t.z, __fmul_rn(beta, acc[c].z));
            t.w = __fadd_rn(t.w, __fmul_rn(beta, acc[c].w));
            *tp = t;
          }
        }
      }
    }
  }
}

template <int LPR, int NCH>
static int launch_spmm(int64_t rows, int F4, const int32_t* rowptr, const int32_t* colidx,
                       const float* vals, float alpha, const float* X, int64_t ldx, float* Y,
                       int64_t ldy, float* T, int64_t ldt, float beta, int col4_base,
                       cudaStream_t s) {
  unsigned grid = (unsigned)cdiv(rows, SPMM_ROWS_PER_CTA);
  {
    ProfileScope prof(PROF_SPMM, s);
    k_spmm<LPR, NCH><<<grid, SPMM_THREADS, 0, s>>>(rows, F4, rowptr, colidx, vals, alpha, X, ldx, Y,
                                                   ldy, T, ldt, beta, col4_base);
  }
  GDR_LAUNCHED();
  return GDR_OK;
}

int spmm_launch(int64_t rows, int64_t F, const int32_t* rowptr, const int32_t* colidx,
                const float* vals, float alpha, const float* X, int64_t ldx, float* Y, int64_t ldy,
                float* T, int64_t ldt, float beta, cudaStream_t s) {
  if (rows == 0 || F == 0) return GDR_OK;
  int F4 = (int)cdiv(F, 4);
#define GDR_SPMM(LPR, NCH, base) \
  launch_spmm<LPR, NCH>(rows, F4, rowptr, colidx, vals, alpha, X, ldx, Y, ldy, T, ldt, beta, base, s)
  if (F4 <= 8) return GDR_SPMM(8, 1, 0);
  if (F4 <= 16) return GDR_SPMM(16, 1, 0);
  if (F4 <= 32) return GDR_SPMM(32, 1, 0);
  if (F4 <= 64) return GDR_SPMM(32, 2, 0);
  if (F4 <= 128) return GDR_SPMM(32, 4, 0);
  // wide rows: column super-blocks of 8*32 float4 = 1024 floats
  for (int base = 0; base < F4; base += 256) {
    int rc = GDR_SPMM(32, 8, base);
    if (rc) return rc;
  }
#undef GDR_SPMM
  return GDR_OK;
}

__global__ void k_scale_rows(int64_t rows, int F4, float a, const float* __restrict__ X, int64_t ldx,
                             float* __restrict__ out, int64_t ldo) {
  int64_t total = rows * (int64_t)F4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / F4;
    int c4 = (int)(i - r * F4);
    float4 x = ldg_nc_f4(X + r * ldx + 4 * c4);
    x.x = __fmul_rn(a, x.x);
    x.y = __fmul_rn(a, x.y);
    x.z = __fmul_rn(a, x.z);
    x.w = __fmul_rn(a, x.w);
    *reinterpret_cast<float4*>(out + r * ldo + 4 * c4) = x;
  }
}

}  // namespace gdr

extern "C" {

int gdr_spmm_prop(int64_t rows_local, int64_t F, const int32_t* rowptr, const int32_t* colidx,
                  const float* vals, float alpha, const float* X, int64_t ldx, float* Y, int64_t ldy,
                  float* T, int64_t ldt, float beta, gdr_stream_t stream) {
  GDR_CHECK_ARG(rows_local >= 0 && F >= 0, "spmm_prop: negative size");
  if (rows_local == 0 || F == 0) return GDR_OK;
  GDR_CHECK_ARG(rowptr && colidx && X && Y, "spmm_prop: null pointer");
  GDR_CHECK_ARG(ldx % 4 == 0 && ldy % 4 == 0 && ldx >= gdr::align_up(F, 4) &&
                    ldy >= gdr::align_up(F, 4),
                "spmm_prop: ldx/ldy must be multiples of 4 and >= F rounded up to 4");
  GDR_CHECK_ARG(((uintptr_t)X & 15) == 0 && ((uintptr_t)Y & 15) == 0, "spmm_prop: X/Y not 16B aligned");
  if (T) {
    GDR_CHECK_ARG(ldt % 4 == 0 && ldt >= gdr::align_up(F, 4) && ((uintptr_t)T & 15) == 0,
                  "spmm_prop: T misaligned");
  }
  GDR_CHECK_ARG(X != Y, "spmm_prop: in-place propagation is not supported");
  return gdr::spmm_launch(rows_local, F, rowptr, colidx, vals, alpha, X, ldx, Y, ldy, T, ldt, beta,
                          (cudaStream_t)stream);
}

int gdr_scale_rows(int64_t rows, int64_t F, float a, const float* X, int64_t ldx, float* out,
                   int64_t ldo, gdr_stream_t stream) {
  GDR_CHECK_ARG(rows >= 0 && F >= 0, "scale_rows: negative size");
  if (rows == 0 || F == 0) return GDR_OK;
  GDR_CHECK_ARG(X && out && ldx % 4 == 0 && ldo % 4 == 0 && ((uintptr_t)X & 15) == 0 &&
                    ((uintptr_t)out & 15) == 0,
                "scale_rows: null or misaligned");
  int F4 = (int)gdr::cdiv(F, 4);
  int64_t total = rows * F4;
  unsigned grid = (unsigned)std::min<int64_t>(gdr::cdiv(total, 256), gdr::kSMs * 16);
  gdr::k_scale_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(rows, F4, a, X, ldx, out, ldo);
  GDR_LAUNCHED();
  return GDR_OK;
}

}  // extern "C"
