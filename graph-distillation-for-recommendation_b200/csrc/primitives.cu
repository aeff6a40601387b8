// primitives.cu — device-wide exclusive scan (int32) and a stable LSD radix
// sort of (u64 key, u32 payload) pairs.  Both are HBM-streaming integer work:
// coalesced tile loads, shared-memory staging, warp ballots — no tensor cores.
//
// They carry stage 1 (COO -> CSR with duplicate summation: what scipy's
// coo_tocsr + csr_sum_duplicates do on the host for utils.py:66-67 and
// distill_recsys.py:116-117), the k-means M-step (cluster membership lists) and
// stage 4 (segmented edge counting, distill_recsys.py:196-201).
#include "common.cuh"

namespace gdr {

// ----------------------------------------------------------------------------
// exclusive scan
// ----------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one int per thread; returns exclusive prefix and
// the block total through *total.
__device__ __forceinline__ int block_excl_scan(int v, int* smem_warp /*[32]*/, int* total) {
  int incl = warp_incl_scan(v);
  int w = threadIdx.x >> 5;
  if (lane_id() == 31) smem_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    int x = lane_id() < nw ? smem_warp[lane_id()] : 0;
    int xi = warp_incl_scan(x);
    smem_warp[lane_id()] = xi - x;  // exclusive warp offsets
    if (lane_id() == 31) smem_warp[32] = xi;
  }
  __syncthreads();
  int r = smem_warp[w] + incl - v;
  *total = smem_warp[32];
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int32_t* __restrict__ in,
                                                              int32_t* __restrict__ block_sums,
                                                              int64_t n) {
  __shared__ int sw[33];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    if (idx < n) s += in[idx];
  }
  int total;
  block_excl_scan(s, sw, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int32_t* __restrict__ in,
                                                             const int32_t* __restrict__ block_offs,
                                                             int32_t* __restrict__ out, int64_t n,
                                                             int write_total) {
  __shared__ int sw[33];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  int total;
  int excl = block_excl_scan(s, sw, &total);
  int run = excl + (block_offs ? block_offs[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = run;
    run += v[i];
    if (write_total && idx == n - 1) out[n] = run;
  }
}

// one-launch scan for small inputs (n <= SCAN_SMALL_MAX): a single CTA, each thread owns a
// contiguous run of ceil(n / 1024) elements; out may alias in.
constexpr int SCAN_SMALL_MAX = 1 << 17;
__global__ void __launch_bounds__(1024) k_scan_small(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                     int n, int write_total) {
  __shared__ int sw[33];
  const int per = (n + 1023) / 1024;
  const int b = threadIdx.x * per, e = min(n, b + per);
  int s = 0;
  for (int i = b; i < e; ++i) s += in[i];
  int total;
  int run = block_excl_scan(s, sw, &total);
  for (int i = b; i < e; ++i) {
    int v = in[i];
    out[i] = run;
    run += v;
  }
  if (write_total && threadIdx.x == 0) out[n] = total;
}

__global__ void k_zero_one(int32_t* out) { out[0] = 0; }

int64_t scan_ws_bytes(int64_t n) {
  int64_t total = 0;
  while (n > SCAN_TILE) {
    int64_t nb = cdiv(n, SCAN_TILE);
    total += ws_need(nb + 1, 4);
    n = nb;
  }
  return total + 256;
}

// out may alias in.  out has n+1 entries when write_total != 0.
static int scan_rec(const int32_t* in, int32_t* out, int64_t n, char* ws, int write_total,
                    cudaStream_t s) {
  if (n <= 0) {
    if (write_total) {
      k_zero_one<<<1, 1, 0, s>>>(out);
      GDR_LAUNCHED();
    }
    return GDR_OK;
  }
  int64_t nb = cdiv(n, SCAN_TILE);
  // (measured: the single-CTA scan only wins below ~3 tiles; the radix tables of the k-means
  //  membership sort, 42 K entries, take 35 us this way vs 3 x 3.8 us through the tiled path)
  if (n <= SCAN_SMALL_MAX && nb > 1 && nb <= 3) {
    k_scan_small<<<1, 1024, 0, s>>>(in, out, (int)n, write_total);
    GDR_LAUNCHED();
    return GDR_OK;
  }
  if (nb == 1) {
    k_scan_apply<<<1, SCAN_THREADS, 0, s>>>(in, nullptr, out, n, write_total);
    GDR_LAUNCHED();
    return GDR_OK;
  }
  int32_t* bs = (int32_t*)ws;
  char* ws_next = ws + ws_need(nb + 1, 4);
  k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, bs, n);
  GDR_LAUNCHED();
  int rc = scan_rec(bs, bs, nb, ws_next, 0, s);
  if (rc) return rc;
  k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, bs, out, n, write_total);
  GDR_LAUNCHED();
  return GDR_OK;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, int64_t ws_bytes,
                       cudaStream_t s) {
  if (ws_bytes < scan_ws_bytes(n)) {
    set_error("exclusive_scan: workspace %lld < %lld", (long long)ws_bytes,
              (long long)scan_ws_bytes(n));
    return GDR_EWORKSPACE;
  }
  return scan_rec(in, out, n, (char*)ws, 1, s);
}

// ----------------------------------------------------------------------------
// stable LSD radix sort, digits of up to 11 bits
// ----------------------------------------------------------------------------
// The pass count is what the sort costs (every pass streams the keys and payloads through HBM once in, once
// out), so the key bits are cut into as few digits as fit: ceil(bits / 11) passes of equal width — 10-bit
// cluster labels (K = 1000) sort in ONE pass, the 36-bit (row, col) keys of an arxiv-sized graph in 4 instead
// of 5, the 44-bit keys of a products-sized graph in 4 instead of 6.  Bins live in dynamic shared memory
// (8 warp-private counter rows of 2^bits ints: 64 KB at 11 bits).
int g_rs_match = 0;   // gdr_debug_set("rs_match", 1): rank with MATCH.ANY instead of the per-bit ballots (experiment)
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_BITS = 11;
constexpr int RS_MAX_BINS = 1 << RS_MAX_BITS;
// keys per thread: 16 (4096-key tiles) for large inputs, 4 (1024-key tiles) below 2M keys so
// that mid-size sorts (k-means membership lists, arxiv-sized graphs) still fill the 148 SMs
static inline int rs_rounds(int64_t n) { return n >= (1ll << 21) ? 16 : 4; }
static inline int rs_passes(int key_bits, int max_bits) { return (key_bits + max_bits - 1) / max_bits; }
static inline int rs_digit_bits(int key_bits, int max_bits) {
  return (key_bits + rs_passes(key_bits, max_bits) - 1) / rs_passes(key_bits, max_bits);
}

template <int RS_ROUNDS>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, int64_t n,
                                                        int shift, int bits, int32_t* __restrict__ table,
                                                        int nblocks) {
  extern __shared__ int rs_smem[];
  int* hist = rs_smem;
  const int bins = 1 << bits;
  const uint32_t dmask = (uint32_t)bins - 1u;
  for (int i = threadIdx.x; i < bins; i += RS_THREADS) hist[i] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * (RS_THREADS * RS_ROUNDS);
#pragma unroll 4
  for (int r = 0; r < RS_ROUNDS; ++r) {
    int64_t idx = base + r * RS_THREADS + threadIdx.x;
    if (idx < n) {
      int d = (int)((uint32_t)(keys[idx] >> shift) & dmask);
      atomicAdd(&hist[d], 1);
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < bins; d += RS_THREADS) table[(int64_t)d * nblocks + blockIdx.x] = hist[d];
}

// Scatter pass.  Every key gets its rank inside the tile (warp match + warp-private counters, so the sort is
// stable); the tile is then REORDERED IN SHARED MEMORY into digit order and written out with consecutive threads on
// consecutive addresses of each digit's run.  Writing straight from registers sends every 8-byte key to its own
// 32-byte sector (the runs of one tile in one bin are short), which cost 4.5 ms per pass at 124 M keys = 0.66 TB/s;
// staged, a run of r keys is r * 8 contiguous bytes.
//
// The ranking is written for latency, not instruction count (ncu, round 2: 60 % of the stall samples sat on the
// consumer of MATCH.ANY, one round at a time, at 25 % occupancy):
//   phase A  all ROUNDS match.any are issued back to back (independent);
//   phase B  the lowest lane of every peer group adds the group size to the warp's counter of that digit with a
//            shared-memory ATOMIC that returns the old value — no load / store / __syncwarp chain between rounds, the
//            atomics of one warp reach the shared-memory pipe in program order, so the returned bases are the
//            in-order prefix;
//   phase C  the bases are broadcast with shuffles.
// FULL = the tile holds TILE keys (no bounds checks on the hot path); the last, partial tile runs the <.., false> variant.
// BITS = digit width, a compile-time constant (the per-bit ballot loop and the per-digit scans unroll exactly; with a
// run-time width the compiler emitted all RS_MAX_BITS + 1 predicated votes: 96 instructions per key instead of 27).
template <int RS_ROUNDS, bool HAS_VALS, bool FULL, int BITS>
__global__ void __launch_bounds__(RS_THREADS, HAS_VALS ? 2 : 3)
k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
             uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
             const int32_t* __restrict__ table_scanned, int nblocks, int block0, int use_match) {
  extern __shared__ __align__(16) unsigned char rs_raw[];
  constexpr int TILE = RS_THREADS * RS_ROUNDS;
  constexpr int bits = BITS;
  constexpr int bins = 1 << BITS;
  constexpr uint32_t dmask = (uint32_t)bins - 1u;
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(rs_raw);                                    // [TILE]
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + TILE);                             // [TILE] (HAS_VALS)
  int* cnt = reinterpret_cast<int*>(s_vals + (HAS_VALS ? TILE : 0));                         // [RS_WARPS][bins + 1]
  int* tile_off = cnt + RS_WARPS * (bins + 1);                                               // [bins] first tile position of digit d
  int* delta = tile_off + bins;                                                              // [bins] global position - tile position
  __shared__ int s_wsum[RS_WARPS];
  constexpr int cstride = bins + 1;                   // + 1: the sentinel digit of padding lanes (partial tile)
  for (int i = threadIdx.x; i < RS_WARPS * cstride; i += RS_THREADS) cnt[i] = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  int* mycnt = cnt + w * cstride;
  const int blk = block0 + blockIdx.x;
  const int64_t tile0 = (int64_t)blk * TILE;
  const int tile_n = FULL ? TILE : (int)min((int64_t)TILE, n - tile0);
  const int64_t base = tile0 + (int64_t)w * (RS_ROUNDS * 32) + lane;
  uint64_t k[RS_ROUNDS];
  uint32_t v[HAS_VALS ? RS_ROUNDS : 1];
  unsigned pr[RS_ROUNDS];                             // peers mask, then the rank inside the warp's keys of that digit
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const bool valid = FULL || base + r * 32 < n;
    k[r] = valid ? keys_in[base + r * 32] : 0ull;
    if (HAS_VALS) v[r] = valid ? vals_in[base + r * 32] : 0u;
  }
  // phase A: the lanes holding the same digit — one ballot per digit bit (bits + 1 with the padding sentinel), each a
  // single-issue warp vote, instead of MATCH.ANY, whose cost grows with the number of distinct values in the warp
  // (~31 of 32 at 9-bit digits: measured ~17 us per 4096-key tile, an IPC of 0.1)
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const bool valid = FULL || base + r * 32 < n;
    const unsigned d = valid ? ((uint32_t)(k[r] >> shift) & dmask) : (unsigned)bins;
    unsigned peers = 0xffffffffu;
    if (use_match) {
      peers = __match_any_sync(0xffffffffu, d);
    } else {
#pragma unroll
      for (int b = 0; b < BITS + (FULL ? 0 : 1); ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
      }
    }
    pr[r] = peers;
  }
  // phase B: group leader (lowest lane) reserves the group's slots; the round's state is packed into one register:
  // [base of the group << 10 | leader lane << 5 | peers before me]
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const bool valid = FULL || base + r * 32 < n;
    const int d = valid ? (int)((uint32_t)(k[r] >> shift) & dmask) : bins;
    const unsigned peers = pr[r];
    const unsigned before = __popc(peers & lt_mask);
    unsigned b = 0;
    if (before == 0u) b = (unsigned)atomicAdd(&mycnt[d], __popc(peers));
    pr[r] = (b << 10) | ((unsigned)(__ffs(peers) - 1) << 5) | before;
  }
  // phase C
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const unsigned b = __shfl_sync(0xffffffffu, pr[r] >> 10, (int)((pr[r] >> 5) & 31u));
    pr[r] = b + (pr[r] & 31u);
  }
  __syncthreads();
  // per digit: counts of the warps -> exclusive offsets; tile-wide exclusive scan of the digit totals
  constexpr int per = (bins + RS_THREADS - 1) / RS_THREADS;      // digits owned by this thread (contiguous)
  const int d0 = threadIdx.x * per, d1 = min(bins, d0 + per);
  int mine = 0;
#pragma unroll
  for (int q = 0; q < per; ++q) {
    const int d = d0 + q;
    if (d < d1) {
      int run = 0;
#pragma unroll
      for (int ww = 0; ww < RS_WARPS; ++ww) {
        int t = cnt[ww * cstride + d];
        cnt[ww * cstride + d] = run;
        run += t;
      }
      tile_off[d] = run;          // digit total for now
      mine += run;
    }
  }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_wsum[w] = incl;
  __syncthreads();
  int woff = 0;
#pragma unroll
  for (int ww = 0; ww < RS_WARPS; ++ww) woff += ww < w ? s_wsum[ww] : 0;
  int run = woff + incl - mine;
  for (int d = d0; d < d1; ++d) {
    const int t = tile_off[d];
    tile_off[d] = run;
    delta[d] = table_scanned[(int64_t)d * nblocks + blk] - run;
    run += t;
  }
  __syncthreads();
  // stage the tile in digit order
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    if (FULL || base + r * 32 < n) {
      const int d = (int)((uint32_t)(k[r] >> shift) & dmask);
      const int p = tile_off[d] + mycnt[d] + (int)pr[r];
      s_keys[p] = k[r];
      if (HAS_VALS) s_vals[p] = v[r];
    }
  }
  __syncthreads();
  // consecutive threads -> consecutive positions of a digit's run
#pragma unroll 4
  for (int t = threadIdx.x; t < tile_n; t += RS_THREADS) {
    const uint64_t kk = s_keys[t];
    const int d = (int)((uint32_t)(kk >> shift) & dmask);
    const int64_t pos = (int64_t)(delta[d] + t);
    keys_out[pos] = kk;
    if (HAS_VALS) vals_out[pos] = s_vals[t];
  }
}

static inline size_t rs_scatter_smem(int rounds, int bits, bool has_vals) {
  return (size_t)RS_THREADS * rounds * (has_vals ? 12 : 8) + (size_t)RS_WARPS * ((1 << bits) + 1) * 4 + (size_t)2 * (1 << bits) * 4;
}

template <int ROUNDS, bool HAS_VALS, int BITS>
static int launch_rs_scatter_b(const uint64_t* kin, const uint32_t* vin, uint64_t* kout, uint32_t* vout, int64_t n, int shift,
                               const int32_t* table, int64_t nb, cudaStream_t s) {
  static PerDevice<bool> attr_set_dev;
  bool& attr_set = attr_set_dev.get();
  const int smem = (int)rs_scatter_smem(ROUNDS, BITS, HAS_VALS);
  if (!attr_set) {
    GDR_CUDA(cudaFuncSetAttribute(k_rs_scatter<ROUNDS, HAS_VALS, true, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    GDR_CUDA(cudaFuncSetAttribute(k_rs_scatter<ROUNDS, HAS_VALS, false, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int64_t tile = (int64_t)RS_THREADS * ROUNDS;
  const int64_t full = n / tile;
  if (full > 0) {
    k_rs_scatter<ROUNDS, HAS_VALS, true, BITS><<<(unsigned)full, RS_THREADS, smem, s>>>(kin, vin, kout, vout, n, shift, table,
                                                                                     (int)nb, 0, g_rs_match);
    GDR_LAUNCHED();
  }
  if (full < nb) {
    k_rs_scatter<ROUNDS, HAS_VALS, false, BITS><<<(unsigned)(nb - full), RS_THREADS, smem, s>>>(kin, vin, kout, vout, n, shift,
                                                                                             table, (int)nb, (int)full,
                                                                                             g_rs_match);
    GDR_LAUNCHED();
  }
  return GDR_OK;
}

template <int ROUNDS, bool HAS_VALS>
static int launch_rs_scatter(const uint64_t* kin, const uint32_t* vin, uint64_t* kout, uint32_t* vout, int64_t n, int shift,
                             int bits, const int32_t* table, int64_t nb, cudaStream_t s) {
  switch (bits) {
    case 7: return launch_rs_scatter_b<ROUNDS, HAS_VALS, 7>(kin, vin, kout, vout, n, shift, table, nb, s);
    case 8: return launch_rs_scatter_b<ROUNDS, HAS_VALS, 8>(kin, vin, kout, vout, n, shift, table, nb, s);
    case 9: return launch_rs_scatter_b<ROUNDS, HAS_VALS, 9>(kin, vin, kout, vout, n, shift, table, nb, s);
    case 10: return launch_rs_scatter_b<ROUNDS, HAS_VALS, 10>(kin, vin, kout, vout, n, shift, table, nb, s);
    case 11: return launch_rs_scatter_b<ROUNDS, HAS_VALS, 11>(kin, vin, kout, vout, n, shift, table, nb, s);
  }
  set_error("sort_pairs: digit width %d is not one of the built widths (7..11)", bits);
  return GDR_EUNSUPPORTED;
}
// digit width: wide digits save passes, but a tile of T keys leaves runs of T / 2^bits keys per digit, and a run
// is what one coalesced write covers: at most 9 bits for the 4096-key tiles of large sorts (runs of >= 8 keys =
// two full 32-byte sectors), up to 11 for small ones (the whole output stays in L2)
int g_rs_max_bits = 0;   // gdr_debug_set("rs_max_bits", 9 | 10 | 11): digit width cap of large sorts (experiment)
static inline int rs_max_bits(int64_t n) {
  if (n < (1ll << 21)) return RS_MAX_BITS;
  return (g_rs_max_bits >= 7 && g_rs_max_bits <= RS_MAX_BITS) ? g_rs_max_bits : 9;
}

int64_t sort_pairs_ws_bytes(int64_t n) {
  if (n <= 0) return 256;
  int64_t nb = cdiv(n, RS_THREADS * rs_rounds(n));
  int64_t tbl = (int64_t)RS_MAX_BINS * nb;
  return ws_need(n, 8) + ws_need(n, 4) + ws_need(tbl + 1, 4) + scan_ws_bytes(tbl) + 256;
}

int sort_pairs(int64_t n, int key_bits, uint64_t* keys, uint32_t* vals, void* ws, int64_t ws_bytes,
               cudaStream_t s) {
  return sort_pairs_ex(n, key_bits, keys, vals, ws, ws_bytes, nullptr, nullptr, s);
}

// keys_sorted / vals_sorted != nullptr: the caller accepts the result wherever the last pass left it (the input
// arrays or their twins inside the workspace) and the copy back after an odd number of passes (n * 12 bytes
// read + written) is skipped.
int sort_pairs_ex(int64_t n, int key_bits, uint64_t* keys, uint32_t* vals, void* ws, int64_t ws_bytes,
                  uint64_t** keys_sorted, uint32_t** vals_sorted, cudaStream_t s) {
  if (keys_sorted) *keys_sorted = keys;
  if (vals_sorted) *vals_sorted = vals;
  if (n <= 1 || key_bits <= 0) return GDR_OK;
  if (ws_bytes < sort_pairs_ws_bytes(n)) {
    set_error("sort_pairs: workspace %lld < %lld", (long long)ws_bytes,
              (long long)sort_pairs_ws_bytes(n));
    return GDR_EWORKSPACE;
  }
  if (n >= (1ll << 31)) {
    set_error("sort_pairs: n=%lld exceeds int32 positions", (long long)n);
    return GDR_ERANGE;
  }
  const int rounds = rs_rounds(n);
  int64_t nb = cdiv(n, RS_THREADS * rounds);
  const int passes = rs_passes(key_bits, rs_max_bits(n));
  // the scatter kernel is built for digit widths 7..11; a narrower plan runs as 7 bits (its digits then overlap the
  // next pass's, which sorts those bits again: still a stable LSD sort).  A digit never reaches above bit key_bits - 1 as
  // long as key_bits >= 7 (the last pass is shifted DOWN onto already sorted bits instead), so callers may keep unrelated
  // bits above the sorted field; below 7 bits the bits above the field must be zero.
  const int bits = std::max(7, rs_digit_bits(key_bits, rs_max_bits(n)));
  const int step = rs_digit_bits(key_bits, rs_max_bits(n));
  const int bins = 1 << bits;
  int64_t tbl = (int64_t)bins * nb;
  Workspace W(ws, ws_bytes);
  uint64_t* kalt = W.take<uint64_t>(n);
  uint32_t* valt = W.take<uint32_t>(n);
  int32_t* table = W.take<int32_t>((int64_t)RS_MAX_BINS * nb + 1);
  void* sws = W.take<char>(scan_ws_bytes((int64_t)RS_MAX_BINS * nb));
  uint64_t* kin = keys;
  uint32_t* vin = vals;
  uint64_t* kout = kalt;
  uint32_t* vout = vals ? valt : nullptr;
  const size_t hist_smem = (size_t)bins * 4;
  for (int p = 0; p < passes; ++p) {
    int shift = std::max(0, std::min(step * p, key_bits - bits));
    if (rounds == 16) k_rs_hist<16><<<(unsigned)nb, RS_THREADS, hist_smem, s>>>(kin, n, shift, bits, table, (int)nb);
    else k_rs_hist<4><<<(unsigned)nb, RS_THREADS, hist_smem, s>>>(kin, n, shift, bits, table, (int)nb);
    GDR_LAUNCHED();
    int rc = exclusive_scan_i32(table, table, tbl, sws, scan_ws_bytes(tbl), s);
    if (rc) return rc;
    if (rounds == 16)
      rc = vals ? launch_rs_scatter<16, true>(kin, vin, kout, vout, n, shift, bits, table, nb, s)
                : launch_rs_scatter<16, false>(kin, vin, kout, vout, n, shift, bits, table, nb, s);
    else
      rc = vals ? launch_rs_scatter<4, true>(kin, vin, kout, vout, n, shift, bits, table, nb, s)
                : launch_rs_scatter<4, false>(kin, vin, kout, vout, n, shift, bits, table, nb, s);
    if (rc) return rc;
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  if (kin != keys) {
    if (keys_sorted) {
      *keys_sorted = kin;
      if (vals_sorted) *vals_sorted = vin;
      else if (vals) GDR_CUDA(cudaMemcpyAsync(vals, vin, n * 4, cudaMemcpyDeviceToDevice, s));
    } else {
      GDR_CUDA(cudaMemcpyAsync(keys, kin, n * 8, cudaMemcpyDeviceToDevice, s));
      if (vals) GDR_CUDA(cudaMemcpyAsync(vals, vin, n * 4, cudaMemcpyDeviceToDevice, s));
    }
  }
  return GDR_OK;
}

// first position of every digit value in the output of a pass: starts[d] = table_scanned[d * nblocks], starts[bins] = n
__global__ void k_digit_starts(int bins, int nblocks, int64_t n, const int32_t* __restrict__ table_scanned,
                               int64_t* __restrict__ starts) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < bins) starts[d] = table_scanned[(int64_t)d * nblocks];
  if (d == bins) starts[d] = n;
}

// ONE stable pass on the `bits`-wide digit at `shift` (bits in 7..11): a stable partition into 2^bits buckets.
// The result is left in the twin arrays of the workspace (*keys_sorted / *vals_sorted); starts_dev (nullable,
// int64[2^bits + 1]) receives the first position of every bucket.
int sort_pairs_digit(int64_t n, int shift, int bits, uint64_t* keys, uint32_t* vals, void* ws, int64_t ws_bytes,
                     uint64_t** keys_sorted, uint32_t** vals_sorted, int64_t* starts_dev, cudaStream_t s,
                     uint64_t* keys_dst, uint32_t* vals_dst) {
  if (bits < 7 || bits > RS_MAX_BITS || !keys_sorted || (vals && !vals_sorted)) {
    set_error("sort_pairs_digit: bad arguments");
    return GDR_EINVAL;
  }
  if (ws_bytes < sort_pairs_ws_bytes(n) || n >= (1ll << 31)) {
    set_error("sort_pairs_digit: workspace too small or n out of range");
    return GDR_EWORKSPACE;
  }
  const int bins = 1 << bits;
  if (n == 0) {
    *keys_sorted = keys;
    if (vals_sorted) *vals_sorted = vals;
    if (starts_dev) GDR_CUDA(cudaMemsetAsync(starts_dev, 0, (bins + 1) * 8, s));
    return GDR_OK;
  }
  const int rounds = rs_rounds(n);
  const int64_t nb = cdiv(n, RS_THREADS * rounds);
  const int64_t tbl = (int64_t)bins * nb;
  Workspace W(ws, ws_bytes);
  uint64_t* kalt = W.take<uint64_t>(n);
  uint32_t* valt = W.take<uint32_t>(n);
  if (keys_dst) kalt = keys_dst;      // explicit destination instead of the workspace twins
  if (vals_dst) valt = vals_dst;
  int32_t* table = W.take<int32_t>((int64_t)RS_MAX_BINS * nb + 1);
  void* sws = W.take<char>(scan_ws_bytes((int64_t)RS_MAX_BINS * nb));
  const size_t hist_smem = (size_t)bins * 4;
  if (rounds == 16) k_rs_hist<16><<<(unsigned)nb, RS_THREADS, hist_smem, s>>>(keys, n, shift, bits, table, (int)nb);
  else k_rs_hist<4><<<(unsigned)nb, RS_THREADS, hist_smem, s>>>(keys, n, shift, bits, table, (int)nb);
  GDR_LAUNCHED();
  int rc = exclusive_scan_i32(table, table, tbl, sws, scan_ws_bytes(tbl), s);
  if (rc) return rc;
  if (rounds == 16)
    rc = vals ? launch_rs_scatter<16, true>(keys, vals, kalt, valt, n, shift, bits, table, nb, s)
              : launch_rs_scatter<16, false>(keys, vals, kalt, nullptr, n, shift, bits, table, nb, s);
  else
    rc = vals ? launch_rs_scatter<4, true>(keys, vals, kalt, valt, n, shift, bits, table, nb, s)
              : launch_rs_scatter<4, false>(keys, vals, kalt, nullptr, n, shift, bits, table, nb, s);
  if (rc) return rc;
  if (starts_dev) {
    k_digit_starts<<<(unsigned)cdiv(bins + 1, 256), 256, 0, s>>>(bins, (int)nb, n, table, starts_dev);
    GDR_LAUNCHED();
  }
  *keys_sorted = kalt;
  if (vals_sorted) *vals_sorted = vals ? valt : nullptr;
  return GDR_OK;
}

}  // namespace gdr

extern "C" {

int64_t gdr_sort_pairs_ws_bytes(int64_t n) { return gdr::sort_pairs_ws_bytes(n); }

int gdr_sort_pairs(int64_t n, int key_bits, uint64_t* keys_io, uint32_t* vals_io, void* ws,
                   int64_t ws_bytes, gdr_stream_t stream) {
  GDR_CHECK_ARG(n >= 0 && key_bits >= 0 && key_bits <= 64, "sort_pairs: bad n/key_bits");
  GDR_CHECK_ARG(n == 0 || keys_io, "sort_pairs: null keys");
  return gdr::sort_pairs(n, key_bits, keys_io, vals_io, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
