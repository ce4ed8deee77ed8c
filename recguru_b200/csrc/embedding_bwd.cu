// Piece 2 (K2): deterministic embedding-table backward.
//
// Replaces ATen embedding_dense_backward at every lookup site of the reference plus the
// dense add_ that sums the per-site results (triggered by loss.backward():
// /root/reference/GURU/train_auto.py:124, gan_training.py:491-564,:860-861; lookup sites
// AutoEnc4Rec.py:74,:92,:115-127,:180,:192,:219-221; AutoEnc4Rec_cross.py:98-102,:203-210).
//
//   dW[i,:] = sum over gathered occurrences j with id_j == i of  coef_j * X[row_j,:]
//
// Pipeline (all on one stream, no atomics on floating point => bit-reproducible):
//   1. build_keys      : concatenate the sites' ids as uint32 keys (occurrence order)
//   2. LSD radix sort  : ceil(bits/9) passes of {tile histogram, scan, stable scatter};
//                        stable => equal ids stay in occurrence order, identical to
//                        torch.sort(ids, stable=True)
//   3. zero_fill       : dense-grad contract of the reference's Adam (optional)
//   4. chunk_reduce    : one warp per kChunk sorted occurrences; runs that lie inside the
//                        chunk are written straight to dW, runs that cross a chunk edge
//                        leave a partial row in scratch
//   5. span_fixup      : one warp per run that crosses chunk edges sums its partials in
//                        chunk order
// The order of every floating-point addition depends only on the (sorted) data.
//
// HBM roofline: per occurrence 8 B id + 4d B source row read; 4d B per distinct row
// written; + 4*V1*d B for the zero fill.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace rg {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;                            // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;     // 2048 keys per CTA
constexpr int kCompactItems = 4;                         // occurrences per thread (compaction)
constexpr int kCompactTile = kSortThreads * kCompactItems;
constexpr int kMaxRadixBits = 9;
constexpr int kMaxBins = 1 << kMaxRadixBits;
constexpr int kChunk = 64;                               // sorted occurrences per warp
constexpr int kReduceWarps = 8;

inline bool env_row_mode_off() {   // RECGURU_BWD_ROWMODE=0: always compact by gathered index
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RECGURU_BWD_ROWMODE");
    v = (e && e[0] == '0') ? 1 : 0;
  }
  return v != 0;
}

struct SiteTable {
  const int64_t* ids[RECGURU_MAX_SITES];
  const float* X[RECGURU_MAX_SITES];
  const float* coef[RECGURU_MAX_SITES];
  const float* scale[RECGURU_MAX_SITES];
  const float* scale2[RECGURU_MAX_SITES];
  const float* live[RECGURU_MAX_SITES];   // optional liveness mask, one entry per row of X
  int64_t offset[RECGURU_MAX_SITES + 1];  // prefix sum of counts
  int group[RECGURU_MAX_SITES];
  int n_sites;
};

// Site lookup with compile-time indices only: dynamic indexing into the by-value parameter
// struct would make the compiler copy it to local memory in every thread.
struct SiteRef {
  const int64_t* ids;
  const float* X;
  const float* coef;
  const float* scale;
  const float* scale2;
  const float* live;
  int64_t offset;
  int group;
  int index;
};
template <int NS>  // compile-time bound on the number of sites (the select chain is per element)
__device__ __forceinline__ SiteRef find_site(const SiteTable& st, int64_t occ) {
  SiteRef r = {st.ids[0], st.X[0], st.coef[0], st.scale[0], st.scale2[0], st.live[0], 0, st.group[0], 0};
#pragma unroll
  for (int i = 1; i < NS; ++i) {
    if (i < st.n_sites && occ >= st.offset[i]) {
      r.ids = st.ids[i]; r.X = st.X[i]; r.coef = st.coef[i]; r.scale = st.scale[i];
      r.scale2 = st.scale2[i]; r.live = st.live[i]; r.offset = st.offset[i]; r.group = st.group[i]; r.index = i;
    }
  }
  return r;
}
__device__ __forceinline__ const float* site_X(const SiteTable& st, int s) {
  const float* x = st.X[0];
#pragma unroll
  for (int i = 1; i < RECGURU_MAX_SITES; ++i)
    if (s == i) x = st.X[i];
  return x;
}

// Dead occurrences -- id == padding_idx, or an exactly-zero coefficient (masked positions: the
// reference multiplies their rows by 0 too) -- are dropped by a stable stream compaction before
// the sort: on padded batches > 90 % of the gathered indices carry no gradient at all.
template <int NS>
__device__ __forceinline__ bool occurrence_live(const SiteTable& st, int64_t i, int64_t padding_idx,
                                                uint32_t* key) {
  const SiteRef r = find_site<NS>(st, i);
  const int64_t j = i - r.offset;
  const int64_t id = r.ids[j];
  *key = static_cast<uint32_t>(id);
  if (id == padding_idx) return false;
  // a site with a liveness mask is planned from the mask alone (its coefficients may not exist
  // yet: the plan of a step can run while the forward pass is still computing them)
  if (r.live != nullptr) return r.live[j / r.group] != 0.f;
  return !(r.coef != nullptr && r.coef[j] == 0.f);
}

// pass 1: live occurrences per tile (same warp-striped tile layout as the sort)
template <int NS>
__global__ void __launch_bounds__(kSortThreads)
compact_count_kernel(SiteTable st, int64_t total, int64_t padding_idx, uint32_t* __restrict__ tile_count) {
  rg::pdl_wait();
  __shared__ uint32_t s_w[kSortThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kCompactTile + warp * (32 * kCompactItems);
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    uint32_t key;
    c += (i < total && occurrence_live<NS>(st, i, padding_idx, &key)) ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) s_w[warp] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) t += s_w[w];
    tile_count[blockIdx.x] = t;
  }
  rg::pdl_trigger();
}

// pass 2: exclusive scan of the tile counts (one CTA) + the live total
__global__ void __launch_bounds__(1024)
compact_scan_kernel(uint32_t* __restrict__ tile_count, int n_tiles, uint32_t* __restrict__ n_live) {
  rg::pdl_wait();
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n_tiles; t0 += 1024) {
    const int i = t0 + tid;
    const uint32_t v = i < n_tiles ? tile_count[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += x;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_warp[w];
    const uint32_t carry = s_carry;
    if (i < n_tiles) tile_count[i] = carry + wbase + incl - v;
    __syncthreads();
    if (tid == 1023) s_carry = carry + wbase + incl;
    __syncthreads();
  }
  if (tid == 0) *n_live = s_carry;
  rg::pdl_trigger();
}

// pass 3: stable scatter of the live (key, occurrence index) pairs
template <int NS>
__global__ void __launch_bounds__(kSortThreads)
compact_scatter_kernel(SiteTable st, int64_t total, int64_t padding_idx,
                       const uint32_t* __restrict__ tile_offset, uint32_t* __restrict__ keys,
                       uint32_t* __restrict__ vals) {
  rg::pdl_wait();
  __shared__ uint32_t s_w[kSortThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kCompactTile + warp * (32 * kCompactItems);
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t key[kCompactItems];
  uint32_t rank[kCompactItems];   // 0xffffffff = dead
  uint32_t run = 0;
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    const bool live = i < total && occurrence_live<NS>(st, i, padding_idx, &key[j]);
    const uint32_t b = __ballot_sync(0xffffffffu, live);
    rank[j] = live ? run + __popc(b & lt_mask) : 0xffffffffu;
    run += __popc(b);
  }
  if (lane == 0) s_w[warp] = run;
  __syncthreads();
  uint32_t base = tile_offset[blockIdx.x];
#pragma unroll
  for (int w = 0; w < kSortThreads / 32; ++w)
    if (w < warp) base += s_w[w];
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    if (rank[j] != 0xffffffffu) {
      keys[base + rank[j]] = key[j];
      vals[base + rank[j]] = static_cast<uint32_t>(warp0 + j * 32 + lane);
    }
  }
  rg::pdl_trigger();
}

// ---- 1b. compaction by ROW when every site shares one liveness mask -----------------------
// A training step's lookup sites all hang off the same positions (input id, positive, K negatives
// of position r), and the pad mask says per position whether any of them carries gradient.  The
// plan then compacts the 0.2 M positions instead of the 1.4 M gathered indices: 7x less to read
// and scan, and the ids are only read for live positions.  Output order is the same as in the
// generic path -- site-major, then position, then k == ascending occurrence index -- so the sums
// are bit-identical.  An occurrence whose id is padding_idx at a live position (dropped by the
// generic path) gets the key `sentinel` (= V1, sorts last) and is skipped by the reduce kernels.
__global__ void __launch_bounds__(kSortThreads)
row_count_kernel(const float* __restrict__ live, int64_t n_rows, uint32_t* __restrict__ tile_count) {
  rg::pdl_wait();
  __shared__ uint32_t s_w[kSortThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kCompactTile + warp * (32 * kCompactItems);
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    c += (i < n_rows && live[i] != 0.f) ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) s_w[warp] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) t += s_w[w];
    tile_count[blockIdx.x] = t;
  }
  rg::pdl_trigger();
}

// The ascending list of rows with live != 0 on its own (row_count + compact_scan + this): the list the
// row-mode plan leaves as a by-product, for steps whose lookup sites do NOT share one mask.
__global__ void __launch_bounds__(kSortThreads)
rows_only_scatter_kernel(const float* __restrict__ live, int64_t n_rows, const uint32_t* __restrict__ tile_offset,
                         uint32_t* __restrict__ live_rows) {
  rg::pdl_wait();
  __shared__ uint32_t s_w[kSortThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kCompactTile + warp * (32 * kCompactItems);
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t rank[kCompactItems];   // 0xffffffff = dead
  uint32_t run = 0;
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    const bool lv = i < n_rows && live[i] != 0.f;
    const uint32_t b = __ballot_sync(0xffffffffu, lv);
    rank[j] = lv ? run + __popc(b & lt_mask) : 0xffffffffu;
    run += __popc(b);
  }
  if (lane == 0) s_w[warp] = run;
  __syncthreads();
  uint32_t base = tile_offset[blockIdx.x];
#pragma unroll
  for (int w = 0; w < kSortThreads / 32; ++w)
    if (w < warp) base += s_w[w];
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j)
    if (rank[j] != 0xffffffffu) live_rows[base + rank[j]] = static_cast<uint32_t>(warp0 + j * 32 + lane);
  rg::pdl_trigger();
}

template <int NS>
__global__ void __launch_bounds__(kSortThreads)
row_scatter_kernel(SiteTable st, const float* __restrict__ live, int64_t n_rows, int64_t padding_idx,
                   uint32_t sentinel, const uint32_t* __restrict__ tile_offset,
                   const uint32_t* __restrict__ n_live_rows, uint32_t* __restrict__ n_live_occ,
                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ live_rows) {
  rg::pdl_wait();
  __shared__ uint32_t s_w[kSortThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kCompactTile + warp * (32 * kCompactItems);
  const uint32_t lt_mask = (1u << lane) - 1u;
  // (requested together with the mask words: one round trip less in the kernel's dependent chain)
  const uint32_t tile_base = tile_offset[blockIdx.x];
  const uint32_t p_live = *n_live_rows;
  uint32_t rank[kCompactItems];   // 0xffffffff = dead
  uint32_t run = 0;
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    const bool lv = i < n_rows && live[i] != 0.f;
    const uint32_t b = __ballot_sync(0xffffffffu, lv);
    rank[j] = lv ? run + __popc(b & lt_mask) : 0xffffffffu;
    run += __popc(b);
  }
  if (lane == 0) s_w[warp] = run;
  __syncthreads();
  uint32_t base = tile_base;
#pragma unroll
  for (int w = 0; w < kSortThreads / 32; ++w)
    if (w < warp) base += s_w[w];
  uint32_t g_total = 0;
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2)
    if (s2 < st.n_sites) g_total += static_cast<uint32_t>(st.group[s2]);
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_live_occ = p_live * g_total;
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    if (rank[j] == 0xffffffffu) continue;
    const int64_t row = warp0 + j * 32 + lane;
    const uint32_t r = base + rank[j];
    live_rows[r] = static_cast<uint32_t>(row);      // the step's live positions, ascending (scoring kernel)
    uint32_t gpre = 0;
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) {
      if (s2 < st.n_sites) {
        const uint32_t G = static_cast<uint32_t>(st.group[s2]);
        const int64_t* ids = st.ids[s2] + row * G;
        const uint32_t out0 = p_live * gpre + r * G;
        const uint32_t occ0 = static_cast<uint32_t>(st.offset[s2] + row * G);
        for (uint32_t k = 0; k < G; ++k) {
          const int64_t id = ids[k];
          keys[out0 + k] = id == padding_idx ? sentinel : static_cast<uint32_t>(id);
          vals[out0 + k] = occ0 + k;
        }
        gpre += G;
      }
    }
  }
  rg::pdl_trigger();
}

__global__ void __launch_bounds__(256)
narrow_keys_kernel(const int64_t* __restrict__ ids, int64_t total, uint32_t* __restrict__ keys) {
  rg::pdl_wait();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += stride)
    keys[i] = static_cast<uint32_t>(ids[i]);
  rg::pdl_trigger();
}

// ---- 2. radix sort ---------------------------------------------------------------------
// Tile layout: warp w of a CTA owns keys [tile0 + w*512, +512); item j of lane l sits at
// offset j*32 + l, so (j, lane) order == memory order and stability is easy to keep.
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t total_host,
                  const uint32_t* __restrict__ n_live, int shift, int bits,
                  uint32_t* __restrict__ block_hist /* [bins][gridDim.x] */) {
  rg::pdl_wait();
  __shared__ uint32_t s_hist[kMaxBins];
  const int64_t total = n_live ? static_cast<int64_t>(*n_live) : total_host;
  const int64_t tile0 = static_cast<int64_t>(blockIdx.x) * kSortTile;
  // The grid is sized for the CAPACITY (all gathered indices); after compaction only the first
  // ceil(total / tile) tiles hold keys.  The others used to store 512 zeros each into the digit-major
  // histogram (stride = grid: one sector per word) for the scan to read back -- on cfg2's padded batch
  // 648 of 700 tiles.  The scan and the scatter now look at the live tiles only, so dead tiles leave.
  if (tile0 >= total) return;
  const int bins = 1 << bits;
  for (int b = threadIdx.x; b < bins; b += kSortThreads) s_hist[b] = 0;
  __syncthreads();
  const uint32_t mask = bins - 1;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int64_t i = tile0 + j * kSortThreads + threadIdx.x;
    if (i < total) atomicAdd(&s_hist[(keys[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += kSortThreads)
    block_hist[static_cast<int64_t>(b) * gridDim.x + blockIdx.x] = s_hist[b];
  rg::pdl_trigger();
}

// Exclusive scan of one digit's row of the digit-major histogram (CTA b <-> digit b); the row
// total goes to totals[b].  The scatter kernel adds the digit base (scan over totals) itself.
__global__ void __launch_bounds__(256)
radix_scan_rows_kernel(uint32_t* __restrict__ hist, int n_tiles_cap, int64_t total_host,
                       const uint32_t* __restrict__ n_live, uint32_t* __restrict__ totals) {
  rg::pdl_wait();
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* row = hist + static_cast<int64_t>(blockIdx.x) * n_tiles_cap;   // row stride = capacity
  const int64_t total = n_live ? static_cast<int64_t>(*n_live) : total_host;
  const int n_tiles = static_cast<int>((total + kSortTile - 1) / kSortTile);   // live tiles only
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n_tiles; t0 += 256) {
    const int i = t0 + tid;
    const uint32_t v = i < n_tiles ? row[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += x;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
      if (w < warp) wbase += s_warp[w];
    const uint32_t carry = s_carry;
    if (i < n_tiles) row[i] = carry + wbase + incl - v;
    __syncthreads();
    if (tid == 255) s_carry = carry + wbase + incl;
    __syncthreads();
  }
  if (tid == 0) totals[blockIdx.x] = s_carry;
  rg::pdl_trigger();
}

template <bool kFirstPass, typename OutVal>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     int64_t total_host, const uint32_t* __restrict__ n_live, int shift, int bits,
                     const uint32_t* __restrict__ block_base /* row-scanned [bins][grid] */,
                     const uint32_t* __restrict__ totals /* [bins] */,
                     uint32_t* __restrict__ keys_out, OutVal* __restrict__ vals_out,
                     int64_t* __restrict__ keys_out64 /* optional widened copy (last pass) */) {
  rg::pdl_wait();
  __shared__ uint32_t s_cnt[kSortThreads / 32][kMaxBins];
  __shared__ uint32_t s_base[kMaxBins];      // exclusive scan of the digit totals
  __shared__ uint32_t s_wsum[kSortThreads / 32];
  const int64_t total = n_live ? static_cast<int64_t>(*n_live) : total_host;
  if (static_cast<int64_t>(blockIdx.x) * kSortTile >= total) return;   // tile beyond the live range
  const int bins = 1 << bits;
  const uint32_t mask = bins - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Every global load of the tile -- digit totals, the tile's scanned bases, keys, values -- is
  // requested here, one round trip behind the live count: on a padded batch the kernel is 52 CTAs
  // whose duration is its chain of dependent loads (it was five of them).
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kSortTile + warp * (32 * kSortItems);
  uint32_t key[kSortItems];
  uint32_t val_in[kSortItems];
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    key[j] = i < total ? keys_in[i] : 0xffffffffu;
    val_in[j] = (!kFirstPass && i < total) ? vals_in[i] : 0u;
  }
  uint32_t bb[kMaxBins / kSortThreads];
#pragma unroll
  for (int q = 0; q < kMaxBins / kSortThreads; ++q) {
    const int b = q * kSortThreads + threadIdx.x;
    bb[q] = b < bins ? block_base[static_cast<int64_t>(b) * gridDim.x + blockIdx.x] : 0u;
  }
  for (int b = threadIdx.x; b < (kSortThreads / 32) * kMaxBins; b += kSortThreads)
    (&s_cnt[0][0])[b] = 0;
  {
    // 256 threads x 2 consecutive digits = 512 = kMaxBins
    const int b0 = threadIdx.x * 2;
    const uint32_t t0 = b0 < bins ? totals[b0] : 0u;
    const uint32_t t1 = b0 + 1 < bins ? totals[b0 + 1] : 0u;
    uint32_t incl = t0 + t1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += x;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w)
      if (w < warp) wbase += s_wsum[w];
    const uint32_t excl = wbase + incl - (t0 + t1);
    s_base[b0] = excl;
    s_base[b0 + 1] = excl + t0;
  }
  __syncthreads();

  uint32_t rank[kSortItems];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    const bool valid = i < total;
    const uint32_t dig = valid ? ((key[j] >> shift) & mask) : 0xffffffffu;  // invalid: own group
    const uint32_t peers = __match_any_sync(0xffffffffu, dig);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) {
      base = s_cnt[warp][dig];
      s_cnt[warp][dig] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[j] = base + __popc(peers & lt_mask);
  }
  __syncthreads();
  // per-digit exclusive prefix over the warps of this CTA, seeded with the global base
#pragma unroll
  for (int q = 0; q < kMaxBins / kSortThreads; ++q) {
    const int b = q * kSortThreads + threadIdx.x;
    if (b >= bins) break;
    uint32_t run = s_base[b] + bb[q];
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) {
      const uint32_t c = s_cnt[w][b];
      s_cnt[w][b] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    if (i < total) {
      const uint32_t dig = (key[j] >> shift) & mask;
      const uint32_t dst = s_cnt[warp][dig] + rank[j];
      keys_out[dst] = key[j];
      vals_out[dst] = static_cast<OutVal>(kFirstPass ? static_cast<uint32_t>(i) : val_in[j]);
      if (keys_out64) keys_out64[dst] = static_cast<int64_t>(key[j]);
    }
  }
  rg::pdl_trigger();
}

// ---- 2b. the plan of the table backward in 2 + passes launches (round 2; RECGURU_BWD_PLAN=2) --
// MEASURED SLOWER than the chain above and therefore not the default -- kept as the record of the
// experiment and for sizes where launch count matters more than atomic throughput.
// Round 1 planned a step's table backward with ten launches of 3-13 us (pad mask, row count, scan,
// row scatter, then {tile histogram, row scan, stable scatter} per radix pass): a latency chain
// that kept < 20 % of the SMs busy and was what `apply` waited for.  Same algorithm -- stable
// compaction of the live (id, occurrence) pairs, stable LSD radix sort by id -- with the
// bookkeeping folded into the kernels that already touch the data:
//   plan_count    liveness per compaction tile (optionally DERIVING the shared pad mask and its
//                 population count from the ids: recguru_pad_mask_count folded in), and the zero fill
//                 of every pass's histogram;
//   plan_scatter  each CTA sums the tile counts before it itself (a few hundred words: no scan
//                 kernel), scatters its live pairs and adds them to pass 0's digit histogram with
//                 integer atomics (exact and order independent, so determinism is untouched);
//   radix_pass    each 2048-key CTA forms its digit bases from the tile-major histogram itself -- the
//                 32-tile group counts before its group, the tiles before it inside the group, and
//                 all live groups for the digit totals (coalesced 2 KB rows: no scan kernel, and
//                 O(tiles/32 + 32) rows per CTA, so it scales) -- ranks its keys (match_any per
//                 warp, stable), scatters, and adds the scattered keys to the NEXT pass's histogram.
// Output (sorted keys / occurrence indices, live count) is bit-identical to the round-1 chain.
constexpr int kP2Threads = 256;
constexpr int kP2Items = 8;
constexpr int kP2Tile = kP2Threads * kP2Items;            // 2048 keys per radix CTA
constexpr int kC2Items = 4;
constexpr int kC2Tile = kP2Threads * kC2Items;            // 1024 units (rows or indices) per compaction CTA
constexpr int kP2Group = 32;                              // radix tiles per histogram group
constexpr int kMaxPasses = 4;
// histogram words of one pass: [tiles][512] per-tile counts, then [groups][512] per-group counts
__host__ __device__ inline int64_t p2_hist_words(int64_t tiles) {
  return (tiles + (tiles + kP2Group - 1) / kP2Group) * kMaxBins;
}
// one key of digit `dig` lands at sorted position `dst` of the NEXT pass's input: count it there
__device__ __forceinline__ void p2_count(uint32_t* __restrict__ hist, int64_t tiles, uint32_t dst, uint32_t dig) {
  const int64_t t = dst / kP2Tile;
  atomicAdd(&hist[t * kMaxBins + dig], 1u);
  atomicAdd(&hist[(tiles + t / kP2Group) * kMaxBins + dig], 1u);
}

template <int NS, bool ROW>
__global__ void __launch_bounds__(kP2Threads)
plan_count_kernel(SiteTable st, const float* __restrict__ live, const int64_t* __restrict__ mask_ids,
                  int64_t pad_index, float* __restrict__ mask_out, uint32_t* __restrict__ count3,
                  int64_t n_units, int64_t padding_idx, uint32_t* __restrict__ tile_count,
                  uint32_t* __restrict__ hist, int64_t hist_words) {
  rg::pdl_wait();
  __shared__ uint32_t s_w[kP2Threads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // histograms of every pass start at zero (plan_scatter / radix_pass add to them)
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kP2Threads + threadIdx.x; i < hist_words;
       i += static_cast<int64_t>(gridDim.x) * kP2Threads)
    hist[i] = 0u;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kC2Tile + warp * (32 * kC2Items);
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < kC2Items; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    bool lv = false;
    if (i < n_units) {
      if (ROW) {
        if (mask_ids != nullptr) {
          lv = mask_ids[i] != pad_index;
          mask_out[i] = lv ? 1.f : 0.f;
        } else {
          lv = live[i] != 0.f;
        }
      } else {
        uint32_t key;
        lv = occurrence_live<NS>(st, i, padding_idx, &key);
      }
    }
    c += lv ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) s_w[warp] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kP2Threads / 32; ++w) t += s_w[w];
    tile_count[blockIdx.x] = t;
    if (ROW && mask_ids != nullptr && count3 != nullptr) {
      // population count of the mask, published by the last CTA (same protocol as pad_mask_count_kernel)
      atomicAdd(&count3[0], t);
      __threadfence();
      const uint32_t ticket = atomicAdd(&count3[1], 1u);
      if (ticket == gridDim.x - 1) {
        __threadfence();
        count3[2] = atomicExch(&count3[0], 0u);
        count3[1] = 0u;
      }
    }
  }
  rg::pdl_trigger();
}

template <int NS, bool ROW>
__global__ void __launch_bounds__(kP2Threads)
plan_scatter_kernel(SiteTable st, const float* __restrict__ live, int64_t n_units, int64_t padding_idx,
                    uint32_t sentinel, const uint32_t* __restrict__ tile_count, int n_ctiles,
                    uint32_t* __restrict__ n_live /* [0] live occurrences, [1] live rows */,
                    uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                    uint32_t* __restrict__ hist0, int64_t rtiles, uint32_t digit_mask) {
  rg::pdl_wait();
  __shared__ uint32_t s_w[kP2Threads / 32];
  __shared__ uint32_t s_red[2][kP2Threads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // live units in the tiles before this one, and in all tiles
  uint32_t before = 0, all = 0;
  for (int t = threadIdx.x; t < n_ctiles; t += kP2Threads) {
    const uint32_t v = tile_count[t];
    all += v;
    before += t < static_cast<int>(blockIdx.x) ? v : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    before += __shfl_xor_sync(0xffffffffu, before, o);
    all += __shfl_xor_sync(0xffffffffu, all, o);
  }
  if (lane == 0) { s_red[0][warp] = before; s_red[1][warp] = all; }
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kC2Tile + warp * (32 * kC2Items);
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t key[kC2Items];
  uint32_t rank[kC2Items];   // 0xffffffff = dead
  uint32_t run = 0;
#pragma unroll
  for (int j = 0; j < kC2Items; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    bool lv = false;
    key[j] = 0;
    if (i < n_units) {
      if (ROW) lv = live[i] != 0.f;
      else lv = occurrence_live<NS>(st, i, padding_idx, &key[j]);
    }
    const uint32_t b = __ballot_sync(0xffffffffu, lv);
    rank[j] = lv ? run + __popc(b & lt_mask) : 0xffffffffu;
    run += __popc(b);
  }
  if (lane == 0) s_w[warp] = run;
  __syncthreads();
  uint32_t base = 0, p_live = 0;
#pragma unroll
  for (int w = 0; w < kP2Threads / 32; ++w) {
    base += s_red[0][w];
    p_live += s_red[1][w];
    if (w < warp) base += s_w[w];
  }
  if (ROW) {
    uint32_t g_total = 0;
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2)
      if (s2 < st.n_sites) g_total += static_cast<uint32_t>(st.group[s2]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { n_live[0] = p_live * g_total; n_live[1] = p_live; }
#pragma unroll
    for (int j = 0; j < kC2Items; ++j) {
      if (rank[j] == 0xffffffffu) continue;
      const int64_t row = warp0 + j * 32 + lane;
      const uint32_t r = base + rank[j];
      uint32_t gpre = 0;
#pragma unroll
      for (int s2 = 0; s2 < NS; ++s2) {
        if (s2 < st.n_sites) {
          const uint32_t G = static_cast<uint32_t>(st.group[s2]);
          const int64_t* ids = st.ids[s2] + row * G;
          const uint32_t out0 = p_live * gpre + r * G;
          const uint32_t occ0 = static_cast<uint32_t>(st.offset[s2] + row * G);
          for (uint32_t k = 0; k < G; ++k) {
            const int64_t id = ids[k];
            const uint32_t kk = id == padding_idx ? sentinel : static_cast<uint32_t>(id);
            keys[out0 + k] = kk;
            vals[out0 + k] = occ0 + k;
            p2_count(hist0, rtiles, out0 + k, kk & digit_mask);
          }
          gpre += G;
        }
      }
    }
  } else {
    if (blockIdx.x == 0 && threadIdx.x == 0) { n_live[0] = p_live; n_live[1] = 0u; }
#pragma unroll
    for (int j = 0; j < kC2Items; ++j) {
      if (rank[j] != 0xffffffffu) {
        const uint32_t dst = base + rank[j];
        keys[dst] = key[j];
        vals[dst] = static_cast<uint32_t>(warp0 + j * 32 + lane);
        p2_count(hist0, rtiles, dst, key[j] & digit_mask);
      }
    }
  }
  rg::pdl_trigger();
}

// One LSD pass over 2048-key tiles.  hist_cur (tile-major per-tile + per-group counts) was filled by
// the kernel that wrote keys_in; hist_next (may be NULL) is filled here for the following pass.
__global__ void __launch_bounds__(kP2Threads)
radix_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  const uint32_t* __restrict__ n_live, int shift, int bits,
                  const uint32_t* __restrict__ hist_cur, uint32_t* __restrict__ hist_next, int64_t rtiles,
                  int next_shift, uint32_t next_mask,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  rg::pdl_wait();
  __shared__ uint32_t s_cnt[kP2Threads / 32][kMaxBins];
  __shared__ uint32_t s_base[kMaxBins];
  __shared__ uint32_t s_wsum[kP2Threads / 32];
  const int64_t total = static_cast<int64_t>(*n_live);
  const int64_t tile0 = static_cast<int64_t>(blockIdx.x) * kP2Tile;
  if (tile0 >= total) return;                                          // tile beyond the live range
  const int live_tiles = static_cast<int>((total + kP2Tile - 1) / kP2Tile);
  const int live_groups = (live_tiles + kP2Group - 1) / kP2Group;
  const int my_group = static_cast<int>(blockIdx.x) / kP2Group;
  const int bins = 1 << bits;
  const uint32_t mask = bins - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = threadIdx.x; b < (kP2Threads / 32) * kMaxBins; b += kP2Threads) (&s_cnt[0][0])[b] = 0;
  // digit bases.  Thread t owns digits 2t and 2t+1 (256 x 2 = 512): keys of that digit in the tiles
  // before this one (whole groups, then the tiles of this group) and in all live tiles.
  {
    const uint32_t* grp = hist_cur + rtiles * kMaxBins;
    const int b0 = threadIdx.x * 2;
    uint32_t before0 = 0, before1 = 0, tot0 = 0, tot1 = 0;
    if (b0 < bins) {
      for (int g = 0; g < live_groups; ++g) {
        const uint2 v = *reinterpret_cast<const uint2*>(grp + static_cast<int64_t>(g) * kMaxBins + b0);
        tot0 += v.x; tot1 += v.y;
        if (g < my_group) { before0 += v.x; before1 += v.y; }
      }
      for (int t = my_group * kP2Group; t < static_cast<int>(blockIdx.x); ++t) {
        const uint2 v = *reinterpret_cast<const uint2*>(hist_cur + static_cast<int64_t>(t) * kMaxBins + b0);
        before0 += v.x; before1 += v.y;
      }
    }
    uint32_t incl = tot0 + tot1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += x;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < kP2Threads / 32; ++w)
      if (w < warp) wbase += s_wsum[w];
    const uint32_t excl = wbase + incl - (tot0 + tot1);
    s_base[b0] = excl + before0;
    s_base[b0 + 1] = excl + tot0 + before1;
  }
  __syncthreads();

  const int64_t warp0 = tile0 + warp * (32 * kP2Items);
  uint32_t key[kP2Items];
  uint32_t rank[kP2Items];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < kP2Items; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    const bool valid = i < total;
    key[j] = valid ? keys_in[i] : 0xffffffffu;
    const uint32_t dig = valid ? ((key[j] >> shift) & mask) : 0xffffffffu;  // invalid: own group
    const uint32_t peers = __match_any_sync(0xffffffffu, dig);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) {
      base = s_cnt[warp][dig];
      s_cnt[warp][dig] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[j] = base + __popc(peers & lt_mask);
  }
  __syncthreads();
  // per-digit exclusive prefix over the warps of this CTA, seeded with the global base
  for (int b = threadIdx.x; b < bins; b += kP2Threads) {
    uint32_t run = s_base[b];
#pragma unroll
    for (int w = 0; w < kP2Threads / 32; ++w) {
      const uint32_t c = s_cnt[w][b];
      s_cnt[w][b] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kP2Items; ++j) {
    const int64_t i = warp0 + j * 32 + lane;
    if (i < total) {
      const uint32_t dig = (key[j] >> shift) & mask;
      const uint32_t dst = s_cnt[warp][dig] + rank[j];
      keys_out[dst] = key[j];
      vals_out[dst] = vals_in[i];
      if (hist_next != nullptr) p2_count(hist_next, rtiles, dst, (key[j] >> next_shift) & next_mask);
    }
  }
  rg::pdl_trigger();
}

// ---- 3. zero fill -----------------------------------------------------------------------
__global__ void __launch_bounds__(256) zero_fill_kernel(float4* __restrict__ p, int64_t n_vec) {
  rg::pdl_wait();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
       i += stride)
    p[i] = z;
  rg::pdl_trigger();
}

// ---- 4. chunked segmented reduce -----------------------------------------------------
// VPL = float4 per lane (ceil(d/4/32)).
template <int VPL, int NS>
__global__ void __launch_bounds__(kReduceWarps * 32, VPL == 1 ? 2 : 1)
chunk_reduce_kernel(SiteTable st, const uint32_t* __restrict__ skeys,
                    const uint32_t* __restrict__ svals, const uint32_t* __restrict__ n_live,
                    int vec /* d/4 */,
                    int64_t padding_idx, int accumulate, float4* __restrict__ dW,
                    float4* __restrict__ partial /* [n_chunks][2][vec] */) {
  rg::pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total = *n_live;
  const int64_t chunk = static_cast<int64_t>(blockIdx.x) * kReduceWarps + warp;
  const int64_t c0 = chunk * kChunk;
  if (c0 >= total) return;
  const int n = static_cast<int>(rg::min64(kChunk, total - c0));

  // lane-parallel decode of this chunk's occurrences (kChunk/32 per lane)
  constexpr int kPer = kChunk / 32;
  uint32_t l_key[kPer];
  int l_site[kPer];
  int l_row[kPer];
  float l_coef[kPer];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int i = q * 32 + lane;
    l_key[q] = 0xffffffffu;
    l_site[q] = 0;
    l_row[q] = 0;
    l_coef[q] = 0.f;
    if (i < n) {
      l_key[q] = skeys[c0 + i];
      const int64_t occ = svals[c0 + i];
      const SiteRef r = find_site<NS>(st, occ);
      const int64_t j = occ - r.offset;
      l_site[q] = r.index;
      l_row[q] = static_cast<int>(j / r.group);
      float c = r.coef ? r.coef[j] : 1.f;
      if (r.scale) c *= *r.scale;
      if (r.scale2) c *= *r.scale2;
      l_coef[q] = c;
    }
  }
  const uint32_t first_key = __shfl_sync(0xffffffffu, l_key[0], 0);
  const uint32_t last_key = __shfl_sync(0xffffffffu, l_key[(n - 1) / 32], (n - 1) & 31);
  const uint32_t pad = static_cast<uint32_t>(padding_idx);  // -1 -> 0xffffffff never matches
  const bool has_pad = padding_idx >= 0;
  if (has_pad && first_key == pad) return;  // sentinel keys sort last: the whole chunk is dead
  const bool lead_open = c0 > 0 && skeys[c0 - 1] == first_key;           // run started earlier
  const bool trail_open = c0 + n < total && skeys[c0 + n] == last_key;   // run continues

  float4 acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t cur = first_key;
  bool cur_is_lead = lead_open;

  constexpr int kBatch = 8;                  // rows in flight per warp (12/16 measured slower: spills)
  for (int b0 = 0; b0 < n; b0 += kBatch) {
    float4 x[kBatch][VPL];
    uint32_t k[kBatch];
    float cf[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = b0 + u;
      const int q = (i >> 5) < kPer ? (i >> 5) : kPer - 1;
      // select the register of slot q (kPer is tiny; keep it branch-free)
      uint32_t kk = l_key[0];
      int ss = l_site[0];
      int rr = l_row[0];
      float cc = l_coef[0];
#pragma unroll
      for (int t = 1; t < kPer; ++t)
        if (q == t) { kk = l_key[t]; ss = l_site[t]; rr = l_row[t]; cc = l_coef[t]; }
      k[u] = __shfl_sync(0xffffffffu, kk, i & 31);
      const int s = __shfl_sync(0xffffffffu, ss, i & 31);
      const int64_t row = __shfl_sync(0xffffffffu, rr, i & 31);
      cf[u] = __shfl_sync(0xffffffffu, cc, i & 31);
      const bool live = i < n && !(has_pad && k[u] == pad);
      const float4* src = reinterpret_cast<const float4*>(site_X(st, s)) + row * vec;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = v * 32 + lane;
        x[u][v] = (live && col < vec) ? ld_stream_f4(src + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = b0 + u;
      if (i < n) {
        if (k[u] != cur) {
          // close the run `cur` (it ends inside this chunk)
          if (!(has_pad && cur == pad)) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              const int col = v * 32 + lane;
              if (col < vec) {
                if (cur_is_lead) {
                  partial[(chunk * 2 + 0) * vec + col] = acc[v];
                } else {
                  float4* dst = dW + static_cast<int64_t>(cur) * vec + col;
                  float4 o = acc[v];
                  if (accumulate) {
                    const float4 old = *dst;
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                  }
                  *dst = o;
                }
              }
            }
          }
#pragma unroll
          for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          cur = k[u];
          cur_is_lead = false;
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          acc[v].x = fmaf(cf[u], x[u][v].x, acc[v].x);
          acc[v].y = fmaf(cf[u], x[u][v].y, acc[v].y);
          acc[v].z = fmaf(cf[u], x[u][v].z, acc[v].z);
          acc[v].w = fmaf(cf[u], x[u][v].w, acc[v].w);
        }
      }
    }
  }
  // close the last run of the chunk
  if (!(has_pad && cur == pad)) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = v * 32 + lane;
      if (col < vec) {
        if (cur_is_lead) {
          partial[(chunk * 2 + 0) * vec + col] = acc[v];       // continues from the left
        } else if (trail_open) {
          partial[(chunk * 2 + 1) * vec + col] = acc[v];       // starts here, continues right
        } else {
          float4* dst = dW + static_cast<int64_t>(cur) * vec + col;
          float4 o = acc[v];
          if (accumulate) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *dst = o;
        }
      }
    }
  }
  rg::pdl_trigger();
}

// ---- 4b. the same reduce with the source rows staged through shared memory by TMA --------
// chunk_reduce_kernel keeps 8 rows per warp in flight in REGISTERS (124 of them): at 11 resident
// warps per SM that is 45 KB in flight, and a chunk is eight dependent load round trips -- a
// latency-bound gather (0.3 of the HBM roofline).  Here every lane issues ONE 1-D bulk copy
// (cp.async.bulk, completing on the warp's mbarrier) for its own occurrence's row straight from the
// lane-parallel decode -- up to 32 rows = 16 KB per warp in flight, no registers, no shuffles to
// issue -- and the warp then adds the rows out of shared memory in sorted order.  Same additions in
// the same order as chunk_reduce_kernel => bit-identical dW and partial rows.
constexpr int kCrWarps = 4;                   // warps per CTA (16 KB of staging each)
constexpr int kCrStageBytes = 16384;
template <int VPL, int NS>
__global__ void __launch_bounds__(kCrWarps * 32)
chunk_reduce_tma_kernel(SiteTable st, const uint32_t* __restrict__ skeys,
                        const uint32_t* __restrict__ svals, const uint32_t* __restrict__ n_live,
                        int vec /* d/4 */, int group_rows /* rows staged at once: min(32, 1024/vec) */,
                        int64_t padding_idx, int accumulate, float4* __restrict__ dW,
                        float4* __restrict__ partial /* [n_chunks][2][vec] */) {
  rg::pdl_wait();
  extern __shared__ __align__(128) unsigned char cr_smem[];
  __shared__ __align__(8) uint64_t s_bar[kCrWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total = *n_live;
  const int64_t chunk = static_cast<int64_t>(blockIdx.x) * kCrWarps + warp;
  const int64_t c0 = chunk * kChunk;
  if (c0 >= total) return;
  const int n = static_cast<int>(rg::min64(kChunk, total - c0));
  float4* buf = reinterpret_cast<float4*>(cr_smem + warp * kCrStageBytes);
  uint64_t* bar = &s_bar[warp];
  if (lane == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_fence_init();
  }
  __syncwarp();

  constexpr int kPer = kChunk / 32;
  uint32_t l_key[kPer];
  const float4* l_src[kPer];
  float l_coef[kPer];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int i = q * 32 + lane;
    l_key[q] = 0xffffffffu;
    l_src[q] = nullptr;
    l_coef[q] = 0.f;
    if (i < n) {
      l_key[q] = skeys[c0 + i];
      const int64_t occ = svals[c0 + i];
      const SiteRef r = find_site<NS>(st, occ);
      const int64_t j = occ - r.offset;
      l_src[q] = reinterpret_cast<const float4*>(r.X) + (j / r.group) * vec;
      float c = r.coef ? r.coef[j] : 1.f;
      if (r.scale) c *= *r.scale;
      if (r.scale2) c *= *r.scale2;
      l_coef[q] = c;
    }
  }
  const uint32_t first_key = __shfl_sync(0xffffffffu, l_key[0], 0);
  const uint32_t last_key = __shfl_sync(0xffffffffu, l_key[(n - 1) / 32], (n - 1) & 31);
  const uint32_t pad = static_cast<uint32_t>(padding_idx);  // -1 -> 0xffffffff never matches
  const bool has_pad = padding_idx >= 0;
  if (has_pad && first_key == pad) return;  // sentinel keys sort last: the whole chunk is dead
  const bool lead_open = c0 > 0 && skeys[c0 - 1] == first_key;           // run started earlier
  const bool trail_open = c0 + n < total && skeys[c0 + n] == last_key;   // run continues

  float4 acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t cur = first_key;
  bool cur_is_lead = lead_open;
  uint32_t phase = 0;
  const uint32_t row_bytes = static_cast<uint32_t>(vec) * 16u;

#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    for (int r0 = 0; r0 < 32 && q * 32 + r0 < n; r0 += group_rows) {
      const bool mine = lane >= r0 && lane < r0 + group_rows && q * 32 + lane < n && !(has_pad && l_key[q] == pad);
      const uint32_t live_mask = __ballot_sync(0xffffffffu, mine);
      if (live_mask != 0u) {
        ptx::fence_proxy_async_smem();         // the stage's earlier generic reads precede the async writes
        if (lane == 0) ptx::mbar_arrive_expect_tx(bar, __popc(live_mask) * row_bytes);
        __syncwarp();
        if (mine) ptx::bulk_copy_g2s(buf + (lane - r0) * vec, l_src[q], row_bytes, bar);
        ptx::mbar_wait(bar, phase);
        phase ^= 1u;
      }
      const int r_end = min(min(r0 + group_rows, 32), n - q * 32);
      for (int r = r0; r < r_end; ++r) {
        const uint32_t k = __shfl_sync(0xffffffffu, l_key[q], r);
        const float cf = __shfl_sync(0xffffffffu, l_coef[q], r);
        if (k != cur) {
          // close the run `cur` (it ends inside this chunk)
          if (!(has_pad && cur == pad)) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              const int col = v * 32 + lane;
              if (col < vec) {
                if (cur_is_lead) {
                  partial[(chunk * 2 + 0) * vec + col] = acc[v];
                } else {
                  float4* dst = dW + static_cast<int64_t>(cur) * vec + col;
                  float4 o = acc[v];
                  if (accumulate) {
                    const float4 old = *dst;
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                  }
                  *dst = o;
                }
              }
            }
          }
#pragma unroll
          for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          cur = k;
          cur_is_lead = false;
        }
        // (a dead occurrence -- sentinel key -- contributed cf * 0 in the register version)
        const bool live = (live_mask >> r) & 1u;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = v * 32 + lane;
          const float4 x = (live && col < vec) ? buf[(r - r0) * vec + col] : make_float4(0.f, 0.f, 0.f, 0.f);
          acc[v].x = fmaf(cf, x.x, acc[v].x);
          acc[v].y = fmaf(cf, x.y, acc[v].y);
          acc[v].z = fmaf(cf, x.z, acc[v].z);
          acc[v].w = fmaf(cf, x.w, acc[v].w);
        }
      }
      __syncwarp();                            // every lane has read the stage before it is refilled
    }
  }
  // close the last run of the chunk
  if (!(has_pad && cur == pad)) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = v * 32 + lane;
      if (col < vec) {
        if (cur_is_lead) {
          partial[(chunk * 2 + 0) * vec + col] = acc[v];       // continues from the left
        } else if (trail_open) {
          partial[(chunk * 2 + 1) * vec + col] = acc[v];       // starts here, continues right
        } else {
          float4* dst = dW + static_cast<int64_t>(cur) * vec + col;
          float4 o = acc[v];
          if (accumulate) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *dst = o;
        }
      }
    }
  }
  rg::pdl_trigger();
}

// ---- 5. runs that cross chunk edges --------------------------------------------------
// One warp per chunk edge b (between chunks b-1 and b).  The warp owns the run iff the
// edge is the FIRST edge that run crosses; it then walks right summing partials.
template <int VPL>
__global__ void __launch_bounds__(kReduceWarps * 32)
span_fixup_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ n_live, int vec,
                  int64_t padding_idx, int accumulate, float4* __restrict__ dW,
                  const float4* __restrict__ partial) {
  rg::pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total = *n_live;
  const int64_t n_chunks = (total + kChunk - 1) / kChunk;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kReduceWarps + warp + 1;
  if (b >= n_chunks) return;
  // Every key the decision needs -- either side of this edge and of the previous one, the probes of
  // the next edges (one lane each) -- is requested together, one round trip behind the live count;
  // then, for an owner, the two partial rows of a run that crosses exactly this edge together: three
  // dependent round trips instead of six (apply 32.5 -> 30.3 us on cfg2's padded batch; requesting
  // the rows along with the keys as well measured the same).
  constexpr int kBatch = VPL == 1 ? 16 : 8;
  const uint32_t key = skeys[b * kChunk];
  const uint32_t key_l = skeys[b * kChunk - 1];
  const uint32_t prev_r = b >= 2 ? skeys[(b - 1) * kChunk] : 0u;
  const uint32_t prev_l = b >= 2 ? skeys[(b - 1) * kChunk - 1] : 0u;
  const int64_t e0 = b + 1 + lane;
  const bool probe0 = lane < kBatch && e0 < n_chunks;
  const uint32_t e0_r = probe0 ? skeys[e0 * kChunk] : 0u;
  const uint32_t e0_l = probe0 ? skeys[e0 * kChunk - 1] : 0u;
  if (key_l != key) return;                                       // no run crosses this edge
  if (padding_idx >= 0 && key == static_cast<uint32_t>(padding_idx)) return;
  // run started before chunk b-1?  then an earlier edge owns it
  if (b >= 2 && prev_r == key && prev_l == key) return;
  float4 acc[VPL], first_lead[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = v * 32 + lane;
    // first piece: chunk b-1.  It is the trailing partial unless the run fills chunk b-1
    // from its first element *and* crossed edge b-1 (excluded above), so slot 1 always.
    acc[v] = col < vec ? partial[((b - 1) * 2 + 1) * vec + col] : make_float4(0.f, 0.f, 0.f, 0.f);
    first_lead[v] = col < vec ? partial[(b * 2 + 0) * vec + col] : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  int64_t c = b;  // invariant: edge c is crossed by this run => chunk c has a leading partial
  bool done = false;
  bool first = true;
  while (!done) {
    // how many of the next kBatch chunks still belong to this run?  One lane per edge, so the
    // key probes are a single round trip instead of a dependent chain.
    const int64_t e = c + 1 + lane;  // edge index == index of the chunk on its right
    bool crossed;
    if (first) crossed = probe0 && e0_r == key && e0_l == key;
    else crossed = lane < kBatch && e < n_chunks && skeys[e * kChunk] == key && skeys[e * kChunk - 1] == key;
    const uint32_t open_mask = __ballot_sync(0xffffffffu, !crossed);   // first lane whose edge is not crossed
    const int first_closed = __ffs(open_mask) - 1;                     // in [0, kBatch]
    int cnt = first_closed + 1;                                        // chunks c .. c+first_closed
    if (cnt > kBatch) cnt = kBatch; else done = true;
    float4 x[kBatch][VPL];
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = v * 32 + lane;
        if (u == 0 && first) x[u][v] = first_lead[v];           // (already here)
        else x[u][v] = (u < cnt && col < vec) ? partial[((c + u) * 2 + 0) * vec + col]
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    first = false;
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
      if (u < cnt) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          acc[v].x += x[u][v].x; acc[v].y += x[u][v].y;
          acc[v].z += x[u][v].z; acc[v].w += x[u][v].w;
        }
      }
    c += cnt;
  }
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = v * 32 + lane;
    if (col < vec) {
      float4* dst = dW + static_cast<int64_t>(key) * vec + col;
      float4 o = acc[v];
      if (accumulate) {
        const float4 old = *dst;
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *dst = o;
    }
  }
  rg::pdl_trigger();
}

// ---- host orchestration -------------------------------------------------------------
struct SortPlan {
  int total_bits, passes, bits_per_pass;
  int64_t n_tiles;
  // workspace carve-up (bytes, 256-aligned)
  int64_t off_keys_a, off_keys_b, off_vals_a, off_vals_b, off_hist, off_totals, off_tilecnt, off_nlive, off_rows, end;
};

static SortPlan make_sort_plan(int64_t count, int64_t V1) {
  SortPlan p;
  int bits = 1;
  while ((1ll << bits) < V1) ++bits;
  p.total_bits = bits;
  p.passes = (bits + kMaxRadixBits - 1) / kMaxRadixBits;
  p.bits_per_pass = (bits + p.passes - 1) / p.passes;
  p.n_tiles = ceil_div(count > 0 ? count : 1, kSortTile);
  int64_t o = 0;
  const int64_t kb = align_up(count * 4, 256);
  p.off_keys_a = o; o += kb;
  p.off_keys_b = o; o += kb;
  p.off_vals_a = o; o += kb;
  p.off_vals_b = o; o += kb;
  {
    // the round-1 chain keeps one digit-major histogram over 2048-key tiles; the fused plan keeps one
    // tile-major histogram per pass over 8192-key tiles
    const int64_t old_b = static_cast<int64_t>(kMaxBins) * p.n_tiles * 4;
    const int64_t new_b = static_cast<int64_t>(p.passes) * p2_hist_words(ceil_div(count > 0 ? count : 1, kP2Tile)) * 4;
    p.off_hist = o; o += align_up(old_b > new_b ? old_b : new_b, 256);
  }
  p.off_totals = o; o += align_up(kMaxBins * 4, 256);
  p.off_tilecnt = o; o += align_up(ceil_div(count > 0 ? count : 1, kCompactTile) * 4, 256);
  p.off_nlive = o; o += 256;
  p.off_rows = o; o += kb;                 // compacted list of live rows (row mode; rows <= count)
  p.end = o;
  return p;
}

// Sorts keys (already in ws keys_a).  On return *skeys / *svals point at the sorted arrays.
// If out_keys64/out_perm64 are given the final pass writes widened copies there instead.
// n_live (device) overrides `count` when given; first_pass_vals: pass 0 takes vals from keys' twin
// array (compacted input) instead of generating the identity.
static int run_radix_sort(const SortPlan& p, char* ws, int64_t count, const uint32_t* n_live,
                          bool identity_vals, cudaStream_t st, uint32_t** skeys, uint32_t** svals,
                          int64_t* out_keys64, int64_t* out_perm64) {
  uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(ws + p.off_keys_a),
                       reinterpret_cast<uint32_t*>(ws + p.off_keys_b)};
  uint32_t* vals[2] = {reinterpret_cast<uint32_t*>(ws + p.off_vals_a),
                       reinterpret_cast<uint32_t*>(ws + p.off_vals_b)};
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + p.off_hist);
  uint32_t* totals = reinterpret_cast<uint32_t*>(ws + p.off_totals);
  const unsigned grid = static_cast<unsigned>(p.n_tiles);
  int cur = 0;
  for (int pass = 0; pass < p.passes; ++pass) {
    const int shift = pass * p.bits_per_pass;
    const int bits = min(p.bits_per_pass, p.total_bits - shift);
    rg::launch_pdl((radix_hist_kernel), grid, kSortThreads, 0, st, keys[cur], count, n_live, shift, bits, hist);
    RG_CHECK_LAUNCH("radix_hist");
    rg::launch_pdl((radix_scan_rows_kernel), 1u << bits, 256, 0, st, hist, static_cast<int>(p.n_tiles), count, n_live, totals);
    RG_CHECK_LAUNCH("radix_scan_rows");
    const bool last = pass == p.passes - 1;
    if (last && out_perm64) {
      if (pass == 0 && identity_vals)
        rg::launch_pdl((radix_scatter_kernel<true, int64_t>), grid, kSortThreads, 0, st, 
            keys[cur], vals[cur], count, n_live, shift, bits, hist, totals, keys[cur ^ 1], out_perm64, out_keys64);
      else
        rg::launch_pdl((radix_scatter_kernel<false, int64_t>), grid, kSortThreads, 0, st, 
            keys[cur], vals[cur], count, n_live, shift, bits, hist, totals, keys[cur ^ 1], out_perm64, out_keys64);
    } else {
      if (pass == 0 && identity_vals)
        rg::launch_pdl((radix_scatter_kernel<true, uint32_t>), grid, kSortThreads, 0, st, 
            keys[cur], vals[cur], count, n_live, shift, bits, hist, totals, keys[cur ^ 1], vals[cur ^ 1], nullptr);
      else
        rg::launch_pdl((radix_scatter_kernel<false, uint32_t>), grid, kSortThreads, 0, st, 
            keys[cur], vals[cur], count, n_live, shift, bits, hist, totals, keys[cur ^ 1], vals[cur ^ 1], nullptr);
    }
    RG_CHECK_LAUNCH("radix_scatter");
    cur ^= 1;
  }
  *skeys = keys[cur];
  *svals = vals[cur];
  return RECGURU_OK;
}

}  // namespace rg

extern "C" int64_t recguru_sort_ids_workspace_bytes(int64_t count, int64_t V1) {
  if (count < 0 || V1 <= 0) return RECGURU_E_SHAPE;
  return rg::make_sort_plan(count, V1).end;
}

extern "C" int recguru_sort_ids(const int64_t* ids, int64_t count, int64_t V1, int64_t* sorted_ids,
                                int64_t* perm, void* workspace, int64_t workspace_bytes,
                                recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(count >= 0 && V1 > 0 && V1 < (1ll << 31), RECGURU_E_SHAPE, "sort_ids: bad sizes");
  RG_REQUIRE(count < (1ll << 32), RECGURU_E_UNSUPPORTED, "sort_ids: count must be < 2^32");
  if (count == 0) return RECGURU_OK;
  RG_REQUIRE(ids && sorted_ids && perm && workspace, RECGURU_E_SHAPE, "sort_ids: null pointer");
  const SortPlan p = make_sort_plan(count, V1);
  RG_REQUIRE(workspace_bytes >= p.end, RECGURU_E_WORKSPACE, "sort_ids: workspace %lld < %lld",
             (long long)workspace_bytes, (long long)p.end);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  int64_t blocks = rg::min64(ceil_div(count, 256 * 4), static_cast<int64_t>(num_sms()) * 8);
  rg::launch_pdl((narrow_keys_kernel), static_cast<unsigned>(blocks), 256, 0, st, 
      ids, count, reinterpret_cast<uint32_t*>(ws + p.off_keys_a));
  RG_CHECK_LAUNCH("narrow_keys");
  uint32_t *sk, *sv;
  return run_radix_sort(p, ws, count, nullptr, true, st, &sk, &sv, sorted_ids, perm);
}

extern "C" int64_t recguru_embedding_bwd_workspace_bytes(int64_t total_count, int64_t V1,
                                                         int64_t d) {
  if (total_count < 0 || V1 <= 0 || d <= 0) return RECGURU_E_SHAPE;
  const rg::SortPlan p = rg::make_sort_plan(total_count, V1);
  const int64_t n_chunks = rg::ceil_div(total_count > 0 ? total_count : 1, rg::kChunk);
  return p.end + rg::align_up(n_chunks * 2 * d * 4, 256);
}

namespace rg {
// validate + flatten the host site array
static int build_site_table(const char* who, const recguru_lookup_site_t* sites_host, int n_sites,
                            bool need_x, SiteTable* tab, int64_t* total_out) {
  RG_REQUIRE(sites_host && n_sites >= 1 && n_sites <= RECGURU_MAX_SITES, RECGURU_E_SHAPE,
             "%s: n_sites=%d out of range [1,%d]", who, n_sites, RECGURU_MAX_SITES);
  memset(tab, 0, sizeof(*tab));
  tab->n_sites = n_sites;
  int64_t total = 0;
  for (int s = 0; s < n_sites; ++s) {
    const recguru_lookup_site_t& site = sites_host[s];
    RG_REQUIRE(site.count >= 0 && site.group >= 1 && site.count % site.group == 0, RECGURU_E_SHAPE,
               "%s: site %d count=%lld group=%lld", who, s, (long long)site.count, (long long)site.group);
    RG_REQUIRE(site.count == 0 || (site.ids && (!need_x || (site.X && aligned16(site.X)))), RECGURU_E_SHAPE,
               "%s: site %d null/unaligned pointer", who, s);
    tab->ids[s] = site.ids; tab->X[s] = site.X; tab->coef[s] = site.coef; tab->scale[s] = site.scale;
    tab->scale2[s] = site.scale2; tab->live[s] = site.live;
    tab->group[s] = static_cast<int>(site.group);
    tab->offset[s] = total;
    total += site.count;
  }
  for (int s = n_sites; s <= RECGURU_MAX_SITES; ++s) tab->offset[s] = total;
  RG_REQUIRE(total < (1ll << 32), RECGURU_E_UNSUPPORTED, "%s: total count >= 2^32", who);
  *total_out = total;
  return RECGURU_OK;
}
}  // namespace rg

namespace rg {
// All sites carry the same liveness mask over the same number of rows => compaction by row.
static bool row_mode(const recguru_lookup_site_t* sites, int n_sites, int64_t* n_rows) {
  if (env_row_mode_off()) return false;
  const float* lv = sites[0].live;
  if (lv == nullptr || sites[0].group < 1) return false;
  const int64_t rows = sites[0].count / sites[0].group;
  for (int s = 0; s < n_sites; ++s)
    if (sites[s].live != lv || sites[s].group < 1 || sites[s].count != rows * sites[s].group) return false;
  *n_rows = rows;
  return rows > 0;
}
}  // namespace rg

extern "C" int recguru_fill_zero(float* p, int64_t n_floats, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(n_floats >= 0 && n_floats % 4 == 0 && (n_floats == 0 || (p && aligned16(p))), RECGURU_E_SHAPE,
             "fill_zero: need a 16-byte aligned buffer of a multiple of 4 floats");
  if (n_floats == 0) return RECGURU_OK;
  const int64_t n_vec = n_floats / 4;
  const int64_t blocks = rg::min64(ceil_div(n_vec, 256 * 4), static_cast<int64_t>(num_sms()) * 16);
  rg::launch_pdl((zero_fill_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<float4*>(p), n_vec);
  RG_CHECK_LAUNCH("zero_fill");
  return RECGURU_OK;
}

namespace rg {
// RECGURU_BWD_PLAN=2 selects the fused 2 + passes launch plan (plan_count / plan_scatter / radix_pass).
// It is bit-identical to the default chain but was measured SLOWER on B200 (cfg2 padded batch: 78 us
// against 65 us; the per-key global histogram atomics cost more than the launches they replace --
// plan_scatter 20.8 us vs row_scatter 12.5, a radix pass with next-pass counting 31.8 us vs 15.2
// without), so the default stays the separate histogram / scan / scatter kernels.
inline bool env_plan_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RECGURU_BWD_PLAN");
    v = (e && e[0] == '2') ? 0 : 1;
  }
  return v != 0;
}

static int plan_v1(const recguru_lookup_site_t* sites_host, int n_sites, const SiteTable& tab, int64_t total,
                   int64_t V1, int64_t padding_idx, bool by_row, int64_t n_rows, const SortPlan& p, char* ws,
                   int stage, cudaStream_t st) {
  uint32_t* tile_cnt = reinterpret_cast<uint32_t*>(ws + p.off_tilecnt);
  uint32_t* n_live = reinterpret_cast<uint32_t*>(ws + p.off_nlive);
  uint32_t* ka = reinterpret_cast<uint32_t*>(ws + p.off_keys_a);
  uint32_t* va = reinterpret_cast<uint32_t*>(ws + p.off_vals_a);
  if (stage == RECGURU_PLAN_SORT) {
    uint32_t *sk0, *sv0;
    return run_radix_sort(p, ws, total, n_live, false, st, &sk0, &sv0, nullptr, nullptr);
  }
  if (by_row) {
    uint32_t* rows = reinterpret_cast<uint32_t*>(ws + p.off_rows);
    const unsigned rtiles = static_cast<unsigned>(ceil_div(n_rows, kCompactTile));
    const float* lv = sites_host[0].live;
    rg::launch_pdl((row_count_kernel), rtiles, kSortThreads, 0, st, lv, n_rows, tile_cnt);
    RG_CHECK_LAUNCH("row_count");
    rg::launch_pdl((compact_scan_kernel), 1, 1024, 0, st, tile_cnt, static_cast<int>(rtiles), n_live + 1);
    RG_CHECK_LAUNCH("compact_scan");
    const uint32_t sentinel = static_cast<uint32_t>(V1);
    if (n_sites <= 1) rg::launch_pdl((row_scatter_kernel<1>), rtiles, kSortThreads, 0, st, tab, lv, n_rows, padding_idx, sentinel, tile_cnt, n_live + 1, n_live, ka, va, rows);
    else if (n_sites <= 3) rg::launch_pdl((row_scatter_kernel<3>), rtiles, kSortThreads, 0, st, tab, lv, n_rows, padding_idx, sentinel, tile_cnt, n_live + 1, n_live, ka, va, rows);
    else rg::launch_pdl((row_scatter_kernel<RECGURU_MAX_SITES>), rtiles, kSortThreads, 0, st, tab, lv, n_rows, padding_idx, sentinel, tile_cnt, n_live + 1, n_live, ka, va, rows);
    RG_CHECK_LAUNCH("row_scatter");
    if (stage == RECGURU_PLAN_COMPACT) return RECGURU_OK;
    uint32_t *sk2, *sv2;
    return run_radix_sort(p, ws, total, n_live, false, st, &sk2, &sv2, nullptr, nullptr);
  }
  const unsigned tiles = static_cast<unsigned>(ceil_div(total, kCompactTile));
  if (n_sites <= 1) rg::launch_pdl((compact_count_kernel<1>), tiles, kSortThreads, 0, st, tab, total, padding_idx, tile_cnt);
  else if (n_sites <= 3) rg::launch_pdl((compact_count_kernel<3>), tiles, kSortThreads, 0, st, tab, total, padding_idx, tile_cnt);
  else rg::launch_pdl((compact_count_kernel<RECGURU_MAX_SITES>), tiles, kSortThreads, 0, st, tab, total, padding_idx, tile_cnt);
  RG_CHECK_LAUNCH("compact_count");
  rg::launch_pdl((compact_scan_kernel), 1, 1024, 0, st, tile_cnt, static_cast<int>(tiles), n_live);
  RG_CHECK_LAUNCH("compact_scan");
  if (n_sites <= 1) rg::launch_pdl((compact_scatter_kernel<1>), tiles, kSortThreads, 0, st, tab, total, padding_idx, tile_cnt, ka, va);
  else if (n_sites <= 3) rg::launch_pdl((compact_scatter_kernel<3>), tiles, kSortThreads, 0, st, tab, total, padding_idx, tile_cnt, ka, va);
  else rg::launch_pdl((compact_scatter_kernel<RECGURU_MAX_SITES>), tiles, kSortThreads, 0, st, tab, total, padding_idx, tile_cnt, ka, va);
  RG_CHECK_LAUNCH("compact_scatter");
  if (stage == RECGURU_PLAN_COMPACT) return RECGURU_OK;
  uint32_t *sk, *sv;
  return run_radix_sort(p, ws, total, n_live, false, st, &sk, &sv, nullptr, nullptr);
}

// The fused plan: plan_count, plan_scatter, one radix_pass per digit (4 launches for <= 18-bit ids).
// mask_ids != NULL (row mode only): the shared liveness mask is derived here (mask_out, count3).
static int plan_v2(const recguru_lookup_site_t* sites_host, int n_sites, const SiteTable& tab, int64_t total,
                   int64_t V1, int64_t padding_idx, bool by_row, int64_t n_rows, const SortPlan& p, char* ws,
                   const int64_t* mask_ids, int64_t pad_index, float* mask_out, uint32_t* count3, cudaStream_t st) {
  uint32_t* tile_cnt = reinterpret_cast<uint32_t*>(ws + p.off_tilecnt);
  uint32_t* n_live = reinterpret_cast<uint32_t*>(ws + p.off_nlive);
  uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(ws + p.off_keys_a), reinterpret_cast<uint32_t*>(ws + p.off_keys_b)};
  uint32_t* vals[2] = {reinterpret_cast<uint32_t*>(ws + p.off_vals_a), reinterpret_cast<uint32_t*>(ws + p.off_vals_b)};
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + p.off_hist);
  const int64_t n_units = by_row ? n_rows : total;
  const unsigned ctiles = static_cast<unsigned>(ceil_div(n_units, kC2Tile));
  const int64_t rtiles = ceil_div(total, kP2Tile);                     // radix tiles (capacity)
  const int64_t hist_pass = p2_hist_words(rtiles);                     // words per pass
  const int64_t hist_words = hist_pass * p.passes;
  const float* lv = by_row ? (mask_ids ? mask_out : sites_host[0].live) : nullptr;
  const uint32_t sentinel = static_cast<uint32_t>(V1);
  const int bits0 = p.bits_per_pass < p.total_bits ? p.bits_per_pass : p.total_bits;
  const uint32_t mask0 = (1u << bits0) - 1u;
#define RG_PLAN_LAUNCH(NS)                                                                             \
  do {                                                                                                 \
    if (by_row) {                                                                                      \
      rg::launch_pdl((plan_count_kernel<NS, true>), ctiles, kP2Threads, 0, st, tab, lv, mask_ids, pad_index, \
                     mask_out, count3, n_units, padding_idx, tile_cnt, hist, hist_words);              \
      RG_CHECK_LAUNCH("plan_count");                                                                   \
      rg::launch_pdl((plan_scatter_kernel<NS, true>), ctiles, kP2Threads, 0, st, tab, lv, n_units, padding_idx, \
                     sentinel, tile_cnt, static_cast<int>(ctiles), n_live, keys[0], vals[0], hist, rtiles, mask0); \
      RG_CHECK_LAUNCH("plan_scatter");                                                                 \
    } else {                                                                                           \
      rg::launch_pdl((plan_count_kernel<NS, false>), ctiles, kP2Threads, 0, st, tab, lv, mask_ids, pad_index, \
                     mask_out, count3, n_units, padding_idx, tile_cnt, hist, hist_words);              \
      RG_CHECK_LAUNCH("plan_count");                                                                   \
      rg::launch_pdl((plan_scatter_kernel<NS, false>), ctiles, kP2Threads, 0, st, tab, lv, n_units, padding_idx, \
                     sentinel, tile_cnt, static_cast<int>(ctiles), n_live, keys[0], vals[0], hist, rtiles, mask0); \
      RG_CHECK_LAUNCH("plan_scatter");                                                                 \
    }                                                                                                  \
  } while (0)
  if (n_sites <= 1) RG_PLAN_LAUNCH(1);
  else if (n_sites <= 3) RG_PLAN_LAUNCH(3);
  else RG_PLAN_LAUNCH(RECGURU_MAX_SITES);
#undef RG_PLAN_LAUNCH
  int cur = 0;
  for (int pass = 0; pass < p.passes; ++pass) {
    const int shift = pass * p.bits_per_pass;
    const int bits = min(p.bits_per_pass, p.total_bits - shift);
    const bool last = pass == p.passes - 1;
    const int nshift = shift + p.bits_per_pass;
    const int nbits = last ? 0 : min(p.bits_per_pass, p.total_bits - nshift);
    rg::launch_pdl((radix_pass_kernel), static_cast<unsigned>(rtiles), kP2Threads, 0, st, keys[cur], vals[cur],
                   n_live, shift, bits, hist + pass * hist_pass, last ? nullptr : hist + (pass + 1) * hist_pass,
                   rtiles, nshift, last ? 0u : (1u << nbits) - 1u, keys[cur ^ 1], vals[cur ^ 1]);
    RG_CHECK_LAUNCH("radix_pass");
    cur ^= 1;
  }
  return RECGURU_OK;
}

static int plan_impl(const char* who, const recguru_lookup_site_t* sites_host, int n_sites, int64_t V1,
                     int64_t padding_idx, void* workspace, int64_t workspace_bytes, const int64_t* mask_ids,
                     int64_t pad_index, float* mask_out, uint32_t* count3, int stage, recguru_stream_t stream) {
  RG_REQUIRE(V1 > 0 && V1 < (1ll << 31), RECGURU_E_SHAPE, "%s: bad V1", who);
  RG_REQUIRE(stage >= RECGURU_PLAN_ALL && stage <= RECGURU_PLAN_SORT, RECGURU_E_SHAPE, "%s: stage=%d", who, stage);
  SiteTable tab;
  int64_t total = 0;
  int rc = build_site_table(who, sites_host, n_sites, false, &tab, &total);
  if (rc) return rc;
  if (total == 0) return RECGURU_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (stage != RECGURU_PLAN_SORT)
    for (int s = 0; s < n_sites; ++s)
      RG_CHECK_IDS("embedding_bwd_plan (site ids)", sites_host[s].ids, sites_host[s].count, 0, V1, st);
  int64_t n_rows = 0;
  const bool by_row = row_mode(sites_host, n_sites, &n_rows);
  RG_REQUIRE(mask_ids == nullptr || (by_row && mask_out != nullptr && sites_host[0].live == mask_out),
             RECGURU_E_SHAPE, "%s: deriving the mask here needs every site's live == mask_out (one entry per row)", who);
  // by_row: one more key value (the sentinel V1) must sort correctly
  const SortPlan p = make_sort_plan(total, by_row ? V1 + 1 : V1);
  RG_REQUIRE(workspace && workspace_bytes >= p.end, RECGURU_E_WORKSPACE,
             "%s: workspace %lld < %lld", who, (long long)workspace_bytes, (long long)p.end);
  char* ws = static_cast<char*>(workspace);
  if (env_plan_v1()) {
    if (mask_ids != nullptr && stage != RECGURU_PLAN_SORT) {
      rc = count3 ? recguru_pad_mask_count(mask_ids, n_rows, pad_index, mask_out, count3, stream)
                  : recguru_pad_mask(mask_ids, n_rows, pad_index, mask_out, stream);
      if (rc) return rc;
    }
    return plan_v1(sites_host, n_sites, tab, total, V1, padding_idx, by_row, n_rows, p, ws, stage, st);
  }
  RG_REQUIRE(stage == RECGURU_PLAN_ALL, RECGURU_E_UNSUPPORTED,
             "%s: the two-stage plan needs the default plan (unset RECGURU_BWD_PLAN)", who);
  return plan_v2(sites_host, n_sites, tab, total, V1, padding_idx, by_row, n_rows, p, ws, mask_ids, pad_index,
                 mask_out, count3, st);
}
}  // namespace rg

// Stage A of the table backward: stable compaction of the live (id, occurrence) pairs and the
// LSD radix sort.  Reads ids and the liveness source only (site.live, else site.coef).
extern "C" int recguru_embedding_bwd_plan(const recguru_lookup_site_t* sites_host, int n_sites,
                                          int64_t V1, int64_t padding_idx, void* workspace,
                                          int64_t workspace_bytes, recguru_stream_t stream) {
  return rg::plan_impl("embedding_bwd_plan", sites_host, n_sites, V1, padding_idx, workspace, workspace_bytes,
                       nullptr, 0, nullptr, nullptr, RECGURU_PLAN_ALL, stream);
}

// Stage A in two calls, so that a caller can hang an event between them: RECGURU_PLAN_COMPACT runs
// the stable compaction (in row mode it also leaves the ascending list of live rows, see
// recguru_embedding_bwd_plan_live_rows), RECGURU_PLAN_SORT the radix sort over it.  COMPACT then
// SORT on one stream == RECGURU_PLAN_ALL == recguru_embedding_bwd_plan.
extern "C" int recguru_embedding_bwd_plan_stage(const recguru_lookup_site_t* sites_host, int n_sites,
                                                int64_t V1, int64_t padding_idx, void* workspace,
                                                int64_t workspace_bytes, int stage, recguru_stream_t stream) {
  return rg::plan_impl("embedding_bwd_plan_stage", sites_host, n_sites, V1, padding_idx, workspace,
                       workspace_bytes, nullptr, 0, nullptr, nullptr, stage, stream);
}

// Where the row-mode compaction leaves the live rows: *live_rows -> ascending row indices,
// *n_live_rows -> their count (both device pointers into `workspace`, valid once the compaction of
// the same sites has run on the stream).  RECGURU_E_UNSUPPORTED when the sites are not in row mode.
extern "C" int recguru_embedding_bwd_plan_live_rows(const recguru_lookup_site_t* sites_host, int n_sites,
                                                    int64_t V1, void* workspace, int64_t workspace_bytes,
                                                    const uint32_t** live_rows, const uint32_t** n_live_rows) {
  using namespace rg;
  RG_REQUIRE(live_rows && n_live_rows, RECGURU_E_SHAPE, "embedding_bwd_plan_live_rows: null pointer");
  RG_REQUIRE(V1 > 0 && V1 < (1ll << 31), RECGURU_E_SHAPE, "embedding_bwd_plan_live_rows: bad V1");
  SiteTable tab;
  int64_t total = 0;
  int rc = build_site_table("embedding_bwd_plan_live_rows", sites_host, n_sites, false, &tab, &total);
  if (rc) return rc;
  int64_t n_rows = 0;
  RG_REQUIRE(total > 0 && row_mode(sites_host, n_sites, &n_rows) && env_plan_v1(), RECGURU_E_UNSUPPORTED,
             "embedding_bwd_plan_live_rows: the sites do not share one per-row liveness mask");
  const SortPlan p = make_sort_plan(total, V1 + 1);
  RG_REQUIRE(workspace && workspace_bytes >= p.end, RECGURU_E_WORKSPACE,
             "embedding_bwd_plan_live_rows: workspace %lld < %lld", (long long)workspace_bytes, (long long)p.end);
  char* ws = static_cast<char*>(workspace);
  *live_rows = reinterpret_cast<const uint32_t*>(ws + p.off_rows);
  *n_live_rows = reinterpret_cast<const uint32_t*>(ws + p.off_nlive) + 1;
  return RECGURU_OK;
}

// Stage A with the shared pad mask derived inside: mask_out[i] = mask_ids[i] != pad_index (one entry
// per row; every site's `live` must point at mask_out), count3 as in recguru_pad_mask_count (may be
// NULL).  One launch fewer than recguru_pad_mask_count + recguru_embedding_bwd_plan.
extern "C" int recguru_embedding_bwd_plan_padmask(const recguru_lookup_site_t* sites_host, int n_sites,
                                                  int64_t V1, int64_t padding_idx, const int64_t* mask_ids,
                                                  int64_t pad_index, float* mask_out, uint32_t* count3,
                                                  void* workspace, int64_t workspace_bytes,
                                                  recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(mask_ids != nullptr && mask_out != nullptr, RECGURU_E_SHAPE, "embedding_bwd_plan_padmask: null pointer");
  return rg::plan_impl("embedding_bwd_plan_padmask", sites_host, n_sites, V1, padding_idx, workspace,
                       workspace_bytes, mask_ids, pad_index, mask_out, count3, RECGURU_PLAN_ALL, stream);
}

extern "C" int64_t recguru_live_rows_workspace_bytes(int64_t n_rows) {
  if (n_rows < 0) return RECGURU_E_SHAPE;
  return rg::align_up(rg::ceil_div(n_rows > 0 ? n_rows : 1, rg::kCompactTile) * 4, 256);
}

// live_rows[0 .. *n_live_rows) = ascending indices i with mask[i] != 0 (three launches: tile counts, one-CTA
// scan, stable scatter).  Same list as recguru_embedding_bwd_plan_live_rows, without a plan.
extern "C" int recguru_live_rows(const float* mask, int64_t n_rows, uint32_t* live_rows, uint32_t* n_live_rows,
                                 void* workspace, int64_t workspace_bytes, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(n_rows >= 0 && n_rows < (1ll << 32), RECGURU_E_SHAPE, "live_rows: bad n_rows");
  RG_REQUIRE(n_live_rows != nullptr, RECGURU_E_SHAPE, "live_rows: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_rows == 0) {
    const cudaError_t e = cudaMemsetAsync(n_live_rows, 0, 4, st);
    RG_REQUIRE(e == cudaSuccess, RECGURU_E_CUDA, "live_rows: %s", cudaGetErrorString(e));
    return RECGURU_OK;
  }
  RG_REQUIRE(mask && live_rows, RECGURU_E_SHAPE, "live_rows: null pointer");
  const int64_t need = recguru_live_rows_workspace_bytes(n_rows);
  RG_REQUIRE(workspace && workspace_bytes >= need, RECGURU_E_WORKSPACE, "live_rows: workspace %lld < %lld",
             (long long)workspace_bytes, (long long)need);
  uint32_t* tile_cnt = static_cast<uint32_t*>(workspace);
  const unsigned tiles = static_cast<unsigned>(ceil_div(n_rows, kCompactTile));
  rg::launch_pdl((row_count_kernel), tiles, kSortThreads, 0, st, mask, n_rows, tile_cnt);
  RG_CHECK_LAUNCH("row_count");
  rg::launch_pdl((compact_scan_kernel), 1, 1024, 0, st, tile_cnt, static_cast<int>(tiles), n_live_rows);
  RG_CHECK_LAUNCH("compact_scan");
  rg::launch_pdl((rows_only_scatter_kernel), tiles, kSortThreads, 0, st, mask, n_rows, tile_cnt, live_rows);
  RG_CHECK_LAUNCH("rows_only_scatter");
  return RECGURU_OK;
}

// Stage B: (optional zero fill,) chunked segmented reduce + span fix-up over the plan that
// recguru_embedding_bwd_plan left in `workspace` for the SAME sites (ids, counts, groups).
extern "C" int recguru_embedding_bwd_apply(const recguru_lookup_site_t* sites_host, int n_sites,
                                           int64_t d, int64_t V1, float* dW, int zero_fill,
                                           void* workspace, int64_t workspace_bytes,
                                           recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(dW && aligned16(dW), RECGURU_E_SHAPE, "embedding_bwd: dW null or unaligned");
  RG_REQUIRE(V1 > 0 && V1 < (1ll << 31), RECGURU_E_SHAPE, "embedding_bwd: bad V1");
  RG_REQUIRE(d > 0 && d % 4 == 0 && d <= 512, RECGURU_E_UNSUPPORTED,
             "embedding_bwd: d=%lld must be a multiple of 4 and <= 512", (long long)d);
  SiteTable tab;
  int64_t total = 0;
  int rc = build_site_table("embedding_bwd", sites_host, n_sites, true, &tab, &total);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec = static_cast<int>(d / 4);
  RG_REQUIRE(zero_fill >= 0 && zero_fill <= RECGURU_DW_PREZEROED, RECGURU_E_SHAPE,
             "embedding_bwd: zero_fill=%d", zero_fill);
  if (zero_fill == RECGURU_DW_ZERO_FILL) {
    rc = recguru_fill_zero(dW, V1 * d, stream);
    if (rc) return rc;
  }
  if (total == 0) return RECGURU_OK;
  int64_t n_rows = 0;
  const bool by_row = row_mode(sites_host, n_sites, &n_rows);   // the same decision the plan took
  const int64_t pad_key = by_row ? V1 : -1;                       // sentinel key of dropped occurrences
  const SortPlan p = make_sort_plan(total, by_row ? V1 + 1 : V1);
  const int64_t n_chunks = ceil_div(total, kChunk);
  const int64_t need = p.end + align_up(n_chunks * 2 * d * 4, 256);
  RG_REQUIRE(workspace && workspace_bytes >= need, RECGURU_E_WORKSPACE,
             "embedding_bwd: workspace %lld < %lld", (long long)workspace_bytes, (long long)need);
  char* ws = static_cast<char*>(workspace);
  const uint32_t* n_live = reinterpret_cast<const uint32_t*>(ws + p.off_nlive);
  // the sort ping-pongs between the two key/value arrays: after `passes` passes the result
  // sits in array (passes & 1)
  const uint32_t* sk = reinterpret_cast<const uint32_t*>(ws + ((p.passes & 1) ? p.off_keys_b : p.off_keys_a));
  const uint32_t* sv = reinterpret_cast<const uint32_t*>(ws + ((p.passes & 1) ? p.off_vals_b : p.off_vals_a));
  float4* partial = reinterpret_cast<float4*>(ws + p.end);
  float4* dW4 = reinterpret_cast<float4*>(dW);
  const int accumulate = zero_fill == RECGURU_DW_ACCUMULATE ? 1 : 0;
  const unsigned grid = static_cast<unsigned>(ceil_div(n_chunks, kReduceWarps));
  const int vpl = (vec + 31) / 32;
  // RECGURU_REDUCE=regs: the register-staged chunk_reduce_kernel (A/B timing); default: TMA-staged
  static int env_reduce_regs = -1;
  if (env_reduce_regs < 0) {
    const char* e = getenv("RECGURU_REDUCE");
    env_reduce_regs = (e && e[0] == 'r') ? 1 : 0;
  }
  const bool reduce_tma = env_reduce_regs == 0;
  const int group_rows = static_cast<int>(rg::min64(32, kCrStageBytes / (vec * 16)));
  const unsigned grid_tma = static_cast<unsigned>(ceil_div(n_chunks, kCrWarps));
#define RG_LAUNCH_REDUCE_TMA(V, NSV)                                                          \
  do {                                                                                        \
    static rg::PerDeviceOnce once;                                                            \
    const int dev = rg::current_device();                                                     \
    if (once.need(dev)) {                                                                     \
      const cudaError_t ea = cudaFuncSetAttribute(chunk_reduce_tma_kernel<V, NSV>,            \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, kCrWarps * kCrStageBytes); \
      RG_REQUIRE(ea == cudaSuccess, RECGURU_E_CUDA, "embedding_bwd: cudaFuncSetAttribute: %s", \
                 cudaGetErrorString(ea));                                                     \
      once.done(dev);                                                                         \
    }                                                                                         \
    rg::launch_pdl((chunk_reduce_tma_kernel<V, NSV>), grid_tma, kCrWarps * 32, kCrWarps * kCrStageBytes, st, tab, sk, \
                   sv, n_live, vec, group_rows, pad_key, accumulate, dW4, partial);            \
  } while (0)
#define RG_LAUNCH_REDUCE(V)                                                                   \
  do {                                                                                        \
    if (reduce_tma) {                                                                         \
      if (n_sites <= 1) RG_LAUNCH_REDUCE_TMA(V, 1);                                           \
      else if (n_sites <= 3) RG_LAUNCH_REDUCE_TMA(V, 3);                                      \
      else RG_LAUNCH_REDUCE_TMA(V, RECGURU_MAX_SITES);                                        \
    } else if (n_sites <= 1)                                                                  \
      rg::launch_pdl((chunk_reduce_kernel<V, 1>), grid, kReduceWarps * 32, 0, st, tab, sk, sv, n_live, vec, pad_key, \
                                                                    accumulate, dW4, partial);    \
    else if (n_sites <= 3)                                                                    \
      rg::launch_pdl((chunk_reduce_kernel<V, 3>), grid, kReduceWarps * 32, 0, st, tab, sk, sv, n_live, vec, pad_key, \
                                                                    accumulate, dW4, partial);    \
    else                                                                                      \
      rg::launch_pdl((chunk_reduce_kernel<V, RECGURU_MAX_SITES>), grid, kReduceWarps * 32, 0, st,           \
          tab, sk, sv, n_live, vec, pad_key, accumulate, dW4, partial);                            \
    RG_CHECK_LAUNCH("chunk_reduce");                                                          \
    if (n_chunks > 1) {                                                                       \
      const unsigned g2 = static_cast<unsigned>(ceil_div(n_chunks - 1, kReduceWarps));        \
      rg::launch_pdl((span_fixup_kernel<V>), g2, kReduceWarps * 32, 0, st, sk, n_live, vec,                 \
                                                             pad_key, accumulate, dW4,        \
                                                             partial);                        \
      RG_CHECK_LAUNCH("span_fixup");                                                          \
    }                                                                                         \
  } while (0)
  switch (vpl) {
    case 1: RG_LAUNCH_REDUCE(1); break;
    case 2: RG_LAUNCH_REDUCE(2); break;
    case 3: RG_LAUNCH_REDUCE(3); break;
    case 4: RG_LAUNCH_REDUCE(4); break;
    default:
      set_error("embedding_bwd: d=%lld unsupported", (long long)d);
      return RECGURU_E_UNSUPPORTED;
  }
#undef RG_LAUNCH_REDUCE
#undef RG_LAUNCH_REDUCE_TMA
  return RECGURU_OK;
}

// Row liveness for the sparse-gradient consumers (recguru_table_adam_rows): every table row that
// owns a run of the sorted plan -- exactly the rows chunk_reduce / span_fixup write -- gets
// row_epoch[row] = epoch.  One thread per sorted occurrence; run starts only.
namespace rg {
__global__ void __launch_bounds__(256)
mark_rows_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ n_live, uint32_t key_limit,
                 uint32_t epoch, uint32_t* __restrict__ row_epoch) {
  rg::pdl_wait();
  const int64_t total = *n_live;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint32_t k = skeys[i];
    if (k < key_limit && (i == 0 || skeys[i - 1] != k)) row_epoch[k] = epoch;
  }
  rg::pdl_trigger();
}
}  // namespace rg

extern "C" int recguru_embedding_bwd_mark_rows(const recguru_lookup_site_t* sites_host, int n_sites, int64_t V1,
                                               const void* workspace, int64_t workspace_bytes,
                                               uint32_t* row_epoch, uint32_t epoch, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(V1 > 0 && V1 < (1ll << 31), RECGURU_E_SHAPE, "embedding_bwd_mark_rows: bad V1");
  RG_REQUIRE(row_epoch != nullptr, RECGURU_E_SHAPE, "embedding_bwd_mark_rows: null row_epoch");
  SiteTable tab;
  int64_t total = 0;
  int rc = build_site_table("embedding_bwd_mark_rows", sites_host, n_sites, false, &tab, &total);
  if (rc) return rc;
  if (total == 0) return RECGURU_OK;
  int64_t n_rows = 0;
  const bool by_row = row_mode(sites_host, n_sites, &n_rows);   // the same decision the plan took
  const SortPlan p = make_sort_plan(total, by_row ? V1 + 1 : V1);
  RG_REQUIRE(workspace && workspace_bytes >= p.end, RECGURU_E_WORKSPACE,
             "embedding_bwd_mark_rows: workspace %lld < %lld", (long long)workspace_bytes, (long long)p.end);
  const char* ws = static_cast<const char*>(workspace);
  const uint32_t* n_live = reinterpret_cast<const uint32_t*>(ws + p.off_nlive);
  const uint32_t* sk = reinterpret_cast<const uint32_t*>(ws + ((p.passes & 1) ? p.off_keys_b : p.off_keys_a));
  const unsigned blocks = static_cast<unsigned>(rg::min64(ceil_div(total, 256), static_cast<int64_t>(num_sms()) * 8));
  // keys >= V1 are the by-row sentinel of dropped occurrences (padding ids at live rows)
  rg::launch_pdl((mark_rows_kernel), blocks, 256, 0, static_cast<cudaStream_t>(stream), sk, n_live,
                 static_cast<uint32_t>(V1), epoch, row_epoch);
  RG_CHECK_LAUNCH("mark_rows");
  return RECGURU_OK;
}

extern "C" int recguru_embedding_bwd(const recguru_lookup_site_t* sites_host, int n_sites,
                                     int64_t d, int64_t V1, int64_t padding_idx, float* dW,
                                     int zero_fill, void* workspace, int64_t workspace_bytes,
                                     recguru_stream_t stream) {
  int rc = recguru_embedding_bwd_plan(sites_host, n_sites, V1, padding_idx, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return recguru_embedding_bwd_apply(sites_host, n_sites, d, V1, dW, zero_fill, workspace, workspace_bytes,
                                     stream);
}
