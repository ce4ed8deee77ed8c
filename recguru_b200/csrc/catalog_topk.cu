// Piece 4 (K4): full-catalog scoring as a TMA-fed tcgen05/TMEM GEMM fused with a running
// per-user top-K and the target's rank.  The [users x items] logit matrix is never written.
//
// Replaces (scaled up to the whole catalog)
//   torch.matmul(dec_inputs, self.src_emb.weight.transpose(1, 0))       AutoEnc4Rec.py:230
//   torch.argsort(torch.argsort(-score, dim=1), dim=1)[:, 0]            train_auto.py:234-237,
//                                                       gan_training.py:130-133, tools/utils.py:168-169
//   torch.topk(prediction_i, top_k)                                     tools/metrics.py:93
//
// Shape of the work: S[u, j] = sum_k U[u,k] * W[j,k]  with K = d in {64,128} -- a GEMM whose
// reduction dimension is a single short K loop.  Per 128x128 output tile the tensor core
// needs d/16 MMAs (64 cycles each), so the epilogue, not the MMA, is the thing to engineer:
// every score has to be looked at once (rank compare + top-K threshold).
//
// CTA = 128 users (one M tile) sweeping a slice of the catalog in 128-item N tiles.
//   warp 0      TMA producer: user tile once, then item tiles through a kStages smem ring
//   warp 1      MMA issuer (one elected lane) + TMEM allocation (4 x 128 fp32 columns = 512)
//   warps 2-5   epilogue set 0: even tiles      each set = 4 warps = the 4 TMEM lane quarters;
//   warps 6-9   epilogue set 1: odd tiles       thread <-> user row, registers <-> item columns
// Each epilogue thread keeps for its user: a rank counter (register), the current K-th best
// score (register) and a sorted top-K list in shared memory.  Fast path per 32 columns:
// 32 compares against the target score and one max-reduction against the K-th best; the list
// is only touched when the max beats the threshold (expected K*ln(N/K) times per user).
// Ties: lower id wins -- for the rank, tiles left of the target's tile compare with >=
// (implemented as > against the next float below), tiles right of it with >; the tile that
// holds the target takes an exact per-column path.
#include <cuda_bf16.h>
#include <limits.h>
#include <math_constants.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace rg {

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kStages = 4;       // smem ring of item tiles
constexpr int kAccum = 4;        // TMEM accumulator buffers (4 x 128 columns)
constexpr int kCatThreads = 320; // 10 warps
constexpr int kEpiThreads = 256;
constexpr int kMaxK = 32;

enum CatalogMode { kModeTopK = 0, kModeDense = 1 };

struct CatalogParams {
  int64_t U;
  int64_t n_items;       // rows of the item slice on this device
  int64_t id_lo;         // catalog id of slice row 0
  const int64_t* target; // [U]
  const float* tscore;   // [U]
  int k;
  int n_splits;          // grid.y
  int tiles_per_split;
  float* out_val;        // [n_splits, U, k]
  int64_t* out_id;       // [n_splits, U, k]
  int64_t* out_cnt;      // [n_splits, U]
  float* dense;          // kModeDense: [U, n_items]
};

__device__ __forceinline__ float next_below(float t) {
  const uint32_t b = __float_as_uint(t);
  if (t > 0.f) return __uint_as_float(b - 1u);
  if (t == 0.f) return __uint_as_float(0x80000001u);
  return __uint_as_float(b + 1u);
}

// Shared-memory carve-up (dynamic, base aligned to 1024 for the 128-byte swizzle).
template <int D>
struct CatSmem {
  static constexpr int kKBlocks = D / 64;                 // 64-wide (128 B) K blocks
  static constexpr int kTileBytes = kBlockN * D * 2;      // one operand tile
  static constexpr int kOffA = 0;
  static constexpr int kOffB = kTileBytes;
  static constexpr int kOffBar = kOffB + kStages * kTileBytes;
  // barriers: full[kStages], empty[kStages], acc_full[kAccum], acc_empty[kAccum], a_full
  static constexpr int kNumBars = 2 * kStages + 2 * kAccum + 1;
  static constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr int kOffCnt = kOffTmemPtr + 8;         // int[128] (set-1 rank counts)
  static constexpr int kOffList = kOffCnt + 128 * 4;      // float[2*k*128] then int[2*k*128]
  static constexpr int bytes(int k) { return kOffList + 2 * k * 128 * 8 + 1024; }
};

template <int D, int MODE>
__global__ void __launch_bounds__(kCatThreads, 1)
catalog_sweep_kernel(const __grid_constant__ CUtensorMap map_users,
                     const __grid_constant__ CUtensorMap map_items, CatalogParams p) {
  using L = CatSmem<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + kStages;
  uint64_t* bar_acc_full = bars + 2 * kStages;
  uint64_t* bar_acc_empty = bars + 2 * kStages + kAccum;
  uint64_t* bar_a = bars + 2 * kStages + 2 * kAccum;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kOffTmemPtr);
  int* s_cnt = reinterpret_cast<int*>(smem + L::kOffCnt);
  float* s_val = reinterpret_cast<float*>(smem + L::kOffList);
  int* s_id = reinterpret_cast<int*>(smem + L::kOffList + 2 * p.k * 128 * 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kBlockM;
  const int total_tiles = static_cast<int>((p.n_items + kBlockN - 1) / kBlockN);
  const int tile_begin = blockIdx.y * p.tiles_per_split;
  const int tile_end = min(total_tiles, tile_begin + p.tiles_per_split);
  const int n_tiles = max(0, tile_end - tile_begin);

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&map_users);
    ptx::tma_prefetch_desc(&map_items);
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&bar_full[i], 1); ptx::mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < kAccum; ++i) { ptx::mbar_init(&bar_acc_full[i], 1); ptx::mbar_init(&bar_acc_empty[i], 4); }
    ptx::mbar_init(bar_a, 1);
    ptx::mbar_fence_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar_a, L::kTileBytes);
      for (int kb = 0; kb < L::kKBlocks; ++kb)
        ptx::tma_load_2d(&map_users, bar_a, smem + L::kOffA + kb * (kBlockM * 128), kb * 64,
                         static_cast<int32_t>(m0));
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % kStages;
        const uint32_t ph = (j / kStages) & 1;
        ptx::mbar_wait(&bar_empty[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&bar_full[s], L::kTileBytes);
        uint8_t* dst = smem + L::kOffB + s * L::kTileBytes;
        const int32_t n0 = (tile_begin + j) * kBlockN;
        for (int kb = 0; kb < L::kKBlocks; ++kb)
          ptx::tma_load_2d(&map_items, &bar_full[s], dst + kb * (kBlockN * 128), kb * 64, n0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(kBlockM, kBlockN);
      const uint32_t a_addr = ptx::smem_u32(smem + L::kOffA);
      ptx::mbar_wait(bar_a, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % kStages;
        const uint32_t ph = (j / kStages) & 1;
        const int a = j % kAccum;
        const uint32_t aph = (j / kAccum) & 1;
        ptx::mbar_wait(&bar_acc_empty[a], aph ^ 1);
        ptx::mbar_wait(&bar_full[s], ph);
        ptx::tc_fence_after();
        const uint32_t b_addr = ptx::smem_u32(smem + L::kOffB + s * L::kTileBytes);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t koff = (k / 4) * (kBlockM * 128) + (k % 4) * 32;  // bytes
          ptx::umma_bf16(tmem_base + a * kBlockN, ptx::umma_smem_desc_sw128(a_addr + koff),
                         ptx::umma_smem_desc_sw128(b_addr + koff), idesc, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit(&bar_empty[s]);     // smem stage free once these MMAs retire
        ptx::umma_commit(&bar_acc_full[a]);  // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue =====================
    const int e = (warp - 2) >> 2;            // set: 0 even tiles, 1 odd tiles
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int64_t u = m0 + row;
    const bool u_ok = u < p.U;
    const int K = p.k;
    const int slot = e * 128 + row;           // list slot; entry i at [i * 256 + slot]

    float t = 0.f;
    int64_t tgt_local = -1;                   // target's row index inside the slice
    if (MODE == kModeTopK && u_ok) {
      t = p.tscore[u];
      tgt_local = p.target[u] - p.id_lo;
    }
    const float t_below = next_below(t);
    int tgt_tile, tgt_col = -1;               // tile relative to this device's slice
    if (tgt_local < 0) tgt_tile = -1;
    else if (tgt_local >= p.n_items) tgt_tile = INT_MAX;
    else { tgt_tile = static_cast<int>(tgt_local / kBlockN); tgt_col = static_cast<int>(tgt_local % kBlockN); }

    int cnt = 0;
    float thr = -CUDART_INF_F;
    if (MODE == kModeTopK) {
      for (int i = 0; i < K; ++i) { s_val[i * 256 + slot] = -CUDART_INF_F; s_id[i * 256 + slot] = -1; }
    }
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;

    for (int j = e; j < n_tiles; j += 2) {
      const int a = j % kAccum;
      const uint32_t aph = (j / kAccum) & 1;
      const int tile = tile_begin + j;
      const int n0 = tile * kBlockN;                                   // slice-local item index
      const int valid = static_cast<int>(rg::min64(kBlockN, p.n_items - n0));
      ptx::mbar_wait(&bar_acc_full[a], aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + a * kBlockN;
      const bool exact = (tile == tgt_tile) || (valid < kBlockN);
      const float tcmp = tile < tgt_tile ? t_below : t;

#pragma unroll 1
      for (int c0 = 0; c0 < kBlockN; c0 += 32) {
        uint32_t vr[32];
        ptx::tmem_ld_32x32(taddr + c0, vr);
        ptx::tmem_ld_wait();
        if (MODE == kModeDense) {
          if (u_ok) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c0 + c < valid) p.dense[u * p.n_items + n0 + c0 + c] = __uint_as_float(vr[c]);
          }
          continue;
        }
        if (!exact) {
          float mx = __uint_as_float(vr[0]);
          int add = 0;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float v = __uint_as_float(vr[c]);
            add += (v > tcmp) ? 1 : 0;
            mx = fmaxf(mx, v);
          }
          cnt += add;
          if (mx > thr) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float v = __uint_as_float(vr[c]);
              if (v > thr) {
                int pos = K - 1;
                while (pos > 0 && s_val[(pos - 1) * 256 + slot] < v) {
                  s_val[pos * 256 + slot] = s_val[(pos - 1) * 256 + slot];
                  s_id[pos * 256 + slot] = s_id[(pos - 1) * 256 + slot];
                  --pos;
                }
                s_val[pos * 256 + slot] = v;
                s_id[pos * 256 + slot] = n0 + c0 + c;
                thr = s_val[(K - 1) * 256 + slot];
              }
            }
          }
        } else {
          // exact path: tile holds the target and/or runs past the end of the slice
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int col = c0 + c;
            const float v = __uint_as_float(vr[c]);
            if (col < valid) {
              if (tile == tgt_tile) {
                if (col < tgt_col) cnt += (v >= t) ? 1 : 0;
                else if (col > tgt_col) cnt += (v > t) ? 1 : 0;
              } else {
                cnt += (v > tcmp) ? 1 : 0;
              }
              if (v > thr) {
                int pos = K - 1;
                while (pos > 0 && s_val[(pos - 1) * 256 + slot] < v) {
                  s_val[pos * 256 + slot] = s_val[(pos - 1) * 256 + slot];
                  s_id[pos * 256 + slot] = s_id[(pos - 1) * 256 + slot];
                  --pos;
                }
                s_val[pos * 256 + slot] = v;
                s_id[pos * 256 + slot] = n0 + col;
                thr = s_val[(K - 1) * 256 + slot];
              }
            }
          }
        }
      }
      // this warp is done reading accumulator `a`
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar_acc_empty[a]);
    }

    if (MODE == kModeTopK) {
      if (e == 1) s_cnt[row] = cnt;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (e == 0 && u_ok) {
        // merge the two sets' lists: (score desc, id asc)
        const int slot1 = 128 + row;
        int i0 = 0, i1 = 0;
        const int64_t out = (static_cast<int64_t>(blockIdx.y) * p.U + u) * K;
        for (int i = 0; i < K; ++i) {
          const float v0 = i0 < K ? s_val[i0 * 256 + slot] : -CUDART_INF_F;
          const float v1 = i1 < K ? s_val[i1 * 256 + slot1] : -CUDART_INF_F;
          const int d0 = i0 < K ? s_id[i0 * 256 + slot] : -1;
          const int d1 = i1 < K ? s_id[i1 * 256 + slot1] : -1;
          bool take0;
          if (d1 < 0) take0 = true;
          else if (d0 < 0) take0 = false;
          else take0 = (v0 > v1) || (v0 == v1 && d0 < d1);
          const float v = take0 ? v0 : v1;
          const int d = take0 ? d0 : d1;
          if (take0) ++i0; else ++i1;
          p.out_val[out + i] = d >= 0 ? v : -CUDART_INF_F;
          p.out_id[out + i] = d >= 0 ? p.id_lo + d : -1;
        }
        p.out_cnt[static_cast<int64_t>(blockIdx.y) * p.U + u] = cnt + s_cnt[row];
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// ---- target scores through the same tensor-core path -----------------------------------
// One CTA = 128 users.  The 128 target rows are gathered into a swizzled B tile by hand,
// one 128x128 MMA is issued and thread r keeps D[r][r].
template <int D>
__global__ void __launch_bounds__(192, 1)
catalog_target_kernel(const __grid_constant__ CUtensorMap map_users,
                      const __nv_bfloat16* __restrict__ items, int64_t id_lo, int64_t n_items,
                      const int64_t* __restrict__ target, int64_t U, float* __restrict__ tscore) {
  constexpr int kKBlocks = D / 64;
  constexpr int kTileBytes = kBlockN * D * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTileBytes;
  uint64_t* bar_a = reinterpret_cast<uint64_t*>(smem + 2 * kTileBytes);
  uint64_t* bar_mma = bar_a + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_a + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kBlockM;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&map_users);
    ptx::mbar_init(bar_a, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::mbar_fence_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, 128);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && lane == 0) {
    ptx::mbar_arrive_expect_tx(bar_a, kTileBytes);
    for (int kb = 0; kb < kKBlocks; ++kb)
      ptx::tma_load_2d(&map_users, bar_a, sA + kb * (kBlockM * 128), kb * 64,
                       static_cast<int32_t>(m0));
  }
  // gather target rows: row r, 16-byte chunk c of K block kb lands at
  //   kb*16KB + r*128 + ((c ^ (r & 7)) << 4)        (128-byte swizzle)
  constexpr int kChunksPerRow = D / 8;  // 16-byte chunks per row
  for (int idx = threadIdx.x; idx < kBlockN * kChunksPerRow; idx += blockDim.x) {
    const int r = idx / kChunksPerRow, c = idx % kChunksPerRow;
    const int64_t uu = m0 + r;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (uu < U) {
      const int64_t loc = target[uu] - id_lo;
      if (loc >= 0 && loc < n_items)
        val = __ldg(reinterpret_cast<const uint4*>(items + loc * D) + c);
    }
    const int kb = c / 8, cc = c % 8;
    *reinterpret_cast<uint4*>(sB + kb * (kBlockN * 128) + r * 128 + ((cc ^ (r & 7)) << 4)) = val;
  }
  ptx::fence_proxy_async_smem();
  __syncthreads();

  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(kBlockM, kBlockN);
    ptx::mbar_wait(bar_a, 0);
    ptx::tc_fence_after();
    const uint32_t a_addr = ptx::smem_u32(sA), b_addr = ptx::smem_u32(sB);
#pragma unroll
    for (int k = 0; k < D / 16; ++k) {
      const uint32_t koff = (k / 4) * (kBlockM * 128) + (k % 4) * 32;
      ptx::umma_bf16(tmem_base, ptx::umma_smem_desc_sw128(a_addr + koff),
                     ptx::umma_smem_desc_sw128(b_addr + koff), idesc, k > 0 ? 1u : 0u);
    }
    ptx::umma_commit(bar_mma);
  }
  if (warp >= 2) {
    const int quarter = warp & 3;  // warps 2,3,4,5 -> quarters 2,3,0,1
    ptx::mbar_wait(bar_mma, 0);
    ptx::tc_fence_after();
    uint32_t vr[32];
    ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + quarter * 32, vr);
    ptx::tmem_ld_wait();
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c == lane) v = __uint_as_float(vr[c]);
    const int64_t uu = m0 + quarter * 32 + lane;
    if (uu < U) tscore[uu] = v;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 128);
}

// ---- host side -----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// [rows, d] bf16 row-major -> box {64 columns, 128 rows}, 128-byte swizzle, OOB rows = 0
static int make_rowmajor_map(CUtensorMap* map, const void* base, int64_t rows, int64_t d) {
  EncodeTiledFn enc = get_encoder();
  RG_REQUIRE(enc != nullptr, RECGURU_E_CUDA, "catalog: cuTensorMapEncodeTiled unavailable");
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d) * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                         gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RG_REQUIRE(r == CUDA_SUCCESS, RECGURU_E_CUDA, "catalog: cuTensorMapEncodeTiled failed (%d)",
             static_cast<int>(r));
  return RECGURU_OK;
}

static int pick_splits(int64_t U, int64_t n_items) {
  const int64_t m_tiles = ceil_div(U, kBlockM);
  const int64_t n_tiles = ceil_div(n_items, kBlockN);
  if (m_tiles >= 2 * kNumSMs) return 1;
  int64_t s = ceil_div(2 * kNumSMs, m_tiles);
  const int64_t max_s = n_tiles / 8 > 0 ? n_tiles / 8 : 1;  // >= 8 tiles per split
  if (s > max_s) s = max_s;
  if (s > 64) s = 64;
  return static_cast<int>(s);
}

template <int D, int MODE>
static int launch_sweep(const CUtensorMap& mu, const CUtensorMap& mi, const CatalogParams& p,
                        cudaStream_t st) {
  const int smem = CatSmem<D>::bytes(MODE == kModeTopK ? p.k : 0);
  static bool attr_set[kMaxK + 1] = {};
  const int key = MODE == kModeTopK ? p.k : 0;
  if (!attr_set[key]) {
    cudaError_t e = cudaFuncSetAttribute(catalog_sweep_kernel<D, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    RG_REQUIRE(e == cudaSuccess, RECGURU_E_CUDA, "catalog: cudaFuncSetAttribute: %s",
               cudaGetErrorString(e));
    attr_set[key] = true;
  }
  const dim3 grid(static_cast<unsigned>(ceil_div(p.U, kBlockM)), static_cast<unsigned>(p.n_splits));
  catalog_sweep_kernel<D, MODE><<<grid, kCatThreads, smem, st>>>(mu, mi, p);
  RG_CHECK_LAUNCH("catalog_sweep");
  return RECGURU_OK;
}

static int check_catalog(const char* who, const void* users, int64_t U, const void* items,
                         int64_t n_items, int64_t d) {
  RG_REQUIRE(users && items, RECGURU_E_SHAPE, "%s: null pointer", who);
  RG_REQUIRE(aligned16(users) && aligned16(items), RECGURU_E_SHAPE, "%s: operands must be 16-byte aligned", who);
  RG_REQUIRE(U > 0 && n_items > 0 && n_items < (1ll << 31) - kBlockN && U < (1ll << 31),
             RECGURU_E_SHAPE, "%s: bad sizes U=%lld n_items=%lld", who, (long long)U, (long long)n_items);
  RG_REQUIRE(d == 64 || d == 128, RECGURU_E_UNSUPPORTED, "%s: d=%lld not in {64,128}", who, (long long)d);
  return RECGURU_OK;
}

}  // namespace rg

extern "C" int64_t recguru_catalog_workspace_bytes(int64_t U, int64_t n_items, int64_t d,
                                                   int64_t k) {
  (void)d;
  if (U <= 0 || n_items <= 0 || k < 1 || k > rg::kMaxK) return RECGURU_E_SHAPE;
  const int s = rg::pick_splits(U, n_items);
  if (s == 1) return 256;
  return rg::align_up(static_cast<int64_t>(s) * U * k * 12, 256) +
         rg::align_up(static_cast<int64_t>(s) * U * 8, 256) + 512;
}

extern "C" int recguru_catalog_target_scores(const void* users_bf16, int64_t U,
                                             const void* items_bf16, int64_t id_lo,
                                             int64_t n_items, int64_t d, const int64_t* target,
                                             float* target_score, recguru_stream_t stream) {
  using namespace rg;
  int rc = check_catalog("catalog_target_scores", users_bf16, U, items_bf16, n_items, d);
  if (rc) return rc;
  RG_REQUIRE(target && target_score, RECGURU_E_SHAPE, "catalog_target_scores: null pointer");
  CUtensorMap mu;
  rc = make_rowmajor_map(&mu, users_bf16, U, d);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(ceil_div(U, kBlockM));
  const int smem = static_cast<int>(2 * kBlockN * d * 2 + 64 + 1024);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(catalog_target_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(catalog_target_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  const __nv_bfloat16* items = static_cast<const __nv_bfloat16*>(items_bf16);
  if (d == 64)
    catalog_target_kernel<64><<<grid, 192, smem, st>>>(mu, items, id_lo, n_items, target, U, target_score);
  else
    catalog_target_kernel<128><<<grid, 192, smem, st>>>(mu, items, id_lo, n_items, target, U, target_score);
  RG_CHECK_LAUNCH("catalog_target");
  return RECGURU_OK;
}

extern "C" int recguru_catalog_topk_rank(const void* users_bf16, int64_t U, const void* items_bf16,
                                         int64_t id_lo, int64_t n_items, int64_t d,
                                         const int64_t* target, const float* target_score,
                                         int64_t k, float* topk_val, int64_t* topk_id,
                                         int64_t* rank_count, void* workspace,
                                         int64_t workspace_bytes, recguru_stream_t stream) {
  using namespace rg;
  int rc = check_catalog("catalog_topk_rank", users_bf16, U, items_bf16, n_items, d);
  if (rc) return rc;
  RG_REQUIRE(target && target_score && topk_val && topk_id && rank_count, RECGURU_E_SHAPE,
             "catalog_topk_rank: null pointer");
  RG_REQUIRE(k >= 1 && k <= kMaxK, RECGURU_E_UNSUPPORTED, "catalog_topk_rank: k=%lld not in [1,%d]",
             (long long)k, kMaxK);
  CUtensorMap mu, mi;
  rc = make_rowmajor_map(&mu, users_bf16, U, d);
  if (rc) return rc;
  rc = make_rowmajor_map(&mi, items_bf16, n_items, d);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CatalogParams p;
  p.U = U; p.n_items = n_items; p.id_lo = id_lo; p.target = target; p.tscore = target_score;
  p.k = static_cast<int>(k); p.dense = nullptr;
  p.n_splits = pick_splits(U, n_items);
  const int64_t n_tiles = ceil_div(n_items, kBlockN);
  p.tiles_per_split = static_cast<int>(ceil_div(n_tiles, p.n_splits));
  if (p.n_splits == 1) {
    p.out_val = topk_val; p.out_id = topk_id; p.out_cnt = rank_count;
  } else {
    const int64_t need = recguru_catalog_workspace_bytes(U, n_items, d, k);
    RG_REQUIRE(workspace && workspace_bytes >= need, RECGURU_E_WORKSPACE,
               "catalog_topk_rank: workspace %lld < %lld", (long long)workspace_bytes, (long long)need);
    char* ws = static_cast<char*>(workspace);
    const int64_t sv = align_up(static_cast<int64_t>(p.n_splits) * U * k * 4, 256);
    const int64_t si = align_up(static_cast<int64_t>(p.n_splits) * U * k * 8, 256);
    p.out_val = reinterpret_cast<float*>(ws);
    p.out_id = reinterpret_cast<int64_t*>(ws + sv);
    p.out_cnt = reinterpret_cast<int64_t*>(ws + sv + si);
  }
  rc = d == 64 ? launch_sweep<64, kModeTopK>(mu, mi, p, st) : launch_sweep<128, kModeTopK>(mu, mi, p, st);
  if (rc) return rc;
  if (p.n_splits > 1)
    return recguru_topk_merge(p.out_val, p.out_id, p.out_cnt, p.n_splits, U, k, topk_val, topk_id,
                              rank_count, stream);
  return RECGURU_OK;
}

extern "C" int recguru_catalog_scores_dense(const void* users_bf16, int64_t U,
                                            const void* items_bf16, int64_t n_items, int64_t d,
                                            float* scores, recguru_stream_t stream) {
  using namespace rg;
  int rc = check_catalog("catalog_scores_dense", users_bf16, U, items_bf16, n_items, d);
  if (rc) return rc;
  RG_REQUIRE(scores, RECGURU_E_SHAPE, "catalog_scores_dense: null pointer");
  CUtensorMap mu, mi;
  rc = make_rowmajor_map(&mu, users_bf16, U, d);
  if (rc) return rc;
  rc = make_rowmajor_map(&mi, items_bf16, n_items, d);
  if (rc) return rc;
  CatalogParams p;
  memset(&p, 0, sizeof(p));
  p.U = U; p.n_items = n_items; p.k = 1; p.dense = scores;
  p.n_splits = pick_splits(U, n_items);
  p.tiles_per_split = static_cast<int>(ceil_div(ceil_div(n_items, kBlockN), p.n_splits));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d == 64 ? launch_sweep<64, kModeDense>(mu, mi, p, st) : launch_sweep<128, kModeDense>(mu, mi, p, st);
}
