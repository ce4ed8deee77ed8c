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
//   warps 1-2   MMA issuers (one elected lane each, alternate tiles); warp 1 owns the TMEM
//               allocation (4 x 128 fp32 columns = 512)
//   warps 3-6   epilogue set 0: even tiles      each set = 4 warps = the 4 TMEM lane quarters;
//   warps 7-10  epilogue set 1: odd tiles       thread <-> user row, registers <-> item columns
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
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace rg {

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;   // target kernel tile; the sweep's item tile is the template parameter BN
constexpr int kStages = 4;       // smem ring of item tiles
constexpr int kAccum = 4;        // TMEM accumulator buffers (4 x 128 columns)
constexpr int kMaxK = 32;
constexpr int kFmaSignCols = 16;  // of every 32 columns, sign bits accumulated on the FMA pipe

enum CatalogMode { kModeTopK = 0, kModeDense = 1, kModeRank = 2 };   // Rank: target rank only (k = 0)

// Build flag RECGURU_DEBUG_KNOBS (python -m recguru_b200.build --debug-knobs): compiles the
// RECGURU_CATALOG_DEBUG bisection knobs, the clock64 profile/trace globals and the experimental
// instantiations (4 epilogue sets, cluster multicast, A-in-TMEM, 256-item tiles, deeper smem ring,
// software-pipelined epilogue) into the library.  The production library carries none of them.
#ifdef RECGURU_DEBUG_KNOBS
constexpr bool kDebugKnobs = true;
#else
constexpr bool kDebugKnobs = false;
#endif
// debug profile (RECGURU_CATALOG_DEBUG & 512): cycles CTA (0,0) spends in each wait, per role
#ifdef RECGURU_DEBUG_KNOBS
__device__ long long g_prof[16];
#define RG_DBG(...) __VA_ARGS__
#else
#define RG_DBG(...)
#endif
// debug flag 65536: per-tile timestamps (clock64, CTA 0) of a window of tiles:
// [tile][0] MMA warp: accumulator slot seen empty   [1] operands seen full   [2] MMAs + commit issued
// [tile][3] epilogue: accumulator seen full         [4] epilogue: slot handed back
// [tile][5] producer: stage seen free               [6] producer: TMA issued
constexpr int kTraceTiles = 64, kTraceFirst = 2000;
// setmaxnreg budget with 12 warps (NSETS == 2): 4*32*kCtlRegs + 8*32*kEpiRegs <= 12*32*168
constexpr int kCtlRegs = 56, kEpiRegs = 224;
static_assert(4 * kCtlRegs + 8 * kEpiRegs <= 12 * 168, "register budget");
RG_DBG(__device__ long long g_trace[kTraceTiles][8];)

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
  const __nv_bfloat16* users;  // [U, d] (TS mode reads the user tile directly)
  int tie_mode;          // RECGURU_TIE_*: how items that tie with the target count towards its rank
  uint32_t two;          // the constant 2, opaque to the compiler (keeps IMAD.HI on the FMA pipe)
  uint32_t debug_flags;  // bit 0: TMEM-load-only epilogue (timing floor experiments; wrong results)
};

__device__ __forceinline__ float next_below(float t) {
  const uint32_t b = __float_as_uint(t);
  if (t > 0.f) return __uint_as_float(b - 1u);
  if (t == 0.f) return __uint_as_float(0x80000001u);
  return __uint_as_float(b + 1u);
}

// Shared-memory carve-up (dynamic, base aligned to 1024 for the 128-byte swizzle).
// TMEM holds 512/BN accumulators (4 of 128 items, or 2 of 256); the shared-memory ring of item
// tiles has NST stages (0 = same depth as the accumulator ring, one commit serves both).
template <int D, int BN, int NST = 0>
struct CatSmem {
  static constexpr int kKBlocks = D / 64;                 // 64-wide (128 B) K blocks
  static constexpr int kRing = 512 / BN;                  // TMEM accumulators
  static constexpr int kStages = NST > 0 ? NST : kRing;   // smem stages
  static constexpr int kATileBytes = kBlockM * D * 2;     // user tile
  static constexpr int kTileBytes = BN * D * 2;           // one item tile
  static constexpr int kOffA = 0;
  static constexpr int kOffB = kATileBytes;
  static constexpr int kOffBar = kOffB + kStages * kTileBytes;
  // barriers: full[8], empty[8], acc_full[8], acc_empty[8], a_full (sized for the deepest ring)
  static constexpr int kNumBars = 4 * 8 + 1;
  static constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr int kOffCnt = kOffTmemPtr + 8;         // int[nsets][128] (per-set rank counts)
  // then float[nsets*k*128] scores and int[nsets*k*128] ids of the per-thread top-k lists
  __host__ __device__ static constexpr int off_list(int nsets) { return kOffCnt + nsets * 128 * 4; }
  __host__ __device__ static constexpr int bytes(int k, int nsets) { return off_list(nsets) + nsets * k * 128 * 8 + 1024; }
};

// CL = cluster size (1 or 2).  With CL == 2 the two CTAs of a cluster sweep the same item tiles
// for different user tiles: each loads half of every item tile and TMA-multicasts it to both,
// halving the L2 -> SM operand traffic (the item table is streamed once per CTA otherwise).
// TS: the user tile (A operand) lives in TMEM for the whole sweep (written once with tcgen05.st),
// so every MMA reads only the item tile from shared memory -- in SS mode the 128 x 16 A slice is
// re-read from smem by each instruction and the N=128 MMA is shared-memory-bandwidth bound.
// TMEM then holds 3 accumulators (384 columns) + A (64 columns).
// NST: stages of the shared-memory item-tile ring.  0 couples it to the accumulator ring (one
// tcgen05.commit per tile frees the stage and publishes the accumulator).  NST > NACC decouples
// the two: the TMA for tile j+NST is issued as soon as tile j's MMAs retire, i.e. NST-NACC tiles
// before the accumulator hand-off that gates its MMAs, so the TMA latency leaves the critical
// chain commit -> epilogue -> MMA (measured: 35 -> see DESIGN.md).
// PIPE: software-pipelined epilogue (two 64-column TMEM loads per tile, the second one and the
// first one of the NEXT tile in flight while the rank count of the previous half runs).
// DBG: the RECGURU_CATALOG_DEBUG bisection knobs are compiled in only in this variant; the
// production instantiation carries none of their per-tile branches.
// RL: the per-thread top-k list (k <= 16) lives in registers instead of shared memory.
template <int D, int MODE, int NSETS, int CL, bool TS, int BN, int NST, bool PIPE, bool DBG, bool RL>
__global__ void __launch_bounds__(128 + NSETS * 128, 1)
catalog_sweep_kernel(const __grid_constant__ CUtensorMap map_users,
                     const __grid_constant__ CUtensorMap map_items, CatalogParams p) {
  const uint32_t dflags = DBG ? p.debug_flags : 0u;
  using L = CatSmem<D, BN, NST>;
  static_assert(!TS || BN == 128, "TS mode is only wired for 128-item tiles");
  constexpr int NACC = TS ? 3 : L::kRing;   // accumulator ring
  constexpr int NSTG = NST > 0 ? NST : NACC;   // smem stage ring
  constexpr bool kCoupled = NST == 0 && CL == 1;   // one commit per tile serves both rings
  constexpr uint32_t kTmemA = NACC * BN;    // TMEM column of the A operand (TS mode)
  constexpr int kBarStride = 8;             // barrier arrays are laid out for the deepest ring
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + kBarStride;
  uint64_t* bar_acc_full = bars + 2 * kBarStride;
  uint64_t* bar_acc_empty = bars + 3 * kBarStride;
  uint64_t* bar_a = bars + 4 * kBarStride;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kOffTmemPtr);
  int* s_cnt = reinterpret_cast<int*>(smem + L::kOffCnt);
  float* s_val = reinterpret_cast<float*>(smem + L::off_list(NSETS));
  constexpr int kSlots = NSETS * 128;       // list entry i of slot s lives at [i * kSlots + s]
  int* s_id = reinterpret_cast<int*>(smem + L::off_list(NSETS) + NSETS * p.k * 128 * 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kBlockM;
  const int total_tiles = static_cast<int>((p.n_items + BN - 1) / BN);
  const int tile_begin = blockIdx.y * p.tiles_per_split;
  const int tile_end = min(total_tiles, tile_begin + p.tiles_per_split);
  const int n_tiles = max(0, tile_end - tile_begin);

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&map_users);
    ptx::tma_prefetch_desc(&map_items);
    for (int i = 0; i < NSTG; ++i) { ptx::mbar_init(&bar_full[i], 1); ptx::mbar_init(&bar_empty[i], CL); }
    for (int i = 0; i < NACC; ++i) { ptx::mbar_init(&bar_acc_full[i], 1); ptx::mbar_init(&bar_acc_empty[i], (dflags & 16384u) ? 1 : 4); }
    ptx::mbar_init(bar_a, TS ? 4 : 1);
    ptx::mbar_fence_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();      // peer barriers initialised before any multicast
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t cta_rank = CL > 1 ? ptx::cluster_ctarank() : 0u;
  constexpr uint16_t kClMask = (1u << CL) - 1u;
  const bool spin = (dflags & 256u) != 0;   // debug: busy-poll barriers
  const bool prof = (dflags & 512u) != 0 && blockIdx.x == 0 && blockIdx.y == 0;
  const bool trace = (dflags & 65536u) != 0 && blockIdx.x == 0 && blockIdx.y == 0;

  if constexpr (NSETS == 2) {
    if (warp < 4) ptx::setmaxnreg_dec<kCtlRegs>();
  }
  if (warp == 0) {
    // ===================== TMA producer (whole warp loops; one elected lane issues) ========
    const bool leader = ptx::elect_one();
    if (leader && !TS) {
      ptx::mbar_arrive_expect_tx(bar_a, L::kATileBytes);
      for (int kb = 0; kb < L::kKBlocks; ++kb)
        ptx::tma_load_2d(&map_users, bar_a, smem + L::kOffA + kb * (kBlockM * 128), kb * 64,
                         static_cast<int32_t>(m0));
    }
    for (int j = 0; j < n_tiles; ++j) {
      if (dflags & 32u) break;                 // debug: no smem-ring handshake at all
      const int s = j % NSTG;
      const uint32_t ph = (j / NSTG) & 1;
      const long long tp0 = prof ? clock64() : 0;
      // coupled rings: the stage of tile j-4 is free exactly when that tile's MMAs have retired,
      // which is what acc_full[j % 4] signals to the epilogue -- one tcgen05.commit per tile serves
      // both.  Otherwise (deeper smem ring, or a cluster whose peer's MMAs count too) the stage has
      // its own barrier.
      ptx::mbar_wait_warp(kCoupled ? &bar_acc_full[s] : &bar_empty[s], ph ^ 1, spin);
      RG_DBG(if (prof && lane == 0) g_prof[0] += clock64() - tp0;)
      RG_DBG(if (trace && lane == 0 && j >= kTraceFirst && j < kTraceFirst + kTraceTiles) g_trace[j - kTraceFirst][5] = clock64();)
      if (leader) {
        if ((dflags & 4u) && j >= NSTG) {   // debug: no operand traffic (timing floor)
          ptx::mbar_arrive(&bar_full[s]);
        } else {
          ptx::mbar_arrive_expect_tx(&bar_full[s], L::kTileBytes);
          uint8_t* dst = smem + L::kOffB + s * L::kTileBytes;
          const int32_t n0 = (tile_begin + j) * BN;
          if (CL == 1) {
            for (int kb = 0; kb < L::kKBlocks; ++kb)
              ptx::tma_load_2d(&map_items, &bar_full[s], dst + kb * (BN * 128), kb * 64, n0);
          } else {
            // this CTA fetches rows [rank*64, rank*64+64) of the tile for the whole cluster
            constexpr int kHalf = BN / CL;
            for (int kb = 0; kb < L::kKBlocks; ++kb)
              ptx::tma_load_2d_mcast(&map_items, &bar_full[s],
                                     dst + kb * (BN * 128) + cta_rank * (kHalf * 128), kb * 64,
                                     n0 + static_cast<int32_t>(cta_rank) * kHalf, kClMask);
          }
        }
      }
      __syncwarp();
      RG_DBG(if (trace && lane == 0 && j >= kTraceFirst && j < kTraceFirst + kTraceTiles) g_trace[j - kTraceFirst][6] = clock64();)
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers (whole warp loops; one elected lane issues) =========
    // Two issuer warps take alternate tiles: the per-tile wait / issue / commit chain of a single
    // thread (~1700 cycles) would otherwise pace the whole kernel.
    const bool leader = ptx::elect_one();
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(kBlockM, BN);
    const uint32_t a_addr = ptx::smem_u32(smem + L::kOffA);
    const int reps = (dflags & 16u) ? 0 : ((dflags & 8u) ? 2 : 1);  // debug knobs
    const uint32_t mma_acc_empty_addr = ptx::smem_u32(bar_acc_empty);
    const uint32_t mma_acc_full_addr = ptx::smem_u32(bar_acc_full);
    const uint32_t mma_full_addr = ptx::smem_u32(bar_full);
    ptx::mbar_wait_warp(bar_a, 0);
    ptx::tc_fence_after();
    for (int j = warp - 1; j < n_tiles; j += 2) {
      const int s = j % NSTG;
      const uint32_t ph = (j / NSTG) & 1;
      const int a = j % NACC;
      const uint32_t aph = (j / NACC) & 1;
      const long long tm0 = prof ? clock64() : 0;
      if (!DBG) ptx::mbar_wait_warp_addr(mma_acc_empty_addr + a * 8, aph ^ 1);
      else if (!(dflags & 64u)) ptx::mbar_wait_warp(&bar_acc_empty[a], aph ^ 1, spin);
      const long long tm1 = (prof || trace) ? clock64() : 0;
      if (!DBG) ptx::mbar_wait_warp_addr(mma_full_addr + s * 8, ph);
      else if (!(dflags & 32u)) ptx::mbar_wait_warp(&bar_full[s], ph, spin);
      const long long tm2 = (prof || trace) ? clock64() : 0;
      ptx::tc_fence_after();
      if (leader) {
        const uint32_t b_addr = ptx::smem_u32(smem + L::kOffB + s * L::kTileBytes);
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t koff = (k / 4) * (kBlockM * 128) + (k % 4) * 32;   // bytes into the user tile
            const uint32_t koff_b = (k / 4) * (BN * 128) + (k % 4) * 32;      // bytes into the item tile
            if (dflags & 1024u)   // debug: N=256 instruction rate (operands/accumulators overlap; timing only)
              ptx::umma_bf16(tmem_base + (a & 1) * 256, ptx::umma_smem_desc_sw128(a_addr + koff),
                             ptx::umma_smem_desc_sw128(ptx::smem_u32(smem + L::kOffB) + (k % 4) * 32),
                             ptx::umma_idesc_bf16(kBlockM, 256), (k > 0 || rep > 0) ? 1u : 0u);
            else if (TS)
              ptx::umma_bf16_ts(tmem_base + a * BN, tmem_base + kTmemA + k * 8,
                                ptx::umma_smem_desc_sw128(b_addr + koff_b), idesc, (k > 0 || rep > 0) ? 1u : 0u);
            else
              ptx::umma_bf16(tmem_base + a * BN, ptx::umma_smem_desc_sw128(a_addr + koff),
                             ptx::umma_smem_desc_sw128(b_addr + koff_b), idesc, (k > 0 || rep > 0) ? 1u : 0u);
          }
        }
        if (!(dflags & 32u)) {       // smem stage free (in every CTA of the cluster) once these MMAs retire
          if (CL > 1) ptx::umma_commit_mcast(&bar_empty[s], kClMask);
          else if (!kCoupled) ptx::umma_commit(&bar_empty[s]);
        }
        if (dflags & 128u) ptx::mbar_arrive(&bar_acc_full[a]);     // debug: plain arrive instead of commit
        else if (!DBG) ptx::umma_commit_addr(mma_acc_full_addr + a * 8);
        else ptx::umma_commit(&bar_acc_full[a]);  // accumulator ready for the epilogue (and stage free)
      }
      __syncwarp();
      if (prof && lane == 0 && warp == 1) {
        RG_DBG(g_prof[1] += tm1 - tm0; g_prof[2] += tm2 - tm1; g_prof[3] += clock64() - tm2; g_prof[4] += 1;)
      }
      if (trace && lane == 0 && j >= kTraceFirst && j < kTraceFirst + kTraceTiles) {
        RG_DBG(g_trace[j - kTraceFirst][0] = tm1; g_trace[j - kTraceFirst][1] = tm2; g_trace[j - kTraceFirst][2] = clock64();)
      }
    }
  } else if (warp == 3) {
    // idle: fills the control warpgroup (setmaxnreg acts on aligned groups of four warps)
  } else {
    // ===================== epilogue =====================
    // The kernel is compiled for 128 + NSETS*128 threads (168 registers each with two sets); the
    // control warpgroup gives most of its registers back and the epilogue warpgroups take them,
    // so the 128-register accumulator row, the sign counters and the group maxima fit without
    // spilling.
    if constexpr (NSETS == 2) ptx::setmaxnreg_inc<kEpiRegs>();
    const int e = (warp - 4) >> 2;            // epilogue set: handles tiles j == e (mod NSETS)
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int64_t u = m0 + row;
    const bool u_ok = u < p.U;
    const int K = p.k;
    const uint32_t two = p.two;
    const int slot = e * 128 + row;           // list slot; entry i at [i * kSlots + slot]

    float t = 0.f;
    int64_t tgt_local = -1;                   // target's row index inside the slice
    if (MODE != kModeDense && u_ok) {
      t = p.tscore[u];
      if (t == 0.f) t = 0.f;                // -0 -> +0 (the sign-bit count relies on it)
      tgt_local = p.target[u] - p.id_lo;
    }
    const float t_below = next_below(t);
    int tgt_tile, tgt_col = -1;               // tile relative to this device's slice
    if (tgt_local < 0) tgt_tile = -1;
    else if (tgt_local >= p.n_items) tgt_tile = INT_MAX;
    else { tgt_tile = static_cast<int>(tgt_local / BN); tgt_col = static_cast<int>(tgt_local % BN); }

    int cnt = 0;
    float thr = -CUDART_INF_F;
    if (MODE == kModeTopK && !RL) {
      for (int i = 0; i < K; ++i) { s_val[i * kSlots + slot] = -CUDART_INF_F; s_id[i * kSlots + slot] = -1; }
    }
    // RL: 16 register slots, descending; the k live entries sit at the BOTTOM (slots 16-k .. 15)
    // and the slots above them are blocked with +inf, so the k-th best -- the insertion
    // threshold -- is always lv[15] whatever k is, and an insertion is a fully unrolled
    // compare-exchange chain with no memory access and no data-dependent loop.
    float lv[RL ? 16 : 1];
    int li[RL ? 16 : 1];
    if constexpr (RL) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { lv[i] = i < 16 - K ? CUDART_INF_F : -CUDART_INF_F; li[i] = -1; }
    }
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    if (TS && e == 0) {
      // this thread's user row (128 bf16 = 64 packed words) -> TMEM lane `row`, columns kTmemA..
      uint32_t av[64];
      const uint4* src = reinterpret_cast<const uint4*>(p.users + (u_ok ? u : 0) * D);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const uint4 x = u_ok ? __ldg(src + q) : make_uint4(0u, 0u, 0u, 0u);
        av[4 * q] = x.x; av[4 * q + 1] = x.y; av[4 * q + 2] = x.z; av[4 * q + 3] = x.w;
      }
      ptx::tmem_st_32x64(tmem_base + lane_base + kTmemA, av);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_a);
    }
    uint32_t cf[4] = {0u, 0u, 0u, 0u};        // sign-bit counters fed by the FMA pipe (IMAD.HI)
    uint32_t ca[4] = {0u, 0u, 0u, 0u};        // sign-bit counters fed by the ALU pipe (LEA.HI)
    uint32_t cf2[4] = {0u, 0u, 0u, 0u}, ca2[4] = {0u, 0u, 0u, 0u};

    // sorted insert into this thread's list (desc score; equal scores keep arrival = id order)
    auto insert = [&](float v, int id) {
      if constexpr (RL) {
        bool sw = false;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // strict: an equal score stays below earlier ones; once the new entry is placed every
          // later slot just shifts down (a carried entry must not be compared again, or it would
          // slip below an equal neighbour and break the id order of ties)
          sw = sw || (v > lv[i]);
          const float ov = lv[i];
          const int oi = li[i];
          lv[i] = sw ? v : ov;
          li[i] = sw ? id : oi;
          v = sw ? ov : v;
          id = sw ? oi : id;
        }
        thr = lv[15];
        return;
      }
      int pos = K - 1;
      while (pos > 0 && s_val[(pos - 1) * kSlots + slot] < v) {
        s_val[pos * kSlots + slot] = s_val[(pos - 1) * kSlots + slot];
        s_id[pos * kSlots + slot] = s_id[(pos - 1) * kSlots + slot];
        --pos;
      }
      s_val[pos * kSlots + slot] = v;
      s_id[pos * kSlots + slot] = id;
      thr = s_val[(K - 1) * kSlots + slot];
    };
    // Offer eight scores (columns id0 .. id0+7, bit c of `ok` = column c is a real item) to the list.
    auto offer8 = [&](const uint32_t (&w8)[8], int id0, uint32_t ok) {
      if constexpr (RL) {
        // one inlined insertion site: lanes queue the columns that beat their threshold and the
        // warp drains the queues together (usually one lane, one column)
        uint32_t pend = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) pend |= (__uint_as_float(w8[c]) > thr ? 1u : 0u) << c;
        pend &= ok;
        while (__any_sync(0xffffffffu, pend != 0u)) {
          if (pend) {
            const int c = __ffs(pend) - 1;
            pend &= pend - 1;
            float v = __uint_as_float(w8[0]);
#pragma unroll
            for (int q = 1; q < 8; ++q) v = (c == q) ? __uint_as_float(w8[q]) : v;
            if (v > thr) insert(v, id0 + c);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float v = __uint_as_float(w8[c]);
          if (((ok >> c) & 1u) && v > thr) insert(v, id0 + c);
        }
      }
    };
    // ---- pieces of the fast path on one 64-column half held in registers -------------------
    // (a) 8-column group maxima decide whether the top-k lists must be touched; if some lane's
    //     threshold was beaten the warp re-reads just those groups from TMEM (the accumulator is
    //     still ours; warp-uniform loop => convergent .sync.aligned load)
    auto topk_check64 = [&](const uint32_t (&vr)[64], int id0) {
      float gm[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float* f = reinterpret_cast<const float*>(&vr[g * 8]);
        const float m1 = fmaxf(fmaxf(f[0], f[1]), f[2]);
        const float m2 = fmaxf(fmaxf(f[3], f[4]), f[5]);
        const float m3 = fmaxf(fmaxf(f[6], f[7]), m1);
        gm[g] = fmaxf(m2, m3);
      }
      const float a3 = fmaxf(fmaxf(gm[0], gm[1]), gm[2]);
      const float b3 = fmaxf(fmaxf(gm[3], gm[4]), gm[5]);
      const float mx = fmaxf(fmaxf(gm[6], gm[7]), fmaxf(a3, b3));
      if (__any_sync(0xffffffffu, mx > thr)) {
        uint32_t gmask = 0;
#pragma unroll
        for (int g = 0; g < 8; ++g) gmask |= (gm[g] > thr ? 1u : 0u) << g;
        uint32_t wmask = __reduce_or_sync(0xffffffffu, gmask);
        while (wmask) {
          const int g = __ffs(wmask) - 1;
          wmask &= wmask - 1;
          uint32_t w8[8];
#define RG_GRP(G) case G: _Pragma("unroll") for (int c = 0; c < 8; ++c) w8[c] = vr[(G) * 8 + c]; break;
          switch (g) {
            RG_GRP(0) RG_GRP(1) RG_GRP(2) RG_GRP(3) RG_GRP(4) RG_GRP(5) RG_GRP(6)
            default: _Pragma("unroll") for (int c = 0; c < 8; ++c) w8[c] = vr[7 * 8 + c]; break;
          }
#undef RG_GRP
          offer8(w8, id0 + g * 8, 0xffu);
        }
      }
    };
    // (b) rank count: sign bit of (tcmp - v); two differences per FADD2, sign bits accumulated
    //     half on the FMA pipe (IMAD.HI by a run-time 2), half on the ALU pipe (LEA.HI)
    auto rank_count64 = [&](const uint32_t (&vr)[64], uint64_t tt) {
#pragma unroll
      for (int c = 0; c < 64; c += 8) {
        uint64_t dd[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint64_t vv;
          asm("mov.b64 %0, {%1, %2};" : "=l"(vv) : "r"(vr[c + 2 * q]), "r"(vr[c + 2 * q + 1]));
          asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dd[q]) : "l"(tt), "l"(vv));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t d0, d1;
          asm("mov.b64 {%0, %1}, %2;" : "=r"(d0), "=r"(d1) : "l"(dd[q]));
          if (((c + 2 * q) & 31) < kFmaSignCols) {
            asm("mad.hi.u32 %0, %1, %2, %0;" : "+r"(cf[q & 1]) : "r"(d0), "r"(two));
            asm("mad.hi.u32 %0, %1, %2, %0;" : "+r"(cf2[q & 1]) : "r"(d1), "r"(two));
          } else {
            ca[q & 1] += d0 >> 31;
            ca2[q & 1] += d1 >> 31;
          }
        }
      }
    };
    const uint32_t acc_full_addr = ptx::smem_u32(bar_acc_full);     // + 8 * slot
    const uint32_t acc_empty_addr = ptx::smem_u32(bar_acc_empty);
    const int n_items32 = static_cast<int>(p.n_items);               // < 2^31 (checked on the host)
    uint32_t vrA[PIPE ? 64 : 1], vrB[PIPE ? 64 : 1];   // PIPE: the two halves of a tile's accumulator row
    bool a_inflight = false;                            // PIPE: this tile's first half was requested last iteration
    for (int j = e; j < n_tiles; j += NSETS) {
      if ((dflags & 16384u) && quarter != 0) break;   // debug: one warp per set in the handshake
      const int a = j % NACC;
      const uint32_t aph = (j / NACC) & 1;
      const int tile = tile_begin + j;
      const int n0 = tile * BN;                                   // slice-local item index
      const int valid = min(BN, n_items32 - n0);
      const bool free_run = (dflags & 32768u) != 0;   // debug: epilogue ignores the accumulator barriers
      if ((dflags & 64u) && !free_run) break;  // debug: no accumulator handshake at all
      const long long te0 = prof ? clock64() : 0;
      if (!free_run && !a_inflight) {
        if (DBG && spin) ptx::mbar_wait_warp(&bar_acc_full[a], aph, true);
        else ptx::mbar_wait_warp_addr(acc_full_addr + a * 8, aph);
      }
      const long long te1 = (prof || trace) ? clock64() : 0;
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + a * BN;
      bool released = false;

      if (MODE == kModeDense) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t vr[32];
          ptx::tmem_ld_32x32(taddr + c0, vr);
          ptx::tmem_ld_wait();
          if (u_ok) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c0 + c < valid) p.dense[u * p.n_items + n0 + c0 + c] = __uint_as_float(vr[c]);
          }
        }
      } else if (__any_sync(0xffffffffu, tile == tgt_tile) || valid < BN) {
        // exact tile (warp-uniform): some lane's target lives here and/or the tile runs past
        // the end of the slice.  Per-column rules, 8 columns at a time.
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 8) {
          uint32_t w[8];
          ptx::tmem_ld_32x8(taddr + c0, w);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int col = c0 + c;
            const float v = __uint_as_float(w[c]);
            if (col < valid) {
              if (tile == tgt_tile) {
                if (col < tgt_col) cnt += (v >= t) ? 1 : 0;
                else if (col > tgt_col) cnt += (v > t) ? 1 : 0;
              } else {
                cnt += (v > (tile < tgt_tile ? t_below : t)) ? 1 : 0;
              }
            }
          }
          if constexpr (MODE == kModeTopK)
            offer8(w, n0 + c0, valid - c0 >= 8 ? 0xffu : (valid > c0 ? (1u << (valid - c0)) - 1u : 0u));
        }
      } else if constexpr (PIPE) {
        // fast tile, software pipelined.  TMEM reads are slow next to a running tensor core
        // (a 64 KB tile takes ~1000 cycles) and a warp's rank count another ~900, so neither may
        // wait for the other:   A(j) ready | request B(j) | top-k + count A | B ready | top-k B |
        // release accumulator | request A(j+2) if that tile is already full | count B
        const float tcmp = tile < tgt_tile ? t_below : t;
        uint64_t tt;
        asm("mov.b64 %0, {%1, %1};" : "=l"(tt) : "f"(tcmp));
        if (!a_inflight) ptx::tmem_ld_32x64(taddr, vrA);
        ptx::tmem_ld_wait_regs(vrA);
        ptx::tmem_ld_32x64(taddr + 64, vrB);
        if (!(dflags & 1u)) {
          topk_check64(vrA, n0);
          rank_count64(vrA, tt);
        } else {
          ca[0] += vrA[0] ^ vrA[63];                  // debug: TMEM-load-only floor
        }
        ptx::tmem_ld_wait_regs(vrB);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_addr(acc_empty_addr + a * 8);
        released = true;
        a_inflight = false;
        const int jn = j + NSETS;
        if (jn < n_tiles && !free_run) {
          const int tile_n = tile_begin + jn;
          const bool fast_n = !__any_sync(0xffffffffu, tile_n == tgt_tile) &&
                              static_cast<int64_t>(tile_n + 1) * BN <= p.n_items;
          if (fast_n) {
            const int an = jn % NACC;
            int ready = 0;
            if (lane == 0) ready = ptx::mbar_test(&bar_acc_full[an], (jn / NACC) & 1) ? 1 : 0;
            ready = __shfl_sync(0xffffffffu, ready, 0);
            if (ready) {
              ptx::tc_fence_after();
              ptx::tmem_ld_32x64(tmem_base + lane_base + an * BN, vrA);
              a_inflight = true;
            }
          }
        }
        if (!(dflags & 1u)) {
          topk_check64(vrB, n0 + 64);
          rank_count64(vrB, tt);
        } else {
          ca[1] += vrB[0] ^ vrB[63];
        }
      } else {
        // fast tile: d = tcmp - v has its sign bit set iff v > tcmp (exact: IEEE subtraction is
        // zero only for equal operands, tcmp is never -0).  FADD2 forms two differences per
        // instruction on the FMA pipe; the sign bits are accumulated half on the FMA pipe
        // (IMAD.HI by a run-time 2) and half on the ALU pipe (LEA.HI), the 8-column group maxima
        // with FMNMX3 on the ALU pipe.  TMEM loads are double buffered against the compute.
        const float tcmp = tile < tgt_tile ? t_below : t;
        uint64_t tt;
        asm("mov.b64 %0, {%1, %1};" : "=l"(tt) : "f"(tcmp));
        // NSETS=2: the whole accumulator row in one TMEM load (128 registers, a single wait per
        // tile; the other set's load overlaps this warp's compute).  NSETS=4: two 64-column loads
        // (the register file only allows ~100 registers per thread with 18 warps).
        constexpr int CH = NSETS == 2 ? 128 : 64;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CH) {
          if (dflags & 2u) break;               // debug: no TMEM reads, no compute (timing floor)
          uint32_t vr[CH];
          ptx::tmem_ld_cols<CH>(taddr + c0, vr);
          ptx::tmem_ld_wait();
          // the scores now live in registers only (the top-k re-scan reads registers too): hand
          // the accumulator back to the MMA warp before any arithmetic
          if (c0 + CH == BN) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_addr(acc_empty_addr + a * 8);
            released = true;
          }
          if (dflags & 1u) {
            ca[0] += vr[0] ^ vr[CH - 1];   // debug: TMEM-load-only floor (results meaningless)
            continue;
          }
          if constexpr (MODE == kModeTopK) {
          // (a) 8-column group maxima first: they decide whether the top-k lists must be touched.
          // (A single row maximum + full re-read was measured: 54 vs 32 ms at cfg4 -- with 32 users
          // per warp SOME lane's threshold is beaten in about every second tile of a 1M-item
          // sweep, so the re-read must stay confined to the groups that did it.)
          float gm[CH / 8];
#pragma unroll
          for (int g = 0; g < CH / 8; ++g) {
            const float* f = reinterpret_cast<const float*>(&vr[g * 8]);
            // four 3-input maxima (FMNMX3) per 8 columns
            const float m1 = fmaxf(fmaxf(f[0], f[1]), f[2]);
            const float m2 = fmaxf(fmaxf(f[3], f[4]), f[5]);
            const float m3 = fmaxf(fmaxf(f[6], f[7]), m1);
            gm[g] = fmaxf(m2, m3);
          }
          float mx;
          {
            static_assert(CH / 8 == 16 || CH / 8 == 8, "group count");
            float r[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = CH / 8 == 16 ? fmaxf(gm[q], gm[q + 8]) : gm[q];
            const float a3 = fmaxf(fmaxf(r[0], r[1]), r[2]);
            const float b3 = fmaxf(fmaxf(r[3], r[4]), r[5]);
            mx = fmaxf(fmaxf(r[6], r[7]), fmaxf(a3, b3));
          }
          if (__any_sync(0xffffffffu, mx > thr)) {
            // some lane's threshold was beaten.  Per-lane bitmask of the 8-column groups that did
            // it; the warp walks the union (warp-uniform g) and every lane checks that group's
            // eight scores, fetched from its registers through a jump table -- no TMEM re-read,
            // so the accumulator is not held while the lists are updated.
            uint32_t gmask = 0;
#pragma unroll
            for (int g = 0; g < CH / 8; ++g) gmask |= (gm[g] > thr ? 1u : 0u) << g;
            uint32_t wmask = __reduce_or_sync(0xffffffffu, gmask);
            while (wmask) {
              const int g = __ffs(wmask) - 1;
              wmask &= wmask - 1;
              uint32_t w8[8];
#define RG_GRP(G) case G: _Pragma("unroll") for (int c = 0; c < 8; ++c) w8[c] = vr[(G) * 8 + c]; break;
              switch (g) {
                RG_GRP(0) RG_GRP(1) RG_GRP(2) RG_GRP(3) RG_GRP(4) RG_GRP(5) RG_GRP(6) RG_GRP(7)
                default:
                  if constexpr (CH / 8 == 16) {
                    switch (g) {
                      RG_GRP(8) RG_GRP(9) RG_GRP(10) RG_GRP(11) RG_GRP(12) RG_GRP(13) RG_GRP(14)
                      default: _Pragma("unroll") for (int c = 0; c < 8; ++c) w8[c] = vr[(CH / 8 - 1) * 8 + c]; break;
                    }
                  } else {
                    _Pragma("unroll") for (int c = 0; c < 8; ++c) w8[c] = vr[7 * 8 + c];
                  }
                  break;
              }
#undef RG_GRP
              offer8(w8, n0 + c0 + g * 8, 0xffu);
            }
          }
          }
          // (c) rank count: four differences (two FADD2) ahead of their sign accumulation so the
          // FMA-pipe latency overlaps
#pragma unroll
          for (int c = 0; c < CH; c += 8) {
            uint64_t dd[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint64_t vv;
              asm("mov.b64 %0, {%1, %2};" : "=l"(vv) : "r"(vr[c + 2 * q]), "r"(vr[c + 2 * q + 1]));
              asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dd[q]) : "l"(tt), "l"(vv));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t d0, d1;
              asm("mov.b64 {%0, %1}, %2;" : "=r"(d0), "=r"(d1) : "l"(dd[q]));
              if (((c + 2 * q) & 31) < kFmaSignCols) {
                asm("mad.hi.u32 %0, %1, %2, %0;" : "+r"(cf[q & 1]) : "r"(d0), "r"(two));
                asm("mad.hi.u32 %0, %1, %2, %0;" : "+r"(cf2[q & 1]) : "r"(d1), "r"(two));
              } else {
                ca[q & 1] += d0 >> 31;
                ca2[q & 1] += d1 >> 31;
              }
            }
          }
        }
      }
      // this warp is done reading accumulator `a` (the fast path released it already)
      if (!released) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_addr(acc_empty_addr + a * 8);
      }
      if (trace && lane == 0 && quarter == 0 && j >= kTraceFirst && j < kTraceFirst + kTraceTiles) {
        RG_DBG(g_trace[j - kTraceFirst][3] = te1; g_trace[j - kTraceFirst][4] = clock64();)
      }
      if (prof && lane == 0 && (warp == 4 || warp == 8)) {
        const int o = warp == 4 ? 5 : 8;
        RG_DBG(g_prof[o] += te1 - te0; g_prof[o + 1] += clock64() - te1; g_prof[o + 2] += 1;)
      }
    }
    cnt += static_cast<int>(cf[0] + cf[1] + cf[2] + cf[3] + ca[0] + ca[1] + ca[2] + ca[3] + cf2[0] + cf2[1] +
                            cf2[2] + cf2[3] + ca2[0] + ca2[1] + ca2[2] + ca2[3]);

    if constexpr (MODE == kModeTopK && RL) {
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (q >= 16 - K) {
          s_val[(q - (16 - K)) * kSlots + slot] = lv[q];
          s_id[(q - (16 - K)) * kSlots + slot] = li[q];
        }
      }
    }
    if (MODE == kModeRank) {
      s_cnt[e * 128 + row] = cnt;
      asm volatile("bar.sync 1, %0;" ::"n"(NSETS * 128) : "memory");
      if (e == 0 && u_ok) {
        int total = 0;
#pragma unroll
        for (int q = 0; q < NSETS; ++q) total += s_cnt[q * 128 + row];
        p.out_cnt[static_cast<int64_t>(blockIdx.y) * p.U + u] = total;
      }
    }
    if (MODE == kModeTopK) {
      s_cnt[e * 128 + row] = cnt;
      asm volatile("bar.sync 1, %0;" ::"n"(NSETS * 128) : "memory");
      if (e == 0 && u_ok) {
        // merge the NSETS lists of this row: (score desc, id asc)
        int head[NSETS];
#pragma unroll
        for (int q = 0; q < NSETS; ++q) head[q] = 0;
        const int64_t out = (static_cast<int64_t>(blockIdx.y) * p.U + u) * K;
        for (int i = 0; i < K; ++i) {
          float bv = -CUDART_INF_F;
          int bd = -1, bq = -1;
#pragma unroll
          for (int q = 0; q < NSETS; ++q) {
            if (head[q] < K) {
              const float v = s_val[head[q] * kSlots + q * 128 + row];
              const int d = s_id[head[q] * kSlots + q * 128 + row];
              if (d >= 0 && (bq < 0 || v > bv || (v == bv && d < bd))) { bv = v; bd = d; bq = q; }
            }
          }
#pragma unroll
          for (int q = 0; q < NSETS; ++q)
            if (q == bq) ++head[q];
          p.out_val[out + i] = bd >= 0 ? bv : -CUDART_INF_F;
          p.out_id[out + i] = bd >= 0 ? p.id_lo + bd : -1;
        }
        int total = 0;
#pragma unroll
        for (int q = 0; q < NSETS; ++q) total += s_cnt[q * 128 + row];
        p.out_cnt[static_cast<int64_t>(blockIdx.y) * p.U + u] = total;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();      // no CTA leaves while its peer may still signal it
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace rg
#include "catalog_sweep2.cuh"
namespace rg {

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
  // function-local static: initialised once, thread-safe (C++11); the driver entry point is
  // process-wide, not per device
  static const EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(sym);
    return static_cast<EncodeTiledFn>(nullptr);
  }();
  return fn;
}

// [rows, d] bf16 row-major -> box {64 columns, box_rows rows}, 128-byte swizzle, OOB rows = 0
static int make_rowmajor_map(CUtensorMap* map, const void* base, int64_t rows, int64_t d,
                             int box_rows = 128) {
  EncodeTiledFn enc = get_encoder();
  RG_REQUIRE(enc != nullptr, RECGURU_E_CUDA, "catalog: cuTensorMapEncodeTiled unavailable");
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                         gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RG_REQUIRE(r == CUDA_SUCCESS, RECGURU_E_CUDA, "catalog: cuTensorMapEncodeTiled failed (%d)",
             static_cast<int>(r));
  return RECGURU_OK;
}

static int pick_splits(int64_t U, int64_t n_items, int bn) {
  const int64_t m_tiles = ceil_div(U, kBlockM);
  const int64_t n_tiles = ceil_div(n_items, bn);
  if (m_tiles >= 2 * static_cast<int64_t>(num_sms())) return 1;
  int64_t s = ceil_div(2 * static_cast<int64_t>(num_sms()), m_tiles);
  const int64_t max_s = n_tiles / 8 > 0 ? n_tiles / 8 : 1;  // >= 8 tiles per split
  if (s > max_s) s = max_s;
  if (s > 64) s = 64;
  return static_cast<int>(s);
}

constexpr int kSweepWarpsDefault = 2;    // epilogue warps per TMEM lane quarter of the production sweep

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
// Items per accumulator hand-off in the top-k sweep: 128 (4 accumulators x 128 TMEM columns, 4 smem
// stages of 32 KB).  256 (2 x 256 columns) halves the hand-offs per flop but leaves room for only
// two 64 KB smem stages, which exposes the TMA latency: measured 46 ms vs 34 ms on cfg4
// (profiles/catalog_exp23.log); kept as a knob.
static int catalog_bn() {
#ifdef RECGURU_DEBUG_KNOBS
  const int bn = env_int("RECGURU_CATALOG_BN", 128);
  return bn == 256 ? 256 : 128;
#else
  return 128;
#endif
}

template <int D, int MODE, int NSETS, int CL, bool TS, int BN, int NST = 0, bool PIPE = false, bool RL = false>
static int launch_sweep_n(const CUtensorMap& mu, const CUtensorMap& mi, const CatalogParams& p,
                          cudaStream_t st) {
  const int smem = CatSmem<D, BN, NST>::bytes(MODE == kModeTopK ? p.k : 0, NSETS);
  RG_REQUIRE(smem <= 227 * 1024, RECGURU_E_UNSUPPORTED, "catalog: k=%d needs %d B of shared memory", p.k, smem);
  auto kern = catalog_sweep_kernel<D, MODE, NSETS, CL, TS, BN, NST, PIPE, false, RL>;
#ifdef RECGURU_DEBUG_KNOBS
  if (p.debug_flags) kern = catalog_sweep_kernel<D, MODE, NSETS, CL, TS, BN, NST, PIPE, true, RL>;
#endif
  // the dynamic shared-memory limit is a per-device function attribute
  static PerDeviceOnce once;
  const int dev = current_device();
  if (once.need(dev)) {
    cudaError_t e = cudaFuncSetAttribute(catalog_sweep_kernel<D, MODE, NSETS, CL, TS, BN, NST, PIPE, false, RL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
#ifdef RECGURU_DEBUG_KNOBS
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(catalog_sweep_kernel<D, MODE, NSETS, CL, TS, BN, NST, PIPE, true, RL>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
#endif
    RG_REQUIRE(e == cudaSuccess, RECGURU_E_CUDA, "catalog: cudaFuncSetAttribute: %s",
               cudaGetErrorString(e));
    once.done(dev);
  }
#ifdef RECGURU_DEBUG_KNOBS
  if (p.debug_flags & 512u) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_prof, z, sizeof(z));
  }
#endif
  cudaLaunchConfig_t cfg = {};
  const int64_t m_tiles = ceil_div(p.U, kBlockM);
  cfg.gridDim = dim3(static_cast<unsigned>(ceil_div(m_tiles, CL) * CL), static_cast<unsigned>(p.n_splits));
  cfg.blockDim = dim3(128 + NSETS * 128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mu, mi, p);
  RG_REQUIRE(e == cudaSuccess, RECGURU_E_CUDA, "catalog_sweep: launch failed: %s", cudaGetErrorString(e));
  RG_CHECK_LAUNCH("catalog_sweep");
#ifdef RECGURU_DEBUG_KNOBS
  if (p.debug_flags & 65536u) {
    static long long tr[kTraceTiles][8];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(tr, g_trace, sizeof(tr));
    const long long t0 = tr[0][2];
    fprintf(stderr, "[trace cta0] tile: mma_saw_empty mma_saw_full mma_issued | epi_saw_full epi_released | prod_saw_free prod_issued  (cycles rel. to first)\n");
    for (int i = 0; i < kTraceTiles; ++i)
      fprintf(stderr, "[trace] %4d: %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld\n", kTraceFirst + i, tr[i][0] - t0,
              tr[i][1] - t0, tr[i][2] - t0, tr[i][3] - t0, tr[i][4] - t0, tr[i][5] - t0, tr[i][6] - t0);
  }
  if (p.debug_flags & 512u) {
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
    fprintf(stderr, "[prof cta0] producer wait_empty %lld | mma wait_acc_empty %lld wait_full %lld issue %lld tiles %lld | "
            "epi set0 wait_full %lld work %lld tiles %lld | epi set1 wait_full %lld work %lld tiles %lld\n",
            h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10]);
  }
#endif
  return RECGURU_OK;
}

template <int D, int MODE, int KS, int NW, int CH>
static int launch_sweep2_n(const CUtensorMap& mu, const CUtensorMap& mi, const CatalogParams& p, cudaStream_t st) {
  const int smem = Cat2Smem<D>::bytes();
  auto kern = catalog_sweep2_kernel<D, MODE, KS, NW, CH>;
  static PerDeviceOnce once;          // the dynamic shared-memory limit is a per-device function attribute
  const int dev = current_device();
  if (once.need(dev)) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    RG_REQUIRE(e == cudaSuccess, RECGURU_E_CUDA, "catalog: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    once.done(dev);
  }
  const dim3 grid(static_cast<unsigned>(ceil_div(p.U, kBlockM)), static_cast<unsigned>(p.n_splits));
  kern<<<grid, 128 + NW * 128, smem, st>>>(mu, mi, p);
  RG_CHECK_LAUNCH("catalog_sweep2");
  return RECGURU_OK;
}

template <int D, int MODE, int NW, int CH>
static int launch_sweep2_k(const CUtensorMap& mu, const CUtensorMap& mi, const CatalogParams& p, cudaStream_t st) {
  if (MODE == kModeRank || p.k <= 8) return launch_sweep2_n<D, MODE, 8, NW, CH>(mu, mi, p, st);
  if (p.k <= 12) return launch_sweep2_n<D, MODE, 12, NW, CH>(mu, mi, p, st);
  return launch_sweep2_n<D, MODE, 16, NW, CH>(mu, mi, p, st);
}

// round-2 production sweep (catalog_sweep2.cuh): rank-only, or top-k with k <= 16 in KS register slots.
// RECGURU_CATALOG_WARPS = epilogue warps per TMEM lane quarter: 2 (128-column loads), 3 (64), 4 (32).
template <int D, int MODE>
static int launch_sweep2(const void* items, const CUtensorMap& mu, CatalogParams p, cudaStream_t st) {
  p.two = 2u;
  p.debug_flags = 0u;
  p.tiles_per_split = static_cast<int>(ceil_div(ceil_div(p.n_items, 128), p.n_splits));
  CUtensorMap mi;
  const int rc = make_rowmajor_map(&mi, items, p.n_items, D, 128);
  if (rc) return rc;
#ifdef RECGURU_DEBUG_KNOBS
  // 3 / 4 epilogue warps per lane quarter (64- / 32-column TMEM loads): measured SLOWER than 2 (cfg4: 24.4 /
  // 26.3 ms against 22.7; profiles/r02_catalog_warps_per_quarter.log) -- the sweep is bound by board power,
  // not by latency hiding, and the extra chunk loops are extra instructions.  Debug builds only.
  const int nw = env_int("RECGURU_CATALOG_WARPS", kSweepWarpsDefault);
  if (nw == 3) return launch_sweep2_k<D, MODE, 3, 64>(mu, mi, p, st);
  if (nw == 4) return launch_sweep2_k<D, MODE, 4, 32>(mu, mi, p, st);
#endif
  return launch_sweep2_k<D, MODE, 2, 128>(mu, mi, p, st);
}

// Round-1 form: 2 epilogue sets (8 warps, one 128-register TMEM load per tile), 128-item tiles, 4
// coupled smem/TMEM stages, top-k lists in registers for k <= 16.  Everything else that was
// measured without gain is compiled only with RECGURU_DEBUG_KNOBS.
template <int D, int MODE>
static int launch_sweep(const void* items, const CUtensorMap& mu, CatalogParams p, int bn, cudaStream_t st) {
  p.two = 2u;
  p.debug_flags = static_cast<uint32_t>(env_int("RECGURU_CATALOG_DEBUG", 0));
  RG_REQUIRE(kDebugKnobs || p.debug_flags == 0, RECGURU_E_UNSUPPORTED,
             "catalog: RECGURU_CATALOG_DEBUG needs a library built with RECGURU_DEBUG_KNOBS "
             "(python -m recguru_b200.build --debug-knobs)");
  p.tiles_per_split = static_cast<int>(ceil_div(ceil_div(p.n_items, bn), p.n_splits));
  CUtensorMap mi;
#ifdef RECGURU_DEBUG_KNOBS
  int nsets = env_int("RECGURU_CATALOG_SETS", 2);
  if (nsets != 2 && nsets != 4) nsets = 2;
  int cl = env_int("RECGURU_CATALOG_CLUSTER", 1);
  if (cl != 1 && cl != 2) cl = 1;
  if (p.U <= kBlockM) cl = 1;
  if (bn == 256) {
    // 2 epilogue sets <-> 2 accumulators of 256 columns; knobs for sets / cluster / TS apply to
    // the 128-item variant only
    int rc = make_rowmajor_map(&mi, items, p.n_items, D, 256);
    if (rc) return rc;
    return launch_sweep_n<D, MODE, 2, 1, false, 256>(mu, mi, p, st);
  }
  {
    int rc = make_rowmajor_map(&mi, items, p.n_items, D, 128 / cl);
    if (rc) return rc;
  }
  const bool ts = D == 128 && env_int("RECGURU_CATALOG_TS", 0) != 0;
  if (ts) {   // measured no faster than SS; 2 epilogue sets only
    if constexpr (D == 128) return launch_sweep_n<D, MODE, 2, 1, true, 128>(mu, mi, p, st);
  }
  if (nsets == 4) return cl == 2 ? launch_sweep_n<D, MODE, 4, 2, false, 128>(mu, mi, p, st) : launch_sweep_n<D, MODE, 4, 1, false, 128>(mu, mi, p, st);
  if (cl == 2) return launch_sweep_n<D, MODE, 2, 2, false, 128>(mu, mi, p, st);
  // RECGURU_CATALOG_STAGES=5|6: item-tile ring deeper than (and decoupled from) the accumulator ring
  const int nst = env_int("RECGURU_CATALOG_STAGES", 0);
  const int k_eff = MODE == kModeTopK ? p.k : 0;
  const bool pipe = MODE == kModeTopK && env_int("RECGURU_CATALOG_PIPE", 0) != 0;
  if (nst == 6 && CatSmem<D, 128, 6>::bytes(k_eff, 2) <= 227 * 1024)
    return pipe ? launch_sweep_n<D, MODE, 2, 1, false, 128, 6, MODE == kModeTopK>(mu, mi, p, st)
                : launch_sweep_n<D, MODE, 2, 1, false, 128, 6>(mu, mi, p, st);
  if (nst >= 5 && CatSmem<D, 128, 5>::bytes(k_eff, 2) <= 227 * 1024)
    return pipe ? launch_sweep_n<D, MODE, 2, 1, false, 128, 5, MODE == kModeTopK>(mu, mi, p, st)
                : launch_sweep_n<D, MODE, 2, 1, false, 128, 5>(mu, mi, p, st);
  if (pipe && p.k <= 16) return launch_sweep_n<D, MODE, 2, 1, false, 128, 0, MODE == kModeTopK, MODE == kModeTopK>(mu, mi, p, st);
  if (pipe) return launch_sweep_n<D, MODE, 2, 1, false, 128, 0, MODE == kModeTopK>(mu, mi, p, st);
  if (MODE == kModeTopK && p.k <= 16 && env_int("RECGURU_CATALOG_REGLIST", 1) == 0)
    return launch_sweep_n<D, MODE, 2, 1, false, 128>(mu, mi, p, st);
#else
  (void)bn;
  {
    int rc = make_rowmajor_map(&mi, items, p.n_items, D, 128);
    if (rc) return rc;
  }
#endif
  // k <= 16 keeps the top-k lists in registers; larger k uses shared-memory lists
  if (MODE == kModeTopK && p.k <= 16)
    return launch_sweep_n<D, MODE, 2, 1, false, 128, 0, false, MODE == kModeTopK>(mu, mi, p, st);
  return launch_sweep_n<D, MODE, 2, 1, false, 128>(mu, mi, p, st);
}

static int check_catalog(const char* who, const void* users, int64_t U, const void* items,
                         int64_t n_items, int64_t d) {
  RG_REQUIRE(users && items, RECGURU_E_SHAPE, "%s: null pointer", who);
  RG_REQUIRE(aligned16(users) && aligned16(items), RECGURU_E_SHAPE, "%s: operands must be 16-byte aligned", who);
  RG_REQUIRE(U > 0 && n_items > 0 && n_items < (1ll << 31) - 256 && U < (1ll << 31),
             RECGURU_E_SHAPE, "%s: bad sizes U=%lld n_items=%lld", who, (long long)U, (long long)n_items);
  RG_REQUIRE(d == 64 || d == 128, RECGURU_E_UNSUPPORTED, "%s: d=%lld not in {64,128}", who, (long long)d);
  return RECGURU_OK;
}

}  // namespace rg

extern "C" int64_t recguru_catalog_workspace_bytes(int64_t U, int64_t n_items, int64_t d,
                                                   int64_t k) {
  (void)d;
  if (U <= 0 || n_items <= 0 || k < 0 || k > rg::kMaxK) return RECGURU_E_SHAPE;
  const int s = rg::pick_splits(U, n_items, rg::catalog_bn());
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
  static PerDeviceOnce once;          // per-device function attribute
  const int dev = current_device();
  if (once.need(dev)) {
    cudaError_t e = cudaFuncSetAttribute(catalog_target_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(catalog_target_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    RG_REQUIRE(e == cudaSuccess, RECGURU_E_CUDA, "catalog_target: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    once.done(dev);
  }
  const __nv_bfloat16* items = static_cast<const __nv_bfloat16*>(items_bf16);
  if (d == 64)
    catalog_target_kernel<64><<<grid, 192, smem, st>>>(mu, items, id_lo, n_items, target, U, target_score);
  else
    catalog_target_kernel<128><<<grid, 192, smem, st>>>(mu, items, id_lo, n_items, target, U, target_score);
  RG_CHECK_LAUNCH("catalog_target");
  return RECGURU_OK;
}

static int catalog_topk_rank_impl(const void* users_bf16, int64_t U, const void* items_bf16,
                                  int64_t id_lo, int64_t n_items, int64_t d,
                                  const int64_t* target, const float* target_score,
                                  int64_t k, int tie_mode, float* topk_val, int64_t* topk_id,
                                  int64_t* rank_count, void* workspace,
                                  int64_t workspace_bytes, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(tie_mode >= RECGURU_TIE_LOWER_ID && tie_mode <= RECGURU_TIE_TARGET_LAST, RECGURU_E_SHAPE,
             "catalog_topk_rank: tie_mode=%d", tie_mode);
  int rc = check_catalog("catalog_topk_rank", users_bf16, U, items_bf16, n_items, d);
  if (rc) return rc;
  RG_REQUIRE(target && target_score && rank_count && (k == 0 || (topk_val && topk_id)), RECGURU_E_SHAPE,
             "catalog_topk_rank: null pointer");
  RG_REQUIRE(k >= 0 && k <= kMaxK, RECGURU_E_UNSUPPORTED, "catalog_topk_rank: k=%lld not in [0,%d]",
             (long long)k, kMaxK);
  CUtensorMap mu;
  rc = make_rowmajor_map(&mu, users_bf16, U, d);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CatalogParams p;
  p.U = U; p.n_items = n_items; p.id_lo = id_lo; p.target = target; p.tscore = target_score;
  p.k = static_cast<int>(k); p.dense = nullptr; p.two = 2u; p.tie_mode = tie_mode;
  p.users = static_cast<const __nv_bfloat16*>(users_bf16);
  const int bn = catalog_bn();
  p.n_splits = pick_splits(U, n_items, bn);
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
  // RECGURU_CATALOG_V=1: the round-1 kernel (A/B measurements); k > 16 (shared-memory lists) always uses it
  const bool v2 = k <= 16 && env_int("RECGURU_CATALOG_V", 2) != 1 && env_int("RECGURU_CATALOG_DEBUG", 0) == 0;
  RG_REQUIRE(v2 || tie_mode == RECGURU_TIE_LOWER_ID, RECGURU_E_UNSUPPORTED,
             "catalog_topk_rank: tie modes other than lower-id-first need k <= 16 (the round-2 sweep)");
  if (v2 && k == 0)
    rc = d == 64 ? launch_sweep2<64, kModeRank>(items_bf16, mu, p, st) : launch_sweep2<128, kModeRank>(items_bf16, mu, p, st);
  else if (v2)
    rc = d == 64 ? launch_sweep2<64, kModeTopK>(items_bf16, mu, p, st) : launch_sweep2<128, kModeTopK>(items_bf16, mu, p, st);
  else if (k == 0)   // rank only: no maxima, no lists -- what HR/NDCG/MRR need
    rc = d == 64 ? launch_sweep<64, kModeRank>(items_bf16, mu, p, 128, st) : launch_sweep<128, kModeRank>(items_bf16, mu, p, 128, st);
  else
    rc = d == 64 ? launch_sweep<64, kModeTopK>(items_bf16, mu, p, bn, st) : launch_sweep<128, kModeTopK>(items_bf16, mu, p, bn, st);
  if (rc) return rc;
  if (p.n_splits > 1)
    return recguru_topk_merge(p.out_val, p.out_id, p.out_cnt, p.n_splits, U, k, topk_val, topk_id,
                              rank_count, stream);
  return RECGURU_OK;
}

extern "C" int recguru_catalog_topk_rank(const void* users_bf16, int64_t U, const void* items_bf16,
                                         int64_t id_lo, int64_t n_items, int64_t d,
                                         const int64_t* target, const float* target_score,
                                         int64_t k, float* topk_val, int64_t* topk_id,
                                         int64_t* rank_count, void* workspace,
                                         int64_t workspace_bytes, recguru_stream_t stream) {
  return catalog_topk_rank_impl(users_bf16, U, items_bf16, id_lo, n_items, d, target, target_score, k,
                                RECGURU_TIE_LOWER_ID, topk_val, topk_id, rank_count, workspace, workspace_bytes,
                                stream);
}

extern "C" int recguru_catalog_topk_rank_ties(const void* users_bf16, int64_t U, const void* items_bf16,
                                              int64_t id_lo, int64_t n_items, int64_t d,
                                              const int64_t* target, const float* target_score,
                                              int64_t k, int tie_mode, float* topk_val, int64_t* topk_id,
                                              int64_t* rank_count, void* workspace,
                                              int64_t workspace_bytes, recguru_stream_t stream) {
  return catalog_topk_rank_impl(users_bf16, U, items_bf16, id_lo, n_items, d, target, target_score, k, tie_mode,
                                topk_val, topk_id, rank_count, workspace, workspace_bytes, stream);
}

extern "C" int recguru_catalog_scores_dense(const void* users_bf16, int64_t U,
                                            const void* items_bf16, int64_t n_items, int64_t d,
                                            float* scores, recguru_stream_t stream) {
  using namespace rg;
  int rc = check_catalog("catalog_scores_dense", users_bf16, U, items_bf16, n_items, d);
  if (rc) return rc;
  RG_REQUIRE(scores, RECGURU_E_SHAPE, "catalog_scores_dense: null pointer");
  CUtensorMap mu;
  rc = make_rowmajor_map(&mu, users_bf16, U, d);
  if (rc) return rc;
  CatalogParams p;
  memset(&p, 0, sizeof(p));
  p.U = U; p.n_items = n_items; p.k = 1; p.dense = scores; p.two = 2u; p.tie_mode = 0;
  p.users = static_cast<const __nv_bfloat16*>(users_bf16);
  p.n_splits = pick_splits(U, n_items, 128);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d == 64 ? launch_sweep<64, kModeDense>(items_bf16, mu, p, 128, st) : launch_sweep<128, kModeDense>(items_bf16, mu, p, 128, st);
}
