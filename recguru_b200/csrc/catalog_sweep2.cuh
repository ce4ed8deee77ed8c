// Catalog sweep, production form (round 2): the same data path as catalog_sweep_kernel in
// catalog_topk.cu -- TMA item-tile ring -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> epilogue
// warps that count ranks and keep register-resident top-k lists -- with the three changes the
// round-1 ncu source counters (profiles/r02_sweep_stall_breakdown.md) asked for:
//
//  1. Tile tickets instead of a static even/odd split.  Round 1 gave epilogue set e the tiles
//     j == e (mod 2).  A tile whose scores beat some lane's k-th best costs 3-5x a quiet tile, and
//     everything upstream is in order (one TMA stream, a 4-deep accumulator ring), so one slow set
//     starved the other within two tiles: the epilogue warps spent 22 % of their time waiting for an
//     accumulator while the MMA warps spent 31 % waiting for the epilogue.  Now the two warps that
//     share a TMEM lane quarter draw tile numbers from a per-quarter shared-memory counter: whoever is
//     free takes the next tile's quarter.  (Each tile's quarter is still processed exactly once, the
//     accumulator still needs four arrivals -- one per quarter -- to be handed back.)
//  2. A lean tile loop.  The round-1 loop head re-derived the shared-window base (S2UR
//     SR_CgaCtaId / SR_SWINHI, 64-bit tail arithmetic, LDCU of kernel parameters): ~50 instructions
//     and ~390 cycles per tile before the first useful one.  Here every address the loop needs is
//     formed once and pinned in a register behind an opaque move; the ragged-tail test is a 32-bit
//     compare against a precomputed tile index.
//  3. The smem item-tile ring is decoupled from the accumulator ring and one stage deeper (NST = 5):
//     the TMA for tile j+5 is issued as soon as tile j's MMAs retire, so a burst of slow tiles no
//     longer drains the operand pipeline behind it.
//  4. NW epilogue warps per lane quarter (2, 3 or 4) reading the accumulator row in chunks of CH
//     columns (128, 64, 32).  The tile loop is latency-, not pipe-bound (round-1 counters: ~0.4
//     instructions per cycle per scheduler with two epilogue warps on it, neither the FMA nor the
//     ALU pipe above 35 %): a third / fourth warp per scheduler hides the fixed-latency stalls of
//     the others; smaller chunks are what makes their register budget fit (setmaxnreg: 224 / 152 /
//     112 registers per epilogue thread).
// Top-k lists: KS register slots (the smallest of 8 / 12 / 16 that holds k), so an insertion is a
// KS-step compare-exchange chain instead of always 16.  At the end of the sweep the lists are
// exchanged through the (then idle) item-tile ring in shared memory and merged per user.
//
// Included by catalog_topk.cu (shares CatalogParams, CatSmem, kMode*, next_below, the ptx layer).
#pragma once

namespace rg {

constexpr int kV2Stages = 5;       // smem item-tile stages (decoupled from the 4 TMEM accumulators)

template <int D>
struct Cat2Smem {
  static constexpr int kKBlocks = D / 64;
  static constexpr int kATileBytes = kBlockM * D * 2;
  static constexpr int kTileBytes = 128 * D * 2;
  static constexpr int kOffA = 0;
  static constexpr int kOffB = kATileBytes;
  static constexpr int kOffBar = kOffB + kV2Stages * kTileBytes;
  // full[8], empty[8], acc_full[4], acc_empty[4], a_full
  static constexpr int kNumBars = 8 + 8 + 4 + 4 + 1;
  static constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr int kOffTicket = kOffTmemPtr + 8;        // uint32[4]: next tile of each lane quarter
  static constexpr int kOffDone = kOffTicket + 16;          // uint32[4 quarters][4 accumulators]: uses released
  static constexpr int kOffCnt = kOffDone + 64;             // int[4][128]
  static constexpr int kEnd = kOffCnt + 4 * 128 * 4;
  // the per-warp top-k lists (NW * k * 128 * 8 B <= 64 KB) alias the item-tile ring after the sweep
  static constexpr int kOffList = kOffB;
  static constexpr int kListRoom = kV2Stages * kTileBytes;
  __host__ __device__ static constexpr int bytes() { return kEnd + 1024; }
};

// a value the compiler must keep in a register (it would otherwise re-derive shared-window
// addresses from special registers inside the tile loop)
__device__ __forceinline__ uint32_t pin_u32(uint32_t x) {
  uint32_t y;
  asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// setmaxnreg budget.  The pool is what the CTA was LAUNCHED with ((4 + 4 NW) warps x the per-thread
// count ptxas reports under __launch_bounds__: 168 / 128 / 96), not the SM's register file: an
// increase that does not fit that pool blocks forever.  4 * ctl + 4 NW * epi <= (4 + 4 NW) * launch.
template <int NW> struct EpiRegs;
template <> struct EpiRegs<2> { static constexpr int ctl = 56, epi = 224, launch = 168; };
template <> struct EpiRegs<3> { static constexpr int ctl = 56, epi = 152, launch = 128; };
template <> struct EpiRegs<4> { static constexpr int ctl = 56, epi = 104, launch = 96; };

// MODE: kModeTopK (k <= KS, lists in registers) or kModeRank (k = 0).
// NW epilogue warps per TMEM lane quarter, accumulator rows read CH columns at a time.
template <int D, int MODE, int KS, int NW, int CH>
__global__ void __launch_bounds__(128 + NW * 128, 1)
catalog_sweep2_kernel(const __grid_constant__ CUtensorMap map_users,
                      const __grid_constant__ CUtensorMap map_items, CatalogParams p) {
  using L = Cat2Smem<D>;
  constexpr int BN = 128, NACC = 4, NSTG = kV2Stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + 8;
  uint64_t* bar_acc_full = bars + 16;
  uint64_t* bar_acc_empty = bars + 20;
  uint64_t* bar_a = bars + 24;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kOffTmemPtr);
  uint32_t* s_ticket = reinterpret_cast<uint32_t*>(smem + L::kOffTicket);
  uint32_t* s_done = reinterpret_cast<uint32_t*>(smem + L::kOffDone);
  int* s_cnt = reinterpret_cast<int*>(smem + L::kOffCnt);
  float* s_val = reinterpret_cast<float*>(smem + L::kOffList);
  int* s_id = reinterpret_cast<int*>(smem + L::kOffList + NW * p.k * 128 * 4);
  constexpr int kSlots = NW * 128;          // list entry i of slot s lives at [i * kSlots + s]
  static_assert(NW * 16 * 128 * 8 <= L::kListRoom, "top-k lists must fit the item-tile ring");
  static_assert(CH == 128 || CH == 64 || CH == 32, "chunk");
  static_assert(4 * EpiRegs<NW>::ctl + 4 * NW * EpiRegs<NW>::epi <= (4 + 4 * NW) * EpiRegs<NW>::launch,
                "setmaxnreg budget exceeds the registers the CTA is launched with");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kBlockM;
  const int n_items32 = static_cast<int>(p.n_items);               // < 2^31 (checked on the host)
  const int total_tiles = (n_items32 + BN - 1) / BN;
  const int tile_begin = blockIdx.y * p.tiles_per_split;
  const int n_tiles = max(0, min(total_tiles, tile_begin + p.tiles_per_split) - tile_begin);

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&map_users);
    ptx::tma_prefetch_desc(&map_items);
    for (int i = 0; i < NSTG; ++i) { ptx::mbar_init(&bar_full[i], 1); ptx::mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { ptx::mbar_init(&bar_acc_full[i], 1); ptx::mbar_init(&bar_acc_empty[i], 4); }
    ptx::mbar_init(bar_a, 1);
    ptx::mbar_fence_init();
  }
  if (threadIdx.x < 4) s_ticket[threadIdx.x] = 0u;
  if (threadIdx.x < 16) s_done[threadIdx.x] = 0u;
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) ptx::setmaxnreg_dec<EpiRegs<NW>::ctl>();
  if (warp == 0) {
    // ===================== TMA producer (whole warp loops; one elected lane issues) ========
    const bool leader = ptx::elect_one();
    if (leader) {
      ptx::mbar_arrive_expect_tx(bar_a, L::kATileBytes);
      for (int kb = 0; kb < L::kKBlocks; ++kb)
        ptx::tma_load_2d(&map_users, bar_a, smem + L::kOffA + kb * (kBlockM * 128), kb * 64,
                         static_cast<int32_t>(m0));
    }
    const uint32_t empty_addr = pin_u32(ptx::smem_u32(bar_empty));
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      ptx::mbar_wait_warp_addr(empty_addr + s * 8, ph ^ 1);       // stage free: its previous tile's MMAs retired
      if (leader) {
        ptx::mbar_arrive_expect_tx(&bar_full[s], L::kTileBytes);
        uint8_t* dst = smem + L::kOffB + s * L::kTileBytes;
        const int32_t n0 = (tile_begin + j) * BN;
        for (int kb = 0; kb < L::kKBlocks; ++kb)
          ptx::tma_load_2d(&map_items, &bar_full[s], dst + kb * (BN * 128), kb * 64, n0);
      }
      __syncwarp();
      if (++s == NSTG) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers (alternate tiles; one elected lane issues) ===========
    const bool leader = ptx::elect_one();
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(kBlockM, BN);
    const uint32_t a_addr = ptx::smem_u32(smem + L::kOffA);
    const uint32_t b_base = ptx::smem_u32(smem + L::kOffB);
    const uint32_t acc_empty_addr = pin_u32(ptx::smem_u32(bar_acc_empty));
    const uint32_t acc_full_addr = pin_u32(ptx::smem_u32(bar_acc_full));
    const uint32_t full_addr = pin_u32(ptx::smem_u32(bar_full));
    const uint32_t empty_addr = pin_u32(ptx::smem_u32(bar_empty));
    ptx::mbar_wait_warp(bar_a, 0);
    ptx::tc_fence_after();
    for (int j = warp - 1; j < n_tiles; j += 2) {
      const int s = j % NSTG;
      const uint32_t ph = (j / NSTG) & 1;
      const int a = j & (NACC - 1);
      const uint32_t aph = (j / NACC) & 1;
      ptx::mbar_wait_warp_addr(acc_empty_addr + a * 8, aph ^ 1);
      ptx::mbar_wait_warp_addr(full_addr + s * 8, ph);
      ptx::tc_fence_after();
      if (leader) {
        const uint32_t b_addr = b_base + s * L::kTileBytes;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t koff = (k / 4) * (kBlockM * 128) + (k % 4) * 32;   // bytes into the user tile
          const uint32_t koff_b = (k / 4) * (BN * 128) + (k % 4) * 32;      // bytes into the item tile
          ptx::umma_bf16(tmem_base + a * BN, ptx::umma_smem_desc_sw128(a_addr + koff),
                         ptx::umma_smem_desc_sw128(b_addr + koff_b), idesc, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit_addr(empty_addr + s * 8);        // smem stage free once these MMAs retire
        ptx::umma_commit_addr(acc_full_addr + a * 8);     // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // idle: fills the control warpgroup (setmaxnreg acts on aligned groups of four warps)
  } else {
    // ===================== epilogue: 4 TMEM lane quarters x NW warps ========================
    ptx::setmaxnreg_inc<EpiRegs<NW>::epi>();
    const int e = (warp - 4) >> 2;            // which of the quarter's NW warps (list slot only)
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int64_t u = m0 + row;
    const bool u_ok = u < p.U;
    const int K = p.k;
    const uint32_t two = p.two;
    const int slot = e * 128 + row;

    float t = 0.f;
    int64_t tgt_local = -1;                   // target's row index inside the slice
    if (u_ok) {
      t = p.tscore[u];
      if (t == 0.f) t = 0.f;                  // -0 -> +0 (the sign-bit count relies on it)
      tgt_local = p.target[u] - p.id_lo;
    }
    // ties: an item that scores exactly t counts towards the rank iff ... its id is lower (default: the
    // stable descending sort), never (RECGURU_TIE_TARGET_FIRST: the target wins every tie, what the
    // reference's column-0 layout gives it), always (RECGURU_TIE_TARGET_LAST: the pessimistic rank).
    // "counts" = compare >= t, done as > against the next float below t.
    const float t_below = next_below(t);
    const float t_left = p.tie_mode == RECGURU_TIE_TARGET_FIRST ? t : t_below;    // items left of the target
    const float t_right = p.tie_mode == RECGURU_TIE_TARGET_LAST ? t_below : t;    // items right of it
    int tgt_tile, tgt_col = -1;               // tile relative to this device's slice
    if (tgt_local < 0) tgt_tile = -1;
    else if (tgt_local >= p.n_items) tgt_tile = INT_MAX;
    else { tgt_tile = static_cast<int>(tgt_local / BN); tgt_col = static_cast<int>(tgt_local % BN); }

    int cnt = 0;
    float thr = -CUDART_INF_F;
    // KS register slots, descending; the k live entries sit at the BOTTOM (slots KS-k .. KS-1), the
    // slots above them are blocked with +inf, so the k-th best -- the insertion threshold -- is
    // always lv[KS-1] and an insertion is a fully unrolled compare-exchange chain.
    float lv[KS];
    int li[KS];
#pragma unroll
    for (int i = 0; i < KS; ++i) { lv[i] = i < KS - K ? CUDART_INF_F : -CUDART_INF_F; li[i] = -1; }
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t cf[2] = {0u, 0u}, cf2[2] = {0u, 0u};   // sign-bit counters fed by the FMA pipe (IMAD.HI)
    uint32_t ca[2] = {0u, 0u}, ca2[2] = {0u, 0u};   // sign-bit counters fed by the ALU pipe (LEA.HI)

    // sorted insert (desc score; equal scores keep arrival = id order)
    auto insert = [&](float v, int id) {
      bool sw = false;
#pragma unroll
      for (int i = 0; i < KS; ++i) {
        // strict: an equal score stays below earlier ones; once the new entry is placed every later
        // slot just shifts down (a carried entry must not be compared again, or it would slip below
        // an equal neighbour and break the id order of ties)
        sw = sw || (v > lv[i]);
        const float ov = lv[i];
        const int oi = li[i];
        lv[i] = sw ? v : ov;
        li[i] = sw ? id : oi;
        v = sw ? ov : v;
        id = sw ? oi : id;
      }
      thr = lv[KS - 1];
    };
    // Offer eight scores (columns id0 .. id0+7, bit c of `ok` = column c is a real item) to the list:
    // lanes queue the columns that beat their threshold and the warp drains the queues together.
    auto offer8 = [&](const uint32_t (&w8)[8], int id0, uint32_t ok) {
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
    };

    const uint32_t acc_full_addr = pin_u32(ptx::smem_u32(bar_acc_full));
    const uint32_t acc_empty_addr = pin_u32(ptx::smem_u32(bar_acc_empty));
    const uint32_t ticket_addr = pin_u32(ptx::smem_u32(&s_ticket[quarter]));
    const uint32_t done_addr = pin_u32(ptx::smem_u32(&s_done[quarter * 4]));
    const uint32_t tmem_row = pin_u32(tmem_base + lane_base);
    const int full_tiles = n_items32 / BN - tile_begin;          // local tiles [0, full_tiles) are complete
    const int tgt_j = tgt_tile == INT_MAX ? INT_MAX : tgt_tile - tile_begin;   // local index of the target's tile

    for (;;) {
      // next tile of this lane quarter: whichever of the quarter's two warps is free takes it
      int j = 0;
      if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(j) : "r"(ticket_addr) : "memory");
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= n_tiles) break;
      const int a = j & (NACC - 1);
      const uint32_t aph = (j >> 2) & 1;
      if constexpr (NW > 2) {
        // A parity wait cannot tell "use n of this accumulator is full" from "use n-1 is not full yet".
        // With two warps per quarter the tickets in flight are never four tiles apart; with more, a
        // warp can draw tile j while a sibling still waits for tile j-4 (the two MMA warps run the
        // even and the odd tiles independently).  So: this quarter must have RELEASED use n-1 of the
        // accumulator before anyone waits for use n.
        if (lane == 0) {
          const uint32_t want = static_cast<uint32_t>(j >> 2);
          uint32_t got;
          do {
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(got) : "r"(done_addr + a * 4) : "memory");
          } while (got != want);
        }
        __syncwarp();
      }
      ptx::mbar_wait_warp_addr(acc_full_addr + a * 8, aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_row + a * BN;
      const int n0 = (tile_begin + j) * BN;                       // slice-local index of column 0

      if (__any_sync(0xffffffffu, j == tgt_j) || j >= full_tiles) {
        // exact tile (warp-uniform): some lane's target lives here and/or the tile runs past the
        // end of the slice.  Per-column rules, 8 columns at a time.
        const int valid = min(BN, n_items32 - n0);
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
              if (j == tgt_j) {
                if (col < tgt_col) cnt += (v > t_left) ? 1 : 0;
                else if (col > tgt_col) cnt += (v > t_right) ? 1 : 0;
              } else {
                cnt += (v > (j < tgt_j ? t_left : t_right)) ? 1 : 0;
              }
            }
          }
          if constexpr (MODE == kModeTopK)
            offer8(w, n0 + c0, valid - c0 >= 8 ? 0xffu : (valid > c0 ? (1u << (valid - c0)) - 1u : 0u));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive_addr(acc_empty_addr + a * 8);
          if constexpr (NW > 2)
            asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(done_addr + a * 4), "r"((j >> 2) + 1) : "memory");
        }
        continue;
      }

      // fast tile: the accumulator row in BN/CH TMEM loads; after the last one the scores live in
      // registers only, so the accumulator goes back to the MMA warps before its arithmetic
      const float tcmp = j < tgt_j ? t_left : t_right;
      uint64_t tt;
      asm("mov.b64 %0, {%1, %1};" : "=l"(tt) : "f"(tcmp));
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t vr[CH];
        ptx::tmem_ld_cols<CH>(taddr + c0, vr);
        ptx::tmem_ld_wait();
        if (c0 + CH == BN) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive_addr(acc_empty_addr + a * 8);
            if constexpr (NW > 2)
              asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(done_addr + a * 4), "r"((j >> 2) + 1) : "memory");
          }
        }
        if constexpr (MODE == kModeTopK) {
          // 8-column group maxima (four FMNMX3 each) decide whether the list must be touched
          constexpr int NG = CH / 8;
          float gm[NG];
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const float* f = reinterpret_cast<const float*>(&vr[g * 8]);
            const float m1 = fmaxf(fmaxf(f[0], f[1]), f[2]);
            const float m2 = fmaxf(fmaxf(f[3], f[4]), f[5]);
            const float m3 = fmaxf(fmaxf(f[6], f[7]), m1);
            gm[g] = fmaxf(m2, m3);
          }
          float mx;
          if constexpr (NG == 16) {
            float r[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = fmaxf(gm[q], gm[q + 8]);
            const float a3 = fmaxf(fmaxf(r[0], r[1]), r[2]);
            const float b3 = fmaxf(fmaxf(r[3], r[4]), r[5]);
            mx = fmaxf(fmaxf(r[6], r[7]), fmaxf(a3, b3));
          } else if constexpr (NG == 8) {
            const float a3 = fmaxf(fmaxf(gm[0], gm[1]), gm[2]);
            const float b3 = fmaxf(fmaxf(gm[3], gm[4]), gm[5]);
            mx = fmaxf(fmaxf(gm[6], gm[7]), fmaxf(a3, b3));
          } else {
            mx = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
          }
          if (__any_sync(0xffffffffu, mx > thr)) {
            // per-lane bitmask of the groups that beat the threshold; the warp walks the union and
            // every lane checks that group's eight scores, fetched from its registers via a jump table
            uint32_t gmask = 0;
#pragma unroll
            for (int g = 0; g < NG; ++g) gmask |= (gm[g] > thr ? 1u : 0u) << g;
            uint32_t wmask = __reduce_or_sync(0xffffffffu, gmask);
            while (wmask) {
              const int g = __ffs(wmask) - 1;
              wmask &= wmask - 1;
              uint32_t w8[8];
#define RG_GRP(G) case G: if constexpr ((G) < NG) { _Pragma("unroll") for (int c = 0; c < 8; ++c) w8[c] = vr[((G) < NG ? (G) : 0) * 8 + c]; } break;
              switch (g) {
                RG_GRP(0) RG_GRP(1) RG_GRP(2) RG_GRP(3) RG_GRP(4) RG_GRP(5) RG_GRP(6) RG_GRP(7)
                RG_GRP(8) RG_GRP(9) RG_GRP(10) RG_GRP(11) RG_GRP(12) RG_GRP(13) RG_GRP(14)
                default: _Pragma("unroll") for (int c = 0; c < 8; ++c) w8[c] = vr[(NG - 1) * 8 + c]; break;
              }
#undef RG_GRP
              offer8(w8, n0 + c0 + g * 8, 0xffu);
            }
          }
        }
        // rank count: d = tcmp - v has its sign bit set iff v > tcmp (exact: IEEE subtraction is zero
        // only for equal operands, tcmp is never -0).  FADD2 forms two differences per instruction; the
        // sign bits are accumulated half on the FMA pipe (IMAD.HI by a run-time 2) and half on the ALU
        // pipe (LEA.HI).  Tiles left of the target's compare >= (as > against the next float below).
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
    cnt += static_cast<int>(cf[0] + cf[1] + ca[0] + ca[1] + cf2[0] + cf2[1] + ca2[0] + ca2[1]);

    s_cnt[e * 128 + row] = cnt;
    // every epilogue warp has drawn a ticket past the last tile: all accumulators were seen full, so
    // every MMA has retired and nothing reads the item-tile ring any more -- it carries the lists now
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 128) : "memory");
    if constexpr (MODE == kModeTopK) {
#pragma unroll
      for (int q = 0; q < KS; ++q) {
        if (q >= KS - K) {
          s_val[(q - (KS - K)) * kSlots + slot] = lv[q];
          s_id[(q - (KS - K)) * kSlots + slot] = li[q];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NW * 128) : "memory");
    }
    if (e == 0 && u_ok) {
      if constexpr (MODE == kModeTopK) {
        // merge the NW lists of this row: (score desc, id asc); an empty entry has id -1
        int head[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) head[q] = 0;
        const int64_t out = (static_cast<int64_t>(blockIdx.y) * p.U + u) * K;
        for (int i = 0; i < K; ++i) {
          float bv = -CUDART_INF_F;
          int bd = -1, bq = -1;
#pragma unroll
          for (int q = 0; q < NW; ++q) {
            if (head[q] < K) {
              const float v = s_val[head[q] * kSlots + q * 128 + row];
              const int id = s_id[head[q] * kSlots + q * 128 + row];
              if (id >= 0 && (bq < 0 || v > bv || (v == bv && id < bd))) { bv = v; bd = id; bq = q; }
            }
          }
#pragma unroll
          for (int q = 0; q < NW; ++q)
            if (q == bq) ++head[q];
          p.out_val[out + i] = bd >= 0 ? bv : -CUDART_INF_F;
          p.out_id[out + i] = bd >= 0 ? p.id_lo + bd : -1;
        }
      }
      int total = 0;
#pragma unroll
      for (int q = 0; q < NW; ++q) total += s_cnt[q * 128 + row];
      p.out_cnt[static_cast<int64_t>(blockIdx.y) * p.U + u] = total;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace rg
