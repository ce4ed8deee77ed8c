// Piece 3 (K3): sampled-negative scoring as one fused gather-dot-loss kernel (+ backward).
//
// Replaces, per training step,
//   src_emb(dec_out).view(-1,L,1,d) ; src_emb(n_items).view(-1,L,K,d) ; 2x batched matmul ;
//   loss module
//   /root/reference/GURU/AutoEnc4Rec.py:115-131,:219-227
//   /root/reference/GURU/AutoEnc4Rec_cross.py:203-215 ; tools/utils.py:103-126
//   /root/reference/GURU/tools/lossfunctions.py:36-49 (SampledCrossEntropyLoss, label 0),
//                                               :56-72 (BPRLoss), :79-96 (BPRLoss_sas)
//
// HBM-bound: per position 4d B (h) + (1+K)(8 + 4d) B (ids + rows) + 4 B mask read,
// 4(1+K) B written; 2d(1+K) flop.  The [B,L,K,d] gathered-row intermediate of the
// reference is never materialised.  One warp per position; rows narrower than a warp's
// 512 B (d < 128) are handled by sub-warp groups so that all 32 lanes keep loading.
// Loss reduction is a fixed-order two-level sum (warp -> CTA -> one finalising CTA), so
// the scalar is bit-reproducible run to run.
#include <math_constants.h>

#include "common.cuh"

namespace rg {

constexpr int kScoreWarps = 8;
constexpr int kPosPerWarp = 4;
constexpr int kPosPerBlock = kScoreWarps * kPosPerWarp;

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }

// Multi-value butterfly: every lane of a LANES-wide group holds partial sums for UN rows; after
// log2(UN) halving exchanges + log2(LANES/UN) plain steps, lane (sub) holds the full sum of row
// (sub / (LANES/UN)).  UN + log2(LANES/UN) - 1 shuffles instead of UN * log2(LANES).
template <int LANES, int UN>
__device__ __forceinline__ float group_multi_reduce(float (&v)[UN], int sub) {
#pragma unroll
  for (int n = UN, off = LANES / 2; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (sub & off) != 0;
#pragma unroll
    for (int k = 0; k < n / 2; ++k) {
      const float send = upper ? v[k] : v[k + n / 2];
      const float keep = upper ? v[k + n / 2] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  float r = v[0];
#pragma unroll
  for (int off = LANES / UN / 2; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

// Loss + unnormalised derivatives of one position from its 1+K scores (row[0] = positive).
// Overwrites row[] with d l / d score.
__device__ __forceinline__ float position_loss(int kind, float* row, int K) {
  const float p = row[0];
  if (kind == RECGURU_LOSS_SCE) {
    float m = p;
    for (int k = 1; k <= K; ++k) m = fmaxf(m, row[k]);
    float sum = 0.f;
    for (int k = 0; k <= K; ++k) sum += expf(row[k] - m);
    for (int k = 0; k <= K; ++k) row[k] = expf(row[k] - m) / sum - (k == 0 ? 1.f : 0.f);
    return (m + logf(sum)) - p;
  }
  float nsum = 0.f;
  for (int k = 1; k <= K; ++k) nsum += row[k];
  const float nbar = nsum / static_cast<float>(K);
  float l, dp, dn;
  if (kind == RECGURU_LOSS_BPR) {
    const float sg = sigmoidf_ref(p - nbar);
    l = -logf(sg);
    dp = sg - 1.f;
    dn = (1.f - sg) / static_cast<float>(K);
  } else {  // RECGURU_LOSS_BPR_SAS
    const float a = sigmoidf_ref(p), b = sigmoidf_ref(nbar);
    const float ya = a + 1e-24f, yb = (1.f - b) + 1e-24f;
    l = -(logf(ya) + logf(yb));
    dp = -(a * (1.f - a)) / ya;
    dn = (b * (1.f - b)) / yb / static_cast<float>(K);
  }
  row[0] = dp;
  for (int k = 1; k <= K; ++k) row[k] = dn;
  return l;
}

constexpr int kRowPad = 33;   // 1+K <= 32 scores, +1 against bank conflicts
constexpr int kDensePW = 16;  // positions per warp (measured best of 8/16/32)

// Dot products + butterfly reduction of one (position, 8-row step) work item; the lane that
// ends up with row c's sum stores it to dst[c].
template <int VEC>
__device__ __forceinline__ void score_item(const float4 (&hv)[(VEC + 31) / 32],
                                           const float4 (&w)[VEC < 32 ? 8 * VEC / 32 : 8][(VEC + 31) / 32],
                                           int c0, int C, int g, int sub, float* dst) {
  constexpr int LANES = VEC < 32 ? VEC : 32;
  constexpr int G = 32 / LANES;
  constexpr int VPL = VEC / LANES;
  constexpr int UN = 8 / G;
  float acc[UN];
#pragma unroll
  for (int u = 0; u < UN; ++u) {
    float a = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      a = fmaf(hv[v].x, w[u][v].x, a);
      a = fmaf(hv[v].y, w[u][v].y, a);
      a = fmaf(hv[v].z, w[u][v].z, a);
      a = fmaf(hv[v].w, w[u][v].w, a);
    }
    acc[u] = a;
  }
  const float r = group_multi_reduce<LANES, UN>(acc, sub);
  constexpr int kRep = LANES / UN;          // lanes holding the same row's sum
  const int c = c0 + (sub / kRep) * G + g;  // row (candidate) this lane ended up with
  if (sub % kRep == 0 && c < C) dst[c] = r;
}

// Scores + loss, every position scored.  One warp owns 16 consecutive positions.
// Phase 1 (per position, whole warp): gather the 1+K table rows (8 rows in flight; sub-warp
// groups when a row is narrower than 512 B), dot with h, reduce all partial sums at once with a
// multi-value butterfly, park the scores in shared memory.  Phase 2 (lane = position): loss and
// derivatives with lane-parallel transcendental math, every output written coalesced;
// fixed-order warp -> CTA reduction of the loss.  40 registers, 6 CTAs/SM: the bandwidth-bound
// case wants occupancy.
template <int VEC>  // float4 per row
__global__ void __launch_bounds__(kScoreWarps * 32)
sampled_score_fwd_kernel(const float4* __restrict__ h, const float4* __restrict__ weight,
                         const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                         int64_t n_pos, int K, const float* __restrict__ mask, int kind,
                         float* __restrict__ p_out, float* __restrict__ n_out,
                         float* __restrict__ dp_unit, float* __restrict__ dn_unit,
                         float2* __restrict__ block_partial) {
  rg::pdl_wait();
  constexpr int LANES = VEC < 32 ? VEC : 32;  // lanes cooperating on one row
  constexpr int G = 32 / LANES;               // row groups per warp step
  constexpr int VPL = VEC / LANES;            // float4 per lane per row
  constexpr int UN = 8 / G;                   // rows per group per step: 8 rows in flight per warp
  constexpr int PW = kDensePW;
  __shared__ float s_sc[kScoreWarps][PW][kRowPad];
  __shared__ float2 s_part[kScoreWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / LANES, sub = lane % LANES;
  const int C = K + 1;
  const int64_t base = (static_cast<int64_t>(blockIdx.x) * kScoreWarps + warp) * PW;
  const int n_here = static_cast<int>(rg::min64(PW, n_pos - base));  // may be <= 0

  for (int q = 0; q < n_here; ++q) {
    const int64_t i = base + q;
    int64_t my_id = 0;
    if (lane == 0) my_id = pos[i];
    else if (lane <= K) my_id = neg[i * K + (lane - 1)];
    float4 hv[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) hv[v] = ld_stream_f4(h + i * VEC + v * LANES + sub);
    for (int c0 = 0; c0 < C; c0 += 8) {
      float4 w[UN][VPL];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int c = c0 + u * G + g;
        const int64_t id = __shfl_sync(0xffffffffu, my_id, c < C ? c : 0);
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          w[u][v] = c < C ? __ldg(weight + id * VEC + v * LANES + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      score_item<VEC>(hv, w, c0, C, g, sub, s_sc[warp][q]);
    }
  }
  __syncwarp();

  // ---- phase 2: lane <-> position ------------------------------------------------------
  const int64_t i = base + lane;
  const bool ok = lane < n_here;
  float* row = s_sc[warp][lane < PW ? lane : 0];
  if (p_out && ok) p_out[i] = row[0];
  if (n_out) {
    const int tot = max(n_here, 0) * K;
    for (int idx = lane; idx < tot; idx += 32) {
      const int q = idx / K, k = idx - q * K;
      n_out[base * K + idx] = s_sc[warp][q][1 + k];
    }
  }
  float lm = 0.f, m = 0.f;
  if (kind >= 0) {
    __syncwarp();
    if (ok) {
      m = mask ? __ldg(mask + i) : 1.f;
      lm = position_loss(kind, row, K) * m;
      if (dp_unit) dp_unit[i] = row[0] * m;
#pragma unroll 1
      for (int k = 1; k <= K; ++k) row[k] *= m;
    }
    __syncwarp();
    if (dn_unit) {
      const int tot = max(n_here, 0) * K;
      for (int idx = lane; idx < tot; idx += 32) {
        const int q = idx / K, k = idx - q * K;
        dn_unit[base * K + idx] = s_sc[warp][q][1 + k];
      }
    }
    // fixed-order reduction: lanes by shuffle tree, warps sequentially
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lm += __shfl_xor_sync(0xffffffffu, lm, o);
      m += __shfl_xor_sync(0xffffffffu, m, o);
    }
    if (lane == 0) s_part[warp] = make_float2(lm, m);
    __syncthreads();
    if (threadIdx.x == 0) {
      float2 t = make_float2(0.f, 0.f);
#pragma unroll
      for (int w = 0; w < kScoreWarps; ++w) { t.x += s_part[w].x; t.y += s_part[w].y; }
      block_partial[blockIdx.x] = t;
    }
  }
  rg::pdl_trigger();
}

constexpr int kSparseThreadsDefault = 1024;
// Padding-aware variant (RECGURU_LOSS_SKIP_MASKED): positions under the pad mask are not scored.
// On a padded batch (cfg2's Amazon-shaped one: 7 % of the positions are live, in runs of ~15 at
// the tail of each sequence) the dense kernel's layout leaves the live work in a few warps, each
// with a serial chain of dependent gathers.  Here one persistent 1024-thread CTA per SM owns a
// contiguous range of positions and walks it in chunks:
//   phase 0  thread = position: read the mask, stable compaction of the live positions into a
//            CTA work list (ballot + warp-count scan), zeros written for the dead positions'
//            logits/derivatives (coalesced), ids of the live ones staged as int32;
//   phase 1  the 32 warps drain the work list round-robin: item = (live position, 8-row step),
//            whole warp per item as in the dense kernel -- every SM has ~100 items for 32 warps;
//   phase 2  thread = live position: loss + derivatives, outputs scattered to their positions.
// Loss partials are reduced in a fixed order (lane tree, warps, chunks) => bit-reproducible.
// NT = threads per CTA: 1024 (one CTA per SM), 512 or 256 (two / four CTAs per SM: the block-wide
// phase barriers of one CTA overlap the gathers of its neighbours).
template <int VEC, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT)
sampled_score_fwd_sparse_kernel(const float4* __restrict__ h, const float4* __restrict__ weight,
                                const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                                int64_t n_pos, int K, const float* __restrict__ mask, int kind,
                                float* __restrict__ p_out, float* __restrict__ n_out,
                                float* __restrict__ dp_unit, float* __restrict__ dn_unit,
                                float2* __restrict__ block_partial, int chunk, int64_t pos_per_cta,
                                float4* __restrict__ dh, const float* __restrict__ dh_scale,
                                const uint32_t* __restrict__ live_count) {
  rg::pdl_wait();
  constexpr int LANES = VEC < 32 ? VEC : 32;
  constexpr int G = 32 / LANES;
  constexpr int VPL = VEC / LANES;
  constexpr int UN = 8 / G;
  constexpr int kWarps = NT / 32;
  extern __shared__ float s_dyn[];
  const int C = K + 1;
  const int RS = C | 1;                                       // odd row stride: no bank conflicts
  float* s_sc = s_dyn;                                        // [chunk][RS] scores of the live list
  int* s_id = reinterpret_cast<int*>(s_sc + chunk * RS);      // [chunk][RS] ids of the live list
  int* s_rank = s_id + chunk * RS;                            // [chunk] position -> list slot, -1 = dead
  int* s_list = s_rank + chunk;                               // [chunk] list slot -> position
  __shared__ int s_wcnt[kWarps];
  __shared__ float2 s_part[kWarps];
  __shared__ int s_nlive;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane / LANES, sub = lane % LANES;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.x) * pos_per_cta;
  const int64_t r_end = rg::min64(n_pos, r_begin + pos_per_cta);
  float tot_l = 0.f, tot_m = 0.f;                             // thread 0: running CTA partial

  for (int64_t cb = r_begin; cb < r_end; cb += chunk) {
    const int n_c = static_cast<int>(rg::min64(chunk, r_end - cb));
    // ---- phase 0: compaction of the live positions (stable: list order = position order) ----
    int my_slot = -1;
    {
      // chunk <= 1024 = one position per thread
      const bool live = tid < n_c && __ldg(mask + cb + tid) != 0.f;
      const uint32_t bal = __ballot_sync(0xffffffffu, live);
      if (lane == 0) s_wcnt[warp] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const int c = s_wcnt[w];
        before += w < warp ? c : 0;
        total += c;
      }
      if (tid == 0) s_nlive = total;
      if (live) {
        my_slot = before + __popc(bal & ((1u << lane) - 1u));
        s_list[my_slot] = tid;
        s_id[my_slot * RS] = static_cast<int>(pos[cb + tid]);
      }
      if (tid < n_c) {
        s_rank[tid] = my_slot;
        if (!live) {
          if (p_out) p_out[cb + tid] = 0.f;
          if (dp_unit && kind >= 0) dp_unit[cb + tid] = 0.f;
        }
      }
      __syncthreads();
      // negatives: coalesced sweep over the chunk's [n_c, K] id block
      const int tot = n_c * K;
      for (int idx = tid; idx < tot; idx += NT) {
        const int q = idx / K, k = idx - q * K;
        const int slot = s_rank[q];
        if (slot >= 0) {
          s_id[slot * RS + 1 + k] = static_cast<int>(neg[cb * K + idx]);
        } else {
          if (n_out) n_out[cb * K + idx] = 0.f;
          if (dn_unit && kind >= 0) dn_unit[cb * K + idx] = 0.f;
        }
      }
      __syncthreads();
    }
    const int n_live = s_nlive;
    // ---- phase 1: warps drain the work list ------------------------------------------------
    const int steps = (C + 7) >> 3;
    const int n_items = n_live * steps;
    for (int t = warp; t < n_items; t += kWarps) {
      const int slot = t / steps;
      const int c0 = (t - slot * steps) << 3;
      const int64_t i = cb + s_list[slot];
      float4 hv[VPL];
#pragma unroll
      for (int v = 0; v < VPL; ++v) hv[v] = ld_stream_f4(h + i * VEC + v * LANES + sub);
      float4 w[UN][VPL];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int c = c0 + u * G + g;
        const int64_t id = s_id[slot * RS + (c < C ? c : 0)];
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          w[u][v] = c < C ? __ldg(weight + id * VEC + v * LANES + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      score_item<VEC>(hv, w, c0, C, g, sub, s_sc + slot * RS);
    }
    // fused dh: the rows of the dead positions are zero (coalesced 16-byte streaming stores).  Issued
    // HERE, behind this warp's gathers: ahead of them (where they sat in round 1) the 32 stores per
    // thread filled the load/store queue in front of the dependent row gathers of phase 1; now the
    // warps that drew few or no work items stream zeros while the others wait for their rows.
    // (Streaming them from a side kernel instead was also measured: no gain.)
    if (dh != nullptr) {
      const int tot4 = n_c * VEC;
      for (int idx = tid; idx < tot4; idx += NT) {
        const int q = idx / VEC;
        if (s_rank[q] < 0)
          st_stream_f4(dh + (cb + q) * VEC + (idx - q * VEC), make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    __syncthreads();
    // ---- phase 2: thread = live position ---------------------------------------------------
    float lm = 0.f, m = 0.f;
    if (tid < n_live) {
      float* row = s_sc + tid * RS;
      const int64_t i = cb + s_list[tid];
      if (p_out) p_out[i] = row[0];
      if (n_out)
        for (int k = 0; k < K; ++k) n_out[i * K + k] = row[1 + k];
      if (kind >= 0) {
        m = __ldg(mask + i);
        lm = position_loss(kind, row, K) * m;
        if (dp_unit) dp_unit[i] = row[0] * m;
        if (dn_unit)
          for (int k = 0; k < K; ++k) dn_unit[i * K + k] = row[1 + k] * m;
      }
    }
    if (kind >= 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lm += __shfl_xor_sync(0xffffffffu, lm, o);
        m += __shfl_xor_sync(0xffffffffu, m, o);
      }
      if (lane == 0) s_part[warp] = make_float2(lm, m);
    }
    __syncthreads();                                // s_part complete; s_sc/s_id/s_list reusable
    if (kind >= 0 && tid == 0) {
#pragma unroll 1
      for (int w = 0; w < kWarps; ++w) { tot_l += s_part[w].x; tot_m += s_part[w].y; }
    }
    // ---- phase 3 (fused dh): dh[i,:] = s * sum_c (d_c * m) * W[id_c,:] for the live positions ----
    // The unit derivatives are still in s_sc, the ids in s_id, the table rows just read are L1/L2
    // hits.  Same coefficient products, same fma chain and group combine as
    // sampled_score_bwd_dh_kernel: the result is bit-identical to the two-kernel path.
    if (dh != nullptr) {
      float sc = dh_scale ? __ldg(dh_scale) : 1.f;
      sc *= 1.f / static_cast<float>(__ldg(live_count));      // == loss_out[1] of loss_finalize_kernel
      for (int t = warp; t < n_live; t += kWarps) {
        const int64_t i = cb + s_list[t];
        const float m = __ldg(mask + i);
        float4 acc[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c0 = 0; c0 < C; c0 += G * UN) {
          float4 w[UN][VPL];
          float cf[UN];
#pragma unroll
          for (int u = 0; u < UN; ++u) {
            const int c = c0 + u * G + g;
            const int src = c < C ? c : 0;
            const int64_t id = s_id[t * RS + src];
            cf[u] = c < C ? (s_sc[t * RS + src] * m) * sc : 0.f;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
              w[u][v] = c < C ? __ldg(weight + id * VEC + v * LANES + sub)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              acc[v].x = fmaf(cf[u], w[u][v].x, acc[v].x);
              acc[v].y = fmaf(cf[u], w[u][v].y, acc[v].y);
              acc[v].z = fmaf(cf[u], w[u][v].z, acc[v].z);
              acc[v].w = fmaf(cf[u], w[u][v].w, acc[v].w);
            }
        }
#pragma unroll
        for (int o = 16; o >= LANES; o >>= 1)
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
            acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
            acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
            acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
          }
        if (g == 0) {
#pragma unroll
          for (int v = 0; v < VPL; ++v) st_stream_f4(dh + i * VEC + v * LANES + sub, acc[v]);
        }
      }
      __syncthreads();                              // s_sc/s_id/s_list are rewritten by the next chunk
    }
  }
  if (kind >= 0 && tid == 0) block_partial[blockIdx.x] = make_float2(tot_l, tot_m);
  rg::pdl_trigger();
}

// The padding-aware scoring when the step already owns the list of live positions (the row-mode
// compaction of the table-backward plan, recguru_embedding_bwd_plan_live_rows): no per-CTA
// compaction and no block-wide phase.  One warp per live position, start to finish: ids, the h row
// and the 1+K table rows (8 rows in flight), scores, loss + unit derivatives (lane 0; a handful of
// transcendentals), dh from the table rows still in registers (1+K <= 8) or re-read from L1.  Same
// dot-product tree, same loss routine, same coefficient products and fma chain as the compacting
// kernel above => bit-identical outputs.  After its live positions every thread streams zeros into
// the outputs of the dead ones.  The loss partials are per warp (positions t = warp, warp + W, ..),
// reduced in a fixed order.
constexpr int kRowsWarps = 8;
// ONE: 1+K <= 8 (one 8-row step per position: no step loop, the table rows stay in registers for dh)
template <int VEC, bool ONE>
__global__ void __launch_bounds__(kRowsWarps * 32, VEC <= 32 ? 4 : 2)
sampled_score_rows_kernel(const float4* __restrict__ h, const float4* __restrict__ weight,
                          const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                          int64_t n_pos, int K, const float* __restrict__ mask, int kind,
                          float* __restrict__ p_out, float* __restrict__ n_out,
                          float* __restrict__ dp_unit, float* __restrict__ dn_unit,
                          float2* __restrict__ block_partial, float4* __restrict__ dh,
                          const float* __restrict__ dh_scale, const uint32_t* __restrict__ live_count,
                          const uint32_t* __restrict__ live_rows, const uint32_t* __restrict__ n_live_rows) {
  rg::pdl_wait();
  constexpr int LANES = VEC < 32 ? VEC : 32;
  constexpr int G = 32 / LANES;
  constexpr int VPL = VEC / LANES;
  constexpr int UN = 8 / G;
  __shared__ float s_sc[kRowsWarps][kRowPad];
  __shared__ float2 s_part[kRowsWarps];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane / LANES, sub = lane % LANES;
  const int C = K + 1;
  constexpr bool one_step = ONE;
  const uint32_t n_live = __ldg(n_live_rows);
  const uint32_t gw = blockIdx.x * kRowsWarps + warp, nw = gridDim.x * kRowsWarps;
  float sc = dh_scale ? __ldg(dh_scale) : 1.f;
  sc *= 1.f / static_cast<float>(__ldg(live_count));         // == loss_out[1] of loss_finalize_kernel
  float* row = s_sc[warp];
  float tl = 0.f, tm = 0.f;                                   // lane 0: this warp's loss partial
  for (uint32_t t = gw; t < n_live; t += nw) {
    const int64_t i = __ldg(live_rows + t);
    const int64_t my_id = lane == 0 ? __ldg(pos + i) : (lane < C ? __ldg(neg + i * K + (lane - 1)) : 0);
    const float m = __ldg(mask + i);
    float4 hv[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) hv[v] = ld_stream_f4(h + i * VEC + v * LANES + sub);
    float4 w[UN][VPL];
    for (int c0 = 0; c0 < (ONE ? 1 : C); c0 += 8) {
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int c = c0 + u * G + g;
        const int64_t id = __shfl_sync(0xffffffffu, my_id, c < C ? c : 0);
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          w[u][v] = c < C ? __ldg(weight + id * VEC + v * LANES + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      score_item<VEC>(hv, w, c0, C, g, sub, row);
    }
    __syncwarp();
    const float sv = lane < C ? row[lane] : 0.f;
    if (lane == 0) { if (p_out) p_out[i] = sv; }
    else if (lane < C && n_out) n_out[i * K + (lane - 1)] = sv;
    __syncwarp();
    if (lane == 0) {
      const float l = position_loss(kind, row, K);
      tl += l * m;
      tm += m;
    }
    __syncwarp();
    const float dv = lane < C ? row[lane] * m : 0.f;           // unit derivative of candidate `lane`
    if (lane == 0) { if (dp_unit) dp_unit[i] = dv; }
    else if (lane < C && dn_unit) dn_unit[i * K + (lane - 1)] = dv;
    if (dh != nullptr) {
      const float my_c = dv * sc;
      float4 acc[VPL];
#pragma unroll
      for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c0 = 0; c0 < (ONE ? 1 : C); c0 += 8) {
        float cf[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int c = c0 + u * G + g;
          const int src = c < C ? c : 0;
          cf[u] = __shfl_sync(0xffffffffu, my_c, src);
          if (c >= C) cf[u] = 0.f;
          if (!one_step) {
            const int64_t id = __shfl_sync(0xffffffffu, my_id, src);
#pragma unroll
            for (int v = 0; v < VPL; ++v)
              w[u][v] = c < C ? __ldg(weight + id * VEC + v * LANES + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            acc[v].x = fmaf(cf[u], w[u][v].x, acc[v].x);
            acc[v].y = fmaf(cf[u], w[u][v].y, acc[v].y);
            acc[v].z = fmaf(cf[u], w[u][v].z, acc[v].z);
            acc[v].w = fmaf(cf[u], w[u][v].w, acc[v].w);
          }
      }
#pragma unroll
      for (int o = 16; o >= LANES; o >>= 1)
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
          acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
          acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
          acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
        }
      if (g == 0) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) st_stream_f4(dh + i * VEC + v * LANES + sub, acc[v]);
      }
    }
    __syncwarp();                                             // `row` is rewritten by the next position
  }
  if (lane == 0) s_part[warp] = make_float2(tl, tm);
  // ---- the dead positions: every output is zero (after the gathers: letting half of the warps start
  // with the zeros instead was measured, 46.1 vs 44.0 us).  A warp takes 32 consecutive positions: ONE
  // coalesced mask load + ballot, then up to 32 independent streaming stores per lane (a loop of
  // "load mask word, then store" is a chain of dependent latencies: 112 us at one CTA per SM) ----
  const int64_t n_blk = (n_pos + 31) >> 5;
  for (int64_t blk = gw; blk < n_blk; blk += nw) {
    const int64_t p0 = blk << 5;
    const bool dead = p0 + lane < n_pos && __ldg(mask + p0 + lane) == 0.f;
    const uint32_t bal = __ballot_sync(0xffffffffu, dead);
    if (bal == 0u) continue;
    if (dead) {
      if (p_out) p_out[p0 + lane] = 0.f;
      if (dp_unit) dp_unit[p0 + lane] = 0.f;
    }
    if (n_out != nullptr || dn_unit != nullptr) {
      for (int idx = lane; idx < 32 * K; idx += 32)
        if ((bal >> (idx / K)) & 1u) {
          if (n_out) n_out[p0 * K + idx] = 0.f;
          if (dn_unit) dn_unit[p0 * K + idx] = 0.f;
        }
    }
    if (dh != nullptr) {
#pragma unroll 8
      for (int idx = lane; idx < 32 * VEC; idx += 32)
        if ((bal >> (idx / VEC)) & 1u) st_stream_f4(dh + p0 * VEC + idx, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  __syncthreads();
  if (tid == 0) {
    float2 t2 = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int w2 = 0; w2 < kRowsWarps; ++w2) { t2.x += s_part[w2].x; t2.y += s_part[w2].y; }
    block_partial[blockIdx.x] = t2;
  }
  rg::pdl_trigger();
}

// Loss from already computed logits: one thread per position.
__global__ void __launch_bounds__(256)
sampled_loss_kernel(const float* __restrict__ p, const float* __restrict__ n, int64_t n_pos, int K,
                    const float* __restrict__ mask, int kind, float* __restrict__ dp_unit,
                    float* __restrict__ dn_unit, float2* __restrict__ block_partial) {
  rg::pdl_wait();
  __shared__ float2 s_part[256 / 32];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  float l = 0.f, m = 0.f;
  if (i < n_pos) {
    m = mask ? mask[i] : 1.f;
    const float pv = p[i];
    const float* nr = n + i * K;
    float dpv = 0.f;
    if (kind == RECGURU_LOSS_SCE) {
      float mx = pv;
      for (int k = 0; k < K; ++k) mx = fmaxf(mx, nr[k]);
      const float e0 = expf(pv - mx);
      float sum = e0;
      for (int k = 0; k < K; ++k) sum += expf(nr[k] - mx);
      l = (mx + logf(sum)) - pv;
      dpv = e0 / sum - 1.f;
      if (dn_unit)
        for (int k = 0; k < K; ++k) dn_unit[i * K + k] = expf(nr[k] - mx) / sum * m;
    } else {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += nr[k];
      const float nbar = s / static_cast<float>(K);
      float dnv;
      if (kind == RECGURU_LOSS_BPR) {
        const float sg = sigmoidf_ref(pv - nbar);
        l = -logf(sg);
        dpv = sg - 1.f;
        dnv = (1.f - sg) / static_cast<float>(K);
      } else {
        const float a = sigmoidf_ref(pv), b = sigmoidf_ref(nbar);
        const float ya = a + 1e-24f, yb = (1.f - b) + 1e-24f;
        l = -(logf(ya) + logf(yb));
        dpv = -(a * (1.f - a)) / ya;
        dnv = (b * (1.f - b)) / yb / static_cast<float>(K);
      }
      if (dn_unit)
        for (int k = 0; k < K; ++k) dn_unit[i * K + k] = dnv * m;
    }
    if (dp_unit) dp_unit[i] = dpv * m;
  }
  float lm = l * m;
  // fixed-order CTA reduction: lanes by shuffle tree, warps sequentially
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lm += __shfl_xor_sync(0xffffffffu, lm, o);
    m += __shfl_xor_sync(0xffffffffu, m, o);
  }
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = make_float2(lm, m);
  __syncthreads();
  if (threadIdx.x == 0) {
    float2 t = make_float2(0.f, 0.f);
    for (int w = 0; w < 256 / 32; ++w) { t.x += s_part[w].x; t.y += s_part[w].y; }
    block_partial[blockIdx.x] = t;
  }
  rg::pdl_trigger();
}

// loss_out[0] = sum(l*m)/sum(m);  loss_out[1] = 1/sum(m)  (the factor still missing from
// dp_unit/dn_unit; consumers multiply it in through their `scale2` argument).
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const float2* __restrict__ block_partial, int64_t n_blocks,
                     float* __restrict__ loss_out) {
  rg::pdl_wait();
  __shared__ float2 s[256];
  float2 t = make_float2(0.f, 0.f);
  for (int64_t b = threadIdx.x; b < n_blocks; b += 256) {
    const float2 v = block_partial[b];
    t.x += v.x; t.y += v.y;
  }
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s[threadIdx.x].x += s[threadIdx.x + o].x; s[threadIdx.x].y += s[threadIdx.x + o].y; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    loss_out[0] = s[0].x / s[0].y;
    loss_out[1] = 1.f / s[0].y;
  }
  rg::pdl_trigger();
}

// dh[i,:] = s * (dp[i] * W[pos[i],:] + sum_k dn[i,k] * W[neg[i,k],:])
template <int VEC>
__global__ void __launch_bounds__(kScoreWarps * 32)
sampled_score_bwd_dh_kernel(const float* __restrict__ dp, const float* __restrict__ dn,
                            const float* __restrict__ scale, const float* __restrict__ scale2,
                            const float4* __restrict__ weight, const int64_t* __restrict__ pos,
                            const int64_t* __restrict__ neg, int64_t n_pos, int K,
                            float4* __restrict__ dh) {
  rg::pdl_wait();
  constexpr int LANES = VEC < 32 ? VEC : 32;
  constexpr int G = 32 / LANES;
  constexpr int VPL = VEC / LANES;
  constexpr int UN = 8 / G > 0 ? 8 / G : 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / LANES, sub = lane % LANES;
  const int C = K + 1;
  float s = scale ? __ldg(scale) : 1.f;
  if (scale2) s *= __ldg(scale2);
  // ids and coefficients of all the warp's positions first: one load latency instead of four
  const int64_t base = (static_cast<int64_t>(blockIdx.x) * kScoreWarps + warp) * kPosPerWarp;
  int64_t ids_q[kPosPerWarp];
  float c_q[kPosPerWarp];
#pragma unroll
  for (int q = 0; q < kPosPerWarp; ++q) {
    const int64_t i = base + q;
    ids_q[q] = 0;
    c_q[q] = 0.f;
    if (i < n_pos) {
      if (lane == 0) { ids_q[q] = pos[i]; c_q[q] = dp[i]; }
      else if (lane <= K) { ids_q[q] = neg[i * K + (lane - 1)]; c_q[q] = dn[i * K + (lane - 1)]; }
    }
  }
#pragma unroll
  for (int q = 0; q < kPosPerWarp; ++q) {
    const int64_t i = base + q;
    if (i >= n_pos) break;
    const int64_t my_id = ids_q[q];
    const float my_c = c_q[q] * s;
    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    // masked positions have all-zero coefficients: their dh row is zero, no table row is read
    const bool any_nz = __any_sync(0xffffffffu, my_c != 0.f);
    for (int c0 = 0; any_nz && c0 < C; c0 += G * UN) {
      float4 w[UN][VPL];
      float cf[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int c = c0 + u * G + g;
        const int src = c < C ? c : 0;
        const int64_t id = __shfl_sync(0xffffffffu, my_id, src);
        cf[u] = __shfl_sync(0xffffffffu, my_c, src);
        if (c >= C) cf[u] = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          w[u][v] = c < C ? __ldg(weight + id * VEC + v * LANES + sub)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u)
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          acc[v].x = fmaf(cf[u], w[u][v].x, acc[v].x);
          acc[v].y = fmaf(cf[u], w[u][v].y, acc[v].y);
          acc[v].z = fmaf(cf[u], w[u][v].z, acc[v].z);
          acc[v].w = fmaf(cf[u], w[u][v].w, acc[v].w);
        }
    }
    // combine the G sub-warp groups (fixed order: xor tree over group index)
#pragma unroll
    for (int o = 16; o >= LANES; o >>= 1)
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
        acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
        acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
        acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
      }
    if (g == 0) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) st_stream_f4(dh + i * VEC + v * LANES + sub, acc[v]);
    }
  }
  // (no explicit launch_dependents here: code after the position loop makes ptxas interleave the
  // unrolled iterations at 74 registers instead of 40; exiting triggers the dependents just as well)
}

// E1 + R1: the reference's sampled-candidate evaluation (get_scores + double argsort), fused.
//   scores[b, 0] = u[b] . W[target[b]],  scores[b, 1+c] = u[b] . W[cand[b, c]]
//   rank[b] = #{c : scores[b, 1+c] > scores[b, 0]}        (the target wins every tie)
// One CTA per user; its 8 warps stride over the candidates in steps of 8 rows per warp with the
// same 8-rows-in-flight gather + butterfly reduction as the training kernel.  The [B, C+1, d]
// candidate-embedding tensor of the reference (2.6 GB at C=19999, B=256) never exists.
template <int VEC>
__global__ void __launch_bounds__(kScoreWarps * 32)
candidate_rank_kernel(const float4* __restrict__ u, const float4* __restrict__ weight,
                      const int64_t* __restrict__ target, const int64_t* __restrict__ cand, int C,
                      float* __restrict__ scores, int64_t* __restrict__ rank) {
  rg::pdl_wait();
  constexpr int LANES = VEC < 32 ? VEC : 32;
  constexpr int G = 32 / LANES;
  constexpr int VPL = VEC / LANES;
  constexpr int UN = 8 / G;
  __shared__ int s_cnt[kScoreWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / LANES, sub = lane % LANES;
  const int64_t b = blockIdx.x;
  float4 hv[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) hv[v] = __ldg(u + b * VEC + v * LANES + sub);
  // every warp computes the target score itself (same instruction sequence => identical bits)
  float s0;
  {
    const int64_t tid = target[b];
    float a = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const float4 w = __ldg(weight + tid * VEC + v * LANES + sub);
      a = fmaf(hv[v].x, w.x, a); a = fmaf(hv[v].y, w.y, a);
      a = fmaf(hv[v].z, w.z, a); a = fmaf(hv[v].w, w.w, a);
    }
    float acc[UN];
#pragma unroll
    for (int q = 0; q < UN; ++q) acc[q] = a;           // same reduction tree as the candidates
    const float r = group_multi_reduce<LANES, UN>(acc, sub);
    s0 = __shfl_sync(0xffffffffu, r, 0);
    if (scores && warp == 0 && lane == 0) scores[b * (C + 1)] = s0;
  }
  int cnt = 0;
  constexpr int kRep = LANES / UN;
  for (int c0 = warp * 8; c0 < C; c0 += kScoreWarps * 8) {
    float4 w[UN][VPL];
    int64_t my_id = 0;
    if (lane < 8 && c0 + lane < C) my_id = cand[b * C + c0 + lane];
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      const int c = q * G + g;                             // candidate slot within this step (0..7)
      const int64_t id = __shfl_sync(0xffffffffu, my_id, c);
#pragma unroll
      for (int v = 0; v < VPL; ++v)
        w[q][v] = c0 + c < C ? __ldg(weight + id * VEC + v * LANES + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float acc[UN];
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      float a = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        a = fmaf(hv[v].x, w[q][v].x, a); a = fmaf(hv[v].y, w[q][v].y, a);
        a = fmaf(hv[v].z, w[q][v].z, a); a = fmaf(hv[v].w, w[q][v].w, a);
      }
      acc[q] = a;
    }
    const float r = group_multi_reduce<LANES, UN>(acc, sub);
    const int c = c0 + (sub / kRep) * G + g;
    if (sub % kRep == 0 && c < C) {
      if (scores) scores[b * (C + 1) + 1 + c] = r;
      cnt += r > s0 ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w2 = 0; w2 < kScoreWarps; ++w2) t += s_cnt[w2];
    rank[b] = t;
  }
  rg::pdl_trigger();
}

static int check_common(const char* who, int64_t V1, int64_t d, int64_t n_pos, int64_t K) {
  RG_REQUIRE(V1 > 0 && V1 < (1ll << 31) && n_pos >= 0, RECGURU_E_SHAPE, "%s: bad sizes", who);
  RG_REQUIRE(d == 32 || d == 64 || d == 128 || d == 256, RECGURU_E_UNSUPPORTED,
             "%s: d=%lld not in {32,64,128,256}", who, (long long)d);
  RG_REQUIRE(K >= 1 && K <= 31, RECGURU_E_UNSUPPORTED, "%s: K=%lld not in [1,31]", who,
             (long long)K);
  return RECGURU_OK;
}

}  // namespace rg

extern "C" int64_t recguru_sampled_workspace_bytes(int64_t n_pos, int64_t K) {
  (void)K;
  if (n_pos < 0) return RECGURU_E_SHAPE;
  // block partials of whichever kernel uses more CTAs (kPosPerBlock = 32 < 256)
  return rg::align_up((rg::ceil_div(n_pos > 0 ? n_pos : 1, rg::kPosPerBlock) + 1) * 8, 256);
}

static int sampled_score_loss_fwd_impl(const float* h, const float* weight, int64_t V1,
                                       int64_t d, const int64_t* pos, const int64_t* neg,
                                       int64_t n_pos, int64_t K, const float* mask, int kind,
                                       float* p_out, float* n_out, float* loss_out,
                                       float* dp_unit, float* dn_unit, void* workspace,
                                       int64_t workspace_bytes, float* dh, const float* dh_scale,
                                       const uint32_t* live_count, const uint32_t* live_rows,
                                       const uint32_t* n_live_rows, recguru_stream_t stream) {
  using namespace rg;
  const bool skip_masked = kind >= 0 && (kind & RECGURU_LOSS_SKIP_MASKED) && mask != nullptr;
  if (dh != nullptr) {
    RG_REQUIRE(skip_masked, RECGURU_E_UNSUPPORTED,
               "sampled_score_loss_fwd_dh: the fused dh needs a loss kind with RECGURU_LOSS_SKIP_MASKED and a mask");
    RG_REQUIRE(live_count != nullptr && aligned16(dh), RECGURU_E_SHAPE,
               "sampled_score_loss_fwd_dh: live_count is null or dh is not 16-byte aligned");
  }
  if (live_rows != nullptr)
    RG_REQUIRE(skip_masked && n_live_rows != nullptr && live_count != nullptr && (dh == nullptr || aligned16(dh)),
               RECGURU_E_SHAPE, "sampled_score_loss_fwd_dh_rows: needs a SKIP_MASKED kind, a mask, live_count and n_live_rows");
  if (kind >= 0) kind &= ~RECGURU_LOSS_SKIP_MASKED;
  int rc = check_common("sampled_score_loss_fwd", V1, d, n_pos, K);
  if (rc) return rc;
  RG_REQUIRE(h && weight && pos && neg, RECGURU_E_SHAPE, "sampled_score_loss_fwd: null pointer");
  RG_REQUIRE(aligned16(h) && aligned16(weight), RECGURU_E_SHAPE,
             "sampled_score_loss_fwd: h/weight must be 16-byte aligned");
  RG_REQUIRE(kind <= RECGURU_LOSS_SCE, RECGURU_E_UNSUPPORTED, "sampled_score_loss_fwd: kind=%d", kind);
  RG_REQUIRE(kind < 0 || loss_out, RECGURU_E_SHAPE, "sampled_score_loss_fwd: loss_out is null");
  if (n_pos == 0) return RECGURU_OK;
  RG_CHECK_IDS("sampled_score_loss_fwd (pos)", pos, n_pos, 0, V1, static_cast<cudaStream_t>(stream));
  RG_CHECK_IDS("sampled_score_loss_fwd (neg)", neg, n_pos * K, 0, V1, static_cast<cudaStream_t>(stream));
  // dense: one CTA per 128 positions; padding-aware: one persistent CTA per SM
  const int RS = static_cast<int>(K + 1) | 1;
  int chunk = RS <= 13 ? 1024 : 512;                         // 2 * chunk * RS * 4 B of shared memory
  {
    // RECGURU_SPARSE_CHUNK=<n>: smaller chunks for the padding-aware kernel (a multiple of 32, at
    // most the default), e.g. 704 = two equal chunks per CTA at cfg2.  Measurement knob, default off.
    static int env_chunk = -1;
    if (env_chunk < 0) {
      const char* e = getenv("RECGURU_SPARSE_CHUNK");
      env_chunk = e ? atoi(e) : 0;
    }
    if (env_chunk >= 32 && env_chunk <= chunk && env_chunk % 32 == 0) chunk = env_chunk;
  }
  int64_t blocks = ceil_div(n_pos, kScoreWarps * kDensePW);
  int64_t pos_per_cta = 0;
  int nt = kSparseThreadsDefault;
  if (skip_masked) {
    // RECGURU_SCORE_RESERVE_SMS=<r>: leave r SMs to the kernels of the other branches of a step (the
    // table-backward plan runs on a side stream while this persistent kernel owns every SM it gets)
    static int env_reserve = -1;
    if (env_reserve < 0) {
      const char* e = getenv("RECGURU_SCORE_RESERVE_SMS");
      env_reserve = e ? atoi(e) : 0;
      if (env_reserve < 0 || env_reserve > 64) env_reserve = 0;
    }
    static int env_nt = -1;       // RECGURU_SPARSE_THREADS = 1024 | 512 | 256 threads per CTA
    if (env_nt < 0) {
      const char* e = getenv("RECGURU_SPARSE_THREADS");
      env_nt = e ? atoi(e) : kSparseThreadsDefault;
      if (env_nt != 1024 && env_nt != 512 && env_nt != 256) env_nt = kSparseThreadsDefault;
    }
    nt = env_nt;
    if (chunk > nt) chunk = nt;                                // phase 0 maps one position to one thread
    blocks = rg::min64(static_cast<int64_t>(num_sms() - env_reserve) * (1024 / nt), ceil_div(n_pos, chunk));
    pos_per_cta = align_up(ceil_div(n_pos, blocks), 32);
    blocks = ceil_div(n_pos, pos_per_cta);
    // live-row list given: 4 CTAs of 8 warps per SM, one warp per live position
    if (live_rows != nullptr) {
      // RECGURU_ROWS_CTAS=<c>: CTAs per SM (1..8; default 4 = every register of the SM at d <= 128)
      static int env_ctas = -1;
      if (env_ctas < 0) {
        const char* e = getenv("RECGURU_ROWS_CTAS");
        env_ctas = e ? atoi(e) : 4;
        if (env_ctas < 1 || env_ctas > 8) env_ctas = 4;
      }
      blocks = rg::min64(static_cast<int64_t>(num_sms()) * env_ctas, ceil_div(n_pos, 32));
    }
  }
  if (kind >= 0)
    RG_REQUIRE(workspace && workspace_bytes >= blocks * 8, RECGURU_E_WORKSPACE,
               "sampled_score_loss_fwd: workspace %lld < %lld", (long long)workspace_bytes,
               (long long)(blocks * 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float2* part = static_cast<float2*>(workspace);
  const float4* h4 = reinterpret_cast<const float4*>(h);
  const float4* w4 = reinterpret_cast<const float4*>(weight);
  const size_t sparse_smem = static_cast<size_t>(chunk) * (2 * RS + 2) * 4;
#define RG_LAUNCH_SPARSE(V, T)                                                                    \
  do {                                                                                           \
    static rg::PerDeviceOnce once;           /* the attribute is per device */                  \
    const int dev = rg::current_device();                                                        \
    if (once.need(dev)) {                                                                        \
      const cudaError_t ea = cudaFuncSetAttribute(sampled_score_fwd_sparse_kernel<V, T>,         \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);              \
      RG_REQUIRE(ea == cudaSuccess, RECGURU_E_CUDA, "sampled_score_fwd: cudaFuncSetAttribute: %s", \
                 cudaGetErrorString(ea));                                                        \
      once.done(dev);                                                                            \
    }                                                                                            \
    rg::launch_pdl((sampled_score_fwd_sparse_kernel<V, T>), static_cast<unsigned>(blocks), T, sparse_smem, st,  \
        h4, w4, pos, neg, n_pos, static_cast<int>(K), mask, kind, p_out, n_out, dp_unit, dn_unit, part,     \
        chunk, pos_per_cta, reinterpret_cast<float4*>(dh), dh_scale, live_count);                \
  } while (0)
  // 512 / 256 threads per CTA (2 / 4 CTAs per SM) measured slower (59 / 69 us vs 56 us at cfg2):
  // compiled only with RECGURU_DEBUG_KNOBS
#ifdef RECGURU_DEBUG_KNOBS
#define RG_SPARSE_VARIANTS(V) if (nt == 256) RG_LAUNCH_SPARSE(V, 256); else if (nt == 512) RG_LAUNCH_SPARSE(V, 512); else
#else
#define RG_SPARSE_VARIANTS(V)
#endif
#define RG_LAUNCH_SCORE(V)                                                                       \
  do {                                                                                           \
    if (live_rows != nullptr) {                                                                  \
      if (K + 1 <= 8)                                                                            \
        rg::launch_pdl((sampled_score_rows_kernel<V, true>), static_cast<unsigned>(blocks), kRowsWarps * 32, 0, st, \
            h4, w4, pos, neg, n_pos, static_cast<int>(K), mask, kind, p_out, n_out, dp_unit, dn_unit, part,   \
            reinterpret_cast<float4*>(dh), dh_scale, live_count, live_rows, n_live_rows);          \
      else                                                                                       \
        rg::launch_pdl((sampled_score_rows_kernel<V, false>), static_cast<unsigned>(blocks), kRowsWarps * 32, 0, st, \
            h4, w4, pos, neg, n_pos, static_cast<int>(K), mask, kind, p_out, n_out, dp_unit, dn_unit, part,   \
            reinterpret_cast<float4*>(dh), dh_scale, live_count, live_rows, n_live_rows);          \
    } else if (skip_masked) {                                                                    \
      RG_SPARSE_VARIANTS(V)                                                                      \
      RG_LAUNCH_SPARSE(V, 1024);                                                                 \
    } else {                                                                                     \
      rg::launch_pdl((sampled_score_fwd_kernel<V>), static_cast<unsigned>(blocks), kScoreWarps * 32, 0, st,    \
          h4, w4, pos, neg, n_pos, static_cast<int>(K), mask, kind, p_out, n_out, dp_unit, dn_unit, part);    \
    }                                                                                            \
  } while (0)
  switch (d / 4) {
    case 8: RG_LAUNCH_SCORE(8); break;
    case 16: RG_LAUNCH_SCORE(16); break;
    case 32: RG_LAUNCH_SCORE(32); break;
    default: RG_LAUNCH_SCORE(64); break;
  }
#undef RG_LAUNCH_SCORE
#undef RG_LAUNCH_SPARSE
#undef RG_SPARSE_VARIANTS
  RG_CHECK_LAUNCH("sampled_score_fwd");
  if (kind >= 0) {
    rg::launch_pdl((loss_finalize_kernel), 1, 256, 0, st, part, blocks, loss_out);
    RG_CHECK_LAUNCH("loss_finalize");
  }
  return RECGURU_OK;
}

extern "C" int recguru_sampled_score_loss_fwd(const float* h, const float* weight, int64_t V1,
                                              int64_t d, const int64_t* pos, const int64_t* neg,
                                              int64_t n_pos, int64_t K, const float* mask, int kind,
                                              float* p_out, float* n_out, float* loss_out,
                                              float* dp_unit, float* dn_unit, void* workspace,
                                              int64_t workspace_bytes, recguru_stream_t stream) {
  return sampled_score_loss_fwd_impl(h, weight, V1, d, pos, neg, n_pos, K, mask, kind, p_out, n_out, loss_out,
                                     dp_unit, dn_unit, workspace, workspace_bytes, nullptr, nullptr, nullptr,
                                     nullptr, nullptr, stream);
}

extern "C" int recguru_sampled_score_loss_fwd_dh(const float* h, const float* weight, int64_t V1,
                                                 int64_t d, const int64_t* pos, const int64_t* neg,
                                                 int64_t n_pos, int64_t K, const float* mask, int kind,
                                                 float* p_out, float* n_out, float* loss_out,
                                                 float* dp_unit, float* dn_unit, void* workspace,
                                                 int64_t workspace_bytes, const float* grad_scale,
                                                 const uint32_t* live_count, float* dh,
                                                 recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(dh != nullptr, RECGURU_E_SHAPE, "sampled_score_loss_fwd_dh: dh is null");
  return sampled_score_loss_fwd_impl(h, weight, V1, d, pos, neg, n_pos, K, mask, kind, p_out, n_out, loss_out,
                                     dp_unit, dn_unit, workspace, workspace_bytes, dh, grad_scale, live_count,
                                     nullptr, nullptr, stream);
}

// The same, given the ascending list of live positions (mask[i] != 0 <=> i in live_rows[0 .. *n_live_rows)):
// one warp per live position, no compaction inside.  dh may be NULL (forward only).
extern "C" int recguru_sampled_score_loss_fwd_dh_rows(const float* h, const float* weight, int64_t V1,
                                                      int64_t d, const int64_t* pos, const int64_t* neg,
                                                      int64_t n_pos, int64_t K, const float* mask, int kind,
                                                      float* p_out, float* n_out, float* loss_out,
                                                      float* dp_unit, float* dn_unit, void* workspace,
                                                      int64_t workspace_bytes, const float* grad_scale,
                                                      const uint32_t* live_count, const uint32_t* live_rows,
                                                      const uint32_t* n_live_rows, float* dh,
                                                      recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(live_rows != nullptr && n_live_rows != nullptr, RECGURU_E_SHAPE,
             "sampled_score_loss_fwd_dh_rows: live_rows is null");
  return sampled_score_loss_fwd_impl(h, weight, V1, d, pos, neg, n_pos, K, mask, kind, p_out, n_out, loss_out,
                                     dp_unit, dn_unit, workspace, workspace_bytes, dh, grad_scale, live_count,
                                     live_rows, n_live_rows, stream);
}

extern "C" int recguru_sampled_loss_fwd(const float* p, const float* n, int64_t n_pos, int64_t K,
                                        const float* mask, int kind, float* loss_out,
                                        float* dp_unit, float* dn_unit, void* workspace,
                                        int64_t workspace_bytes, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(p && n && loss_out, RECGURU_E_SHAPE, "sampled_loss_fwd: null pointer");
  RG_REQUIRE(n_pos > 0, RECGURU_E_SHAPE, "sampled_loss_fwd: n_pos must be > 0");
  RG_REQUIRE(K >= 1 && K <= 4096, RECGURU_E_UNSUPPORTED, "sampled_loss_fwd: K=%lld", (long long)K);
  RG_REQUIRE(kind >= 0 && kind <= RECGURU_LOSS_SCE, RECGURU_E_UNSUPPORTED,
             "sampled_loss_fwd: kind=%d", kind);
  const int64_t blocks = ceil_div(n_pos, 256);
  RG_REQUIRE(workspace && workspace_bytes >= blocks * 8, RECGURU_E_WORKSPACE,
             "sampled_loss_fwd: workspace %lld < %lld", (long long)workspace_bytes,
             (long long)(blocks * 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float2* part = static_cast<float2*>(workspace);
  rg::launch_pdl((sampled_loss_kernel), static_cast<unsigned>(blocks), 256, 0, st, 
      p, n, n_pos, static_cast<int>(K), mask, kind, dp_unit, dn_unit, part);
  RG_CHECK_LAUNCH("sampled_loss");
  rg::launch_pdl((loss_finalize_kernel), 1, 256, 0, st, part, blocks, loss_out);
  RG_CHECK_LAUNCH("loss_finalize");
  return RECGURU_OK;
}

extern "C" int recguru_sampled_score_bwd_dh(const float* dp, const float* dn, const float* scale,
                                            const float* scale2, const float* weight, int64_t V1,
                                            int64_t d, const int64_t* pos, const int64_t* neg,
                                            int64_t n_pos, int64_t K, float* dh,
                                            recguru_stream_t stream) {
  using namespace rg;
  int rc = check_common("sampled_score_bwd_dh", V1, d, n_pos, K);
  if (rc) return rc;
  RG_REQUIRE(dp && dn && weight && pos && neg && dh, RECGURU_E_SHAPE,
             "sampled_score_bwd_dh: null pointer");
  RG_REQUIRE(aligned16(weight) && aligned16(dh), RECGURU_E_SHAPE,
             "sampled_score_bwd_dh: weight/dh must be 16-byte aligned");
  if (n_pos == 0) return RECGURU_OK;
  RG_CHECK_IDS("sampled_score_bwd_dh (pos)", pos, n_pos, 0, V1, static_cast<cudaStream_t>(stream));
  RG_CHECK_IDS("sampled_score_bwd_dh (neg)", neg, n_pos * K, 0, V1, static_cast<cudaStream_t>(stream));
  const int64_t blocks = ceil_div(n_pos, kPosPerBlock);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float4* w4 = reinterpret_cast<const float4*>(weight);
  float4* dh4 = reinterpret_cast<float4*>(dh);
#define RG_LAUNCH_DH(V)                                                                        \
  rg::launch_pdl((sampled_score_bwd_dh_kernel<V>), static_cast<unsigned>(blocks), kScoreWarps * 32, 0, st,   \
      dp, dn, scale, scale2, w4, pos, neg, n_pos, static_cast<int>(K), dh4)
  switch (d / 4) {
    case 8: RG_LAUNCH_DH(8); break;
    case 16: RG_LAUNCH_DH(16); break;
    case 32: RG_LAUNCH_DH(32); break;
    default: RG_LAUNCH_DH(64); break;
  }
#undef RG_LAUNCH_DH
  RG_CHECK_LAUNCH("sampled_score_bwd_dh");
  return RECGURU_OK;
}

extern "C" int recguru_candidate_scores_rank(const float* u, const float* weight, int64_t V1, int64_t d,
                                             const int64_t* target, const int64_t* cand, int64_t B,
                                             int64_t C, float* scores, int64_t* rank,
                                             recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(u && weight && target && cand && rank, RECGURU_E_SHAPE, "candidate_scores_rank: null pointer");
  RG_REQUIRE(V1 > 0 && V1 < (1ll << 31) && B >= 0 && C >= 1 && C < (1ll << 30), RECGURU_E_SHAPE,
             "candidate_scores_rank: bad sizes");
  RG_REQUIRE(d == 32 || d == 64 || d == 128 || d == 256, RECGURU_E_UNSUPPORTED,
             "candidate_scores_rank: d=%lld not in {32,64,128,256}", (long long)d);
  RG_REQUIRE(aligned16(u) && aligned16(weight), RECGURU_E_SHAPE,
             "candidate_scores_rank: u/weight must be 16-byte aligned");
  if (B == 0) return RECGURU_OK;
  RG_CHECK_IDS("candidate_scores_rank (target)", target, B, 0, V1, static_cast<cudaStream_t>(stream));
  RG_CHECK_IDS("candidate_scores_rank (candidates)", cand, B * C, 0, V1, static_cast<cudaStream_t>(stream));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float4* u4 = reinterpret_cast<const float4*>(u);
  const float4* w4 = reinterpret_cast<const float4*>(weight);
#define RG_LAUNCH_CAND(V)                                                                         \
  rg::launch_pdl((candidate_rank_kernel<V>), static_cast<unsigned>(B), kScoreWarps * 32, 0, st,                 \
      u4, w4, target, cand, static_cast<int>(C), scores, rank)
  switch (d / 4) {
    case 8: RG_LAUNCH_CAND(8); break;
    case 16: RG_LAUNCH_CAND(16); break;
    case 32: RG_LAUNCH_CAND(32); break;
    default: RG_LAUNCH_CAND(64); break;
  }
#undef RG_LAUNCH_CAND
  RG_CHECK_LAUNCH("candidate_rank");
  return RECGURU_OK;
}
