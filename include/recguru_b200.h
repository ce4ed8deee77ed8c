/*
 * recguru_b200.h -- C ABI of the B200-native RecGURU embedding-and-scoring hot path.
 *
 * This is the drop-in boundary.  The reference (Chain123/RecGURU)
 * is pure PyTorch and has no FFI of its own; each entry point below names the reference
 * call site (file:line under /root/reference/GURU) whose ATen op sequence it replaces.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types.  Every pointer is a DEVICE pointer unless
 *     the parameter name ends in _host.
 *   - the caller owns every buffer (inputs, outputs, scratch).  No allocation, no global
 *     state, no implicit synchronisation inside; work is enqueued on `stream`.
 *   - scratch: call the matching *_workspace_bytes() first and pass a buffer of at least
 *     that size, 256-byte aligned.
 *   - return value: 0 on success, negative on error
 *       RECGURU_E_SHAPE      (-1)  bad shape / alignment / null pointer
 *       RECGURU_E_UNSUPPORTED(-2)  unsupported d / K / size
 *       RECGURU_E_CUDA       (-3)  CUDA launch error (see recguru_last_error())
 *       RECGURU_E_WORKSPACE  (-4)  workspace too small
 *       RECGURU_E_INDEX      (-5)  an item id outside [0, V1) -- only reported when the
 *                                  environment has RECGURU_CHECK_IDS=1 (debug aid: the check
 *                                  synchronises the stream; nn.Embedding raises IndexError /
 *                                  a device assert for the same input)
 *   - thread / device model: entry points may be called concurrently from several host
 *     threads and on several devices of one process (nn.DataParallel, train_auto.py:350-351);
 *     work goes to the CURRENT device (cudaGetDevice) on `stream`.  Host-side caches (SM
 *     count, kernel attributes) are kept per device; the only process-wide state is the
 *     launch counter and the thread-local error string.
 *   - row-major fp32 tables; ids are int64 on arrival (numpy int64 -> default_collate,
 *     data/data_loader.py:36,418-421) and are narrowed to int32 internally (V1 < 2^31).
 */
#ifndef RECGURU_B200_H_
#define RECGURU_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RECGURU_OK 0
#define RECGURU_E_SHAPE (-1)
#define RECGURU_E_UNSUPPORTED (-2)
#define RECGURU_E_CUDA (-3)
#define RECGURU_E_WORKSPACE (-4)
#define RECGURU_E_INDEX (-5)

/* loss kinds for the sampled-scoring ops */
#define RECGURU_LOSS_BPR 0     /* tools/lossfunctions.py:52-72  BPRLoss               */
#define RECGURU_LOSS_BPR_SAS 1 /* tools/lossfunctions.py:75-96  BPRLoss_sas           */
#define RECGURU_LOSS_SCE 2     /* tools/lossfunctions.py:29-49  SampledCrossEntropyLoss (label 0) */
/* OR-ed into `kind` of recguru_sampled_score_loss_fwd: positions with mask[i] == 0 are not scored.
 * The reference scores them and then multiplies their loss term by 0 (lossfunctions.py:44-48,
 * :66-71,:90-95), so loss, dp_unit and dn_unit are unchanged; only p[i], n[i,:] of those
 * positions differ (reported as 0).  On a padded batch this removes most of the row gathers. */
#define RECGURU_LOSS_SKIP_MASKED 0x100

typedef void* recguru_stream_t; /* cudaStream_t */

/* ABI version of this header (bumped on any signature change). */
int recguru_abi_version(void);
/* Thread-local description of the last error returned on this thread ("" if none). */
const char* recguru_last_error(void);
/* Number of kernels this library has launched since load (all threads). */
int64_t recguru_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Piece 1 -- fused item-embedding gather + positional add + pad mask      (K1: G1,M1,P1)
 *   replaces  src_emb(ids) ; mask = 1-(ids==pad) ; pos_emb(x, mask)  before dropout
 *             AutoEnc4Rec.py:71-75,:89-93,:176-181,:189-193
 *             AutoEnc4Rec_cross.py:98-102,:124-128,:156-160 ; Transformer/transformer.py:104-106
 *   out[t,:]  = (weight[ids[t],:] + pe[t % L,:]) * mask[t]      (fp32, add then mul: bit-exact)
 *   mask[t]   = ids[t] != pad_index ? 1 : 0   when mask_in == NULL, else mask_in[t]
 *   pe == NULL  -> plain gather  out[t,:] = weight[ids[t],:]   (S1/S2/E1 gathers)
 *   mask_out may be NULL.  d % 4 == 0, weight/pe/out 16-byte aligned.
 * ---------------------------------------------------------------------------------- */
int recguru_gather_pe_mask(const float* weight, int64_t V1, int64_t d,
                           const int64_t* ids, int64_t n_tokens, int64_t L,
                           const float* pe, const float* mask_in, int64_t pad_index,
                           float* out, float* mask_out, recguru_stream_t stream);

/* PositionalEncoding.forward(x, mask) before dropout (Transformer/transformer.py:104-106;
 * learned variant :337-342 by passing the learned table as pe):
 *   out[t,:] = (x[t,:] + pe[t % L,:]) * mask[t]                                        */
/* M1 alone: mask[i] = ids[i] != pad_index ? 1 : 0   (AutoEnc4Rec.py:71,:89; lets the plan of the
 * table backward start before the gather)                                                  */
int recguru_pad_mask(const int64_t* ids, int64_t n, int64_t pad_index, float* mask_out,
                     recguru_stream_t stream);
/* M1 + its population count (the sum(mask) of the masked-mean losses, known before any activation
 * exists): count3 = three uint32 words zeroed once by the caller; after the call count3[2] holds
 * #{i : ids[i] != pad_index}; words 0-1 are scratch that the kernel leaves cleared.             */
int recguru_pad_mask_count(const int64_t* ids, int64_t n, int64_t pad_index, float* mask_out,
                           uint32_t* count3, recguru_stream_t stream);
int recguru_pe_mask(const float* x, const float* pe, const float* mask, int64_t n_tokens,
                    int64_t L, int64_t d, float* out, recguru_stream_t stream);

/* SURVEY 8f rank 4 -- the nn.Dropout that closes PositionalEncoding.forward
 * (Transformer/transformer.py:104-106, p = 0.5: config_auto4rec.py:225), fused into K1 and
 * BIT-COMPATIBLE with torch's own CUDA dropout kernel and generator stream:
 *   out = dropout((weight[ids] + pe) * mask, p)  ==  torch.nn.functional.dropout(K1_out, p, True)
 * when torch's CUDA generator holds (seed, philox_offset) and its launch has torch_threads =
 * 256 * min(SMs * (maxThreadsPerSM / 256), ceil(numel / 256)) threads; the caller then advances the
 * generator by ((numel - 1) / (torch_threads * 4) + 1) * 4, as torch does (recguru_b200/ops.py:
 * torch_dropout_state).  recguru_dropout is the same dropout on an existing tensor (rows whose
 * row_mask entry is 0 pass through untouched; row_mask may be NULL) -- also its own backward:
 * grad_in = recguru_dropout(grad_out) with the state of the forward call. */
int recguru_gather_pe_mask_dropout(const float* weight, int64_t V1, int64_t d, const int64_t* ids,
                                   int64_t n_tokens, int64_t L, const float* pe, const float* mask_in,
                                   int64_t pad_index, double p_drop, uint64_t seed, uint64_t philox_offset,
                                   int64_t torch_threads, float* out, float* mask_out,
                                   recguru_stream_t stream);
int recguru_dropout(const float* x, const float* row_mask, int64_t n_rows, int64_t d, double p_drop,
                    uint64_t seed, uint64_t philox_offset, int64_t torch_threads, float* out,
                    recguru_stream_t stream);
/* P1' -- gradient of a LEARNED position table (PositionalEncodingM, Transformer/transformer.py:327-342:
 * x + pos_emb(arange(L)), then the mask):  dP[l,:] (+)= sum_b mask[b,l] * g[b,l,:], fixed summation
 * order (bit-reproducible).  g [B,L,d], mask [B,L], dP [L,d]. */
int recguru_pe_grad(const float* g, const float* mask, int64_t B, int64_t L, int64_t d, int accumulate,
                    float* dP, recguru_stream_t stream);
/* SURVEY 8f rank 4 -- user representation: u[b,:] = h[b,L-1,:]  (dec_embedding[:, -1, :],
 * train_auto.py:178,:194; get_user_embed, gan_training.py:152-162), as fp32 (discriminator /
 * get_scores input) and/or bf16 (operand of the catalog sweep) in one pass; either may be NULL. */
int recguru_extract_last(const float* h, int64_t B, int64_t L, int64_t d, float* out_f32, void* out_bf16,
                         recguru_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Piece 2 -- deterministic embedding-table backward                               (K2: B1)
 *   replaces  ATen embedding_dense_backward at every lookup site + the dense add_
 *             accumulation (loss.backward(): train_auto.py:124, gan_training.py:491-564)
 *   A lookup *site* s contributes, for each of its `count` gathered indices j,
 *       dW[ids[j],:] += coef[j] * (*scale) * (*scale2) * X[j / group, :]
 *   (coef NULL -> 1, scale/scale2 NULL -> 1).  Input sites use X = upstream grad, group = 1,
 *   coef = pad mask; scoring sites use X = h, group = K, coef = dLoss/dlogit.
 *   Indices equal to padding_idx are skipped (nn.Embedding padding_idx; pass -1 for none).
 *   Algorithm: stable LSD radix sort of (id, occurrence) over all sites in one pass,
 *   run-length segments, fixed-order segmented sum (no atomics) -> bit-reproducible.
 *   dW is dense [V1,d]; zero_fill != 0 zero-fills it first (dense-grad contract of the
 *   reference's Adam), zero_fill == 0 accumulates (+=) into the touched rows only.
 * ---------------------------------------------------------------------------------- */
#define RECGURU_MAX_SITES 8
typedef struct {
  const int64_t* ids; /* [count] */
  int64_t count;
  const float* X;     /* [count / group, d] */
  const float* coef;  /* [count] or NULL */
  const float* scale; /* device scalar or NULL (e.g. upstream grad of the scalar loss) */
  const float* scale2;/* second device scalar or NULL (e.g. 1/sum(mask), loss_out[1]) */
  int64_t group;      /* gathered indices per row of X (1, or K) */
  const float* live;  /* [count / group] or NULL: liveness mask per row of X (e.g. the pad mask).
                         An occurrence carries gradient iff id != padding_idx and
                         (live ? live[j/group] != 0 : coef == NULL || coef[j] != 0).           */
} recguru_lookup_site_t;

/* `zero_fill` modes of dW */
#define RECGURU_DW_ACCUMULATE 0 /* dW += gradient (rows read-modify-written)                    */
#define RECGURU_DW_ZERO_FILL 1  /* dW = gradient, untouched rows zeroed here (dense-grad contract
                                   of the reference's Adam)                                     */
#define RECGURU_DW_PREZEROED 2  /* dW = gradient, the caller zeroed dW (recguru_fill_zero), e.g.
                                   on another stream                                            */

int64_t recguru_embedding_bwd_workspace_bytes(int64_t total_count, int64_t V1, int64_t d);
int recguru_embedding_bwd(const recguru_lookup_site_t* sites_host, int n_sites, int64_t d,
                          int64_t V1, int64_t padding_idx, float* dW, int zero_fill,
                          void* workspace, int64_t workspace_bytes, recguru_stream_t stream);
/* The same computation in two stages, so that a training step can overlap the first with its
 * forward pass (it depends on the ids and the pad mask only, not on any activation):
 *   plan  : stable compaction of the live (id, occurrence) pairs + LSD radix sort -> workspace.
 *           Reads site.ids and site.live (else site.coef); X/scale may still be unset.
 *   apply : segmented reduce of coef * X rows over the plan, for the SAME ids/counts/groups and the
 *           same `live` configuration (all sites sharing one mask selects compaction by row in both).
 * recguru_embedding_bwd == plan followed by apply on one stream.                              */
int recguru_embedding_bwd_plan(const recguru_lookup_site_t* sites_host, int n_sites, int64_t V1,
                               int64_t padding_idx, void* workspace, int64_t workspace_bytes,
                               recguru_stream_t stream);
/* recguru_embedding_bwd_plan with the shared pad mask of a training step derived inside (row mode:
 * every site's `live` must be mask_out, one entry per row of X): mask_out[i] = mask_ids[i] != pad_index,
 * and -- if count3 != NULL -- its population count as recguru_pad_mask_count publishes it.  Saves the
 * separate pad-mask launch at the head of the plan branch (AutoEnc4Rec.py:71,:89 + train_auto.py:124). */
int recguru_embedding_bwd_plan_padmask(const recguru_lookup_site_t* sites, int n_sites, int64_t V1,
                                       int64_t padding_idx, const int64_t* mask_ids, int64_t pad_index,
                                       float* mask_out, uint32_t* count3, void* workspace,
                                       int64_t workspace_bytes, recguru_stream_t stream);
/* The plan in two calls, so that a caller can record an event between the compaction and the sort:
 * RECGURU_PLAN_COMPACT (stable compaction; in row mode it also leaves the ascending list of live rows
 * that recguru_embedding_bwd_plan_live_rows points at) then RECGURU_PLAN_SORT (the radix sort) on the
 * same stream == RECGURU_PLAN_ALL == recguru_embedding_bwd_plan. */
#define RECGURU_PLAN_ALL 0
#define RECGURU_PLAN_COMPACT 1
#define RECGURU_PLAN_SORT 2
int recguru_embedding_bwd_plan_stage(const recguru_lookup_site_t* sites, int n_sites, int64_t V1,
                                     int64_t padding_idx, void* workspace, int64_t workspace_bytes,
                                     int stage, recguru_stream_t stream);
/* Host-only: device pointers (into `workspace`) of the live-row list the row-mode compaction of these
 * sites writes -- *live_rows[0 .. **n_live_rows) ascending rows with live != 0.  It is the list of
 * positions the loss keeps (train_auto.py:124 tgt_padding_mask), consumed by
 * recguru_sampled_score_loss_fwd_dh_rows.  RECGURU_E_UNSUPPORTED if the sites are not in row mode. */
int recguru_embedding_bwd_plan_live_rows(const recguru_lookup_site_t* sites, int n_sites, int64_t V1,
                                         void* workspace, int64_t workspace_bytes,
                                         const uint32_t** live_rows, const uint32_t** n_live_rows);
/* The same list without a plan, for steps whose lookup sites do not share one mask (RecGURU: the inputs
 * and the loss are masked by different id tensors, AutoEnc4Rec_cross.py:98-128): live_rows[0 .. *n_live_rows)
 * = ascending i with mask[i] != 0.  Three launches; workspace = recguru_live_rows_workspace_bytes(n_rows). */
int64_t recguru_live_rows_workspace_bytes(int64_t n_rows);
int recguru_live_rows(const float* mask, int64_t n_rows, uint32_t* live_rows, uint32_t* n_live_rows,
                      void* workspace, int64_t workspace_bytes, recguru_stream_t stream);
int recguru_embedding_bwd_apply(const recguru_lookup_site_t* sites_host, int n_sites, int64_t d,
                                int64_t V1, float* dW, int zero_fill, void* workspace,
                                int64_t workspace_bytes, recguru_stream_t stream);
/* p[0 .. n_floats) = 0 (n_floats a multiple of 4, p 16-byte aligned). */
int recguru_fill_zero(float* p, int64_t n_floats, recguru_stream_t stream);

/* The sort stage on its own (parity: bit-exact vs torch.sort(ids, stable=True)).
 *   sorted_ids[i] = ids[perm[i]], ascending, ties in ascending perm.  ids in [0, V1).   */
int64_t recguru_sort_ids_workspace_bytes(int64_t count, int64_t V1);
int recguru_sort_ids(const int64_t* ids, int64_t count, int64_t V1, int64_t* sorted_ids,
                     int64_t* perm, void* workspace, int64_t workspace_bytes,
                     recguru_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Piece 3 -- sampled-negative scoring: fused gather-dot(-loss)               (K3: S1,S2,S3,L*)
 *   replaces  src_emb(dec_out).view(-1,L,1,d), src_emb(n_items).view(-1,L,K,d), the two
 *             batched matmuls and the loss module
 *             AutoEnc4Rec.py:115-131,:219-227 ; AutoEnc4Rec_cross.py:203-215 ;
 *             tools/utils.py:103-126 ; tools/lossfunctions.py:36-49,:56-72,:79-96
 *   p[i]   = h[i,:] . weight[pos[i],:]            i in [0, n_pos)   (n_pos = B*L)
 *   n[i,k] = h[i,:] . weight[neg[i*K+k],:]
 *   loss   = sum_i l_i * mask[i] / sum_i mask[i]   (mask NULL -> mean)
 *   loss_out[0] = loss, loss_out[1] = 1 / sum_i mask[i]  (two floats).
 *   dp_unit[i] = mask[i] * d l_i / d p[i], dn_unit likewise: the per-position derivatives
 *   WITHOUT the 1/sum(mask) factor; consumers pass loss_out+1 as `scale2`.
 *   kind < 0: scores only (loss, dp_unit, dn_unit may be NULL).
 * ---------------------------------------------------------------------------------- */
int64_t recguru_sampled_workspace_bytes(int64_t n_pos, int64_t K);
int recguru_sampled_score_loss_fwd(const float* h, const float* weight, int64_t V1, int64_t d,
                                   const int64_t* pos, const int64_t* neg, int64_t n_pos,
                                   int64_t K, const float* mask, int kind, float* p_out,
                                   float* n_out, float* loss_out, float* dp_unit,
                                   float* dn_unit, void* workspace, int64_t workspace_bytes,
                                   recguru_stream_t stream);
/* Forward and d loss / d h in ONE kernel (training step: padding-aware kinds only, i.e. `kind` carries
 * RECGURU_LOSS_SKIP_MASKED and mask != NULL):
 *   dh[i,:] = grad_scale * (1 / *live_count) * (dp_unit[i] * W[pos[i],:] + sum_k dn_unit[i,k] * W[neg[i,k],:])
 * -- bit-identical to recguru_sampled_score_bwd_dh(dp_unit, dn_unit, grad_scale, loss_out + 1, ...)
 * run afterwards, without re-reading ids, derivatives and table rows.  It needs sum(mask) BEFORE the
 * loss is reduced: live_count points at the count recguru_pad_mask_count produced for this mask
 * (count3 + 2).  grad_scale: device scalar d(total)/d(loss), NULL = 1.                          */
int recguru_sampled_score_loss_fwd_dh(const float* h, const float* weight, int64_t V1, int64_t d,
                                      const int64_t* pos, const int64_t* neg, int64_t n_pos,
                                      int64_t K, const float* mask, int kind, float* p_out,
                                      float* n_out, float* loss_out, float* dp_unit,
                                      float* dn_unit, void* workspace, int64_t workspace_bytes,
                                      const float* grad_scale, const uint32_t* live_count,
                                      float* dh, recguru_stream_t stream);
/* The same with the live positions given as a list: live_rows[0 .. *n_live_rows) = the ascending
 * indices i with mask[i] != 0 (device pointers; the row-mode compaction of the table-backward plan
 * leaves exactly this list, recguru_embedding_bwd_plan_live_rows).  One warp per live position and no
 * compaction inside the kernel; outputs bit-identical to recguru_sampled_score_loss_fwd_dh.  dh may
 * be NULL (forward only). */
int recguru_sampled_score_loss_fwd_dh_rows(const float* h, const float* weight, int64_t V1, int64_t d,
                                           const int64_t* pos, const int64_t* neg, int64_t n_pos,
                                           int64_t K, const float* mask, int kind, float* p_out,
                                           float* n_out, float* loss_out, float* dp_unit,
                                           float* dn_unit, void* workspace, int64_t workspace_bytes,
                                           const float* grad_scale, const uint32_t* live_count,
                                           const uint32_t* live_rows, const uint32_t* n_live_rows,
                                           float* dh, recguru_stream_t stream);
/* Loss only, from logits (drop-in for loss_bpr(p,n,mask) / loss_s_ce(logits,...)). */
int recguru_sampled_loss_fwd(const float* p, const float* n, int64_t n_pos, int64_t K,
                             const float* mask, int kind, float* loss_out, float* dp_unit,
                             float* dn_unit, void* workspace, int64_t workspace_bytes,
                             recguru_stream_t stream);
/* dh[i,:] = s * (dp[i] * weight[pos[i],:] + sum_k dn[i,k] * weight[neg[i*K+k],:]),
 * s = (*scale) * (*scale2), device scalars, each NULL -> 1.  The table gradient goes through
 * recguru_embedding_bwd with sites (pos, X=h, coef=dp, group 1), (neg, h, dn, K).        */
int recguru_sampled_score_bwd_dh(const float* dp, const float* dn, const float* scale,
                                 const float* scale2, const float* weight, int64_t V1, int64_t d, const int64_t* pos,
                                 const int64_t* neg, int64_t n_pos, int64_t K, float* dh,
                                 recguru_stream_t stream);

/* E1 + R1 -- the reference's own (sampled-candidate) evaluation protocol, fused:
 *   get_scores (train_auto.py:164-199; gan_training.py:58-87; tools/utils.py:130-141) followed by
 *   argsort(argsort(-score))[:, 0] (train_auto.py:234-237).
 *   scores[b,0] = u[b].W[target[b]], scores[b,1+c] = u[b].W[cand[b,c]]  (fp32; scores may be NULL)
 *   rank[b] = #{c : scores[b,1+c] > scores[b,0]}   -- column 0 (the target) wins every tie, the
 *   stable reading of the reference idiom.  C = candidate_size (199 / 19999).              */
int recguru_candidate_scores_rank(const float* u, const float* weight, int64_t V1, int64_t d,
                                  const int64_t* target, const int64_t* cand, int64_t B, int64_t C,
                                  float* scores, int64_t* rank, recguru_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Piece 4 -- full-catalog scoring: tcgen05/TMEM GEMM + running top-K + target rank
 *                                                                     (K4: E1', R1, R2)
 *   replaces  torch.matmul(u, src_emb.weight.T)  (AutoEnc4Rec.py:230) followed by
 *             argsort(argsort(-score))[:,0]      (train_auto.py:234-237; gan_training.py:130-133)
 *   Operands are bf16 (cast once with recguru_cast_bf16), accumulation fp32 in TMEM.
 *   items_bf16 holds n_items rows; row r is catalog id  id_lo + r.  Candidates are all of
 *   them (caller excludes pad row 0 / EOS row by choosing the slice, data_loader.py:117-118).
 *   target_score[u] = score of (u, target[u]) from recguru_catalog_target_scores (same
 *   tensor-core data path => exact ties stay exact).
 *   rank_count[u] = #{j in slice: s_j > s_t} + #{j in slice, j < t: s_j == s_t}
 *   topk: descending score, ties -> lower id; padded with (-inf, -1) if n_items < k.
 *   The logit matrix [U x n_items] is never written.  d in {64,128}; 1 <= k <= 32.
 * ---------------------------------------------------------------------------------- */
int recguru_cast_bf16(const float* in, int64_t n, void* out_bf16, recguru_stream_t stream);
int recguru_catalog_target_scores(const void* users_bf16, int64_t U, const void* items_bf16,
                                  int64_t id_lo, int64_t n_items, int64_t d,
                                  const int64_t* target, float* target_score,
                                  recguru_stream_t stream);
int64_t recguru_catalog_workspace_bytes(int64_t U, int64_t n_items, int64_t d, int64_t k);
/* k == 0: rank only (topk_val / topk_id may be NULL) -- the sweep skips the maxima and the lists. */
int recguru_catalog_topk_rank(const void* users_bf16, int64_t U, const void* items_bf16,
                              int64_t id_lo, int64_t n_items, int64_t d, const int64_t* target,
                              const float* target_score, int64_t k, float* topk_val,
                              int64_t* topk_id, int64_t* rank_count, void* workspace,
                              int64_t workspace_bytes, recguru_stream_t stream);
/* recguru_catalog_topk_rank with an explicit rule for items whose score EQUALS the target's
 * (a collapsed model or low-precision scores make exact ties common, and the reference's
 * argsort leaves their order unspecified):
 *   RECGURU_TIE_LOWER_ID     rank = #{s > t} + #{j < target : s_j == t}   (stable descending sort; default)
 *   RECGURU_TIE_TARGET_FIRST rank = #{s > t}            -- the target wins every tie (what the reference's
 *                                                          target-in-column-0 layout yields on CPU)
 *   RECGURU_TIE_TARGET_LAST  rank = #{s >= t} - 1       -- the pessimistic rank
 * TARGET_LAST - TARGET_FIRST = the number of items tied with the target (degenerate-evaluation check).
 * The top-k lists do not depend on the mode.  k <= 16. */
#define RECGURU_TIE_LOWER_ID 0
#define RECGURU_TIE_TARGET_FIRST 1
#define RECGURU_TIE_TARGET_LAST 2
int recguru_catalog_topk_rank_ties(const void* users_bf16, int64_t U, const void* items_bf16,
                                   int64_t id_lo, int64_t n_items, int64_t d, const int64_t* target,
                                   const float* target_score, int64_t k, int tie_mode, float* topk_val,
                                   int64_t* topk_id, int64_t* rank_count, void* workspace,
                                   int64_t workspace_bytes, recguru_stream_t stream);

/* Debug/parity only: dense scores [U, n_items] fp32 through the same tcgen05 path. */
int recguru_catalog_scores_dense(const void* users_bf16, int64_t U, const void* items_bf16,
                                 int64_t n_items, int64_t d, float* scores,
                                 recguru_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Piece 5 -- merge of per-shard results (after the NCCL all-gather)                 (K5)
 *   vals/ids [G,U,k], counts [G,U]  ->  top-k [U,k] (desc score, asc id), rank[U] = sum counts
 * ---------------------------------------------------------------------------------- */
int recguru_topk_merge(const float* vals, const int64_t* ids, const int64_t* counts, int64_t G,
                       int64_t U, int64_t k, float* out_val, int64_t* out_id, int64_t* out_rank,
                       recguru_stream_t stream);

/* ------------------------------------------------------------------------------------
 * H1 -- rank -> HR/NDCG/MRR                                    (tools/metrics.py:26-68)
 *   hist[r] = #{u: rank[u] == r} for r < kmax, hist[kmax] = #{u: rank[u] >= kmax}
 *   (exact integer histogram; the host folds it into HR@k / NDCG@k / MRR@k in fp64).
 *   hist must be zeroed by the caller if accumulate == 0 is not requested.
 * ---------------------------------------------------------------------------------- */
int recguru_rank_histogram(const int64_t* rank, int64_t U, int64_t kmax, int64_t* hist,
                           int accumulate, recguru_stream_t stream);

/* ------------------------------------------------------------------------------------
 * SURVEY 8f rank 2 -- optimizer step of the item table: torch.optim.Adam's arithmetic
 *   (train_auto.py:354-364: betas (0.9, 0.99), eps 1e-9, dense gradient, no weight decay /
 *   amsgrad) in one pass over weight, exp_avg, exp_avg_sq and grad.  `step` counts from 1;
 *   bias corrections are formed in fp64 on the host as torch does.  All buffers [n_floats].
 * ---------------------------------------------------------------------------------- */
int recguru_table_adam(float* weight, float* exp_avg, float* exp_avg_sq, const float* grad,
                       int64_t n_floats, double lr, double beta1, double beta2, double eps,
                       int64_t step, recguru_stream_t stream);

/* Sparse-gradient form of the two calls above ("directly from K2's segmented
 * output without materialising a dense grad"): run recguru_embedding_bwd_apply with
 * RECGURU_DW_PREZEROED on a dW buffer that is NOT zeroed (a row store: only the rows of this
 * step's lookups are written), then
 *   recguru_embedding_bwd_mark_rows: row_epoch[row] = epoch for every row the apply stage wrote
 *     (same sites / workspace as the plan; row_epoch uint32 [V1], zeroed once at allocation,
 *      epochs count from 1 so it never needs clearing);
 *   recguru_table_adam_rows: the Adam arithmetic of recguru_table_adam with
 *     g = row_epoch[row] == epoch ? grad_rows[row, :] : 0 -- bit-identical to the dense call on
 *     a zero-filled gradient, minus the 4*V1*d zero fill and the reads of untouched rows. */
int recguru_embedding_bwd_mark_rows(const recguru_lookup_site_t* sites, int n_sites, int64_t V1,
                                    const void* workspace, int64_t workspace_bytes,
                                    uint32_t* row_epoch, uint32_t epoch, recguru_stream_t stream);
int recguru_table_adam_rows(float* weight, float* exp_avg, float* exp_avg_sq, const float* grad_rows,
                            const uint32_t* row_epoch, uint32_t epoch, int64_t V1, int64_t d,
                            double lr, double beta1, double beta2, double eps, int64_t step,
                            recguru_stream_t stream);

/* ------------------------------------------------------------------------------------
 * SURVEY 8f rank 3 -- negative sampler of the training loader (data/data_loader.py:289-314):
 *   n_items_out[b, :] ~ multinomial(weights_b, num_n, replacement=True), where weights_b is the
 *   base weight vector over ids [0, v_size) -- uniform (alias_* NULL) or any non-negative weights
 *   given as a Walker alias table over v_size entries (item_freq^0.75, :251-254) -- with id 0,
 *   every id of history[b, :] (entries == pad_index ignored), val[b] and test[b] zeroed.
 *   Deterministic in (seed, step): counter-based Philox, independent of launch geometry.
 * ---------------------------------------------------------------------------------- */
int recguru_sample_negatives(const int64_t* history, int64_t B, int64_t H, int64_t pad_index,
                             const int64_t* val, const int64_t* test, int64_t v_size,
                             const float* alias_prob, const int32_t* alias_idx, int64_t num_n,
                             uint64_t seed, uint64_t step, int64_t* n_items_out,
                             recguru_stream_t stream);
/* Same, with the step read from device memory at kernel time (its low 32 bits are the Philox
 * counter word, as above) so that the call can sit in a CUDA graph: replay, then
 * recguru_counter_add(step_dev, 1) inside the same graph, and every replay draws fresh negatives
 * identical to the immediate-step call with that step value. */
int recguru_sample_negatives_devstep(const int64_t* history, int64_t B, int64_t H, int64_t pad_index,
                                     const int64_t* val, const int64_t* test, int64_t v_size,
                                     const float* alias_prob, const int32_t* alias_idx,
                                     int64_t num_n, uint64_t seed, const uint64_t* step_dev,
                                     int64_t* n_items_out, recguru_stream_t stream);
/* Same law with the exclusion list read from the ragged store of recguru_build_batches: the WHOLE
 * sequence of user u = user_idx[b] (items[offsets[u] .. offsets[u+1]), as `for val in seq:
 * weights[seq] = 0` does, data_loader.py:291-293,:301-303 -- not just the window that fits
 * enc_in), val[u] and test[u] (per-user arrays, either may be NULL).  max_len >= the longest stored
 * sequence (<= 12000).  step_dev NULL: the immediate `step` is used.  Same Philox counters as the
 * matrix form: for a user whose sequence fits its enc_in row both calls draw identical negatives. */
int recguru_sample_negatives_csr(const int64_t* items, const int64_t* offsets, const int64_t* user_idx,
                                 int64_t B, int64_t max_len, const int64_t* val, const int64_t* test,
                                 int64_t v_size, const float* alias_prob, const int32_t* alias_idx,
                                 int64_t num_n, uint64_t seed, uint64_t step, const uint64_t* step_dev,
                                 int64_t* n_items_out, recguru_stream_t stream);
/* *slot = %globaltimer (ns) when the stream reaches this point: a one-thread kernel that can sit
 * inside a captured CUDA graph, where events cannot be timed -- the step-timeline tool of
 * scripts/step_timeline.py (which branch of a multi-stream step is the critical one). */
int recguru_debug_timestamp(uint64_t* slot, recguru_stream_t stream);
/* *counter += inc on the stream (a one-thread kernel; graph-capturable step bookkeeping). */
int recguru_counter_add(uint64_t* counter, uint64_t inc, recguru_stream_t stream);
/* ids_out[i] = (int64) ids32_in[i].  The reference's loaders hand over int64 ids (numpy int64 ->
 * default_collate, data/data_loader.py:36,:418-421) although every id is < V1 < 2^31; a loader
 * that collates into pinned int32 halves the per-step host->device traffic (11.5 -> 5.7 MB at
 * cfg2, the PCIe-bound part of the end-to-end step) and widens here.  16-byte aligned buffers. */
int recguru_widen_ids(const int32_t* ids32_in, int64_t n, int64_t* ids_out, recguru_stream_t stream);

/* ------------------------------------------------------------------------------------
 * SURVEY 8f rank 3 (second half) -- batch builder of the loaders:
 *   seq_padding  (data/data_loader.py:25-36)  mode RECGURU_BATCH_TRAIN -> enc_in, dec_in, dec_out
 *   test_seq_gen (data/data_loader.py:39-55)  mode RECGURU_BATCH_EVAL  -> enc_in, dec_in
 *       (call once with append = NULL for (eval_enc_in, dec_in) and once with append = val for
 *        (test_enc_in, t_dec_in): the appended item is the per-user array indexed like val/test)
 *   The dataset is resident in ragged form: items[] (all histories back to back, the "seq" lists
 *   of the pickles) and offsets[n_users + 1]; user_idx[B] selects the batch (NULL = users 0..B-1).
 *   enc_in [B, enc_len], dec_in / dec_out [B, dec_len] int64, any of them may be NULL;
 *   requires 2 <= enc_len, 1 <= dec_len <= enc_len.  val_out / test_out [B] (optional) receive
 *   val[user_idx[b]] / test[user_idx[b]] (the targets the loaders return beside the rows).
 * ---------------------------------------------------------------------------------- */
#define RECGURU_BATCH_TRAIN 0
#define RECGURU_BATCH_EVAL 1
int recguru_build_batches(const int64_t* items, const int64_t* offsets, const int64_t* user_idx,
                          const int64_t* append, int64_t B, int64_t enc_len, int64_t dec_len,
                          int64_t eos, int mode, int64_t* enc_in, int64_t* dec_in, int64_t* dec_out,
                          const int64_t* val, const int64_t* test, int64_t* val_out,
                          int64_t* test_out, recguru_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RECGURU_B200_H_ */
