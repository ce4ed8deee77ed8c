"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (torch-CPU / numpy) of the reference's embedding-and-scoring hot path, one
function per row of SURVEY.md section 8(a), each citing the reference lines it follows.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package; nothing under ``recguru_b200/`` does.

Pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against the reference ITSELF: ``oracle/make_golden.py`` imports the unmodified modules
from /root/reference/GURU in the authoring container, runs them on seeded inputs and stores
inputs + outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function
here against those fixtures (bit-exact where the op is exact).  The arithmetic itself lives in
PyTorch ATen (torch 2.11.0+cu128 CPU kernels here; the reference does not pin a version).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---- P1: positional encoding buffer --- Transformer/transformer.py:95-102 ----------------
def pe_table(max_len: int, d_model: int) -> Tensor:
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0., max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)  # [1, max_len, d]


# ---- M1: pad mask --- AutoEnc4Rec.py:71-72,:89-91,:176-177 ---------------------------------
def pad_mask(ids: Tensor, pad_index: int) -> Tensor:
    mask = ids == pad_index
    return (1 - mask.to(int)).to(torch.float32)


# ---- G1 + P1: src_emb(ids) -> pos_emb(x, mask) before dropout -------------------------------
#      AutoEnc4Rec.py:74-75,:92-93,:180-181 ; Transformer/transformer.py:104-106
def gather_pe_mask(weight: Tensor, ids: Tensor, pe: Tensor, pad_index: int = 0,
                   mask: Optional[Tensor] = None, padding_idx: Optional[int] = None):
    if mask is None:
        mask = pad_mask(ids, pad_index)
    x = F.embedding(ids, weight, padding_idx=padding_idx)
    x = x + pe[:, :x.size(1)]
    return x * torch.unsqueeze(mask, 2), mask


# ---- B1: embedding backward (dense) --- ATen embedding_dense_backward via autograd ----------
def embedding_backward(ids: Tensor, grad_rows: Tensor, V1: int, padding_idx: Optional[int]) -> Tensor:
    """dW[i] = sum of grad_rows over occurrences of i; row padding_idx skipped."""
    w = torch.zeros(V1, grad_rows.shape[-1], requires_grad=True)
    out = F.embedding(ids.reshape(-1), w, padding_idx=padding_idx)
    out.backward(grad_rows.reshape(out.shape))
    return w.grad


def embedding_backward_sites(sites, V1: int, d: int, padding_idx: Optional[int]) -> Tensor:
    """Sum over lookup sites; site = (ids, X[rows,d], coef or None, group).  fp64 accumulation
    (order-free reference value for the toleranced comparison)."""
    dW = np.zeros((V1, d), dtype=np.float64)
    for ids, X, coef, group in sites:
        ids = ids.reshape(-1).numpy()
        rows = np.repeat(X.reshape(-1, d).numpy().astype(np.float64), group, axis=0)
        if coef is not None:
            rows = rows * coef.reshape(-1, 1).numpy().astype(np.float64)
        keep = np.ones_like(ids, dtype=bool) if padding_idx is None else ids != padding_idx
        np.add.at(dW, ids[keep], rows[keep])
    return torch.from_numpy(dW)


def stable_sort_ids(ids: Tensor) -> Tuple[Tensor, Tensor]:
    """The permutation the deterministic backward must reproduce bit-exactly."""
    s, p = torch.sort(ids.reshape(-1), stable=True)
    return s, p


# ---- S1/S2/S3: sampled scoring --- AutoEnc4Rec.py:115-131 ; tools/utils.py:103-124 -----------
def sampled_scores(h: Tensor, weight: Tensor, dec_out: Tensor, n_items: Tensor, L: int, K: int,
                   padding_idx: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    d = weight.shape[1]
    n_emb = F.embedding(n_items, weight, padding_idx=padding_idx).view(-1, L, K, d)
    p_emb = F.embedding(dec_out, weight, padding_idx=padding_idx).view(-1, L, 1, d)
    hh = h.view(-1, L, 1, d)
    p_logits = torch.squeeze(torch.matmul(hh, p_emb.transpose(2, 3)), 2)  # [b, len, 1]
    n_logits = torch.squeeze(torch.matmul(hh, n_emb.transpose(2, 3)), 2)  # [b, len, K]
    return p_logits, n_logits


# ---- Lbpr --- tools/lossfunctions.py:56-72 ----------------------------------------------------
def bpr_loss(p_logit: Tensor, n_logit: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    p = p_logit.view(-1)
    n = torch.mean(n_logit, 2).view(-1)
    loss = -(p - n).sigmoid().log()
    if mask is not None:
        return torch.sum(torch.mul(loss, mask)) / torch.sum(mask)
    return torch.mean(loss)


# ---- Lsas --- tools/lossfunctions.py:79-96 ----------------------------------------------------
def bpr_sas_loss(p_logit: Tensor, n_logit: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    p = p_logit.view(-1)
    n = torch.mean(n_logit, 2).view(-1)
    loss = -((p.sigmoid() + 1e-24).log() + (1 - n.sigmoid() + 1e-24).log())
    if mask is not None:
        return torch.sum(torch.mul(loss, mask)) / torch.sum(mask)
    return torch.mean(loss)


# ---- Lce --- tools/lossfunctions.py:36-49 with label 0 (train_auto.py:44-51) ------------------
def sampled_ce_loss(logit: Tensor, num_class: int, mask: Optional[Tensor] = None) -> Tensor:
    lg = logit.view(-1, num_class)
    label = torch.zeros(lg.shape[0], dtype=torch.long)
    loss = F.cross_entropy(lg, label, reduction="none")
    if mask is not None:
        return torch.sum(torch.mul(loss, mask)) / torch.sum(mask)
    return torch.mean(loss)


LOSSES = {
    "bpr": lambda p, n, m: bpr_loss(p, n, m),
    "bpr_sas": lambda p, n, m: bpr_sas_loss(p, n, m),
    "sce": lambda p, n, m: sampled_ce_loss(torch.cat([p, n], dim=2), n.shape[2] + 1, m),
}


# ---- E1 --- get_scores tail, train_auto.py:196-199 ---------------------------------------------
def candidate_scores(u: Tensor, weight: Tensor, target: Tensor, cand: Tensor, C: int) -> Tensor:
    d = weight.shape[1]
    n_emb = F.embedding(cand, weight).view(-1, C, d)
    p_emb = F.embedding(target, weight).view(-1, d)
    cand_emb = torch.cat([torch.unsqueeze(p_emb, 1), n_emb], dim=1)
    scores = torch.matmul(torch.unsqueeze(u, 1), cand_emb.transpose(1, 2))
    return torch.squeeze(scores)


# ---- R1 --- train_auto.py:234-237 -----------------------------------------------------------------
def rank_of_column0(score: Tensor) -> Tensor:
    """The reference idiom.  On CPU torch.argsort is stable here, so column 0 wins ties."""
    return torch.argsort(torch.argsort(-score, dim=1), dim=1)[:, 0]


# ---- E1' --- AutoEnc4Rec.py:230 at catalog scale, operands rounded to bf16 ---------------------
def catalog_scores(users: Tensor, items: Tensor) -> Tensor:
    """S = bf16(users) @ bf16(items)^T with exact products and fp32 accumulation semantics.
    Computed in fp64 and rounded once to fp32: on grid-valued data (k/16) every partial sum is
    exactly representable, so this equals any fp32 accumulation order bit for bit."""
    a = users.to(torch.bfloat16).to(torch.float64)
    b = items.to(torch.bfloat16).to(torch.float64)
    return (a @ b.t()).to(torch.float32)


# ---- R1 in catalog layout --- SURVEY.md section 8a row R1 -------------------------------------
def catalog_rank(scores: Tensor, target_col: Tensor) -> Tensor:
    """rank = #{j: s_j > s_t} + #{j < t: s_j == s_t}  (ties -> lower id first)."""
    U = scores.shape[0]
    st = scores[torch.arange(U), target_col].unsqueeze(1)
    cols = torch.arange(scores.shape[1]).unsqueeze(0)
    gt = (scores > st).sum(1)
    tie = ((scores == st) & (cols < target_col.unsqueeze(1))).sum(1)
    return (gt + tie).to(torch.int64)


def catalog_rank_argsort(scores: Tensor, target_col: Tensor) -> Tensor:
    """Same rank through the reference's own double-argsort idiom applied to the catalog row
    (stable sort => ties -> lower column first); pins catalog_rank to train_auto.py:234-237."""
    order = torch.argsort(-scores, dim=1, stable=True)
    pos = torch.argsort(order, dim=1, stable=True)
    return pos[torch.arange(scores.shape[0]), target_col]


# ---- R2 --- top-K, descending, ties -> lower id ------------------------------------------------
def catalog_topk(scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    order = torch.argsort(-scores, dim=1, stable=True)[:, :k]
    return torch.gather(scores, 1, order), order


# ---- H1 --- tools/metrics.py:26-68 ---------------------------------------------------------------
def hit_at_k_batch(r, k):
    r = (np.array(r) < k) + 0
    return float(sum(r)) / len(r)


def NDCG_at_k_batch(r, k):
    ndcg = 0
    for rank in r:
        ndcg += (1 / np.log2(rank + 2)) if rank < k else 0
    return float(ndcg) / len(r)


def mrr_at_k_batch(r, k):
    mrr = 0
    for rank in r:
        mrr += (1.0 / (rank + 1)) if rank < k else 0.0
    return float(mrr) / len(r)


# ---- eval driver --- train_auto.py:202-253 (gan_training.py:90-149 is the same loop per domain) --
def reference_evaluation(user_rep, weight: Tensor, batches, candidate_size: int, k_val=(5, 10, 20, 30)):
    """The reference loop restated: four get_scores calls per batch (each recomputing the user
    representation), double argsort, ranks appended to Python lists, metrics by per-user loops.
    ``user_rep(enc_in, dec_in) -> [B, d]`` stands for ``get_embedding*(...)[:, -1, :]``."""
    ranks = {"eval_f": [], "eval_r": [], "test_f": [], "test_r": []}
    for eval_data, test_data, n_items_f, n_items_r in batches:
        for split, (enc_in, dec_in, target) in (("eval", eval_data), ("test", test_data)):
            for s, cand in (("f", n_items_f), ("r", n_items_r)):
                u = user_rep(enc_in, dec_in)
                score = candidate_scores(u, weight, target.reshape(-1), cand.reshape(-1), candidate_size)
                ranks[f"{split}_{s}"].extend(rank_of_column0(score.reshape(u.shape[0], -1)).numpy())
    out = []
    for s in ("f", "r"):
        res = {}
        for k in k_val:
            res[str(k)] = {}
            for split in ("eval", "test"):
                r = ranks[f"{split}_{s}"]
                res[str(k)][f"ht_{split}"] = [hit_at_k_batch(r, k)]
                res[str(k)][f"ndcg_{split}"] = [NDCG_at_k_batch(r, k)]
                res[str(k)][f"mrr_{split}"] = [mrr_at_k_batch(r, k)]
        out.append(res)
    return out


# ---- data layout helpers --- data/data_loader.py:25-36,:39-55 ---------------------------------
def seq_padding(seq, length_enc, len_dec, eos):
    """Left-pad with 0, EOS last; dec_in / dec_out are enc_in shifted by two / one."""
    seq = list(seq)
    if len(seq) >= length_enc:
        enc_in = seq[-length_enc + 1:] + [eos]
    else:
        enc_in = [0] * (length_enc - len(seq) - 1) + seq + [eos]
    dec_in = ([0, 0] + enc_in[0:-2])[-len_dec:]
    dec_out = ([0, 0] + enc_in[1:-1])[-len_dec:]
    return np.array(enc_in), np.array(dec_in), np.array(dec_out)


def test_seq_gen(seq, length_enc, len_dec, eos, val):
    seq = list(seq)
    seq_test = seq + [val]

    def enc(s):
        if len(s) >= length_enc:
            return s[-length_enc + 1:] + [eos]
        return [0] * (length_enc - len(s) - 1) + s + [eos]

    eval_enc_in = enc(seq)
    dec_in = ([0] + eval_enc_in[0:-1])[-len_dec:]
    test_enc_in = enc(seq_test)
    t_dec_in = ([0] + test_enc_in[0:-1])[-len_dec:]
    return np.array(eval_enc_in), np.array(dec_in), np.array(test_enc_in), np.array(t_dec_in)


test_seq_gen.__test__ = False  # not a pytest test


def sampling_weights(seq, val, test, v_size, item_freq=None, exclude_val=False):
    """The weight vector ``pickle_loader.__getitem__`` hands to ``torch.multinomial``
    (data/data_loader.py:289-314), restated.  The loop ``for val in seq: weights[seq] = 0`` rebinds
    ``val`` to ``seq[-1]`` (:291-292,:303-304), so the following ``weights[val] = 0`` does NOT zero the
    validation item; in the ``item_freq`` branch (:289-295) ``weights[0]`` keeps ``wf[0]**0.75 / sum``.
    ``exclude_val=True`` is the repaired law (validation item excluded) that NegativeSampler defaults to."""
    import numpy as np
    seq = np.asarray(seq, dtype=np.int64)
    if item_freq is not None:
        w = np.power(np.asarray(item_freq, dtype=np.float64), 0.75)
        w = (w / w.sum()).astype(np.float32).astype(np.float64)       # FT(wf): a float32 tensor (:252-254)
    else:
        w = np.ones(v_size, dtype=np.float64)
        w[0] = 0
    w[seq] = 0
    w[seq[-1] if (len(seq) and not exclude_val) else val] = 0
    w[test] = 0
    return w
