"""Evaluation driver: the caller of the scoring/ranking kernels.

Drop-in for the reference's ``evaluation(model_train, data_loader, de, param_config, k_val, sas)``
(train_auto.py:202-253; gan_training.py:90-149 is the same loop per domain): same arguments, same
return value ``[result_freq, result_rand]`` with keys ``ht_/ndcg_/mrr_`` x ``eval/test`` per k.

What changes underneath:

* the reference runs the transformer FOUR times per batch (``get_scores`` for eval/test x
  popularity-sampled/random candidates recomputes the same user representation); here the user
  representation is computed once per split and both candidate sets are ranked from it;
* ``get_scores`` + double argsort is one kernel (``recguru_candidate_scores_rank``): the
  ``[B, C+1, d]`` candidate tensor and the ``[B, C+1]`` score matrix are never materialised;
* ranks never leave the device: each batch is folded into four exact integer rank histograms
  (``recguru_rank_histogram`` in accumulate mode); ONE ``4 x (kmax+1)`` int64 D2H copy at the end
  replaces 4 ``.cpu()`` syncs per batch and the per-user Python loops of tools/metrics.py:49-68.

The transformer itself (``get_embedding`` / ``get_embedding_sas``) is the reference's and is only
called here.  ``full_catalog=True`` ranks against every item with the tcgen05 sweep instead of the
sampled candidates (then ``n_items_f/_r`` are ignored and both result dicts are equal).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

from . import catalog, ops

Tensor = torch.Tensor
_SPLITS = ("eval", "test")


def _unwrap(model):
    """nn.DataParallel wraps the model when several GPUs are visible (train_auto.py:350-351)."""
    return model.module if hasattr(model, "module") else model


def _last_position(h: Tensor) -> Tensor:
    """``h[:, -1, :]`` -- one small kernel (recguru_extract_last) on the GPU, a view elsewhere."""
    if h.is_cuda and h.dtype == torch.float32 and h.dim() == 3 and not h.requires_grad:
        return ops.raw_extract_last(h)[0]
    return h[:, -1, :]


def reference_user_rep(model, sas: bool) -> Callable[[Tensor, Tensor], Tensor]:
    """``u = get_embedding*(...)[:, -1, :]`` exactly as ``get_scores`` takes it (train_auto.py:164-194)."""
    m = _unwrap(model)
    if sas:
        return lambda enc_in, dec_in: _last_position(m.get_embedding_sas(dec_in))
    return lambda enc_in, dec_in: _last_position(m.get_embedding(enc_in, dec_in))


def reference_item_table(model, sas: bool) -> Tensor:
    m = _unwrap(model)
    return (m.src_emb if sas else m.AutoEnc.src_emb).weight


class RankAccumulator:
    """Exact integer rank histograms kept on the device, one per (sampler, split)."""

    def __init__(self, k_val: Sequence[int], device):
        self.k_val = [int(k) for k in k_val]
        self.kmax = max(self.k_val)
        self.hist: Dict[str, Tensor] = {}
        self.count: Dict[str, int] = {}
        self.device = device

    # the two device operations (host-logic tests override them with the oracle)
    def _histogram(self, rank: Tensor, hist: Optional[Tensor]) -> Tensor:
        return ops.raw_rank_histogram(rank, self.kmax, hist=hist, accumulate=hist is not None)

    def add(self, key: str, rank: Tensor) -> None:
        self.hist[key] = self._histogram(rank, self.hist.get(key))
        self.count[key] = self.count.get(key, 0) + int(rank.numel())

    def metrics(self, key: str) -> Dict[int, Dict[str, float]]:
        return catalog.metrics_from_histogram(self.hist[key], self.count[key], self.k_val)

    def all_metrics(self) -> Dict[str, Dict[int, Dict[str, float]]]:
        keys = sorted(self.hist)
        if not keys:
            return {}
        stacked = torch.stack([self.hist[k] for k in keys]).cpu()      # the one D2H copy
        return {k: catalog.metrics_from_histogram(stacked[i], self.count[k], self.k_val)
                for i, k in enumerate(keys)}


class EvalDriver:
    """Accumulates one evaluation pass.  ``rank_fn(u, weight, target, cand) -> rank`` defaults to
    the fused candidate kernel; ``evaluator`` (a :class:`CatalogEvaluator`) switches to full-catalog
    ranking."""

    def __init__(self, user_rep: Callable[[Tensor, Tensor], Tensor], weight: Tensor, device,
                 k_val: Optional[Sequence[int]] = None, evaluator: Optional[catalog.CatalogEvaluator] = None,
                 accumulator: Optional[RankAccumulator] = None):
        self.user_rep = user_rep
        self.weight = weight
        self.device = device
        self.k_val = [5, 10, 20, 30] if k_val is None else list(k_val)     # train_auto.py:203-204
        self.evaluator = evaluator
        self.acc = accumulator or RankAccumulator(self.k_val, device)

    def _rank(self, u: Tensor, target: Tensor, cand: Optional[Tensor]) -> Tensor:
        if self.evaluator is not None:
            return self.evaluator.target_rank(u, target)
        return catalog.get_scores_rank(u, self.weight, target, cand.view(u.shape[0], -1))[1]

    @torch.no_grad()
    def add_batch(self, data_batch) -> None:
        """One element of the reference's eval loader: ``(eval_data, test_data, n_items_f, n_items_r)``
        with ``*_data = (enc_in, dec_in, target)`` (data_loader.py:319-384)."""
        eval_data, test_data, n_items_f, n_items_r = data_batch[0], data_batch[1], data_batch[2], data_batch[3]
        de = self.device
        cands = {"f": n_items_f.to(de, non_blocking=True), "r": n_items_r.to(de, non_blocking=True)}
        for split, data in zip(_SPLITS, (eval_data, test_data)):
            enc_in, dec_in, target = (t.to(de, non_blocking=True) for t in data[:3])
            u = self.user_rep(enc_in, dec_in).contiguous().float()        # ONE transformer forward per split
            target = target.view(-1)
            if self.evaluator is not None:
                r = self._rank(u, target, None)
                self.acc.add(f"f_{split}", r)
                self.acc.add(f"r_{split}", r)
            else:
                for s in ("f", "r"):
                    self.acc.add(f"{s}_{split}", self._rank(u, target, cands[s]))

    def results(self) -> List[Dict[str, Dict[str, List[float]]]]:
        """``[result_freq, result_rand]`` in the reference's layout (train_auto.py:213-252)."""
        met = self.acc.all_metrics()
        out = []
        for s in ("f", "r"):
            res = {}
            for k in self.k_val:
                d = {}
                for split in _SPLITS:
                    m = met[f"{s}_{split}"][k]
                    d[f"ht_{split}"] = [m["hr"]]
                    d[f"ndcg_{split}"] = [m["ndcg"]]
                    d[f"mrr_{split}"] = [m["mrr"]]
                res[str(k)] = d
            out.append(res)
        return out


def evaluation(model_train, data_loader: Iterable, de, param_config=None, k_val: Optional[Sequence[int]] = None,
               sas: bool = False, full_catalog: bool = False, tie: str = "id"):
    """Same call as the reference's ``evaluation`` (train_auto.py:202); see the module docstring.

    Protocol notes.  Default (``full_catalog=False``): the reference's own protocol -- the target against
    the loader's sampled candidates, the target winning ties (its column-0 position under a stable
    argsort; the reference's CUDA argsort leaves tie order unspecified).  ``full_catalog=True`` is a
    DIFFERENT protocol and its numbers are not comparable with the reference's: every user is ranked
    against all items, including the items of the user's own history, so ``result_freq`` and
    ``result_rand`` (which differ only in how candidates were sampled) come out identical; ties are
    resolved by ``tie`` ("id": lower id first, "target_first", "target_last" -- see
    :class:`recguru_b200.catalog.CatalogEvaluator`; ``CatalogEvaluator.tie_bracket`` measures how many
    items tie with each target, the check for a degenerate evaluation)."""
    model_train.eval()
    weight = reference_item_table(model_train, sas).detach()
    ev = catalog.CatalogEvaluator(weight, k=10, tie=tie) if full_catalog else None
    drv = EvalDriver(reference_user_rep(model_train, sas), weight, de, k_val, evaluator=ev)
    for data_batch in data_loader:
        drv.add_batch(data_batch)
    return drv.results()
