"""Drop-in modules mirroring the reference's PyTorch surface for the hot path.

Same names / argument meaning / return shapes as the reference so that
``model.AutoEnc.src_emb = FusedItemEmbedding.from_embedding(model.AutoEnc.src_emb)`` and
``model.AutoEnc.pos_emb = FusedPositionalEncoding.from_module(model.AutoEnc.pos_emb)`` keep every
call site in AutoEnc4Rec.py / AutoEnc4Rec_cross.py / train_auto.py / tools/utils.py working
(SURVEY.md section 8b).  All compute goes through ``torch.ops.recguru.*`` (CUDA only).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import ops  # noqa: F401  (registers torch.ops.recguru)

Tensor = torch.Tensor


def sinusoid_table(max_len: int, d_model: int) -> Tensor:
    """The reference's fixed PE buffer, built with the same fp32 torch ops so that the values
    are bit-identical (Transformer/transformer.py:95-102).  Returns [1, max_len, d_model]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0., max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


class FusedItemEmbedding(nn.Module):
    """``nn.Embedding(vocab_size + 1, d_model, padding_idx=...)`` replacement (T1/G1/S1/S2).

    Reference tables: AutoEnc4Rec.py:31,:153 (padding_idx=pad_index) and
    AutoEnc4Rec_cross.py:32,:35 (no padding_idx).  ``weight`` keeps its name so
    ``state_dict()`` keys are unchanged."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None,
                 _weight: Optional[Tensor] = None, one_pass_backward: Optional[bool] = None):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.padding_idx = padding_idx
        # one deterministic table backward per backward() for ALL lookups of this table (input ids,
        # positives, negatives) instead of one dense [V1, d] gradient per lookup site summed by
        # autograd: see ops.enable_one_pass_backward.  RECGURU_ONE_PASS_BWD=0 turns the default off.
        self.one_pass_backward = ops.ONE_PASS_DEFAULT if one_pass_backward is None else bool(one_pass_backward)
        if _weight is None:
            w = torch.empty(num_embeddings, embedding_dim)
            nn.init.normal_(w)  # nn.Embedding.reset_parameters
            if padding_idx is not None:
                with torch.no_grad():
                    w[padding_idx].fill_(0)
            self.weight = nn.Parameter(w)
        else:
            self.weight = nn.Parameter(_weight)

    @classmethod
    def from_embedding(cls, emb: nn.Embedding) -> "FusedItemEmbedding":
        m = cls(emb.num_embeddings, emb.embedding_dim, emb.padding_idx, _weight=emb.weight.data)
        return m

    def _register(self) -> None:
        # keyed by the storage address: re-registered on every call because .to()/.cuda() move it
        if self.one_pass_backward and self.weight.is_cuda and torch.is_grad_enabled() and self.weight.requires_grad \
                and type(self.weight.data) is Tensor:
            ops.enable_one_pass_backward(self.weight, self.padding_idx)

    def forward(self, ids: Tensor) -> Tensor:
        pad = -1 if self.padding_idx is None else int(self.padding_idx)
        self._register()
        return torch.ops.recguru.gather_rows(self.weight, ids, pad)

    def lookup_pe_mask(self, ids: Tensor, pos_emb: "FusedPositionalEncoding",
                       pad_index: Optional[int] = 0, mask: Optional[Tensor] = None,
                       fused_dropout: Optional[bool] = None):
        """Fused ``pos_emb(src_emb(ids), mask)`` (K1).  With ``mask=None`` the 0/1 mask is derived
        from ``ids != pad_index`` inside the kernel and returned: ``(out, mask)``.

        In training mode the module's ``nn.Dropout`` is fused into the same kernel
        (``fused_dropout``, default on): it draws from torch's CUDA generator exactly like
        ``nn.Dropout`` would -- same kept elements, same generator offset afterwards."""
        self._register()
        table = pos_emb.table(ids.shape[-1])
        p = float(pos_emb.dropout.p)
        if fused_dropout is None:
            fused_dropout = pos_emb.fused_dropout
        if fused_dropout and pos_emb.training and 0.0 < p < 1.0 and type(self.weight.data) is Tensor \
                and self.weight.is_cuda and not torch.cuda.is_current_stream_capturing():
            seed, off, threads = ops.torch_dropout_state(ids.numel() * self.embedding_dim, self.weight.device)
            out, m = torch.ops.recguru.gather_pe_mask_dropout(self.weight, ids, table, mask, int(pad_index), p,
                                                              seed, off, threads)
            return out, m
        if mask is None:
            out, m = torch.ops.recguru.gather_pe_mask(self.weight, ids, table, int(pad_index))
        else:
            out = torch.ops.recguru.gather_pe_mask_m(self.weight, ids, table, mask)
            m = mask
        return pos_emb.dropout(out), m

    def extra_repr(self) -> str:
        return f"{self.num_embeddings}, {self.embedding_dim}, padding_idx={self.padding_idx}"


class FusedPositionalEncoding(nn.Module):
    """``PositionalEncoding(d_model, dropout, max_len)`` (Transformer/transformer.py:87-106):
    ``forward(x, mask) = dropout((x + pe[:, :L]) * mask[..., None])``.  The add+mask runs in one
    kernel; dropout stays with torch so the RNG stream matches the reference."""

    def __init__(self, d_model: int, dropout: float, max_len: int = 5000, fused_dropout: bool = True):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.fused_dropout = fused_dropout
        self.register_buffer("pe", sinusoid_table(max_len, d_model))

    def table(self, L: int) -> Tensor:
        """The [>=L, d] position table the kernels add (fixed sinusoid buffer)."""
        return self.pe

    @classmethod
    def from_module(cls, mod: nn.Module) -> "FusedPositionalEncoding":
        pe = mod.pe
        m = cls(pe.shape[-1], mod.dropout.p, max_len=pe.shape[1])
        m.pe.copy_(pe)
        return m

    def forward(self, x: Tensor, mask: Tensor) -> Tensor:
        y = torch.ops.recguru.pe_mask(x, self.table(x.shape[-2]), mask)
        p = float(self.dropout.p)
        if self.fused_dropout and self.training and 0.0 < p < 1.0 and x.is_cuda and x.shape[-1] % 4 == 0 \
                and type(x.data) is Tensor and not torch.cuda.is_current_stream_capturing():
            # same kept elements and same generator consumption as self.dropout(y), without the byte mask
            seed, off, threads = ops.torch_dropout_state(y.numel(), y.device)
            return torch.ops.recguru.dropout(y, None, p, seed, off, threads)
        return self.dropout(y)


class FusedPositionalEncodingM(FusedPositionalEncoding):
    """``PositionalEncodingM(d_model, dropout, max_len, device)`` (Transformer/transformer.py:327-342):
    LEARNED positions -- ``forward(x, mask) = dropout((x + pos_emb(arange(L))) * mask[..., None])``.
    Same kernels as the sinusoid module (they add any ``[L, d]`` table); the table's gradient
    ``dP[l] = sum_b mask[b,l] * g[b,l]`` comes from ``recguru_pe_grad`` (fixed summation order).
    ``pos_emb`` keeps its name, so ``state_dict()`` keys match the reference (``pos_emb.weight``)."""

    def __init__(self, d_model: int, dropout: float, max_len: int = 5000, device=None, fused_dropout: bool = True):
        nn.Module.__init__(self)
        self.dropout = nn.Dropout(p=dropout)
        self.fused_dropout = fused_dropout
        self.device = device
        self.pos_emb = nn.Embedding(max_len, d_model)

    @classmethod
    def from_module(cls, mod: nn.Module) -> "FusedPositionalEncodingM":
        w = mod.pos_emb.weight
        m = cls(w.shape[1], mod.dropout.p, max_len=w.shape[0], device=getattr(mod, "device", None))
        m.pos_emb.weight = w                          # shares the Parameter
        return m

    def table(self, L: int) -> Tensor:
        return self.pos_emb.weight[:L]


class _FusedLogitLoss(nn.Module):
    kind = ""

    def forward(self, p_logit: Tensor, n_logit: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        aux, _, _ = torch.ops.recguru.sampled_loss(p_logit, n_logit, mask, self.kind)
        return aux[0]


class FusedBPRLoss(_FusedLogitLoss):
    """tools/lossfunctions.py:52-72 -- ``-log(sigmoid(p - mean_k n_k))``, masked mean."""
    kind = "bpr"


class FusedBPRLoss_sas(_FusedLogitLoss):
    """tools/lossfunctions.py:75-96 -- ``-(log(sig(p)+1e-24) + log(1-sig(mean n)+1e-24))``."""
    kind = "bpr_sas"


class FusedSampledCrossEntropyLoss(nn.Module):
    """tools/lossfunctions.py:29-49 with the sampled-negative convention of loss_ae
    (train_auto.py:44-51): logits [B, L, 1+K], label 0 == the positive in column 0."""

    def forward(self, logit: Tensor, label: Optional[Tensor], num_class: int,
                mask: Optional[Tensor] = None) -> Tensor:
        lg = logit.reshape(-1, num_class)
        p = lg[:, :1].contiguous()
        n = lg[:, 1:].contiguous()
        aux, _, _ = torch.ops.recguru.sampled_loss(p, n, mask, "sce")
        return aux[0]


def sampled_logits(h: Tensor, table: FusedItemEmbedding, dec_out: Tensor, n_items: Tensor):
    """The scoring tail of MyRec.forward (AutoEnc4Rec.py:115-131) / loss_bpr_func
    (tools/utils.py:103-124) without the gathered-row intermediates:
    returns ``(p_logits [B,L,1], n_logits [B,L,K])``."""
    pad = -1 if table.padding_idx is None else int(table.padding_idx)
    table._register()
    return torch.ops.recguru.sampled_scores(h, table.weight, dec_out, n_items, pad)


def sampled_logits_cat(h: Tensor, table: FusedItemEmbedding, dec_out: Tensor, n_items: Tensor) -> Tensor:
    """MyAuto4Rec.forward / MyAuto4Rec_c.forward tail (AutoEnc4Rec.py:219-227,
    AutoEnc4Rec_cross.py:203-215): ``cat([p, n], dim=2)`` -> [B, L, 1+K]."""
    p, n = sampled_logits(h, table, dec_out, n_items)
    return torch.cat([p, n], dim=2)


def sampled_loss_fused(h: Tensor, table: FusedItemEmbedding, dec_out: Tensor, n_items: Tensor,
                       mask: Optional[Tensor], kind: str, skip_masked: bool = True):
    """Scores + loss in one kernel (K3).  ``kind`` in {"bpr", "bpr_sas", "sce"}.
    Returns ``(loss, p_logits, n_logits)``; backward produces dh and the dense table grad.

    ``skip_masked`` (default): positions with ``mask == 0`` are not scored -- the reference scores them and
    multiplies their loss term by 0 (tools/lossfunctions.py:44-48,:66-71,:90-95), so the loss and every
    gradient are unchanged; only the returned logits of those positions are 0 instead of values nothing
    downstream may use.  Pass ``skip_masked=False`` to get every logit."""
    pad = -1 if table.padding_idx is None else int(table.padding_idx)
    table._register()
    aux, p, n, _, _ = torch.ops.recguru.sampled_score_loss(h, table.weight, dec_out, n_items, mask,
                                                           kind + "/skip" if (skip_masked and mask is not None) else kind,
                                                           pad)
    return aux[0], p, n
