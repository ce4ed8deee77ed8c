"""The training-step hot path as one object: all lookup sites of a step, fused.

A reference training step touches the item table at its two edges
(SURVEY.md section 3.5): ids -> (W[ids] + pe) * mask -> [transformer, not ours] -> h ->
(h . W[pos], h . W[neg]) -> loss, and back.  ``TrainHotPath`` runs exactly those edges with
pre-allocated buffers and no per-step allocation, and produces ONE dense table gradient for all
lookup sites (input + positive + negative) in a single deterministic sort/segmented-reduce pass,
instead of one dense [V1, d] gradient per site summed by autograd.

It is driven by whatever sits between the edges (the reference's EncoderM/DecoderM): call
:meth:`forward_in`, run the transformer on the returned activations, call :meth:`forward_loss`
on its output, then :meth:`backward` with the gradient that came back to the input activations.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import ops
from .ops import Site

Tensor = torch.Tensor


class TrainHotPath:
    """SASRec / AutoRec BPR step edges (MyRec.forward, AutoEnc4Rec.py:98-133 + BPRLoss_sas)."""

    def __init__(self, weight: Tensor, pe: Tensor, B: int, L: int, K: int, kind: str = "bpr_sas",
                 pad_index: int = 0, padding_idx: Optional[int] = 0, skip_masked: bool = True,
                 overlap_plan: bool = True, dense_grad: bool = True, fuse_dh: Optional[bool] = None):
        assert weight.is_cuda and weight.dtype == torch.float32
        self.weight = weight
        self.pe = pe[0] if pe.dim() == 3 else pe
        self.B, self.L, self.K, self.kind = B, L, K, kind
        self.pad_index = pad_index
        self.skip_masked = skip_masked
        self.padding_idx = -1 if padding_idx is None else padding_idx
        V1, d = weight.shape
        self.V1, self.d = V1, d
        dev = weight.device
        n = B * L
        f32 = dict(dtype=torch.float32, device=dev)
        self.x0 = torch.empty(B, L, d, **f32)
        self.mask = torch.empty(B, L, **f32)
        self.sbuf = {
            "p": torch.empty(B, L, 1, **f32), "n": torch.empty(B, L, K, **f32),
            "aux": torch.empty(2, **f32), "dp": torch.empty(n, **f32), "dn": torch.empty(n, K, **f32),
            "ws": torch.empty(max(256, int(ops._lib.load().recguru_sampled_workspace_bytes(n, K))),
                              dtype=torch.uint8, device=dev),
        }
        self.dh = torch.empty(B, L, d, **f32)
        self.dW = torch.empty(V1, d, **f32)
        # dense_grad=False: dW is a row store -- no zero fill, only the rows of this
        # step's lookups are written and stamped row_epoch[row] = epoch; consumed by
        # FusedTableAdam.step_rows.  dense_grad=True keeps the reference's dense-gradient contract.
        self.dense_grad = dense_grad
        self.row_epoch = None if dense_grad else torch.zeros(V1, dtype=torch.int32, device=dev)
        self.epoch = 0
        total = n * (2 + K)
        self.ws_bwd = torch.empty(ops.embedding_bwd_workspace_bytes(total, V1, d), dtype=torch.uint8, device=dev)
        self.one = torch.ones(1, **f32)
        # side stream for the part of the table backward that does not depend on activations
        self.overlap_plan = overlap_plan
        # high priority: its short kernels slip in between the waves of the bandwidth-bound kernels
        self._side = torch.cuda.Stream(device=dev, priority=-1)
        # the dense zero fill of dW is bandwidth-bound and independent of the plan's latency-bound
        # chain: its own branch (RECGURU_ZERO_STREAM=0 puts it back at the head of the plan stream)
        self._zero_own = os.environ.get("RECGURU_ZERO_STREAM", "1") != "0"
        self._zstream = torch.cuda.Stream(device=dev)
        self._zero_done = torch.cuda.Event()
        self._fork = torch.cuda.Event()
        self._join = torch.cuda.Event()
        self._planned = False
        self._side_mask = torch.empty(B, L, **f32)
        # fuse_dh: d loss / d h comes out of the padding-aware scoring kernel itself
        # (recguru_sampled_score_loss_fwd_dh) instead of a second kernel that re-reads ids,
        # derivatives and table rows.  It needs sum(mask) before the loss is reduced: the pad-mask
        # kernel that runs at the start of the step anyway also counts (recguru_pad_mask_count).
        if fuse_dh is None:                                # RECGURU_FUSE_DH=0: two-kernel path by default (A/B timing)
            fuse_dh = skip_masked and os.environ.get("RECGURU_FUSE_DH", "1") != "0"
        self.fuse_dh = bool(fuse_dh)
        if self.fuse_dh and not skip_masked:
            raise RuntimeError("recguru: fuse_dh needs the padding-aware scoring kernel (skip_masked=True)")
        self._cnt = torch.zeros(3, dtype=torch.int32, device=dev)
        self._cnt_ready = torch.cuda.Event()
        self._cnt_state = 0            # 0 = not issued this step, 1 = on the current stream, 2 = on the side stream
        self._dh_done = False
        self._dh_scale = None

    # -- edge 1: ids -> activations ------------------------------------------------------
    def forward_in(self, dec_in: Tensor) -> Tensor:
        ops.raw_gather_pe_mask(self.weight, dec_in, self.pe, self.pad_index, out=self.x0, mask_out=self.mask)
        if self.fuse_dh and self._cnt_state == 0:          # no plan forked: count on this stream
            ops.raw_pad_mask_count(dec_in, self.pad_index, self._side_mask, self._cnt)
            self._cnt_state = 1
        return self.x0

    # -- edge 2: hidden states -> loss ---------------------------------------------------
    def forward_loss(self, h: Tensor, dec_out: Tensor, n_items: Tensor, mask: Optional[Tensor] = None,
                     grad_loss: Optional[Tensor] = None) -> Tensor:
        m = self.mask if mask is None else mask
        # the loss is all this edge exports: positions under the pad mask are not scored
        if self.fuse_dh and mask is None and self._cnt_state != 0:
            if self._cnt_state == 2:
                torch.cuda.current_stream().wait_event(self._cnt_ready)
            ops.raw_sampled_score_loss(h, self.weight, dec_out, n_items, m, self.kind, bufs=self.sbuf,
                                       skip_masked=True, dh_out=self.dh, live_count=self._cnt,
                                       grad_scale=grad_loss)
            self._dh_done = True
            self._dh_scale = grad_loss
        else:
            ops.raw_sampled_score_loss(h, self.weight, dec_out, n_items, m, self.kind, bufs=self.sbuf,
                                       skip_masked=self.skip_masked)
        return self.sbuf["aux"][0]

    # -- table-backward plan: depends on the ids and the pad mask only -----------------------
    def plan_backward(self, dec_in: Tensor, dec_out: Tensor, n_items: Tensor, mask: Optional[Tensor] = None):
        """Issue stage A of the table backward (zero fill of dW, compaction of the live lookups,
        stable radix sort) on a side stream.  It needs nothing but the ids and the pad mask, so it
        runs UNDER the forward pass (in a full model: under the whole transformer forward/backward)
        instead of on the critical path; :meth:`backward` joins it."""
        n = self.B * self.L
        cur = torch.cuda.current_stream()
        self._fork.record(cur)
        self._side.wait_event(self._fork)
        if self.dense_grad and self._zero_own:
            self._zstream.wait_event(self._fork)
            with torch.cuda.stream(self._zstream):
                ops.raw_fill_zero(self.dW)
                self._zero_done.record(self._zstream)
        with torch.cuda.stream(self._side):
            if mask is None:
                # the side stream derives the pad mask itself (a 2 us kernel), so the plan does not
                # have to wait for K1 and can be issued at the very start of the step
                if self.fuse_dh and self._cnt_state == 0:   # (state 1: forward_in already counted on its stream)
                    m = ops.raw_pad_mask_count(dec_in, self.pad_index, self._side_mask, self._cnt).view(n)
                    self._cnt_ready.record(self._side)
                    self._cnt_state = 2
                else:
                    m = ops.raw_pad_mask(dec_in, self.pad_index, out=self._side_mask).view(n)
            else:
                m = mask.view(n)
            if self.dense_grad and not self._zero_own:
                ops.raw_fill_zero(self.dW)
            sites = [Site(dec_in, None, live=m), Site(dec_out, None, live=m),
                     Site(n_items, None, group=self.K, live=m)]
            ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd)
            if not self.dense_grad:
                self.epoch += 1
                ops.raw_embedding_bwd_mark_rows(sites, self.V1, self.ws_bwd, self.row_epoch, self.epoch)
            self._join.record(self._side)
        self._planned = True

    # -- backward of both edges ------------------------------------------------------------
    def backward(self, h: Tensor, dec_in: Tensor, dec_out: Tensor, n_items: Tensor, g_x0: Tensor,
                 grad_loss: Optional[Tensor] = None):
        """Returns (dh, dW).  ``g_x0`` is d loss / d x0 as it comes back through the transformer."""
        go = self.one if grad_loss is None else grad_loss
        inv_m = self.sbuf["aux"][1:2]
        if self._dh_done and (grad_loss is None) == (self._dh_scale is None) and \
                (grad_loss is None or grad_loss.data_ptr() == self._dh_scale.data_ptr()):
            pass                                           # dh was written by the forward kernel
        else:
            ops.raw_sampled_score_bwd_dh(self.sbuf["dp"], self.sbuf["dn"], self.weight, dec_out, n_items,
                                         scale=go, scale2=inv_m, out=self.dh)
        self._dh_done = False
        self._cnt_state = 0
        n = self.B * self.L
        h2 = h.view(n, self.d)
        m = self.mask.view(n)
        sites = [
            Site(dec_in, g_x0.view(n, self.d), coef=m, live=m),                                       # input site
            Site(dec_out, h2, coef=self.sbuf["dp"], scale=go, scale2=inv_m, live=m),                   # positives
            Site(n_items, h2, coef=self.sbuf["dn"].view(-1), scale=go, scale2=inv_m, group=self.K, live=m),  # negatives
        ]
        if self._planned:
            torch.cuda.current_stream().wait_event(self._join)
            if self.dense_grad and self._zero_own:
                torch.cuda.current_stream().wait_event(self._zero_done)
            ops.raw_embedding_bwd_apply(sites, self.V1, self.d, self.dW, self.ws_bwd, mode=ops.DW_PREZEROED)
            self._planned = False
        elif self.dense_grad:
            ops.raw_embedding_bwd(sites, self.V1, self.d, self.padding_idx, out=self.dW, zero_fill=True,
                                  workspace=self.ws_bwd)
        else:
            ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd)
            ops.raw_embedding_bwd_apply(sites, self.V1, self.d, self.dW, self.ws_bwd, mode=ops.DW_PREZEROED)
            self.epoch += 1
            ops.raw_embedding_bwd_mark_rows(sites, self.V1, self.ws_bwd, self.row_epoch, self.epoch)
        return self.dh, self.dW

    def step(self, dec_in: Tensor, dec_out: Tensor, n_items: Tensor, h: Tensor, g_x0: Tensor):
        """All edges of one step with the transformer replaced by given (h, g_x0)."""
        if self.overlap_plan:
            self.plan_backward(dec_in, dec_out, n_items)      # needs the ids only: forked first
        self.forward_in(dec_in)
        loss = self.forward_loss(h, dec_out, n_items)
        dh, dW = self.backward(h, dec_in, dec_out, n_items, g_x0)
        return loss, dh, dW

    def capture_step(self, dec_in: Tensor, dec_out: Tensor, n_items: Tensor, h: Tensor, g_x0: Tensor,
                     warmup: int = 2) -> torch.cuda.CUDAGraph:
        """Record :meth:`step` on these (static) input buffers into a CUDA graph.  The step is a
        chain of ~17 short kernels (4-90 us each); replaying it as one graph removes the per-launch
        host cost and the inter-kernel gaps.  Every entry point on the path is capture-safe: no
        allocation, no synchronisation, the live-index count of the table backward stays on the
        device.  Refill the input buffers in place, then ``graph.replay()``; results land in
        ``sbuf['aux'][0]`` (loss), ``dh`` and ``dW``."""
        if not self.dense_grad:
            # the row-store epoch is a host-side counter passed by value: a replayed graph would stamp
            # every step with the captured epoch and stale rows of earlier replays would look live
            raise RuntimeError("recguru: capture_step needs dense_grad=True (the row-store epoch is a launch argument)")
        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            for _ in range(warmup):
                self.step(dec_in, dec_out, n_items, h, g_x0)
        torch.cuda.current_stream().wait_stream(warm)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.step(dec_in, dec_out, n_items, h, g_x0)
        return graph

    # algorithmic bytes per step, for roofline accounting
    def algorithmic_bytes(self, distinct_rows: Optional[int] = None, live: Optional[int] = None,
                          live_pos: Optional[int] = None) -> dict:
        """Per-stage algorithmic bytes (cache hits never discounted).  For the table backward,
        ``live`` = gathered indices that carry a gradient (id != padding_idx and coefficient != 0);
        dead ones cost their id + coefficient only (they are compacted away before the sort), which
        is the honest numerator on padded batches.  ``live=None`` counts every index as live."""
        n, d, K, V1 = self.B * self.L, self.d, self.K, self.V1
        total = n * (2 + K)
        live = total if live is None else live
        b = {
            "gather_pe_mask": n * (8 + 4 * d + 4 * d + 4),
            # per position: mask + logits + derivatives; per SCORED position: h row, ids, table rows
            "sampled_score_loss_fwd": n * (4 + 2 * 4 * (1 + K))
            + (n if (live_pos is None or not self.skip_masked) else live_pos) * (4 * d + (1 + K) * (8 + 4 * d)),
            # positions whose coefficients are all zero (masked) need no table rows for dh
            "sampled_score_bwd_dh": n * ((1 + K) * (8 + 4) + 4 * d)
            + (n if live_pos is None else live_pos) * (1 + K) * 4 * d,
            # plan: id + liveness/coefficient of every gathered index, dense zero fill of dW
            "embedding_bwd_plan": total * (8 + 4) + 4 * V1 * d,
            # apply: one source row per live index, one dW row per distinct id
            "embedding_bwd_apply": live * 4 * d + (4 * d * distinct_rows if distinct_rows is not None else 0),
        }
        if self.fuse_dh:
            # dh comes out of the scoring kernel: its dense write moves there, the re-reads of ids,
            # derivatives and table rows by the second kernel disappear
            b["sampled_score_loss_fwd"] += n * 4 * d
            b["sampled_score_bwd_dh_standalone"] = b.pop("sampled_score_bwd_dh")
            b["total"] = sum(v for k, v in b.items() if k != "sampled_score_bwd_dh_standalone")
            return b
        b["total"] = sum(b.values())
        return b
