"""The training-step hot path as one object: all lookup sites of a step, fused.

A reference training step touches the item table at its two edges
(SURVEY.md section 3.5): ids -> (W[ids] + pe) * mask -> [transformer, not ours] -> h ->
(h . W[pos], h . W[neg]) -> loss, and back.  :class:`TrainStep` runs exactly those edges with
pre-allocated buffers and no per-step allocation, and produces ONE dense table gradient for all
lookup sites (inputs + positives + negatives) in a single deterministic sort/segmented-reduce
pass, instead of one dense [V1, d] gradient per site summed by autograd.

The three step topologies of the reference are instances of it:

=====================  ===========================  =============================  ==========
reference step         input sites (ids -> mask)     scoring edge                   loss mask
=====================  ===========================  =============================  ==========
SASRec BPR             dec_in -> dec_in != pad       pos = dec_out, K = 5 negs,     dec_in
 AutoEnc4Rec.py:87-131                               BPRLoss_sas
AutoRec recon / BPR    enc_in -> enc_in != pad       pos = dec_out; K = n_negs      dec_in
 AutoEnc4Rec.py:175-230  dec_in -> dec_in != pad     (30) sampled softmax, or K = 5
 :55-133                                             BPRLoss_sas
RecGURU, per domain    enc_in -> dec_out != pad      pos = dec_out; K = n_negs      dec_out
 AutoEnc4Rec_cross.py    (the mask argument,         sampled softmax (recon) or
 :93-221;                gan_training.py:347-350)    K = n_bpr_neg BPRLoss
 tools/utils.py:60-127 dec_in -> enc_in != pad (:118)
                         or dec_in != pad (:150)
=====================  ===========================  =============================  ==========

A *mask source* is the id tensor whose ``!= pad_index`` pattern masks a site; sites that share a
source share one mask (and, when ALL sites share it, the table backward compacts by row).

It is driven by whatever sits between the edges (the reference's EncoderM/DecoderM): call
:meth:`forward_in` per input site, run the transformer on the returned activations, call
:meth:`forward_loss` on its output, then :meth:`backward` with the gradients that came back to the
input activations.  :class:`TrainHotPath` is the single-input (SASRec) form with the argument
order the round-1 callers use; :class:`RecGuruStep` runs the two domains of a RecGURU step on two
streams.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch

from . import ops
from .ops import Site

Tensor = torch.Tensor


class TrainStep:
    """Edges of one training step over ONE item table with ``n_inputs`` input lookup sites."""

    def __init__(self, weight: Tensor, pe: Tensor, B: int, L: int, K: int, kind: str = "bpr_sas",
                 n_inputs: int = 1, pad_index: int = 0, padding_idx: Optional[int] = 0,
                 skip_masked: bool = True, overlap_plan: bool = True, dense_grad: bool = True,
                 fuse_dh: Optional[bool] = None):
        assert weight.is_cuda and weight.dtype == torch.float32
        if not 1 <= n_inputs <= ops._lib.MAX_SITES - 2:
            raise ValueError(f"recguru: n_inputs must be in [1, {ops._lib.MAX_SITES - 2}]")
        self.weight = weight
        self.pe = pe[0] if pe.dim() == 3 else pe
        self.B, self.L, self.K, self.kind = B, L, K, kind
        self.n_inputs = n_inputs
        self.pad_index = pad_index
        self.skip_masked = skip_masked
        self.padding_idx = -1 if padding_idx is None else padding_idx
        V1, d = weight.shape
        self.V1, self.d = V1, d
        dev = weight.device
        self.device = dev
        n = B * L
        f32 = dict(dtype=torch.float32, device=dev)
        self.x = [torch.empty(B, L, d, **f32) for _ in range(n_inputs)]         # activations per input site
        self._in_mask_buf = [torch.empty(B, L, **f32) for _ in range(n_inputs)]
        self.in_mask = list(self._in_mask_buf)                                   # mask each site was built with
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
        total = n * (n_inputs + 1 + K)
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
        # RECGURU_ZERO_AFTER=score: start the fill behind the scoring kernel instead of behind K1 (A/B knob)
        self._zero_late = os.environ.get("RECGURU_ZERO_AFTER", "k1") == "score"
        self._zero_done = torch.cuda.Event()
        self._fork = torch.cuda.Event()
        self._join = torch.cuda.Event()
        self._zero_issued = False
        self._k1_done = torch.cuda.Event()
        # up to n_inputs + 1 distinct mask sources per step; masks derived on the side stream (plan)
        # and on the main stream (forward) live in separate buffers: no cross-stream hazard
        self._side_masks = [torch.empty(B, L, **f32) for _ in range(n_inputs + 1)]
        self._main_masks = [torch.empty(B, L, **f32) for _ in range(n_inputs + 1)]
        # fuse_dh: d loss / d h comes out of the padding-aware scoring kernel itself
        # (recguru_sampled_score_loss_fwd_dh) instead of a second kernel that re-reads ids,
        # derivatives and table rows.  It needs sum(mask) before the loss is reduced: the pad-mask
        # kernel that runs at the start of the step anyway also counts (recguru_pad_mask_count).
        if fuse_dh is None:                                # RECGURU_FUSE_DH=0: two-kernel path (A/B timing)
            fuse_dh = skip_masked and os.environ.get("RECGURU_FUSE_DH", "1") != "0"
        self.fuse_dh = bool(fuse_dh)
        if self.fuse_dh and not skip_masked:
            raise RuntimeError("recguru: fuse_dh needs the padding-aware scoring kernel (skip_masked=True)")
        self._cnt = torch.zeros(3, dtype=torch.int32, device=dev)
        self._cnt_ready = torch.cuda.Event()
        # score_rows: when every lookup site of the step shares one pad mask (SASRec), the plan's row
        # compaction yields the list of live positions anyway: the scoring kernel walks that list, one warp
        # per live position, instead of compacting the mask per CTA (RECGURU_SCORE_ROWS=0: A/B timing)
        self.score_rows = self.fuse_dh and os.environ.get("RECGURU_SCORE_ROWS", "1") != "0"
        self._rows_ready = torch.cuda.Event()
        # (steps whose sites do not share one mask: the list of the LOSS mask from recguru_live_rows)
        self._rows_buf = torch.empty(n, dtype=torch.int32, device=dev)
        self._nrows_buf = torch.zeros(1, dtype=torch.int32, device=dev)
        self._rows_ws = torch.empty(max(256, int(ops._lib.load().recguru_live_rows_workspace_bytes(n))),
                                    dtype=torch.uint8, device=dev)
        self.begin_step()

    # -- step timeline (debug) -----------------------------------------------------------
    TRACE_POINTS = ("start", "pad_mask", "plan_done", "dw_zero_done", "k1_done", "score_done", "apply_done",
                    "rows_ready")

    def enable_trace(self) -> Tensor:
        """Timestamp kernels (recguru_debug_timestamp) at the ends of the step's branches: they sit in the
        captured graph too, so a replay yields the real multi-stream timeline (scripts/step_timeline.py)."""
        self._trace = torch.zeros(len(self.TRACE_POINTS), dtype=torch.int64, device=self.device)
        return self._trace

    def _stamp(self, name: str) -> None:
        if getattr(self, "_trace", None) is not None:
            ops.raw_debug_timestamp(self._trace, self.TRACE_POINTS.index(name))

    # -- per-step state -------------------------------------------------------------------
    def begin_step(self) -> None:
        """Forget everything a previous (possibly unfinished: forward-only, failed, skipped) step
        left behind.  Called by :meth:`step`; the individual edges call it themselves when they see
        that a new step has begun (an input site produced twice, or an edge entered after the loss
        of the previous step), so a validation forward without a backward cannot leak its
        live-count or its plan into the next step."""
        self._planned = False
        self._zero_issued = False
        self._cnt_state = 0            # 0 = not issued this step, 1 = on the current stream, 2 = on the side stream
        self._dh_done = False
        self._dh_scale = None
        self._loss_done = False
        self._in_done = set()
        self._in_ids: List[Optional[Tensor]] = [None] * self.n_inputs
        self._in_src: List[Optional[Tensor]] = [None] * self.n_inputs
        self._main_src: List[Tensor] = []      # mask sources whose mask exists on the main stream (index = buffer)
        self._main_ext = {}                    # id(source) -> mask tensor produced by K1 itself
        self._loss_mask = None
        self._live_rows = None                 # (rows, n_rows) device addresses once the plan compacted by row
        self._rows_src = None
        self._compacted = None                 # (pos, neg) when forward_loss ran the compaction itself

    # -- mask plumbing --------------------------------------------------------------------
    def _mask_on_main(self, src: Tensor) -> Tensor:
        """The fp32 0/1 mask of ``src != pad_index`` on the current stream (computed once per step)."""
        m = self._main_ext.get(id(src))
        if m is not None:
            return m
        for j, s in enumerate(self._main_src):
            if s is src:
                return self._main_masks[j]
        j = len(self._main_src)
        self._main_src.append(src)
        return ops.raw_pad_mask(src, self.pad_index, out=self._main_masks[j])

    # -- edge 1: ids -> activations ------------------------------------------------------
    def forward_in(self, i: int, ids: Tensor, mask_ids: Optional[Tensor] = None) -> Tensor:
        """``x_i = (W[ids] + pe) * (mask_ids != pad_index)``; ``mask_ids`` defaults to ``ids``."""
        if i in self._in_done or self._loss_done:
            self.begin_step()
        src = ids if mask_ids is None else mask_ids
        if src is ids:
            ops.raw_gather_pe_mask(self.weight, ids, self.pe, self.pad_index, out=self.x[i],
                                   mask_out=self._in_mask_buf[i])
            # sites with the same mask source share ONE mask tensor (the table backward compacts by
            # row when every site does)
            self.in_mask[i] = self._main_ext.setdefault(id(ids), self._in_mask_buf[i])
        else:
            m = self._mask_on_main(src)
            ops.raw_gather_pe_mask(self.weight, ids, self.pe, -1, mask_in=m, out=self.x[i], want_mask=False)
            self.in_mask[i] = m
        self._in_done.add(i)
        self._in_ids[i], self._in_src[i] = ids, src
        if i == self.n_inputs - 1:
            self._stamp("k1_done")
        return self.x[i]

    # -- edge 2: hidden states -> loss ---------------------------------------------------
    def forward_loss(self, h: Tensor, pos: Tensor, neg: Tensor, mask_ids: Optional[Tensor] = None,
                     mask: Optional[Tensor] = None, grad_loss: Optional[Tensor] = None) -> Tensor:
        """Sampled scores + masked-mean loss.  The loss mask is ``mask_ids != pad_index``
        (``mask_ids`` defaults to the first input site's mask source) unless an explicit fp32
        ``mask`` is given (then the live count is unknown up front and dh takes its own kernel)."""
        if self._loss_done:
            self.begin_step()
        explicit = mask is not None
        if not explicit:
            src = mask_ids if mask_ids is not None else self._in_src[0]
            if src is None:
                raise RuntimeError("recguru: forward_loss needs mask_ids (no input site ran this step)")
            if self.fuse_dh and self._cnt_state == 0:      # no plan forked: count on this stream
                j = len(self._main_src)
                self._main_src.append(src)
                ops.raw_pad_mask_count(src, self.pad_index, self._main_masks[j], self._cnt)
                self._cnt_state = 1
            mask = self._mask_on_main(src)
        self._loss_mask = mask
        # the loss is all this edge exports: positions under the pad mask are not scored
        if self.fuse_dh and not explicit and self._cnt_state != 0:
            cur = torch.cuda.current_stream()
            if self._cnt_state == 2:
                cur.wait_event(self._cnt_ready)
            self._issue_zero_fill(cur)
            rows = None
            if self._live_rows is not None and src is self._rows_src:
                cur.wait_event(self._rows_ready)           # the plan's compaction: the live-position list
                rows = self._live_rows
            elif self.score_rows and not self._planned and self._cnt_state == 1 and \
                    len(self._in_done) == self.n_inputs and all(s_ is src for s_ in self._in_src):
                # no plan forked (serial schedule): the compaction the table backward needs anyway runs
                # here, ahead of the scoring kernel, so that the same kernel scores in every schedule
                n_ = self.B * self.L
                m_ = mask.view(n_)
                sites = [Site(self._in_ids[i], None, live=m_) for i in range(self.n_inputs)]
                sites += [Site(pos, None, live=m_), Site(neg, None, group=self.K, live=m_)]
                ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd, stage=ops.PLAN_COMPACT)
                rows = ops.raw_embedding_bwd_plan_live_rows(sites, self.V1, self.ws_bwd)
                self._compacted = (pos, neg)
            elif self.score_rows and not self._planned and self._cnt_state == 1:
                # serial schedule, several mask sources: the loss mask's live rows on this stream
                ops.raw_live_rows(mask.view(-1), self._rows_buf, self._nrows_buf, self._rows_ws)
                rows = (self._rows_buf.data_ptr(), self._nrows_buf.data_ptr())
            ops.raw_sampled_score_loss(h, self.weight, pos, neg, mask, self.kind, bufs=self.sbuf,
                                       skip_masked=True, dh_out=self.dh, live_count=self._cnt,
                                       grad_scale=grad_loss, live_rows=rows)
            self._dh_done = True
            self._dh_scale = grad_loss
        else:
            self._issue_zero_fill(torch.cuda.current_stream())
            ops.raw_sampled_score_loss(h, self.weight, pos, neg, mask, self.kind, bufs=self.sbuf,
                                       skip_masked=self.skip_masked)
        if self._zero_late:
            self._issue_zero_fill(torch.cuda.current_stream(), late=True)
        self._stamp("score_done")
        self._loss_done = True
        return self.sbuf["aux"][0]

    def _issue_zero_fill(self, cur, late: bool = False) -> None:
        """The dense zero fill of dW (dense-gradient contract) of a planned step, on its own stream,
        BEHIND the input gathers (bandwidth-bound like them: side by side they stretched the gathers
        from 29 to 43 us -- scripts/step_timeline.py) and BESIDE the scoring kernel (latency-bound: it
        leaves the HBM pins idle)."""
        if not (self._planned and self.dense_grad and self._zero_own and not self._zero_issued):
            return
        if self._zero_late != late:
            return
        self._k1_done.record(cur)
        self._zstream.wait_event(self._k1_done)
        with torch.cuda.stream(self._zstream):
            ops.raw_fill_zero(self.dW)
            self._stamp("dw_zero_done")
            self._zero_done.record(self._zstream)
        self._zero_issued = True

    # -- table-backward plan: depends on the ids and the pad masks only -----------------------
    def plan_backward(self, in_ids: Sequence[Tensor], pos: Tensor, neg: Tensor,
                      in_mask_ids: Optional[Sequence[Optional[Tensor]]] = None,
                      loss_mask_ids: Optional[Tensor] = None, mask: Optional[Tensor] = None) -> None:
        """Issue stage A of the table backward (zero fill of dW, compaction of the live lookups,
        stable radix sort) on a side stream.  It needs nothing but the ids and the pad masks, so it
        runs UNDER the forward pass (in a full model: under the whole transformer forward/backward)
        instead of on the critical path; :meth:`backward` joins it.  Must precede
        :meth:`forward_loss` of its step.  ``mask``: one explicit fp32 mask for every site instead
        of the id-derived ones."""
        if self._planned or self._loss_done:
            self.begin_step()
        if len(in_ids) != self.n_inputs:
            raise ValueError(f"recguru: expected {self.n_inputs} input id tensors, got {len(in_ids)}")
        n = self.B * self.L
        srcs = [in_ids[i] if (in_mask_ids is None or in_mask_ids[i] is None) else in_mask_ids[i]
                for i in range(self.n_inputs)]
        loss_src = srcs[0] if loss_mask_ids is None else loss_mask_ids
        cur = torch.cuda.current_stream()
        self._stamp("start")
        self._fork.record(cur)
        self._side.wait_event(self._fork)
        # (the dense zero fill of dW is bandwidth-bound like the input gathers: it is issued by
        # forward_loss, behind the gathers and beside the latency-bound scoring kernel)
        self._zero_issued = False
        with torch.cuda.stream(self._side):
            fused_src = None
            if mask is None:
                # the side stream derives the pad masks itself, so the plan does not have to wait for K1
                # and can be issued at the very start of the step.  One shared mask source (SASRec): the
                # plan's first kernel derives mask and count (recguru_embedding_bwd_plan_padmask);
                # otherwise one 2 us kernel per distinct source.
                distinct: List[Tensor] = []
                for s in [loss_src] + srcs:
                    if not any(s is t for t in distinct):
                        distinct.append(s)
                side = {}
                need_cnt = self.fuse_dh and self._cnt_state == 0
                if len(distinct) == 1 and not need_cnt and os.environ.get("RECGURU_BWD_PLAN", "1") == "2":
                    # nobody waits for the population count: the plan's first kernel derives the mask
                    fused_src = distinct[0]
                    side[id(fused_src)] = self._side_masks[0].view(n)
                else:
                    # (the fused scoring kernel needs sum(mask) early: a 2 us kernel of its own, so that
                    # the event it waits for sits at the head of the plan branch, not at its end)
                    for j, s in enumerate(distinct):
                        if s is loss_src and self.fuse_dh and self._cnt_state == 0:
                            # (state 1: the forward already counted on its stream)
                            side[id(s)] = ops.raw_pad_mask_count(s, self.pad_index, self._side_masks[j], self._cnt).view(n)
                            self._cnt_ready.record(self._side)
                            self._cnt_state = 2
                        else:
                            side[id(s)] = ops.raw_pad_mask(s, self.pad_index, out=self._side_masks[j]).view(n)
                in_live = [side[id(s)] for s in srcs]
                loss_live = side[id(loss_src)]
                self._stamp("pad_mask")
            else:
                in_live = [mask.view(n)] * self.n_inputs
                loss_live = mask.view(n)
            if self.dense_grad and not self._zero_own:
                ops.raw_fill_zero(self.dW)
            sites = [Site(in_ids[i], None, live=in_live[i]) for i in range(self.n_inputs)]
            sites += [Site(pos, None, live=loss_live), Site(neg, None, group=self.K, live=loss_live)]
            one_mask = all(t is loss_live for t in in_live)
            if fused_src is not None:
                ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd, mask_ids=fused_src,
                                           pad_index=self.pad_index)
            elif self.score_rows and one_mask and mask is None and self._cnt_state != 0:
                # compaction | event | sort: the scoring kernel of the main stream consumes the live rows
                ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd, stage=ops.PLAN_COMPACT)
                self._stamp("rows_ready")
                self._rows_ready.record(self._side)
                self._live_rows = ops.raw_embedding_bwd_plan_live_rows(sites, self.V1, self.ws_bwd)
                self._rows_src = loss_src
                ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd, stage=ops.PLAN_SORT)
            else:
                if self.score_rows and mask is None and self._cnt_state != 0:
                    # several mask sources: the live rows of the LOSS mask from their own three small kernels
                    ops.raw_live_rows(loss_live, self._rows_buf, self._nrows_buf, self._rows_ws)
                    self._rows_ready.record(self._side)
                    self._live_rows = (self._rows_buf.data_ptr(), self._nrows_buf.data_ptr())
                    self._rows_src = loss_src
                ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd)
            if not self.dense_grad:
                self.epoch += 1
                ops.raw_embedding_bwd_mark_rows(sites, self.V1, self.ws_bwd, self.row_epoch, self.epoch)
            self._stamp("plan_done")
            self._join.record(self._side)
        self._planned = True

    # -- backward of both edges ------------------------------------------------------------
    def backward(self, h: Tensor, pos: Tensor, neg: Tensor, g_x: Sequence[Tensor],
                 grad_loss: Optional[Tensor] = None):
        """Returns (dh, dW).  ``g_x[i]`` is d loss / d x_i as it comes back through the transformer.
        The ids and masks of the input sites are the ones :meth:`forward_in` saw this step."""
        if len(g_x) != self.n_inputs or len(self._in_done) != self.n_inputs or self._loss_mask is None:
            raise RuntimeError("recguru: backward needs every forward_in and forward_loss of this step first")
        go = self.one if grad_loss is None else grad_loss
        inv_m = self.sbuf["aux"][1:2]
        if self._dh_done and (grad_loss is None) == (self._dh_scale is None) and \
                (grad_loss is None or grad_loss.data_ptr() == self._dh_scale.data_ptr()):
            pass                                           # dh was written by the forward kernel
        else:
            ops.raw_sampled_score_bwd_dh(self.sbuf["dp"], self.sbuf["dn"], self.weight, pos, neg,
                                         scale=go, scale2=inv_m, out=self.dh)
        n = self.B * self.L
        h2 = h.view(n, self.d)
        lm = self._loss_mask.view(n)
        sites = []
        for i in range(self.n_inputs):                                                          # input sites
            m = self.in_mask[i].view(n)
            sites.append(Site(self._in_ids[i], g_x[i].view(n, self.d), coef=m, live=m))
        sites.append(Site(pos, h2, coef=self.sbuf["dp"], scale=go, scale2=inv_m, live=lm))                   # positives
        sites.append(Site(neg, h2, coef=self.sbuf["dn"].view(-1), scale=go, scale2=inv_m, group=self.K, live=lm))  # negatives
        if self._planned:
            torch.cuda.current_stream().wait_event(self._join)
            if self.dense_grad and self._zero_own:
                if not self._zero_issued:                  # (forward_loss of this step did not run planned)
                    ops.raw_fill_zero(self.dW)
                else:
                    torch.cuda.current_stream().wait_event(self._zero_done)
            ops.raw_embedding_bwd_apply(sites, self.V1, self.d, self.dW, self.ws_bwd, mode=ops.DW_PREZEROED)
        elif self.dense_grad:
            if self._compacted is not None and self._compacted[0] is pos and self._compacted[1] is neg:
                # forward_loss compacted these sites already: sort + reduce
                ops.raw_fill_zero(self.dW)
                ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd, stage=ops.PLAN_SORT)
                ops.raw_embedding_bwd_apply(sites, self.V1, self.d, self.dW, self.ws_bwd, mode=ops.DW_PREZEROED)
            else:
                ops.raw_embedding_bwd(sites, self.V1, self.d, self.padding_idx, out=self.dW, zero_fill=True,
                                      workspace=self.ws_bwd)
        elif self._compacted is not None and self._compacted[0] is pos and self._compacted[1] is neg:
            ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd, stage=ops.PLAN_SORT)
            ops.raw_embedding_bwd_apply(sites, self.V1, self.d, self.dW, self.ws_bwd, mode=ops.DW_PREZEROED)
            self.epoch += 1
            ops.raw_embedding_bwd_mark_rows(sites, self.V1, self.ws_bwd, self.row_epoch, self.epoch)
        else:
            ops.raw_embedding_bwd_plan(sites, self.V1, self.padding_idx, self.ws_bwd)
            ops.raw_embedding_bwd_apply(sites, self.V1, self.d, self.dW, self.ws_bwd, mode=ops.DW_PREZEROED)
            self.epoch += 1
            ops.raw_embedding_bwd_mark_rows(sites, self.V1, self.ws_bwd, self.row_epoch, self.epoch)
        self._stamp("apply_done")
        self.begin_step()
        return self.dh, self.dW

    def run(self, in_ids: Sequence[Tensor], pos: Tensor, neg: Tensor, h: Tensor, g_x: Sequence[Tensor],
            in_mask_ids: Optional[Sequence[Optional[Tensor]]] = None, loss_mask_ids: Optional[Tensor] = None):
        """All edges of one step with the transformer replaced by given (h, g_x)."""
        self.begin_step()
        if self.overlap_plan:
            self.plan_backward(in_ids, pos, neg, in_mask_ids, loss_mask_ids)     # needs the ids only: forked first
        for i in range(self.n_inputs):
            self.forward_in(i, in_ids[i], None if in_mask_ids is None else in_mask_ids[i])
        loss = self.forward_loss(h, pos, neg, mask_ids=loss_mask_ids)
        dh, dW = self.backward(h, pos, neg, g_x)
        return loss, dh, dW

    def capture(self, fn, warmup: int = 2) -> torch.cuda.CUDAGraph:
        """Record ``fn()`` (a closure over static input buffers that drives this object's edges)
        into a CUDA graph.  Every entry point on the path is capture-safe: no allocation, no
        synchronisation, the live-index count of the table backward stays on the device."""
        if not self.dense_grad:
            # the row-store epoch is a host-side counter passed by value: a replayed graph would stamp
            # every step with the captured epoch and stale rows of earlier replays would look live
            raise RuntimeError("recguru: graph capture needs dense_grad=True (the row-store epoch is a launch argument)")
        warm = torch.cuda.Stream(device=self.device)
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(warm)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        return graph

    def capture_run(self, in_ids, pos, neg, h, g_x, in_mask_ids=None, loss_mask_ids=None, warmup: int = 2):
        """:meth:`run` on these (static) buffers as one CUDA graph: refill the buffers in place, then
        ``graph.replay()``; results land in ``sbuf['aux'][0]`` (loss), ``dh`` and ``dW``."""
        return self.capture(lambda: self.run(in_ids, pos, neg, h, g_x, in_mask_ids, loss_mask_ids), warmup)

    # algorithmic bytes per step, for roofline accounting
    def algorithmic_bytes(self, distinct_rows: Optional[int] = None, live: Optional[int] = None,
                          live_pos: Optional[int] = None) -> dict:
        """Per-stage algorithmic bytes (cache hits never discounted).  For the table backward,
        ``live`` = gathered indices that carry a gradient (id != padding_idx and coefficient != 0);
        dead ones cost their id + coefficient only (they are compacted away before the sort), which
        is the honest numerator on padded batches.  ``live=None`` counts every index as live."""
        n, d, K, V1 = self.B * self.L, self.d, self.K, self.V1
        total = n * (self.n_inputs + 1 + K)
        live = total if live is None else live
        b = {
            "gather_pe_mask": self.n_inputs * n * (8 + 4 * d + 4 * d + 4),
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


class TrainHotPath(TrainStep):
    """SASRec BPR step edges (MyRec.forward with sas=True, AutoEnc4Rec.py:87-131 + BPRLoss_sas):
    one input site ``dec_in`` that is also the mask source of the loss."""

    def __init__(self, weight: Tensor, pe: Tensor, B: int, L: int, K: int, kind: str = "bpr_sas",
                 pad_index: int = 0, padding_idx: Optional[int] = 0, skip_masked: bool = True,
                 overlap_plan: bool = True, dense_grad: bool = True, fuse_dh: Optional[bool] = None):
        super().__init__(weight, pe, B, L, K, kind, n_inputs=1, pad_index=pad_index, padding_idx=padding_idx,
                         skip_masked=skip_masked, overlap_plan=overlap_plan, dense_grad=dense_grad,
                         fuse_dh=fuse_dh)

    @property
    def x0(self) -> Tensor:
        return self.x[0]

    @property
    def mask(self) -> Tensor:
        return self.in_mask[0]

    @property
    def _side_mask(self) -> Tensor:
        return self._side_masks[0]

    def forward_in(self, dec_in: Tensor) -> Tensor:          # noqa: D102 (single-site signature)
        if 0 in self._in_done or self._loss_done:
            self.begin_step()
        x = super().forward_in(0, dec_in)
        if self.fuse_dh and self._cnt_state == 0:          # no plan forked: count on this stream
            ops.raw_pad_mask_count(dec_in, self.pad_index, self._main_masks[0], self._cnt)
            self._cnt_state = 1
        return x

    def forward_loss(self, h: Tensor, dec_out: Tensor, n_items: Tensor, mask: Optional[Tensor] = None,
                     grad_loss: Optional[Tensor] = None) -> Tensor:
        return super().forward_loss(h, dec_out, n_items, mask=mask, grad_loss=grad_loss)

    def plan_backward(self, dec_in: Tensor, dec_out: Tensor, n_items: Tensor, mask: Optional[Tensor] = None):
        return super().plan_backward([dec_in], dec_out, n_items, mask=mask)

    def backward(self, h: Tensor, dec_in: Tensor, dec_out: Tensor, n_items: Tensor, g_x0: Tensor,
                 grad_loss: Optional[Tensor] = None):
        return super().backward(h, dec_out, n_items, [g_x0], grad_loss=grad_loss)

    def step(self, dec_in: Tensor, dec_out: Tensor, n_items: Tensor, h: Tensor, g_x0: Tensor):
        """All edges of one step with the transformer replaced by given (h, g_x0)."""
        self.begin_step()
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
        host cost and the inter-kernel gaps.  Refill the input buffers in place, then
        ``graph.replay()``; results land in ``sbuf['aux'][0]`` (loss), ``dh`` and ``dW``."""
        return self.capture(lambda: self.step(dec_in, dec_out, n_items, h, g_x0), warmup)


def autorec_step(weight: Tensor, pe: Tensor, B: int, L: int, K: int, kind: str, **kw) -> TrainStep:
    """AutoRec (MyAuto4Rec / MyRec non-sas, AutoEnc4Rec.py:55-133,:175-230): input sites
    ``enc_in`` and ``dec_in`` each masked by its own pad pattern; reconstruction phase
    ``kind='sce'`` with K = n_negs (30), fine-tune phase ``kind='bpr_sas'`` with K = 5; the loss
    mask is ``dec_in != pad`` (train_auto.py:108-110).  Drive it with
    ``run([enc_in, dec_in], dec_out, n_items, h, [g_enc, g_dec], loss_mask_ids=dec_in)``."""
    return TrainStep(weight, pe, B, L, K, kind, n_inputs=2, **kw)


class RecGuruStep:
    """One RecGURU reconstruction / BPR step (train_recon_x, gan_training.py:836-861): the same
    edges on the two domains' tables (AutoEnc4Rec_cross.py:93-221), no ``padding_idx``
    (:32,:35).  Per domain: ``enc_in`` masked by the caller's mask = ``dec_out != pad``
    (gan_training.py:347-350,:843-844 -> :98-102), ``dec_in`` masked by ``enc_in != pad``
    (get_dec_out, :118-128) or by ``dec_in != pad`` (recommend_forward, :150-160), loss masked by
    ``dec_out != pad``.  The domains share no table, so they run on two streams."""

    def __init__(self, weight_a: Tensor, weight_b: Tensor, pe: Tensor, B: int, L: int, K: int, kind: str = "sce",
                 recommend: bool = False, **kw):
        kw.setdefault("padding_idx", None)
        self.dom = [TrainStep(weight_a, pe, B, L, K, kind, n_inputs=2, **kw),
                    TrainStep(weight_b, pe, B, L, K, kind, n_inputs=2, **kw)]
        self.recommend = recommend
        dev = weight_a.device
        self._sb = torch.cuda.Stream(device=dev)
        self._fork, self._join = torch.cuda.Event(), torch.cuda.Event()

    def _run_domain(self, st: TrainStep, b, h, g):
        dec_src = b.dec_in if self.recommend else b.enc_in
        return st.run([b.enc_in, b.dec_in], b.dec_out, b.n_items, h, g,
                      in_mask_ids=[b.dec_out, dec_src], loss_mask_ids=b.dec_out)

    def run(self, batch_a, batch_b, h_a: Tensor, h_b: Tensor, g_a: Sequence[Tensor], g_b: Sequence[Tensor]):
        """batches: objects with enc_in/dec_in/dec_out/n_items.  Returns ((loss, dh, dW) per domain)."""
        cur = torch.cuda.current_stream()
        self._fork.record(cur)
        self._sb.wait_event(self._fork)
        with torch.cuda.stream(self._sb):
            rb = self._run_domain(self.dom[1], batch_b, h_b, g_b)
            self._join.record(self._sb)
        ra = self._run_domain(self.dom[0], batch_a, h_a, g_a)
        cur.wait_event(self._join)
        return ra, rb

    def capture_run(self, batch_a, batch_b, h_a, h_b, g_a, g_b, warmup: int = 2) -> torch.cuda.CUDAGraph:
        return self.dom[0].capture(lambda: self.run(batch_a, batch_b, h_a, h_b, g_a, g_b), warmup)
