"""torch.library custom ops (namespace ``recguru``) over the C ABI in include/recguru_b200.h.

Host-side mirror of the reference's PyTorch surface for the embedding-and-scoring hot path
(SURVEY.md section 8b).  Every op is CUDA-only: there is no CPU implementation, no Triton,
no multi-backend dispatch.  PyTorch is used for device memory, streams and autograd glue.

Raw layer   : ``raw_*`` functions -- validate, allocate outputs/scratch with torch's caching
              allocator, pass raw device pointers to the shared library.
Op layer    : ``torch.ops.recguru.*`` with register_fake (shape inference) and
              register_autograd where the reference op is differentiable.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import weakref
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import LOSS_KINDS, LookupSite, check

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: Tensor, dtype: torch.dtype, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"recguru: {name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise RuntimeError(f"recguru: {name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _workspace(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------
# raw layer
# --------------------------------------------------------------------------------------
def raw_gather_pe_mask(weight: Tensor, ids: Tensor, pe: Optional[Tensor], pad_index: int,
                       mask_in: Optional[Tensor] = None, out: Optional[Tensor] = None,
                       mask_out: Optional[Tensor] = None, want_mask: bool = True
                       ) -> Tuple[Tensor, Optional[Tensor]]:
    """out[..., :] = (weight[ids] + pe[l]) * mask ; mask = (ids != pad_index) unless mask_in.

    Reference: AutoEnc4Rec.py:71-75,:89-93,:176-181 ; Transformer/transformer.py:104-106.
    ``pe`` may be ``[>=L, d]`` or the reference's ``[1, max_len, d]`` buffer; ``None`` gives a
    plain gather (S1/S2/E1 sites)."""
    lib = _lib.load()
    weight = _need(weight, torch.float32, "weight")
    ids = _need(ids, torch.int64, "ids")
    V1, d = weight.shape
    L = ids.shape[-1] if ids.dim() >= 1 else 1
    n_tokens = ids.numel()
    if pe is not None:
        pe = _need(pe, torch.float32, "pe")
        if pe.dim() == 3:
            pe = pe[0]
        if pe.shape[0] < L or pe.shape[1] != d:
            raise RuntimeError(f"recguru: pe shape {tuple(pe.shape)} incompatible with L={L}, d={d}")
    if mask_in is not None:
        mask_in = _need(mask_in, torch.float32, "mask")
        if mask_in.numel() != n_tokens:
            raise RuntimeError("recguru: mask must have one entry per token")
    if out is None:
        out = torch.empty(*ids.shape, d, dtype=torch.float32, device=weight.device)
    if mask_out is None and want_mask and pe is not None:
        mask_out = torch.empty(ids.shape, dtype=torch.float32, device=weight.device)
    check(lib.recguru_gather_pe_mask(_ptr(weight), V1, d, _ptr(ids), n_tokens, max(L, 1), _ptr(pe),
                                     _ptr(mask_in), int(pad_index), _ptr(out), _ptr(mask_out),
                                     _stream()), "gather_pe_mask")
    return out, mask_out


def torch_dropout_state(numel: int, device) -> Tuple[int, int, int]:
    """Consume torch's CUDA generator exactly as ``torch.nn.functional.dropout`` would for a fp32
    tensor of ``numel`` elements (ATen/native/cuda/Dropout.cu: dropout_cuda): returns
    ``(seed, philox_offset, torch_threads)`` and advances the generator's offset by
    ``((numel - 1) // (torch_threads * 4) + 1) * 4``.  Host-side bookkeeping: not for use inside a
    CUDA-graph capture."""
    torch.cuda.init()                                   # default_generators is empty before the lazy init
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    prop = torch.cuda.get_device_properties(idx)
    block = 256
    grid = min(prop.multi_processor_count * (prop.max_threads_per_multi_processor // block), (numel + block - 1) // block)
    threads = grid * block
    inc = ((numel - 1) // (threads * 4) + 1) * 4
    gen = torch.cuda.default_generators[idx]
    seed, off = gen.initial_seed(), gen.get_offset()
    gen.set_offset(off + inc)
    return int(seed), int(off), int(threads)


def raw_gather_pe_mask_dropout(weight: Tensor, ids: Tensor, pe: Tensor, pad_index: int, p: float, seed: int,
                               offset: int, threads: int, mask_in: Optional[Tensor] = None,
                               out: Optional[Tensor] = None, mask_out: Optional[Tensor] = None
                               ) -> Tuple[Tensor, Optional[Tensor]]:
    """K1 with the nn.Dropout of PositionalEncoding.forward fused in (Transformer/transformer.py:104-106):
    ``dropout((weight[ids] + pe) * mask, p)``, bit-identical to torch's dropout kernel for the generator
    state ``(seed, offset)`` (see :func:`torch_dropout_state`)."""
    lib = _lib.load()
    weight = _need(weight, torch.float32, "weight")
    ids = _need(ids, torch.int64, "ids")
    pe = _need(pe, torch.float32, "pe")
    if pe.dim() == 3:
        pe = pe[0]
    V1, d = weight.shape
    L = ids.shape[-1]
    if pe.shape[0] < L or pe.shape[1] != d:
        raise RuntimeError(f"recguru: pe shape {tuple(pe.shape)} incompatible with L={L}, d={d}")
    if mask_in is not None:
        mask_in = _need(mask_in, torch.float32, "mask")
    if out is None:
        out = torch.empty(*ids.shape, d, dtype=torch.float32, device=weight.device)
    if mask_out is None and mask_in is None:
        mask_out = torch.empty(ids.shape, dtype=torch.float32, device=weight.device)
    check(lib.recguru_gather_pe_mask_dropout(_ptr(weight), V1, d, _ptr(ids), ids.numel(), max(L, 1), _ptr(pe),
                                             _ptr(mask_in), int(pad_index), float(p), int(seed), int(offset),
                                             int(threads), _ptr(out), _ptr(mask_out), _stream()),
          "gather_pe_mask_dropout")
    return out, mask_out


def raw_dropout(x: Tensor, p: float, seed: int, offset: int, threads: int, row_mask: Optional[Tensor] = None,
                out: Optional[Tensor] = None) -> Tensor:
    """``torch.nn.functional.dropout(x, p, True)`` for the generator state ``(seed, offset)``, bit-identical
    to torch's kernel; rows with ``row_mask == 0`` pass through.  Its own backward (same state)."""
    x = _need(x, torch.float32, "x")
    d = x.shape[-1]
    n_rows = x.numel() // d
    if row_mask is not None:
        row_mask = _need(row_mask, torch.float32, "row_mask")
        if row_mask.numel() != n_rows:
            raise RuntimeError("recguru: row_mask must have one entry per row of x")
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().recguru_dropout(_ptr(x), _ptr(row_mask), n_rows, d, float(p), int(seed), int(offset),
                                      int(threads), _ptr(out), _stream()), "dropout")
    return out


def raw_pe_grad(g: Tensor, mask: Tensor, out: Optional[Tensor] = None, accumulate: bool = False) -> Tensor:
    """dP[l,:] = sum_b mask[b,l] * g[b,l,:]: gradient of a learned position table (P1')."""
    g = _need(g, torch.float32, "g")
    mask = _need(mask, torch.float32, "mask")
    B, L, d = g.shape
    if mask.numel() != B * L:
        raise RuntimeError("recguru: mask must be [B, L]")
    if out is None:
        out = torch.empty(L, d, dtype=torch.float32, device=g.device)
        accumulate = False
    check(_lib.load().recguru_pe_grad(_ptr(g), _ptr(mask), B, L, d, 1 if accumulate else 0, _ptr(out), _stream()),
          "pe_grad")
    return out


def raw_extract_last(h: Tensor, want_f32: bool = True, want_bf16: bool = False):
    """u = h[:, -1, :] (dec_embedding[:, -1, :], train_auto.py:178; get_user_embed, gan_training.py:152-162)
    as fp32 and/or bf16 (the catalog sweep's operand) in one pass.  Returns (u_f32 | None, u_bf16 | None)."""
    h = _need(h, torch.float32, "h")
    B, L, d = h.shape
    u32 = torch.empty(B, d, dtype=torch.float32, device=h.device) if want_f32 else None
    u16 = torch.empty(B, d, dtype=torch.bfloat16, device=h.device) if want_bf16 else None
    check(_lib.load().recguru_extract_last(_ptr(h), B, L, d, _ptr(u32), _ptr(u16), _stream()), "extract_last")
    return u32, u16


def raw_pe_mask(x: Tensor, pe: Tensor, mask: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """PositionalEncoding.forward before dropout (Transformer/transformer.py:104-106)."""
    lib = _lib.load()
    x = _need(x, torch.float32, "x")
    pe = _need(pe, torch.float32, "pe")
    mask = _need(mask, torch.float32, "mask")
    if pe.dim() == 3:
        pe = pe[0]
    L, d = x.shape[-2], x.shape[-1]
    n_tokens = x.numel() // d
    if mask.numel() != n_tokens or pe.shape[0] < L or pe.shape[1] != d:
        raise RuntimeError("recguru: pe_mask shape mismatch")
    if out is None:
        out = torch.empty_like(x)
    check(lib.recguru_pe_mask(_ptr(x), _ptr(pe), _ptr(mask), n_tokens, L, d, _ptr(out), _stream()),
          "pe_mask")
    return out


class Site:
    """One lookup site of the table backward: dW[ids[j]] += coef[j]*scale*scale2 * X[j//group].
    ``live`` (optional, one entry per row of X, e.g. the pad mask) decides which occurrences
    carry gradient when the plan is made before the coefficients exist."""
    __slots__ = ("ids", "X", "coef", "scale", "scale2", "group", "live")

    def __init__(self, ids: Tensor, X: Optional[Tensor], coef: Optional[Tensor] = None,
                 scale: Optional[Tensor] = None, scale2: Optional[Tensor] = None, group: int = 1,
                 live: Optional[Tensor] = None):
        self.ids, self.X, self.coef, self.scale, self.scale2, self.group = ids, X, coef, scale, scale2, group
        self.live = live


DW_ACCUMULATE, DW_ZERO_FILL, DW_PREZEROED = 0, 1, 2


def embedding_bwd_workspace_bytes(total: int, V1: int, d: int) -> int:
    return int(_lib.load().recguru_embedding_bwd_workspace_bytes(total, V1, d))


def _marshal_sites(sites: Sequence[Site], d: Optional[int]):
    """-> (ctypes array, tensors kept alive, total count).  ``d=None``: plan only, X not needed."""
    if not 1 <= len(sites) <= _lib.MAX_SITES:
        raise RuntimeError(f"recguru: 1..{_lib.MAX_SITES} sites supported, got {len(sites)}")
    arr = (LookupSite * len(sites))()
    keep = []
    total = 0
    for i, s in enumerate(sites):
        ids = _need(s.ids, torch.int64, "site.ids")
        n = ids.numel()
        if n % s.group:
            raise RuntimeError(f"recguru: site {i}: {n} ids is not a multiple of group {s.group}")
        X = None
        if d is not None:
            X = _need(s.X, torch.float32, "site.X")
            if X.numel() != (n // s.group) * d:
                raise RuntimeError(f"recguru: site {i}: ids {n}, group {s.group}, X {tuple(X.shape)}, d {d}")
        coef = None if s.coef is None else _need(s.coef, torch.float32, "site.coef")
        scale = None if s.scale is None else _need(s.scale, torch.float32, "site.scale")
        scale2 = None if s.scale2 is None else _need(s.scale2, torch.float32, "site.scale2")
        live = None if s.live is None else _need(s.live, torch.float32, "site.live")
        if coef is not None and coef.numel() != n:
            raise RuntimeError(f"recguru: site {i}: coef must have one entry per id")
        if live is not None and live.numel() != n // s.group:
            raise RuntimeError(f"recguru: site {i}: live must have one entry per row of X")
        keep += [ids, X, coef, scale, scale2, live]
        arr[i] = LookupSite(_ptr(ids), n, _ptr(X), _ptr(coef), _ptr(scale), _ptr(scale2), s.group, _ptr(live))
        total += n
    return arr, keep, total


def raw_embedding_bwd(sites: Sequence[Site], V1: int, d: int, padding_idx: int = -1,
                      out: Optional[Tensor] = None, zero_fill: bool = True,
                      workspace: Optional[Tensor] = None) -> Tensor:
    """Deterministic dense table gradient over all lookup sites in one pass.

    Replaces ATen embedding_dense_backward per site + the dense accumulation
    (train_auto.py:124).  ``padding_idx=-1`` means none (cross-domain tables,
    AutoEnc4Rec_cross.py:32,:35)."""
    lib = _lib.load()
    arr, keep, total = _marshal_sites(sites, d)
    dev = sites[0].X.device
    if out is None:
        out = torch.empty(V1, d, dtype=torch.float32, device=dev)
    else:
        out = _need(out, torch.float32, "out")
        if tuple(out.shape) != (V1, d):
            raise RuntimeError("recguru: out must be [V1, d]")
    need = embedding_bwd_workspace_bytes(total, V1, d)
    if workspace is None or workspace.numel() < need:
        workspace = _workspace(need, dev)
    check(lib.recguru_embedding_bwd(arr, len(sites), d, V1, int(padding_idx), _ptr(out),
                                    DW_ZERO_FILL if zero_fill else DW_ACCUMULATE, _ptr(workspace),
                                    workspace.numel(), _stream()), "embedding_bwd")
    return out


PLAN_ALL, PLAN_COMPACT, PLAN_SORT = 0, 1, 2


def raw_embedding_bwd_plan_live_rows(sites: Sequence[Site], V1: int, workspace: Tensor) -> Tuple[int, int]:
    """Device addresses ``(live_rows, n_live_rows)`` inside ``workspace`` of the ascending live-row list
    the row-mode compaction of these sites writes (every site shares one per-row ``live`` mask).
    Host-only; the list is valid once the compaction has run on the stream."""
    lib = _lib.load()
    arr, keep, total = _marshal_sites(sites, None)
    workspace = _need(workspace, torch.uint8, "workspace")
    rows, n_rows = C.c_void_p(), C.c_void_p()
    check(lib.recguru_embedding_bwd_plan_live_rows(arr, len(sites), V1, _ptr(workspace), workspace.numel(),
                                                   C.byref(rows), C.byref(n_rows)), "embedding_bwd_plan_live_rows")
    return int(rows.value), int(n_rows.value)


def raw_live_rows(mask: Tensor, rows_out: Optional[Tensor] = None, n_rows_out: Optional[Tensor] = None,
                  workspace: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """``rows[0 .. n_rows[0])`` = ascending flat indices with ``mask != 0`` (int32 storage read as uint32).
    Returns ``(rows, n_rows)``; pass ``(rows.data_ptr(), n_rows.data_ptr())`` as ``live_rows`` to
    :func:`raw_sampled_score_loss`."""
    lib = _lib.load()
    mask = _need(mask, torch.float32, "mask")
    n = mask.numel()
    dev = mask.device
    if rows_out is None:
        rows_out = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    if n_rows_out is None:
        n_rows_out = torch.empty(1, dtype=torch.int32, device=dev)
    if rows_out.dtype != torch.int32 or rows_out.numel() < n or n_rows_out.dtype != torch.int32:
        raise RuntimeError("recguru: rows_out must be an int32 tensor with one entry per mask element")
    need = int(lib.recguru_live_rows_workspace_bytes(n))
    if workspace is None or workspace.numel() < need:
        workspace = _workspace(need, dev)
    check(lib.recguru_live_rows(_ptr(mask), n, _ptr(rows_out), _ptr(n_rows_out), _ptr(workspace), workspace.numel(),
                                _stream()), "live_rows")
    return rows_out, n_rows_out


def raw_embedding_bwd_plan(sites: Sequence[Site], V1: int, padding_idx: int, workspace: Tensor,
                           mask_ids: Optional[Tensor] = None, pad_index: int = 0,
                           count3: Optional[Tensor] = None, stage: int = PLAN_ALL) -> None:
    """Stage A (compaction + stable sort) on the current stream; needs ids and live/coef only.
    ``stage``: ``PLAN_COMPACT`` then ``PLAN_SORT`` = the same work in two calls (an event can be
    recorded between them: the live-row list of the compaction feeds the scoring kernel).

    ``mask_ids``: derive the shared pad mask inside the plan (``recguru_embedding_bwd_plan_padmask``):
    every site's ``live`` must be the SAME tensor, which is written here as ``mask_ids != pad_index``
    (and its population count to ``count3``, the int32 ``[3]`` buffer of :func:`raw_pad_mask_count`)."""
    lib = _lib.load()
    arr, keep, total = _marshal_sites(sites, None)
    workspace = _need(workspace, torch.uint8, "workspace")
    if stage != PLAN_ALL:
        if mask_ids is not None:
            raise RuntimeError("recguru: the two-stage plan takes a ready mask (mask_ids must be None)")
        check(lib.recguru_embedding_bwd_plan_stage(arr, len(sites), V1, int(padding_idx), _ptr(workspace),
                                                   workspace.numel(), int(stage), _stream()),
              "embedding_bwd_plan_stage")
        return
    if mask_ids is None:
        check(lib.recguru_embedding_bwd_plan(arr, len(sites), V1, int(padding_idx), _ptr(workspace),
                                             workspace.numel(), _stream()), "embedding_bwd_plan")
        return
    mask_ids = _need(mask_ids, torch.int64, "mask_ids")
    mask_out = sites[0].live
    if mask_out is None or any(s.live is None or s.live.data_ptr() != mask_out.data_ptr() for s in sites):
        raise RuntimeError("recguru: mask_ids needs every site's live to be one shared mask tensor")
    if mask_out.numel() != mask_ids.numel() or not mask_out.is_contiguous():
        raise RuntimeError("recguru: the shared mask must be contiguous with one entry per mask id")
    if count3 is not None and (count3.dtype != torch.int32 or count3.numel() != 3 or not count3.is_cuda):
        raise RuntimeError("recguru: count3 must be a CUDA int32 [3] tensor")
    check(lib.recguru_embedding_bwd_plan_padmask(arr, len(sites), V1, int(padding_idx), _ptr(mask_ids),
                                                 int(pad_index), _ptr(mask_out), _ptr(count3), _ptr(workspace),
                                                 workspace.numel(), _stream()), "embedding_bwd_plan_padmask")


def raw_embedding_bwd_apply(sites: Sequence[Site], V1: int, d: int, out: Tensor, workspace: Tensor,
                            mode: int = DW_ZERO_FILL) -> Tensor:
    """Stage B (segmented reduce) over the plan left in ``workspace`` for the same ids."""
    lib = _lib.load()
    arr, keep, total = _marshal_sites(sites, d)
    out = _need(out, torch.float32, "out")
    if tuple(out.shape) != (V1, d):
        raise RuntimeError("recguru: out must be [V1, d]")
    workspace = _need(workspace, torch.uint8, "workspace")
    check(lib.recguru_embedding_bwd_apply(arr, len(sites), d, V1, _ptr(out), int(mode), _ptr(workspace),
                                          workspace.numel(), _stream()), "embedding_bwd_apply")
    return out


def raw_embedding_bwd_mark_rows(sites: Sequence[Site], V1: int, workspace: Tensor, row_epoch: Tensor,
                                epoch: int) -> Tensor:
    """``row_epoch[row] = epoch`` for every table row the apply stage writes for the plan left in
    ``workspace`` (same sites as the plan).  ``row_epoch``: int32/uint32-sized ``[V1]`` tensor, zeroed
    once; epochs count from 1.  Feeds :func:`raw_table_adam_rows`."""
    arr, keep, total = _marshal_sites(sites, None)
    workspace = _need(workspace, torch.uint8, "workspace")
    row_epoch = _need(row_epoch, torch.int32, "row_epoch")
    if row_epoch.numel() != V1 or not row_epoch.is_contiguous():
        raise RuntimeError("recguru: row_epoch must be a contiguous int32 [V1] tensor")
    check(_lib.load().recguru_embedding_bwd_mark_rows(arr, len(sites), V1, _ptr(workspace), workspace.numel(),
                                                      _ptr(row_epoch), int(epoch) & 0xFFFFFFFF, _stream()),
          "embedding_bwd_mark_rows")
    return row_epoch


def raw_pad_mask(ids: Tensor, pad_index: int, out: Optional[Tensor] = None) -> Tensor:
    """mask = (ids != pad_index) as fp32 (AutoEnc4Rec.py:71)."""
    ids = _need(ids, torch.int64, "ids")
    if out is None:
        out = torch.empty(ids.shape, dtype=torch.float32, device=ids.device)
    check(_lib.load().recguru_pad_mask(_ptr(ids), ids.numel(), int(pad_index), _ptr(out), _stream()), "pad_mask")
    return out


def raw_pad_mask_count(ids: Tensor, pad_index: int, out: Tensor, count3: Tensor) -> Tensor:
    """:func:`raw_pad_mask` plus the number of non-pad ids in ``count3[2]`` (``count3``: CUDA int32
    ``[3]``, zeroed once; the kernel leaves words 0-1 cleared)."""
    ids = _need(ids, torch.int64, "ids")
    if count3.dtype != torch.int32 or count3.numel() != 3 or not count3.is_cuda or not count3.is_contiguous():
        raise RuntimeError("recguru: count3 must be a contiguous CUDA int32 [3] tensor")
    out = _need(out, torch.float32, "mask_out")
    if out.numel() != ids.numel() or not out.is_contiguous():
        raise RuntimeError("recguru: mask_out must be contiguous with one entry per id")
    check(_lib.load().recguru_pad_mask_count(_ptr(ids), ids.numel(), int(pad_index), _ptr(out), _ptr(count3),
                                             _stream()), "pad_mask_count")
    return out


def raw_widen_ids(ids32: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """int32 ids (as a loader that collates into pinned int32 ships them) -> the int64 ids every
    entry point takes (data/data_loader.py:36,:418-421 delivers int64)."""
    if not ids32.is_cuda or ids32.dtype != torch.int32:
        raise RuntimeError("recguru: ids32 must be a CUDA int32 tensor")
    ids32 = ids32 if ids32.is_contiguous() else ids32.contiguous()
    if out is None:
        out = torch.empty(ids32.shape, dtype=torch.int64, device=ids32.device)
    elif out.dtype != torch.int64 or out.numel() != ids32.numel() or not out.is_contiguous():
        raise RuntimeError("recguru: out must be a contiguous int64 tensor of the same size")
    check(_lib.load().recguru_widen_ids(_ptr(ids32), ids32.numel(), _ptr(out), _stream()), "widen_ids")
    return out


def raw_debug_timestamp(slots: Tensor, index: int) -> None:
    """slots[index] = %globaltimer (ns) when the current stream gets here (int64 CUDA tensor); capturable."""
    if slots.dtype != torch.int64 or not slots.is_cuda:
        raise RuntimeError("recguru: slots must be a CUDA int64 tensor")
    check(_lib.load().recguru_debug_timestamp(slots.data_ptr() + 8 * int(index), _stream()), "debug_timestamp")


def raw_fill_zero(x: Tensor) -> Tensor:
    x = _need(x, torch.float32, "x")
    check(_lib.load().recguru_fill_zero(_ptr(x), x.numel(), _stream()), "fill_zero")
    return x


def raw_sort_ids(ids: Tensor, V1: int) -> Tuple[Tensor, Tensor]:
    """Stable ascending sort of ids; returns (sorted_ids, perm) == torch.sort(ids, stable=True)."""
    lib = _lib.load()
    ids = _need(ids, torch.int64, "ids").reshape(-1)
    n = ids.numel()
    sorted_ids = torch.empty_like(ids)
    perm = torch.empty_like(ids)
    ws = _workspace(lib.recguru_sort_ids_workspace_bytes(n, V1), ids.device)
    check(lib.recguru_sort_ids(_ptr(ids), n, V1, _ptr(sorted_ids), _ptr(perm), _ptr(ws), ws.numel(),
                               _stream()), "sort_ids")
    return sorted_ids, perm


def raw_sampled_score_loss(h: Tensor, weight: Tensor, pos: Tensor, neg: Tensor,
                           mask: Optional[Tensor], kind: Optional[str], want_grads: bool = True,
                           bufs: Optional[dict] = None, skip_masked: bool = False,
                           dh_out: Optional[Tensor] = None, live_count: Optional[Tensor] = None,
                           grad_scale: Optional[Tensor] = None,
                           live_rows: Optional[Tuple[int, int]] = None):
    """p = h.W[pos], n = h.W[neg]; loss per tools/lossfunctions.py:36-49,:56-72,:79-96.
    ``skip_masked``: do not score positions with mask == 0 (loss and derivatives unchanged, the
    logits of those positions are reported as 0 instead of the values nobody consumes).

    ``dh_out`` (with ``skip_masked``): also write d loss / d h in the same kernel
    (``recguru_sampled_score_loss_fwd_dh``); ``live_count`` = the int32 ``[3]`` buffer
    :func:`raw_pad_mask_count` filled for this mask, ``grad_scale`` an optional device scalar.
    ``live_rows`` (with ``skip_masked`` and ``live_count``): the pair of device addresses
    :func:`raw_embedding_bwd_plan_live_rows` returns -- the kernel walks that list of live positions
    instead of compacting the mask itself (``recguru_sampled_score_loss_fwd_dh_rows``; same bits).

    Returns (p [..,1], n [..,K], loss_aux [2] or None, dp_unit, dn_unit).  loss_aux[0] is the
    loss, loss_aux[1] = 1/sum(mask)."""
    lib = _lib.load()
    h = _need(h, torch.float32, "h")
    weight = _need(weight, torch.float32, "weight")
    pos = _need(pos, torch.int64, "pos")
    neg = _need(neg, torch.int64, "neg")
    V1, d = weight.shape
    n_pos = pos.numel()
    if n_pos == 0 or neg.numel() % n_pos:
        raise RuntimeError("recguru: neg must hold K ids per position")
    K = neg.numel() // n_pos
    if h.numel() != n_pos * d:
        raise RuntimeError(f"recguru: h has {h.numel()} elements, expected {n_pos}*{d}")
    if mask is not None:
        mask = _need(mask, torch.float32, "mask")
        if mask.numel() != n_pos:
            raise RuntimeError("recguru: mask must have one entry per position")
    dev = h.device
    b = bufs or {}
    p = b.get("p") if "p" in b else torch.empty(*pos.shape, 1, dtype=torch.float32, device=dev)
    n = b.get("n") if "n" in b else torch.empty(*pos.shape, K, dtype=torch.float32, device=dev)
    kind_i = -1 if kind is None else LOSS_KINDS[kind]
    if skip_masked and kind is not None and mask is not None:
        kind_i |= _lib.LOSS_SKIP_MASKED
    aux = dp = dn = ws = None
    if kind is not None:
        aux = b.get("aux") if "aux" in b else torch.empty(2, dtype=torch.float32, device=dev)
        if want_grads:
            dp = b.get("dp") if "dp" in b else torch.empty(n_pos, dtype=torch.float32, device=dev)
            dn = b.get("dn") if "dn" in b else torch.empty(n_pos, K, dtype=torch.float32, device=dev)
        ws = b.get("ws") if "ws" in b else _workspace(lib.recguru_sampled_workspace_bytes(n_pos, K), dev)
    if dh_out is not None or live_rows is not None:
        if dh_out is not None:
            dh_out = _need(dh_out, torch.float32, "dh_out")
            if dh_out.numel() != n_pos * d or not dh_out.is_contiguous():
                raise RuntimeError("recguru: dh_out must be a contiguous [n_pos, d] tensor")
        if live_count is None or live_count.dtype != torch.int32 or live_count.numel() != 3 or not live_count.is_cuda:
            raise RuntimeError("recguru: live_count must be the CUDA int32 [3] buffer of raw_pad_mask_count")
        gs = None if grad_scale is None else _need(grad_scale, torch.float32, "grad_scale")
        if live_rows is not None:
            rows_ptr, n_rows_ptr = live_rows
            check(lib.recguru_sampled_score_loss_fwd_dh_rows(
                _ptr(h), _ptr(weight), V1, d, _ptr(pos), _ptr(neg), n_pos, K, _ptr(mask), kind_i, _ptr(p), _ptr(n),
                _ptr(aux), _ptr(dp), _ptr(dn), _ptr(ws), 0 if ws is None else ws.numel(), _ptr(gs),
                live_count.data_ptr() + 8, rows_ptr, n_rows_ptr, _ptr(dh_out), _stream()),
                "sampled_score_loss_fwd_dh_rows")
            return p, n, aux, dp, dn
        check(lib.recguru_sampled_score_loss_fwd_dh(_ptr(h), _ptr(weight), V1, d, _ptr(pos), _ptr(neg), n_pos,
                                                    K, _ptr(mask), kind_i, _ptr(p), _ptr(n), _ptr(aux),
                                                    _ptr(dp), _ptr(dn), _ptr(ws), 0 if ws is None else ws.numel(),
                                                    _ptr(gs), live_count.data_ptr() + 8, _ptr(dh_out), _stream()),
              "sampled_score_loss_fwd_dh")
        return p, n, aux, dp, dn
    check(lib.recguru_sampled_score_loss_fwd(_ptr(h), _ptr(weight), V1, d, _ptr(pos), _ptr(neg), n_pos,
                                             K, _ptr(mask), kind_i, _ptr(p), _ptr(n), _ptr(aux),
                                             _ptr(dp), _ptr(dn), _ptr(ws),
                                             0 if ws is None else ws.numel(), _stream()),
          "sampled_score_loss_fwd")
    return p, n, aux, dp, dn


def raw_sampled_loss(p: Tensor, n: Tensor, mask: Optional[Tensor], kind: str):
    """Loss from logits only.  Returns (loss_aux [2], dp_unit [n_pos], dn_unit [n_pos, K])."""
    lib = _lib.load()
    p = _need(p, torch.float32, "p_logit")
    n = _need(n, torch.float32, "n_logit")
    n_pos = p.numel()
    if n_pos == 0 or n.numel() % n_pos:
        raise RuntimeError("recguru: n_logit must hold K logits per position")
    K = n.numel() // n_pos
    if mask is not None:
        mask = _need(mask, torch.float32, "mask")
        if mask.numel() != n_pos:
            raise RuntimeError("recguru: mask must have one entry per position")
    dev = p.device
    aux = torch.empty(2, dtype=torch.float32, device=dev)
    dp = torch.empty(n_pos, dtype=torch.float32, device=dev)
    dn = torch.empty(n_pos, K, dtype=torch.float32, device=dev)
    ws = _workspace(lib.recguru_sampled_workspace_bytes(n_pos, K), dev)
    check(lib.recguru_sampled_loss_fwd(_ptr(p), _ptr(n), n_pos, K, _ptr(mask), LOSS_KINDS[kind],
                                       _ptr(aux), _ptr(dp), _ptr(dn), _ptr(ws), ws.numel(), _stream()),
          "sampled_loss_fwd")
    return aux, dp, dn


def raw_sampled_score_bwd_dh(dp: Tensor, dn: Tensor, weight: Tensor, pos: Tensor, neg: Tensor,
                             scale: Optional[Tensor] = None, scale2: Optional[Tensor] = None,
                             out: Optional[Tensor] = None) -> Tensor:
    lib = _lib.load()
    dp = _need(dp, torch.float32, "dp")
    dn = _need(dn, torch.float32, "dn")
    weight = _need(weight, torch.float32, "weight")
    pos = _need(pos, torch.int64, "pos")
    neg = _need(neg, torch.int64, "neg")
    V1, d = weight.shape
    n_pos = pos.numel()
    K = neg.numel() // n_pos
    if out is None:
        out = torch.empty(*pos.shape, d, dtype=torch.float32, device=weight.device)
    check(lib.recguru_sampled_score_bwd_dh(_ptr(dp), _ptr(dn), _ptr(scale), _ptr(scale2), _ptr(weight),
                                           V1, d, _ptr(pos), _ptr(neg), n_pos, K, _ptr(out), _stream()),
          "sampled_score_bwd_dh")
    return out


def raw_candidate_scores_rank(u: Tensor, weight: Tensor, target: Tensor, cand: Tensor,
                              want_scores: bool = True):
    """get_scores + rank-of-column-0 (train_auto.py:164-199,:234-237) in one kernel.
    u [B,d] fp32, target [B], cand [B,C] int64 -> (scores [B,C+1] or None, rank [B] int64)."""
    lib = _lib.load()
    u = _need(u, torch.float32, "u")
    weight = _need(weight, torch.float32, "weight")
    target = _need(target, torch.int64, "target")
    cand = _need(cand, torch.int64, "n_items")
    B, d = u.shape
    V1 = weight.shape[0]
    if weight.shape[1] != d or target.numel() != B or cand.numel() % max(B, 1):
        raise RuntimeError("recguru: candidate_scores_rank shape mismatch")
    Cn = cand.numel() // max(B, 1)
    scores = torch.empty(B, Cn + 1, dtype=torch.float32, device=u.device) if want_scores else None
    rank = torch.empty(B, dtype=torch.int64, device=u.device)
    check(lib.recguru_candidate_scores_rank(_ptr(u), _ptr(weight), V1, d, _ptr(target), _ptr(cand), B, Cn,
                                            _ptr(scores), _ptr(rank), _stream()), "candidate_scores_rank")
    return scores, rank


def raw_cast_bf16(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    lib = _lib.load()
    x = _need(x, torch.float32, "x")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib.recguru_cast_bf16(_ptr(x), x.numel(), _ptr(out), _stream()), "cast_bf16")
    return out


def raw_catalog_target_scores(users_bf16: Tensor, items_bf16: Tensor, target: Tensor, id_lo: int,
                              out: Optional[Tensor] = None) -> Tensor:
    """Score of (u, target[u]) through the tensor-core path; 0 where the target is not in
    [id_lo, id_lo + n_items) (so that a sum over shards yields the owner's value)."""
    lib = _lib.load()
    users_bf16 = _need(users_bf16, torch.bfloat16, "users")
    items_bf16 = _need(items_bf16, torch.bfloat16, "items")
    target = _need(target, torch.int64, "target")
    U, d = users_bf16.shape
    if out is None:
        out = torch.empty(U, dtype=torch.float32, device=users_bf16.device)
    check(lib.recguru_catalog_target_scores(_ptr(users_bf16), U, _ptr(items_bf16), int(id_lo),
                                            items_bf16.shape[0], d, _ptr(target), _ptr(out), _stream()),
          "catalog_target_scores")
    return out


def raw_catalog_topk_rank(users_bf16: Tensor, items_bf16: Tensor, target: Tensor,
                          target_score: Tensor, k: int, id_lo: int, bufs: Optional[dict] = None,
                          tie: str = "id"):
    """Fused scoring GEMM + running top-k + rank count over items [id_lo, id_lo+n_items).
    ``tie``: how items scoring exactly the target's score count -- ``"id"`` (lower id first: the stable
    sort, default), ``"target_first"`` (rank = #{s > t}), ``"target_last"`` (rank = #{s >= t} - 1)."""
    lib = _lib.load()
    users_bf16 = _need(users_bf16, torch.bfloat16, "users")
    items_bf16 = _need(items_bf16, torch.bfloat16, "items")
    target = _need(target, torch.int64, "target")
    target_score = _need(target_score, torch.float32, "target_score")
    U, d = users_bf16.shape
    n_items = items_bf16.shape[0]
    if items_bf16.shape[1] != d or target.numel() != U or target_score.numel() != U:
        raise RuntimeError("recguru: catalog_topk_rank shape mismatch")
    dev = users_bf16.device
    b = bufs or {}
    val = b.get("val") if "val" in b else torch.empty(U, k, dtype=torch.float32, device=dev)
    idx = b.get("idx") if "idx" in b else torch.empty(U, k, dtype=torch.int64, device=dev)
    cnt = b.get("cnt") if "cnt" in b else torch.empty(U, dtype=torch.int64, device=dev)
    need = lib.recguru_catalog_workspace_bytes(U, n_items, d, k)
    if need < 0:
        raise _lib.RecGuruError(f"catalog_workspace_bytes rejected U={U} n_items={n_items} k={k}")
    ws = b.get("ws") if "ws" in b and b["ws"].numel() >= need else _workspace(need, dev)
    if tie == "id":
        check(lib.recguru_catalog_topk_rank(_ptr(users_bf16), U, _ptr(items_bf16), int(id_lo), n_items, d,
                                            _ptr(target), _ptr(target_score), k, _ptr(val), _ptr(idx),
                                            _ptr(cnt), _ptr(ws), ws.numel(), _stream()),
              "catalog_topk_rank")
    else:
        check(lib.recguru_catalog_topk_rank_ties(_ptr(users_bf16), U, _ptr(items_bf16), int(id_lo), n_items, d,
                                                 _ptr(target), _ptr(target_score), k, _lib.TIE_MODES[tie],
                                                 _ptr(val), _ptr(idx), _ptr(cnt), _ptr(ws), ws.numel(), _stream()),
              "catalog_topk_rank_ties")
    return val, idx, cnt


def raw_catalog_scores_dense(users_bf16: Tensor, items_bf16: Tensor) -> Tensor:
    """Debug/parity: the dense [U, n_items] score matrix from the same tcgen05 data path."""
    lib = _lib.load()
    users_bf16 = _need(users_bf16, torch.bfloat16, "users")
    items_bf16 = _need(items_bf16, torch.bfloat16, "items")
    U, d = users_bf16.shape
    n_items = items_bf16.shape[0]
    out = torch.empty(U, n_items, dtype=torch.float32, device=users_bf16.device)
    check(lib.recguru_catalog_scores_dense(_ptr(users_bf16), U, _ptr(items_bf16), n_items, d, _ptr(out),
                                           _stream()), "catalog_scores_dense")
    return out


def raw_topk_merge(vals: Tensor, ids: Tensor, counts: Optional[Tensor]):
    """vals/ids [G,U,k] (+ counts [G,U]) -> (top-k vals [U,k], ids [U,k], rank [U] or None)."""
    lib = _lib.load()
    vals = _need(vals, torch.float32, "vals")
    ids = _need(ids, torch.int64, "ids")
    G, U, k = vals.shape
    dev = vals.device
    out_v = torch.empty(U, k, dtype=torch.float32, device=dev)
    out_i = torch.empty(U, k, dtype=torch.int64, device=dev)
    out_r = None
    if counts is not None:
        counts = _need(counts, torch.int64, "counts")
        out_r = torch.empty(U, dtype=torch.int64, device=dev)
    check(lib.recguru_topk_merge(_ptr(vals), _ptr(ids), _ptr(counts), G, U, k, _ptr(out_v), _ptr(out_i),
                                 _ptr(out_r), _stream()), "topk_merge")
    return out_v, out_i, out_r


def raw_rank_histogram(rank: Tensor, kmax: int, hist: Optional[Tensor] = None,
                       accumulate: bool = False) -> Tensor:
    """hist[r] = #{rank == r} for r < kmax; hist[kmax] = #{rank >= kmax}  (int64, exact)."""
    lib = _lib.load()
    rank = _need(rank, torch.int64, "rank").reshape(-1)
    if hist is None:
        hist = torch.empty(kmax + 1, dtype=torch.int64, device=rank.device)
        accumulate = False
    check(lib.recguru_rank_histogram(_ptr(rank), rank.numel(), kmax, _ptr(hist), 1 if accumulate else 0,
                                     _stream()), "rank_histogram")
    return hist


def raw_table_adam(weight: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, grad: Tensor, lr: float,
                   beta1: float, beta2: float, eps: float, step: int) -> None:
    """One torch.optim.Adam step of the item table (train_auto.py:354-364) in a single pass."""
    lib = _lib.load()
    bufs = [_need(t, torch.float32, n) for t, n in ((weight, "weight"), (exp_avg, "exp_avg"),
                                                    (exp_avg_sq, "exp_avg_sq"), (grad, "grad"))]
    if any(not t.is_contiguous() for t in (weight, exp_avg, exp_avg_sq)):
        raise RuntimeError("recguru: table_adam updates in place: weight and moments must be contiguous")
    if len({t.numel() for t in bufs}) != 1:
        raise RuntimeError("recguru: table_adam: weight, moments and grad must have the same size")
    check(lib.recguru_table_adam(_ptr(weight), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(bufs[3]), weight.numel(),
                                 float(lr), float(beta1), float(beta2), float(eps), int(step), _stream()),
          "table_adam")


def raw_table_adam_rows(weight: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, grad_rows: Tensor,
                        row_epoch: Tensor, epoch: int, lr: float, beta1: float, beta2: float, eps: float,
                        step: int) -> None:
    """:func:`raw_table_adam` with the gradient given as a row store: rows whose ``row_epoch`` differs
    from ``epoch`` take g = 0 and are not read (bit-identical to the dense call on a zero-filled
    gradient)."""
    lib = _lib.load()
    bufs = [_need(t, torch.float32, n) for t, n in ((weight, "weight"), (exp_avg, "exp_avg"),
                                                    (exp_avg_sq, "exp_avg_sq"), (grad_rows, "grad_rows"))]
    if any(not t.is_contiguous() for t in (weight, exp_avg, exp_avg_sq, grad_rows)):
        raise RuntimeError("recguru: table_adam_rows: weight, moments and the row store must be contiguous")
    if weight.dim() != 2 or len({tuple(t.shape) for t in bufs}) != 1:
        raise RuntimeError("recguru: table_adam_rows: weight, moments and grad_rows must all be [V1, d]")
    V1, d = weight.shape
    row_epoch = _need(row_epoch, torch.int32, "row_epoch")
    if row_epoch.numel() != V1 or not row_epoch.is_contiguous():
        raise RuntimeError("recguru: row_epoch must be a contiguous int32 [V1] tensor")
    check(lib.recguru_table_adam_rows(_ptr(weight), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(grad_rows),
                                      _ptr(row_epoch), int(epoch) & 0xFFFFFFFF, V1, d, float(lr), float(beta1),
                                      float(beta2), float(eps), int(step), _stream()), "table_adam_rows")


# --------------------------------------------------------------------------------------
# device guard: every raw_* call runs on the device of its tensors
# --------------------------------------------------------------------------------------
def _tensor_devices(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            yield a
        elif isinstance(a, Site):
            yield from (t for t in (a.ids, a.X, a.coef, a.scale, a.scale2, a.live) if t is not None)
        elif isinstance(a, (list, tuple)) and a and isinstance(a[0], Site):
            for s in a:
                yield from (t for t in (s.ids, s.X, s.coef, s.scale, s.scale2, s.live) if t is not None)
        elif isinstance(a, dict):
            yield from (t for t in a.values() if isinstance(t, torch.Tensor))


def _on_tensor_device(fn):
    """The C ABI launches on the CURRENT device and the stream it is handed.  A caller that drives
    several GPUs from one process (nn.DataParallel replicas, train_auto.py:350-351) may call with
    tensors of a device that is not current: run under that device's guard (and its current stream)
    and refuse tensors that live on different devices -- a raw pointer of the wrong device would be
    an illegal address, not an exception."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = args[0] if args else None
        if not isinstance(t, torch.Tensor):
            t = next(_tensor_devices(args, kwargs), None)
        # fast path (one comparison per call): the first tensor lives on the current device
        if t is None or not t.is_cuda or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        dev = t.device
        for o in _tensor_devices(args, kwargs):
            if o.is_cuda and o.device != dev:
                raise RuntimeError(f"recguru: {fn.__name__}: tensors on different devices ({dev} and {o.device})")
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


for _name, _f in list(globals().items()):
    if _name.startswith("raw_") and callable(_f):
        globals()[_name] = _on_tensor_device(_f)
del _name, _f


# --------------------------------------------------------------------------------------
# op layer: torch.library custom ops + autograd
# --------------------------------------------------------------------------------------
_lib_def = torch.library.Library("recguru", "DEF")
_lib_def.define("gather_pe_mask(Tensor weight, Tensor ids, Tensor pe, int pad_index) -> (Tensor, Tensor)")
_lib_def.define("gather_pe_mask_m(Tensor weight, Tensor ids, Tensor pe, Tensor mask) -> Tensor")
_lib_def.define("gather_rows(Tensor weight, Tensor ids, int padding_idx) -> Tensor")
_lib_def.define("pe_mask(Tensor x, Tensor pe, Tensor mask) -> Tensor")
_lib_def.define("gather_pe_mask_dropout(Tensor weight, Tensor ids, Tensor pe, Tensor? mask, int pad_index, float p,"
                " int seed, int offset, int threads) -> (Tensor, Tensor)")
_lib_def.define("dropout(Tensor x, Tensor? row_mask, float p, int seed, int offset, int threads) -> Tensor")
_lib_def.define("pe_grad(Tensor g, Tensor mask) -> Tensor")
_lib_def.define("extract_last(Tensor h) -> Tensor")
_lib_def.define("embedding_bwd(Tensor grad, Tensor ids, Tensor? coef, int V1, int padding_idx) -> Tensor")
_lib_def.define("sort_ids(Tensor ids, int V1) -> (Tensor, Tensor)")
_lib_def.define("sampled_scores(Tensor h, Tensor weight, Tensor pos, Tensor neg, int padding_idx) -> (Tensor, Tensor)")
_lib_def.define("sampled_loss(Tensor p, Tensor n, Tensor? mask, str kind) -> (Tensor, Tensor, Tensor)")
_lib_def.define("sampled_score_loss(Tensor h, Tensor weight, Tensor pos, Tensor neg, Tensor? mask, str kind, int padding_idx)"
                " -> (Tensor, Tensor, Tensor, Tensor, Tensor)")
_lib_def.define("sampled_score_bwd(Tensor dp, Tensor dn, Tensor? scale, Tensor? scale2, Tensor h, Tensor weight,"
                " Tensor pos, Tensor neg, int padding_idx) -> (Tensor, Tensor)")
_lib_def.define("candidate_scores_rank(Tensor u, Tensor weight, Tensor target, Tensor cand) -> (Tensor, Tensor)")
_lib_def.define("cast_bf16(Tensor x) -> Tensor")
_lib_def.define("catalog_topk_rank(Tensor users, Tensor items_bf16, Tensor target, int k, int id_lo)"
                " -> (Tensor, Tensor, Tensor)")
_lib_def.define("topk_merge(Tensor vals, Tensor ids, Tensor counts) -> (Tensor, Tensor, Tensor)")
_lib_def.define("rank_histogram(Tensor rank, int kmax) -> Tensor")


def _impl(name):
    def deco(fn):
        torch.library.impl(_lib_def, name, "CUDA")(fn)
        return fn
    return deco


@_impl("gather_pe_mask")
def _gather_pe_mask(weight, ids, pe, pad_index):
    out, mask = raw_gather_pe_mask(weight, ids, pe, pad_index)
    return out, mask


@_impl("gather_pe_mask_m")
def _gather_pe_mask_m(weight, ids, pe, mask):
    out, _ = raw_gather_pe_mask(weight, ids, pe, -1, mask_in=mask, want_mask=False)
    return out


@_impl("gather_rows")
def _gather_rows(weight, ids, padding_idx):
    out, _ = raw_gather_pe_mask(weight, ids, None, -1, want_mask=False)
    return out


@_impl("pe_mask")
def _pe_mask(x, pe, mask):
    return raw_pe_mask(x, pe, mask)


@_impl("gather_pe_mask_dropout")
def _gather_pe_mask_dropout(weight, ids, pe, mask, pad_index, p, seed, offset, threads):
    out, m = raw_gather_pe_mask_dropout(weight, ids, pe, pad_index, p, seed, offset, threads, mask_in=mask)
    return out, (mask if mask is not None else m)


@_impl("dropout")
def _dropout(x, row_mask, p, seed, offset, threads):
    return raw_dropout(x, p, seed, offset, threads, row_mask=row_mask)


@_impl("pe_grad")
def _pe_grad(g, mask):
    return raw_pe_grad(g, mask)


@_impl("extract_last")
def _extract_last(h):
    return raw_extract_last(h)[0]


@_impl("embedding_bwd")
def _embedding_bwd(grad, ids, coef, V1, padding_idx):
    d = grad.shape[-1]
    return raw_embedding_bwd([Site(ids, grad, coef)], V1, d, padding_idx)


@_impl("sort_ids")
def _sort_ids(ids, V1):
    return raw_sort_ids(ids, V1)


@_impl("sampled_scores")
def _sampled_scores(h, weight, pos, neg, padding_idx):
    p, n, _, _, _ = raw_sampled_score_loss(h, weight, pos, neg, None, None)
    return p, n


@_impl("sampled_loss")
def _sampled_loss(p, n, mask, kind):
    aux, dp, dn = raw_sampled_loss(p, n, mask, kind)
    return aux, dp, dn


@_impl("sampled_score_loss")
def _sampled_score_loss(h, weight, pos, neg, mask, kind, padding_idx):
    # kind "bpr_sas/skip": positions under the mask are not scored (padding-aware kernel)
    base, _, flag = kind.partition("/")
    p, n, aux, dp, dn = raw_sampled_score_loss(h, weight, pos, neg, mask, base,
                                               skip_masked=flag == "skip" and mask is not None)
    return aux, p, n, dp, dn


@_impl("sampled_score_bwd")
def _sampled_score_bwd(dp, dn, scale, scale2, h, weight, pos, neg, padding_idx):
    V1, d = weight.shape
    n_pos = pos.numel()
    K = neg.numel() // n_pos
    dh = raw_sampled_score_bwd_dh(dp, dn, weight, pos, neg, scale, scale2)
    dh = dh.view(h.shape)
    h2 = h.reshape(n_pos, d)
    dW = raw_embedding_bwd([Site(pos, h2, dp, scale, scale2, 1), Site(neg, h2, dn, scale, scale2, K)],
                           V1, d, padding_idx)
    return dh, dW


@_impl("candidate_scores_rank")
def _candidate_scores_rank(u, weight, target, cand):
    return raw_candidate_scores_rank(u, weight, target, cand, True)


@_impl("cast_bf16")
def _cast_bf16(x):
    return raw_cast_bf16(x)


@_impl("catalog_topk_rank")
def _catalog_topk_rank(users, items_bf16, target, k, id_lo):
    u = users if users.dtype == torch.bfloat16 else raw_cast_bf16(users)
    ts = raw_catalog_target_scores(u, items_bf16, target, id_lo)
    return raw_catalog_topk_rank(u, items_bf16, target, ts, k, id_lo)


@_impl("topk_merge")
def _topk_merge(vals, ids, counts):
    return raw_topk_merge(vals, ids, counts)


@_impl("rank_histogram")
def _rank_histogram(rank, kmax):
    return raw_rank_histogram(rank, kmax)


# ---- fake (meta) kernels: shape inference -----------------------------------------------
def _fake(name):
    def deco(fn):
        torch.library.register_fake(f"recguru::{name}")(fn)
        return fn
    return deco


@_fake("gather_pe_mask")
def _(weight, ids, pe, pad_index):
    return weight.new_empty(*ids.shape, weight.shape[1]), weight.new_empty(ids.shape)


@_fake("gather_pe_mask_m")
def _(weight, ids, pe, mask):
    return weight.new_empty(*ids.shape, weight.shape[1])


@_fake("gather_rows")
def _(weight, ids, padding_idx):
    return weight.new_empty(*ids.shape, weight.shape[1])


@_fake("pe_mask")
def _(x, pe, mask):
    return torch.empty_like(x)


@_fake("gather_pe_mask_dropout")
def _(weight, ids, pe, mask, pad_index, p, seed, offset, threads):
    return weight.new_empty(*ids.shape, weight.shape[1]), weight.new_empty(ids.shape)


@_fake("dropout")
def _(x, row_mask, p, seed, offset, threads):
    return torch.empty_like(x)


@_fake("pe_grad")
def _(g, mask):
    return g.new_empty(g.shape[-2], g.shape[-1])


@_fake("extract_last")
def _(h):
    return h.new_empty(h.shape[0], h.shape[-1])


@_fake("embedding_bwd")
def _(grad, ids, coef, V1, padding_idx):
    return grad.new_empty(V1, grad.shape[-1])


@_fake("sort_ids")
def _(ids, V1):
    return ids.new_empty(ids.numel()), ids.new_empty(ids.numel())


@_fake("sampled_scores")
def _(h, weight, pos, neg, padding_idx):
    K = neg.numel() // max(pos.numel(), 1)
    return h.new_empty(*pos.shape, 1), h.new_empty(*pos.shape, K)


@_fake("sampled_loss")
def _(p, n, mask, kind):
    n_pos = p.numel()
    return p.new_empty(2), p.new_empty(n_pos), p.new_empty(n_pos, n.numel() // max(n_pos, 1))


@_fake("sampled_score_loss")
def _(h, weight, pos, neg, mask, kind, padding_idx):
    n_pos = pos.numel()
    K = neg.numel() // max(n_pos, 1)
    return (h.new_empty(2), h.new_empty(*pos.shape, 1), h.new_empty(*pos.shape, K),
            h.new_empty(n_pos), h.new_empty(n_pos, K))


@_fake("sampled_score_bwd")
def _(dp, dn, scale, scale2, h, weight, pos, neg, padding_idx):
    return torch.empty_like(h), torch.empty_like(weight)


@_fake("candidate_scores_rank")
def _(u, weight, target, cand):
    B = u.shape[0]
    return u.new_empty(B, cand.numel() // max(B, 1) + 1), target.new_empty(B)


@_fake("cast_bf16")
def _(x):
    return x.new_empty(x.shape, dtype=torch.bfloat16)


@_fake("catalog_topk_rank")
def _(users, items_bf16, target, k, id_lo):
    U = users.shape[0]
    return (users.new_empty(U, k, dtype=torch.float32), target.new_empty(U, k), target.new_empty(U))


@_fake("topk_merge")
def _(vals, ids, counts):
    _, U, k = vals.shape
    return vals.new_empty(U, k), ids.new_empty(U, k), ids.new_empty(U)


@_fake("rank_histogram")
def _(rank, kmax):
    return rank.new_empty(kmax + 1)


# ---- one table backward per backward pass (drop-in surface) -----------------------------------
# The reference's unchanged call sites look a table up several times per step (input ids,
# positives, negatives; AutoEnc4Rec.py:92,:115,:118).  Through plain autograd every lookup op
# would return its own dense [V1, d] gradient and autograd would add them (what ATen does:
# embedding_dense_backward per site).  For a table registered here the lookup
# ops' backward functions return NO weight gradient: they queue their (ids, source rows,
# coefficients) as sites, and one engine callback -- run when the whole backward pass has
# finished -- reduces all of them in ONE deterministic sort + segmented-reduce pass and
# accumulates the dense result into ``param.grad`` (the optimizer sees what it always saw).
_one_pass = {}        # weight.data_ptr() -> (weakref to the Parameter, padding_idx)
_pending = {}         # weight.data_ptr() -> list[Site] of the running backward pass
ONE_PASS_DEFAULT = os.environ.get("RECGURU_ONE_PASS_BWD", "1") != "0"


def enable_one_pass_backward(param: Tensor, padding_idx: Optional[int] = None) -> None:
    """Register ``param`` (an item table): gradients of the recguru lookup ops w.r.t. it are reduced
    in one pass per ``backward()`` and written to ``param.grad``.  (``torch.autograd.grad`` w.r.t.
    the table then returns nothing for these ops -- use ``backward()``, as the reference does.)"""
    _one_pass[param.data_ptr()] = (weakref.ref(param), -1 if padding_idx is None else int(padding_idx))


def disable_one_pass_backward(param: Tensor) -> None:
    _one_pass.pop(param.data_ptr(), None)


def _flush_table(key: int) -> None:
    sites = _pending.pop(key, [])
    ent = _one_pass.get(key)
    param = ent[0]() if ent is not None else None
    if not sites or param is None:
        return
    V1, d = param.shape
    dW = None
    for i in range(0, len(sites), _lib.MAX_SITES):          # > 8 lookup sites: several passes, accumulated
        grp = sites[i:i + _lib.MAX_SITES]
        dW = raw_embedding_bwd(grp, V1, d, ent[1], out=dW, zero_fill=dW is None)
    with torch.no_grad():
        if param.grad is None:
            param.grad = dW
        else:
            param.grad.add_(dW)


def _defer_sites(weight: Tensor, sites) -> bool:
    """Queue ``sites`` for the table ``weight`` if it is registered; True if queued."""
    if not _one_pass:
        return False
    key = weight.data_ptr()
    ent = _one_pass.get(key)
    if ent is None:
        return False
    if ent[0]() is None:
        _one_pass.pop(key, None)
        return False
    lst = _pending.get(key)
    if lst is None:
        lst = _pending[key] = []
        torch.autograd.Variable._execution_engine.queue_callback(lambda: _flush_table(key))
    lst.extend(sites)
    return True


# ---- autograd ------------------------------------------------------------------------------
def _setup_gather_pe_mask(ctx, inputs, output):
    weight, ids, pe, pad_index = inputs
    _, mask = output
    ctx.save_for_backward(ids, mask)
    ctx.mark_non_differentiable(mask)   # a 0/1 function of the ids: keep it out of the graph
    ctx.V1 = weight.shape[0]
    ctx.wref = weight.detach()          # same storage, no graph: identifies the table in backward
    ctx.pe_needs_grad = pe.requires_grad   # learned positions (PositionalEncodingM, transformer.py:327-342)
    ctx.pe_shape = tuple(pe.shape)


def _fit_pe_grad(d_pe: Tensor, pe_shape) -> Tensor:
    """[L, d] gradient -> the shape of the position table the op was given ([>=L, d] or [1, >=L, d])."""
    L, d = d_pe.shape
    rows = pe_shape[-2]
    if rows != L:
        full = d_pe.new_zeros(rows, d)
        full[:L] = d_pe
        d_pe = full
    return d_pe.reshape(pe_shape)


def _pe_grad_of(ctx, g_out, ids, mask):
    if not ctx.pe_needs_grad:
        return None
    L, d = ids.shape[-1], g_out.shape[-1]
    return _fit_pe_grad(torch.ops.recguru.pe_grad(g_out.reshape(-1, L, d).contiguous(), mask.reshape(-1, L)),
                        ctx.pe_shape)


def _bwd_gather_pe_mask(ctx, g_out, g_mask):
    ids, mask = ctx.saved_tensors
    d = g_out.shape[-1]
    d_pe = _pe_grad_of(ctx, g_out, ids, mask)
    if _defer_sites(ctx.wref, [Site(ids.reshape(-1), g_out.reshape(-1, d), coef=mask.reshape(-1))]):
        return None, None, d_pe, None
    # d out / d W[id] = mask ; pad rows get coef 0 (== the reference: x*mask kills their grad,
    # and nn.Embedding(padding_idx) would skip them anyway)
    dW = torch.ops.recguru.embedding_bwd(g_out.reshape(-1, d), ids.reshape(-1), mask.reshape(-1),
                                         ctx.V1, -1)
    return dW, None, d_pe, None


torch.library.register_autograd("recguru::gather_pe_mask", _bwd_gather_pe_mask,
                                setup_context=_setup_gather_pe_mask)


def _setup_gather_pe_mask_m(ctx, inputs, output):
    weight, ids, pe, mask = inputs
    ctx.save_for_backward(ids, mask)
    ctx.V1 = weight.shape[0]
    ctx.wref = weight.detach()
    ctx.pe_needs_grad = pe.requires_grad
    ctx.pe_shape = tuple(pe.shape)


def _bwd_gather_pe_mask_m(ctx, g_out):
    ids, mask = ctx.saved_tensors
    d = g_out.shape[-1]
    d_pe = _pe_grad_of(ctx, g_out, ids, mask)
    if _defer_sites(ctx.wref, [Site(ids.reshape(-1), g_out.reshape(-1, d), coef=mask.reshape(-1))]):
        return None, None, d_pe, None
    dW = torch.ops.recguru.embedding_bwd(g_out.reshape(-1, d), ids.reshape(-1), mask.reshape(-1),
                                         ctx.V1, -1)
    return dW, None, d_pe, None


torch.library.register_autograd("recguru::gather_pe_mask_m", _bwd_gather_pe_mask_m,
                                setup_context=_setup_gather_pe_mask_m)


def _setup_gather_pe_mask_dropout(ctx, inputs, output):
    weight, ids, pe, mask_in, pad_index, p, seed, offset, threads = inputs
    _, mask = output
    ctx.save_for_backward(ids, mask)
    ctx.mark_non_differentiable(mask)
    ctx.V1 = weight.shape[0]
    ctx.wref = weight.detach()
    ctx.drop = (p, seed, offset, threads)
    ctx.pe_needs_grad = pe.requires_grad
    ctx.pe_shape = tuple(pe.shape)


def _bwd_gather_pe_mask_dropout(ctx, g_out, g_mask):
    ids, mask = ctx.saved_tensors
    d = g_out.shape[-1]
    # through the dropout first (same Philox state => same keep pattern), then as gather_pe_mask
    p, seed, offset, threads = ctx.drop
    g = g_out.contiguous()
    if p > 0:
        g = torch.ops.recguru.dropout(g, mask, p, seed, offset, threads)
    d_pe = _pe_grad_of(ctx, g, ids, mask)
    m1 = mask.reshape(-1)
    if _defer_sites(ctx.wref, [Site(ids.reshape(-1), g.reshape(-1, d), coef=m1, live=m1)]):
        dW = None
    else:
        dW = torch.ops.recguru.embedding_bwd(g.reshape(-1, d), ids.reshape(-1), mask.reshape(-1), ctx.V1, -1)
    return dW, None, d_pe, None, None, None, None, None, None


torch.library.register_autograd("recguru::gather_pe_mask_dropout", _bwd_gather_pe_mask_dropout,
                                setup_context=_setup_gather_pe_mask_dropout)


def _setup_dropout(ctx, inputs, output):
    x, row_mask, p, seed, offset, threads = inputs
    if row_mask is not None:
        ctx.save_for_backward(row_mask)
    ctx.has_mask = row_mask is not None
    ctx.drop = (p, seed, offset, threads)


def _bwd_dropout(ctx, g):
    row_mask = ctx.saved_tensors[0] if ctx.has_mask else None
    p, seed, offset, threads = ctx.drop
    return torch.ops.recguru.dropout(g.contiguous(), row_mask, p, seed, offset, threads), None, None, None, None, None


torch.library.register_autograd("recguru::dropout", _bwd_dropout, setup_context=_setup_dropout)


def _setup_extract_last(ctx, inputs, output):
    (h,) = inputs
    ctx.shape = h.shape


def _bwd_extract_last(ctx, g):
    dh = g.new_zeros(ctx.shape)
    dh[:, -1, :] = g
    return dh


torch.library.register_autograd("recguru::extract_last", _bwd_extract_last, setup_context=_setup_extract_last)


def _setup_pe_mask(ctx, inputs, output):
    _, pe, mask = inputs
    ctx.save_for_backward(mask)
    ctx.pe_needs_grad = pe.requires_grad
    ctx.pe_shape = tuple(pe.shape)


def _bwd_pe_mask(ctx, g):
    (mask,) = ctx.saved_tensors
    d_pe = None
    if ctx.pe_needs_grad:
        L, d = g.shape[-2], g.shape[-1]
        d_pe = _fit_pe_grad(torch.ops.recguru.pe_grad(g.reshape(-1, L, d).contiguous(), mask.reshape(-1, L)),
                            ctx.pe_shape)
    return g * mask.reshape(*g.shape[:-1], 1), d_pe, None


torch.library.register_autograd("recguru::pe_mask", _bwd_pe_mask, setup_context=_setup_pe_mask)


def _setup_sampled_scores(ctx, inputs, output):
    h, weight, pos, neg, padding_idx = inputs
    ctx.save_for_backward(h, weight, pos, neg)
    ctx.padding_idx = padding_idx


def _bwd_sampled_scores(ctx, g_p, g_n):
    h, weight, pos, neg = ctx.saved_tensors
    n_pos = pos.numel()
    K = neg.numel() // n_pos
    dp = g_p.reshape(n_pos).contiguous()
    dn = g_n.reshape(n_pos, K).contiguous()
    if weight.data_ptr() in _one_pass:
        dh = raw_sampled_score_bwd_dh(dp, dn, weight, pos, neg).view(h.shape)
        h2 = h.reshape(n_pos, h.shape[-1])
        if _defer_sites(weight, [Site(pos.reshape(-1), h2, dp), Site(neg.reshape(-1), h2, dn.view(-1), group=K)]):
            return dh, None, None, None, None
    dh, dW = torch.ops.recguru.sampled_score_bwd(dp, dn, None, None, h, weight, pos, neg,
                                                 ctx.padding_idx)
    return dh, dW, None, None, None


torch.library.register_autograd("recguru::sampled_scores", _bwd_sampled_scores,
                                setup_context=_setup_sampled_scores)


def _setup_sampled_loss(ctx, inputs, output):
    p, n, mask, kind = inputs
    aux, dp, dn = output
    ctx.save_for_backward(aux, dp, dn)
    ctx.shapes = (p.shape, n.shape)
    ctx.mark_non_differentiable(dp, dn)


def _bwd_sampled_loss(ctx, g_aux, g_dp, g_dn):
    aux, dp, dn = ctx.saved_tensors
    s = g_aux[0] * aux[1]
    return (dp * s).view(ctx.shapes[0]), (dn * s).view(ctx.shapes[1]), None, None


torch.library.register_autograd("recguru::sampled_loss", _bwd_sampled_loss,
                                setup_context=_setup_sampled_loss)


def _setup_sampled_score_loss(ctx, inputs, output):
    h, weight, pos, neg, mask, kind, padding_idx = inputs
    aux, p, n, dp, dn = output
    if mask is not None:
        ctx.save_for_backward(h, weight, pos, neg, aux, dp, dn, mask)
    else:
        ctx.save_for_backward(h, weight, pos, neg, aux, dp, dn)
    ctx.has_mask = mask is not None
    ctx.padding_idx = padding_idx
    ctx.mark_non_differentiable(dp, dn)
    ctx.set_materialize_grads(False)  # g_p / g_n stay None unless p / n are used downstream


def _bwd_sampled_score_loss(ctx, g_aux, g_p, g_n, g_dp, g_dn):
    h, weight, pos, neg, aux, dp, dn = ctx.saved_tensors[:7]
    # the loss mask doubles as the sites' liveness mask: when the input lookup of the same step hangs off
    # the SAME mask tensor, the one-pass table backward compacts by row (7x less to scan at cfg2)
    live = ctx.saved_tensors[7].reshape(-1) if ctx.has_mask else None
    n_pos = pos.numel()
    K = neg.numel() // n_pos
    inv_m = aux[1:2]                      # 1/sum(mask), still missing from dp/dn
    if g_p is None and g_n is None and g_aux is not None:
        go = g_aux[0:1].contiguous()      # upstream grad of the scalar loss (device scalar)
        if weight.data_ptr() in _one_pass:
            dh = raw_sampled_score_bwd_dh(dp, dn, weight, pos, neg, go, inv_m).view(h.shape)
            h2 = h.reshape(n_pos, h.shape[-1])
            if _defer_sites(weight, [Site(pos.reshape(-1), h2, dp, go, inv_m, live=live),
                                     Site(neg.reshape(-1), h2, dn.view(-1), go, inv_m, group=K, live=live)]):
                return dh, None, None, None, None, None, None
        dh, dW = torch.ops.recguru.sampled_score_bwd(dp, dn, go, inv_m, h, weight, pos, neg,
                                                     ctx.padding_idx)
    else:
        # logits were also used downstream: fold their grads in (rare; plain torch arithmetic)
        dp_t = torch.zeros_like(dp) if g_aux is None else dp * (g_aux[0] * aux[1])
        dn_t = torch.zeros_like(dn) if g_aux is None else dn * (g_aux[0] * aux[1])
        if g_p is not None:
            dp_t = dp_t + g_p.reshape(n_pos)
        if g_n is not None:
            dn_t = dn_t + g_n.reshape(n_pos, K)
        dh, dW = torch.ops.recguru.sampled_score_bwd(dp_t.contiguous(), dn_t.contiguous(), None, None,
                                                     h, weight, pos, neg, ctx.padding_idx)
    return dh, dW, None, None, None, None, None


torch.library.register_autograd("recguru::sampled_score_loss", _bwd_sampled_score_loss,
                                setup_context=_setup_sampled_score_loss)


def _setup_gather_rows(ctx, inputs, output):
    weight, ids, padding_idx = inputs
    ctx.save_for_backward(ids)
    ctx.V1 = weight.shape[0]
    ctx.padding_idx = padding_idx
    ctx.wref = weight.detach()


def _bwd_gather_rows(ctx, g):
    (ids,) = ctx.saved_tensors
    d = g.shape[-1]
    if _defer_sites(ctx.wref, [Site(ids.reshape(-1), g.reshape(-1, d).contiguous())]):
        return None, None, None
    dW = torch.ops.recguru.embedding_bwd(g.reshape(-1, d).contiguous(), ids.reshape(-1), None, ctx.V1,
                                         ctx.padding_idx)
    return dW, None, None


torch.library.register_autograd("recguru::gather_rows", _bwd_gather_rows, setup_context=_setup_gather_rows)
