"""Device-resident sequence store and on-GPU batch builder.

The reference's loaders (``pickle_loader`` / ``pickle_loader_eval``, data/data_loader.py:203-316,
:58-135) keep the ragged item histories of the pickles (``{"seq", "len", "val", "test"}``,
data/amazon_csv.py:251-280) in Python lists and build every sample on the host:
``seq_padding`` (:25-36) or ``test_seq_gen`` (:39-55) per sequence, ``default_collate`` per batch,
then 3 x [B, L] int64 over PCIe per step -- 4.9 MB at cfg2 for about twelve real items per row.

Here the dataset sits in HBM in ragged form (an Amazon 5-core domain is a few MB; 180 GB holds
billions of interactions) and a step ships only the B user indices of its batch: :class:`SequenceStore`
packs the lists once, :meth:`SequenceStore.train_batch` / :meth:`eval_batch` run
``recguru_build_batches`` and hand back exactly the tensors ``default_collate`` would have
produced (int64, same layout, same values -- ``tests/test_gpu_batch_builder.py`` checks them
bit for bit against the reference functions), on the device.  Together with
:class:`recguru_b200.sampler.NegativeSampler` no per-sample host work is left.

There is no CPU path: the store refuses CPU devices and raises if the CUDA library is missing.
"""
from __future__ import annotations

import itertools
from typing import NamedTuple, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .ops import _need, _ptr, _stream, check

Tensor = torch.Tensor
BATCH_TRAIN, BATCH_EVAL = 0, 1


def pack_sequences(seqs: Sequence[Sequence[int]]):
    """Ragged lists -> (items int64[sum len], offsets int64[n+1]) numpy arrays (host, once)."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    # one pass over all items (C loop inside numpy) instead of a slice assignment per sequence
    items = np.fromiter(itertools.chain.from_iterable(seqs), dtype=np.int64, count=int(offsets[-1]))
    return items, offsets


def raw_build_batches(items: Tensor, offsets: Tensor, user_idx: Optional[Tensor], append: Optional[Tensor],
                      B: int, enc_len: int, dec_len: int, eos: int, mode: int, enc_in: Optional[Tensor],
                      dec_in: Optional[Tensor], dec_out: Optional[Tensor], val: Optional[Tensor] = None,
                      test: Optional[Tensor] = None, val_out: Optional[Tensor] = None,
                      test_out: Optional[Tensor] = None) -> None:
    """Thin checked call of ``recguru_build_batches`` (include/recguru_b200.h) on the current stream."""
    items = _need(items, torch.int64, "items")
    offsets = _need(offsets, torch.int64, "offsets")
    n_users = offsets.numel() - 1
    if user_idx is not None:
        user_idx = _need(user_idx, torch.int64, "user_idx")
        if user_idx.numel() != B:
            raise RuntimeError("recguru: user_idx must hold B entries")
    elif B > n_users:
        raise RuntimeError(f"recguru: B={B} exceeds the {n_users} stored sequences")
    for name, t, shape in (("enc_in", enc_in, (B, enc_len)), ("dec_in", dec_in, (B, dec_len)),
                           ("dec_out", dec_out, (B, dec_len)), ("val_out", val_out, (B,)),
                           ("test_out", test_out, (B,))):
        if t is not None:
            if not (t.is_cuda and t.dtype == torch.int64 and t.is_contiguous() and tuple(t.shape) == shape):
                raise RuntimeError(f"recguru: {name} must be a contiguous CUDA int64 tensor of shape {shape}")
    for name, t in (("append", append), ("val", val), ("test", test)):
        if t is not None:
            _need(t, torch.int64, name)
            if not t.is_contiguous() or t.numel() != n_users:
                raise RuntimeError(f"recguru: {name} must hold one contiguous entry per stored sequence")
    check(_lib.load().recguru_build_batches(_ptr(items), _ptr(offsets), _ptr(user_idx), _ptr(append), B, enc_len,
                                            dec_len, eos, mode, _ptr(enc_in), _ptr(dec_in), _ptr(dec_out),
                                            _ptr(val), _ptr(test), _ptr(val_out), _ptr(test_out), _stream()),
          "build_batches")


class TrainRows(NamedTuple):
    """What ``pickle_loader.__getitem__`` + ``default_collate`` yield, minus the negatives."""
    enc_in: Tensor
    dec_in: Tensor
    dec_out: Tensor
    val: Optional[Tensor]
    test: Optional[Tensor]


class EvalRows(NamedTuple):
    """``pickle_loader_eval``'s ``(eval_enc_in, dec_in, val), (test_enc_in, t_dec_in, test)``."""
    eval_enc_in: Tensor
    dec_in: Tensor
    val: Tensor
    test_enc_in: Tensor
    t_dec_in: Tensor
    test: Tensor


class SequenceStore:
    """The ``seq`` / ``val`` / ``test`` lists of one domain, resident on the device.

    ``eos`` is the loader's ``vocab_size_a/_b`` (data_loader.py:72-75); ``enc_len`` / ``dec_len`` its
    ``enc_maxlen`` / ``rec_maxlen``."""

    def __init__(self, seqs: Sequence[Sequence[int]], val: Optional[Sequence[int]], test: Optional[Sequence[int]],
                 eos: int, enc_len: int, dec_len: Optional[int] = None, device="cuda"):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("recguru: SequenceStore lives on a CUDA device (no CPU path exists)")
        _lib.load()
        items, offsets = pack_sequences(seqs)
        self.n_users = len(seqs)
        self.max_len = int(np.diff(offsets).max()) if len(seqs) else 0
        if self.max_len > 12000:
            raise RuntimeError("recguru: sequences longer than 12000 items are not supported by the device sampler")
        self.items = torch.from_numpy(items).to(device)
        self.offsets = torch.from_numpy(offsets).to(device)
        self.val = None if val is None else torch.as_tensor(np.asarray(val, dtype=np.int64)).to(device)
        self.test = None if test is None else torch.as_tensor(np.asarray(test, dtype=np.int64)).to(device)
        for t in (self.val, self.test):
            if t is not None and t.numel() != self.n_users:
                raise ValueError("val / test need one item per sequence")
        self.eos, self.enc_len = int(eos), int(enc_len)
        self.dec_len = self.enc_len if dec_len is None else int(dec_len)
        self.device = device

    def __len__(self) -> int:
        return self.n_users

    def _rows(self, B: int, n: int):
        return torch.empty(B, n, dtype=torch.int64, device=self.device)

    def train_batch(self, user_idx: Optional[Tensor], B: Optional[int] = None,
                    out: Optional[TrainRows] = None) -> TrainRows:
        """``seq_padding`` rows of the users in ``user_idx`` (device int64 [B]; ``None`` = the first
        ``B`` users).  ``out`` = a previous result to overwrite in place (static buffers for CUDA graphs)."""
        B = int(user_idx.numel()) if user_idx is not None else int(self.n_users if B is None else B)
        if out is None:
            vec = lambda src: None if src is None else torch.empty(B, dtype=torch.int64, device=self.device)
            out = TrainRows(self._rows(B, self.enc_len), self._rows(B, self.dec_len), self._rows(B, self.dec_len),
                            vec(self.val), vec(self.test))
        raw_build_batches(self.items, self.offsets, user_idx, None, B, self.enc_len, self.dec_len, self.eos,
                          BATCH_TRAIN, out.enc_in, out.dec_in, out.dec_out, self.val, self.test, out.val, out.test)
        return out

    def eval_batch(self, user_idx: Optional[Tensor], B: Optional[int] = None) -> EvalRows:
        """``test_seq_gen`` rows: the validation layout, and the test layout with ``val`` appended."""
        if self.val is None or self.test is None:
            raise RuntimeError("recguru: eval_batch needs the val and test items")
        B = int(user_idx.numel()) if user_idx is not None else int(self.n_users if B is None else B)
        ev_enc, ev_dec = self._rows(B, self.enc_len), self._rows(B, self.dec_len)
        te_enc, te_dec = self._rows(B, self.enc_len), self._rows(B, self.dec_len)
        v = torch.empty(B, dtype=torch.int64, device=self.device)
        t = torch.empty(B, dtype=torch.int64, device=self.device)
        raw_build_batches(self.items, self.offsets, user_idx, None, B, self.enc_len, self.dec_len, self.eos,
                          BATCH_EVAL, ev_enc, ev_dec, None, self.val, self.test, v, t)
        raw_build_batches(self.items, self.offsets, user_idx, self.val, B, self.enc_len, self.dec_len, self.eos,
                          BATCH_EVAL, te_enc, te_dec, None)
        return EvalRows(ev_enc, ev_dec, v, te_enc, te_dec, t)
