"""On-GPU negative sampler.

The reference draws the L*K negatives of every training sample on the CPU inside
``pickle_loader.__getitem__`` (data/data_loader.py:289-314): a ``np.ones(v_size)`` (or
``item_freq ** 0.75``, :251-254) weight vector per sample, the user's sequence / val / test item
zeroed, then ``torch.multinomial(weights, enc_maxlen * num_n, replacement=True)`` -- O(v_size)
host work per sequence and 8 MB of H2D traffic per 1024-sequence batch at cfg2.
:class:`NegativeSampler` draws from the same distribution on the device (rejection sampling,
counter-based Philox: reproducible from ``(seed, step)``); only the id sequences cross PCIe.

The random stream differs from ``torch.multinomial``'s, the distribution does not
(``tests/test_gpu_sampler.py``: exclusions never drawn, chi-square against the reference weights).

Two quirks of the reference loader decide what "the reference's distribution" is (pinned by
``tests/golden/sampler_weights.npz``, generated from ``pickle_loader.__getitem__`` itself):

* ``for val in seq: weights[seq] = 0`` (data_loader.py:291-292,:303-304) REBINDS ``val`` to the last
  history item, so ``weights[val] = 0`` zeroes ``seq[-1]`` a second time and the user's true
  validation item stays a legal negative.  ``exclude_val=True``
  (default) fixes that -- the validation item is excluded, as the code plainly intends;
  ``exclude_val=False`` reproduces the reference's law exactly.
* with ``item_freq`` the reference does not zero ``weights[0]`` (:289-295 vs :301-302): the pad id can be
  drawn if ``wf[0] > 0``.  Here id 0 is never a negative; pass frequencies with ``wf[0] == 0`` (item counts
  never include the pad id) and the two laws coincide.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .ops import _need, _ptr, _stream, check

Tensor = torch.Tensor


def build_alias(weights: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Walker/Vose alias table of a non-negative weight vector: (prob float32[n], alias int32[n])."""
    w = np.asarray(weights, dtype=np.float64)
    n = w.size
    if n == 0 or w.min() < 0 or w.sum() <= 0:
        raise ValueError("weights must be non-negative with a positive sum")
    p = w * (n / w.sum())
    prob = np.ones(n, dtype=np.float64)
    alias = np.arange(n, dtype=np.int64)
    small = [i for i in range(n) if p[i] < 1.0]
    large = [i for i in range(n) if p[i] >= 1.0]
    while small and large:
        s, l = small.pop(), large.pop()
        prob[s], alias[s] = p[s], l
        p[l] = p[l] + p[s] - 1.0
        (small if p[l] < 1.0 else large).append(l)
    return prob.astype(np.float32), alias.astype(np.int32)


class NegativeSampler:
    """``sample(history, val, test)`` -> ``n_items [B, num_n]`` int64 on the device.

    ``v_size``: the loader's ``vocab_size_a/_b`` (ids ``1 .. v_size-1`` are candidates; 0 is the pad
    id, ``v_size`` the EOS id).  ``item_freq``: the ``wf`` vector of the reference (raised to 0.75
    here, data_loader.py:251-254); ``None`` = uniform."""

    def __init__(self, v_size: int, num_n: int, seed: int = 0, item_freq: Optional[np.ndarray] = None,
                 device="cuda", exclude_val: bool = True):
        self.v_size, self.num_n, self.seed, self.step = int(v_size), int(num_n), int(seed), 0
        self.exclude_val = bool(exclude_val)       # False: the reference's law (its val exclusion is a no-op)
        self.device = torch.device(device)
        self.alias_prob = self.alias_idx = None
        if item_freq is not None:
            w = np.power(np.asarray(item_freq, dtype=np.float64)[:self.v_size], 0.75)
            w[0] = 0.0
            prob, alias = build_alias(w)
            self.alias_prob = torch.from_numpy(prob).to(self.device)
            self.alias_idx = torch.from_numpy(alias).to(self.device)

    def sample(self, history: Tensor, val: Optional[Tensor] = None, test: Optional[Tensor] = None,
               pad_index: int = 0, step: Optional[int] = None, out: Optional[Tensor] = None,
               step_dev: Optional[Tensor] = None) -> Tensor:
        """``history [B, H]``: the ids to exclude per sequence (the padded sequence / ``enc_in``;
        entries equal to ``pad_index`` or outside ``[1, v_size)`` -- the EOS id -- are ignored).
        ``step_dev``: a device int64 ``[1]`` step counter read at kernel time instead of ``step``
        (CUDA graphs: bump it inside the graph with :meth:`advance`)."""
        history = _need(history, torch.int64, "history")
        B, H = history.shape
        if not self.exclude_val:
            val = None
        val = None if val is None else _need(val.view(-1), torch.int64, "val")
        test = None if test is None else _need(test.view(-1), torch.int64, "test")
        if out is None:
            out = torch.empty(B, self.num_n, dtype=torch.int64, device=history.device)
        if step_dev is not None:
            step_dev = _need(step_dev, torch.int64, "step_dev")
            check(_lib.load().recguru_sample_negatives_devstep(
                _ptr(history), B, H, int(pad_index), _ptr(val), _ptr(test), self.v_size, _ptr(self.alias_prob),
                _ptr(self.alias_idx), self.num_n, self.seed, _ptr(step_dev), _ptr(out), _stream()),
                "sample_negatives_devstep")
            return out
        if step is None:
            step = self.step
            self.step += 1
        check(_lib.load().recguru_sample_negatives(_ptr(history), B, H, int(pad_index), _ptr(val), _ptr(test),
                                                   self.v_size, _ptr(self.alias_prob), _ptr(self.alias_idx),
                                                   self.num_n, self.seed, int(step), _ptr(out), _stream()),
              "sample_negatives")
        return out

    def sample_store(self, store, user_idx: Optional[Tensor], B: Optional[int] = None, step: Optional[int] = None,
                     out: Optional[Tensor] = None, step_dev: Optional[Tensor] = None) -> Tensor:
        """Negatives for the users ``user_idx`` of a :class:`recguru_b200.loader.SequenceStore`: the
        exclusion list is the user's WHOLE stored sequence plus ``store.val`` / ``store.test`` -- what
        the reference zeroes (data_loader.py:291-295) -- not only the window that fits ``enc_in``."""
        if user_idx is not None:
            user_idx = _need(user_idx, torch.int64, "user_idx")
            B = int(user_idx.numel())
        else:
            B = int(len(store) if B is None else B)
            if B > len(store):
                raise RuntimeError(f"recguru: B={B} exceeds the {len(store)} stored sequences")
        if out is None:
            out = torch.empty(B, self.num_n, dtype=torch.int64, device=store.items.device)
        if step_dev is not None:
            step_dev = _need(step_dev, torch.int64, "step_dev")
            step = 0
        elif step is None:
            step = self.step
            self.step += 1
        check(_lib.load().recguru_sample_negatives_csr(
            _ptr(store.items), _ptr(store.offsets), _ptr(user_idx), B, max(int(store.max_len), 1),
            _ptr(store.val if self.exclude_val else None),
            _ptr(store.test), self.v_size, _ptr(self.alias_prob), _ptr(self.alias_idx), self.num_n, self.seed,
            int(step), _ptr(step_dev), _ptr(out), _stream()), "sample_negatives_csr")
        return out

    @staticmethod
    def advance(step_dev: Tensor, inc: int = 1) -> None:
        """``step_dev += inc`` on the current stream (one-thread kernel; capturable)."""
        step_dev = _need(step_dev, torch.int64, "step_dev")
        check(_lib.load().recguru_counter_add(_ptr(step_dev), int(inc), _stream()), "counter_add")
