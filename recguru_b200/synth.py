"""Synthetic Amazon-shaped inputs for the hot path.

Zipf(alpha=1.05) item popularity over a permuted id space, log-normal history lengths,
the reference's own padded-sequence layout (data/data_loader.py:25-36: left-pad 0, EOS = N+1
last, dec_in / dec_out = enc_in shifted by two / one), uniform negatives
(data/data_loader.py:304-314), N(0,1) fp32 tables with a zero pad row, and the grid-valued
table (k/16) on which every bf16 x bf16 -> fp32 score is exact in any summation order.
Everything is generated on the CPU from fixed seeds so the CPU baseline and the GPU see the
same bytes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

SEED_DATA, SEED_WEIGHTS, SEED_NEG = 1234, 1235, 1236
ZIPF_ALPHA = 1.05


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def zipf_ids(n: int, n_items: int, seed: int = SEED_DATA, alpha: float = ZIPF_ALPHA) -> torch.Tensor:
    """n item ids in 1..n_items, popularity rank^-alpha assigned to a random permutation of ids."""
    g = _gen(seed)
    w = torch.arange(1, n_items + 1, dtype=torch.float64).pow(-alpha)
    perm = torch.randperm(n_items, generator=g)
    # multinomial over >2^24 categories is not supported: sample by inverse CDF instead
    cdf = torch.cumsum(w / w.sum(), 0)
    u = torch.rand(n, generator=g, dtype=torch.float64)
    r = torch.searchsorted(cdf, u).clamp_(max=n_items - 1)
    return perm[r] + 1


def item_table(n_items: int, d: int, seed: int = SEED_WEIGHTS, grid: bool = False) -> torch.Tensor:
    """[n_items + 2, d] fp32: row 0 = pad (zeros), 1..n_items items, n_items+1 = EOS."""
    g = _gen(seed)
    if grid:
        w = torch.randint(-32, 33, (n_items + 2, d), generator=g).to(torch.float32) / 16.0
    else:
        w = torch.randn(n_items + 2, d, generator=g)
    w[0].zero_()
    return w


def user_reps(n_users: int, d: int, seed: int = SEED_DATA + 7, grid: bool = False) -> torch.Tensor:
    g = _gen(seed)
    if grid:
        return torch.randint(-32, 33, (n_users, d), generator=g).to(torch.float32) / 16.0
    return torch.randn(n_users, d, generator=g)


@dataclass
class TrainBatch:
    enc_in: torch.Tensor    # [B, L] int64
    dec_in: torch.Tensor    # [B, L] int64
    dec_out: torch.Tensor   # [B, L] int64
    n_items: torch.Tensor   # [B, L*K] int64
    mask: torch.Tensor      # [B*L] fp32, 1 - (dec_in == 0)   (train_auto.py:109-110)
    n_items_total: int
    L: int
    K: int

    def to(self, device) -> "TrainBatch":
        return TrainBatch(self.enc_in.to(device), self.dec_in.to(device), self.dec_out.to(device),
                          self.n_items.to(device), self.mask.to(device), self.n_items_total, self.L, self.K)


def train_batch(B: int, L: int, n_items: int, K: int, seed: int = SEED_DATA) -> TrainBatch:
    g = _gen(seed)
    # history length ~ clip(3 + floor(LogNormal(ln 8, 1)), 3, L-1)
    ln = torch.empty(B, dtype=torch.float64).log_normal_(mean=2.0794415416798357, std=1.0, generator=g)
    length = (3 + ln.floor()).clamp_(3, L - 1).to(torch.int64)
    items = zipf_ids(B * L, n_items, seed=seed + 1).view(B, L)
    j = torch.arange(L).unsqueeze(0)
    start = (L - 1 - length).unsqueeze(1)          # first history slot in enc_in
    enc_in = torch.where((j >= start) & (j <= L - 2), items, torch.zeros_like(items))
    enc_in[:, L - 1] = n_items + 1                 # EOS
    zeros2 = torch.zeros(B, 2, dtype=torch.int64)
    dec_in = torch.cat([zeros2, enc_in[:, :-2]], dim=1)[:, -L:]
    dec_out = torch.cat([zeros2, enc_in[:, 1:-1]], dim=1)[:, -L:]
    gn = _gen(SEED_NEG if seed == SEED_DATA else seed + 2)
    neg = torch.randint(1, n_items + 1, (B, L * K), generator=gn)
    mask = (dec_in != 0).to(torch.float32).view(-1)
    return TrainBatch(enc_in, dec_in, dec_out, neg, mask, n_items, L, K)


def hidden_states(B: int, L: int, d: int, seed: int = SEED_DATA + 11) -> torch.Tensor:
    return torch.randn(B, L, d, generator=_gen(seed))
