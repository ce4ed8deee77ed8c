"""Timing experiments for the catalog sweep kernel (env knobs are read at every launch)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recguru_b200 import ops  # noqa: E402

dev = torch.device("cuda")
d, k = 128, 10
U = 148 * 4 * 128
g = torch.Generator(device=dev); g.manual_seed(1)
users = ops.raw_cast_bf16(torch.randn(U, d, generator=g, device=dev))


def run(n_items, label, k=k, **env):
    for kk, v in env.items():
        os.environ[kk] = str(v)
    items = ops.raw_cast_bf16(torch.randn(n_items, d, generator=g, device=dev))
    target = torch.randint(1, n_items + 1, (U,), generator=g, device=dev)
    ts = ops.raw_catalog_target_scores(users, items, target, 1)
    bufs = {}
    for _ in range(2):
        ops.raw_catalog_topk_rank(users, items, target, ts, k, 1, bufs=bufs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        ops.raw_catalog_topk_rank(users, items, target, ts, k, 1, bufs=bufs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = 2.0 * U * n_items * d / (ms * 1e-3) / 1e12
    print(f"{label:40s} N={n_items:>9d} {ms:9.3f} ms  {tf:7.1f} TFLOP/s", flush=True)
    for kk in env:
        os.environ.pop(kk, None)


n_items = 1_000_000
run(n_items, "v1 top-10 + rank", RECGURU_CATALOG_V=1)
for nw in (2, 3, 4):
    run(n_items, f"v2 top-10 + rank, {nw} warps/quarter", RECGURU_CATALOG_WARPS=nw)
    run(n_items, f"v2 rank only, {nw} warps/quarter", k=0, RECGURU_CATALOG_WARPS=nw)
for nw in (2, 3, 4):
    run(10_000_000, f"v2 top-10 + rank, 10M items, {nw} warps/quarter", RECGURU_CATALOG_WARPS=nw)
run(10_000_000, "v2 rank only, 10M items, 3 warps/quarter", k=0, RECGURU_CATALOG_WARPS=3)
