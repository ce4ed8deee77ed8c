#!/bin/bash
# does halving the L2 -> SM operand traffic (cluster-of-2 TMA multicast, round-1 kernel, debug build) help now?
mkdir -p gpurun_out
python -m recguru_b200.build --debug-knobs > /dev/null 2>&1; echo build rc=$?
python - <<'PY'
import os, sys, torch
sys.path.insert(0, ".")
from recguru_b200 import ops
dev = torch.device("cuda"); d = 128; U = 148 * 4 * 128
g = torch.Generator(device=dev); g.manual_seed(1)
users = ops.raw_cast_bf16(torch.randn(U, d, generator=g, device=dev))
def run(n_items, label, k=10, reps=6, **env):
    for kk, v in env.items(): os.environ[kk] = str(v)
    items = ops.raw_cast_bf16(torch.randn(n_items, d, generator=g, device=dev))
    target = torch.randint(1, n_items + 1, (U,), generator=g, device=dev)
    ts = ops.raw_catalog_target_scores(users, items, target, 1)
    bufs = {}
    for _ in range(3): ops.raw_catalog_topk_rank(users, items, target, ts, k, 1, bufs=bufs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.raw_catalog_topk_rank(users, items, target, ts, k, 1, bufs=bufs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{label:46s} N={n_items:>9d} {ms:9.3f} ms  {2.0*U*n_items*d/(ms*1e-3)/1e12:7.1f} TFLOP/s", flush=True)
    for kk in env: os.environ.pop(kk, None)
for n in (1_000_000, 10_000_000):
    r = 6 if n == 1_000_000 else 3
    run(n, "v1 cluster 1", reps=r, RECGURU_CATALOG_V=1)
    run(n, "v1 cluster 2 (TMA multicast)", reps=r, RECGURU_CATALOG_V=1, RECGURU_CATALOG_CLUSTER=2)
    run(n, "v1 A in TMEM (TS)", reps=r, RECGURU_CATALOG_V=1, RECGURU_CATALOG_TS=1)
    run(n, "v2", reps=r)
    run(n, "v1 rank-only cluster 1", k=0, reps=r, RECGURU_CATALOG_V=1)
    run(n, "v1 rank-only cluster 2", k=0, reps=r, RECGURU_CATALOG_V=1, RECGURU_CATALOG_CLUSTER=2)
PY
