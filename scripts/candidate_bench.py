"""Times the fused get_scores+rank kernel at the reference's evaluation shapes
(candidate_size 199 and 19999; config_auto4rec.py / train_auto.py:164-199) and, beside it, the
reference's own op sequence run on the same GPU with stock torch ops (embedding, cat, matmul,
double argsort).  Prints algorithmic GB/s = (ids + gathered rows + u + rank) / time."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from recguru_b200 import catalog  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def torch_ref(u, W, target, cand):
    B, C = cand.shape
    d = W.shape[1]
    n_emb = torch.nn.functional.embedding(cand.reshape(-1), W).view(-1, C, d)
    p_emb = torch.nn.functional.embedding(target, W).view(-1, d)
    ce = torch.cat([p_emb.unsqueeze(1), n_emb], dim=1)
    s = torch.matmul(u.unsqueeze(1), ce.transpose(1, 2)).squeeze(1)
    return torch.argsort(torch.argsort(-s, dim=1), dim=1)[:, 0]


out = []
for (B, C, V, d) in [(256, 199, 100_000, 128), (256, 19999, 100_000, 128), (1024, 19999, 1_000_000, 128)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    W = torch.randn(V + 2, d, device="cuda", generator=g)
    u = torch.randn(B, d, device="cuda", generator=g)
    target = torch.randint(1, V + 1, (B,), device="cuda", generator=g)
    cand = torch.randint(1, V + 1, (B, C), device="cuda", generator=g)
    ms = timeit(lambda: catalog.get_scores_rank(u, W, target, cand, want_scores=False))
    ms_t = timeit(lambda: torch_ref(u, W, target, cand), iters=5, warm=2)
    r = catalog.get_scores_rank(u, W, target, cand)[1]
    same = float((r == torch_ref(u, W, target, cand)).float().mean())
    bytes_ = B * (C + 1) * (8 + 4 * d) + B * 4 * d + B * 8
    out.append({"B": B, "C": C, "V": V, "d": d, "ms": round(ms, 4), "algorithmic_GBps": round(bytes_ / ms / 1e6, 1),
                "users_per_s": round(B / ms * 1e3), "torch_ops_ms": round(ms_t, 4), "speedup_vs_torch_ops": round(ms_t / ms, 1),
                "rank_agreement_with_torch_ops": same})
    print(json.dumps(out[-1]), flush=True)
