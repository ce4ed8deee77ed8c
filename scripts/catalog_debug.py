"""Diagnostics for the tcgen05 path: prints how the dense scores differ from the oracle so a
descriptor / layout mistake can be read off the pattern."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import ref_path as O  # noqa: E402
from recguru_b200 import ops  # noqa: E402


def grid(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-32, 33, shape, generator=g).float() / 16.0


for d in (128, 64):
    for U, N in ((128, 128), (200, 1000)):
        users, items = grid((U, d), 1), grid((N, d), 2)
        ref = O.catalog_scores(users, items)
        out = ops.raw_catalog_scores_dense(users.to(torch.bfloat16).cuda(), items.to(torch.bfloat16).cuda()).cpu()
        torch.cuda.synchronize()
        eq = (out == ref)
        print(f"d={d} U={U} N={N}: equal {eq.float().mean():.4f} maxerr {(out - ref).abs().max():.4f}")
        if not bool(eq.all()):
            print(" row-wise equal frac (first 16 rows):", [round(float(x), 2) for x in eq.float().mean(1)[:16]])
            print(" col-wise equal frac (first 16 cols):", [round(float(x), 2) for x in eq.float().mean(0)[:16]])
            print(" out[0,:8]", out[0, :8].tolist())
            print(" ref[0,:8]", ref[0, :8].tolist())
            # is out a K-truncated product?  try partial K sums
            for kk in (16, 32, 64, 96):
                if kk <= d:
                    part = O.catalog_scores(users[:, :kk], items[:, :kk])
                    print(f"  matches first-{kk}-of-K partial product: {(out == part).float().mean():.4f}")
            # transposed?
            if U == N:
                print("  matches transpose:", float((out == ref.t()).float().mean()))
print("dense debug done")
