"""Single catalog sweep launch (for ncu source-level profiling); env knobs as in catalog_exp.py."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recguru_b200 import ops  # noqa: E402
dev = torch.device("cuda"); d, k = 128, 10
U = 148 * 4 * 128; n_items = 1_000_000
g = torch.Generator(device=dev); g.manual_seed(1)
users = ops.raw_cast_bf16(torch.randn(U, d, generator=g, device=dev))
items = ops.raw_cast_bf16(torch.randn(n_items, d, generator=g, device=dev))
target = torch.randint(1, n_items + 1, (U,), generator=g, device=dev)
ts = ops.raw_catalog_target_scores(users, items, target, 1)
for _ in range(2):
    ops.raw_catalog_topk_rank(users, items, target, ts, k, 1)
torch.cuda.synchronize()
