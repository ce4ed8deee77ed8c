"""The table-backward plan of the cfg2 step on its own (for ncu launch lists): RECGURU_BWD_PLAN=1 = round-1 chain."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recguru_b200 import ops, synth  # noqa: E402
B, L, n_items, K, d = 1024, 200, 100_000, 5, 128
dense = len(sys.argv) > 1 and sys.argv[1] == "dense"
tb = synth.train_batch(B, L, n_items, K).to("cuda")
if dense:
    g = torch.Generator().manual_seed(3)
    tb.dec_in = torch.randint(1, n_items + 1, (B, L), generator=g).cuda()
    tb.dec_out = torch.randint(1, n_items + 1, (B, L), generator=g).cuda()
n = B * L
mask = torch.empty(n, device="cuda")
cnt = torch.zeros(3, dtype=torch.int32, device="cuda")
V1 = n_items + 2
ws = torch.empty(ops.embedding_bwd_workspace_bytes(n * (2 + K), V1, d), dtype=torch.uint8, device="cuda")
sites = [ops.Site(tb.dec_in, None, live=mask), ops.Site(tb.dec_out, None, live=mask),
         ops.Site(tb.n_items, None, group=K, live=mask)]
for _ in range(3):
    ops.raw_pad_mask_count(tb.dec_in, 0, mask, cnt)
    ops.raw_embedding_bwd_plan(sites, V1, 0, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.raw_pad_mask_count(tb.dec_in, 0, mask, cnt)
    ops.raw_embedding_bwd_plan(sites, V1, 0, ws)
e1.record(); torch.cuda.synchronize()
print("plan (+pad mask) us:", e0.elapsed_time(e1) / 20 * 1e3, "live", int(mask.sum()))
