"""Stand-alone duration of the two padding-aware scoring kernels on the cfg2 batch (compacting per CTA vs
walking the plan's live-row list), then the graph-replayed step."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recguru_b200 import synth, modules, ops  # noqa: E402
from recguru_b200.hotpath import TrainHotPath  # noqa: E402
B, L, n_items, K, d = 1024, 200, 100_000, 5, 128
dev = torch.device("cuda")
W = synth.item_table(n_items, d).to(dev)
pe = modules.sinusoid_table(256, d)[0].to(dev)
tb = synth.train_batch(B, L, n_items, K).to(dev)
h = synth.hidden_states(B, L, d, seed=50).to(dev)
g = synth.hidden_states(B, L, d, seed=90).to(dev)
hp = TrainHotPath(W, pe, B, L, K, "bpr_sas")
n = B * L
mask = torch.empty(n, device=dev)
cnt = torch.zeros(3, dtype=torch.int32, device=dev)
ops.raw_pad_mask_count(tb.dec_in, 0, mask.view(B, L), cnt)
sites = [ops.Site(tb.dec_in, None, live=mask), ops.Site(tb.dec_out, None, live=mask),
         ops.Site(tb.n_items, None, group=K, live=mask)]
ops.raw_embedding_bwd_plan(sites, hp.V1, 0, hp.ws_bwd, stage=ops.PLAN_COMPACT)
rows = ops.raw_embedding_bwd_plan_live_rows(sites, hp.V1, hp.ws_bwd)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=30):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def score(rows_):
    ops.raw_sampled_score_loss(h, W, tb.dec_out, tb.n_items, mask, "bpr_sas", bufs=hp.sbuf, skip_masked=True,
                               dh_out=hp.dh, live_count=cnt, live_rows=rows_)


for _ in range(3):
    score(None); score(rows)
print(f"compacting kernel + finalize, cold L2: {timed(lambda: score(None)):.1f} us")
print(f"live-row-list kernel + finalize, cold L2: {timed(lambda: score(rows)):.1f} us")
graph = hp.capture_step(tb.dec_in, tb.dec_out, tb.n_items, h, g)
for _ in range(5):
    graph.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200):
    graph.replay()
b.record(); torch.cuda.synchronize()
print(f"step (graph replay, same batch): {a.elapsed_time(b) / 200 * 1e3:.1f} us")
