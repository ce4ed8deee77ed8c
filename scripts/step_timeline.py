"""Timeline of ONE CUDA-graph replay of the cfg2 training step: timestamp kernels at the end of every
branch (main: K1, scoring, apply; side: pad mask, plan; zero fills) -> which branch is critical."""
import statistics
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recguru_b200 import synth, modules  # noqa: E402
from recguru_b200.hotpath import TrainHotPath  # noqa: E402
B, L, n_items, K, d = 1024, 200, 100_000, 5, 128
dev = torch.device("cuda")
W = synth.item_table(n_items, d).to(dev)
pe = modules.sinusoid_table(256, d)[0].to(dev)
tb = synth.train_batch(B, L, n_items, K).to(dev)
h = synth.hidden_states(B, L, d, seed=50).to(dev)
g = synth.hidden_states(B, L, d, seed=90).to(dev)
hp = TrainHotPath(W, pe, B, L, K, "bpr_sas")
tr = hp.enable_trace()
graph = hp.capture_step(tb.dec_in, tb.dec_out, tb.n_items, h, g)
rows = []
for _ in range(30):
    graph.replay()
    torch.cuda.synchronize()
    t = tr.cpu().tolist()
    rows.append([(x - t[0]) / 1e3 if x else float("nan") for x in t])
print("median time since step start (us) at which each point was reached, 30 graph replays:")
for i, name in enumerate(hp.TRACE_POINTS):
    col = [r[i] for r in rows[5:]]
    print(f"  {name:14s} {statistics.median(col):8.1f}   (min {min(col):7.1f} max {max(col):7.1f})")
