"""A few eager TrainHotPath steps on the cfg2 batch (for ncu captures of the step's kernels)."""
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
hp = TrainHotPath(W, pe, B, L, K, "bpr_sas", overlap_plan=len(sys.argv) > 1 and sys.argv[1] == "overlap")
for _ in range(4):
    hp.step(tb.dec_in, tb.dec_out, tb.n_items, h, g)
torch.cuda.synchronize()
