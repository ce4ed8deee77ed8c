"""Timeline of ONE CUDA-graph replay of the cfg3 (RecGURU, two domains) step and of one domain alone:
timestamp kernels at the end of every branch of each domain's TrainStep."""
import statistics
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recguru_b200 import synth, modules  # noqa: E402
from recguru_b200.hotpath import RecGuruStep  # noqa: E402
B, L, n_items, K, d = 1024, 100, 500_000, 30, 128
dev = torch.device("cuda")
Wa = synth.item_table(n_items, d, seed=synth.SEED_WEIGHTS).to(dev)
Wb = synth.item_table(n_items, d, seed=synth.SEED_WEIGHTS + 1).to(dev)
pe = modules.sinusoid_table(256, d)[0].to(dev)
ta = synth.train_batch(B, L, n_items, K, seed=synth.SEED_DATA).to(dev)
tb = synth.train_batch(B, L, n_items, K, seed=synth.SEED_DATA + 5000).to(dev)
ha, hb = synth.hidden_states(B, L, d, seed=1).to(dev), synth.hidden_states(B, L, d, seed=2).to(dev)
ga = [(synth.hidden_states(B, L, d, seed=10 + i) * 0.01).to(dev) for i in range(2)]
gb = [(synth.hidden_states(B, L, d, seed=20 + i) * 0.01).to(dev) for i in range(2)]
print("live positions (dec_out != pad):", int((ta.dec_out != 0).sum()), "of", B * L,
      "| enc_in != pad:", int((ta.enc_in != 0).sum()))
rg = RecGuruStep(Wa, Wb, pe, B, L, K, "sce")
traces = [st.enable_trace() for st in rg.dom]
graph = rg.capture_run(ta, tb, ha, hb, ga, gb)
rows = [[], []]
for _ in range(30):
    graph.replay()
    torch.cuda.synchronize()
    t0 = min(int(traces[0][0]), int(traces[1][0]))
    for k in range(2):
        t = traces[k].cpu().tolist()
        rows[k].append([(x - t0) / 1e3 if x else float("nan") for x in t])
for k in range(2):
    print(f"domain {k}: median time since step start (us)")
    for i, name in enumerate(rg.dom[k].TRACE_POINTS):
        col = [r[i] for r in rows[k][5:]]
        print(f"  {name:14s} {statistics.median(col):8.1f}")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100):
    graph.replay()
b.record(); torch.cuda.synchronize()
print(f"two-domain step (graph replay, same batch): {a.elapsed_time(b) / 100 * 1e3:.1f} us")
