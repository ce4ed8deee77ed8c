"""Graph-replay time of the training-step edges under different schedules of the table-backward plan."""
import sys, torch
sys.path.insert(0, ".")
from recguru_b200 import synth, modules, ops
from recguru_b200.hotpath import TrainHotPath

B, L, d, K, NI = 1024, 200, 128, 5, 100_000
dev = torch.device("cuda")
W = synth.item_table(NI, d).to(dev)
pe = modules.sinusoid_table(256, d)[0].to(dev)
batches = []
for b in range(4):
    tb = synth.train_batch(B, L, NI, K, seed=synth.SEED_DATA + 100 * b).to(dev)
    batches.append((tb, synth.hidden_states(B, L, d, seed=50 + b).to(dev), synth.hidden_states(B, L, d, seed=90 + b).to(dev)))


def timeit(graphs, iters=40, warm=8):
    for i in range(warm):
        graphs[i % 4].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        graphs[i % 4].replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for name, kw in (("plan on side stream (default)", dict(overlap_plan=True)), ("everything serial", dict(overlap_plan=False))):
    hp = TrainHotPath(W, pe, B, L, K, "bpr_sas", **kw)
    graphs = [hp.capture_step(tb.dec_in, tb.dec_out, tb.n_items, h, g) for tb, h, g in batches]
    print(f"{name:36s} {timeit(graphs):8.1f} us", flush=True)
