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


# lower bound of the overlapped schedule: the plan is not launched at all (apply re-uses the plan a
# previous step left in the workspace: timing only)
class NoPlan(TrainHotPath):
    def plan_backward(self, dec_in, dec_out, n_items, mask=None):
        cur = torch.cuda.current_stream()
        self._fork.record(cur)
        self._side.wait_event(self._fork)
        with torch.cuda.stream(self._side):
            ops.raw_fill_zero(self.dW)
            self._join.record(self._side)
        self._planned = True


hp = NoPlan(W, pe, B, L, K, "bpr_sas")
TrainHotPath.step(hp, batches[0][0].dec_in, batches[0][0].dec_out, batches[0][0].n_items, batches[0][1], batches[0][2])
tb, h, g = batches[0]
TrainHotPath.plan_backward(hp, tb.dec_in, tb.dec_out, tb.n_items); torch.cuda.synchronize()
graphs = [hp.capture_step(tb.dec_in, tb.dec_out, tb.n_items, h, g) for _ in range(4)]
print(f"{'no plan launched (lower bound)':36s} {timeit(graphs):8.1f} us", flush=True)
