"""Times the sampled scoring forward (K3) at cfg2 on the padded (Amazon-shaped) batch and on a
dense batch (no padding), with and without RECGURU_LOSS_SKIP_MASKED; rotating inputs > L2."""
import sys
import torch
sys.path.insert(0, ".")
from recguru_b200 import ops, synth, _lib  # noqa: E402

B, L, d, K, NI = 1024, 200, 128, 5, 100_000
dev = torch.device("cuda")
W = synth.item_table(NI, d).to(dev)
batches = []
for b in range(4):
    tb = synth.train_batch(B, L, NI, K, seed=synth.SEED_DATA + 100 * b).to(dev)
    h = synth.hidden_states(B, L, d, seed=50 + b).to(dev)
    batches.append((tb, h))
n = B * L
bufs = {"p": torch.empty(B, L, 1, device=dev), "n": torch.empty(B, L, K, device=dev), "aux": torch.empty(2, device=dev),
        "dp": torch.empty(n, device=dev), "dn": torch.empty(n, K, device=dev),
        "ws": torch.empty(max(256, int(_lib.load().recguru_sampled_workspace_bytes(n, K))), dtype=torch.uint8, device=dev)}
ones = torch.ones(B, L, device=dev)


def timeit(fn, iters=40, warm=8):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


zeros = torch.zeros(B, L, device=dev)
short = torch.zeros(B, L, device=dev)
short[:, -8:] = 1.0          # 8 live positions at the end of every sequence
lens = batches[0][0].mask.view(B, L).sum(1)
print("live positions per sequence: mean %.1f max %d; sequences with >100 live: %d" % (
    float(lens.mean()), int(lens.max()), int((lens > 100).sum())))
for name, dense in (("padded", False), ("dense", True), ("allmasked", zeros), ("last8", short)):
    for skip in (False, True):
        def step(i):
            tb, h = batches[i % 4]
            m = dense if isinstance(dense, torch.Tensor) else (ones if dense else tb.mask)
            ids_pos = tb.n_items.view(B, L, K)[:, :, 0].contiguous() if dense is True else tb.dec_out
            ops.raw_sampled_score_loss(h, W, ids_pos, tb.n_items, m, "bpr_sas", bufs=bufs, skip_masked=skip)
        print(f"fwd {name:7s} skip_masked={skip!s:5s}: {timeit(step):8.1f} us", flush=True)
