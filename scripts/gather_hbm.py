"""K1 on a 512 MB table (cfg4's 1M x 128 fp32) for the ncu DRAM-traffic capture: 8 cold launches with uniform
ids (rotating id/out sets), then 8 with Zipf(1.05) ids.  Launch order is what scripts/ncu_summary2.py keys on."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recguru_b200 import ops, synth, modules  # noqa: E402
dev = torch.device("cuda")
n_items, d, B, L = 1_000_000, 128, 1024, 200
gen = torch.Generator(device=dev).manual_seed(4242)
W = torch.randn(n_items + 2, d, generator=gen, device=dev)
pe = modules.sinusoid_table(256, d)[0].to(dev)
n_sets = 8
outs = [torch.empty(B, L, d, device=dev) for _ in range(n_sets)]
masks = [torch.empty(B, L, device=dev) for _ in range(n_sets)]
ids_u = torch.randint(1, n_items + 1, (n_sets, B, L), generator=gen, device=dev)
ids_z = synth.zipf_ids(n_sets * B * L, n_items, seed=99).view(n_sets, B, L).to(dev)
torch.cuda.synchronize()
for ids in (ids_u, ids_z):
    for j in range(n_sets):
        ops.raw_gather_pe_mask(W, ids[j], pe, 0, out=outs[j], mask_out=masks[j])
torch.cuda.synchronize()
