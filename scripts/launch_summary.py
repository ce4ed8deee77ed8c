"""Per-kernel device time from an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list.
usage: python scripts/launch_summary.py gpurun_out/r01_launches.csv > profiles/r01_launch_summary.txt
Only our kernels (namespace rg::) are listed; share_of_train_step = avg time x launches per step of the
kernels TrainHotPath.step launches, over their sum (cold-cache, serialised: compare SHARES, not absolutes)."""
import csv
import sys
from collections import OrderedDict

# launches per TrainHotPath.step (fused forward+dh, dense gradient, plan by row, 17-bit keys = 2 radix passes)
PER_STEP = OrderedDict([
    ("pad_mask_count_kernel", 1), ("zero_fill_kernel", 1), ("row_count_kernel", 1), ("compact_scan_kernel", 1),
    ("row_scatter_kernel", 1), ("radix_hist_kernel", 2), ("radix_scan_rows_kernel", 2), ("radix_scatter_kernel", 2),
    ("gather_pe_mask_kernel", 1), ("sampled_score_rows_kernel", 1), ("loss_finalize_kernel", 1),
    ("chunk_reduce_tma_kernel", 1), ("span_fixup_kernel", 1)])

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= iv or "rg::" not in r[ik]:
        continue
    name = r[ik].split("(")[0].replace("void ", "")
    agg.setdefault(name, []).append(float(r[iv].replace(",", "")) / 1e3)
step_us = 0.0
avg = {}
for name, v in agg.items():
    avg[name] = sum(v) / len(v)
    base = name.replace("rg::", "").split("<")[0]
    if base in PER_STEP:
        step_us += avg[name] * PER_STEP[base]
print("# per-kernel device time, ncu --metrics gpu__time_duration.sum --clock-control none, "
      "bench.py --steps 3 --warmup 3 --only train,cfg4 --no-graph")
print("# (cold-cache, serialised: compare SHARES, not absolutes; eager launches so that every kernel is listed;")
print("#  n counts every launch of the run incl. the stage-breakdown, eager, e2e and device-batch loops;")
print(f"#  train step = {step_us:.1f} us summed over the kernels of one TrainHotPath.step)")
for name, v in agg.items():
    base = name.replace("rg::", "").split("<")[0]
    share = f"{100 * avg[name] * PER_STEP[base] / step_us:5.1f}%" if base in PER_STEP else "  -  "
    print(f"{name:76s} n={len(v):3d} avg={avg[name]:10.2f} us  share_of_train_step={share}")
