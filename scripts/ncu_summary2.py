"""Summarise .ncu-rep captures (ncu --set full) into text + the traffic table bench.py reads.
usage: python scripts/ncu_summary2.py OUT_JSON <rep> [<rep> ...] > summary.txt
A rep whose name contains 'gather_hbm' holds 16 launches of gather_pe_mask_kernel: 8 uniform-id, then 8 Zipf-id
(scripts/gather_hbm.py); they are keyed gather_hbm_uniform_cold / gather_hbm_zipf_cold."""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg.per_second"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


out_json = sys.argv[1]
try:
    traffic = json.load(open(out_json), object_pairs_hook=OrderedDict)
except (OSError, ValueError):
    traffic = OrderedDict()
for rep in sys.argv[2:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"# {rep}: empty")
        continue
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    agg = OrderedDict()
    for i, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("rg::", "")
        if "gather_hbm" in rep:
            name = "gather_hbm_uniform_cold" if i < 8 else "gather_hbm_zipf_cold"
        agg.setdefault(name, []).append(r)
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    for name, rs in agg.items():
        print(f"{name}  (n={len(rs)})")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals = [float(r[i].replace(",", "")) for r in rs if r[i] not in ("", "n/a")]
                if vals:
                    print(f"    {k:85s} {sum(vals) / len(vals):16.3f} {units[i]}")
        t = sum(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in rs) / len(rs)
        traffic[name] = {"dram_bytes_per_launch": t, "launches_captured": len(rs),
                         "us_under_ncu": sum(float(r[it].replace(",", "")) for r in rs) / len(rs)
                         * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[it], 1.0), "source": rep}
json.dump(traffic, open(out_json, "w"), indent=1)
