"""Dynamic instruction count per warp-tile by opcode class + top stall lines, from `ncu --page source --csv`."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
wt = float(sys.argv[2]) if len(sys.argv) > 2 else 592 * 7813 * 4
hdr = rows[1]
iS, iI, isrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
data = [(n, int(r[iS] or 0), int(r[iI]), r[isrc].strip()) for n, r in enumerate(rows[2:]) if len(r) > iI and r[iI].isdigit()]
tot_e, tot_s = sum(d[2] for d in data), sum(d[1] for d in data)
print(f"static instrs {len(data)}  executed/warp-tile {tot_e / wt:.1f}  samples {tot_s}")
ops, smp = Counter(), Counter()
for n, s, e, src in data:
    op = src.split()[1] if src.startswith('@') else src.split()[0]
    ops[op] += e; smp[op] += s
for op, e in ops.most_common(28):
    print(f"  {op:28s} {e / wt:7.1f} /warp-tile   samples {100 * smp[op] / tot_s:5.1f}%")
