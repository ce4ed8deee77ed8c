"""Top stall-sample lines of one kernel from `ncu --page source --csv` output (first instance)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
iS, iI, isrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
data = []
for n, r in enumerate(rows[2:]):
    if r and r[0] == "Kernel Name":
        break
    if len(r) > iI and r[iS].isdigit():
        data.append((int(r[iS]), int(r[iI] or 0), r[isrc].strip(), n))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data), "executed", sum(d[1] for d in data))
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]:6d} {100 * d[0] / tot:5.1f}%  ex={d[1]:8d}  #{d[3]:5d} {d[2][:100]}")
