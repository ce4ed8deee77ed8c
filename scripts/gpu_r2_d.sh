#!/bin/bash
mkdir -p gpurun_out
for v in 2 1; do
echo "== plan only, RECGURU_BWD_PLAN=$v"
RECGURU_BWD_PLAN=$v timeout 60 python scripts/plan_only.py
RECGURU_BWD_PLAN=$v timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_plan_v$v.csv python scripts/plan_only.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_plan_v$v.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
seen=[]
for r in rows[1:]:
    seen.append((r[ik].split("(")[0][:60], float(r[iv].replace(",",""))/1e3))
# last full iteration
names=[s[0] for s in seen]
first=names.index(names[0],1) if names.count(names[0])>1 else len(names)
for nme,us in seen[-first:]: print(f"   {nme:60s} {us:8.2f} us")
print("   sum", round(sum(u for _,u in seen[-first:]),2))
PY
done
echo "== dense batch"
timeout 60 python scripts/plan_only.py dense
RECGURU_BWD_PLAN=1 timeout 60 python scripts/plan_only.py dense
