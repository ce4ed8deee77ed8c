#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train_path.py -m gpu -q --no-header 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --only train,gather --skip-cpu 2>gpurun_out/bt.err > gpurun_out/bt.json; tail -2 gpurun_out/bt.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/bt.json"))
print("train seqs/s", round(j["value"]), "ms/step", round(j["ms_per_step"],4), "eager", round(j.get("ms_per_step_eager",0),4), "e2e ms", round(j["e2e"]["ms_per_step"],4), "e2e", round(j["e2e"]["value"]), "launches", j["gpu_launches"])
for k,v in j["stages"].items(): print(" ", k, v)
PY
timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 3 --warmup 3 --only train --skip-cpu --no-graph > /dev/null 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches_train.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    n=row['Kernel Name'].split('(')[0]; v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    agg.setdefault(n,[]).append(v)
for k,v in agg.items():
    if 'rg::' in k: print(f"{k[:60]:60s} n={len(v):3d} avg={sum(v)/len(v):8.2f} us")
PY
