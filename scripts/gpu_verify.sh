#!/bin/bash
# Short GPU-box check after a change: all parity tests, smoke, the default bench line.
mkdir -p gpurun_out
echo "== tests"
timeout 600 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -25 | tee gpurun_out/tests_verify.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_verify.log
echo "== bench"
timeout 600 python bench.py > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err; echo "rc=$?"
tail -5 gpurun_out/bench_verify.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_verify.json").read().strip().splitlines()[-1])
print("value", round(j["value"]), "ms", round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["value"]), j["e2e"]["ms_per_step"])
for k in ("e2e_device_negatives", "e2e_device_batches", "table_adam"):
    print(k, j.get(k))
print("clocks", j.get("clocks"))
PY
