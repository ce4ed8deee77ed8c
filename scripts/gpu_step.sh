#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_path.py tests/test_gpu_hotpath.py -m gpu -q --no-header -x 2>&1 | tail -8
timeout 200 python scripts/fwd_bench.py 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --only train,gather --skip-cpu 2>gpurun_out/bt.err > gpurun_out/bt.json; tail -2 gpurun_out/bt.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/bt.json"))
print("train seqs/s", round(j["value"]), "ms/step", round(j["ms_per_step"],4), "eager", round(j.get("ms_per_step_eager",0),4), "serial", round(j.get("ms_per_step_serialised_stages",0),4), "roof", j["roofline"]["kernel"], j["roofline"]["frac"], j["roofline"]["whole_step"], "e2e ms", round(j["e2e"]["ms_per_step"],4), "e2e", round(j["e2e"]["value"]), "launches", j["gpu_launches"])
for k,v in j["stages"].items(): print(" ", k, v)
PY
