#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_train_path.py tests/test_gpu_hotpath.py tests/test_gpu_dropout.py tests/test_gpu_topologies.py tests/test_gpu_full_size.py -m gpu -q --no-header -x 2>&1 | tail -4
echo "== timing"
timeout 120 python scripts/score_rows_bench.py | tail -1
timeout 100 python scripts/step_timeline.py | tail -8
timeout 600 python bench.py --only train,gather --skip-cpu --steps 100 > gpurun_out/r2_bench_w.json 2> gpurun_out/r2_bench_w.err; echo "rc=$?"; tail -2 gpurun_out/r2_bench_w.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench_w.json").read().strip().splitlines()[-1])
print("train", round(j["value"]), "seqs/s", round(j["ms_per_step"],4), "roof", j["roofline"]["frac"], j["roofline"]["kernel"], j["roofline"]["whole_step"])
for k,v in j["stages"].items(): print("  ", k, v["ms"], v["frac_of_hbm_peak"])
print("gather", j["gather"]["ms"], j["gather"]["dense_ids"], j["gather"]["hbm_table"]["uniform"], j["gather"]["hbm_table"]["zipf"])
print("dropout", j["fused_dropout"])
PY
