#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== tests (TMA-staged reduce)"
timeout 900 python -m pytest tests/test_gpu_train_path.py tests/test_gpu_hotpath.py tests/test_gpu_topologies.py tests/test_gpu_full_size.py tests/test_gpu_table_adam.py -m gpu -q --no-header -x 2>&1 | tail -6
echo "== timeline"
timeout 100 python scripts/step_timeline.py | tail -8
timeout 120 python scripts/score_rows_bench.py | tail -1
echo "== register-staged reduce for comparison"
RECGURU_REDUCE=regs timeout 120 python scripts/score_rows_bench.py | tail -1
echo "== bench train only"
timeout 600 python bench.py --only train --skip-cpu --steps 100 > gpurun_out/r2_bench_s.json 2> gpurun_out/r2_bench_s.err; echo "rc=$?"; tail -2 gpurun_out/r2_bench_s.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench_s.json").read().strip().splitlines()[-1])
print("train", round(j["value"]), "seqs/s", round(j["ms_per_step"],4), "eager", round(j["ms_per_step_eager"],4), "roof", j["roofline"]["frac"], j["roofline"]["kernel"], j["roofline"]["whole_step"])
for k,v in j["stages"].items(): print("  ", k, v["ms"], v["frac_of_hbm_peak"])
print("dense", j["dense_batch"]["ms_per_step"], j["dense_batch"]["k2_table_backward"])
print("dropin", j["dropin_surface"]["one_pass_backward"], j["dropin_surface"]["ratio_graph_vs_hotpath_graph"])
PY
