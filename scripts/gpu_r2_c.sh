#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== tests"
timeout 400 python -m pytest tests/test_gpu_train_path.py tests/test_gpu_hotpath.py tests/test_gpu_topologies.py "tests/test_gpu_full_size.py::test_cfg2_step_vs_oracle_fp32_and_fp64" -m gpu -q --no-header -x 2>&1 | tail -4
for ch in 0 704; do
echo "== bench train, RECGURU_SPARSE_CHUNK=$ch"
RECGURU_SPARSE_CHUNK=$ch timeout 300 python bench.py --steps 20 --warmup 5 --only train --skip-cpu > gpurun_out/r2_bench_c_ch$ch.json 2> gpurun_out/r2_bench_c.err; echo rc=$?
python - <<PY
import json
j=json.loads(open("gpurun_out/r2_bench_c_ch$ch.json").read().strip().splitlines()[-1])
print("train", round(j["value"]), "seqs/s", round(j["ms_per_step"],4), "ms; sustained", round(j["sustained"]["ms_per_step"],4), "| score", j["stages"]["sampled_score_loss_fwd"]["ms"], "apply", j["stages"]["embedding_bwd_apply"]["ms"], "plan", j["stages"]["embedding_bwd_plan"]["ms"])
PY
done
