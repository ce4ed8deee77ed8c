#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== tests"
timeout 600 python -m pytest tests/test_gpu_catalog.py tests/test_gpu_sampler.py tests/test_gpu_topologies.py tests/test_gpu_train_path.py tests/test_gpu_dropout.py -m gpu -q --no-header -x 2>&1 | tail -8 | tee gpurun_out/r2_tests_g.log
echo "== bench train"
timeout 300 python bench.py --steps 20 --warmup 5 --only train --skip-cpu > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; echo rc=$?
tail -3 gpurun_out/r2_bench_g.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench_g.json").read().strip().splitlines()[-1])
print("train", round(j["value"]), "seqs/s", round(j["ms_per_step"],4), "ms; eager", round(j["ms_per_step_eager"],4))
print("dropin", j["dropin_surface"])
print("dense", j["dense_batch"])
print("dropout", j["fused_dropout"])
PY
