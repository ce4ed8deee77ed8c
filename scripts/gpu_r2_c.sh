#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
for nt in 512 256; do
echo "== tests, RECGURU_SPARSE_THREADS=$nt"
RECGURU_SPARSE_THREADS=$nt timeout 300 python -m pytest tests/test_gpu_train_path.py tests/test_gpu_hotpath.py tests/test_gpu_topologies.py -m gpu -q --no-header -x 2>&1 | tail -4
done
for cfg in "1024 0" "512 0" "256 0" "512 16" "256 16" "256 8"; do
set -- $cfg
echo "== bench train, RECGURU_SPARSE_THREADS=$1 RECGURU_SCORE_RESERVE_SMS=$2 (DH_ZERO_STREAM=0)"
RECGURU_DH_ZERO_STREAM=0 RECGURU_SPARSE_THREADS=$1 RECGURU_SCORE_RESERVE_SMS=$2 timeout 300 python bench.py --steps 20 --warmup 5 --only train --skip-cpu > gpurun_out/r2_bench_c_t$1_r$2.json 2> gpurun_out/r2_bench_c.err; echo rc=$?
python - <<PY
import json
j=json.loads(open("gpurun_out/r2_bench_c_t$1_r$2.json").read().strip().splitlines()[-1])
print("train", round(j["value"]), "seqs/s", round(j["ms_per_step"],4), "ms; sustained", round(j["sustained"]["ms_per_step"],4), "eager", round(j["ms_per_step_eager"],4), "| score", j["stages"]["sampled_score_loss_fwd"]["ms"], "e2e", round(j["e2e"]["ms_per_step"],4), "devbatch", round(j["e2e_device_batches"]["ms_per_step"],4))
PY
done
