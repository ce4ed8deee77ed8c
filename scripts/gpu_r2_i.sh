#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== rows kernel tests"
timeout 600 python -m pytest tests/test_gpu_train_path.py tests/test_gpu_topologies.py tests/test_gpu_hotpath.py -m gpu -q --no-header -x 2>&1 | tail -8
echo "== timeline rows"
timeout 100 python scripts/step_timeline.py
echo "== timeline compacting kernel"
RECGURU_SCORE_ROWS=0 timeout 100 python scripts/step_timeline.py
echo "== bench"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err; echo "rc=$?"
tail -3 gpurun_out/r2_bench_i.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench_i.json").read().strip().splitlines()[-1])
print("train", round(j["value"]), "seqs/s", round(j["ms_per_step"],4), "ms; eager", round(j["ms_per_step_eager"],4), "e2e", round(j["e2e"]["value"]), j["e2e"], "roof", j["roofline"])
for k,v in j["stages"].items(): print("  ", k, v["ms"], v["frac_of_hbm_peak"])
print("dropin", j["dropin_surface"])
print("topo", {k: {vk: (round(vv["value"]), vv["roofline"]["frac"]) for vk, vv in v["variants"].items()} for k, v in j["train_topologies"].items()})
PY
