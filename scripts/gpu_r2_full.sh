#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== whole gpu suite"
timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -6 | tee gpurun_out/r2_tests_all.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
echo "== bench (default flags)"
timeout 1500 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo "rc=$?"
tail -3 gpurun_out/r02_bench_full.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "rc=$?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02_bench_full.json").read().strip().splitlines()[-1])
print("train", round(j["value"]), "seqs/s", round(j["ms_per_step"],4), "ms; eager", round(j["ms_per_step_eager"],4), "e2e", round(j["e2e"]["value"]), "roof", j["roofline"]["frac"], j["roofline"]["kernel"], j["roofline"]["whole_step"], "launches", j["gpu_launches"], "steps", j["steps"])
for k,v in j["stages"].items(): print("  ", k, v["ms"], v["frac_of_hbm_peak"])
print("dropin", j["dropin_surface"])
print("dense", j["dense_batch"]["ms_per_step"], j["dense_batch"]["frac_of_hbm_peak"], j["dense_batch"]["k2_table_backward"])
for k,c in j["catalog"].items(): print(k, round(c["value"]), "users/s", c["roofline"]["frac"], "rank-only", c["rank_only"]["users_per_s"], "cpu", c.get("cpu_baseline",{}).get("value"))
print("topo", {k: {vk: (round(vv["value"]), vv["roofline"]["frac"]) for vk, vv in v["variants"].items()} for k, v in j["train_topologies"].items()})
print("cpu", j["cpu_baseline"])
print("ref", open("gpurun_out/r02_bench_reference_arm.json").read()[:400])
PY
