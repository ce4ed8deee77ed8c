#!/bin/bash
# final round-2 evidence: whole gpu suite, smoke, default bench + reference arm, launch list, ncu --set full of the train kernels
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
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "rc=$?"
echo "== launch list (train + cfg4, eager so that every kernel is listed)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --only train,cfg4 --skip-cpu --no-graph \
    > gpurun_out/r02_launches_bench.log 2>&1; echo "rc=$?"
echo "== full: train-step kernels (eager step, overlap schedule)"
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'sampled_score_rows|chunk_reduce|span_fixup|gather_pe_mask|radix_|row_scatter|row_count|compact_scan|pad_mask|zero_fill|loss_finalize' -s 16 -c 16 \
    -o gpurun_out/r02_train2 -f python scripts/one_step.py overlap > gpurun_out/r02_train2.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r02_train2* gpurun_out/r02_launches.csv
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02_bench_full.json").read().strip().splitlines()[-1])
print("train", round(j["value"]), "seqs/s", round(j["ms_per_step"],4), "ms; eager", round(j["ms_per_step_eager"],4), "e2e", round(j["e2e"]["value"]), round(j["e2e"]["ms_per_step"],4), "int64", round(j["e2e"]["int64_ids"]["value"]), "roof", j["roofline"]["frac"], j["roofline"]["kernel"], j["roofline"]["traffic"], j["roofline"]["whole_step"], "launches", j["gpu_launches"], "steps", j["steps"], "sustained", j["sustained"]["value"])
for k,v in j["stages"].items(): print("  ", k, v["ms"], v["frac_of_hbm_peak"])
print("dropin", j["dropin_surface"])
print("dense", j["dense_batch"]["ms_per_step"], j["dense_batch"]["value"], j["dense_batch"]["frac_of_hbm_peak"], j["dense_batch"]["k2_table_backward"])
print("devbatches", j["e2e_device_batches"]["value"], j["e2e_device_batches"]["ms_per_step"], "devneg", j["e2e_device_negatives"]["value"])
print("adam", j["table_adam"]["ms"], j["table_adam"]["frac_of_hbm_peak"], j["table_adam"]["row_store"]["ms"])
print("dropout", j["fused_dropout"])
print("gather", j["gather"]["ms"], j["gather"]["hbm_table"]["uniform"]["cold"], j["gather"]["hbm_table"]["zipf"]["cold"])
for k,c in j["catalog"].items(): print(k, round(c["value"]), "users/s", c["ms_per_step"], c["roofline"]["frac"], c["roofline"]["achieved"], "rank-only", c["rank_only"], "cpu", c.get("cpu_baseline",{}).get("value"), c["clocks"], "e2e", c["e2e"]["value"])
print("topo", {k: {vk: (round(vv["value"]), vv["ms_per_step"], vv["roofline"]["frac"]) for vk, vv in v["variants"].items()} for k, v in j["train_topologies"].items()})
print("cpu", j["cpu_baseline"])
print("ref", open("gpurun_out/r02_bench_reference_arm.json").read()[:300])
PY
