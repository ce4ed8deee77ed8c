#!/bin/bash
# driver-style launch at N GPUs: reference arm first, then ours
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref rc=$? lines=$(wc -l < gpurun_out/bench_ref_n$N.json)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "ours rc=$? lines=$(wc -l < gpurun_out/bench_n$N.json)"
tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_n$N.json"))
print("train", round(j["value"]), "seqs/s n_gpus", j["n_gpus"], "ms/step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["value"]))
for k,c in j["catalog"].items(): print(k, round(c["value"]), "users/s", round(c["ms_per_step"],2), "ms", c["config"].get("sharding"), "e2e", round(c["e2e"]["value"]), "rank_only", c["rank_only"]["users_per_s"])
r=json.load(open("gpurun_out/bench_ref_n$N.json")); print("ref", r["value"], r["n_gpus"])
PY
