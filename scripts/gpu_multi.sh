#!/bin/bash
# 2-GPU round: NCCL parity test of the sharded catalog + bench at N=2 (driver launch line).
mkdir -p gpurun_out
nvidia-smi -L
echo "== sharded parity"; timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --no-header 2>&1 | tail -5
N=${1:-2}
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
j=[json.loads(l) for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1]; n_lines=sum(1 for _ in open("gpurun_out/bench_n$N.json")); print("stdout lines:", n_lines)
print("train", round(j["value"]), "seqs/s n_gpus", j["n_gpus"], "ms/step", round(j["ms_per_step"],4))
for k,c in j["catalog"].items(): print(k, round(c["value"]), "users/s", round(c["ms_per_step"],2), "ms", c["roofline"]["achieved"], "TF/GPU-kernel", c["config"]["sharding"], "e2e", round(c["e2e"]["value"]))
PY
echo "== bench --workload catalog --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 --workload catalog --only cfg5 > gpurun_out/bench_cat_n$N.json 2> gpurun_out/bench_cat_n$N.err; echo rc=$?
tail -2 gpurun_out/bench_cat_n$N.err; cut -c1-400 gpurun_out/bench_cat_n$N.json
