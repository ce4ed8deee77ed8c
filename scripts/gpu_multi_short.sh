#!/bin/bash
# short 2-GPU check: NCCL parity test of the sharded catalog + the default bench line at N=2
mkdir -p gpurun_out
N=${1:-2}
echo "== sharded parity"; timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q --no-header 2>&1 | tail -3
echo "== bench --gpus $N"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
j=[json.loads(l) for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1]; n_lines=sum(1 for _ in open("gpurun_out/bench_n$N.json")); print("stdout lines:", n_lines)
print("train", round(j["value"]), "seqs/s n_gpus", j["n_gpus"], "ms/step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["value"]), "dev_batches", round(j["e2e_device_batches"]["value"]))
for k,c in j["catalog"].items(): print(k, round(c["value"]), "users/s", round(c["ms_per_step"],2), "ms", c["roofline"]["achieved"], "TF/GPU-kernel", c["config"]["sharding"], "e2e", round(c["e2e"]["value"]))
PY
