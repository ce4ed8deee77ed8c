#!/bin/bash
# N-GPU round: NCCL parity of the sharded catalog (+ two-devices-in-one-process tests), then the bench the
# way the driver launches it.  usage: gpu_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
nvidia-smi -L | tee gpurun_out/r2_multi_n${N}_gpus.txt
if [ "$N" = "2" ]; then
echo "== sharded NCCL parity + multi-device tests (2 GPUs)"
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_multidevice.py -m gpu -v --no-header 2>&1 | tail -25 | tee gpurun_out/r2_sharded_2gpu_tests.log
fi
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo rc=$?
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
j=[json.loads(l) for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1]
print("train", round(j["value"]), "seqs/s n_gpus", j["n_gpus"], "ms/step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["value"]), "e2e int64", round(j["e2e"]["int64_ids"]["value"]), "devbatch", round(j["e2e_device_batches"]["value"]))
for k,c in j["catalog"].items():
    print(k, round(c["value"]), "users/s", round(c["ms_per_step"],2), "ms", c["roofline"]["achieved"], "TF/GPU", c["config"]["sharding"], "e2e", round(c["e2e"]["value"]), "clocks", c["clocks"]["sm_mhz"])
    print("   parity", {kk: vv for kk, vv in c["parity"].items() if kk != "data"})
PY
