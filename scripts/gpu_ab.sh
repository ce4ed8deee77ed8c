#!/bin/bash
# A/B of one env knob on the training-step bench: $1 = "VAR=value" for arm B (arm A = default).
# $2 (optional) = pytest selection run first.
mkdir -p gpurun_out
echo "== tests"
timeout 600 python -m pytest ${2:-tests/test_gpu_train_path.py tests/test_gpu_hotpath.py tests/test_gpu_table_adam.py tests/test_gpu_full_size.py} -m gpu -q --no-header -x 2>&1 | tail -30
for arm in A B A B; do
  if [ "$arm" = "B" ]; then export "$1"; else unset "${1%%=*}"; fi
  timeout 300 python bench.py --steps 50 --warmup 10 --only train --skip-cpu 2>gpurun_out/ab.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('arm $arm', 'ms/step', round(j['ms_per_step'],4), 'eager', round(j['ms_per_step_eager'],4), {k: v['ms'] for k, v in j['stages'].items()}, 'e2e', round(j['e2e']['ms_per_step'],4), 'dev_batches', round(j['e2e_device_batches']['ms_per_step'],4))
"
done
tail -2 gpurun_out/ab.err
