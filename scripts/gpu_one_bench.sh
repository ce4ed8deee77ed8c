#!/bin/bash
# one short training-step bench under an env knob: $1 = "VAR=value"
mkdir -p gpurun_out
export "$1"
timeout 35 python bench.py --steps 30 --warmup 5 --only train --skip-cpu 2>gpurun_out/one.err | tee gpurun_out/one.json | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step', round(j['ms_per_step'],4), {k: v['ms'] for k, v in j['stages'].items()})
"
tail -2 gpurun_out/one.err
