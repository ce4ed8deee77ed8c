#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_path.py tests/test_gpu_hotpath.py tests/test_gpu_candidates.py tests/test_gpu_table_adam.py -q --no-header -x 2>&1 | tail -3
for pdl in 1 0; do
RECGURU_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --only train --skip-cpu 2>gpurun_out/bp$pdl.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('PDL=$pdl', round(j['value']), 'ms', round(j['ms_per_step'],4), 'eager', round(j['ms_per_step_eager'],4), 'e2e', round(j['e2e']['value']), {k:v['ms'] for k,v in j['stages'].items()})"
tail -2 gpurun_out/bp$pdl.err
done
