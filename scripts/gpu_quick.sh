#!/bin/bash
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -15 | tee gpurun_out/tests.log
echo "== catalog experiments"
timeout 600 python scripts/catalog_exp.py 2>&1 | tee gpurun_out/catalog_exp.log
echo "== bench (train only)"
timeout 600 python bench.py --steps 20 --warmup 5 --only train,gather --skip-cpu 2>gpurun_out/bench_quick.err | tee gpurun_out/bench_quick.json | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('train seqs/s', round(j['value']), 'ms/step', round(j['ms_per_step'],4))
for k,v in j['stages'].items(): print(' ', k, v)
print(' e2e', j['e2e']['value'], 'gather', j.get('gather'))
"
tail -3 gpurun_out/bench_quick.err
