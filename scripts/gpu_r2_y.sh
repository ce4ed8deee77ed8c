#!/bin/bash
python -m recguru_b200.build > /dev/null
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_train_path.py tests/test_gpu_hotpath.py tests/test_gpu_topologies.py tests/test_gpu_full_size.py tests/test_abi.py -m gpu -q --no-header -x 2>&1 | tail -4
echo "== cfg3 timeline: rows kernel"
timeout 200 python scripts/step_timeline_cfg3.py
echo "== cfg3 timeline: compacting kernel"
RECGURU_SCORE_ROWS=0 timeout 200 python scripts/step_timeline_cfg3.py | tail -3
echo "== bench topologies"
timeout 600 python bench.py --only cfg1,cfg3 --skip-cpu --steps 100 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: {vk: (round(vv['value']), round(vv['ms_per_step'],4), vv['roofline']['frac']) for vk, vv in v['variants'].items()} for k, v in j['train_topologies'].items()})
"
