#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== tests"
timeout 300 python -m pytest tests/test_gpu_train_path.py tests/test_gpu_hotpath.py -m gpu -q --no-header -x 2>&1 | tail -3
echo "== timeline"
timeout 100 python scripts/step_timeline.py | tail -8
timeout 120 python scripts/score_rows_bench.py
