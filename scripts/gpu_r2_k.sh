#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
timeout 300 python -m pytest tests/test_gpu_train_path.py -m gpu -q --no-header -x -k "live_row_list" 2>&1 | tail -3
for c in 4 2; do
  echo "== RECGURU_ROWS_CTAS=$c"
  RECGURU_ROWS_CTAS=$c timeout 120 python scripts/score_rows_bench.py
  RECGURU_ROWS_CTAS=$c timeout 100 python scripts/step_timeline.py | tail -8
done
