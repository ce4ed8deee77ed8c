#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
for c in 4 3 2 1; do
  echo "== RECGURU_ROWS_CTAS=$c"
  RECGURU_ROWS_CTAS=$c timeout 120 python scripts/score_rows_bench.py
  RECGURU_ROWS_CTAS=$c timeout 100 python scripts/step_timeline.py | tail -8
done
echo "== compacting kernel (RECGURU_SCORE_ROWS=0)"
RECGURU_SCORE_ROWS=0 timeout 120 python scripts/score_rows_bench.py | tail -1
