#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== zero after k1 (default)"
timeout 100 python scripts/step_timeline.py | tail -8
timeout 120 python scripts/score_rows_bench.py | tail -1
echo "== zero after score"
RECGURU_ZERO_AFTER=score timeout 100 python scripts/step_timeline.py | tail -8
RECGURU_ZERO_AFTER=score timeout 120 python scripts/score_rows_bench.py | tail -1
echo "== zero on the plan stream head (RECGURU_ZERO_STREAM=0)"
RECGURU_ZERO_STREAM=0 timeout 100 python scripts/step_timeline.py | tail -8
RECGURU_ZERO_STREAM=0 timeout 120 python scripts/score_rows_bench.py | tail -1
