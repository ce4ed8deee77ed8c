#!/bin/bash
# round-2 A/B timing of the training step's knobs (each line: graph-replayed cfg2 step + the branch timeline)
#   RECGURU_SCORE_ROWS=0   compacting scoring kernel instead of the one over the plan's live-row list
#   RECGURU_ROWS_CTAS=1..8 CTAs per SM of the live-row-list kernel (default 4)
#   RECGURU_ZERO_AFTER=score | RECGURU_ZERO_STREAM=0   where the dense zero fill of dW runs
#   RECGURU_REDUCE=regs    register-staged chunk_reduce instead of the TMA-staged one
#   RECGURU_BWD_PLAN=2     fused plan (4 launches) instead of the 9-launch chain
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
for knob in "" "RECGURU_SCORE_ROWS=0" "RECGURU_ROWS_CTAS=3" "RECGURU_ROWS_CTAS=2" "RECGURU_ZERO_AFTER=score" "RECGURU_ZERO_STREAM=0" "RECGURU_REDUCE=regs" "RECGURU_BWD_PLAN=2"; do
  echo "== ${knob:-default}"
  env $knob timeout 120 python scripts/score_rows_bench.py | tail -2
  env $knob timeout 100 python scripts/step_timeline.py | tail -8
done
