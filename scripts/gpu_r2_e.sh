#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== source counters: sparse scoring kernel"
timeout 200 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section Occupancy --section LaunchStats --clock-control none --import-source on \
    -k regex:sampled_score_fwd_sparse -s 2 -c 1 -o gpurun_out/r2_score_src -f python scripts/one_step.py > gpurun_out/r2_score_src.log 2>&1; echo rc=$?
echo "== source counters: chunk_reduce + span_fixup"
timeout 200 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section Occupancy --section LaunchStats --clock-control none --import-source on \
    -k regex:'chunk_reduce|span_fixup' -s 4 -c 2 -o gpurun_out/r2_apply_src -f python scripts/one_step.py > gpurun_out/r2_apply_src.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep | tail -3
