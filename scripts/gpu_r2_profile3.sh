#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== full: rows scoring kernel + cluster sort (eager cfg2 steps, 3rd step)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sampled_score_rows" -s 2 -c 1 \
    -o gpurun_out/r02_rows -f python scripts/one_step.py overlap > gpurun_out/r02_rows.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r02_rows.log
ls -la gpurun_out/r02_rows*
