#!/bin/bash
# source-level counters of the production sweep kernel (one launch)
mkdir -p gpurun_out
timeout 600 ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on \
    -k regex:catalog_sweep -s 1 -c 1 -o gpurun_out/sweep_src -f python scripts/catalog_one.py > gpurun_out/sweep_src.log 2>&1; echo rc=$?
tail -3 gpurun_out/sweep_src.log
