#!/bin/bash
# source-level counters of the sweep kernel (one launch); env knobs pass through
mkdir -p gpurun_out
timeout 600 ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on \
    -k regex:catalog_sweep -s 1 -c 1 -o gpurun_out/sweep_src${1} -f python scripts/catalog_one.py > gpurun_out/sweep_src${1}.log 2>&1; echo rc=$?
tail -2 gpurun_out/sweep_src${1}.log
