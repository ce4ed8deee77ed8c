#!/bin/bash
# run one test file (or any pytest selection) on the GPU box: $@ = pytest arguments
mkdir -p gpurun_out
timeout 600 python -m pytest "$@" -m gpu -q --no-header -x 2>&1 | tail -40 | tee gpurun_out/onefile.log
