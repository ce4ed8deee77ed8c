#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_candidates.py -q --no-header -x 2>&1 | tail -15 | tee gpurun_out/tests_cand.log
timeout 300 python scripts/candidate_bench.py 2>&1 | tee gpurun_out/candidate_bench.log
