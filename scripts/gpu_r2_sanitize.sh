#!/bin/bash
# memcheck over the parity tests of the kernels added or rewritten in round 2: live-row-list scoring, two-stage plan,
# TMA-staged reduce, span fix-up, fused dropout, learned-position gradient, last-position extraction, the step
# topologies, and the rewritten catalog sweep (small catalogs)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_gpu_train_path.py tests/test_gpu_dropout.py tests/test_gpu_topologies.py tests/test_gpu_hotpath.py \
    -q --no-header -x -k "live_row_list or embedding_bwd or plan or fused_forward or dropout or learned or extract or topolog or autorec or recguru or step_matches" \
    > gpurun_out/r02_memcheck_train.log 2>&1; echo "memcheck train rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r02_memcheck_train.log | head -20
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_gpu_catalog.py -q --no-header -x -k "not large and not property" \
    > gpurun_out/r02_memcheck_catalog.log 2>&1; echo "memcheck catalog rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r02_memcheck_catalog.log | head -20
