#!/bin/bash
# memcheck over the parity tests of the kernels added last (batch builder, row-store Adam, device-step sampler)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_gpu_batch_builder.py tests/test_gpu_table_adam.py tests/test_gpu_sampler.py -q --no-header -x \
    -k "not distribution" > gpurun_out/sanitize_new.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/sanitize_new.log | head -20
