#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
timeout 400 python -m pytest tests/test_gpu_dropout.py tests/test_evaldriver.py tests/test_gpu_train_path.py -m gpu -q --no-header -x 2>&1 | tail -30 | tee gpurun_out/r2_tests_f.log
