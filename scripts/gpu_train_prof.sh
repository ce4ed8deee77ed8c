#!/bin/bash
# train-path check: tests, bench, warm-cache launch list, full ncu captures of the train kernels
mkdir -p gpurun_out
bash scripts/gpu_train_check.sh
echo "== full ncu: train kernels (eager, warm caches)"
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on \
    -k regex:'chunk_reduce|radix_|span_fixup|compact_|zero_fill|sampled_score|gather_pe' -s 34 -c 17 \
    -o gpurun_out/prof_train -f python bench.py --steps 3 --warmup 3 --only train --skip-cpu --no-graph \
    > gpurun_out/prof_train.log 2>&1; echo "rc=$?"
