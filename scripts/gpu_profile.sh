#!/bin/bash
# ncu evidence: launch list of one bench run + full captures of the dominant kernels.
mkdir -p gpurun_out
echo "== launch list (train + cfg4)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --only train,cfg4 --skip-cpu \
    > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"
echo "== full: catalog sweep"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:catalog_sweep -s 3 -c 1 \
    -o gpurun_out/prof_sweep -f python bench.py --steps 3 --warmup 3 --only cfg4 --skip-cpu \
    > gpurun_out/prof_sweep.log 2>&1; echo "rc=$?"
echo "== full: table backward kernels"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'chunk_reduce|radix_|span_fixup|build_keys|zero_fill' -s 40 -c 12 \
    -o gpurun_out/prof_bwd -f python bench.py --steps 3 --warmup 3 --only train --skip-cpu \
    > gpurun_out/prof_bwd.log 2>&1; echo "rc=$?"
echo "== full: sampled score fwd/bwd + gather"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'sampled_score|gather_pe' -s 12 -c 3 \
    -o gpurun_out/prof_score -f python bench.py --steps 3 --warmup 3 --only train --skip-cpu \
    > gpurun_out/prof_score.log 2>&1; echo "rc=$?"
ls -la gpurun_out/
