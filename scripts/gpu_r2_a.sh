#!/bin/bash
# round 2, run A: new parity tests first, then the whole GPU suite, smoke, and a full bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== new tests"
timeout 1200 python -m pytest tests/test_gpu_topologies.py tests/test_gpu_multidevice.py tests/test_gpu_full_size.py -m gpu -q --no-header -x 2>&1 | tail -30 | tee gpurun_out/r2_tests_new.log
echo "== whole gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -25 | tee gpurun_out/r2_tests_all.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r2_smoke.log
echo "== bench"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "rc=$?"
tail -5 gpurun_out/r2_bench_a.err; cut -c1-3000 gpurun_out/r2_bench_a.json
