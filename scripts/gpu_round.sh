#!/bin/bash
# One GPU-box round: smoke, parity tests, bench, launch list.  Every step has its own timeout
# so that a hung kernel cannot eat the lease.  Output lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== catalog debug" | tee gpurun_out/debug.log
timeout 300 python scripts/catalog_debug.py >> gpurun_out/debug.log 2>&1; echo "rc=$?" >> gpurun_out/debug.log
tail -40 gpurun_out/debug.log
echo "== train-path tests"
timeout 900 python -m pytest tests/test_gpu_train_path.py -m gpu -q -x --no-header 2>&1 | tail -40 | tee gpurun_out/test_train.log
echo "== catalog tests"
timeout 900 python -m pytest tests/test_gpu_catalog.py -m gpu -q --no-header 2>&1 | tail -60 | tee gpurun_out/test_catalog.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
