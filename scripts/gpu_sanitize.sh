#!/bin/bash
# memcheck over the (small-shape) parity tests of the non-tensor-core kernels + a bench sanity run
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_gpu_train_path.py tests/test_gpu_hotpath.py tests/test_gpu_candidates.py -q --no-header -x \
    -k "not 19999 and not 30000" > gpurun_out/sanitize.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/sanitize.log | head -20
echo "== bench sanity (N=1, clean stdout?)"
timeout 600 python bench.py --steps 5 --warmup 3 --only train,gather --skip-cpu > gpurun_out/b1.json 2> gpurun_out/b1.err; echo rc=$? lines=$(wc -l < gpurun_out/b1.json)
python -c "import json;j=json.load(open('gpurun_out/b1.json'));print(j['value'], j['e2e']['value'], j['roofline']['frac'])"
