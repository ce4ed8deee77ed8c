#!/bin/bash
# round 2, run B: v2 sweep variants (epilogue warps per lane quarter) -- quick parity, timing, full parity.
# Every step under a SHORT timeout: a deadlocked experimental kernel must not eat the lease.
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
ok=1
for nw in 3 4; do
echo "== quick parity, RECGURU_CATALOG_WARPS=$nw"
RECGURU_CATALOG_WARPS=$nw timeout 90 python -m pytest tests/test_gpu_catalog.py -m gpu -q --no-header -x 2>&1 | tail -4 | tee gpurun_out/r2_tests_b_nw$nw.log
grep -q "passed" gpurun_out/r2_tests_b_nw$nw.log && ! grep -q "failed" gpurun_out/r2_tests_b_nw$nw.log || ok=0
done
if [ $ok = 1 ]; then
echo "== catalog experiments"
timeout 150 python scripts/catalog_exp.py 2>&1 | tee gpurun_out/r2_catalog_exp_b2.log
for nw in 3 4; do
echo "== timed-shape parity, RECGURU_CATALOG_WARPS=$nw"
RECGURU_CATALOG_WARPS=$nw timeout 150 python -m pytest tests/test_evaldriver.py "tests/test_gpu_full_size.py::test_catalog_cfg4_timed_shape_single_split" "tests/test_gpu_full_size.py::test_catalog_10m_items" -m gpu -q --no-header -x 2>&1 | tail -4 | tee -a gpurun_out/r2_tests_b_nw$nw.log
done
fi
