#!/bin/bash
# Full round: tests, smoke, bench (+reference arm), ncu launch list + full captures for profiles/.
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -4 | tee gpurun_out/tests_full.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; tail -2 gpurun_out/bench_full.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json | cut -c1-300
echo "== candidate protocol"; timeout 200 python scripts/candidate_bench.py 2>&1 | tee gpurun_out/candidate_bench.log
echo "== catalog knobs"; timeout 300 python scripts/catalog_exp.py 2>&1 | tee gpurun_out/catalog_exp_final.log
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 3 --warmup 3 --only train,cfg4 --skip-cpu --no-graph > /dev/null 2>&1; echo rc=$?
echo "== full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:catalog_sweep -s 3 -c 1 -o gpurun_out/r01_sweep -f python bench.py --steps 3 --warmup 3 --only cfg4 --skip-cpu > /dev/null 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gather_pe_mask|sampled_score|chunk_reduce|radix_scatter|compact_scatter|span_fixup' -s 27 -c 9 -o gpurun_out/r01_train -f python bench.py --steps 3 --warmup 3 --only train --skip-cpu --no-graph > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out | tail -8
