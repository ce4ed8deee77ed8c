#!/bin/bash
# round-2 ncu evidence: launch list of one bench run + full captures of the dominant kernels
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== launch list (train + cfg4, eager so that every kernel is listed)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --only train,cfg4 --skip-cpu --no-graph \
    > gpurun_out/r02_launches_bench.log 2>&1; echo "rc=$?"
echo "== full: catalog sweep (cfg4 shape, top-10 and rank-only)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:catalog_sweep2 -s 2 -c 1 \
    -o gpurun_out/r02_sweep -f python scripts/catalog_one.py > gpurun_out/r02_sweep.log 2>&1; echo "rc=$?"
echo "== full: train-step kernels"
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'sampled_score_fwd_sparse|chunk_reduce|span_fixup|gather_pe_mask|radix_|row_scatter|row_count|compact_scan|pad_mask|zero_fill|loss_finalize' -s 17 -c 17 \
    -o gpurun_out/r02_train -f python scripts/one_step.py > gpurun_out/r02_train.log 2>&1; echo "rc=$?"
echo "== full: K1 on the 512 MB table (HBM gather)"
timeout 400 ncu --set full --clock-control none -k regex:gather_pe_mask -c 16 \
    -o gpurun_out/r02_gather_hbm -f python scripts/gather_hbm.py > gpurun_out/r02_gather_hbm.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r02_*
