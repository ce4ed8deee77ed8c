#!/bin/bash
mkdir -p gpurun_out
python -m recguru_b200.build > /dev/null
echo "== full: catalog sweep v2 (cfg4 shape, top-10)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:catalog_sweep -s 1 -c 1 \
    -o gpurun_out/r02_sweep -f python scripts/catalog_one.py > gpurun_out/r02_sweep.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r02_sweep.log
echo "== source counters: catalog sweep v2"
timeout 400 ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on \
    -k regex:catalog_sweep -s 1 -c 1 -o gpurun_out/r02_sweep_src -f python scripts/catalog_one.py > gpurun_out/r02_sweep_src.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r02_sweep*
