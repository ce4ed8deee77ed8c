#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <script> [log]   -- retries while the pod answers busy (rc 3)
T=$1; S=$2; LOG=${3:-gpurun_out/last_run.log}
mkdir -p gpurun_out
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "bash $S" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)"; tail -90 "$LOG"; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
