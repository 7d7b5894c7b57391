#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout_s> '<command>'  -- retries while the pod has no free slot (exit 3 / transient)
G=$1; T=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1; fi
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then echo "attempt $i: no slot, retrying in 120 s"; sleep 120; continue; fi
  cat /tmp/gpurun_last.log | tail -60
  exit $rc
done
echo "gave up"; exit 3
