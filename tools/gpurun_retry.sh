#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient): tools/gpurun_retry.sh <timeout> '<command>' [--gpus N]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|retry in a few minutes\|no box or slot"; then sleep 90; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; exit 3
