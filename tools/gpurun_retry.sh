#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 / "transient"): tools/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for attempt in $(seq 1 30); do
  out=$(gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 100; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
