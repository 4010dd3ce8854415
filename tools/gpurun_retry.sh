#!/bin/bash
# gpurun with retries while the pod is busy: tools/gpurun_retry.sh <timeout> [--gpus N] <command>
T=$1; shift
OPTS=""
if [ "$1" = "--gpus" ]; then OPTS="--gpus $2"; shift 2; fi
for attempt in $(seq 1 40); do
  out=$(gpurun $OPTS --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -qE "status=transient|rc=3|busy"; then sleep 100; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
