#!/bin/bash
# Runs bench.py variants given as quoted argument strings; prints value and kernel ms per launch.
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
n=0
for args in "$@"; do
  n=$((n+1))
  echo "== bench.py $args"
  eval "timeout 600 python bench.py --no-e2e --cpu-sample 100000 $args" > "$OUT/bench_$n.json" 2> "$OUT/err_$n.log"; tail -2 "$OUT/err_$n.log"
  python - "$OUT/bench_$n.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("value %.3e reads/s  ms/step %.3f  kernel ms/launch %.4f reads/launch %d -> %.3e reads/s kernel" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["reads_per_launch"], r["reads_per_launch"]/r["kernel_ms_per_launch"]*1e3))
PY
done
