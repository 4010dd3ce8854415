#!/bin/bash
# A/B of environment-selected kernel variants: parity once, then a short bench per variant.
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
if [ -z "$SKIP" ]; then
echo "== pytest -m gpu (default)"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?"; tail -12 "$OUT/pytest_gpu.log"
fi
shift
for variant in "$@"; do
  echo "== bench $variant"
  name=$(echo $variant | tr '= ,/.' '_____')
  env $variant timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --cpu-sample 100000 > "$OUT/bench_$name.json" 2> "$OUT/err.log"
  tail -2 "$OUT/err.log"
  python - "$OUT/bench_$name.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("value %.3e reads/s  ms/step %.3f  kernel ms/launch %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"]))
PY
done
