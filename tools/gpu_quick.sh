#!/bin/bash
# Quick GPU visit: parity tests + one bench line. Usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?"; tail -15 "$OUT/pytest_gpu.log"
echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS} > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "rc=$?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
