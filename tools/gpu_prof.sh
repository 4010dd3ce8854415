#!/bin/bash
# ncu: launch list + full capture of the counting kernels. Usage: bash tools/gpu_prof.sh <tag> [kernel regex] [extra bench args]
TAG=${1:-prof}
PAT=${2:-count_}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?"; tail -5 "$OUT/pytest_gpu.log"
echo "== bench"; timeout 900 python bench.py $3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "rc=$?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --cpu-sample 100000 $3 > "$OUT/bench_under_ncu.log" 2>&1; echo "rc=$?"
echo "== ncu full capture"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$PAT -s 26 -c 2 -f -o "$OUT/prof" \
    python bench.py --steps 1 --warmup 1 --no-e2e --cpu-sample 100000 $3 > "$OUT/ncu_full.log" 2>&1; echo "rc=$?"
ls -la "$OUT"
