#!/bin/bash
# One GPU-box visit: parity tests, smoke, a bench line, the ncu launch list and one full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?"; tail -5 "$OUT/pytest_gpu.log"
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "rc=$?"; tail -3 "$OUT/smoke.log"
echo "== bench"; timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "rc=$?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "rc=$?"; cat "$OUT/bench_ref.json"
if [ -z "$SKIP_NCU" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --cpu-sample 100000 > "$OUT/bench_under_ncu.log" 2>&1; echo "rc=$?"
echo "== ncu full capture"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:count_ -s 13 -c 2 -f -o "$OUT/prof_count" \
    python bench.py --steps 1 --warmup 1 --no-e2e --cpu-sample 100000 > "$OUT/ncu_full.log" 2>&1; echo "rc=$?"
fi
ls -la "$OUT"
