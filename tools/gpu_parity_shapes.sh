#!/bin/bash
OUT=gpurun_out/${1:-v12}; mkdir -p $OUT
echo "== shapes"; timeout 120 python tools/bench_shapes.py se100 se50-150 "c3 1 lib" "c3 2 libs" "se100 2 libs" 2>&1 | tail -5; echo "rc=$?"
echo "== pytest"; timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest.log
