#!/bin/bash
# careful visit: every step under its own short timeout (a hung kernel must not eat the GPU budget)
OUT=gpurun_out/${1:-v5}; mkdir -p $OUT
echo "== shapes"; timeout 90 python tools/bench_shapes.py se100 se50-150 "c3 1 lib" 2>&1 | tail -3; echo "rc=$?"
echo "== pytest"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "${PYTEST_K:-mode_switches or identical or sorted}" > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest.log
echo "== g3"; timeout 120 python tools/prof_g3.py 2>&1 | tail -1; timeout 120 python tools/prof_g3.py sorted 2>&1 | tail -1
