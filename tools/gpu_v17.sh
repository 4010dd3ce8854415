#!/bin/bash
OUT=gpurun_out/${1:-v17}; mkdir -p $OUT
echo "== pytest (indel-heavy + golden)"; timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "indel or golden or synthetic or libraries_in_one or mode_switches" > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest.log
for v in "X=1" "MDG_PLANES_INDELS=0"; do
echo "== shapes $v"; env $v timeout 120 python tools/bench_shapes.py se100 se50-150 "se50-150+indels" "c3 1 lib" "c3 2 libs" 2>&1 | tail -5
done
