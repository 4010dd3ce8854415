#!/bin/bash
# visit: the sorted-reads test, planes parity variants, A/B of planes launch variants by read shape
OUT=gpurun_out/${1:-v1}; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest.log
for v in "X=1" "MDG_PLANES_PREFETCH=0" "MDG_PLANES_THREADS=256" "MDG_PLANES_THREADS=256 MDG_PLANES_PREFETCH=0"; do
  echo "== shapes $v"; env $v timeout 300 python tools/bench_shapes.py se100 se50-150 "c3 1 lib" "c3 2 libs" 2>&1 | tail -4
done
