#!/bin/bash
# visit: parity of the warp-specialised planes kernel and its timing by shape
OUT=gpurun_out/${1:-v2}; mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mode_switches or identical or synthetic or golden" > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest.log
for v in "MDG_PLANES_WS=0" "MDG_PLANES_WS=2x8+8" "MDG_PLANES_WS=2x8+4" "MDG_PLANES_WS=3x6+8" "MDG_PLANES_WS=4x4+8"; do
  echo "== shapes $v"; env $v timeout 300 python tools/bench_shapes.py se100 se50-150 "c3 1 lib" "c3 2 libs" 2>&1 | tail -4
done
