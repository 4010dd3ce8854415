#!/bin/bash
# visit: parity of the planes kernels (mode switches, identical reads), timing by shape for the variants in $VARIANTS
OUT=gpurun_out/${1:-v3}; mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${PYTEST_K:-mode_switches or identical}" > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest.log
IFS=';' read -ra VS <<< "${VARIANTS:-MDG_PLANES_WS=0;MDG_PLANES_WS=2x8+8}"
for v in "${VS[@]}"; do
  echo "== shapes $v"; env $v timeout 300 python tools/bench_shapes.py ${SHAPES:-se100 se50-150} 2>&1 | tail -4
done
