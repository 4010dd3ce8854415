#!/bin/bash
OUT=gpurun_out/${1:-v19}; mkdir -p $OUT
echo "== pytest"; timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest.log
