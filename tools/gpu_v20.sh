#!/bin/bash
OUT=gpurun_out/${1:-v20}; mkdir -p $OUT
echo "== shapes"; timeout 120 python tools/bench_shapes.py se100 "se100 -Q 20" "se50-150 -Q 20" 2>&1 | tail -3
echo "== shapes MDG_PLANES_QUAL=0"; MDG_PLANES_QUAL=0 timeout 120 python tools/bench_shapes.py "se100 -Q 20" 2>&1 | tail -1
echo "== pytest -Q cases"; timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py -m gpu -x -q --timeout 120 -k "25 or 20 or 13 or 17 or q20 or q13 or minqual or golden or synthetic" > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest.log
