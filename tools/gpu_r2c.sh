#!/bin/bash
# round 2, third visit: bit-plane kernel parity + bench; device BAM path again with stage timings
OUT=gpurun_out/r2c; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --configs c3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cut -c1-6000 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench_bam 8M"; MDG_BAM_TIMING=1 timeout 900 python tools/bench_bam.py --reads 8000000 > $OUT/bench_bam.json 2> $OUT/bench_bam.err; echo "rc=$?"; cat $OUT/bench_bam.json; grep "device slab" $OUT/bench_bam.err | tail -12
