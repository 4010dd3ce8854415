#!/bin/bash
# ncu full capture of the bit-plane kernel (one launch of 4.17 M reads)
OUT=gpurun_out/r2d; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:count_planes -s 13 -c 1 -f -o $OUT/prof_planes \
    python bench.py --steps 1 --warmup 1 --no-e2e --cpu-sample 100000 --configs none > $OUT/ncu_full.log 2>&1; echo "rc=$?"
tail -3 $OUT/ncu_full.log
ls -la $OUT
