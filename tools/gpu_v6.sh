#!/bin/bash
OUT=gpurun_out/${1:-v6}; mkdir -p $OUT
for v in 1 3 5 7; do
echo "== MDG_PLANES_PREFETCH=$v"; MDG_PLANES_PREFETCH=$v timeout 90 python tools/bench_shapes.py se100 se50-150 2>&1 | tail -2
MDG_PLANES_PREFETCH=$v timeout 120 python tools/prof_g3.py 2>&1 | tail -1; MDG_PLANES_PREFETCH=$v timeout 120 python tools/prof_g3.py sorted 2>&1 | tail -1
done
