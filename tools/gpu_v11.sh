#!/bin/bash
for v in "2x8+8" "2x9+8" "2x9+4"; do
echo "== MDG_PLANES_WS=$v"; MDG_PLANES_WS=$v timeout 120 python tools/bench_shapes.py se100 se50-150 "c3 1 lib" 2>&1 | tail -3
done
