#!/bin/bash
for v in "X=1" "MDG_PLANES_PREFETCH=1" "MDG_PLANES_WS=2x8+8" "MDG_PLANES_WS=2x8+8 MDG_PLANES_PREFETCH=1" "MDG_PLANES_WS=0"; do
echo "== $v"; env $v timeout 120 python tools/prof_g3.py 2>&1 | tail -1; env $v timeout 120 python tools/prof_g3.py sorted 2>&1 | tail -1
done
