#!/bin/bash
# ncu full capture (with source) of the warp-specialised planes kernel on the se100 shape
OUT=gpurun_out/${1:-pws}; mkdir -p $OUT
V=${WS:-2x8+8}
MDG_PLANES_WS=$V timeout 900 ncu --set full --clock-control none --import-source on -k regex:count_planes_ws -s 4 -c 1 -f -o $OUT/prof_ws \
    python tools/bench_shapes.py ${SHAPE:-se100} > $OUT/ncu.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu.log
ls -la $OUT
