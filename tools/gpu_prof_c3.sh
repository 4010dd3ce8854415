#!/bin/bash
# ncu on the configs[2] shape (50-150 bp pairs, indels, clips, two libraries): launch list and a full capture of the
# warp-specialised kernel's two-library, indel-staging variant
OUT=gpurun_out/${1:-pc3}; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c3.csv python tools/bench_shapes.py "c3 2 libs" > $OUT/shapes_under_ncu.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:count_planes_ws -s 6 -c 1 -f -o $OUT/prof_ws_c3 python tools/bench_shapes.py "c3 2 libs" > $OUT/ncu_c3.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_c3.log
ls -la $OUT
