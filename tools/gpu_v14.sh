#!/bin/bash
timeout 120 python tools/bench_shapes.py se100 se50-150 "c3 1 lib" "se100 2 libs" 2>&1 | tail -4
