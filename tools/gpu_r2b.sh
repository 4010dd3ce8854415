#!/bin/bash
# round 2, second visit: the device BAM path (tests first, then the file benchmarks)
OUT=gpurun_out/r2b; mkdir -p $OUT
echo "== pytest bamdev"; timeout 900 python -m pytest tests/test_bamdev.py -m gpu -x -q > $OUT/pytest_bamdev.log 2>&1; echo "rc=$?"; tail -30 $OUT/pytest_bamdev.log
echo "== pytest host mirror (bam)"; timeout 900 python -m pytest tests/test_host_mirror.py -m gpu -x -q -k "bam" > $OUT/pytest_mirror.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_mirror.log
echo "== bench_bam 8M"; timeout 900 python tools/bench_bam.py --reads 8000000 > $OUT/bench_bam.json 2> $OUT/bench_bam.err; echo "rc=$?"; cat $OUT/bench_bam.json; tail -5 $OUT/bench_bam.err
