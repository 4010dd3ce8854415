#!/bin/bash
# configs[3] at its full size through the file path: a 200 M-read BAM -> count tables, and -> rescaled BAM (decode, rescale, deflate on the GPU)
OUT=gpurun_out/${1:-c4}; mkdir -p $OUT
df -h /dev/shm | tail -1
timeout 540 python tools/bench_bam.py --reads 200000000 --skip-host > $OUT/bam_200M.json 2> $OUT/bam_200M.err; echo "rc=$?"; tail -3 $OUT/bam_200M.err; cut -c1-2500 $OUT/bam_200M.json
