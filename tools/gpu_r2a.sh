#!/bin/bash
# round 2, first visit (2 GPUs): parity tests incl. the NCCL test, bench at N=1 with sub-configs, bench at N=2
OUT=gpurun_out/r2a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt; df -h /tmp /dev/shm >> $OUT/nproc.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== bench N=1"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench2.json 2> $OUT/bench2.err; echo "rc=$?"; cat $OUT/bench2.json; tail -5 $OUT/bench2.err
