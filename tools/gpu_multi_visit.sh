#!/bin/bash
# 8-GPU visit at the end of round 2: the NCCL all-reduce test (skipped on one GPU), the bench at N=8 (configs[4] = 1 B reads) and N=2
OUT=gpurun_out/${1:-multi}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== pytest multigpu"; timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 200 > $OUT/pytest_multigpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_multigpu.log
echo "== bench N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/bench8.json 2> $OUT/bench8.err; echo "rc=$?"; tail -3 $OUT/bench8.err; cut -c1-1500 $OUT/bench8.json
echo "== bench N=2"; timeout 400 $TR --nproc-per-node 2 --master-port 29713 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench2.json 2> $OUT/bench2.err; echo "rc=$?"; cut -c1-1200 $OUT/bench2.json
