#!/bin/bash
# 8-GPU visit: H2D bandwidth per rank at 1/2/4/8 ranks (bound / unbound), the bench at N=8 (configs[4] = 1 B reads) and N=4
OUT=gpurun_out/r2g; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; lscpu | head -30 > $OUT/lscpu.txt; numactl -H >> $OUT/lscpu.txt 2>&1; free -g >> $OUT/lscpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do for bind in 1 0; do
  echo "== h2d n=$n bind=$bind"; timeout 300 $TR --nproc-per-node $n --master-port $((29600+n+bind)) tools/bench_h2d.py --bind $bind --gb 2 2>/dev/null | grep '^{' | tee -a $OUT/h2d.jsonl | cut -c1-400
done; done
echo "== bench N=8"; timeout 1500 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/bench8.json 2> $OUT/bench8.err; echo "rc=$?"; tail -3 $OUT/bench8.err; cut -c1-3000 $OUT/bench8.json
echo "== bench N=4"; timeout 900 $TR --nproc-per-node 4 --master-port 29712 bench.py --gpus 4 --steps 5 --warmup 3 > $OUT/bench4.json 2> $OUT/bench4.err; echo "rc=$?"; cut -c1-1200 $OUT/bench4.json
echo "== bench N=2"; timeout 900 $TR --nproc-per-node 2 --master-port 29713 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench2.json 2> $OUT/bench2.err; echo "rc=$?"; cut -c1-1200 $OUT/bench2.json
