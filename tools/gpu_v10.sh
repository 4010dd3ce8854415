#!/bin/bash
OUT=gpurun_out/${1:-v10}; mkdir -p $OUT
for v in 0 1; do
echo "== MDG_RESCALE_SORT=$v"
MDG_RESCALE_SORT=$v timeout 150 python tools/bench_rescale.py --reads 8000000 --steps 3 --warmup 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SE kernel %.3f G reads/s, %.3f ms per batch'%(d['kernel_only_value']/1e9,d['kernel_ms_per_batch']))"
MDG_RESCALE_SORT=$v timeout 150 python tools/bench_rescale.py --reads 8000000 --steps 3 --warmup 1 --paired 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PE kernel %.3f G reads/s, %.3f ms per batch'%(d['kernel_only_value']/1e9,d['kernel_ms_per_batch']))"
done
echo "== pytest rescale"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py -m gpu -x -q --timeout 120 -k "rescale" > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest.log
