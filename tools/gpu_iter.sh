#!/bin/bash
# kernel iteration: parity of the counting paths, one bench line (headline only), optional extra bench args in $BENCH_ARGS
TAG=${1:-iter}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest.log
echo "== bench"; timeout 900 python bench.py --configs ${CONFIGS:-none} --no-e2e $BENCH_ARGS > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print("c2: value %.3e reads/s  kernel ms/launch %.4f  frac %.4f" % (d["value"], r["kernel_ms_per_launch"], r["frac"]))
for k,v in d.get("configs",{}).items():
    if "roofline" in v: print("%s: value %.3e reads/s kernel ms/launch %.4f" % (k, v["value"], v["roofline"]["kernel_ms_per_launch"]))
PY
