#!/bin/bash
# full default bench (headline + configs[2] + configs[3]) and the BAM file benchmarks
OUT=gpurun_out/r2e; mkdir -p $OUT
echo "== pytest bamdev"; timeout 900 python -m pytest tests/test_bamdev.py -m gpu -x -q 2>&1 | tail -3
echo "== bench (auto configs)"; timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -5 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print("c2: value %.3e e2e %.3e kernel ms/launch %.4f frac %.4f check %s" % (d["value"], d["e2e"]["value"], r["kernel_ms_per_launch"], r["frac"], d["check"]))
for k,v in d.get("configs",{}).items():
    print(k, json.dumps({kk:vv for kk,vv in v.items() if kk in ("value","e2e","file_to_file","check","ms_per_step")})[:1800])
PY
echo "== bench_bam 8M"; MDG_BAM_TIMING=1 timeout 900 python tools/bench_bam.py --reads ${BAM_READS:-32000000} > $OUT/bench_bam.json 2> $OUT/bench_bam.err; echo "rc=$?"; cat $OUT/bench_bam.json; grep "device slab" $OUT/bench_bam.err | tail -4
