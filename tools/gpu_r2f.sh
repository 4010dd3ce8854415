#!/bin/bash
# device BAM path at scale: 64 M reads file -> tables / rescaled file, with stage timings; then the default bench again
OUT=gpurun_out/r2f; mkdir -p $OUT
echo "== bench_bam"; MDG_TIMING=1 MDG_BAM_TIMING=1 timeout 1500 python tools/bench_bam.py --reads ${BAM_READS:-64000000} --skip-host > $OUT/bench_bam.json 2> $OUT/bench_bam.err; echo "rc=$?"; cat $OUT/bench_bam.json; grep -E "device slab|count_alignments" $OUT/bench_bam.err | tail -8
echo "== bench (auto configs)"; timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -5 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print("c2: value %.3e e2e %.3e kernel ms/launch %.4f frac %.4f" % (d["value"], d["e2e"]["value"], r["kernel_ms_per_launch"], r["frac"]))
for k,v in d.get("configs",{}).items():
    print(k, json.dumps({kk:vv for kk,vv in v.items() if kk in ("value","e2e","file_to_file","check","ms_per_step","shuffled","sorted")})[:1600])
PY
