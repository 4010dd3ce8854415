#!/bin/bash
# final single-GPU visit of the round: all GPU tests, smoke, the two bench arms, ncu launch list and full capture
OUT=gpurun_out/${1:-full}; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -2 $OUT/smoke.log
echo "== bench"; timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-600 $OUT/bench_ref.json
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print("c2: value %.3e e2e %.3e kernel ms/launch %.4f frac %.4f" % (d["value"], d["e2e"]["value"], r["kernel_ms_per_launch"], r["frac"]))
for k,v in d.get("configs",{}).items():
    print(k, json.dumps({kk:vv for kk,vv in v.items() if kk in ("value","e2e","file_to_file","check","ms_per_step","shuffled","sorted")})[:1500])
PY
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --cpu-sample 100000 --configs none > $OUT/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full capture (count_planes_kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:count_planes -s 13 -c 1 -f -o $OUT/prof_planes \
    python bench.py --steps 1 --warmup 1 --no-e2e --cpu-sample 100000 --configs none > $OUT/ncu_full.log 2>&1; echo "rc=$?"

echo "== ncu full capture on the 3.1 Gbp genome (shuffled reads)"
timeout 900 ncu --set full --clock-control none -k regex:count_planes -s 4 -c 1 -f -o $OUT/prof_planes_g3 python tools/prof_g3.py > $OUT/ncu_g3.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_g3.log
echo "== shapes"; timeout 600 python tools/bench_shapes.py > $OUT/shapes.txt 2>&1; cat $OUT/shapes.txt
ls -la $OUT
