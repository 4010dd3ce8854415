#!/bin/bash
# compute-sanitizer memcheck over the smoke pass and a small counting run with every planes variant
OUT=gpurun_out/${1:-mc}; mkdir -p $OUT
cat > /tmp/mc.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, oracle
from mapdamage_b200 import synth
from mapdamage_b200.engine import DamageEngine
reference = synth.make_reference([200_000, 3_000], seed=3, other_rate=0.001)
for n_libs in (1, 2):
    batch = synth.simulate_reads(reference, 30_000, seed=4 + n_libs, length=(35, 140), mix=(5, 2, 2, 1), paired=True, n_libs=n_libs, read_n_rate=0.01)
    want = oracle.count(batch, reference, n_lib=n_libs, lg_bins=8192, threads=2)
    with DamageEngine(n_libraries=n_libs, max_reads=0) as engine:
        engine.set_reference(reference)
        dev = engine.upload(batch)
        engine.count_resident(dev)
        got = engine.tables()
    assert all(np.array_equal(a, b) for a, b in zip(got, want)), n_libs
    uni = synth.simulate_reads(reference, 20_000, seed=9, length=(100, 100), n_libs=n_libs)
    want = oracle.count(uni, reference, n_lib=n_libs, lg_bins=8192, threads=2)
    with DamageEngine(n_libraries=n_libs, max_reads=0) as engine:
        engine.set_reference(reference)
        dev = engine.upload(uni)
        engine.count_resident(dev)
        got = engine.tables()
    assert all(np.array_equal(a, b) for a, b in zip(got, want)), n_libs
print("memcheck workload ok")
PY
for v in "MDG_PLANES_INDELS=1" "MDG_PLANES_INDELS=0 MDG_PLANES_GATHER=1 MDG_PLANES_PREFETCH=3"; do
echo "== synccheck $v"; env $v timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_$(echo $v | tr ' =' '__').log python /tmp/mc.py 2>&1 | tail -2; echo "rc=$?"
for f in $OUT/synccheck_*.log; do tail -n 3 $f; done
done
