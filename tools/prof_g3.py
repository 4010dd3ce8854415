#!/usr/bin/env python
"""A few counting launches over 4.17 M reads on a 3.1 Gbp synthetic genome (for ncu: DRAM bytes per read when the
genome does not fit L2).  Usage: prof_g3.py [sorted]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mapdamage_b200.engine import DamageEngine  # noqa: E402

order = sys.argv[1] if len(sys.argv) > 1 else "shuffled"
n = 4_166_666
with DamageEngine(lg_bins=8192, max_reads=0) as engine:
    engine.synth_reference([100_000_000] * 31, seed=20260101)
    dev = engine.synth_batch(n, seed=20263001, length=(100, 100), with_qual=False, sorted_positions=order == "sorted")
    for _ in range(3):
        engine.count_resident(dev)
    engine.sync()
    engine.kernel_ms()
    for _ in range(5):
        engine.count_resident(dev)
    ms = engine.kernel_ms() / 5
    print("g3 %s: %.3f ms per %d reads = %.2f G reads/s" % (order, ms, n, n / ms / 1e6))
