#!/usr/bin/env python
"""Kernel time of the counting pass over one resident 4 M-read batch, for several read shapes (diagnostics)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mapdamage_b200 import synth  # noqa: E402
from mapdamage_b200.engine import DamageEngine  # noqa: E402

SHAPES = {
    "se100": dict(length=(100, 100)),
    "se50-150": dict(length=(50, 150)),
    "se30-90": dict(length=(30, 90)),
    "se50-150+clips": dict(length=(50, 150), mix=(9, 0, 0, 1)),
    "se50-150+indels": dict(length=(50, 150), mix=(8, 1, 1, 0)),
    "pe50-150": dict(length=(50, 150), paired=True),
    "c3 1 lib": dict(length=(50, 150), mix=(7, 1, 1, 1), paired=True),
    "c3 2 libs": dict(length=(50, 150), mix=(7, 1, 1, 1), paired=True, n_libs=2),
    "se100 2 libs": dict(length=(100, 100), n_libs=2),
    "se100 -Q 20": dict(length=(100, 100), min_qual=20),
    "se50-150 -Q 20": dict(length=(50, 150), min_qual=20),
}

reference = synth.make_reference([1_000_000], seed=5)
n = 4_000_000
only = sys.argv[1:]  # shape names; none: all of them
for name, kw in SHAPES.items():
    if only and name not in only:
        continue
    kw = dict(kw)
    min_qual = kw.pop("min_qual", 0)
    with DamageEngine(n_libraries=kw.get("n_libs", 1), max_reads=1024, min_qual=min_qual) as engine:
        engine.set_reference(reference)
        dev = engine.synth_batch(n, seed=7, with_qual=min_qual > 0, **kw)
        for _ in range(3):
            engine.count_resident(dev)
        engine.sync()
        engine.kernel_ms()
        for _ in range(5):
            engine.count_resident(dev)
        ms = engine.kernel_ms() / 5
        print("%-18s %.3f ms per 4 M reads  = %.2f G reads/s" % (name, ms, n / ms / 1e6))
        dev.free()
