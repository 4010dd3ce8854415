#!/usr/bin/env python
"""Secondary benchmark: the rescale pass (BASELINE.json configs[3] shape: 100 bp SE reads with qualities).

Prints one JSON line: reads/s through ``DamageEngine.rescale`` from pinned host batches (host->device copy of the
records, kernel, device->host copy of the new qualities / MR / status inside the timed region), the summed kernel time,
and the CPU oracle on a sample.  Not the headline metric; kept for DESIGN.md.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "oracle"):
    sys.path.insert(0, str(p))

import oracle  # noqa: E402
from mapdamage_b200 import synth  # noqa: E402
from mapdamage_b200.engine import DamageEngine  # noqa: E402
from mapdamage_b200.rescale_model import RescaleModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=16_000_000)
    ap.add_argument("--batch-reads", type=int, default=1 << 21)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--paired", action="store_true")
    args = ap.parse_args()
    n_batches = -(-args.reads // args.batch_reads)
    sizes = [args.reads // n_batches] * n_batches
    corr = {("C", "T", p): 0.9 * 0.67 ** (p - 1) for p in range(1, 13)}
    corr.update({("G", "A", -p): 0.85 * 0.6 ** (p - 1) for p in range(1, 13)})
    corr.update({("G", "A", p): 0.013 for p in range(1, 13)})
    corr.update({("C", "T", -p): 0.021 for p in range(1, 13)})
    model = RescaleModel(corr, 12, 12)
    reference = synth.make_reference([1_000_000], seed=5)
    cap = max(sizes)
    with DamageEngine(max_reads=cap, max_cigar_ops=3 * cap, max_bases=152 * cap) as engine:
        engine.set_reference(reference)
        engine.set_rescale_model(model)
        kw = dict(length=(50, 150), mix=(7, 1, 1, 1), paired=True) if args.paired else dict(length=(100, 100))
        resident = [engine.synth_batch(n, seed=31 + i, with_qual=True, **kw) for i, n in enumerate(sizes)]
        host = [engine.download(dev, pinned=True) for dev in resident]
        for dev in resident:
            dev.free()
        most = max(b.total_bases for b in host)
        outs = [(engine.arena.empty(most, np.uint8), engine.arena.empty(cap, np.float32),
                 engine.arena.empty(cap, np.uint8)) for _ in range(2)]
        # parity on a sample first
        sample = host[0].slice(0, 100_000)
        want_qual, want_mr, want_status, _, rc = oracle.rescale(sample, reference, corr)
        qual, mr, status = engine.rescale(sample)
        engine.sync()
        assert rc == 0 and np.array_equal(status, want_status) and np.array_equal(mr[status == 1], want_mr[status == 1])
        # quality bytes of real bases only: the pad slot behind an odd-length read is unspecified
        starts, lens = sample.base_off.astype(np.int64), sample.l_seq.astype(np.int64)
        idx = np.repeat(starts, lens) + (np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens))
        assert np.array_equal(qual[idx], want_qual[idx])

        def one_pass():
            for i, b in enumerate(host):
                engine.rescale(b, out=outs[i & 1])
            engine.sync()

        for _ in range(args.warmup):
            one_pass()
        engine.kernel_ms()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one_pass()
        dt = time.perf_counter() - t0
        kernel_ms = engine.kernel_ms()
        total = sum(sizes)
        h2d = sum(engine.h2d_bytes(b, rescale=True) for b in host)
        d2h = sum(b.total_bases + 5 * b.n for b in host)
        t1 = time.perf_counter()
        oracle.rescale(host[0].slice(0, 1_000_000), reference, corr)
        cpu_dt = time.perf_counter() - t1
        print(json.dumps({
            "metric": "reads/sec (rescale pass)", "e2e_value": total * args.steps / dt, "unit": "reads/s",
            "kernel_only_value": total * args.steps / (kernel_ms * 1e-3), "kernel_ms_per_batch": kernel_ms / (args.steps * n_batches),
            "reads_per_step": total, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "algorithmic_GBps_kernel": 331 * total * args.steps / (kernel_ms * 1e-3) / 1e9,
            "cpu_oracle_reads_per_s_1_thread": 1_000_000 / cpu_dt, "paired": args.paired,
        }))


if __name__ == "__main__":
    main()
