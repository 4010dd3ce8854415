#!/usr/bin/env python
"""Headline benchmark: aligned reads/s of the misincorporation + composition counting pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one pass of the counting path over the whole workload (BASELINE.json
configs[1]: 50 M synthetic 100 bp single-end reads with C->T / G->A damage on a 1 Mb random
reference, ``-l 70 -a 10 -Q 0``), cut into batches that fit the 32-bit offsets of one
``mdg_batch``.  With N > 1 every rank counts its own 50 M-read shard (weak scaling) and each
step ends with the NCCL all-reduce of the count tables.

* ``value``  -- whole-job reads/s with the batches resident in HBM, timed with CUDA events on
  the stream the kernels run on, max over ranks.
* ``e2e``    -- the same pass through the public API (``DamageEngine.count`` ->
  ``mdg_count_submit``) from pinned HOST batches: host->device copies, kernels and the
  device->host read of the tables are all inside the timed region.
* ``roofline`` -- algorithmic bytes per launch / mean device time of a counting launch, against
  the measured HBM copy bandwidth in MEASURED_PEAKS.json.
* ``cpu_baseline`` -- the CPU oracle (a C port of the reference's algorithm, oracle/) timed on
  this box's host cores over a bounded sample of the same workload.

``--impl reference`` times the CPU implementation alone (the reference itself is pure Python
on pysam and cannot travel to the GPU box; the arm runs the oracle port on every host thread).
The oracle is only ever the checker / the CPU baseline here, never the thing measured as ours.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

# stdout carries one JSON line and nothing else: whatever libraries write to file descriptor 1 (NCCL prints its
# version banner there when NCCL_DEBUG is VERSION or WARN, as it is on the GPU boxes) is sent to stderr, and the result
# line goes to a private copy of the original stdout
RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    RESULT_OUT.write(json.dumps(line) + "\n")
    RESULT_OUT.flush()


ROOT = Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "oracle"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

METRIC = "aligned reads/sec (misincorporation+composition pass)"
UNIT = "reads/s"
LENGTH, AROUND = 70, 10
READ_LEN = 100
REF_BASES = 1_000_000
N_LIBS_C3 = 2
# 15 B of record fields + 4 B per CIGAR op + packed read + packed reference span with flanks
ALGO_BYTES_PER_READ = 15 + 4 * 1 + (READ_LEN + 1) // 2 + (READ_LEN + 2 * AROUND + 1) // 2  # = 129


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--reads", type=int, default=50_000_000, help="reads per GPU per step")
    ap.add_argument("--batch-reads", type=int, default=1 << 22)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--cpu-sample", type=int, default=4_000_000, help="reads timed on the CPU oracle")
    ap.add_argument("--check-sample", type=int, default=200_000, help="reads checked against the oracle")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", choices=("c2", "c3"), default="c2",
                    help="c2 = configs[1] (the headline); c3 = configs[2]: 50-150 bp pairs, mixed CIGARs, two libraries")
    ap.add_argument("--configs", default="auto",
                    help="other BASELINE.json configurations measured after the headline and reported under \"configs\": "
                         "comma list of c3,c4,c5,g3, 'none', or 'auto' (c3, c4 and g3 on one GPU, c5 on eight)")
    ap.add_argument("--c4-reads", type=int, default=200_000_000, help="reads per step of the rescale configuration")
    ap.add_argument("--c4-file-reads", type=int, default=16_000_000,
                    help="reads of the BAM file the file -> rescaled-file leg of c4 runs on")
    return ap.parse_args()


# ---------------------------------------------------------------------------
class ClockSampler:
    """``nvidia-smi`` clocks and throttle reasons sampled while a region is timed."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, device):
        self.rows = []
        self.proc = None
        self.device = device

    def __enter__(self):
        if self.device is None:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in line.split(",")])

    def wait_for_samples(self, n=1, timeout=4.0):
        """nvidia-smi takes a moment to print its first line; the load loop should not start before it."""
        end = time.perf_counter() + timeout
        while len(self.rows) < n and time.perf_counter() < end and self.proc is not None:
            time.sleep(0.02)

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=5)
        return False

    def summary(self, t0=None, t1=None):
        """Median SM clock and the throttle reasons seen while the load ran (``t0``..``t1``: the part of the
        samples inside the timed region is reported separately; the load before it is the same loop)."""
        sm, mx, reasons, inside = [], [], set(), 0
        for row in self.rows:
            if len(row) < 8:
                continue
            try:
                sm.append(float(row[1]))
                mx.append(float(row[2]))
            except ValueError:
                continue
            if t0 is not None and t0 <= row[0] <= t1 + 0.1:
                inside += 1
            for name, state in zip(self.REASONS, row[4:8]):
                if state.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "samples_in_timed_region": inside,
                "window": "sampled every 100 ms over the same counting loop: >= 1 s of untimed passes, then the timed steps"}


def measured_peak():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.is_file():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload):
    """DRAM bytes per counting launch from the committed ncu capture of THIS workload (profiles/traffic.json is keyed by
    workload); None for a workload that was never captured -- a figure measured on another shape is not reported."""
    path = ROOT / "profiles" / "traffic.json"
    if path.is_file():
        try:
            return json.loads(path.read_text()).get(workload)
        except ValueError:
            pass
    return None


WORKLOADS = {
    "c2": "configs[1]: 50M x 100bp SE aDNA reads (C->T/G->A damage), 1 Mb reference, -l 70 -a 10 -Q 0",
    "c3": "configs[2]: 50M x 50-150bp PE aDNA reads, mixed CIGARs (70% match, 10% each insertion / deletion / "
          "soft clips), 2 libraries, 1 Mb reference, -l 70 -a 10 -Q 0",
    "c4": "configs[3]: 200M x 100bp SE reads with qualities, --rescale pass producing a rescaled BAM",
    "c5": "configs[4]: 1B x 100bp SE aDNA reads sharded over 8 GPUs (125M per rank), NCCL all-reduce of the count tables",
    "g3": "the reads of configs[1] on a 3.1 Gbp genome (31 contigs of 100 Mbp, made on the device): reference gathers "
          "come from DRAM instead of L2; reads in random order and in coordinate order",
}
# the reference's own Python loop (main.py:165-220) cannot run on the GPU box (pysam is not installable, /root/reference
# does not travel); this is the figure measured in the survey container through oracle/pysam_shim.py
PYTHON_REFERENCE = {"value": 18200.0, "unit": UNIT, "cores": 1,
                    "provenance": "unmodified reference main() under oracle/pysam_shim.py, 200 k reads, one core of the "
                                  "survey container (SURVEY.md section 6); not re-measured on this box: kind 'reference' "
                                  "is not obtainable there (no pysam wheel, no network, /root/reference absent)"}


# ---------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The CPU implementation alone, on every host thread (rank 0 only)."""
    if rank != 0:
        return
    import oracle
    from mapdamage_b200 import synth

    cores = os.cpu_count() or 1
    sample = min(args.reads, 2_000_000)
    reference = synth.make_reference([REF_BASES], seed=args.seed)
    batch = synth.simulate_reads(reference, sample, seed=args.seed + 1, length=(READ_LEN, READ_LEN),
                                 with_qual=False, threads=min(cores, 16))
    for _ in range(args.warmup):
        oracle.count(batch, reference, length=LENGTH, around=AROUND, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.count(batch, reference, length=LENGTH, around=AROUND, threads=cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    what = ("each step counts a bounded sample of %d reads of the workload (host-generated with the same read shape, "
            "seed %d); throughput does not depend on the sample size" % (sample, args.seed + 1))
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64",
        "data": "synthetic",
        "config": {"workload": WORKLOADS["c2"], "reads_per_gpu_per_step": args.reads, "seed": args.seed,
                   "sample": what, "sample_reads_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": what,
                         "note": "C port of the reference's Python loop (oracle/mdg_oracle.c), pinned to the "
                                 "unmodified reference by tests/golden",
                         "python_reference": PYTHON_REFERENCE},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---------------------------------------------------------------------------
class Ranks:
    """torch.distributed plumbing of one bench process (one per GPU)."""

    def __init__(self, rank, local_rank, world):
        import torch

        self.torch, self.rank, self.local_rank, self.world = torch, rank, local_rank, world
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok):
        return self.max(0.0 if ok else 1.0) == 0.0

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def same_tables(got, want):
    return all(np.array_equal(a, b) for a, b in zip(got, want))


def counting_config(args, ranks, workload, reads, steps, warmup, sample_clocks=False, with_cpu=False, with_e2e=True,
                    e2e_host_reads=None):
    """One counting workload measured the way the headline is: parity first, then ``value`` (resident batches, CUDA
    events), ``e2e`` (pinned host batches through DamageEngine.count) and the roofline of the dominant kernel."""
    from mapdamage_b200 import multigpu, synth
    from mapdamage_b200.engine import DamageEngine

    rank, world, local_rank = ranks.rank, ranks.world, ranks.local_rank
    c3 = workload == "c3"
    n_lib = N_LIBS_C3 if c3 else 1
    n_batches = max(1, -(-reads // args.batch_reads))
    sizes = [reads // n_batches + (1 if i < reads % n_batches else 0) for i in range(n_batches)]
    if c3:
        sizes = [n + (n & 1) for n in sizes]
    cap = max(sizes)
    engine = DamageEngine(length=LENGTH, around=AROUND, min_qual=0, n_libraries=n_lib, lg_bins=8192,
                          device=local_rank, n_slots=2, max_reads=cap + 1, max_cigar_ops=3 * cap + 3 if c3 else cap,
                          max_bases=(cap + 1) * (152 if c3 else READ_LEN + (READ_LEN & 1)))
    reference = synth.make_reference([REF_BASES], seed=args.seed)
    engine.set_reference(reference)
    if world > 1:
        multigpu.connect(engine, ranks.dist, rank, world)
    shape = dict(length=(50, 150), mix=(7, 1, 1, 1), paired=True, n_libs=N_LIBS_C3) if c3 else dict(length=(READ_LEN, READ_LEN))
    resident = [engine.synth_batch(n, seed=args.seed + 1000 * rank + i, with_qual=False, **shape)
                for i, n in enumerate(sizes)]
    total = sum(sizes)

    def one_pass():
        for dev in resident:
            engine.count_resident(dev)
        if world > 1:
            engine.allreduce_tables()

    # ---- correctness before speed ----
    host0 = engine.download(resident[0])
    check = {}
    if rank == 0:
        import oracle

        threads = min(16, os.cpu_count() or 1)
        # (1) the streamed path (mdg_count_submit) on a sample
        sample = host0.slice(0, min(args.check_sample, host0.n))
        want = oracle.count(sample, reference, length=LENGTH, around=AROUND, lg_bins=8192, n_lib=n_lib, threads=threads)
        engine.reset()
        engine.count(sample)
        if not same_tables(engine.tables(), want):
            raise SystemExit("bench[%s]: tables differ from the oracle on the %d-read sample (submit path)" % (workload, sample.n))
        check["oracle_sample_reads_submit_path"] = sample.n
        # (2) the resident path that is timed: one WHOLE resident batch, every table cell against the oracle
        want = oracle.count(host0, reference, length=LENGTH, around=AROUND, lg_bins=8192, n_lib=n_lib, threads=threads)
        engine.reset()
        engine.count_resident(resident[0])
        if not same_tables(engine.tables(), want):
            raise SystemExit("bench[%s]: tables differ from the oracle on the whole resident batch (%d reads)" % (workload, host0.n))
        check["oracle_whole_resident_batch_reads"] = host0.n
    # (3) every batch at full size: size-independent invariants
    engine.reset()
    for dev in resident:
        engine.count_resident(dev)
    mis, comp, lg = local = engine.tables()
    if not c3:
        # every read is 100M on an N-free genome: each end / position sees every read exactly once
        per_pos = mis[0, :, :, 0:4, :].sum(axis=(1, 2))
        ok = (np.all(per_pos == total) and int(lg.sum()) == total and int(lg[0, 1, :, READ_LEN].sum()) == total
              and np.all(comp[0, :, :, :, :LENGTH].sum(axis=(1, 2)) == total))
    else:
        # every record is a proper pair: one histogram entry per first mate; reads of >= 50 bp fill position 1 of
        # both ends with a read base
        ok = int(lg.sum()) == total // 2 and int(comp[:, :, :, :, 0].sum()) == 2 * total
    if not ok:
        raise SystemExit("bench[%s]: full-size invariants failed" % workload)
    check["full_size_invariants"] = "ok"
    # (4) several GPUs: the all-reduced tables equal the sum of the per-rank tables taken before the collective
    #     (summed independently through torch.distributed), on every rank
    if world > 1:
        engine.allreduce_tables()
        reduced = engine.tables()
        want = multigpu.sum_tables_host(ranks.dist, local)
        ok = same_tables(reduced, want)
        if not c3:
            ok = ok and bool(np.all(reduced[0][0, :, :, 0:4, :].sum(axis=(1, 2)) == world * total))
            ok = ok and int(reduced[2].sum()) == world * total
        engine.allreduce_tables()  # a second reduction must not add the sums to themselves
        ok = ok and same_tables(engine.tables(), reduced)
        if not ranks.all_true(ok):
            raise SystemExit("bench[%s]: all-reduced tables differ from the sum of the per-rank tables" % workload)
        check["allreduced_parity"] = "ok"
        check["allreduced_reads"] = world * total

    # ---- value: resident batches, CUDA events on the compute stream ----
    t_warm = time.perf_counter()
    for _ in range(warmup):
        one_pass()
    engine.sync()
    t_pass = ranks.max((time.perf_counter() - t_warm) / max(1, warmup))
    # the same number of extra passes on every rank (each pass ends in a collective): about one second of load
    n_sustain = int(min(2000, max(1, np.ceil(1.0 / max(t_pass, 1e-4))))) if sample_clocks else 0
    ranks.barrier()
    with ClockSampler(local_rank if sample_clocks else None) as clocks:
        # the timed region is tens of milliseconds, shorter than one nvidia-smi period: keep the same load
        # running untimed until the sampler has seen it, then time the K steps without a gap
        if sample_clocks:
            clocks.wait_for_samples(1)
        for _ in range(n_sustain):
            one_pass()
        engine.sync()
        engine.kernel_ms()
        launches0 = engine.launch_count()
        ranks.barrier()
        t_start = time.perf_counter()
        engine.event_record(0)
        for _ in range(steps):
            one_pass()
        engine.event_record(1)
        ms = engine.event_elapsed_ms()
        t_stop = time.perf_counter()
        ranks.barrier()
    ms = ranks.max(ms)
    launches = engine.launch_count() - launches0
    kernel_ms = engine.kernel_ms()
    value = world * total * steps / (ms * 1e-3)

    # ---- e2e: pinned host batches through the public API ----
    e2e = None
    if with_e2e:
        # e2e_host_reads: pin only that many reads and stream them repeatedly (same bytes and copies per read)
        n_host = len(resident) if e2e_host_reads is None else max(1, min(len(resident), -(-e2e_host_reads // cap)))
        host = [engine.download(dev, pinned=True) for dev in resident[:n_host]]
        order = [host[i % n_host] for i in range(len(resident))]
        h2d = sum(engine.h2d_bytes(b) for b in order)
        d2h = int(mis.nbytes + comp.nbytes + lg.nbytes)
        e2e_reads = sum(b.n for b in order)
        engine.reset()

        def e2e_pass():
            for b in order:
                engine.count(b)
            if world > 1:
                engine.allreduce_tables()
            return engine.tables()

        for _ in range(min(warmup, 3)):
            e2e_pass()
        ranks.barrier()
        engine.event_record(0)
        t0 = time.perf_counter()
        for _ in range(steps):
            tables = e2e_pass()
        engine.event_record(1)
        e2e_ms = engine.event_elapsed_ms()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ranks.barrier()
        e2e_ms = ranks.max(max(e2e_ms, wall_ms))
        e2e = {"value": world * e2e_reads * steps / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / steps,
               "timing": "max(CUDA events on the compute stream, host wall clock) around submit..fetch, max over ranks"}
        if n_host < len(resident):
            e2e["host_batches"] = "%d pinned batches streamed %d times per step" % (n_host, len(resident))
        del tables

    # ---- CPU baseline on a bounded sample (rank 0, single GPU runs only) ----
    cpu = None
    if with_cpu and rank == 0 and world == 1:
        import oracle

        sample = host0.slice(0, min(args.cpu_sample, host0.n))
        passes = 3 if sample.n >= 1_000_000 else 1  # about 10 s of CPU work at the default sample
        t0 = time.perf_counter()
        for _ in range(passes):
            oracle.count(sample, reference, length=LENGTH, around=AROUND, lg_bins=8192, n_lib=n_lib, threads=1)
        dt = time.perf_counter() - t0
        cpu = {"value": passes * sample.n / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "first %d reads of batch 0 of this workload, %d passes, one thread, %.1f s"
                         % (sample.n, passes, dt),
               "python_reference": PYTHON_REFERENCE}

    out = None
    if rank == 0:
        peak, peak_source = measured_peak()
        launch_ms = kernel_ms / (steps * n_batches)
        algo = ALGO_BYTES_PER_READ
        if c3:  # SURVEY 8(d) formula, averaged over batch 0
            algo = float(15 + 4 * host0.cigar.shape[0] / host0.n + (host0.l_seq.mean() + 1) / 2
                         + (host0.l_seq.mean() + 2 * AROUND + 1) / 2)
        bytes_per_launch = algo * total / n_batches
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        traffic = recorded_traffic("c2" if workload == "c5" else workload)
        out = {
            "value": value, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
            "config": {
                "workload": WORKLOADS[workload],
                "reads_per_gpu_per_step": total, "batches_per_step": n_batches, "seed": args.seed,
                "l2": "inputs larger than L2 (%.1f GB of resident batches per pass vs 126 MB)" % (total * 90 / 1e9),
                "parallelism": "reads sharded per GPU; NCCL all-reduce of the count tables per step" if world > 1
                else "single GPU",
            },
            "clocks": clocks.summary(t_start, t_stop) if sample_clocks else None,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                "traffic_source": None if traffic is None else traffic.get("source"),
                "peak_source": peak_source, "algorithmic_bytes_per_read": algo,
                "reads_per_launch": total / n_batches, "kernel_ms_per_launch": launch_ms,
                "kernel_share_of_step": kernel_ms / ms,
                "frac_of_nominal_8TBps": achieved / 8000.0,
            },
            "cpu_baseline": cpu,
            "check": check,
        }
    for dev in resident:
        dev.free()
    engine.close()
    return out


def gpu_arm(args, rank, local_rank, world):
    from mapdamage_b200.engine import bind_host_to_device

    cpus = bind_host_to_device(local_rank)  # pinned batches on the GPU's own NUMA node
    ranks = Ranks(rank, local_rank, world)
    main_cfg = counting_config(args, ranks, args.workload, args.reads, args.steps, args.warmup, sample_clocks=True,
                               with_cpu=True, with_e2e=not args.no_e2e)
    wanted = [c for c in args.configs.split(",") if c and c != "none"]
    if "auto" in wanted:
        # the other configurations BASELINE.json names: configs[2] and configs[3] are single-GPU cases,
        # configs[4] is the 8-GPU case
        wanted = (["c3", "c4", "g3"] if world == 1 else []) + (["c5"] if world == 8 else [])
    configs = {}
    side_steps, side_warm = max(2, min(args.steps, 5)), 3
    for name in wanted:
        if name == args.workload:
            continue
        if name == "c3":
            configs[name] = counting_config(args, ranks, "c3", args.reads, side_steps, side_warm,
                                            with_e2e=not args.no_e2e)
        elif name == "c5":
            # 1 B reads over the ranks: each rank counts its own 1e9 / world reads per step
            configs[name] = counting_config(args, ranks, "c5", 1_000_000_000 // world, side_steps, side_warm,
                                            with_e2e=not args.no_e2e, e2e_host_reads=args.reads)
        elif name == "c4":
            configs[name] = rescale_config(args, ranks)
        elif name == "g3":
            configs[name] = big_genome_config(args, ranks)
    if rank == 0:
        line = {"metric": METRIC, "n_gpus": world, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/int64", "data": "synthetic"}
        line.update(main_cfg)
        line["config"]["host_cpus_bound"] = None if cpus is None else len(cpus)
        line["configs"] = {k: v for k, v in configs.items() if v is not None}
        emit(line)
    ranks.close()


def big_genome_config(args, ranks):
    """the synthetic reference of configs[1] is 1 Mb and lives in L2; a human-sized genome does not.
    12.5 M reads of the configs[1] shape on a 3.1 Gbp genome, drawn uniformly (an unsorted file) and in coordinate
    order (a sorted one); parity on a sample against the oracle on the downloaded genome."""
    from mapdamage_b200.engine import DamageEngine

    if ranks.rank != 0:
        return None
    import oracle

    reads, n_batches = 12_500_000, 3
    per = reads // n_batches
    out = {"unit": UNIT, "config": {"workload": WORKLOADS["g3"], "reads_per_step": per * n_batches, "batches_per_step": n_batches,
                                    "genome_bases": 31 * 100_000_000}}
    with DamageEngine(length=LENGTH, around=AROUND, lg_bins=8192, device=ranks.local_rank, max_reads=0) as engine:
        names, lengths = engine.synth_reference([100_000_000] * 31, seed=args.seed)
        for order in ("shuffled", "sorted"):
            resident = [engine.synth_batch(per, seed=args.seed + 3000 + i, length=(READ_LEN, READ_LEN), with_qual=False,
                                           sorted_positions=order == "sorted") for i in range(n_batches)]
            if order == "shuffled":
                # parity: 100 k reads against the oracle on the genome as the device holds it
                reference = engine.reference_host(names, lengths)
                sample = engine.download(resident[0]).slice(0, 100_000)
                want = oracle.count(sample, reference, length=LENGTH, around=AROUND, lg_bins=8192, threads=min(16, os.cpu_count() or 1))
                sub = engine.upload(sample)
                engine.reset()
                engine.count_resident(sub)
                if not same_tables(engine.tables(), want):
                    raise SystemExit("bench[g3]: tables differ from the oracle on the 3.1 Gbp genome")
                sub.free()
                del reference
                out["check"] = {"oracle_sample_reads": sample.n}
            for _ in range(2):
                for dev in resident:
                    engine.count_resident(dev)
            engine.sync()
            engine.kernel_ms()
            steps = 5
            engine.event_record(0)
            for _ in range(steps):
                for dev in resident:
                    engine.count_resident(dev)
            engine.event_record(1)
            ms = engine.event_elapsed_ms()
            kernel_ms = engine.kernel_ms()
            out[order] = {"value": per * n_batches * steps / (ms * 1e-3), "kernel_ms_per_launch": kernel_ms / (steps * n_batches),
                          "reads_per_launch": per}
            for dev in resident:
                dev.free()
    out["value"] = out["shuffled"]["value"]
    traffic = recorded_traffic("g3")
    out["dram_bytes_per_read"] = None if traffic is None else traffic
    return out


def rescale_model_terms():
    """The synthetic damage model of configs[3] (24 rows of the shape rescale.py:23-46 reads)."""
    corr = {("C", "T", p): 0.9 * 0.67 ** (p - 1) for p in range(1, 13)}
    corr.update({("G", "A", -p): 0.85 * 0.6 ** (p - 1) for p in range(1, 13)})
    corr.update({("G", "A", p): 0.013 for p in range(1, 13)})
    corr.update({("C", "T", -p): 0.021 for p in range(1, 13)})
    return corr


def rescale_config(args, ranks):
    """configs[3]: the rescale pass (rescale.py:285-365) over 200 M reads, three ways:
    kernels alone on resident batches, through the public API from pinned host batches (only the changed quality
    bytes come back), and BAM file -> rescaled BAM file through ``rescale.rescale_qual`` (decode, rescale, re-emission
    and BGZF deflate on the GPU; the host reads one file and writes the other)."""
    import argparse as _argparse
    import shutil
    import tempfile

    from mapdamage_b200 import rescale, synth
    from mapdamage_b200.bamio import BamReader, BamWriter
    from mapdamage_b200.engine import DamageEngine
    from mapdamage_b200.rescale_model import RescaleModel
    from mapdamage_b200.samtext import SamHeader

    if ranks.rank != 0:
        return None
    import oracle

    corr = rescale_model_terms()
    model = RescaleModel(corr, 12, 12)
    reference = synth.make_reference([REF_BASES], seed=args.seed)
    n_batches = max(1, -(-args.c4_reads // args.batch_reads))
    per = args.c4_reads // n_batches
    total = per * n_batches
    out = {"unit": UNIT, "config": {"workload": WORKLOADS["c4"], "reads_per_step": total, "batches_per_step": n_batches,
                                    "model": "synthetic Stats_out_MCMC_correct_prob.csv, 12 positions per end"}}
    check = {}
    with DamageEngine(device=ranks.local_rank, n_slots=2, max_reads=per + 1, max_cigar_ops=per + 1,
                      max_bases=(per + 1) * READ_LEN) as engine:
        engine.set_reference(reference)
        engine.set_rescale_model(model)
        resident = [engine.synth_batch(per, seed=args.seed + 7000 + i, length=(READ_LEN, READ_LEN), with_qual=True)
                    for i in range(n_batches)]
        host0 = engine.download(resident[0])
        # parity first: 100 k reads, exact qualities, MR and status (the contract allows +-1 Phred; the LUT is exact)
        sample = host0.slice(0, min(100_000, host0.n))
        want_qual, want_mr, want_status, _, rc = oracle.rescale(sample, reference, corr)
        sub = engine.upload(sample)
        mr, status = engine.rescale_resident(sub, want_results=True)
        engine.sync()
        got = engine.download(sub)
        sub.free()
        n_b = sample.total_bases
        if rc != 0 or not (np.array_equal(status, want_status) and np.array_equal(mr[status == 1], want_mr[status == 1])
                           and np.array_equal(got.qual[:n_b], want_qual[:n_b])):
            raise SystemExit("bench[c4]: rescaled qualities / MR differ from the oracle on the %d-read sample" % sample.n)
        check["oracle_sample_reads"] = sample.n
        check["qualities"] = "exact"
        # ---- kernels alone: resident batches, CUDA events on the compute stream ----
        for dev in resident:  # one untimed pass: every batch gets its result arrays, the kernels their first launch
            engine.rescale_resident(dev)
        engine.sync()
        engine.kernel_ms()
        launches0 = engine.launch_count()
        engine.event_record(0)
        for dev in resident:
            engine.rescale_resident(dev)
        engine.event_record(1)
        ms = engine.event_elapsed_ms()
        kernel_ms = engine.kernel_ms()
        launches = engine.launch_count() - launches0
        peak, peak_source = measured_peak()
        algo = 15 + 8 + 4 + (READ_LEN + 1) // 2 + READ_LEN + (READ_LEN + 1) // 2 + READ_LEN + 4  # 331 B at 100 bp
        achieved = algo * total / (kernel_ms * 1e-3) / 1e9
        out.update({"value": total / (ms * 1e-3), "ms_per_step": ms, "steps": 1, "gpu_launches": int(launches),
                    "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                 "traffic": None, "peak_source": peak_source, "algorithmic_bytes_per_read": algo,
                                 "kernel_ms_per_launch": kernel_ms / n_batches, "reads_per_launch": per,
                                 "note": "qualities are rewritten in place: after the first pass the inputs are the "
                                         "rescaled qualities (the same columns are looked up)"}})
        for dev in resident:
            dev.free()
        # ---- e2e: pinned host batches through DamageEngine.rescale_sparse / rescale_collect ----
        if not args.no_e2e:
            n_host = min(4, n_batches)
            devs = [engine.synth_batch(per, seed=args.seed + 7000 + i, length=(READ_LEN, READ_LEN), with_qual=True)
                    for i in range(n_host)]
            host = [engine.download(dev, pinned=True) for dev in devs]
            for dev in devs:
                dev.free()
            outs = [(engine.arena.empty(per + 1, np.float32), engine.arena.empty(per + 1, np.uint8)) for _ in range(2)]
            scratch = (np.empty(host[0].total_bases // 4 + 4096, np.uint32), np.empty(host[0].total_bases // 4 + 4096, np.uint8))
            h2d = n_batches * engine.h2d_bytes(host[0], rescale=True)

            # the pinned batches are streamed again and again, so the changes are fetched but not written into them: a
            # patched batch would come back unchanged the next time round.  Writing them (a scatter over the quality
            # array, four host threads inside mdg_rescale_collect) is timed on its own below and added per batch.
            def e2e_pass(apply=False):
                pending, changed = None, 0
                for i in range(n_batches):
                    batch = host[i % n_host]
                    _, _, ticket = engine.rescale_sparse(batch, out=outs[i & 1])
                    if pending is not None:
                        changed += engine.rescale_collect(pending[0], pending[1], scratch, apply=apply)
                    pending = (ticket, batch)
                changed += engine.rescale_collect(pending[0], pending[1], scratch, apply=apply)
                engine.sync()
                return changed

            e2e_pass()
            t0 = time.perf_counter()
            changed = e2e_pass()
            dt = time.perf_counter() - t0
            # the scatter alone, on one batch (the last collect left its change list in `scratch`)
            n_last = changed // n_batches
            t1 = time.perf_counter()
            _, _, ticket = engine.rescale_sparse(host[0], out=outs[0])
            engine.rescale_collect(ticket, host[0], scratch, apply=False)
            t_fetch = time.perf_counter() - t1
            t1 = time.perf_counter()
            _, _, ticket = engine.rescale_sparse(host[1], out=outs[1])
            engine.rescale_collect(ticket, host[1], scratch, apply=True)
            t_apply = max(0.0, (time.perf_counter() - t1) - t_fetch)
            dt += n_batches * t_apply
            out["e2e"] = {"value": total / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                          "d2h_bytes_per_step": int(5 * total + 5 * changed), "ms_per_step": dt * 1e3,
                          "changed_quality_bytes_per_step": int(changed),
                          "host_batches": "%d pinned batches streamed %d times per step; only the quality bytes that "
                                          "changed come back (index + score); writing them into the host array is "
                                          "timed on one batch (%.1f ms for about %d bytes) and added for every batch"
                                          % (n_host, n_batches, t_apply * 1e3, n_last),
                          "timing": "host wall clock around submit..collect of every batch"}
    # ---- BAM file -> rescaled BAM file ----
    n_file = int(args.c4_file_reads)
    if n_file > 0:
        base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 600 * n_file else None
        tmp = Path(tempfile.mkdtemp(prefix="mdg_c4_", dir=base))
        try:
            bam, fasta = tmp / "reads.bam", tmp / "ref.fa"
            reference.write_fasta(fasta)
            rows = ['"","Position","C.T","G.A"']
            for k, pos in enumerate(list(range(1, 13)) + list(range(-12, 0))):
                rows.append('"%d",%d,%r,%r' % (k + 1, pos, corr[("C", "T", pos)], corr[("G", "A", pos)]))
            (tmp / "Stats_out_MCMC_correct_prob.csv").write_text("\n".join(rows) + "\n")
            header = SamHeader()
            header.add("@HD\tVN:1.6\tSO:unsorted")
            header.add("@SQ\tSN:%s\tLN:%d" % (reference.names[0], REF_BASES))
            with DamageEngine(device=ranks.local_rank, max_reads=1024) as engine:
                engine.set_reference(reference)
                with BamWriter(bam, header) as writer:
                    done = 0
                    while done < n_file:
                        n = min(1 << 21, n_file - done)
                        dev = engine.synth_batch(n, seed=args.seed + 9000 + done, length=(READ_LEN, READ_LEN), with_qual=True)
                        writer.write_soa(engine.download(dev), first_index=done)
                        dev.free()
                        done += n

            def run(name):
                timings = {}
                options = _argparse.Namespace(folder=tmp, filename=bam, rescale_out=tmp / name, rescale_length_5p=12,
                                              rescale_length_3p=12, timings=timings, device=ranks.local_rank)
                t0 = time.perf_counter()
                rc = rescale.rescale_qual(fasta, options)
                dt = time.perf_counter() - t0
                if rc != 0:
                    raise SystemExit("bench[c4]: rescale_qual failed on the synthetic BAM")
                return dt, timings

            run("warm.bam")  # CUDA context, page-locked buffers, page cache
            (tmp / "warm.bam").unlink()
            dt, timings = run("rescaled.bam")
            # round trip: the output re-read by the host decoder; every record there, in order; the first 100 k against the oracle
            with BamReader(bam, merge_libraries=True, apply_filter=False) as a, \
                    BamReader(tmp / "rescaled.bam", merge_libraries=True, apply_filter=False) as b:
                first_in = a.read_batch(max_reads=100_000, keep_raw=True)
                first_out = b.read_batch(max_reads=100_000, keep_raw=True)
                n_out, n_mr = first_out.n, int(first_out.has_mr.sum())
                while True:
                    part = b.read_batch(max_reads=1 << 20, keep_raw=True)
                    if part is None:
                        break
                    n_out += part.n
                    n_mr += int(part.has_mr.sum())
            want_qual, _, want_status, _, rc = oracle.rescale(first_in, reference, corr)
            n_b = first_in.total_bases
            same = (rc == 0 and n_out == n_file and np.array_equal(first_out.qual[:n_b], want_qual[:n_b])
                    and np.array_equal(first_out.has_mr.astype(np.uint8), want_status & 1)
                    and np.array_equal(first_out.flag, first_in.flag) and np.array_equal(first_out.pos, first_in.pos)
                    and np.array_equal(first_out.seq4, first_in.seq4))
            if not same:
                raise SystemExit("bench[c4]: the rescaled BAM does not read back as the input with the oracle's qualities")
            check["file_round_trip"] = "ok: %d records re-read by BamReader, %d with MR; first %d against the oracle" % (
                n_out, n_mr, first_in.n)
            out["file_to_file"] = {"value": n_file / dt, "unit": UNIT, "reads": n_file, "seconds": dt,
                                   "input_bam_bytes": bam.stat().st_size,
                                   "output_bam_bytes": (tmp / "rescaled.bam").stat().st_size, "stages": timings,
                                   "projected_seconds_for_200M_reads": 200e6 / (n_file / dt),
                                   "where": str(base or tempfile.gettempdir())}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    # the same leg at the configuration's own size is a separate, committed run (tools/gpu_c4_200m.sh: 55 GB of files,
    # 2.5 minutes with the making of the input): reported with its provenance, not measured here
    recorded = ROOT / "profiles" / "r02_bam_200M.json"
    if recorded.is_file():
        try:
            big = json.loads(recorded.read_text())
            out["file_to_file_at_200M_reads"] = {
                "value": big["file_to_rescaled_file_device_reads_per_s"], "unit": UNIT, "reads": big["reads"],
                "seconds": big["file_to_rescaled_file_device_s"], "input_bam_bytes": big["bam_bytes"],
                "output_bam_bytes": big["rescaled_bam_bytes_device"], "stages": big["file_to_rescaled_file_device_stages"],
                "file_to_tables_reads_per_s": big["file_to_tables_device_reads_per_s"],
                "source": "recorded: profiles/r02_bam_200M.json (tools/bench_bam.py --reads 200000000 --skip-host on one B200), not measured in this run"}
        except (KeyError, ValueError):
            pass
    out["check"] = check
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
