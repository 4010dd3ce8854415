#!/usr/bin/env python
"""Secondary benchmark: BAM file -> count tables and BAM file -> rescaled BAM file, end to end.

Writes a synthetic BAM (device-generated reads, native encoder), then times the decoders alone, ``counting.count_alignments``
and ``rescale.rescale_qual`` on the file, once through the GPU decoder / encoder (the default) and once through the host
threads (MDG_BAM_HOST=1).
"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from mapdamage_b200 import counting, rescale, synth  # noqa: E402
from mapdamage_b200.bamio import BamReader, BamWriter, DeviceBamStream  # noqa: E402
from mapdamage_b200.engine import DamageEngine  # noqa: E402
from mapdamage_b200.samtext import SamHeader  # noqa: E402


def write_model(folder):
    """A synthetic Stats_out_MCMC_correct_prob.csv of the shape rescale.py:23-46 reads (24 rows)."""
    rows = ['"","Position","C.T","G.A"']
    k = 1
    for p in list(range(1, 13)) + list(range(-12, 0)):
        ct = 0.9 * 0.67 ** (p - 1) if p > 0 else 0.021
        ga = 0.013 if p > 0 else 0.85 * 0.6 ** (-p - 1)
        rows.append('"%d",%d,%.6f,%.6f' % (k, p, ct, ga))
        k += 1
    (Path(folder) / "Stats_out_MCMC_correct_prob.csv").write_text("\n".join(rows) + "\n")


def make_bam(path, fasta, reads, threads=0, level=1, seed=100, folder=None):
    reference = synth.make_reference([1_000_000], seed=5)
    header = SamHeader()
    header.add("@HD\tVN:1.6\tSO:unsorted")
    header.add("@SQ\tSN:chr1\tLN:1000000")
    reference.write_fasta(fasta)
    t0 = time.perf_counter()
    with DamageEngine(max_reads=1024) as engine:
        engine.set_reference(reference)
        with BamWriter(path, header, threads=threads, level=level) as writer:
            done = 0
            while done < reads:
                n = min(1 << 21, reads - done)
                dev = engine.synth_batch(n, seed=seed + done, length=(100, 100), with_qual=True)
                writer.write_soa(engine.download(dev), first_index=done)
                dev.free()
                done += n
    return reference, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=8_000_000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--dir", default=None, help="where the files go (default: /dev/shm when it exists)")
    ap.add_argument("--skip-host", action="store_true")
    args = ap.parse_args()
    base = args.dir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    tmp = Path(tempfile.mkdtemp(prefix="mdg_bam_", dir=base))
    bam, fasta = tmp / "reads.bam", tmp / "ref.fa"
    reference, t_write = make_bam(bam, fasta, args.reads, args.threads, args.level)
    write_model(tmp)
    size = bam.stat().st_size
    out = {"metric": "reads/sec (BAM file -> count tables / rescaled BAM, end to end)", "reads": args.reads, "bam_bytes": size,
           "bytes_per_read_compressed": size / args.reads, "host_threads": args.threads or os.cpu_count(),
           "synthetic_file_written_reads_per_s": args.reads / t_write, "dir": str(tmp)}

    def decode_host():
        t0 = time.perf_counter()
        with BamReader(bam, threads=args.threads, merge_libraries=True) as reader:
            buffers = reader.buffers(1 << 20, with_qual=False)
            n = 0
            while True:
                batch = reader.read_batch(buffers=buffers)
                if batch is None:
                    break
                n += batch.n
        assert n == args.reads
        return time.perf_counter() - t0

    def decode_device():
        with DamageEngine(max_reads=0) as engine:
            t0 = time.perf_counter()
            with DeviceBamStream(engine, bam, merge_libraries=True, with_qual=False) as stream:
                n = sum(dev.n for dev in stream)
                stats = stream.stats()
            dt = time.perf_counter() - t0
        assert n == args.reads
        return dt, stats

    if not args.skip_host:
        out["decode_only_host_reads_per_s"] = args.reads / decode_host()
    decode_device()  # warm-up: CUDA context, page cache
    dt, stats = decode_device()
    out["decode_only_device_reads_per_s"] = args.reads / dt
    out["decode_only_device_stats"] = stats

    def count(host):
        if host:
            os.environ["MDG_BAM_HOST"] = "1"
        try:
            t0 = time.perf_counter()
            _, _, lg = counting.count_alignments(bam, fasta, merge_libraries=True, batch_reads=1 << 20)
            dt = time.perf_counter() - t0
        finally:
            os.environ.pop("MDG_BAM_HOST", None)
        assert sum(sum(t.values()) for t in lg.data[("*", "*")].values()) == args.reads
        return dt

    count(False)  # warm-up
    out["file_to_tables_device_reads_per_s"] = args.reads / count(False)
    if not args.skip_host:
        out["file_to_tables_host_reads_per_s"] = args.reads / count(True)

    def rescale_file(host, name):
        if host:
            os.environ["MDG_BAM_HOST"] = "1"
        timings = {}
        options = argparse.Namespace(folder=tmp, filename=bam, rescale_out=tmp / name, rescale_length_5p=12,
                                     rescale_length_3p=12, timings=timings)
        try:
            t0 = time.perf_counter()
            rc = rescale.rescale_qual(fasta, options)
            dt = time.perf_counter() - t0
        finally:
            os.environ.pop("MDG_BAM_HOST", None)
        assert rc == 0
        return dt, timings, (tmp / name).stat().st_size

    rescale_file(False, "warm.bam")
    (tmp / "warm.bam").unlink()
    dt, timings, out_size = rescale_file(False, "rescaled_dev.bam")
    out["file_to_rescaled_file_device_reads_per_s"] = args.reads / dt
    out["file_to_rescaled_file_device_s"] = dt
    out["file_to_rescaled_file_device_stages"] = timings
    out["rescaled_bam_bytes_device"] = out_size
    if not args.skip_host:
        dt, _, out_size = rescale_file(True, "rescaled_host.bam")
        out["file_to_rescaled_file_host_reads_per_s"] = args.reads / dt
        out["rescaled_bam_bytes_host"] = out_size
    print(json.dumps(out))
    for p in tmp.iterdir():
        p.unlink()
    tmp.rmdir()


if __name__ == "__main__":
    main()
