#!/usr/bin/env python
"""Secondary benchmark: BAM file -> count tables, end to end.

Writes a synthetic BAM (device-generated reads, native encoder), then times (a) the native decoder alone and
(b) ``counting.count_alignments`` on the file: BGZF inflate + SoA build on host threads, H2D, kernels, tables.
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

from mapdamage_b200 import counting, synth  # noqa: E402
from mapdamage_b200.bamio import BamReader, BamWriter  # noqa: E402
from mapdamage_b200.engine import DamageEngine  # noqa: E402
from mapdamage_b200.samtext import SamHeader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=8_000_000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--level", type=int, default=1)
    args = ap.parse_args()
    reference = synth.make_reference([1_000_000], seed=5)
    header = SamHeader()
    header.add("@HD\tVN:1.6\tSO:unsorted")
    header.add("@SQ\tSN:chr1\tLN:1000000")
    tmp = Path(tempfile.mkdtemp(prefix="mdg_bam_"))
    bam, fasta = tmp / "reads.bam", tmp / "ref.fa"
    reference.write_fasta(fasta)
    t0 = time.perf_counter()
    with DamageEngine(max_reads=1024) as engine:
        engine.set_reference(reference)
        with BamWriter(bam, header, threads=args.threads, level=args.level) as writer:
            done = 0
            while done < args.reads:
                n = min(1 << 21, args.reads - done)
                dev = engine.synth_batch(n, seed=100 + done, length=(100, 100), with_qual=True)
                writer.write_soa(engine.download(dev), first_index=done)
                dev.free()
                done += n
    t_write = time.perf_counter() - t0
    size = bam.stat().st_size

    def decode(device):
        t0 = time.perf_counter()
        with BamReader(bam, threads=args.threads, merge_libraries=True, device=device) as reader:
            buffers = reader.buffers(1 << 20, with_qual=False)
            n = 0
            while True:
                batch = reader.read_batch(buffers=buffers)
                if batch is None:
                    break
                n += batch.n
            blocks = reader.device_blocks
        assert n == args.reads
        return time.perf_counter() - t0, blocks

    t_decode, _ = decode(None)
    decode(0)  # warm-up: CUDA context, pinned slabs
    t_decode_gpu, gpu_blocks = decode(0)

    counting.count_alignments(bam, fasta, merge_libraries=True, batch_reads=1 << 18)  # warm-up: CUDA context, page cache
    t0 = time.perf_counter()
    misincorp, _, lg = counting.count_alignments(bam, fasta, merge_libraries=True, batch_reads=1 << 20)
    t_count = time.perf_counter() - t0
    assert sum(sum(t.values()) for t in lg.data[("*", "*")].values()) == args.reads
    os.environ["MDG_BAM_GPU"] = "1"
    t0 = time.perf_counter()
    counting.count_alignments(bam, fasta, merge_libraries=True, batch_reads=1 << 20)
    t_count_gpu = time.perf_counter() - t0
    del os.environ["MDG_BAM_GPU"]
    print(json.dumps({
        "metric": "reads/sec (BAM file -> count tables, end to end)", "reads": args.reads, "bam_bytes": size,
        "bytes_per_read_compressed": size / args.reads, "host_threads": args.threads or os.cpu_count(),
        "encode_reads_per_s": args.reads / t_write, "decode_only_reads_per_s": args.reads / t_decode,
        "decode_only_gpu_inflate_reads_per_s": args.reads / t_decode_gpu, "blocks_inflated_on_gpu": gpu_blocks,
        "end_to_end_reads_per_s": args.reads / t_count, "end_to_end_s": t_count,
        "end_to_end_gpu_inflate_reads_per_s": args.reads / t_count_gpu,
    }))
    for p in (bam, fasta):
        p.unlink()
    tmp.rmdir()


if __name__ == "__main__":
    main()
