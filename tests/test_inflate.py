"""The native DEFLATE decoder behind the BAM reader (``csrc/mdg_inflate.cpp``) against zlib: every block type,
long codes, runs, multi-block streams, short output buffers, truncated and corrupted input."""
import ctypes as C
import random
import zlib

import pytest

from mapdamage_b200 import _native


def inflate(raw, cap):
    lib = _native.load()
    out = (C.c_uint8 * max(cap, 1))()
    n = lib.mdg_inflate_raw(raw, len(raw), out, cap)
    return n, bytes(out[:max(n, 0)])


def deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return c.compress(data) + c.flush()


def payloads():
    rng = random.Random(1)
    for n in (0, 1, 2, 5, 100, 1000, 65280, 150_000):
        yield "random", bytes(rng.getrandbits(8) for _ in range(n))
        yield "acgt", bytes(rng.choice(b"ACGT") for _ in range(n))
        yield "run", b"A" * n
        yield "period7", (b"abcdefg" * (n // 7 + 1))[:n]
        yield "mostly_two", bytes(rng.choice(b"AB") if rng.random() < .9 else rng.getrandbits(8) for _ in range(n))
        yield "skewed", bytes(min(255, int(rng.expovariate(0.05))) for _ in range(n))  # long code lengths


@pytest.mark.parametrize("strategy", [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])
@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_round_trips_like_zlib(level, strategy):
    for name, data in payloads():
        raw = deflate(data, level, strategy)
        n, out = inflate(raw, len(data))
        assert n == len(data) and out == data, (name, len(data))
        if data:
            assert inflate(raw, len(data) - 1)[0] < 0, "output one byte short must fail"
        if len(raw) > 3:
            n, out = inflate(raw[:len(raw) // 2], len(data))
            assert n < 0 or out == data[:n]


def test_multi_block_stream():
    rng = random.Random(2)
    parts = [bytes(rng.getrandbits(8) for _ in range(1000)), b"A" * 5000, bytes(rng.choice(b"ACGT") for _ in range(30_000))]
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    raw = b"".join(c.compress(p) + c.flush(zlib.Z_FULL_FLUSH) for p in parts) + c.flush()
    n, out = inflate(raw, sum(map(len, parts)))
    assert out == b"".join(parts)


def test_damaged_streams_fail_or_stay_in_bounds():
    rng = random.Random(3)
    data = bytes(rng.choice(b"ACGTN") for _ in range(50_000))
    raw = deflate(data, 6)
    for _ in range(2000):
        hurt = bytearray(raw)
        hurt[rng.randrange(len(hurt))] ^= 1 << rng.randrange(8)
        n, out = inflate(bytes(hurt), len(data))
        assert n <= len(data)


def test_reader_takes_either_inflater(tmp_path, monkeypatch):
    """The BAM reader gives the same batches with the native decoder and with zlib only (MDG_BAM_ZLIB=1)."""
    import numpy as np

    import bam_py
    from mapdamage_b200.bamio import BamReader
    from mapdamage_b200.samtext import read_sam
    from conftest import GOLDEN

    header, records = read_sam(GOLDEN / "fuzz_0_l70_a10_q0" / "input.sam")
    bam_py.write_bam(tmp_path / "in.bam", header, records, block_bytes=3000)
    got = {}
    for env in ("0", "1"):
        monkeypatch.setenv("MDG_BAM_ZLIB", env)
        with BamReader(tmp_path / "in.bam", merge_libraries=True) as reader:
            batch = reader.read_batch()
        got[env] = batch
    for name in ("flag", "pos", "l_seq", "cigar", "seq4", "qual"):
        assert np.array_equal(getattr(got["0"], name), getattr(got["1"], name)), name
