"""BAM files decoded and encoded on the GPU (csrc/mdg_bamdev.cuh) against the host decoder
(csrc/mdg_bamio.cpp), the pure-Python codec written from the specification (tests/bam_py.py) and Python's gzip
module, which checks the CRC32 and ISIZE of every BGZF block the device deflated."""
import gzip

import numpy as np
import pytest

import bam_py
from conftest import GOLDEN
from mapdamage_b200 import synth
from mapdamage_b200.bamio import BamReader, BamWriter, DeviceBamStream
from mapdamage_b200.batch import BAMError, concatenate
from mapdamage_b200.engine import DamageEngine
from mapdamage_b200.rescale_model import RescaleModel
from mapdamage_b200.samtext import SamHeader, read_sam
from test_bamio import FIELDS

pytestmark = pytest.mark.gpu


def host_batch(path, merge, apply_filter):
    with BamReader(path, merge_libraries=merge, apply_filter=apply_filter) as reader:
        parts = list(reader)
        seen = reader.records_seen
    return concatenate(parts) if parts else None, seen


def assert_same_arrays(got, want):
    assert got.n == want.n
    for name in FIELDS:
        assert np.array_equal(getattr(got, name), getattr(want, name)), name
    assert np.array_equal(got.qual[:got.total_bases], want.qual[:want.total_bases])


def synthetic_bam(path, n, seed, read_groups=("rgA", "rgB")):
    reference = synth.make_reference([300_000, 20_000], seed=3)
    batch = synth.simulate_reads(reference, n, seed=seed, length=(30, 150), mix=(6, 1, 1, 2), paired=True,
                                 n_libs=len(read_groups), filtered_rate=0.05)
    header = SamHeader()
    header.add("@HD\tVN:1.6\tSO:unsorted")
    for name, length in zip(reference.names, reference.lengths):
        header.add("@SQ\tSN:%s\tLN:%d" % (name, length))
    for k, rg in enumerate(read_groups):
        header.add("@RG\tID:%s\tSM:s\tLB:lib%d" % (rg, k))
    with BamWriter(path, header, threads=2) as writer:
        writer.write_soa(batch, first_index=0, read_groups=list(read_groups))
    return reference, batch


@pytest.mark.parametrize("slab", [1 << 17, 300_001, 0])
@pytest.mark.parametrize("merge,apply_filter", [(False, True), (True, False)])
def test_stream_matches_the_host_decoder(tmp_path, slab, merge, apply_filter):
    """Slabs a little larger than two blocks: every slab ends inside a block and inside a record, both carried over."""
    bam = tmp_path / "reads.bam"
    synthetic_bam(bam, 60_000, seed=11)
    want, seen = host_batch(bam, merge, apply_filter)
    with DamageEngine(max_reads=0) as engine:
        with DeviceBamStream(engine, bam, merge_libraries=merge, apply_filter=apply_filter, slab_bytes=slab) as stream:
            parts = [engine.download(dev) for dev in stream]
            stats = stream.stats()
    assert stats["records_seen"] == seen and stats["blocks_on_device"] > 0
    if slab:
        assert len(parts) > 3
    assert_same_arrays(concatenate(parts), want)


@pytest.mark.parametrize("case", ["kat", "a05_libraries", "fuzz_0_l70_a10_q0", "fuzz_3_l200_a30_q13_merge", "rfuzz_0_12_12"])
def test_stream_on_python_made_files(case, tmp_path):
    """Files from the pure-Python encoder: tiny blocks (records span many), and the golden inputs' odd records."""
    header, records = read_sam(GOLDEN / case / "input.sam")
    bam = tmp_path / "in.bam"
    bam_py.write_bam(bam, header, records * 3, block_bytes=777)
    merge = "merge" in case or case in ("kat", "rfuzz_0_12_12")
    for apply_filter in (True, False):
        merge = merge or not apply_filter
        want, seen = host_batch(bam, merge, apply_filter)
        with DamageEngine(max_reads=0) as engine:
            with DeviceBamStream(engine, bam, merge_libraries=merge, apply_filter=apply_filter, slab_bytes=1 << 17) as stream:
                parts = [engine.download(dev) for dev in stream]
                assert stream.stats()["records_seen"] == seen
        if want is None:
            assert not parts
        else:
            assert_same_arrays(concatenate(parts), want)


def test_stream_errors(tmp_path, monkeypatch):
    monkeypatch.setenv("MDG_BAM_SLAB", "65536")  # host reader (header only): small slabs, so that a cut further on is not met at once
    header, records = read_sam(GOLDEN / "a06_no_readgroup" / "input.sam")
    bam = tmp_path / "in.bam"
    bam_py.write_bam(bam, header, records)
    with DamageEngine(max_reads=0) as engine:
        with DeviceBamStream(engine, bam, merge_libraries=False) as stream:
            with pytest.raises(BAMError) as info:
                list(stream)
            assert str(info.value) == "Read 'a6' has no read-group. Either fix BAM or use --merge-libraries"
        good = tmp_path / "good.bam"
        synthetic_bam(good, 20_000, seed=5)
        data = good.read_bytes()
        (tmp_path / "cut.bam").write_bytes(data[:len(data) // 2])
        with pytest.raises(BAMError):  # a file this small is seen to be cut short as soon as it is opened
            with DeviceBamStream(engine, tmp_path / "cut.bam", merge_libraries=True) as stream:
                list(stream)
        corrupt = bytearray(data)
        corrupt[len(data) // 2] ^= 0x55
        (tmp_path / "corrupt.bam").write_bytes(bytes(corrupt))
        with pytest.raises(BAMError):
            with DeviceBamStream(engine, tmp_path / "corrupt.bam", merge_libraries=True) as stream:
                list(stream)
        # a larger file cut inside a later slab: the header reads fine, the stream fails when it gets there
        big = tmp_path / "big.bam"
        synthetic_bam(big, 120_000, seed=6)
        data = big.read_bytes()
        (tmp_path / "big_cut.bam").write_bytes(data[:len(data) * 3 // 4])
        with DeviceBamStream(engine, tmp_path / "big_cut.bam", merge_libraries=True, slab_bytes=1 << 20) as stream:
            with pytest.raises(BAMError):
                list(stream)


@pytest.mark.parametrize("slab", [1 << 18, 0])
def test_rescale_and_encode_on_the_device(tmp_path, slab):
    """file -> device -> rescale -> device encoder -> file: the output holds the records of the host path byte for
    byte, and Python's gzip module accepts every block (CRC32, ISIZE)."""
    bam = tmp_path / "reads.bam"
    reference, _ = synthetic_bam(bam, 50_000, seed=21)
    corr = {("C", "T", p): 0.8 * 0.7 ** (p - 1) for p in range(1, 13)}
    corr.update({("G", "A", -p): 0.8 * 0.7 ** (p - 1) for p in range(1, 13)})
    model = RescaleModel(corr, 12, 12)
    out_dev, out_host = tmp_path / "dev.bam", tmp_path / "host.bam"
    with DamageEngine(max_reads=1 << 16) as engine:
        engine.set_reference(reference)
        engine.set_rescale_model(model)
        with DeviceBamStream(engine, bam, merge_libraries=True, apply_filter=False, want_mr=True, slab_bytes=slab) as stream, \
                BamWriter(out_dev, stream.header) as writer:
            n = 0
            for dev in stream:
                engine.rescale_resident(dev)
                assert not stream.has_mr(dev).any()
                stream.encode(dev, writer)
                n += dev.n
            raw_bytes, packed_bytes, _, _ = stream.flush()
        assert n == 50_000 and 0 < packed_bytes < raw_bytes
        with BamReader(bam, merge_libraries=True, apply_filter=False) as reader, BamWriter(out_host, reader.header) as writer:
            while True:
                batch = reader.read_batch(max_reads=1 << 16, keep_raw=True)
                if batch is None:
                    break
                qual, mr, status = engine.rescale(batch, compact=False)
                engine.sync()
                writer.write(batch, status=status, qual=qual, mr=mr)
    with gzip.open(out_dev, "rb") as a, gzip.open(out_host, "rb") as b:
        got, want = a.read(), b.read()
    assert got == want
    # and the device decoder reads its own encoder's file: MR tags flagged on the rescaled records
    with DamageEngine(max_reads=0) as engine:
        with DeviceBamStream(engine, out_dev, merge_libraries=True, apply_filter=False, want_mr=True) as stream:
            flagged = np.concatenate([stream.has_mr(dev) for dev in stream])
    assert flagged.shape[0] == 50_000 and flagged.any()


def test_sparse_rescale_returns_only_what_changed():
    reference = synth.make_reference([200_000], seed=3)
    batch = synth.simulate_reads(reference, 30_000, seed=4, length=(40, 120), mix=(6, 1, 1, 2), paired=True)
    corr = {("C", "T", p): 0.8 * 0.7 ** (p - 1) for p in range(1, 13)}
    corr.update({("G", "A", -p): 0.8 * 0.7 ** (p - 1) for p in range(1, 13)})
    for general in (False, True):
        with DamageEngine(max_reads=batch.n) as engine:
            engine.set_reference(reference)
            engine.set_rescale_model(RescaleModel(corr, 12, 12))
            if general:
                import os
                os.environ["MDG_RESCALE_GENERAL"] = "1"
            try:
                want_qual, want_mr, want_status = engine.rescale(batch)
                engine.sync()
                want_qual = want_qual.copy()
                patched = batch.slice(0, batch.n)
                patched.qual = batch.qual.copy()
                mr, status, ticket = engine.rescale_sparse(patched)
                changed = engine.rescale_collect(ticket, patched)
            finally:
                if general:
                    del os.environ["MDG_RESCALE_GENERAL"]
        assert np.array_equal(status, want_status) and np.array_equal(mr[status == 1], want_mr[status == 1])
        assert np.array_equal(patched.qual[:batch.total_bases], want_qual[:batch.total_bases])
        assert 0 < changed == int((want_qual[:batch.total_bases] != batch.qual[:batch.total_bases]).sum())


def test_encoder_limits_code_lengths(tmp_path):
    """Byte frequencies that fall off like Fibonacci numbers make a Huffman tree deeper than DEFLATE's 15 bits: the
    device encoder has to rebalance it (zlib's gen_bitlen), and every block must still inflate (gzip checks CRC32 and
    ISIZE of each)."""
    reference = synth.make_reference([300_000], seed=3)
    batch = synth.simulate_reads(reference, 40_000, seed=9, length=(150, 150), mix=(1, 0, 0, 0), paired=False)
    rng = np.random.default_rng(5)
    weights = np.array([1.618 ** -k for k in range(40)])
    batch.qual[:] = rng.choice(np.arange(2, 42, dtype=np.uint8), size=batch.qual.shape[0], p=weights / weights.sum())
    header = SamHeader()
    header.add("@HD\tVN:1.6\tSO:unsorted")
    header.add("@SQ\tSN:%s\tLN:%d" % (reference.names[0], reference.lengths[0]))
    src, out = tmp_path / "in.bam", tmp_path / "out.bam"
    with BamWriter(src, header, threads=2) as writer:
        writer.write_soa(batch, first_index=0)
    with DamageEngine(max_reads=0) as engine:
        with DeviceBamStream(engine, src, merge_libraries=True, apply_filter=False) as stream, BamWriter(out, stream.header) as writer:
            for dev in stream:
                stream.encode(dev, writer)
            stream.flush()
    with gzip.open(src, "rb") as a, gzip.open(out, "rb") as b:
        assert a.read() == b.read()
    with BamReader(out, merge_libraries=True, apply_filter=False) as reader:  # and the native host decoder agrees
        assert sum(part.n for part in reader) == batch.n
