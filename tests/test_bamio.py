"""Native BGZF/BAM decode and encode (csrc/mdg_bamio.cpp) against an independent pure-Python codec written from
the SAM/BAM specification (tests/bam_py.py) and against the SAM path.  No GPU needed: the decoder is host code."""
import numpy as np
import pytest

import bam_py
from conftest import GOLDEN
from mapdamage_b200 import synth
from mapdamage_b200.bamio import BamReader, BamWriter
from mapdamage_b200.batch import BAMError, BatchBuilder
from mapdamage_b200.samtext import read_sam

FIELDS = ("flag", "tid", "pos", "lib", "l_seq", "base_off", "cigar_off", "cigar", "seq4", "tlen", "mtid", "mpos")


def sam_batch(header, records, merge, apply_filter):
    builder = BatchBuilder(readgroups=None if merge else header.libraries(), merge_libraries=merge,
                           apply_filter=apply_filter)
    for record in records:
        builder.add(record)
    return builder.finish(with_qual=True), builder.libraries


def assert_same_batch(got, want):
    assert got.n == want.n
    for name in FIELDS:
        assert np.array_equal(getattr(got, name), getattr(want, name)), name
    # qualities: compare real bases (the pad slot of odd-length reads is unspecified in the SAM path)
    for i in range(got.n):
        assert got.qualities_of(i) == want.qualities_of(i), i


@pytest.mark.parametrize("case", ["kat", "a05_libraries", "fuzz_0_l70_a10_q0", "fuzz_3_l200_a30_q13_merge", "rfuzz_0_12_12"])
@pytest.mark.parametrize("block_bytes,threads", [(0xff00, 1), (777, 3)])
def test_reader_matches_the_sam_path(case, block_bytes, threads, tmp_path):
    header, records = read_sam(GOLDEN / case / "input.sam")
    bam = tmp_path / "in.bam"
    bam_py.write_bam(bam, header, records, block_bytes=block_bytes)
    merge = "merge" in case or case in ("kat", "rfuzz_0_12_12")
    for apply_filter in (True, False):
        merge = merge or not apply_filter  # the rescale pass (no filter) knows no libraries
        want, libraries = sam_batch(header, records, merge, apply_filter)
        with BamReader(bam, threads=threads, merge_libraries=merge, apply_filter=apply_filter) as reader:
            assert reader.header.references == header.references and reader.header.lengths == header.lengths
            assert reader.libraries == libraries
            parts = []
            while True:
                batch = reader.read_batch(max_reads=97)  # several batches, records straddling refills
                if batch is None:
                    break
                parts.append(batch)
            assert reader.records_seen == len(records)
        assert sum(p.n for p in parts) == want.n
        at = 0
        for part in parts:
            assert_same_batch(part, want.slice(at, at + part.n))
            at += part.n


def test_reader_errors(tmp_path):
    header, records = read_sam(GOLDEN / "a06_no_readgroup" / "input.sam")
    bam = tmp_path / "in.bam"
    bam_py.write_bam(bam, header, records)
    with BamReader(bam, merge_libraries=False) as reader:
        with pytest.raises(BAMError) as info:
            reader.read_batch()
        # the reference's text, read name included (reader.py:67-73)
        assert str(info.value) == "Read 'a6' has no read-group. Either fix BAM or use --merge-libraries"
    with BamReader(bam, merge_libraries=False, lenient_libraries=True) as reader:
        batch = reader.read_batch()
        failures = reader.library_failures()
        assert failures and all(batch.lib[i] == 0xFFFF for i, _ in failures)
        assert failures[0][1].startswith("Read 'a6' has no read-group")
    (tmp_path / "junk.bam").write_bytes(b"this is not a BAM file at all, not even gzip")
    with pytest.raises(BAMError):
        BamReader(tmp_path / "junk.bam")
    data = bam.read_bytes()
    (tmp_path / "cut.bam").write_bytes(data[:len(data) // 2])
    with pytest.raises(BAMError):
        with BamReader(tmp_path / "cut.bam", merge_libraries=True) as reader:
            while reader.read_batch() is not None:
                pass
    corrupt = bytearray(data)
    corrupt[40] ^= 0x55
    (tmp_path / "corrupt.bam").write_bytes(bytes(corrupt))
    with pytest.raises(BAMError):
        with BamReader(tmp_path / "corrupt.bam", merge_libraries=True) as reader:
            while reader.read_batch() is not None:
                pass


def test_large_file_round_trip(tmp_path):
    """40k synthetic pairs through the Python encoder, the C decoder, the C encoder and the Python decoder."""
    reference = synth.make_reference([300_000, 20_000], seed=3)
    batch = synth.simulate_reads(reference, 40_000, seed=4, length=(30, 150), mix=(6, 1, 1, 2), paired=True)
    sam = tmp_path / "in.sam"
    synth.write_sam(batch, reference, sam)
    header, records = read_sam(sam)
    bam = tmp_path / "in.bam"
    bam_py.write_bam(bam, header, records)
    want, _ = sam_batch(header, records, True, False)
    out = tmp_path / "out.bam"
    with BamReader(bam, threads=4, merge_libraries=True, apply_filter=False) as reader, \
            BamWriter(out, reader.header, threads=4) as writer:
        at = 0
        while True:
            part = reader.read_batch(max_reads=16_384, keep_raw=True)
            if part is None:
                break
            assert_same_batch(part, want.slice(at, at + part.n))
            assert not part.has_mr.any()
            # every third record "rescaled": new qualities and an MR tag
            status = (np.arange(at, at + part.n) % 3 == 0).astype(np.uint8)
            new_qual = (part.qual + 1).astype(np.uint8)
            writer.write(part, status=status, qual=new_qual, mr=np.arange(at, at + part.n, dtype=np.float32) / 8)
            at += part.n
        assert at == want.n
    text, refs, decoded = bam_py.read_bam(out)
    assert refs == list(zip(header.references, header.lengths))
    assert text.splitlines() == header.lines
    assert len(decoded) == len(records)
    for i, (got, rec) in enumerate(zip(decoded, records)):
        assert (got["qname"], got["flag"], got["tid"], got["pos"], got["cigar"], got["seq"], got["tlen"]) == \
               (rec.qname, rec.flag, rec.tid, rec.pos, rec.cigar, rec.seq, rec.tlen)
        if i % 3 == 0:
            assert got["qual"] == "".join(chr(ord(c) + 1) for c in rec.qual)
            assert got["tags"]["MR"] == ("f", i / 8)
        else:
            assert got["qual"] == rec.qual and "MR" not in got["tags"]
    # and the written file reads back through the C decoder, MR tags flagged
    with BamReader(out, threads=2, merge_libraries=True, apply_filter=False) as reader:
        again = reader.read_batch(max_reads=50_000, keep_raw=True)
        assert again.n == want.n and np.array_equal(again.has_mr, np.arange(want.n) % 3 == 0)


def test_soa_encoder(tmp_path):
    """A batch written with write_soa decodes (Python codec and C decoder) to the same batch."""
    from mapdamage_b200.samtext import SamHeader

    reference = synth.make_reference([100_000, 5_000], seed=3)
    batch = synth.simulate_reads(reference, 5_001, seed=8, length=(31, 99), mix=(6, 1, 1, 2), paired=False, n_libs=2)
    header = SamHeader()
    header.add("@HD\tVN:1.6\tSO:unsorted")
    for name, length in zip(reference.names, reference.lengths):
        header.add("@SQ\tSN:%s\tLN:%d" % (name, length))
    header.add("@RG\tID:rgA\tSM:s\tLB:libA")
    header.add("@RG\tID:rgB\tSM:s\tLB:libB")
    with BamWriter(tmp_path / "soa.bam", header, threads=2) as writer:
        writer.write_soa(batch.slice(0, 3000), first_index=0, read_groups=["rgA", "rgB"])
        writer.write_soa(batch.slice(3000, batch.n), first_index=3000, read_groups=["rgA", "rgB"])
    _, _, decoded = bam_py.read_bam(tmp_path / "soa.bam")
    assert [d["qname"] for d in decoded[:2] + decoded[-1:]] == ["r0", "r1", "r%d" % (batch.n - 1)]
    for i in (0, 17, 3000, batch.n - 1):
        assert decoded[i]["seq"] == batch.sequence_of(i) and decoded[i]["qual"] == batch.qualities_of(i)
        assert decoded[i]["cigar"] == batch.cigar_of(i) and decoded[i]["tags"]["RG"][1] == ("rgA", "rgB")[batch.lib[i]]
    with BamReader(tmp_path / "soa.bam", merge_libraries=False, apply_filter=False) as reader:
        assert reader.libraries == [("s", "libA"), ("s", "libB")]
        got = reader.read_batch(max_reads=6000)
    assert_same_batch(got, batch)


@pytest.mark.parametrize("slab", ["65536", "70001", "250000"])
def test_many_slabs(tmp_path, monkeypatch, slab):
    """Slabs a little larger than a block (MDG_BAM_SLAB): every slab ends inside a block, which is carried over."""
    import numpy as np

    from conftest import GOLDEN

    header, records = read_sam(GOLDEN / "fuzz_0_l70_a10_q0" / "input.sam")
    bam_py.write_bam(tmp_path / "in.bam", header, records * 4, block_bytes=60_000)
    with BamReader(tmp_path / "in.bam", merge_libraries=True) as reader:
        want = reader.read_batch()
    monkeypatch.setenv("MDG_BAM_SLAB", slab)
    with BamReader(tmp_path / "in.bam", merge_libraries=True, threads=3) as reader:
        parts = []
        while True:
            batch = reader.read_batch(max_reads=500)
            if batch is None:
                break
            parts.append(batch)
    assert sum(p.n for p in parts) == want.n
    for name in ("flag", "pos", "l_seq"):
        assert np.array_equal(np.concatenate([getattr(p, name) for p in parts]), getattr(want, name)), name
    assert np.array_equal(np.concatenate([p.cigar for p in parts]), want.cigar)
