"""Known-answer test of the BAM record layer against bytes assembled BY HAND from the SAM/BAM specification (v1.6).

The codec tests elsewhere compare the native decoder with ``tests/bam_py.py`` -- same author, same reading of the
specification.  Here the example alignments of the specification's section 1.1 are written out field by field as
literal hexadecimal (section 4.2 gives the layout; every value below can be checked against the text of the
specification with a pocket calculator), compressed by Python's zlib into BGZF blocks with the header bytes of
section 4.1, and closed by the end-of-file marker block the specification prints verbatim.  No htslib, samtools
or pysam exists in this image to make such a file; this is the external pin the record layer gets instead.

    r001   99 ref  7 30 8M2I4M1D3M = 37  39 TTAGATAAAGGATACTG *
    r002    0 ref  9 30 3S6M1P1I4M *  0   0 AAAAGATAAGGATA    *
    r003    0 ref  9 30 5S6M       *  0   0 GCCTAAGCTAA       * SA:Z:ref,29,-,6H5M,17,0;
    r004    0 ref 16 30 6M14N5M    *  0   0 ATAGCTTCAGC       *
    r003 2064 ref 29 17 6H5M       *  0   0 TAGGC             * SA:Z:ref,9,+,5S6M,30,1;
    r001  147 ref 37 30 9M         =  7 -39 CAGCGGCAT         * NM:i:1
"""
import struct
import zlib

import numpy as np
import pytest

from mapdamage_b200.bamio import BamReader

HEADER_TEXT = b"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ref\tLN:45\n"


def h(text):
    return bytes.fromhex(text.replace(" ", ""))


# refID pos l_read_name mapq bin n_cigar_op flag l_seq next_refID next_pos tlen | read_name | cigar | seq | qual [| tags]
# bin: every alignment lies inside the first 16 kb of the reference -> 4681 = 0x1249 (section 5.3)
RECORDS = [
    # r001: pos 7 -> 6; 8M2I4M1D3M = 8<<4|0, 2<<4|1, 4<<4|0, 1<<4|2, 3<<4|0; mate at 37 -> 36, tlen 39
    h("00000000 06000000 05 1e 4912 0500 6300 11000000 00000000 24000000 27000000") + b"r001\0"
    + h("80000000 21000000 40000000 12000000 30000000") + h("88 14 18 11 14 41 81 28 40") + b"\xff" * 17,
    # r002: 3S6M1P1I4M = 3<<4|4, 6<<4|0, 1<<4|6, 1<<4|1, 4<<4|0; no mate: -1, -1, 0
    h("00000000 08000000 05 1e 4912 0500 0000 0e000000 ffffffff ffffffff 00000000") + b"r002\0"
    + h("34000000 60000000 16000000 11000000 40000000") + h("11 11 41 81 14 41 81") + b"\xff" * 14,
    # r003: 5S6M; eleven bases, the last nibble is padding
    h("00000000 08000000 05 1e 4912 0200 0000 0b000000 ffffffff ffffffff 00000000") + b"r003\0"
    + h("54000000 60000000") + h("42 28 11 42 81 10") + b"\xff" * 11 + b"SAZref,29,-,6H5M,17,0;\0",
    # r004: 6M14N5M = 6<<4|0, 14<<4|3, 5<<4|0
    h("00000000 0f000000 05 1e 4912 0300 0000 0b000000 ffffffff ffffffff 00000000") + b"r004\0"
    + h("60000000 e3000000 50000000") + h("18 14 28 82 14 20") + b"\xff" * 11,
    # r003, supplementary (2064 = 0x810): 6H5M = 6<<4|5, 5<<4|0; mapq 17
    h("00000000 1c000000 05 11 4912 0200 1008 05000000 ffffffff ffffffff 00000000") + b"r003\0"
    + h("65000000 50000000") + h("81 44 20") + b"\xff" * 5 + b"SAZref,9,+,5S6M,30,1;\0",
    # r001, second mate (147 = 0x93): 9M; mate at 7 -> 6, tlen -39; NM:i:1 stored as an unsigned byte
    h("00000000 24000000 05 1e 4912 0100 9300 09000000 00000000 06000000 d9ffffff") + b"r001\0"
    + h("90000000") + h("21 42 44 21 80") + b"\xff" * 9 + b"NMC\x01",
]
EOF_BLOCK = h("1f 8b 08 04 00 00 00 00 00 ff 06 00 42 43 02 00 1b 00 03 00 00 00 00 00 00 00 00 00")  # section 4.1.2


def bgzf_block(data):
    deflater = zlib.compressobj(6, zlib.DEFLATED, -15)
    payload = deflater.compress(data) + deflater.flush()
    total = 12 + 6 + len(payload) + 8
    return (h("1f 8b 08 04 00 00 00 00 00 ff 06 00 42 43 02 00") + struct.pack("<H", total - 1) + payload
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_example(path, records_per_block=2):
    head = b"BAM\1" + struct.pack("<i", len(HEADER_TEXT)) + HEADER_TEXT + struct.pack("<i", 1) \
        + struct.pack("<i", 4) + b"ref\0" + struct.pack("<i", 45)
    body = [struct.pack("<i", len(r)) + r for r in RECORDS]
    blocks = [bgzf_block(head)]
    for i in range(0, len(body), records_per_block):
        blocks.append(bgzf_block(b"".join(body[i:i + records_per_block])))
    path.write_bytes(b"".join(blocks) + EOF_BLOCK)


EXPECT = dict(
    flag=[99, 0, 0, 0, 2064, 147], pos=[6, 8, 8, 15, 28, 36], tid=[0] * 6, l_seq=[17, 14, 11, 11, 5, 9],
    tlen=[39, 0, 0, 0, 0, -39], mtid=[0, -1, -1, -1, -1, 0], mpos=[36, -1, -1, -1, -1, 6],
    cigar=[[128, 33, 64, 18, 48], [52, 96, 22, 17, 64], [84, 96], [96, 227, 80], [101, 80], [144]],
    seq=["TTAGATAAAGGATACTG", "AAAAGATAAGGATA", "GCCTAAGCTAA", "ATAGCTTCAGC", "TAGGC", "CAGCGGCAT"],
)


def check(batch, keep):
    assert batch.n == len(keep)
    for name in ("flag", "pos", "tid", "l_seq", "tlen", "mtid", "mpos"):
        assert [int(x) for x in getattr(batch, name)] == [EXPECT[name][k] for k in keep], name
    for i, k in enumerate(keep):
        assert [int(w) for w in batch.cigar[batch.cigar_off[i]:batch.cigar_off[i + 1]]] == EXPECT["cigar"][k]
        off = int(batch.base_off[i])
        nibbles = [(int(b) >> s) & 15 for b in batch.seq4[off // 2:off // 2 + (len(EXPECT["seq"][k]) + 1) // 2] for s in (4, 0)]
        assert "".join("=ACMGRSVTWYHKDBN"[n] for n in nibbles[:len(EXPECT["seq"][k])]) == EXPECT["seq"][k]
        assert np.all(batch.qual[off:off + len(EXPECT["seq"][k])] == 0xFF)  # '*': no qualities


@pytest.mark.parametrize("records_per_block", [1, 2, 6])
def test_host_decoder_reads_the_specification_example(tmp_path, records_per_block):
    write_example(tmp_path / "example.bam", records_per_block)
    with BamReader(tmp_path / "example.bam", merge_libraries=True, apply_filter=False) as reader:
        assert reader.header.references == ["ref"] and reader.header.lengths == [45]
        assert [line for line in reader.header.lines] == HEADER_TEXT.decode().splitlines()
        check(reader.read_batch(), [0, 1, 2, 3, 4, 5])
    with BamReader(tmp_path / "example.bam", merge_libraries=True, apply_filter=True) as reader:
        check(reader.read_batch(), [0, 1, 2, 3, 5])  # reader.py:121-132 drops the supplementary alignment (0x800)


@pytest.mark.gpu
def test_device_decoder_reads_the_specification_example(tmp_path):
    from mapdamage_b200.bamio import DeviceBamStream
    from mapdamage_b200.engine import DamageEngine

    write_example(tmp_path / "example.bam", 1)
    with DamageEngine(max_reads=0) as engine:
        for apply_filter, keep in ((False, [0, 1, 2, 3, 4, 5]), (True, [0, 1, 2, 3, 5])):
            with DeviceBamStream(engine, tmp_path / "example.bam", merge_libraries=True, apply_filter=apply_filter) as stream:
                batches = [engine.download(dev) for dev in stream]
            assert len(batches) == 1
            check(batches[0], keep)
