"""Test-only, pure-Python BAM encoder / decoder written from the SAM/BAM specification (v1.6, sections 4.1-4.2),
independent of the product's C decoder: each checks the other.  BGZF is read with the standard ``gzip`` module
(a BGZF file is a series of gzip members) and written member by member with ``zlib``."""
import gzip
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"
SEQ_CODES = "=ACMGRSVTWYHKDBN"


def _bin(beg, end):  # SAM spec 5.3, reg2bin
    end -= 1
    for shift, offset in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return offset + (beg >> shift)
    return 0


def encode_record(rec):
    """``rec``: a samtext.SamRecord (tags given as the SAM text fields in ``tag_text``)."""
    name = rec.qname.encode() + b"\0"
    seq = rec.seq or ""
    l_seq = len(seq)
    ref_len = sum(n for op, n in rec.cigar if op in (0, 2, 3, 7, 8))
    end = rec.pos + (ref_len if ref_len else 1)
    core = struct.pack("<iiBBHHHiiii", rec.tid, rec.pos, len(name), rec.mapq, _bin(max(rec.pos, 0), end),
                       len(rec.cigar), rec.flag, l_seq, rec.mtid, rec.mpos, rec.tlen)
    cigar = b"".join(struct.pack("<I", n << 4 | op) for op, n in rec.cigar)
    nib = [SEQ_CODES.index(c.upper()) if c.upper() in SEQ_CODES else 15 for c in seq] + [0]
    packed = bytes(nib[i] << 4 | nib[i + 1] for i in range(0, l_seq, 2))
    qual = bytes(ord(c) - 33 for c in rec.qual) if rec.qual is not None else b"\xff" * l_seq
    aux = b""
    for item in rec.tag_text:
        tag, typ, value = item.split(":", 2)
        if typ == "Z":
            aux += tag.encode() + b"Z" + value.encode() + b"\0"
        elif typ == "i":
            aux += tag.encode() + b"i" + struct.pack("<i", int(value))
        elif typ == "f":
            aux += tag.encode() + b"f" + struct.pack("<f", float(value))
        elif typ == "A":
            aux += tag.encode() + b"A" + value.encode()[:1]
        else:
            raise ValueError("tag type %r not needed by the tests" % typ)
    body = core + name + cigar + packed + qual + aux
    return struct.pack("<i", len(body)) + body


def bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    payload = comp.compress(data) + comp.flush()
    bsize = 18 + len(payload) + 8 - 1
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + payload
            + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def write_bam(path, header, records, block_bytes=0xff00):
    """``header``: samtext.SamHeader.  ``block_bytes`` small values force records across block boundaries."""
    text = "".join(line + "\n" for line in header.lines).encode()
    data = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(header.references))
    for name, length in zip(header.references, header.lengths):
        raw = name.encode() + b"\0"
        data += struct.pack("<i", len(raw)) + raw + struct.pack("<i", length)
    data += b"".join(encode_record(r) for r in records)
    with open(path, "wb") as handle:
        for at in range(0, len(data), block_bytes):
            handle.write(bgzf_block(data[at:at + block_bytes]))
        handle.write(BGZF_EOF)


def read_bam(path):
    """``(header_text, [(name, length)], [record dict])`` decoded per the specification."""
    with gzip.open(path, "rb") as handle:
        data = handle.read()
    assert data[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", data, 4)
    text = data[8:8 + l_text].rstrip(b"\0").decode()
    at = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, at)
    at += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, at)
        name = data[at + 4:at + 4 + l_name - 1].decode()
        length, = struct.unpack_from("<i", data, at + 4 + l_name)
        refs.append((name, length))
        at += 8 + l_name
    records = []
    while at < len(data):
        size, = struct.unpack_from("<i", data, at)
        body = data[at + 4:at + 4 + size]
        at += 4 + size
        tid, pos, l_name, mapq, _bin_, n_cig, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", body, 0)
        p = 32
        qname = body[p:p + l_name - 1].decode()
        p += l_name
        cigar = [(w & 0xF, w >> 4) for w in struct.unpack_from("<%dI" % n_cig, body, p)]
        p += 4 * n_cig
        packed = body[p:p + (l_seq + 1) // 2]
        p += (l_seq + 1) // 2
        seq = "".join(SEQ_CODES[b >> 4] + SEQ_CODES[b & 15] for b in packed)[:l_seq]
        q = body[p:p + l_seq]
        p += l_seq
        qual = None if (l_seq and q[0] == 0xFF) or not l_seq else "".join(chr(x + 33) for x in q)
        tags = {}
        while p < len(body):
            tag, typ = body[p:p + 2].decode(), chr(body[p + 2])
            p += 3
            if typ == "Z":
                end = body.index(b"\0", p)
                tags[tag] = ("Z", body[p:end].decode())
                p = end + 1
            elif typ in "iI":
                tags[tag] = ("i", struct.unpack_from("<i", body, p)[0])
                p += 4
            elif typ == "f":
                tags[tag] = ("f", struct.unpack_from("<f", body, p)[0])
                p += 4
            elif typ in "AcC":
                tags[tag] = (typ, body[p])
                p += 1
            elif typ in "sS":
                tags[tag] = (typ, struct.unpack_from("<h", body, p)[0])
                p += 2
            else:
                raise ValueError("tag type %r" % typ)
        records.append(dict(qname=qname, flag=flag, tid=tid, pos=pos, mapq=mapq, cigar=cigar, mtid=mtid, mpos=mpos,
                            tlen=tlen, seq=seq if l_seq else None, qual=qual, tags=tags))
    return text, refs, records
