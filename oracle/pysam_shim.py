"""TEST INFRASTRUCTURE ONLY -- a minimal in-memory stand-in for ``pysam``.

The unmodified reference (``/root/reference/mapdamage``) imports ``pysam`` and
``coloredlogs``; neither is installed in this image and there is no network.
``install()`` registers stub modules in ``sys.modules`` so that the reference's
``mapdamage.main.main`` and ``mapdamage.rescale.rescale_qual`` run unchanged on
SAM *text* + FASTA fixtures.  It is used only by ``oracle/gen_golden.py`` and
by tests that run the live reference when ``/root/reference`` is present (it
is absent on the GPU box).  Nothing under ``mapdamage_b200/`` may import this.

htslib / pysam semantics reproduced here (the reference relies on them but
does not vendor or pin pysam -- ``setup.py:55``):

* ``pos`` is 0-based; ``aend = pos + sum(len of M, D, N, =, X)``.
* ``query`` / ``qqual`` strip leading and trailing soft clips (skipping ``H``).
* ``qual`` is the Phred+33 string or ``None``; assigning a string of the wrong
  length raises ``ValueError`` as real pysam does.
* ``get_tag`` raises ``KeyError`` when the tag is absent.

Every attribute the reference touches is listed in SURVEY.md section 8(c).
"""
import sys
import types
from pathlib import Path

CIGAR_OPS = "MIDNSHP=X"
_REF_CONSUMING = (0, 2, 3, 7, 8)


def parse_cigar_string(text):
    if text == "*" or not text:
        return None
    out, num = [], ""
    for ch in text:
        if ch.isdigit():
            num += ch
        else:
            out.append((CIGAR_OPS.index(ch), int(num)))
            num = ""
    return out


def cigar_to_string(cigar):
    if not cigar:
        return "*"
    return "".join("%d%s" % (n, CIGAR_OPS[op]) for op, n in cigar)


class AlignedSegment:
    """One SAM record with the legacy pysam attribute names the reference uses."""

    def __init__(self, fields, header):
        self._header = header
        self.qname = self.query_name = fields[0]
        self.flag = int(fields[1])
        self._rname = fields[2]
        self.tid = self.reference_id = header.tid_of(fields[2])
        self.pos = self.reference_start = int(fields[3]) - 1
        self.mapq = int(fields[4])
        self.cigar = parse_cigar_string(fields[5])
        if fields[6] == "=":
            self.mrnm = self.tid
        else:
            self.mrnm = header.tid_of(fields[6])
        self.pnext = int(fields[7]) - 1
        self.template_length = self.tlen = int(fields[8])
        self.seq = None if fields[9] == "*" else fields[9]
        self._qual = None if fields[10] == "*" else fields[10]
        self._tags = []
        for item in fields[11:]:
            tag, typ, value = item.split(":", 2)
            if typ == "i":
                value = int(value)
            elif typ == "f":
                value = float(value)
            self._tags.append([tag, typ, value])

    # -- flags ---------------------------------------------------------
    is_paired = property(lambda self: bool(self.flag & 0x1))
    is_proper_pair = property(lambda self: bool(self.flag & 0x2))
    is_unmapped = property(lambda self: bool(self.flag & 0x4))
    is_reverse = property(lambda self: bool(self.flag & 0x10))
    mate_is_reverse = property(lambda self: bool(self.flag & 0x20))
    is_read1 = property(lambda self: bool(self.flag & 0x40))

    # -- coordinates ---------------------------------------------------
    @property
    def aend(self):
        if self.is_unmapped or not self.cigar:
            return None
        return self.pos + sum(n for op, n in self.cigar if op in _REF_CONSUMING)

    reference_end = aend

    @property
    def reference_length(self):
        end = self.aend
        return None if end is None else end - self.pos

    def _clip_bounds(self):
        length = len(self.seq)
        start, end = 0, length
        cigar = self.cigar or []
        for op, n in cigar:
            if op == 4:
                start += n
            elif op == 5:
                continue
            else:
                break
        for op, n in reversed(cigar):
            if op == 4:
                end -= n
            elif op == 5:
                continue
            else:
                break
        return start, max(start, end)

    @property
    def query(self):
        if self.seq is None:
            return None
        start, end = self._clip_bounds()
        return self.seq[start:end]

    @property
    def qqual(self):
        if self._qual is None or self.seq is None:
            return None
        start, end = self._clip_bounds()
        return self._qual[start:end]

    @property
    def qual(self):
        return self._qual

    @qual.setter
    def qual(self, value):
        if value is not None and self.seq is not None and len(value) != len(self.seq):
            raise ValueError(
                "quality and sequence mismatch: %i != %i" % (len(value), len(self.seq))
            )
        self._qual = value

    # -- tags ----------------------------------------------------------
    def get_tag(self, tag):
        for key, _, value in self._tags:
            if key == tag:
                return value
        raise KeyError("tag '%s' not present" % tag)

    def has_tag(self, tag):
        return any(key == tag for key, _, _ in self._tags)

    def set_tag(self, tag, value, value_type=None):
        typ = value_type or ("i" if isinstance(value, int) else "Z")
        for item in self._tags:
            if item[0] == tag:
                item[1], item[2] = typ, value
                return
        self._tags.append([tag, typ, value])

    def __str__(self):
        return self.to_sam()

    def to_sam(self):
        rnext = "*"
        if self.mrnm is not None and self.mrnm >= 0:
            rnext = "=" if self.mrnm == self.tid else self._header.references[self.mrnm]
        fields = [
            self.qname,
            str(self.flag),
            self._rname,
            str(self.pos + 1),
            str(self.mapq),
            cigar_to_string(self.cigar),
            rnext,
            str(self.pnext + 1),
            str(self.template_length),
            self.seq if self.seq is not None else "*",
            self._qual if self._qual is not None else "*",
        ]
        for tag, typ, value in self._tags:
            if typ == "f":
                # float32 round trip, as a BAM 'f' tag would store it
                import struct

                value = struct.unpack("<f", struct.pack("<f", value))[0]
                fields.append("%s:f:%r" % (tag, value))
            else:
                fields.append("%s:%s:%s" % (tag, typ, value))
        return "\t".join(fields)


class _Header:
    def __init__(self):
        self.lines = []
        self.references = []
        self.lengths = []
        self.readgroups = []

    def add_line(self, line):
        self.lines.append(line)
        fields = line.split("\t")
        record = dict(f.split(":", 1) for f in fields[1:] if ":" in f)
        if fields[0] == "@SQ":
            self.references.append(record["SN"])
            self.lengths.append(int(record["LN"]))
        elif fields[0] == "@RG":
            self.readgroups.append(record)

    def tid_of(self, name):
        if name == "*":
            return -1
        return self.references.index(name)

    def get(self, key, default=None):
        if key == "RG":
            return self.readgroups if self.readgroups else default
        return default


class AlignmentFile:
    """Reads SAM text; in write mode collects records and writes SAM text."""

    def __init__(self, filepath, mode="r", template=None):
        self._path = Path(filepath)
        self._mode = mode
        self._records = []
        if "w" in mode:
            self.header = template.header
        else:
            self.header = _Header()
            with open(self._path, "rt") as handle:
                for line in handle:
                    line = line.rstrip("\n")
                    if not line:
                        continue
                    if line.startswith("@"):
                        self.header.add_line(line)
                    else:
                        self._records.append(AlignedSegment(line.split("\t"), self.header))

    references = property(lambda self: tuple(self.header.references))
    lengths = property(lambda self: tuple(self.header.lengths))

    def getrname(self, tid):
        return self.header.references[tid]

    get_reference_name = getrname

    def __iter__(self):
        return iter(self._records)

    def write(self, read):
        self._records.append(read)

    def close(self):
        if "w" in self._mode:
            with open(self._path, "wt") as handle:
                for line in self.header.lines:
                    handle.write(line + "\n")
                for read in self._records:
                    handle.write(read.to_sam() + "\n")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class FastaFile:
    def __init__(self, filepath):
        self._seqs = {}
        name = None
        chunks = []
        with open(filepath, "rt") as handle:
            for line in handle:
                line = line.strip()
                if line.startswith(">"):
                    if name is not None:
                        self._seqs[name] = "".join(chunks)
                    name = line[1:].split()[0]
                    chunks = []
                elif line:
                    chunks.append(line)
        if name is not None:
            self._seqs[name] = "".join(chunks)

    references = property(lambda self: tuple(self._seqs))

    def fetch(self, reference, start=None, end=None):
        seq = self._seqs[reference]
        start = 0 if start is None else max(0, start)
        end = len(seq) if end is None else min(len(seq), end)
        if end < start:
            raise ValueError("invalid coordinates: start (%i) > stop (%i)" % (start, end))
        return seq[start:end]

    def close(self):
        pass


def install():
    """Registers the ``pysam`` and ``coloredlogs`` stand-ins (idempotent)."""
    pysam = types.ModuleType("pysam")
    pysam.AlignmentFile = AlignmentFile
    pysam.Samfile = AlignmentFile
    pysam.FastaFile = FastaFile
    pysam.Fastafile = FastaFile
    pysam.AlignedSegment = AlignedSegment
    pysam.set_verbosity = lambda level: 0
    pysam.__shim__ = True
    sys.modules["pysam"] = pysam

    coloredlogs = types.ModuleType("coloredlogs")
    coloredlogs.install = lambda **kwargs: None
    sys.modules["coloredlogs"] = coloredlogs
    return pysam
