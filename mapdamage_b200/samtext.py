"""Plain-text SAM reader (host glue; per-record Python, not a hot path).

pysam/htslib are not available in this image, so the
package reads SAM text itself.  Records expose the fields the batch builder
needs, with htslib's conventions: 0-based ``pos``, ``mtid`` resolved from
``RNEXT`` (``=`` means the read's own reference), ``tags`` as a dict.
"""
from pathlib import Path

CIGAR_OPS = "MIDNSHP=X"


class SamRecord:
    __slots__ = ("qname", "flag", "rname", "tid", "pos", "mapq", "cigar", "mtid", "mpos",
                 "tlen", "seq", "qual", "tags", "tag_text")

    def aend(self):
        return self.pos + sum(n for op, n in self.cigar if op in (0, 2, 3, 7, 8))


def parse_cigar(text):
    if text == "*":
        return []
    out, num = [], 0
    for ch in text:
        if ch.isdigit():
            num = num * 10 + ord(ch) - 48
        else:
            out.append((CIGAR_OPS.index(ch), num))
            num = 0
    return out


def format_cigar(cigar):
    return "".join("%d%s" % (n, CIGAR_OPS[op]) for op, n in cigar) or "*"


class SamHeader:
    def __init__(self):
        self.lines = []
        self.references = []
        self.lengths = []
        self.readgroups = {}
        self._tid = {}

    def add(self, line):
        self.lines.append(line)
        fields = line.split("\t")
        record = dict(f.split(":", 1) for f in fields[1:] if ":" in f)
        if fields[0] == "@SQ":
            self._tid[record["SN"]] = len(self.references)
            self.references.append(record["SN"])
            self.lengths.append(int(record["LN"]))
        elif fields[0] == "@RG":
            self.readgroups[record.get("ID")] = record

    def tid(self, name):
        return -1 if name == "*" else self._tid[name]

    def set_references(self, names, lengths):
        """Replaces the reference list (BAM: the binary list wins over the @SQ lines)."""
        self.references = list(names)
        self.lengths = list(lengths)
        self._tid = {name: i for i, name in enumerate(self.references)}

    def libraries(self):
        """Read-group ID -> (sample, library); KeyError text as reader.py:107-116."""
        from .batch import BAMError

        out = {}
        for rg_id, record in self.readgroups.items():
            try:
                out[rg_id] = (record["SM"], record["LB"])
            except KeyError as error:  # reader.py:107-116
                raise BAMError("Incomplete readgroup found: %s is missing %s. "
                               "Either fix BAM or use --merge-libraries"
                               % (rg_id if rg_id is not None else "Unnamed readgroup", error))
        return out


def parse_record(line, header):
    f = line.rstrip("\n").split("\t")
    r = SamRecord()
    r.qname = f[0]
    r.flag = int(f[1])
    r.rname = f[2]
    r.tid = header.tid(f[2])
    r.pos = int(f[3]) - 1
    r.mapq = int(f[4])
    r.cigar = parse_cigar(f[5])
    r.mtid = r.tid if f[6] == "=" else header.tid(f[6])
    r.mpos = int(f[7]) - 1
    r.tlen = int(f[8])
    r.seq = None if f[9] == "*" else f[9]
    r.qual = None if f[10] == "*" else f[10]
    r.tag_text = f[11:]
    r.tags = {}
    for item in f[11:]:
        tag, _, value = item.split(":", 2)
        r.tags[tag] = value
    return r


def read_sam(path):
    """Returns ``(header, [records])`` for a SAM text file."""
    header = SamHeader()
    records = []
    with open(Path(path), "rt") as handle:
        for line in handle:
            if line.startswith("@"):
                header.add(line.rstrip("\n"))
            elif line.strip():
                records.append(parse_record(line, header))
    return header, records


def read_fasta(path):
    """FASTA -> ordered dict name -> sequence string (as stored, not uppercased)."""
    seqs, name, chunks = {}, None, []
    with open(Path(path), "rt") as handle:
        for line in handle:
            line = line.strip()
            if line.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(chunks)
                name, chunks = line[1:].split()[0], []
            elif line:
                chunks.append(line)
    if name is not None:
        seqs[name] = "".join(chunks)
    return seqs


def iter_sam(path):
    """``(header, record iterator)``: the header is complete on return, records stream."""
    header = SamHeader()
    handle = open(Path(path), "rt")
    first = None
    for line in handle:
        if line.startswith("@"):
            header.add(line.rstrip("\n"))
        elif line.strip():
            first = line
            break

    def records():
        try:
            if first is not None:
                yield parse_record(first, header)
                for line in handle:
                    if line.strip():
                        yield parse_record(line, header)
        finally:
            handle.close()

    return header, records()


def format_record(record, header, qual=None, extra_tags=()):
    """SAM text line of ``record``, optionally with new qualities / appended tags."""
    if record.mtid is None or record.mtid < 0:
        rnext = "*"
    else:
        rnext = "=" if record.mtid == record.tid else header.references[record.mtid]
    fields = [
        record.qname, str(record.flag), record.rname, str(record.pos + 1), str(record.mapq),
        format_cigar(record.cigar), rnext, str(record.mpos + 1), str(record.tlen),
        record.seq if record.seq is not None else "*",
        (qual if qual is not None else record.qual) or "*",
    ]
    fields.extend(record.tag_text)
    fields.extend(extra_tags)
    return "\t".join(fields)
