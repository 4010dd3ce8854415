"""TEST INFRASTRUCTURE ONLY -- adversarial SAM/FASTA generator for the golden vectors.

Per-record Python on purpose: it exercises every CIGAR shape and flag the
reference accepts (both strands, 1-6 M/=/X blocks with I/D/N/P between them,
soft clips and outer hard clips, N/IUPAC bases in reads, N/lower-case/IUPAC in
the reference, reads touching both contig ends, all filter flags, several read
groups, reads without qualities, every pairing orientation).  It never emits
what makes the reference crash: mapped reads without SEQ/CIGAR or
reads hanging over a contig end.
"""
import random

OPS = "MIDNSHP=X"


def make_fasta_text(rng, lengths):
    contigs = []
    for k, length in enumerate(lengths):
        seq = [rng.choice("ACGT") for _ in range(length)]
        for _ in range(max(1, length // 120)):
            start = rng.randrange(length)
            run = rng.randint(1, 6)
            for i in range(start, min(length, start + run)):
                seq[i] = "N"
        for _ in range(max(1, length // 200)):
            start = rng.randrange(length)
            run = rng.randint(1, 25)
            for i in range(start, min(length, start + run)):
                seq[i] = seq[i].lower()
        for _ in range(length // 150):
            seq[rng.randrange(length)] = rng.choice("RYKMSW")
        contigs.append(("ctg%d" % (k + 1), "".join(seq)))
    return contigs


def _random_cigar(rng, allow_skip, max_blocks=6):
    ops = []
    if rng.random() < 0.08:
        ops.append((5, rng.randint(1, 9)))
    if rng.random() < 0.25:
        ops.append((4, rng.randint(1, 12)))
    n_blocks = rng.randint(1, max_blocks)
    for k in range(n_blocks):
        if k:
            x = rng.random()
            if x < 0.45:
                ops.append((1, rng.randint(1, 5)))
            elif x < 0.90:
                ops.append((2, rng.randint(1, 5)))
            elif allow_skip and x < 0.96:
                ops.append((3, rng.randint(1, 30)))
            elif allow_skip:
                ops.append((6, rng.randint(1, 3)))
            else:
                ops.append((2, rng.randint(1, 3)))
        ops.append((rng.choice((0, 0, 0, 7, 8)), rng.randint(1, 40)))
    if rng.random() < 0.25:
        ops.append((4, rng.randint(1, 12)))
    if rng.random() < 0.08:
        ops.append((5, rng.randint(1, 9)))
    return ops


def random_record(rng, name, contigs, readgroups, allow_skip=True, paired_rate=0.4,
                  no_rg_rate=0.0, filtered_rate=0.12, no_qual_rate=0.05):
    """One SAM line (list of fields)."""
    while True:
        cigar = _random_cigar(rng, allow_skip)
        span = sum(n for op, n in cigar if op in (0, 2, 3, 7, 8))
        tid = rng.randrange(len(contigs))
        cname, cseq = contigs[tid]
        if span <= len(cseq):
            break
    x = rng.random()
    if x < 0.08:
        pos = 0
    elif x < 0.16:
        pos = len(cseq) - span
    else:
        pos = rng.randint(0, len(cseq) - span)
    seq = []
    rpos = pos
    for op, n in cigar:
        if op in (0, 7, 8):
            for i in range(n):
                base = cseq[rpos + i].upper()
                if base not in "ACGT" or rng.random() < 0.08:
                    base = rng.choice("ACGT")
                elif base == "C" and rng.random() < 0.15:
                    base = "T"
                elif base == "G" and rng.random() < 0.15:
                    base = "A"
                seq.append(base)
            rpos += n
        elif op in (1, 4):
            seq.extend(rng.choice("ACGT") for _ in range(n))
        elif op in (2, 3):
            rpos += n
    for i in range(len(seq)):
        y = rng.random()
        if y < 0.02:
            seq[i] = "N"
        elif y < 0.025:
            seq[i] = rng.choice("RYKMSWBDHV=")
    seq = "".join(seq)
    if rng.random() < no_qual_rate:
        qual = "*"
    else:
        qual = "".join(chr(33 + rng.randint(0, 41)) for _ in seq)

    flag = 16 if rng.random() < 0.5 else 0
    rnext, pnext, tlen = "*", 0, 0
    if rng.random() < paired_rate:
        flag |= 1
        if rng.random() < 0.7:
            flag |= 2
        if rng.random() < 0.5:
            flag |= 0x20
        flag |= 0x40 if rng.random() < 0.5 else 0x80
        y = rng.random()
        if y < 0.75:
            rnext = "="
        elif y < 0.9 and len(contigs) > 1:
            rnext = contigs[(tid + 1) % len(contigs)][0]
        pnext = max(1, pos + 1 + rng.choice((-1, 1)) * rng.randint(0, 60)) if rnext != "*" else 0
        if rng.random() < 0.1 and rnext != "*":
            pnext = pos + 1
        tlen = rng.choice((-1, 1)) * rng.randint(0, 400)
    if rng.random() < filtered_rate:
        flag |= rng.choice((0x100, 0x200, 0x400, 0x800))
    fields = [name, str(flag), cname, str(pos + 1), str(rng.randint(0, 60)),
              "".join("%d%s" % (n, OPS[op]) for op, n in cigar), rnext, str(pnext), str(tlen),
              seq, qual]
    if readgroups and rng.random() >= no_rg_rate:
        fields.append("RG:Z:%s" % rng.choice(readgroups)[0])
    if rng.random() < 0.3:
        fields.append("NM:i:%d" % rng.randint(0, 5))
    return fields


def unmapped_record(rng, name):
    n = rng.randint(20, 60)
    seq = "".join(rng.choice("ACGT") for _ in range(n))
    qual = "".join(chr(33 + rng.randint(0, 41)) for _ in range(n))
    return [name, "4", "*", "0", "0", "*", "*", "0", "0", seq, qual]


def make_case(seed, n_reads, lengths=(400, 1500, 90), readgroups=None, allow_skip=True,
              no_rg_rate=0.0, paired_rate=0.4, with_unmapped=True):
    """Returns ``(fasta_contigs, sam_text)``."""
    rng = random.Random(seed)
    contigs = make_fasta_text(rng, lengths)
    if readgroups is None:
        readgroups = [("g2", "samB", "lib1"), ("g1", "samA", "lib2"), ("g0", "samA", "lib1"),
                      ("g3", "samA", "lib1")]
    lines = ["@HD\tVN:1.6\tSO:unsorted"]
    for name, seq in contigs:
        lines.append("@SQ\tSN:%s\tLN:%d" % (name, len(seq)))
    for rg_id, sample, library in readgroups:
        lines.append("@RG\tID:%s\tSM:%s\tLB:%s" % (rg_id, sample, library))
    for i in range(n_reads):
        if with_unmapped and rng.random() < 0.02:
            fields = unmapped_record(rng, "u%d" % i)
        else:
            fields = random_record(rng, "r%d" % i, contigs, readgroups, allow_skip=allow_skip,
                                   no_rg_rate=no_rg_rate, paired_rate=paired_rate)
        lines.append("\t".join(fields))
    return contigs, "\n".join(lines) + "\n"


def write_fasta(contigs, path, width=60):
    with open(path, "wt") as handle:
        for name, seq in contigs:
            handle.write(">%s\n" % name)
            for i in range(0, len(seq), width):
                handle.write(seq[i:i + width] + "\n")
