"""Seeded synthetic aDNA alignments, generated directly as SoA batches.

Implements the synthetic inputs of SURVEY.md section 8(d): a uniform random
reference and reads drawn from it with post-mortem damage on the read strand
(C->T with p = 0.3 * 0.7**i at 5' distance i, G->A mirrored at the 3' end; the
profile is truncated to 0 beyond i = 23), a small uniform sequencing-error
rate, and -- for the mixed configuration -- 50-150 bp proper pairs whose CIGARs
mix plain matches, 1-3 bp insertions, 1-3 bp deletions and 1-10 bp soft clips.
Everything is vectorised numpy so that tens of millions of reads can be made
in seconds per million; chunks are seeded independently (``SeedSequence.spawn``)
and can be generated on several threads.

In BAM (forward-strand) orientation the damage looks the same on both
strands: C->T decaying from the left end of the alignment and G->A from the
right end; the strand only decides which table the counts land in.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .batch import ReadBatch
from .refgenome import Reference, CODE_OTHER, _CODE_OF

DAMAGE_REACH = 24
_DAMAGE = np.zeros(DAMAGE_REACH + 1, dtype=np.float32)
_DAMAGE[:DAMAGE_REACH] = 0.3 * 0.7 ** np.arange(DAMAGE_REACH)

_OP_M, _OP_I, _OP_D, _OP_S = 0, 1, 2, 4


def make_reference(lengths, seed=1, names=None, other_rate=0.0):
    """Uniform random A/C/G/T contigs; ``other_rate`` sprinkles ``N``."""
    rng = np.random.default_rng(np.random.SeedSequence([seed, 0x5EF]))
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = []
    for length in lengths:
        seq = letters[rng.integers(0, 4, size=length, dtype=np.uint8)]
        if other_rate:
            seq = seq.copy()
            seq[rng.random(length) < other_rate] = ord("N")
        seqs.append(seq)
    if names is None:
        names = ["chr%d" % (i + 1) for i in range(len(lengths))]
    return Reference(names, seqs)


def _chunk(ref_codes, contig_start, contig_len, n, rng, length, mix, paired, error_rate,
           with_qual, n_libs, read_n_rate, filtered_rate):
    lo, hi = length
    if paired:
        n += n & 1
    l_seq = rng.integers(lo, hi + 1, size=n, dtype=np.int32)
    kind = rng.choice(4, size=n, p=np.asarray(mix, dtype=np.float64) / sum(mix)).astype(np.int8)
    # too-short reads cannot host an indel with 5 bp anchors
    kind[(l_seq < 16) & ((kind == 1) | (kind == 2))] = 0
    k = rng.integers(1, 4, size=n, dtype=np.int32)
    s1 = np.where(kind == 3, rng.integers(0, 11, size=n, dtype=np.int32), 0)
    s2 = np.where(kind == 3, rng.integers(0, 11, size=n, dtype=np.int32), 0)
    s1[(kind == 3) & (s1 == 0) & (s2 == 0)] = 1
    # keep at least 10 aligned bases
    over = (kind == 3) & (l_seq - s1 - s2 < 10)
    s1[over] = 1
    s2[over] = 0
    nq = l_seq - s1 - s2
    is_i, is_d = kind == 1, kind == 2
    k = np.where(is_i | is_d, k, 0)
    span_a = np.where(is_i, nq - k - 10, nq - 10)
    a = 5 + (rng.random(n) * np.maximum(span_a, 1)).astype(np.int32)
    a = np.where(is_i | is_d, a, nq)
    rspan = nq - np.where(is_i, k, 0) + np.where(is_d, k, 0)

    weights = contig_len.astype(np.float64) / contig_len.sum()
    tid = rng.choice(contig_len.shape[0], size=n, p=weights).astype(np.int32)
    room = contig_len[tid] - rspan
    if np.any(room < 0):
        raise ValueError("reference contig shorter than a read")
    pos = (rng.random(n) * (room + 1)).astype(np.int64)
    pos = np.minimum(pos, room).astype(np.int32)
    reverse = rng.random(n) < 0.5

    flag = np.zeros(n, dtype=np.uint16)
    tlen = np.zeros(n, dtype=np.int32)
    mtid = np.full(n, -1, dtype=np.int32)
    mpos = np.full(n, -1, dtype=np.int32)
    if paired:
        # records 2i / 2i+1 are mates on the same contig: leftmost forward,
        # rightmost reverse (inward-facing proper pair)
        left, right = slice(0, n, 2), slice(1, n, 2)
        tid[right] = tid[left]
        gap = rng.integers(0, 301, size=n // 2, dtype=np.int32)
        room_r = contig_len[tid[right]] - rspan[right]
        pos[right] = np.minimum(pos[left] + gap, room_r)
        pos[left] = np.minimum(pos[left], pos[right])
        reverse[left], reverse[right] = False, True
        first_left = rng.random(n // 2) < 0.5
        flag[left] = np.where(first_left, 99, 163)
        flag[right] = np.where(first_left, 147, 83)
        end_r = pos[right] + rspan[right]
        end_l = pos[left] + rspan[left]
        frag = np.maximum(end_r, end_l) - pos[left]
        tlen[left], tlen[right] = frag, -frag
        mtid[:] = tid
        mpos[left], mpos[right] = pos[right], pos[left]
    else:
        flag |= np.where(reverse, 16, 0).astype(np.uint16)
    if filtered_rate:
        hit = rng.random(n) < filtered_rate
        bits = np.array([0x100, 0x200, 0x400, 0x800], dtype=np.uint16)
        flag[hit] |= bits[rng.integers(0, 4, size=int(hit.sum()))]

    # ---- per-base work, flat over all reads of the chunk -------------
    padded = (l_seq + 1) & ~1
    base_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(padded, out=base_off[1:])
    read_start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(l_seq, out=read_start[1:])
    total = int(read_start[-1])
    rid = np.repeat(np.arange(n, dtype=np.int32), l_seq)
    j = (np.arange(total, dtype=np.int64) - read_start[rid]).astype(np.int32)
    jq = j - s1[rid]
    nq_f = nq[rid]
    aligned = (jq >= 0) & (jq < nq_f)
    a_f, k_f = a[rid], k[rid]
    ins = is_i[rid] & (jq >= a_f) & (jq < a_f + k_f)
    aligned &= ~ins
    shift = np.where(is_i[rid] & (jq >= a_f + k_f), -k_f, 0) + np.where(is_d[rid] & (jq >= a_f), k_f, 0)
    gidx = contig_start[tid[rid]] + pos[rid] + jq + shift
    gidx = np.where(aligned, gidx, 0)
    base = ref_codes[gidx]
    rand_base = rng.integers(0, 4, size=total, dtype=np.uint8)
    base = np.where(aligned & (base < 4), base, rand_base)
    u = rng.random(total, dtype=np.float32)
    p5 = _DAMAGE[np.minimum(jq, DAMAGE_REACH).clip(0)]
    p3 = _DAMAGE[np.minimum(nq_f - 1 - jq, DAMAGE_REACH).clip(0)]
    base = np.where(aligned & (base == 1) & (u < p5), 3, base)  # C -> T
    base = np.where(aligned & (base == 2) & (u < p3), 0, base)  # G -> A
    if error_rate:
        n_err = rng.binomial(total, error_rate)
        where = rng.integers(0, total, size=n_err)
        base[where] = (base[where] + rng.integers(1, 4, size=n_err, dtype=np.uint8)) & 3
    nib = (np.uint8(1) << base).astype(np.uint8)
    if read_n_rate:
        nib[rng.random(total) < read_n_rate] = 15

    dest = base_off[rid] + j
    nibbuf = np.zeros(int(base_off[-1]), dtype=np.uint8)
    nibbuf[dest] = nib
    seq4 = (nibbuf[0::2] << 4) | nibbuf[1::2]
    qual = None
    if with_qual:
        qual = np.zeros(int(base_off[-1]), dtype=np.uint8)
        qual[dest] = rng.integers(2, 41, size=total, dtype=np.uint8)

    # ---- CIGARs: up to three ops per read ----------------------------
    ops = np.zeros((n, 3), dtype=np.uint32)
    valid = np.zeros((n, 3), dtype=bool)
    plain = kind == 0
    ops[plain, 0] = (nq[plain].astype(np.uint32) << 4) | _OP_M
    valid[plain, 0] = True
    for mask, op in ((is_i, _OP_I), (is_d, _OP_D)):
        b = nq - a - (k if op == _OP_I else 0)
        ops[mask, 0] = (a[mask].astype(np.uint32) << 4) | _OP_M
        ops[mask, 1] = (k[mask].astype(np.uint32) << 4) | op
        ops[mask, 2] = (b[mask].astype(np.uint32) << 4) | _OP_M
        valid[mask, :] = True
    clip = kind == 3
    ops[clip, 0] = (s1[clip].astype(np.uint32) << 4) | _OP_S
    ops[clip, 1] = (nq[clip].astype(np.uint32) << 4) | _OP_M
    ops[clip, 2] = (s2[clip].astype(np.uint32) << 4) | _OP_S
    valid[clip, 0] = s1[clip] > 0
    valid[clip, 1] = True
    valid[clip, 2] = s2[clip] > 0
    cigar = ops[valid]
    cigar_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(valid.sum(axis=1), out=cigar_off[1:])

    lib = rng.integers(0, n_libs, size=n, dtype=np.uint16) if n_libs > 1 else np.zeros(n, np.uint16)
    return dict(flag=flag, tid=tid, pos=pos, lib=lib, l_seq=l_seq.astype(np.uint32),
                base_off=base_off[:-1], n_bases=int(base_off[-1]), cigar_off=cigar_off,
                cigar=cigar, seq4=seq4, qual=qual, tlen=tlen, mtid=mtid, mpos=mpos)


def simulate_reads(reference, n, seed=1, length=(100, 100), mix=(1, 0, 0, 0), paired=False,
                   error_rate=0.002, with_qual=True, n_libs=1, read_n_rate=0.0,
                   filtered_rate=0.0, chunk=250_000, threads=4):
    """Returns a :class:`ReadBatch` of ``n`` synthetic alignments.

    ``mix`` = relative weights of (plain match, one insertion, one deletion,
    soft-clipped) reads.  ``paired`` makes inward-facing proper pairs
    (flags 99/147 or 163/83).  Deterministic in ``seed`` and ``chunk``.
    """
    if isinstance(length, int):
        length = (length, length)
    ref_codes = np.concatenate([_CODE_OF[s] for s in reference.sequences])
    contig_len = np.array(reference.lengths, dtype=np.int64)
    contig_start = np.zeros(len(contig_len), dtype=np.int64)
    np.cumsum(contig_len[:-1], out=contig_start[1:])
    if paired:
        n += n & 1
        chunk += chunk & 1
    sizes = [min(chunk, n - i) for i in range(0, n, chunk)] or [0]
    seeds = np.random.SeedSequence([seed, 0xADA]).spawn(len(sizes))

    def work(args):
        size, ss = args
        return _chunk(ref_codes, contig_start, contig_len, size, np.random.default_rng(ss),
                      length, mix, paired, error_rate, with_qual, n_libs, read_n_rate,
                      filtered_rate)

    if threads > 1 and len(sizes) > 1:
        with ThreadPoolExecutor(threads) as pool:
            parts = list(pool.map(work, zip(sizes, seeds)))
    else:
        parts = [work(x) for x in zip(sizes, seeds)]
    return concat_parts(parts, with_qual)


def concat_parts(parts, with_qual):
    base_shift = np.cumsum([0] + [p["n_bases"] for p in parts])
    cigar_shift = np.cumsum([0] + [p["cigar"].shape[0] for p in parts])
    if base_shift[-1] >= 1 << 32:
        raise ValueError("batch exceeds 2^32 bases; generate it in several batches")
    cat = np.concatenate
    cigar_off = cat([p["cigar_off"][:-1] + s for p, s in zip(parts, cigar_shift)]
                    + [cigar_shift[-1:]])
    return ReadBatch(
        flag=cat([p["flag"] for p in parts]), tid=cat([p["tid"] for p in parts]),
        pos=cat([p["pos"] for p in parts]), lib=cat([p["lib"] for p in parts]),
        l_seq=cat([p["l_seq"] for p in parts]),
        base_off=cat([p["base_off"] + s for p, s in zip(parts, base_shift)]),
        cigar_off=cigar_off, cigar=cat([p["cigar"] for p in parts]),
        seq4=cat([p["seq4"] for p in parts]),
        qual=cat([p["qual"] for p in parts]) if with_qual else None,
        tlen=cat([p["tlen"] for p in parts]), mtid=cat([p["mtid"] for p in parts]),
        mpos=cat([p["mpos"] for p in parts]),
    )


def write_sam(batch, reference, path, readgroups=None, lib_to_rg=None):
    """SAM text of a batch (tests only; per-record Python).

    ``readgroups``: list of ``(ID, SM, LB)``; ``lib_to_rg[lib]`` names the read
    group written for a read of library index ``lib``.
    """
    from .samtext import format_cigar

    with open(path, "wt") as handle:
        handle.write("@HD\tVN:1.6\tSO:unsorted\n")
        for name, length in zip(reference.names, reference.lengths):
            handle.write("@SQ\tSN:%s\tLN:%d\n" % (name, length))
        for rg_id, sample, library in readgroups or ():
            handle.write("@RG\tID:%s\tSM:%s\tLB:%s\n" % (rg_id, sample, library))
        for i in range(batch.n):
            tid = int(batch.tid[i])
            mtid = int(batch.mtid[i])
            rnext = "*" if mtid < 0 else ("=" if mtid == tid else reference.names[mtid])
            fields = [
                batch.names[i] if batch.names is not None else "r%d" % i,
                str(int(batch.flag[i])), reference.names[tid], str(int(batch.pos[i]) + 1), "37",
                format_cigar(batch.cigar_of(i)), rnext, str(int(batch.mpos[i]) + 1),
                str(int(batch.tlen[i])), batch.sequence_of(i), batch.qualities_of(i) or "*",
            ]
            if lib_to_rg is not None:
                fields.append("RG:Z:%s" % lib_to_rg[int(batch.lib[i])])
            handle.write("\t".join(fields) + "\n")
