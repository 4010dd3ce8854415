"""Host-side SoA read batches: the unit handed across the C ABI.

A :class:`ReadBatch` is the struct-of-arrays image of a run of BAM alignment
records, holding exactly the fields the hot path reads (reference
``main.py:165-217``, ``rescale.py:300-344``): flag, reference id, position,
library id, CIGAR ops (BAM ``len<<4|op``), the 4-bit packed read sequence,
base qualities, and the mate fields used by the fragment-length histogram and
the rescale pairing rule.  Layout (see ``include/mapdamage_b200.h``):

* ``base_off[i]`` is the read's first base in a per-batch *base* coordinate in
  which every read starts on an even base: the read's packed sequence begins at
  byte ``base_off[i] // 2`` of ``seq4`` (BAM nibble order, high nibble first)
  and its qualities at byte ``base_off[i]`` of ``qual``.
* ``qual`` holds raw Phred values (not +33); a read without qualities has
  ``0xFF`` in its first quality byte (the BAM convention).
* ``cigar_off`` has ``n + 1`` entries.

:class:`BatchBuilder` applies the reference's read filter
(``reader.py:121-132``) and library lookup (``reader.py:63-81``).
"""
import numpy as np

from . import seq as _seq

CIGAR_OPS = "MIDNSHP=X"
FILTERED_FLAGS = 0x4 | 0x100 | 0x200 | 0x400 | 0x800  # reader.py:9-13,121-132

_NIBBLE_OF = np.full(256, 15, dtype=np.uint8)
for _i, _ch in enumerate("=ACMGRSVTWYHKDBN"):
    _NIBBLE_OF[ord(_ch)] = _i
    _NIBBLE_OF[ord(_ch.lower())] = _i


class BAMError(RuntimeError):
    """Mirror of the reference's ``reader.BAMError`` (``reader.py:16``)."""


class ReadBatch:
    """Struct-of-arrays batch of alignment records (numpy, host memory)."""

    FIELDS = (
        ("flag", np.uint16), ("tid", np.int32), ("pos", np.int32), ("lib", np.uint16),
        ("l_seq", np.uint32), ("base_off", np.uint32), ("cigar_off", np.uint32),
        ("tlen", np.int32), ("mtid", np.int32), ("mpos", np.int32),
    )

    def __init__(self, **arrays):
        self.flag = np.ascontiguousarray(arrays["flag"], dtype=np.uint16)
        n = self.n = int(self.flag.shape[0])
        self.tid = np.ascontiguousarray(arrays["tid"], dtype=np.int32)
        self.pos = np.ascontiguousarray(arrays["pos"], dtype=np.int32)
        self.lib = np.ascontiguousarray(arrays.get("lib", np.zeros(n)), dtype=np.uint16)
        self.l_seq = np.ascontiguousarray(arrays["l_seq"], dtype=np.uint32)
        self.base_off = np.ascontiguousarray(arrays["base_off"], dtype=np.uint32)
        self.cigar_off = np.ascontiguousarray(arrays["cigar_off"], dtype=np.uint32)
        self.cigar = np.ascontiguousarray(arrays["cigar"], dtype=np.uint32)
        self.seq4 = np.ascontiguousarray(arrays["seq4"], dtype=np.uint8)
        qual = arrays.get("qual")
        self.qual = None if qual is None else np.ascontiguousarray(qual, dtype=np.uint8)
        zeros = np.zeros(n, dtype=np.int32)
        self.tlen = np.ascontiguousarray(arrays.get("tlen", zeros), dtype=np.int32)
        self.mtid = np.ascontiguousarray(arrays.get("mtid", zeros - 1), dtype=np.int32)
        self.mpos = np.ascontiguousarray(arrays.get("mpos", zeros - 1), dtype=np.int32)
        self.names = arrays.get("names")
        self.validate()

    def validate(self):
        n = self.n
        for name in ("tid", "pos", "lib", "l_seq", "base_off", "tlen", "mtid", "mpos"):
            if getattr(self, name).shape != (n,):
                raise ValueError("field %r must have shape (%d,)" % (name, n))
        if self.cigar_off.shape != (n + 1,):
            raise ValueError("cigar_off must have n + 1 entries")
        if n and int(self.cigar_off[-1]) != self.cigar.shape[0]:
            raise ValueError("cigar_off[n] must equal the number of CIGAR ops")
        if n and np.any(self.base_off & 1):
            raise ValueError("every read must start on an even base offset")
        if n:
            end = int(self.base_off[-1]) + int(self.l_seq[-1])
            if self.seq4.shape[0] < (end + 1) // 2:
                raise ValueError("seq4 is shorter than base_off/l_seq imply")
            if self.qual is not None and self.qual.shape[0] < end:
                raise ValueError("qual is shorter than base_off/l_seq imply")

    def droppable(self):
        """Names of the optional ``mdg_batch`` arrays that hold exactly what a NULL pointer stands for
        (``include/mapdamage_b200.h``), so they need not cross PCIe.  Computed once and cached."""
        if getattr(self, "_droppable", None) is None:
            n = self.n
            drop = set()
            if n:
                if not self.tid.any():
                    drop.add("tid")
                consumes = np.isin(self.cigar & 0xF, (0, 1, 4, 7, 8))
                from_cigar = np.add.reduceat(np.where(consumes, self.cigar >> 4, 0).astype(np.int64),
                                             self.cigar_off[:-1].astype(np.int64)) if self.cigar.shape[0] else None
                if from_cigar is not None and np.all(np.diff(self.cigar_off.astype(np.int64)) > 0) \
                        and np.array_equal(from_cigar, self.l_seq):
                    drop.add("l_seq")
                if not self.lib.any():
                    drop.add("lib")
                if not (self.flag & 1).any():
                    drop.add("tlen")
                if (self.mtid == -1).all() and (self.mpos == -1).all():
                    drop.update(("mtid", "mpos"))
                if self.cigar.shape[0] == n and np.array_equal(self.cigar_off, np.arange(n + 1, dtype=np.uint32)):
                    drop.add("cigar_off")
                    if n > 1 and (self.cigar == self.cigar[0]).all():
                        drop.add("cigar")  # one CIGAR word shared by every read: a single word crosses PCIe
                padded = (self.l_seq.astype(np.int64) + 1) & ~1
                packed = np.zeros(n, dtype=np.int64)
                np.cumsum(padded[:-1], out=packed[1:])
                if np.array_equal(self.base_off, packed):
                    drop.add("base_off")
            self._droppable = frozenset(drop)
        return self._droppable

    def invalidate(self):
        """Call after editing the arrays in place: forgets what :meth:`droppable` cached."""
        self._droppable = None

    @property
    def total_bases(self):
        """Base slots spanned by the batch (incl. odd-length pad slots)."""
        if not self.n:
            return 0
        return int(self.base_off[-1]) + int(self.l_seq[-1]) + (int(self.l_seq[-1]) & 1)

    def nbytes(self):
        total = self.cigar.nbytes + self.seq4.nbytes
        total += sum(getattr(self, name).nbytes for name, _ in self.FIELDS)
        if self.qual is not None:
            total += self.qual.nbytes
        return total

    # -- per-record accessors (tests, SAM export) ----------------------
    def cigar_of(self, i):
        ops = self.cigar[self.cigar_off[i]:self.cigar_off[i + 1]]
        return [(int(c) & 0xF, int(c) >> 4) for c in ops]

    def sequence_of(self, i):
        off, n = int(self.base_off[i]), int(self.l_seq[i])
        packed = self.seq4[off // 2:off // 2 + (n + 1) // 2]
        nib = np.empty(packed.shape[0] * 2, dtype=np.uint8)
        nib[0::2] = packed >> 4
        nib[1::2] = packed & 0xF
        return "".join("=ACMGRSVTWYHKDBN"[x] for x in nib[:n])

    def qualities_of(self, i):
        """Phred+33 string, or ``None`` if the read has no qualities."""
        if self.qual is None:
            return None
        off, n = int(self.base_off[i]), int(self.l_seq[i])
        q = self.qual[off:off + n]
        if n and q[0] == 0xFF:
            return None
        return (q + 33).astype(np.uint8).tobytes().decode("latin-1")

    def select(self, index):
        """New batch holding the records at ``index`` (any order)."""
        index = np.asarray(index, dtype=np.int64)
        builder = _Concat()
        for i in index:
            i = int(i)
            off, n = int(self.base_off[i]), int(self.l_seq[i])
            builder.add(
                flag=self.flag[i], tid=self.tid[i], pos=self.pos[i], lib=self.lib[i],
                cigar=self.cigar[self.cigar_off[i]:self.cigar_off[i + 1]],
                seq4=self.seq4[off // 2:off // 2 + (n + 1) // 2], l_seq=n,
                qual=None if self.qual is None else self.qual[off:off + n],
                tlen=self.tlen[i], mtid=self.mtid[i], mpos=self.mpos[i],
                name=None if self.names is None else self.names[i],
            )
        return builder.finish(with_qual=self.qual is not None)

    def split(self, parts):
        """Contiguous split into ``parts`` batches (multi-GPU sharding by range)."""
        bounds = np.linspace(0, self.n, parts + 1).astype(np.int64)
        return [self.slice(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:])]

    def slice(self, start, stop):
        """Contiguous sub-batch ``[start, stop)`` without per-record work."""
        start, stop = int(start), int(stop)
        if stop <= start:
            return empty_batch(with_qual=self.qual is not None)
        b0 = int(self.base_off[start])
        last = stop - 1
        b1 = int(self.base_off[last]) + int(self.l_seq[last])
        b1 += b1 & 1
        c0, c1 = int(self.cigar_off[start]), int(self.cigar_off[stop])
        return ReadBatch(
            flag=self.flag[start:stop], tid=self.tid[start:stop], pos=self.pos[start:stop],
            lib=self.lib[start:stop], l_seq=self.l_seq[start:stop],
            base_off=self.base_off[start:stop] - np.uint32(b0),
            cigar_off=self.cigar_off[start:stop + 1] - np.uint32(c0),
            cigar=self.cigar[c0:c1], seq4=self.seq4[b0 // 2:b1 // 2],
            qual=None if self.qual is None else self.qual[b0:b1],
            tlen=self.tlen[start:stop], mtid=self.mtid[start:stop], mpos=self.mpos[start:stop],
            names=None if self.names is None else self.names[start:stop],
        )


class _Concat:
    """Appends already-encoded records; ``finish`` lays them out as a batch."""

    def __init__(self):
        self.rows = []

    def add(self, **row):
        self.rows.append(row)

    def finish(self, with_qual=True):
        rows = self.rows
        n = len(rows)
        if not n:
            return empty_batch(with_qual=with_qual)
        l_seq = np.array([r["l_seq"] for r in rows], dtype=np.uint32)
        padded = (l_seq.astype(np.int64) + 1) & ~1
        base_off = np.zeros(n, dtype=np.int64)
        np.cumsum(padded[:-1], out=base_off[1:])
        total = int(base_off[-1] + padded[-1])
        if total >= 1 << 32:
            raise ValueError("batch exceeds 2^32 bases; split it")
        n_ops = np.array([len(r["cigar"]) for r in rows], dtype=np.int64)
        cigar_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(n_ops, out=cigar_off[1:])
        cigar = np.zeros(int(cigar_off[-1]), dtype=np.uint32)
        seq4 = np.zeros(total // 2, dtype=np.uint8)
        qual = np.full(total, 0xFF, dtype=np.uint8) if with_qual else None
        for i, r in enumerate(rows):
            cigar[cigar_off[i]:cigar_off[i + 1]] = r["cigar"]
            off = int(base_off[i])
            nb = (int(l_seq[i]) + 1) // 2
            seq4[off // 2:off // 2 + nb] = r["seq4"][:nb]
            if with_qual and r["qual"] is not None:
                qual[off:off + int(l_seq[i])] = r["qual"]
        names = [r.get("name") for r in rows]
        return ReadBatch(
            flag=[r["flag"] for r in rows], tid=[r["tid"] for r in rows],
            pos=[r["pos"] for r in rows], lib=[r["lib"] for r in rows], l_seq=l_seq,
            base_off=base_off, cigar_off=cigar_off, cigar=cigar, seq4=seq4, qual=qual,
            tlen=[r["tlen"] for r in rows], mtid=[r["mtid"] for r in rows],
            mpos=[r["mpos"] for r in rows],
            names=None if any(x is None for x in names) else names,
        )


def concatenate(batches):
    """One batch holding the records of ``batches`` in order (all with or all without qualities)."""
    batches = [b for b in batches if b.n]
    if not batches:
        return empty_batch()
    base_shift = np.cumsum([0] + [b.total_bases for b in batches])
    cigar_shift = np.cumsum([0] + [b.cigar.shape[0] for b in batches])
    if base_shift[-1] >= 1 << 32:
        raise ValueError("batch exceeds 2^32 bases")
    with_qual = all(b.qual is not None for b in batches)
    cat = np.concatenate
    return ReadBatch(
        flag=cat([b.flag for b in batches]), tid=cat([b.tid for b in batches]), pos=cat([b.pos for b in batches]),
        lib=cat([b.lib for b in batches]), l_seq=cat([b.l_seq for b in batches]),
        base_off=cat([b.base_off.astype(np.int64) + s for b, s in zip(batches, base_shift)]),
        cigar_off=cat([b.cigar_off[:-1].astype(np.int64) + s for b, s in zip(batches, cigar_shift)] + [cigar_shift[-1:]]),
        cigar=cat([b.cigar for b in batches]), seq4=cat([b.seq4[:b.total_bases // 2] for b in batches]),
        qual=cat([b.qual[:b.total_bases] for b in batches]) if with_qual else None,
        tlen=cat([b.tlen for b in batches]), mtid=cat([b.mtid for b in batches]), mpos=cat([b.mpos for b in batches]),
    )


def empty_batch(with_qual=True):
    z = np.zeros(0)
    return ReadBatch(flag=z, tid=z, pos=z, lib=z, l_seq=z, base_off=z,
                     cigar_off=np.zeros(1), cigar=z, seq4=z,
                     qual=z if with_qual else None, tlen=z, mtid=z, mpos=z)


def pack_sequence(text):
    """ASCII read sequence -> BAM 4-bit packed bytes (high nibble first)."""
    nib = _NIBBLE_OF[np.frombuffer(text.encode("latin-1"), dtype=np.uint8)]
    if nib.shape[0] & 1:
        nib = np.concatenate([nib, np.zeros(1, dtype=np.uint8)])
    return ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8)


def encode_cigar(cigar):
    """``[(op, len), ...]`` -> BAM ``len << 4 | op`` words."""
    return np.array([(n << 4) | op for op, n in cigar], dtype=np.uint32)


class BatchBuilder:
    """Collects alignment records into a :class:`ReadBatch`.

    ``libraries`` maps a read-group ID to ``(sample, library)`` the way
    ``BAMReader._collect_readgroups`` does (``reader.py:98-118``); with
    ``merge_libraries`` every read lands in ``("*", "*")`` (``reader.py:44-46``).
    ``apply_filter`` reproduces ``BAMReader._filter_reads``; the rescale pass
    sees every record (``rescale.py:300``) and so builds with it off.
    """

    def __init__(self, readgroups=None, merge_libraries=False, apply_filter=True):
        self.merge_libraries = merge_libraries
        self.apply_filter = apply_filter
        if merge_libraries:
            self.readgroups = {None: ("*", "*")}
        else:
            self.readgroups = dict(readgroups or {})
        # library index = rank in the sorted (sample, library) order used when
        # the tables are written (statistics.py:190)
        self.libraries = sorted(set(self.readgroups.values()))
        self._lib_index = {key: i for i, key in enumerate(self.libraries)}
        self._rows = _Concat()
        self._with_qual = False
        self.n_seen = 0

    def library_of(self, record):
        """``BAMReader.get_sample_and_library`` (``reader.py:63-81``)."""
        if self.merge_libraries:
            return self.readgroups[None]
        tags = record.tags
        if "RG" not in tags:
            raise BAMError(
                "Read %r has no read-group. Either fix BAM or use --merge-libraries"
                % (record.qname,)
            )
        try:
            return self.readgroups[tags["RG"]]
        except KeyError:
            raise BAMError(
                "Read %r has read-group not listed in BAM header (%r); either fix BAM "
                "or use --merge-libraries" % (record.qname, tags["RG"])
            )

    def add(self, record):
        """Adds one record; returns False if the read filter dropped it."""
        self.n_seen += 1
        if self.apply_filter and (record.flag & FILTERED_FLAGS):
            return False
        lib = self._lib_index[self.library_of(record)] if self.apply_filter else 0
        if record.seq is None or not record.cigar:
            if self.apply_filter or not (record.flag & 0x4):
                # the reference dies with TypeError here
                raise BAMError(
                    "Read %r is mapped but has no sequence or CIGAR" % (record.qname,)
                )
        seq = record.seq or ""
        qual = None
        if record.qual is not None:
            qual = np.frombuffer(record.qual.encode("latin-1"), dtype=np.uint8) - 33
            self._with_qual = True
        self._rows.add(
            flag=record.flag, tid=record.tid, pos=record.pos, lib=lib,
            cigar=encode_cigar(record.cigar or []), seq4=pack_sequence(seq), l_seq=len(seq),
            qual=qual, tlen=record.tlen, mtid=record.mtid, mpos=record.mpos,
            name=record.qname,
        )
        return True

    def finish(self, with_qual=None):
        with_qual = self._with_qual if with_qual is None else with_qual
        batch = self._rows.finish(with_qual=with_qual)
        self._rows = _Concat()
        return batch


assert _seq.LETTERS == ("A", "C", "G", "T")
