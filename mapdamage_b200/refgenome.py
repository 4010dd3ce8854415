"""Reference genome held as the device wants it: 4 bits per base, packed.

The reference fetches bases per read with ``pysam.FastaFile.fetch(...).upper()``
(``main.py:180``, ``align.py:32-33``, ``rescale.py:213``).  Here the genome is
uploaded once and gathered on the device.  Encoding (one nibble per base,
low nibble = even base, so base ``i`` of the packed stream is
``(word32[i >> 3] >> (4 * (i & 7))) & 15``):

    0..3 = A, C, G, T (case-insensitive: the reference uppercases)   7 = anything else

Only A/C/G/T ever count (``statistics.py:27,102``), so every other
character (N, IUPAC codes) collapses to one "not a base" code.  Contigs are
concatenated, each starting on a multiple of 8 bases so that a contig begins
on a 32-bit word.
"""
import numpy as np

CODE_OTHER = 7
_CODE_OF = np.full(256, CODE_OTHER, dtype=np.uint8)
for _code, _ch in enumerate("ACGT"):
    _CODE_OF[ord(_ch)] = _code
    _CODE_OF[ord(_ch.lower())] = _code


class ReferenceMismatch(ValueError):
    """The FASTA does not hold the contigs of the alignment file (``seq.compare_sequence_dicts`` returning False,
    ``main.py:139-145``)."""


class Reference:
    """Ordered contigs as ASCII ``uint8`` arrays plus the packed device image."""

    def __init__(self, names, sequences):
        self.names = list(names)
        self.sequences = [
            np.frombuffer(s.encode("latin-1"), dtype=np.uint8) if isinstance(s, str)
            else np.ascontiguousarray(s, dtype=np.uint8)
            for s in sequences
        ]
        self.lengths = [int(s.shape[0]) for s in self.sequences]
        self._packed = None

    @classmethod
    def from_fasta(cls, path):
        from .samtext import read_fasta

        seqs = read_fasta(path)
        return cls(list(seqs), list(seqs.values()))

    def reordered(self, names, lengths=None):
        """Contigs in the order of the BAM header (tid order).

        With ``lengths`` (the header's ``LN`` values) the two sequence dictionaries are compared the way
        ``main.py:139-145`` does through ``seq.compare_sequence_dicts`` (``seq.py:75-112``): a contig of the
        alignment file that the FASTA lacks, or whose length differs, is logged with the reference's messages
        and raises :class:`ReferenceMismatch` (the reference's ``main`` returns 1 there); FASTA-only contigs
        only draw the reference's warning.
        """
        index = {name: i for i, name in enumerate(self.names)}
        if lengths is not None:
            import logging

            log = logging.getLogger(__name__)
            bam = dict(zip(names, (int(x) for x in lengths)))
            fasta = dict(zip(self.names, self.lengths))
            if fasta != bam:
                common = set(fasta) & set(bam)
                if not common:
                    log.error("BAM and FASTA file have no sequence names in common")
                    raise ReferenceMismatch("BAM and FASTA file have no sequence names in common")
                different = [(key, fasta[key], bam[key]) for key in sorted(common) if fasta[key] != bam[key]]
                if different:
                    log.error("Length of required FASTA sequences differ:")
                    for values in different:
                        log.error(" - %s: %i vs %i bp" % values)
                bam_only = set(bam) - common
                if bam_only:
                    log.error("Sequences not found in FASTA:")
                    for key in bam_only:
                        log.error("%s (%i bp)", key, bam[key])
                fasta_only = set(fasta) - common
                if fasta_only:
                    log.warning("FASTA file contains extra sequences:")
                    for key in fasta_only:
                        log.warning(" - %s = %i bp", key, fasta[key])
                if different or bam_only:
                    raise ReferenceMismatch(
                        "FASTA and alignment file disagree: %s" % "; ".join(
                            ["%s: %i vs %i bp" % v for v in different]
                            + ["%s (%i bp) not found in FASTA" % (k, bam[k]) for k in sorted(bam_only)]))
        missing = [n for n in names if n not in index]
        if missing:
            raise ReferenceMismatch("Sequences not found in FASTA: %s" % ", ".join(missing))
        return Reference(names, [self.sequences[index[n]] for n in names])

    def packed(self):
        """``(packed_bytes, contig_base_offset[u64], contig_len[u32])``."""
        if self._packed is None:
            offsets, total = [], 0
            for length in self.lengths:
                offsets.append(total)
                total += (length + 7) & ~7
            codes = np.full(total + 8, CODE_OTHER, dtype=np.uint8)
            for off, seq in zip(offsets, self.sequences):
                codes[off:off + seq.shape[0]] = _CODE_OF[seq]
            packed = (codes[0::2] | (codes[1::2] << 4)).astype(np.uint8)
            self._packed = (
                packed,
                np.array(offsets, dtype=np.uint64),
                np.array(self.lengths, dtype=np.uint32),
            )
        return self._packed

    def write_fasta(self, path, width=60):
        with open(path, "wt") as handle:
            for name, seq in zip(self.names, self.sequences):
                handle.write(">%s\n" % name)
                text = seq.tobytes().decode("latin-1")
                for i in range(0, len(text), width):
                    handle.write(text[i:i + width] + "\n")
