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

    def reordered(self, names):
        """Contigs in the order of the BAM header (tid order)."""
        index = {name: i for i, name in enumerate(self.names)}
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
