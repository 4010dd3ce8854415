"""Host-side mirror of the reference's accumulator classes (``statistics.py``).

The reference mutates three objects once per read
(``MisincorporationRates`` ``statistics.py:9-55``, ``DNAComposition``
``:58-103``, ``FragmentLengths`` ``:106-137``) and serialises them with
``write()`` (``main.py:229-231``).  Here the per-read work happens on the GPU;
these classes keep the reference's constructor signatures, the nested ``.data``
dictionaries and the exact text format of ``write()``, and are filled from the
dense ``uint64`` slabs the device returns (``load``) instead of by ``update``.

Slab layouts (shared with the kernels and the oracle):

* misincorporation ``[lib][end 5p,3p][strand +,-][30 classes][L]``
* dnacomp ``[lib][end][strand][A,C,G,T][L + A]``: slot ``d < L`` is the read
  base at distance ``d`` from that end, slot ``L + d - 1`` the flanking
  reference base at distance ``d`` (1..A)
* fragment lengths ``[lib][kind pe,se][strand][lg_bins]``
"""
import csv
import logging
import os

import numpy as np

from . import seq as _seq

ENDS = ("5p", "3p")
STRANDS = ("+", "-")
KINDS = ("pe", "se")


def _check_libraries(libraries):
    libraries = list(libraries)
    if libraries != sorted(libraries):
        raise ValueError("libraries must be in sorted (sample, library) order")
    return libraries


class MisincorporationRates:
    """``statistics.py:9-55``; ``length`` = ``--length``."""

    def __init__(self, libraries, length):
        self.libraries = _check_libraries(libraries)
        self.length = length
        self.slab = np.zeros((len(self.libraries), 2, 2, _seq.N_CLASSES, length), dtype=np.uint64)

    def load(self, slab):
        self.slab[...] = np.asarray(slab, dtype=np.uint64).reshape(self.slab.shape)
        return self

    def column(self, lib, end, strand, name):
        """Counts for one printed column, positions 0..L-1."""
        part = self.slab[lib, ENDS.index(end), STRANDS.index(strand)]
        if name == "Total":
            return part[0:4].sum(axis=0)
        return part[_seq.device_class(name)]

    @property
    def data(self):
        """``data[(sample, library)][end][strand][column][index] -> int`` (``:10-20``)."""
        out = {}
        for li, library in enumerate(self.libraries):
            out[library] = {
                end: {
                    strand: {
                        name: dict(enumerate(int(x) for x in self.column(li, end, strand, name)))
                        for name in _seq.HEADER if name != "Total"
                    }
                    for strand in STRANDS
                }
                for end in ENDS
            }
        return out

    def write(self, filepath):
        """Same bytes as ``_write_freq_table(..., offset=1)`` (``:53-55,187-203``)."""
        lines = ["Sample\tLibrary\tEnd\tStd\tPos\t%s\n" % "\t".join(_seq.HEADER)]
        for li, (sample, library) in enumerate(self.libraries):
            for end in sorted(ENDS):
                for strand in sorted(STRANDS):
                    cols = np.stack(
                        [self.column(li, end, strand, name) for name in _seq.HEADER], axis=1
                    )
                    prefix = "%s\t%s\t%s\t%s\t" % (sample, library, end, strand)
                    for index in range(self.length):
                        lines.append(
                            prefix + str(index + 1) + "\t"
                            + "\t".join(str(int(x)) for x in cols[index]) + "\n"
                        )
        with open(filepath, "wt") as handle:
            handle.writelines(lines)


class DNAComposition:
    """``statistics.py:58-103``; ``around`` = ``--around``, ``length`` = ``--length``."""

    def __init__(self, libraries, around, length):
        self.libraries = _check_libraries(libraries)
        self.around = around
        self.length = length
        self.slab = np.zeros((len(self.libraries), 2, 2, 4, length + around), dtype=np.uint64)

    def load(self, slab):
        self.slab[...] = np.asarray(slab, dtype=np.uint64).reshape(self.slab.shape)
        return self

    def keys(self, end):
        """Printed positions of an end, ascending (``statistics.py:60-63``)."""
        length, around = self.length, self.around
        if end == "3p":
            return list(range(-length, 0)) + list(range(1, around + 1))
        return list(range(-around, 0)) + list(range(1, length + 1))

    def _slot(self, end, key):
        """Slab slot of printed position ``key``."""
        inside = key > 0 if end == "5p" else key < 0
        return abs(key) - 1 if inside else self.length + abs(key) - 1

    @property
    def data(self):
        out = {}
        for li, library in enumerate(self.libraries):
            out[library] = {}
            for ei, end in enumerate(ENDS):
                out[library][end] = {}
                for si, strand in enumerate(STRANDS):
                    out[library][end][strand] = {
                        letter: {
                            key: int(self.slab[li, ei, si, bi, self._slot(end, key)])
                            for key in self.keys(end)
                        }
                        for bi, letter in enumerate(_seq.LETTERS)
                    }
        return out

    def write(self, filepath):
        lines = ["Sample\tLibrary\tEnd\tStd\tPos\t%s\n" % "\t".join(_seq.LETTERS + ("Total",))]
        for li, (sample, library) in enumerate(self.libraries):
            for end in sorted(ENDS):
                ei = ENDS.index(end)
                slots = [self._slot(end, key) for key in self.keys(end)]
                for strand in sorted(STRANDS):
                    part = self.slab[li, ei, STRANDS.index(strand)]
                    prefix = "%s\t%s\t%s\t%s\t" % (sample, library, end, strand)
                    for key, slot in zip(self.keys(end), slots):
                        acgt = [int(part[b, slot]) for b in range(4)]
                        lines.append(
                            prefix + "%d\t%d\t%d\t%d\t%d\t%d\n" % (key, *acgt, sum(acgt))
                        )
        with open(filepath, "wt") as handle:
            handle.writelines(lines)


class FragmentLengths:
    """``statistics.py:106-137``; the dense histogram becomes the sparse dict."""

    def __init__(self, libraries):
        self.libraries = _check_libraries(libraries)
        self.data = {
            library: {(kind, strand): {} for kind in KINDS for strand in STRANDS}
            for library in self.libraries
        }

    def load(self, hist, overflow=()):
        """``hist[lib][kind][strand][length]``; ``overflow`` = extra
        ``(lib, kind, strand, length, count)`` rows beyond the dense bins."""
        hist = np.asarray(hist)
        for li, library in enumerate(self.libraries):
            for ki, kind in enumerate(KINDS):
                for si, strand in enumerate(STRANDS):
                    row = hist[li, ki, si]
                    nz = np.flatnonzero(row)
                    self.data[library][(kind, strand)] = {int(k): int(row[k]) for k in nz}
        for li, ki, si, length, count in overflow:
            table = self.data[self.libraries[li]][(KINDS[ki], STRANDS[si])]
            table[int(length)] = table.get(int(length), 0) + int(count)
        return self

    def write(self, filepath):
        with open(filepath, "wt") as handle:
            handle.write("Sample\tLibrary\tStd\tKind\tLength\tOccurences\n")
            for (sample, library), reads in sorted(self.data.items()):
                for (pe_or_se, strand), lengths in sorted(reads.items()):
                    for length, count in sorted(lengths.items()):
                        handle.write(
                            "%s\t%s\t%s\t%s\t%d\t%d\n"
                            % (sample, library, strand, pe_or_se, length, count)
                        )


def _first_position_counts(path):
    """Sums, over libraries and strands, the ``Pos == 1`` rows of a misincorporation table:
    ``{"5p": (C, C>T), "3p": (G, G>A)}``; ``None`` for a file without a header."""
    wanted = {"5p": ("C", "C>T"), "3p": ("G", "G>A")}
    sums = {end: [0, 0] for end in wanted}
    with open(path, newline="") as handle:
        rows = csv.DictReader(handle, delimiter="\t")
        if not rows.fieldnames:
            return None
        for row in rows:
            if int(row["Pos"]) != 1:
                continue
            base, mutation = wanted[row["End"]]
            sums[row["End"]][0] += int(row[base])
            sums[row["End"]][1] += int(row[mutation])
    return {end: tuple(v) for end, v in sums.items()}


def check_table_and_warn_if_dmg_freq_is_low(folder):
    """Host mirror of the reference's pre-flight check for the Bayesian stage (``statistics.py:140-184``):
    ``True`` when ``misincorporation.txt`` in ``folder`` has C (5') and G (3') counts at the first position;
    warns when C>T + G>A there is below 1 %.  Messages and return values are the reference's."""
    log = logging.getLogger(__name__)
    name = "misincorporation.txt"
    try:
        counts = _first_position_counts(os.path.join(folder, name))
    except (csv.Error, IOError, OSError, KeyError) as error:
        log.error("Error reading misincorporation table: %s", error)
        return False
    if counts is None:
        log.error("%r is empty; please re-run mapDamage", name)
        return False
    (c_total, c_to_t), (g_total, g_to_a) = counts["5p"], counts["3p"]
    if not (c_total and g_total):
        log.error("Insufficient data in %r; cannot perform Bayesian computation", name)
        return False
    damage = 0.0 + c_to_t / c_total + g_to_a / g_total
    if damage < 0.01:
        log.warning("DNA damage levels are too low, the Bayesian computation should not be "
                    "performed (%f < 0.01)", damage)
    return True
