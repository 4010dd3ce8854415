"""Column order of the output tables and nucleotide helpers.

The column order is the on-disk contract consumed by the R stage
(reference ``seq.py:6-30`` defines it; ``r/mapDamage.r`` and
``r/stats/data.r`` read the columns by name).
"""
LETTERS = tuple("ACGT")

# 12 substitutions, 4 deletions, 4 insertions, soft clips -- in the order the
# reference prints them (reference seq.py:7-29)
_SUBSTITUTIONS = "GA CT AG TC AC AT CG CA TG TA GC GT".split()
_DELETED = "ATCG"
_INSERTED = "ATCG"
MUTATIONS = (
    tuple("%s>%s" % (p[0], p[1]) for p in _SUBSTITUTIONS)
    + tuple("%s>-" % b for b in _DELETED)
    + tuple("->%s" % b for b in _INSERTED)
    + ("S",)
)
HEADER = LETTERS + ("Total",) + MUTATIONS

# Device-side class layout of the misincorporation slab:
#   0..3   reference base counts A, C, G, T
#   4+5g+b pair (reference g, read b), g/b in A,C,G,T,gap = 0..4, g != b
#   29     soft clips
N_CLASSES = 30
SOFTCLIP_CLASS = 29
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "-": 4}


def device_class(column):
    """Index in the device slab of a printed column name (not ``Total``)."""
    if column in LETTERS:
        return _CODE[column]
    if column == "S":
        return SOFTCLIP_CLASS
    ref, read = column.split(">")
    return 4 + 5 * _CODE[ref] + _CODE[read]


_COMPLEMENT = str.maketrans("ACGTMRWSYKVHDBacgtmrwsykvhdb", "TGCAKYWSRMBDHVtgcakywsrmbdhv")


def revcomp(text):
    """Reverse complement (IUPAC aware, like reference ``seq.py:4,33-35``)."""
    return text.translate(_COMPLEMENT)[::-1]
