"""Genome-wide base composition for the Bayesian stage (reference ``composition.py``).

``write_base_comp`` keeps the reference's signature and output (``dnacomp_genome.csv``: header
``A,C,G,T`` and one row of frequencies, ``composition.py:6-25``); the per-sequence counting the
reference does in its C extension (``seqtk.comp``, ``seqtk/seqtk.c:56-143``) is a population
count over the genome image already resident on the GPU.
"""
import csv

from .engine import DamageEngine
from .refgenome import Reference


def base_counts(fasta, engine=None):
    """``{"A": n, "C": n, "G": n, "T": n}`` over every sequence of ``fasta`` (either case)."""
    own = engine is None
    if own:
        engine = DamageEngine(max_reads=0)
    try:
        if fasta is not None:
            reference = fasta if isinstance(fasta, Reference) else Reference.from_fasta(fasta)
            engine.set_reference(reference)
        return dict(zip("ACGT", engine.genome_composition()))
    finally:
        if own:
            engine.close()


def write_base_comp(fasta, destination, engine=None):
    """``composition.write_base_comp`` (``composition.py:6-25``).  With ``engine`` and ``fasta=None`` the
    genome that engine already holds is used."""
    bases = base_counts(fasta, engine)
    ba_su = sum(bases.values())
    for key in bases:
        bases[key] = bases[key] / ba_su
    with open(destination, "wt", newline="") as handle:
        writer = csv.writer(handle)
        header = ["A", "C", "G", "T"]
        writer.writerow(header)
        writer.writerow(bases[key] for key in header)


def read_base_comp(filename):
    """``composition.read_base_comp`` (``composition.py:28-35``)."""
    with open(filename, newline="") as csvfile:
        reader = csv.DictReader(csvfile)
        for row in reader:
            return row
    raise csv.Error("No rows found in %r" % (filename,))
