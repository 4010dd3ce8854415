"""Shared test plumbing: golden-case loading and table rendering."""
import hashlib
import json
from pathlib import Path

import numpy as np

from mapdamage_b200 import inputs, statistics, synth
from mapdamage_b200.batch import BatchBuilder
from mapdamage_b200.refgenome import Reference
from mapdamage_b200.samtext import read_sam

TABLES = ("misincorporation.txt", "dnacomp.txt", "lgdistribution.txt")


def materialise_inputs(case_dir, params, tmp_path):
    """Returns (sam_path, fasta_path); regenerates seeded inputs that are not stored."""
    case_dir = Path(case_dir)
    if (case_dir / "input.sam").is_file():
        return case_dir / "input.sam", case_dir / "ref.fa"
    reference = synth.make_reference(**params["reference"])
    reads = dict(params["reads"])
    reads["length"] = tuple(reads["length"])
    reads["mix"] = tuple(reads["mix"])
    batch = synth.simulate_reads(reference, **reads)
    rgs = [tuple(rg) for rg in params["readgroups"]]
    sam, fasta = tmp_path / "input.sam", tmp_path / "ref.fa"
    synth.write_sam(batch, reference, sam, readgroups=rgs, lib_to_rg=[rg[0] for rg in rgs])
    reference.write_fasta(fasta)
    for path in (sam, fasta):
        digest = hashlib.sha256(path.read_bytes()).hexdigest()
        assert digest == params["sha256"][path.name], (
            "regenerated %s differs from the input the golden tables were made from "
            "(numpy RNG stream drift?)" % path.name)
    return sam, fasta


def load_counting_case(case_dir, params, tmp_path):
    sam, fasta = materialise_inputs(case_dir, params, tmp_path)
    return inputs.load_alignments(sam, fasta, merge_libraries=params["merge_libraries"])


def render_tables(out_dir, libraries, length, around, mis, comp, lg):
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    statistics.MisincorporationRates(libraries, length).load(mis).write(out_dir / TABLES[0])
    statistics.DNAComposition(libraries, around, length).load(comp).write(out_dir / TABLES[1])
    statistics.FragmentLengths(libraries).load(lg).write(out_dir / TABLES[2])


def assert_tables_equal(out_dir, case_dir):
    for table in TABLES:
        got = (Path(out_dir) / table).read_text()
        want = (Path(case_dir) / table).read_text()
        if got != want:
            g, w = got.splitlines(), want.splitlines()
            for i, (a, b) in enumerate(zip(g, w)):
                assert a == b, "%s line %d:\n got  %s\n want %s" % (table, i + 1, a, b)
            assert len(g) == len(w), "%s: %d lines, expected %d" % (table, len(g), len(w))


def load_rescale_case(case_dir):
    """(batch, reference, header, records) with every record kept (rescale.py:300)."""
    case_dir = Path(case_dir)
    header, records = read_sam(case_dir / "input.sam")
    reference = Reference.from_fasta(case_dir / "ref.fa").reordered(header.references)
    builder = BatchBuilder(merge_libraries=True, apply_filter=False)
    for record in records:
        builder.add(record)
    return builder.finish(with_qual=True), reference, header, records


def expected_rescale(case_dir):
    """[(qual string or None, MR float32 or None)] from the reference's output."""
    _, records = read_sam(Path(case_dir) / "expected.sam")
    out = []
    for r in records:
        mr = np.float32(r.tags["MR"]) if "MR" in r.tags else None
        out.append((r.qual, mr))
    return out
