"""How an input is recognised (``counting.input_kind``): by content for files, like pysam under the reference
(``reader.py:34-38``); ``-`` and pipes are taken for BAM unless named ``*.sam``."""
import os

import bam_py
from conftest import GOLDEN
from mapdamage_b200 import counting
from mapdamage_b200.samtext import read_sam


def test_files_are_recognised_by_their_first_bytes(tmp_path):
    case = GOLDEN / "kat"
    header, records = read_sam(case / "input.sam")
    bam_py.write_bam(tmp_path / "reads.anything", header, records)
    (tmp_path / "reads.bam").write_bytes((case / "input.sam").read_bytes())  # SAM text under a BAM name
    assert counting.input_kind(tmp_path / "reads.anything") == (tmp_path / "reads.anything", True, False)
    assert counting.input_kind(tmp_path / "reads.bam") == (tmp_path / "reads.bam", False, False)
    assert counting.input_kind(str(case / "input.sam"))[1:] == (False, False)


def test_pipes_and_stdin(tmp_path):
    os.mkfifo(tmp_path / "pipe")
    os.mkfifo(tmp_path / "pipe.sam")
    assert counting.input_kind(tmp_path / "pipe")[1:] == (True, True)
    assert counting.input_kind(tmp_path / "pipe.sam")[1:] == (False, True)
    path, is_bam, is_stream = counting.input_kind("-")
    assert str(path) == "/dev/stdin" and is_bam and is_stream


def test_sequence_dictionaries_are_compared_like_the_reference(caplog):
    """``main.py:139-145`` / ``seq.compare_sequence_dicts`` (``seq.py:75-112``): a contig of the alignment file that
    the FASTA lacks, or whose length differs, stops the run with the reference's messages; extra FASTA contigs warn."""
    import logging

    import pytest

    from mapdamage_b200.refgenome import Reference, ReferenceMismatch

    reference = Reference(["chr1", "chr2", "extra"], ["ACGTACGTAC", "GGGGCCCC", "TT"])
    caplog.set_level(logging.WARNING, logger="mapdamage_b200.refgenome")
    ordered = reference.reordered(["chr2", "chr1"], [8, 10])
    assert ordered.names == ["chr2", "chr1"] and ordered.lengths == [8, 10]
    assert any("FASTA file contains extra sequences" in r.getMessage() for r in caplog.records)
    caplog.clear()
    with pytest.raises(ReferenceMismatch):
        reference.reordered(["chr1", "chr2"], [10, 9])  # wrong genome build: silent garbage before
    assert [r.getMessage() for r in caplog.records][:2] == ["Length of required FASTA sequences differ:",
                                                             " - chr2: 8 vs 9 bp"]
    caplog.clear()
    with pytest.raises(ReferenceMismatch):
        reference.reordered(["chr1", "chrM"], [10, 16569])
    assert "Sequences not found in FASTA:" in [r.getMessage() for r in caplog.records]
    with pytest.raises(ReferenceMismatch):
        reference.reordered(["a", "b"], [1, 2])
    with pytest.raises(ReferenceMismatch):
        reference.reordered(["nope"])  # no lengths given: still a proper error, not a KeyError


def test_incomplete_readgroup_is_a_bam_error():
    """``reader.py:107-116``: an @RG line without SM or LB."""
    import pytest

    from mapdamage_b200.batch import BAMError
    from mapdamage_b200.samtext import SamHeader

    header = SamHeader()
    header.add("@RG\tID:rg1\tSM:sample")
    with pytest.raises(BAMError) as info:
        header.libraries()
    assert str(info.value) == ("Incomplete readgroup found: rg1 is missing 'LB'. "
                               "Either fix BAM or use --merge-libraries")
