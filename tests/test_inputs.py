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
