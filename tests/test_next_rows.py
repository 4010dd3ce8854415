"""genome composition (f3) and the low-damage table check (f4), against
golden vectors produced by the unmodified reference (oracle/gen_golden_extra.py)."""
import json
import logging

import pytest

from conftest import GOLDEN
from mapdamage_b200 import statistics

CHECKS = json.loads((GOLDEN / "low_damage_check.json").read_text())


@pytest.mark.parametrize("case", sorted(CHECKS["golden_cases"]))
def test_low_damage_check_on_golden_tables(case, caplog):
    want = CHECKS["golden_cases"][case]
    caplog.set_level(logging.DEBUG, logger="mapdamage_b200.statistics")
    assert statistics.check_table_and_warn_if_dmg_freq_is_low(GOLDEN / case) is want["result"]
    got = [[r.levelname, r.getMessage()] for r in caplog.records if r.name == "mapdamage_b200.statistics"]
    assert got == want["log"]


@pytest.mark.parametrize("case", sorted(CHECKS["edited"]))
def test_low_damage_check_on_damaged_tables(case, caplog):
    want = CHECKS["edited"][case]
    folder = GOLDEN / "low_damage_tables" / case
    caplog.set_level(logging.DEBUG, logger="mapdamage_b200.statistics")
    assert statistics.check_table_and_warn_if_dmg_freq_is_low(folder) is want["result"]
    got = [[r.levelname, r.getMessage().replace(str(folder), "<folder>")] for r in caplog.records
           if r.name == "mapdamage_b200.statistics"]
    assert got == want["log"]
    if case == "low_damage":
        assert want["result"] is True and "too low" in want["log"][0][1]


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(p.name for p in (GOLDEN / "genome_composition").iterdir()))
def test_genome_composition_csv(case, tmp_path):
    """dnacomp_genome.csv byte for byte (composition.py:6-25 over seqtk.comp, seqtk.c:56-143)."""
    from mapdamage_b200 import composition

    folder = GOLDEN / "genome_composition" / case
    composition.write_base_comp(folder / "ref.fa", tmp_path / "dnacomp_genome.csv")
    assert (tmp_path / "dnacomp_genome.csv").read_text() == (folder / "dnacomp_genome.csv").read_text()
    row = composition.read_base_comp(tmp_path / "dnacomp_genome.csv")
    assert set(row) == {"A", "C", "G", "T"} and abs(sum(float(v) for v in row.values()) - 1) < 1e-12
