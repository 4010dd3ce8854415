"""TEST INFRASTRUCTURE ONLY -- golden vectors for the "next" rows f3 / f4 of SURVEY.md section 8.

f3: ``composition.write_base_comp`` (reference ``composition.py:6-25``) over ``seqtk.comp``
    (``seqtk/seqtk.c:56-143``, compiled here from the reference's own sources by ``make -C oracle ref``)
    -> ``tests/golden/genome_composition/<case>/{ref.fa, dnacomp_genome.csv}``
f4: ``statistics.check_table_and_warn_if_dmg_freq_is_low`` (``statistics.py:140-184``) run on every
    golden ``misincorporation.txt`` (and a few edited ones) -> ``tests/golden/low_damage_check.json``

Runs only where /root/reference exists.
"""
import importlib.util
import json
import logging
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
GOLDEN = ROOT / "tests" / "golden"
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(ROOT))

import run_reference  # noqa: E402


def load_reference_modules():
    mapdamage = run_reference._import_reference()
    built = sorted((HERE / "_ref").glob("seqtk*.so"))
    if not built:
        sys.exit("build the reference's seqtk extension first: make -C oracle ref")
    spec = importlib.util.spec_from_file_location("seqtk", built[0])
    seqtk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(seqtk)
    sys.modules["mapdamage.seqtk"] = seqtk
    import mapdamage.composition
    import mapdamage.statistics

    # mapdamage.main already imported composition with the (empty) namespace package mapdamage/seqtk/ bound
    mapdamage.seqtk = seqtk
    mapdamage.composition.seqtk = seqtk

    return mapdamage


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__()
        self.messages = []

    def emit(self, record):
        self.messages.append([record.levelname, record.getMessage()])


def composition_cases():
    import numpy as np

    from mapdamage_b200 import synth

    rng = np.random.default_rng(77)
    mixed = "".join(rng.choice(list("ACGTacgtNnRYKM"), p=[.2, .2, .2, .2, .03, .03, .03, .03, .02, .02, .01, .01, .01, .01],
                               size=5000))
    yield "mixed_case_iupac", [("chrA", mixed), ("chrB", "ACGT" * 7 + "acgtn" * 3), ("empty_of_bases", "NNNNNNNN")]
    reference = synth.make_reference([70_001, 1_233], seed=9, other_rate=0.01)
    yield "random_two_contigs", [(n, s.tobytes().decode()) for n, s in zip(reference.names, reference.sequences)]
    yield "kat", [("chr1", "ACGTTGCAACCCGGATATCGTTAGCCGTACGGCATCGATCAATTCCGGATCGCGTATACA")]


def main():
    if not run_reference.available():
        sys.exit("reference tree not found")
    mapdamage = load_reference_modules()
    out_root = GOLDEN / "genome_composition"
    for name, contigs in composition_cases():
        out = out_root / name
        out.mkdir(parents=True, exist_ok=True)
        with open(out / "ref.fa", "wt") as handle:
            for contig, seq in contigs:
                handle.write(">%s\n" % contig)
                for i in range(0, len(seq), 60):
                    handle.write(seq[i:i + 60] + "\n")
        mapdamage.composition.write_base_comp(out / "ref.fa", out / "dnacomp_genome.csv")
        print("composition", name, (out / "dnacomp_genome.csv").read_text().splitlines()[1])

    results = {}
    logger = logging.getLogger("mapdamage.statistics")
    handler = _Capture()
    logger.addHandler(handler)
    logger.setLevel(logging.DEBUG)
    cases = [p.parent for p in sorted(GOLDEN.glob("*/misincorporation.txt"))]
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        for case in cases:
            handler.messages.clear()
            result = mapdamage.statistics.check_table_and_warn_if_dmg_freq_is_low(str(case))
            results[case.name] = {"result": bool(result), "log": list(handler.messages)}
        # edited tables: empty file, missing file, header only, missing column
        (tmp / "empty").mkdir()
        (tmp / "empty" / "misincorporation.txt").write_text("")
        (tmp / "missing").mkdir()
        (tmp / "header_only").mkdir()
        header = (GOLDEN / "kat" / "misincorporation.txt").read_text().splitlines()[0]
        (tmp / "header_only" / "misincorporation.txt").write_text(header + "\n")
        (tmp / "no_column").mkdir()
        lines = (GOLDEN / "c1_se100" / "misincorporation.txt").read_text().splitlines()
        drop = lines[0].split("\t").index("C>T")
        (tmp / "no_column" / "misincorporation.txt").write_text(
            "\n".join("\t".join(f for i, f in enumerate(line.split("\t")) if i != drop) for line in lines) + "\n")
        (tmp / "low_damage").mkdir()
        cols = lines[0].split("\t")
        rows = [lines[0]]
        for line in lines[1:]:
            f = line.split("\t")
            if f[cols.index("Pos")] == "1":
                f[cols.index("C>T")] = f[cols.index("G>A")] = "1"
            rows.append("\t".join(f))
        (tmp / "low_damage" / "misincorporation.txt").write_text("\n".join(rows) + "\n")
        edited = {}
        for name in ("empty", "missing", "header_only", "no_column", "low_damage"):
            handler.messages.clear()
            result = mapdamage.statistics.check_table_and_warn_if_dmg_freq_is_low(str(tmp / name))
            log = [[lvl, msg.replace(str(tmp / name), "<folder>")] for lvl, msg in handler.messages]
            edited[name] = {"result": bool(result), "log": log}
            src = tmp / name / "misincorporation.txt"
            if src.is_file():
                dst = GOLDEN / "low_damage_tables" / name
                dst.mkdir(parents=True, exist_ok=True)
                shutil.copy(src, dst / "misincorporation.txt")
        (GOLDEN / "low_damage_tables" / "missing").mkdir(parents=True, exist_ok=True)
        ((GOLDEN / "low_damage_tables" / "missing") / ".keep").write_text("")
    (GOLDEN / "low_damage_check.json").write_text(
        json.dumps({"golden_cases": results, "edited": edited}, indent=1, sort_keys=True) + "\n")
    print("low-damage check:", {k: v["result"] for k, v in results.items()})
    print("edited:", edited)


if __name__ == "__main__":
    main()
