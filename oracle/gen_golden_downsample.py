"""TEST INFRASTRUCTURE ONLY -- golden tables for the reader's down-sampling (``reader.py:84-96,134-164``).

Runs the unmodified reference (``run_reference.run_counting``) with ``-n X --downsample-seed S`` on the inputs of
existing golden cases and stores the three tables under ``tests/golden/<case>/`` with ``kind = counting_downsample``.
The draws come from CPython's ``random.Random``; the tables are therefore pinned to the interpreter's generator
(MT19937, stable across CPython 3.x for ``random()`` and, since 3.2, for ``randint``).

Runs only where /root/reference exists.
"""
import json
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
GOLDEN = ROOT / "tests" / "golden"
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import run_reference  # noqa: E402

TABLES = ("misincorporation.txt", "dnacomp.txt", "lgdistribution.txt")
CASES = (
    ("d0_fraction_fuzz0", "fuzz_0_l70_a10_q0", 0.4, 5),
    ("d1_fixed_fuzz0", "fuzz_0_l70_a10_q0", 150, 5),
    ("d2_fixed_more_than_there_are", "fuzz_2_l7_a1_q0_merge", 5000, 1),
    ("d3_fraction_c1", "c1_se100", 0.25, 11),
    ("d4_fixed_c3", "c3_pe_mixed", 1234, 3),
    ("d5_fixed_one", "fuzz_1_l25_a4_q20", 1, 9),
)


def main():
    if not run_reference.available():
        sys.exit("reference tree not found")
    from helpers import materialise_inputs

    for name, source, downsample, seed in CASES:
        params = json.loads((GOLDEN / source / "params.json").read_text())
        out = GOLDEN / name
        if out.exists():
            shutil.rmtree(out)
        out.mkdir(parents=True)
        with tempfile.TemporaryDirectory() as tmp:
            tmp = Path(tmp)
            sam, fasta = materialise_inputs(GOLDEN / source, params, tmp)
            shutil.copy(fasta, tmp / "reference.fa")  # the .fai is written next to it
            rc = run_reference.run_counting(sam, tmp / "reference.fa", tmp / "out", length=params["length"],
                                            around=params["around"], minqual=params["minqual"],
                                            merge_libraries=params["merge_libraries"],
                                            extra=["-n", repr(downsample), "--downsample-seed", str(seed)])
            assert rc == 0, (name, rc)
            for table in TABLES:
                shutil.copy(tmp / "out" / table, out / table)
        params.update(kind="counting_downsample", source=source, downsample=downsample, downsample_seed=seed)
        (out / "params.json").write_text(json.dumps(params, indent=1, sort_keys=True) + "\n")
        kept = sum(int(line.split("\t")[-1]) for line in (out / "lgdistribution.txt").read_text().splitlines()
                   if line and line[0] not in "#S" and line.split("\t")[-1].isdigit())
        print("downsample", name, "rc", rc, "fragments counted", kept)


if __name__ == "__main__":
    main()
