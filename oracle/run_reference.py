"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference through the shim.

Works only where ``/root/reference`` exists (this dev container); the GPU box
has no reference tree, so nothing that runs there may import this module.
Used by ``oracle/gen_golden.py`` (which writes ``tests/golden/``) and by the
live-reference tests that skip when the tree is absent.

Recipe follows SURVEY.md Appendix C: stub ``pysam``/``coloredlogs`` in
``sys.modules``, put the reference on ``sys.path``, call
``mapdamage.main.main([...])`` (reference ``main.py:49``) for the counting pass
and ``mapdamage.rescale.rescale_qual`` (``rescale.py:368``) for rescaling.
"""
import argparse
import logging
import os
import sys
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("MAPDAMAGE_REFERENCE", "/root/reference"))

_here = Path(__file__).resolve().parent
if str(_here) not in sys.path:
    sys.path.insert(0, str(_here))


def available():
    return (REFERENCE_ROOT / "mapdamage" / "main.py").is_file()


def _import_reference():
    import pysam_shim

    pysam_shim.install()
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import mapdamage.main  # noqa: F401
        import mapdamage.rescale  # noqa: F401
    import mapdamage

    assert Path(mapdamage.__file__).resolve().is_relative_to(REFERENCE_ROOT.resolve())
    return mapdamage


def write_fai(fasta_path):
    """Writes ``<fasta>.fai`` with the 5 tab-separated fields ``seq.py:45`` wants."""
    fasta_path = Path(fasta_path)
    lines = []
    with open(fasta_path, "rb") as handle:
        name, length, offset, linebases, linewidth = None, 0, 0, 0, 0
        pos = 0
        for raw in handle:
            if raw.startswith(b">"):
                if name is not None:
                    lines.append((name, length, offset, linebases, linewidth))
                name = raw[1:].split()[0].decode()
                length, linebases, linewidth = 0, 0, 0
                offset = pos + len(raw)
            else:
                stripped = raw.rstrip(b"\r\n")
                if linebases == 0:
                    linebases, linewidth = len(stripped), len(raw)
                length += len(stripped)
            pos += len(raw)
        if name is not None:
            lines.append((name, length, offset, linebases, linewidth))
    with open(str(fasta_path) + ".fai", "wt") as handle:
        for item in lines:
            handle.write("%s\t%d\t%d\t%d\t%d\n" % item)


def _drop_file_handlers():
    root = logging.getLogger()
    for handler in list(root.handlers):
        if isinstance(handler, logging.FileHandler):
            root.removeHandler(handler)
            handler.close()


def run_counting(sam, fasta, folder, length=70, around=10, minqual=0, merge_libraries=False,
                 extra=()):
    """Runs the reference counting pass; returns its return code.

    ``-m/-b`` are clamped to satisfy ``config.py:417-420``.
    """
    mapdamage = _import_reference()
    if not Path(str(fasta) + ".fai").is_file():
        write_fai(fasta)
    argv = [
        "-i", str(sam), "-r", str(fasta), "-d", str(folder),
        "-l", str(length), "-a", str(around), "-Q", str(minqual),
        "-m", str(min(25, length)), "-b", str(min(10, around)),
        "--no-stats",
    ]
    if merge_libraries:
        argv.append("--merge-libraries")
    argv.extend(extra)
    level = logging.getLogger().level
    try:
        return mapdamage.main.main(argv)
    finally:
        _drop_file_handlers()
        logging.getLogger().setLevel(level)


def run_rescale(sam, fasta, folder, out, length_5p=12, length_3p=12):
    """Runs the reference rescale pass (``rescale.py:368``); returns its rc.

    ``folder`` must hold ``Stats_out_MCMC_correct_prob.csv``.  May raise
    ``SystemExit`` (pre-existing MR tag, ``rescale.py:277-278``).
    """
    mapdamage = _import_reference()
    import pysam

    ref = pysam.FastaFile(str(fasta))
    options = argparse.Namespace(
        folder=Path(folder), filename=Path(sam), rescale_out=Path(out),
        rescale_length_5p=length_5p, rescale_length_3p=length_3p,
    )
    return mapdamage.rescale.rescale_qual(ref, options)
