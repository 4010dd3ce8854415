"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/ by running the
UNMODIFIED reference (/root/reference) through oracle/pysam_shim.py.

Run here (dev container) only:  python oracle/gen_golden.py
The GPU box has no /root/reference; it only ever reads the committed vectors.

Each case directory holds the inputs (input.sam, ref.fa, params.json and, for
rescale cases, Stats_out_MCMC_correct_prob.csv) and what the reference wrote
(misincorporation.txt / dnacomp.txt / lgdistribution.txt, or expected.sam).
Large synthetic inputs (c1_*) are not stored: params.json records the
generator arguments and the SHA-256 of the regenerated input.sam / ref.fa.
"""
import hashlib
import io
import json
import logging
import math
import shutil
import sys
import tempfile
from contextlib import redirect_stderr
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(ROOT))

import fuzz  # noqa: E402
import run_reference  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
KAT_REF = "ACGTTGCAACCCGGATATCGTTAGCCGTACGGCATCGATCAATTCCGGATCGCGTATACA"
TABLES = ("misincorporation.txt", "dnacomp.txt", "lgdistribution.txt")


def sam_text(records, contigs, readgroups=()):
    lines = ["@HD\tVN:1.6\tSO:unsorted"]
    lines += ["@SQ\tSN:%s\tLN:%d" % (n, len(s)) for n, s in contigs]
    lines += ["@RG\tID:%s\tSM:%s\tLB:%s" % rg for rg in readgroups]
    lines += ["\t".join(str(x) for x in r) for r in records]
    return "\n".join(lines) + "\n"


def rec(name, flag, pos0, cigar, seq, qual=None, rnext="*", pnext0=-1, tlen=0, tags=(), rname="chr1"):
    return [name, flag, rname, pos0 + 1, 60, cigar, rnext, pnext0 + 1, tlen, seq,
            qual if qual is not None else "I" * len(seq), *tags]


def corr_csv_text(seq_length=12):
    """The synthetic damage model of SURVEY Appendix A (format: rescale.py:23-46)."""
    rows = ['"","Position","C.T","G.A"']
    positions = list(range(-seq_length, 0)) + list(range(1, seq_length + 1))
    for i, p in enumerate(positions, 1):
        hi = 0.9 * math.exp(-0.4 * (abs(p) - 1))
        ct, ga = (hi, 0.02) if p > 0 else (0.02, hi)
        rows.append('"%d",%d,%r,%r' % (i, p, ct, ga))
    return "\n".join(rows) + "\n"


def sha256(path):
    return hashlib.sha256(Path(path).read_bytes()).hexdigest()


def counting_case(name, contigs, sam, length, around, minqual, merge, store_inputs=True, extra=None):
    out = GOLDEN / name
    if out.exists():
        shutil.rmtree(out)
    out.mkdir(parents=True)
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        fuzz.write_fasta(contigs, tmp / "ref.fa")
        (tmp / "input.sam").write_text(sam)
        err = io.StringIO()
        exc = None
        with redirect_stderr(err):
            try:
                rc = run_reference.run_counting(tmp / "input.sam", tmp / "ref.fa", tmp / "out",
                                                length=length, around=around, minqual=minqual,
                                                merge_libraries=merge)
            except Exception as error:  # e.g. BAMError
                rc, exc = None, "%s: %s" % (type(error).__name__, error)
        params = dict(kind="counting", length=length, around=around, minqual=minqual,
                      merge_libraries=merge, rc=rc, exception=exc)
        if extra:
            params.update(extra)
        if store_inputs:
            shutil.copy(tmp / "ref.fa", out / "ref.fa")
            shutil.copy(tmp / "input.sam", out / "input.sam")
        else:
            params["sha256"] = {"ref.fa": sha256(tmp / "ref.fa"), "input.sam": sha256(tmp / "input.sam")}
        if rc == 0:
            for table in TABLES:
                shutil.copy(tmp / "out" / table, out / table)
        (out / "params.json").write_text(json.dumps(params, indent=1, sort_keys=True) + "\n")
    print("counting", name, "rc", rc, exc or "")


def rescale_case(name, contigs, sam, length_5p=12, length_3p=12, csv_text=None):
    out = GOLDEN / name
    if out.exists():
        shutil.rmtree(out)
    out.mkdir(parents=True)
    fuzz.write_fasta(contigs, out / "ref.fa")
    (out / "input.sam").write_text(sam)
    (out / "Stats_out_MCMC_correct_prob.csv").write_text(csv_text or corr_csv_text())
    messages = []

    class Grab(logging.Handler):
        def emit(self, record):
            messages.append(record.getMessage())

    handler = Grab(level=logging.INFO)
    logging.getLogger().addHandler(handler)
    level = logging.getLogger().level
    logging.getLogger().setLevel(logging.INFO)
    exc = None
    err = io.StringIO()
    with redirect_stderr(err):
        try:
            rc = run_reference.run_rescale(out / "input.sam", out / "ref.fa", out, out / "expected.sam",
                                           length_5p=length_5p, length_3p=length_3p)
        except SystemExit as error:  # pre-existing MR tag (rescale.py:277-278)
            rc, exc = None, "SystemExit: %s" % (str(error).split("\t")[0],)
    logging.getLogger().removeHandler(handler)
    logging.getLogger().setLevel(level)
    if rc != 0 and (out / "expected.sam").exists():
        (out / "expected.sam").unlink()
    keep = [m for m in messages if not m.startswith(("Rescaling BAM", "Reading corrected"))]
    params = dict(kind="rescale", length_5p=length_5p, length_3p=length_3p, rc=rc, exception=exc,
                  log=keep)
    (out / "params.json").write_text(json.dumps(params, indent=1, sort_keys=True) + "\n")
    print("rescale", name, "rc", rc, exc or "")


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    kat = [("chr1", KAT_REF)]

    # -- SURVEY section 8(c) known-answer vector -----------------------
    records = [
        rec("r1", 0, 0, "10M", "ATGTTGCAAC"),
        rec("r2", 16, 10, "2S4M1I3M2D2M", "NNCCGGAATAGT"),
        rec("r3", 1024, 30, "10M", "GGCATCGATC"),
        rec("r4", 67, 50, "8M2S", "CGCGTATAGG", rnext="=", pnext0=9, tlen=-45),
    ]
    counting_case("kat", kat, sam_text(records, kat), 6, 3, 0, True)

    # -- SURVEY Appendix A, counting rows ------------------------------
    a1 = [rec("a1", 0, 0, "4M1I5M", "ACGTGTGCAA", "I#II#IIIII")]
    counting_case("a01_insertion", kat, sam_text(a1, kat), 10, 2, 0, True)
    counting_case("a02_insertion_q20", kat, sam_text(a1, kat), 10, 2, 20, True)
    a3 = [rec("a3", 0, 0, "3M4N3M", "ACGAAC")]
    counting_case("a03_skip", kat, sam_text(a3, kat), 10, 2, 0, True)
    a4 = [rec("a4", 16, 5, "5M", "GNAAC")]
    counting_case("a04_reverse_n", kat, sam_text(a4, kat), 10, 2, 0, True)
    rgs = [("g2", "samB", "lib1"), ("g1", "samA", "lib2"), ("g0", "samA", "lib1")]
    a5 = [rec("a5", 0, 3, "10M", "TTGCAACCCG", tags=("RG:Z:g2",))]
    counting_case("a05_libraries", kat, sam_text(a5, kat, rgs), 10, 2, 0, False)
    a6 = [rec("a6", 0, 3, "10M", "TTGCAACCCG")]
    counting_case("a06_no_readgroup", kat, sam_text(a6, kat, rgs), 10, 2, 0, False)
    a7 = [rec("a7", 0, 2, "2H1S3=1X2M1S", "TGTTACAG")]
    counting_case("a07_clips_eq_x", kat, sam_text(a7, kat), 10, 2, 0, True)
    # skip + insertion + low quality on both strands: the 5'/3' walks use different
    # reference offsets and the mask lands on the left-aligned reference column
    a18 = [rec("a18f", 0, 1, "3M1I2M5N2M1D3M", "CGTAGCAACCC", "I#I#II#I#II"),
           rec("a18r", 16, 1, "3M1I2M5N2M1D3M", "CGTAGCAACCC", "I#I#II#I#II"),
           rec("a18p", 0, 4, "2M2P1I3N4M", "TGAGGAT", "II#IIII")]
    counting_case("a18_skip_indel_q20", kat, sam_text(a18, kat), 10, 2, 20, True)
    counting_case("a18_skip_indel_q0", kat, sam_text(a18, kat), 10, 2, 0, True)

    # -- fuzz: every CIGAR shape, flag and library layout ---------------
    settings = [(70, 10, 0, False), (25, 4, 20, False), (7, 1, 0, True), (200, 30, 13, True)]
    for k, (length, around, minqual, merge) in enumerate(settings):
        contigs, sam = fuzz.make_case(1000 + k, 700, no_rg_rate=0.1 if merge else 0.0)
        counting_case("fuzz_%d_l%d_a%d_q%d%s" % (k, length, around, minqual, "_merge" if merge else ""),
                      contigs, sam, length, around, minqual, merge)

    # -- configs[0]: 10k synthetic 100 bp SE reads on a 1 Mb reference ---
    from mapdamage_b200 import synth

    for tag, kwargs, minqual in (
        ("c1_se100", dict(length=(100, 100), mix=(1, 0, 0, 0), paired=False), 0),
        ("c1_se100_q20", dict(length=(100, 100), mix=(1, 0, 0, 0), paired=False), 20),
        ("c3_pe_mixed", dict(length=(50, 150), mix=(7, 1, 1, 1), paired=True, n_libs=2), 0),
    ):
        ref_args = dict(lengths=[1_000_000], seed=11)
        read_args = dict(n=10_000, seed=12, **kwargs)
        reference = synth.make_reference(**ref_args)
        batch = synth.simulate_reads(reference, **read_args)
        n_libs = kwargs.get("n_libs", 1)
        rgs = [("rg%d" % i, "sample", "lib%d" % i) for i in range(n_libs)]
        with tempfile.TemporaryDirectory() as tmp:
            synth.write_sam(batch, reference, Path(tmp) / "in.sam", readgroups=rgs,
                            lib_to_rg=[r[0] for r in rgs])
            sam = (Path(tmp) / "in.sam").read_text()
        contigs = [(n, s.tobytes().decode()) for n, s in zip(reference.names, reference.sequences)]
        counting_case(tag, contigs, sam, 70, 10, minqual, False, store_inputs=False,
                      extra=dict(reference=ref_args, reads=read_args, readgroups=rgs))

    # -- rescale: SURVEY section 8(c) KAT and Appendix A8-A17 ------------
    rescale_case("r00_kat", kat, sam_text([rec("k", 0, 0, "10M", "ATGTTGCAAT", "IIII5IIII#")], kat))
    rescale_case("r08_deletion", kat, sam_text([rec("a8", 0, 0, "3M2D5M", "ACGGTAAC")], kat))
    rescale_case("r09_insertion", kat, sam_text([rec("a9", 0, 0, "2M1I5M", "ACTGTTGT")], kat))
    rescale_case("r10_hard_soft", kat, sam_text([rec("a10", 0, 1, "2H1S4M", "ACGTT")], kat))
    rescale_case("r11_trailing_del", kat, sam_text([rec("a11", 0, 0, "4M2D", "ATGT")], kat))
    pairs = [
        rec("a12", 0x1 | 0x20 | 0x40, 0, "10M", "ATGTTGCAAT", rnext="=", pnext0=30, tlen=40),
        rec("a13a", 0x1 | 0x20 | 0x40, 0, "10M", "ATGTTGCAAT", rnext="=", pnext0=0, tlen=10),
        rec("a13b", 0x1 | 0x40, 0, "10M", "ATGTTGCAAT", rnext="=", pnext0=30, tlen=40),
        rec("a12r", 0x1 | 0x10 | 0x80, 30, "10M", "GGCATCAATC", rnext="=", pnext0=0, tlen=-40),
    ]
    rescale_case("r12_pairs", kat, sam_text(pairs, kat))
    rescale_case("r14_duplicate", kat, sam_text([rec("a14", 1024, 0, "10M", "ATGTTGCAAT")], kat))
    rescale_case("r15_no_quals", kat, sam_text([rec("a15", 0, 0, "10M", "ATGTTGCAAT", "*")], kat))
    rescale_case("r16_existing_mr", kat,
                 sam_text([rec("a16", 0, 0, "10M", "ATGTTGCAAT", tags=("MR:f:0.5",))], kat))
    rescale_case("r17_tie", kat, sam_text([rec("a17", 0, 0, "11M", "ACGTTACAACC")], kat))
    rescale_case("r18_short_lengths", kat,
                 sam_text([rec("k", 0, 0, "10M", "ATGTTGCAAT", "IIII5IIII#"),
                           rec("k2", 16, 20, "10M", "TTAGTCGTAT", "IIII5IIII#")], kat),
                 length_5p=3, length_3p=1)

    # -- rescale fuzz: SE + every PE orientation, indels, clips, skips ---
    for k, (l5, l3) in enumerate(((12, 12), (12, 9), (4, 0))):
        rng_seed = 2000 + k
        contigs, sam = fuzz.make_case(rng_seed, 900, readgroups=[], paired_rate=0.5)
        # drop records the reference cannot rescale without failing the whole pass:
        # an outer hard clip followed by a soft clip
        keep = []
        for line in sam.splitlines():
            if not line.startswith("@"):
                cigar = line.split("\t")[5]
                ops = [c for c in cigar if not c.isdigit()]
                if ops and ((ops[0] == "H" and len(ops) > 1 and ops[1] == "S")
                            or (ops[-1] == "H" and len(ops) > 1 and ops[-2] == "S")):
                    continue
            keep.append(line)
        rescale_case("rfuzz_%d_%d_%d" % (k, l5, l3), contigs, "\n".join(keep) + "\n",
                     length_5p=l5, length_3p=l3)


if __name__ == "__main__":
    if not run_reference.available():
        sys.exit("reference tree not found; golden vectors can only be regenerated in the dev container")
    main()
