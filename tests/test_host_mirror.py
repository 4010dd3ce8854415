"""The host-side mirrors of the reference's two call sites, end to end on the GPU:
``counting.count_alignments`` (main.py:126-231) and ``rescale.rescale_qual`` (rescale.py:368-383),
against the files and log lines the unmodified reference produced (tests/golden)."""
import argparse
import logging

import numpy as np
import pytest

import oracle
from conftest import golden_cases
from helpers import assert_tables_equal, load_rescale_case, materialise_inputs
from mapdamage_b200 import counting, rescale, synth
from mapdamage_b200.batch import BAMError
from mapdamage_b200.engine import DamageEngine
from mapdamage_b200.rescale_model import RescaleModel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case_dir,params", golden_cases("counting"))
def test_count_alignments_writes_the_reference_tables(case_dir, params, tmp_path):
    sam, fasta = materialise_inputs(case_dir, params, tmp_path)
    kwargs = dict(length=params["length"], around=params["around"], min_basequal=params["minqual"],
                  merge_libraries=params["merge_libraries"], folder=tmp_path / "out", batch_reads=4096)
    if params["exception"]:
        with pytest.raises(BAMError) as info:
            counting.count_alignments(sam, fasta, **kwargs)
        assert str(info.value) == params["exception"].split(": ", 1)[1]
        return
    misincorp, dnacomp, lgdistrib = counting.count_alignments(sam, fasta, **kwargs)
    assert_tables_equal(tmp_path / "out", case_dir)
    # the accumulators keep the reference's nested-dict shape (statistics.py:10-20,59-73,107-115)
    library = misincorp.libraries[0]
    assert set(misincorp.data[library]) == {"5p", "3p"}
    assert set(misincorp.data[library]["5p"]["+"]) >= {"A", "C>T", "G>A", "S", "->A", "T>-"}
    assert sorted(dnacomp.data[library]["5p"]["+"]["A"]) == dnacomp.keys("5p")
    assert set(lgdistrib.data[library]) == {("pe", "+"), ("pe", "-"), ("se", "+"), ("se", "-")}


@pytest.mark.parametrize("case_dir,params", golden_cases("rescale"))
def test_rescale_qual_matches_the_reference(case_dir, params, tmp_path, caplog):
    options = argparse.Namespace(folder=case_dir, filename=case_dir / "input.sam",
                                 rescale_out=tmp_path / "rescaled.sam",
                                 rescale_length_5p=params["length_5p"], rescale_length_3p=params["length_3p"])
    caplog.set_level(logging.INFO, logger="mapdamage_b200.rescale")
    if params["exception"]:
        with pytest.raises(SystemExit) as info:
            rescale.rescale_qual(case_dir / "ref.fa", options)
        assert str(info.value).startswith(params["exception"].split(": ", 1)[1])
        return
    rc = rescale.rescale_qual(case_dir / "ref.fa", options)
    assert rc == params["rc"]
    messages = [r.getMessage() for r in caplog.records if r.name == "mapdamage_b200.rescale"]
    if rc != 0:
        assert any("quality and sequence mismatch" in m for m in messages)
        return
    assert (tmp_path / "rescaled.sam").read_text() == (case_dir / "expected.sam").read_text()
    want = [m for m in params["log"] if not m.startswith("Rescaling BAM")]
    got = [m for m in messages if not m.startswith("Rescaling BAM")]
    assert got == want


@pytest.mark.parametrize("name,kw", [
    ("se100", dict(length=(100, 100))),
    ("pe_mixed", dict(length=(50, 150), mix=(7, 1, 1, 1), paired=True)),
    ("short", dict(length=(20, 45), mix=(6, 1, 1, 2), read_n_rate=0.01)),
])
def test_substitution_summary_vs_oracle(name, kw):
    """Integer part of rescale._record_subs (rescale.py:106-139) from the kernel == the oracle's."""
    reference = synth.make_reference([300_000, 150_000, 4_000], seed=5, other_rate=0.002)
    batch = synth.simulate_reads(reference, 40_000, seed=17, **kw)
    corr = {("C", "T", p): 0.9 * 0.67 ** (p - 1) for p in range(1, 13)}
    corr.update({("G", "A", -p): 0.85 * 0.6 ** (p - 1) for p in range(1, 13)})
    corr.update({("G", "A", p): 0.013 for p in range(1, 13)})
    corr.update({("C", "T", -p): 0.021 for p in range(1, 13)})
    model = RescaleModel(corr, 12, 12)
    _, _, _, subs, rc = oracle.rescale(batch, reference, corr)
    assert rc == 0
    with DamageEngine(max_reads=batch.n, max_cigar_ops=batch.cigar.shape[0], max_bases=batch.total_bases) as engine:
        engine.set_reference(reference)
        engine.set_rescale_model(model)
        engine.rescale(batch)
        engine.sync()
        summary = rescale.SubstitutionSummary(model, *engine.rescale_hist(model.n_slots))
    want = np.array(subs.hist, dtype=np.int64)  # [CT, TC, GA, AG][before, after][130]
    for t, key in enumerate(("CT", "TC", "GA", "AG")):
        assert np.array_equal(summary.data[key + "-before"], want[t, 0]), key
        assert np.array_equal(summary.data[key + "-after"], want[t, 1]), key
    assert [summary.data[b] for b in "ACGT"] == list(subs.ref_count)
    pvals = list(subs.pvals)  # CT, CT_before, TC, GA, GA_before, AG
    for key, value in zip(("CT-pvals", "CT-pvals_before", "TC-pvals", "GA-pvals", "GA-pvals_before", "AG-pvals"),
                          pvals):
        assert summary.data[key] == pytest.approx(value, rel=1e-9), key
    assert want[0, 0].sum() > 100 and want[2, 0].sum() > 100


# ---- the same two call sites fed with BAM (native decode / encode) ----
def _as_bam(sam, path, block_bytes=0xff00):
    import bam_py
    from mapdamage_b200.samtext import read_sam

    header, records = read_sam(sam)
    bam_py.write_bam(path, header, records, block_bytes=block_bytes)
    return header, records


@pytest.fixture(params=["device", "host"])
def bam_path(request, monkeypatch):
    """BAM files go through the GPU decoder / encoder (bamio.DeviceBamStream) unless MDG_BAM_HOST=1."""
    if request.param == "host":
        monkeypatch.setenv("MDG_BAM_HOST", "1")
    monkeypatch.setenv("MDG_BAM_DEVICE_SLAB", "200000")  # several slabs even from small files
    monkeypatch.setenv("MDG_RESCALE_SLAB", "200000")
    return request.param


@pytest.mark.parametrize("case_dir,params", [c for c in golden_cases("counting") if not c.values[1]["exception"]])
def test_count_alignments_from_bam(case_dir, params, tmp_path, bam_path):
    sam, fasta = materialise_inputs(case_dir, params, tmp_path)
    _as_bam(sam, tmp_path / "input.bam", block_bytes=4096)
    counting.count_alignments(tmp_path / "input.bam", fasta, length=params["length"], around=params["around"],
                              min_basequal=params["minqual"], merge_libraries=params["merge_libraries"],
                              folder=tmp_path / "out", batch_reads=1024)
    assert_tables_equal(tmp_path / "out", case_dir)


def test_input_format_is_read_from_the_content(tmp_path, golden_dir):
    """Like pysam under the reference (reader.py:38), not from the file name; a pipe is taken for BAM."""
    import json
    import os
    import threading

    case = golden_dir / "fuzz_0_l70_a10_q0"
    params = json.loads((case / "params.json").read_text())
    _as_bam(case / "input.sam", tmp_path / "alignments.dat", block_bytes=4096)
    (tmp_path / "alignments.txt").write_bytes((case / "input.sam").read_bytes())
    assert counting.input_kind(tmp_path / "alignments.dat")[1:] == (True, False)
    assert counting.input_kind(tmp_path / "alignments.txt")[1:] == (False, False)
    kwargs = dict(length=params["length"], around=params["around"], min_basequal=params["minqual"],
                  merge_libraries=params["merge_libraries"], batch_reads=512)
    for name in ("alignments.dat", "alignments.txt"):
        counting.count_alignments(tmp_path / name, case / "ref.fa", folder=tmp_path / ("out_" + name), **kwargs)
        assert_tables_equal(tmp_path / ("out_" + name), case)
    # through a named pipe
    fifo = tmp_path / "pipe"
    os.mkfifo(fifo)
    writer = threading.Thread(target=lambda: fifo.write_bytes((tmp_path / "alignments.dat").read_bytes()))
    writer.start()
    try:
        assert counting.input_kind(fifo)[1:] == (True, True)
        counting.count_alignments(fifo, case / "ref.fa", folder=tmp_path / "out_pipe", **kwargs)
    finally:
        writer.join()
    assert_tables_equal(tmp_path / "out_pipe", case)
    with pytest.raises(ValueError):
        counting.count_alignments(fifo, case / "ref.fa", downsample=10, **kwargs)


@pytest.mark.parametrize("as_bam", [False, True], ids=["sam", "bam"])
@pytest.mark.parametrize("case_dir,params", golden_cases("counting_downsample"))
def test_count_alignments_downsampled(case_dir, params, as_bam, tmp_path):
    """``-n X --downsample-seed S`` (reader.py:84-96): the reads the reference drew, the tables it wrote."""
    sam, fasta = materialise_inputs(case_dir.parent / params["source"], params, tmp_path)
    if as_bam:
        _as_bam(sam, tmp_path / "input.bam", block_bytes=4096)
        sam = tmp_path / "input.bam"
    counting.count_alignments(sam, fasta, length=params["length"], around=params["around"],
                              min_basequal=params["minqual"], merge_libraries=params["merge_libraries"],
                              folder=tmp_path / "out", batch_reads=256,
                              downsample=params["downsample"], downsample_seed=params["downsample_seed"])
    assert_tables_equal(tmp_path / "out", case_dir)


@pytest.mark.parametrize("case_dir,params", [c for c in golden_cases("rescale")])
def test_rescale_qual_bam_to_bam(case_dir, params, tmp_path, caplog, bam_path):
    import bam_py
    from mapdamage_b200.samtext import read_sam

    _as_bam(case_dir / "input.sam", tmp_path / "input.bam", block_bytes=2048)
    options = argparse.Namespace(folder=case_dir, filename=tmp_path / "input.bam", rescale_out=tmp_path / "rescaled.bam",
                                 rescale_length_5p=params["length_5p"], rescale_length_3p=params["length_3p"])
    caplog.set_level(logging.INFO, logger="mapdamage_b200.rescale")
    if params["exception"]:
        with pytest.raises(SystemExit) as info:
            rescale.rescale_qual(case_dir / "ref.fa", options)
        assert str(info.value).startswith(params["exception"].split(": ", 1)[1])
        return
    rc = rescale.rescale_qual(case_dir / "ref.fa", options)
    assert rc == params["rc"]
    if rc != 0:
        return
    _, want = read_sam(case_dir / "expected.sam")
    text, refs, got = bam_py.read_bam(tmp_path / "rescaled.bam")
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert (g["qname"], g["flag"], g["pos"], g["cigar"], g["seq"], g["qual"]) == \
               (w.qname, w.flag, w.pos, w.cigar, w.seq, w.qual)
        if "MR" in w.tags:
            assert g["tags"]["MR"] == ("f", float(np.float32(w.tags["MR"])))
        else:
            assert "MR" not in g["tags"]
    messages = [r.getMessage() for r in caplog.records if r.name == "mapdamage_b200.rescale"]
    assert [m for m in messages if not m.startswith("Rescaling BAM")] == \
           [m for m in params["log"] if not m.startswith("Rescaling BAM")]
