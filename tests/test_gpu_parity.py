"""GPU parity: the CUDA path, through the C ABI, against the golden vectors of the
unmodified reference and against the CPU oracle on seeded synthetic batches."""
import numpy as np
import pytest

import oracle
from conftest import golden_cases
from helpers import (assert_tables_equal, expected_rescale, load_counting_case,
                     load_rescale_case, render_tables)
from mapdamage_b200 import _native, rescale, synth
from mapdamage_b200.engine import DamageEngine
from mapdamage_b200.rescale_model import RescaleModel, get_corr_prob

pytestmark = pytest.mark.gpu


def run_engine(batch, reference, n_lib=1, length=70, around=10, min_qual=0, lg_bins=8192, chunks=1,
               resident=False):
    with DamageEngine(length=length, around=around, min_qual=min_qual, n_libraries=n_lib,
                      lg_bins=lg_bins, max_reads=max(1024, batch.n),
                      max_cigar_ops=max(4096, batch.cigar.shape[0]),
                      max_bases=max(1 << 16, batch.total_bases)) as engine:
        engine.set_reference(reference)
        if resident:
            dev = engine.upload(batch)
            engine.count_resident(dev)
            engine.sync()
            dev.free()
        else:
            for part in batch.split(chunks):
                engine.count(part)
        mis, comp, lg = engine.tables()
        overflow = engine.lg_overflow()
        assert engine.launch_count() > 0
    return mis, comp, lg, overflow


@pytest.mark.parametrize("case_dir,params", [c for c in golden_cases("counting")
                                              if not c.values[1]["exception"]])
def test_counting_golden(case_dir, params, tmp_path):
    batch, reference, libraries, _ = load_counting_case(case_dir, params, tmp_path)
    L, A = params["length"], params["around"]
    mis, comp, lg, overflow = run_engine(batch, reference, n_lib=len(libraries), length=L, around=A,
                                         min_qual=params["minqual"], chunks=3)
    assert not overflow
    render_tables(tmp_path / "out", libraries, L, A, mis, comp, lg)
    assert_tables_equal(tmp_path / "out", case_dir)


SYNTH = {
    "se100": dict(length=(100, 100), mix=(1, 0, 0, 0), paired=False),
    "se100_noqual": dict(length=(100, 100), mix=(1, 0, 0, 0), paired=False, with_qual=False),
    "pe_mixed": dict(length=(50, 150), mix=(7, 1, 1, 1), paired=True),
    "short": dict(length=(20, 45), mix=(6, 1, 1, 2), paired=False, read_n_rate=0.01, filtered_rate=0.05),
    "long": dict(length=(180, 400), mix=(5, 2, 2, 1), paired=True, read_n_rate=0.002),
}


@pytest.mark.parametrize("name", sorted(SYNTH))
@pytest.mark.parametrize("min_qual,n_lib,resident", [(0, 1, False), (20, 1, True), (0, 3, False), (13, 2, True)])
def test_counting_vs_oracle(name, min_qual, n_lib, resident):
    reference = synth.make_reference([300_000, 150_000, 4_000], seed=5, other_rate=0.002)
    batch = synth.simulate_reads(reference, 60_000, seed=6, n_libs=n_lib, **SYNTH[name])
    want = oracle.count(batch, reference, minqual=min_qual, n_lib=n_lib, lg_bins=8192, threads=4)
    got = run_engine(batch, reference, n_lib=n_lib, min_qual=min_qual, chunks=2, resident=resident)
    for key, a, b in zip(("misincorp", "dnacomp", "lghist"), got, want):
        assert np.array_equal(a, b), key
    assert want[0].sum() > 0 and want[1].sum() > 0


@pytest.mark.parametrize("length,around", [(1, 0), (5, 40), (200, 3), (900, 10)])
def test_counting_table_shapes(length, around):
    """Extreme --length / --around, including a slab too large for shared memory."""
    reference = synth.make_reference([50_000], seed=8)
    batch = synth.simulate_reads(reference, 20_000, seed=9, length=(30, 120), mix=(4, 1, 1, 1))
    want = oracle.count(batch, reference, length=length, around=around, lg_bins=8192, threads=4)
    got = run_engine(batch, reference, length=length, around=around)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("length,around", [(50, 10), (40, 24), (86, 10), (70, 26), (22, 10)])
@pytest.mark.parametrize("n_lib,min_qual", [(1, 0), (2, 0), (1, 18)])
def test_window_geometries_of_the_planes_kernels(length, around, n_lib, min_qual, monkeypatch):
    """--length / --around at the edges of what the warp-specialised kernel is compiled for (two or three 32-position words
    per anchor window; its largest shared-memory footprint at L + A = 96), one word per window (the one-role kernel),
    with two libraries, with -Q, with one-indel reads staged by the planes kernel."""
    monkeypatch.setenv("MDG_PLANES_INDELS", "1")
    reference = synth.make_reference([80_000, 3_000], seed=21, other_rate=0.001)
    batch = synth.simulate_reads(reference, 30_000, seed=22, length=(25, 130), mix=(5, 2, 2, 1), n_libs=n_lib, read_n_rate=0.01)
    want = oracle.count(batch, reference, length=length, around=around, minqual=min_qual, n_lib=n_lib, lg_bins=8192, threads=4)
    got = run_engine(batch, reference, n_lib=n_lib, length=length, around=around, min_qual=min_qual, resident=True)
    for key, a, b in zip(("misincorp", "dnacomp", "lghist"), got, want):
        assert np.array_equal(a, b), key


def test_fragment_length_overflow():
    reference = synth.make_reference([50_000], seed=8)
    batch = synth.simulate_reads(reference, 5_000, seed=9, length=(30, 120), paired=True)
    want = oracle.count(batch, reference, lg_bins=8192)[2]
    mis, comp, lg, overflow = run_engine(batch, reference, lg_bins=64)
    assert np.array_equal(lg, want[..., :64])
    dense = np.zeros_like(want)
    for lib, kind, strand, length, count in overflow:
        dense[lib, kind, strand, length] += count
    assert np.array_equal(dense[..., 64:], want[..., 64:]) and dense[..., :64].sum() == 0


def test_empty_and_reset():
    reference = synth.make_reference([10_000], seed=8)
    batch = synth.simulate_reads(reference, 1_000, seed=9)
    with DamageEngine(max_reads=2048) as engine:
        engine.set_reference(reference)
        engine.count(batch.slice(0, 0))
        assert all(t.sum() == 0 for t in engine.tables())
        engine.count(batch)
        first = engine.tables()
        engine.count(batch)
        twice = engine.tables()
        assert all(np.array_equal(2 * a, b) for a, b in zip(first, twice))
        engine.reset()
        assert all(t.sum() == 0 for t in engine.tables())


def test_errors_are_loud():
    reference = synth.make_reference([10_000], seed=8)
    batch = synth.simulate_reads(reference, 1_000, seed=9)
    with DamageEngine(max_reads=100) as engine:
        with pytest.raises(_native.NativeError) as info:
            engine.count(batch)
        assert info.value.code == _native.ERR_STATE  # no reference yet
        engine.set_reference(reference)
        with pytest.raises(_native.NativeError) as info:
            engine.count(batch)
        assert info.value.code == _native.ERR_CAPACITY
    with DamageEngine(max_reads=2048) as engine:
        engine.set_reference(reference)
        batch.lib[5] = 3
        batch.invalidate()
        engine.count(batch)
        with pytest.raises(_native.NativeError) as info:
            engine.sync()
        assert info.value.code == _native.ERR_DATA


@pytest.mark.parametrize("case_dir,params", [c for c in golden_cases("rescale")
                                              if not c.values[1]["exception"]])
def test_rescale_golden(case_dir, params):
    batch, reference, _, records = load_rescale_case(case_dir)
    model = RescaleModel.from_csv(case_dir / "Stats_out_MCMC_correct_prob.csv",
                                  params["length_5p"], params["length_3p"])
    with DamageEngine(max_reads=max(1024, batch.n), max_cigar_ops=max(4096, batch.cigar.shape[0])) as engine:
        engine.set_reference(reference)
        engine.set_rescale_model(model)
        qual, mr, status = engine.rescale(batch)
        if params["rc"] != 0:
            with pytest.raises(_native.NativeError) as info:
                engine.sync()
            assert info.value.code == _native.ERR_DATA
            assert "quality and sequence mismatch" in info.value.message
            return
        engine.sync()
        stats = engine.rescale_stats()
    want = expected_rescale(case_dir)
    for i, (want_qual, want_mr) in enumerate(want):
        off, n = int(batch.base_off[i]), int(batch.l_seq[i])
        if want_qual is not None:
            got = (qual[off:off + n] + 33).astype(np.uint8).tobytes().decode("latin-1")
            assert got == want_qual, "record %d (%s)" % (i, records[i].qname)
        if want_mr is None:
            assert status[i] == 0
        else:
            assert status[i] == 1 and mr[i] == want_mr, "record %d MR %r != %r" % (i, mr[i], want_mr)
    n_warn = sum("longer than the actual read" in m for m in params["log"])
    assert stats["alignment_longer_than_read"] == n_warn


@pytest.mark.parametrize("name", ["se100", "pe_mixed", "short", "long"])
@pytest.mark.parametrize("l5,l3", [(12, 12), (30, 7)])
def test_rescale_vs_oracle(name, l5, l3):
    reference = synth.make_reference([300_000, 150_000, 4_000], seed=5, other_rate=0.002)
    batch = synth.simulate_reads(reference, 50_000, seed=7, **SYNTH[name])
    corr = {("C", "T", p): 0.9 * 0.67 ** (p - 1) for p in range(1, l5 + 1)}
    corr.update({("G", "A", -p): 0.85 * 0.6 ** (p - 1) for p in range(1, l3 + 1)})
    corr.update({("G", "A", p): 0.013 for p in range(1, l5 + 1)})
    corr.update({("C", "T", -p): 0.021 for p in range(1, l3 + 1)})
    model = RescaleModel(corr, l5, l3)
    want_qual, want_mr, want_status, subs, rc = oracle.rescale(batch, reference, corr)
    assert rc == 0
    with DamageEngine(max_reads=batch.n, max_cigar_ops=batch.cigar.shape[0], max_bases=batch.total_bases) as engine:
        engine.set_reference(reference)
        engine.set_rescale_model(model)
        qual, mr, status = engine.rescale(batch)
        engine.sync()
        stats = engine.rescale_stats()
    assert np.array_equal(status, want_status)
    assert np.array_equal(mr[status == 1], want_mr[want_status == 1])
    # compare the quality bytes of real bases only (pad slots are unspecified)
    mask = np.zeros(qual.shape[0], dtype=bool)
    starts, lens = batch.base_off.astype(np.int64), batch.l_seq.astype(np.int64)
    idx = np.repeat(starts, lens) + (np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens))
    mask[idx] = True
    assert np.array_equal(qual[mask], want_qual[:qual.shape[0]][mask])
    assert (qual[mask] != batch.qual[:qual.shape[0]][mask]).sum() > 100
    assert stats["rescaled"] == int(want_status.sum()) == subs.n_rescaled
    assert stats["pairs"] == subs.n_pairs and stats["improper_pairs"] == subs.n_improper


def test_rescale_long_model_uses_global_histograms():
    """A correction model too long for the block's shared-memory histograms (rescale lengths 45 + 40): same numbers."""
    l5, l3 = 45, 40
    reference = synth.make_reference([200_000], seed=15)
    batch = synth.simulate_reads(reference, 20_000, seed=17, **SYNTH["pe_mixed"])
    corr = {("C", "T", p): 0.8 * 0.9 ** (p - 1) for p in range(1, l5 + 1)}
    corr.update({("G", "A", -p): 0.7 * 0.9 ** (p - 1) for p in range(1, l3 + 1)})
    corr.update({("G", "A", p): 0.01 for p in range(1, l5 + 1)})
    corr.update({("C", "T", -p): 0.02 for p in range(1, l3 + 1)})
    model = RescaleModel(corr, l5, l3)
    want_qual, want_mr, want_status, subs, rc = oracle.rescale(batch, reference, corr)
    assert rc == 0
    with DamageEngine(max_reads=batch.n, max_cigar_ops=batch.cigar.shape[0], max_bases=batch.total_bases) as engine:
        engine.set_reference(reference)
        engine.set_rescale_model(model)
        qual, mr, status = engine.rescale(batch)
        engine.sync()
        summary = rescale.SubstitutionSummary(model, *engine.rescale_hist(model.n_slots))
    assert np.array_equal(status, want_status)
    assert np.array_equal(mr[status == 1], want_mr[want_status == 1])
    starts, lens = batch.base_off.astype(np.int64), batch.l_seq.astype(np.int64)
    idx = np.repeat(starts, lens) + (np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens))
    assert np.array_equal(qual[idx], want_qual[idx])
    want = np.array(subs.hist, dtype=np.int64)  # [CT, TC, GA, AG][before, after][130]
    for t, key in enumerate(("CT", "TC", "GA", "AG")):
        assert np.array_equal(summary.data[key + "-before"], want[t, 0]), key
        assert np.array_equal(summary.data[key + "-after"], want[t, 1]), key
    assert [summary.data[b] for b in "ACGT"] == list(subs.ref_count)
    assert want[0, 0].sum() > 100


DEVICE_SYNTH = {
    "se100": dict(length=(100, 100)),
    "pe_mixed": dict(length=(50, 150), mix=(7, 1, 1, 1), paired=True),
    "short_libs": dict(length=(20, 45), mix=(6, 1, 1, 2), read_n_rate=0.01, filtered_rate=0.05, n_libs=3),
    "noqual": dict(length=(100, 100), with_qual=False),
}


@pytest.mark.parametrize("name", sorted(DEVICE_SYNTH))
@pytest.mark.parametrize("min_qual", [0, 17])
def test_device_generated_batches(name, min_qual):
    """Batches generated in HBM (mdg_synth_batch): well-formed, damaged, and counted like the oracle counts them."""
    kw = dict(DEVICE_SYNTH[name])
    n_lib = kw.get("n_libs", 1)
    reference = synth.make_reference([300_000, 150_000, 4_000], seed=5, other_rate=0.002)
    with DamageEngine(min_qual=min_qual, n_libraries=n_lib, max_reads=1024) as engine:
        engine.set_reference(reference)
        dev = engine.synth_batch(80_000, seed=11, **kw)
        engine.count_resident(dev)
        got = engine.tables()
        host = engine.download(dev)
        again = engine.synth_batch(80_000, seed=11, **kw)
        host2 = engine.download(again)
        dev.free()
        again.free()
    host.validate()
    assert np.array_equal(host.seq4, host2.seq4) and np.array_equal(host.pos, host2.pos)  # deterministic in the seed
    want = oracle.count(host, reference, minqual=min_qual, n_lib=n_lib, lg_bins=8192, threads=4)
    for key, a, b in zip(("misincorp", "dnacomp", "lghist"), got, want):
        assert np.array_equal(a, b), key
    mis = want[0].sum(axis=(0, 2))  # [end][class][pos]
    c_to_t = mis[0, 4 + 5 * 1 + 3, 0] / max(1, mis[0, 1, 0])
    g_to_a = mis[1, 4 + 5 * 2 + 0, 0] / max(1, mis[1, 2, 0])
    assert 0.2 < c_to_t < 0.4 and 0.2 < g_to_a < 0.4  # the injected 5' C>T / 3' G>A damage is there
    if kw.get("paired"):
        assert want[2][:, 0].sum() > 0 and np.all(host.flag & 1)


@pytest.mark.parametrize("name", ["se100_noqual", "pe_mixed", "short"])
def test_optional_arrays_default_on_device(name):
    """NULL lib / tlen / base_off / cigar_off (mdg_batch optional arrays) count like the explicit arrays."""
    reference = synth.make_reference([300_000, 150_000, 4_000], seed=5)
    batch = synth.simulate_reads(reference, 30_000, seed=21, **SYNTH[name])
    assert {"base_off", "lib", "l_seq"} <= batch.droppable()
    if name == "se100_noqual":
        assert {"cigar_off", "tlen", "cigar"} <= batch.droppable()  # "cigar": one word ("100M") shared by every read
    tables = []
    for compact in (False, True):
        with DamageEngine(max_reads=batch.n, max_cigar_ops=batch.cigar.shape[0], max_bases=batch.total_bases) as engine:
            engine.set_reference(reference)
            assert engine.h2d_bytes(batch, compact=True) < engine.h2d_bytes(batch, compact=False)
            engine.count(batch, compact=compact)
            tables.append(engine.tables())
    want = oracle.count(batch, reference, lg_bins=8192, threads=4)
    for a, b, c in zip(tables[0], tables[1], want):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    # a batch that lies about its size is refused, not read out of bounds
    import ctypes as C
    from mapdamage_b200.engine import batch_struct
    with DamageEngine(max_reads=batch.n, max_cigar_ops=batch.cigar.shape[0], max_bases=batch.total_bases) as engine:
        engine.set_reference(reference)
        s = batch_struct(batch, compact=True)
        s.n_bases = (s.n_bases // 4) & ~1
        engine._check(engine._lib.mdg_count_submit(engine._ctx, C.byref(s)))
        with pytest.raises(_native.NativeError) as info:
            engine.sync()
        assert info.value.code == _native.ERR_DATA


@pytest.mark.parametrize("env", [{}, {"MDG_SWAR_FLUSH_TILES": "3"}, {"MDG_SWAR_UNIFORM": "0"}, {"MDG_STAGE_THREADS": "256"},
                                 {"MDG_STAGE_TILE": "96"}, {"MDG_STAGE_INDELS": "1"}, {"MDG_STAGE_INDELS": "0"}, {"MDG_KERNEL": "swar"},
                                 {"MDG_KERNEL": "staged"}, {"MDG_KERNEL": "staged", "MDG_SWAR_FLUSH_TILES": "3"},
                                 {"MDG_PLANES_TILE": "224", "MDG_SWAR_FLUSH_TILES": "2"}, {"MDG_PLANES_SLAB": "0"},
                                 {"MDG_KERNEL": "swar", "MDG_SWAR_VARIANT": "512,1", "MDG_SWAR_FLUSH_TILES": "2"},
                                 {"MDG_PLANES_WS": "0"}, {"MDG_PLANES_WS": "2x8+8"}, {"MDG_PLANES_WS": "2x8+8", "MDG_SWAR_FLUSH_TILES": "3"},
                                 {"MDG_PLANES_WS": "2x8+8", "MDG_PLANES_SLAB": "0"}, {"MDG_PLANES_WS": "2x8+8", "MDG_SWAR_UNIFORM": "0"},
                                 {"MDG_PLANES_WS": "2x9+8", "MDG_SWAR_FLUSH_TILES": "5"},
                                 {"MDG_PLANES_QUAL": "0"}, {"MDG_PLANES_THREADS": "256"},
                                 {"MDG_PLANES_GATHER": "1", "MDG_PLANES_PREFETCH": "3"}, {"MDG_PLANES_GATHER": "1", "MDG_PLANES_WS": "2x8+8"},
                                 {"MDG_PLANES_INDELS": "1"}, {"MDG_PLANES_INDELS": "1", "MDG_PLANES_GATHER": "1", "MDG_SWAR_FLUSH_TILES": "3"},
                                 {"MDG_PLANES_INDELS": "1", "MDG_PLANES_WS": "2x8+8", "MDG_PLANES_SLAB": "0"}, {"MDG_PLANES_INDELS": "0"}])
@pytest.mark.parametrize("min_qual", [0, 25])
def test_mode_switches_and_flushes_at_scale(env, min_qual, monkeypatch):
    """1.3 M reads laid out so that every block of the bit-sliced kernel alternates between equal-length tiles
    (two different lengths) and mixed tiles: exercises the counter flush on mode changes, the periodic flush,
    and the other compiled variants of the bit-sliced kernels (staged: the default; one-phase: MDG_KERNEL=swar)."""
    from mapdamage_b200.batch import concatenate

    for key, value in env.items():
        monkeypatch.setenv(key, value)
    reference = synth.make_reference([400_000, 90_000], seed=5, other_rate=0.001)
    with DamageEngine(min_qual=min_qual, max_reads=1024) as engine:
        engine.set_reference(reference)
        parts = []
        for k, kw in enumerate((dict(length=(100, 100)), dict(length=(35, 140), mix=(6, 1, 1, 2), read_n_rate=0.01),
                                dict(length=(83, 83)), dict(length=(100, 100), filtered_rate=0.2))):
            dev = engine.synth_batch(330_000, seed=50 + k, **kw)
            parts.append(engine.download(dev))
            dev.free()
    batch = concatenate(parts)
    want = oracle.count(batch, reference, minqual=min_qual, lg_bins=8192, threads=8)
    with DamageEngine(min_qual=min_qual, max_reads=batch.n, max_cigar_ops=batch.cigar.shape[0],
                      max_bases=batch.total_bases) as engine:
        engine.set_reference(reference)
        engine.count(batch)
        got = engine.tables()
        dev = engine.upload(batch)
        engine.count_resident(dev)
        twice = engine.tables()
        dev.free()
    for name, a, b, c in zip(("misincorp", "dnacomp", "lghist"), got, want, twice):
        assert np.array_equal(a, b), name
        assert np.array_equal(2 * a, c), name


@pytest.mark.parametrize("n_libs", [1, 2])
@pytest.mark.parametrize("min_qual", [0, 15])
def test_batch_sizes_around_the_tile_size(n_libs, min_qual, monkeypatch):
    """Batches of 1 read up to a few tiles, on and just off the tile sizes of the warp-specialised kernel (288 reads, 256
    with two libraries): a team without a tile, a last tile with one read, bulk copies of a few bytes."""
    monkeypatch.setenv("MDG_PLANES_INDELS", "1")
    reference = synth.make_reference([60_000, 2_000], seed=12, other_rate=0.001)
    sizes = (1, 2, 31, 255, 256, 257, 287, 288, 289, 511, 512, 513, 575, 576, 577, 1153)
    whole = synth.simulate_reads(reference, sum(sizes), seed=13, length=(30, 120), mix=(6, 1, 1, 2), n_libs=n_libs, read_n_rate=0.01)
    with DamageEngine(n_libraries=n_libs, min_qual=min_qual, max_reads=0) as engine:
        engine.set_reference(reference)
        at = 0
        for n in sizes:
            part = whole.slice(at, at + n)
            at += n
            want = oracle.count(part, reference, minqual=min_qual, n_lib=n_libs, lg_bins=8192, threads=1)
            engine.reset()
            dev = engine.upload(part)
            engine.count_resident(dev)
            got = engine.tables()
            dev.free()
            for name, a, b in zip(("misincorp", "dnacomp", "lghist"), got, want):
                assert np.array_equal(a, b), (n, name)


@pytest.mark.parametrize("env", [{}, {"MDG_PLANES_WS_LIBS": "0"}, {"MDG_PLANES_WS": "0"}, {"MDG_SWAR_FLUSH_TILES": "4"},
                                 {"MDG_SWAR_UNIFORM": "0"}, {"MDG_PLANES_GATHER": "1", "MDG_PLANES_PREFETCH": "3"},
                                 {"MDG_PLANES_INDELS": "1"}, {"MDG_PLANES_INDELS": "1", "MDG_PLANES_GATHER": "1"}])
@pytest.mark.parametrize("n_libs", [2, 3])
def test_libraries_in_one_launch(env, n_libs, monkeypatch):
    """Two libraries are counted by ONE launch of the warp-specialised kernel (a read's library picks its counters, event
    tables and the lists of left-over reads); three fall back to a launch per library over index lists.  Equal-length and
    mixed stretches, indels, clips, filtered flags: the tables of every library equal the oracle's."""
    from mapdamage_b200.batch import concatenate

    for key, value in env.items():
        monkeypatch.setenv(key, value)
    reference = synth.make_reference([300_000, 70_000], seed=8, other_rate=0.001)
    with DamageEngine(n_libraries=n_libs, max_reads=1024) as engine:
        engine.set_reference(reference)
        parts = []
        for k, kw in enumerate((dict(length=(100, 100)), dict(length=(35, 140), mix=(6, 1, 1, 2), read_n_rate=0.01, paired=True),
                                dict(length=(100, 100), filtered_rate=0.2))):
            dev = engine.synth_batch(250_000, seed=70 + k, n_libs=n_libs, **kw)
            parts.append(engine.download(dev))
            dev.free()
    batch = concatenate(parts)
    assert len(np.unique(batch.lib)) == n_libs
    want = oracle.count(batch, reference, n_lib=n_libs, lg_bins=8192, threads=8)
    with DamageEngine(n_libraries=n_libs, max_reads=0) as engine:
        engine.set_reference(reference)
        dev = engine.upload(batch)
        engine.count_resident(dev)
        got = engine.tables()
        dev.free()
    for name, a, b in zip(("misincorp", "dnacomp", "lghist"), got, want):
        assert np.array_equal(a, b), name


@pytest.mark.parametrize("env", [{}, {"MDG_PLANES_WS": "0"}, {"MDG_PLANES_WS": "2x8+8"}, {"MDG_KERNEL": "staged"}])
@pytest.mark.parametrize("length", [(100, 100), (60, 140)])
def test_identical_reads_do_not_overflow_the_block_counters(env, length, monkeypatch):
    """12 M forward reads of ONE place of the genome (an amplicon, a tower of PCR duplicates): every read adds to the same
    table cells, so a block's private counters (12-bit vertical counters, 16-bit nibble counters) reach their capacity
    as fast as they ever can and must be reduced into the 64-bit tables in time."""
    from mapdamage_b200.batch import ReadBatch

    for key, value in env.items():
        monkeypatch.setenv(key, value)
    reference = synth.make_reference([50_000], seed=9)
    n = 12_000_000
    lo, hi = length
    lens = np.full(n, lo, dtype=np.int64) if lo == hi else lo + (np.arange(n, dtype=np.int64) * 7) % (hi - lo + 1)
    template = synth.simulate_reads(reference, 1, seed=3, length=(hi, hi))  # one read of the longest length, as packed bases
    row = np.zeros((hi + 1) // 2, dtype=np.uint8)
    row[:] = template.seq4[:row.shape[0]]
    padded = (lens + 1) & ~1
    base_off = np.zeros(n, dtype=np.int64)
    np.cumsum(padded[:-1], out=base_off[1:])
    if lo == hi:
        seq4 = np.tile(row[:(lo + 1) // 2], n)
    else:
        seq4 = np.zeros(int(base_off[-1] + padded[-1]) // 2, dtype=np.uint8)
        for ln in range(lo, hi + 1):
            idx = np.nonzero(lens == ln)[0]
            if idx.size:
                at = (base_off[idx] // 2)[:, None] + np.arange((ln + 1) // 2)[None, :]
                seq4[at] = row[:(ln + 1) // 2][None, :]
    batch = ReadBatch(flag=np.zeros(n, np.uint16), tid=np.zeros(n, np.int32), pos=np.full(n, int(template.pos[0]), np.int32),
                      l_seq=lens, base_off=base_off, cigar_off=np.arange(n + 1), cigar=(lens << 4).astype(np.uint32), seq4=seq4)
    want = oracle.count(batch, reference, lg_bins=8192, threads=8)
    with DamageEngine(max_reads=0) as engine:
        engine.set_reference(reference)
        dev = engine.upload(batch)
        engine.count_resident(dev)
        got = engine.tables()
        dev.free()
    for name, a, b in zip(("misincorp", "dnacomp", "lghist"), got, want):
        assert np.array_equal(a, b), name


@pytest.mark.parametrize("name", ["pe_mixed", "short", "long"])
@pytest.mark.parametrize("n_lib", [1, 2, 3])
@pytest.mark.parametrize("pinned", ["1", "0"])
def test_indel_reads_in_the_planes_kernel(name, n_lib, pinned, monkeypatch):
    """MDG_PLANES_INDELS=1: reads with one insertion / deletion and no clips are staged by the warp-specialised bit-plane
    kernel itself in tiles with two windows per read (the planes of the sequence that is discontinuous at the gap at two
    alignments, gap columns as events); 0: all of them go to count_staged_kernel.  The tables do not change."""
    monkeypatch.setenv("MDG_PLANES_INDELS", pinned)
    reference = synth.make_reference([300_000, 150_000, 4_000], seed=5, other_rate=0.002)
    kw = dict(SYNTH[name])
    kw["mix"] = (2, 3, 3, 2)  # mostly indel reads
    kw.setdefault("read_n_rate", 0.02)
    batch = synth.simulate_reads(reference, 60_000, seed=33, n_libs=n_lib, **kw)
    want = oracle.count(batch, reference, n_lib=n_lib, lg_bins=8192, threads=4)
    got = run_engine(batch, reference, n_lib=n_lib, chunks=2)
    for key, a, b in zip(("misincorp", "dnacomp", "lghist"), got, want):
        assert np.array_equal(a, b), key
    assert want[0][:, :, :, 4 + 5 * 4:4 + 5 * 4 + 4].sum() > 1000  # insertion classes are populated


@pytest.mark.parametrize("name", ["pe_mixed", "short", "long"])
@pytest.mark.parametrize("min_qual,n_lib", [(0, 1), (20, 1), (0, 2), (13, 3)])
def test_indel_reads_in_the_staged_kernel(name, min_qual, n_lib, monkeypatch):
    """MDG_STAGE_INDELS=1: reads with one insertion / deletion are counted by the bit-sliced kernel itself (near block,
    gap columns, lagging far block) instead of the general kernel; the tables do not change."""
    monkeypatch.setenv("MDG_STAGE_INDELS", "1")
    reference = synth.make_reference([300_000, 150_000, 4_000], seed=5, other_rate=0.002)
    kw = dict(SYNTH[name])
    kw["mix"] = (2, 3, 3, 2)  # mostly indel reads
    batch = synth.simulate_reads(reference, 50_000, seed=31, n_libs=n_lib, **kw)
    want = oracle.count(batch, reference, minqual=min_qual, n_lib=n_lib, lg_bins=8192, threads=4)
    got = run_engine(batch, reference, n_lib=n_lib, min_qual=min_qual, chunks=2)
    for key, a, b in zip(("misincorp", "dnacomp", "lghist"), got, want):
        assert np.array_equal(a, b), key
    assert want[0][:, :, :, 4 + 5 * 4:4 + 5 * 4 + 4].sum() > 1000  # insertion classes are populated


@pytest.mark.parametrize("case_dir,params", [c for c in golden_cases("counting")
                                              if not c.values[1]["exception"]])
def test_counting_golden_with_indels_staged(case_dir, params, tmp_path, monkeypatch):
    monkeypatch.setenv("MDG_STAGE_INDELS", "1")
    batch, reference, libraries, _ = load_counting_case(case_dir, params, tmp_path)
    L, A = params["length"], params["around"]
    mis, comp, lg, overflow = run_engine(batch, reference, n_lib=len(libraries), length=L, around=A,
                                         min_qual=params["minqual"], chunks=2)
    render_tables(tmp_path / "out", libraries, L, A, mis, comp, lg)
    assert_tables_equal(tmp_path / "out", case_dir)


@pytest.mark.parametrize("sorted_positions", [False, True])
def test_device_made_genome_and_sorted_reads(sorted_positions):
    """mdg_synth_reference / mdg_reference_download (benchmarks with genomes far larger than L2) and reads placed in
    coordinate order: the tables equal the oracle's on the downloaded genome."""
    lengths = [700_001, 90_000, 1_234_567]
    with DamageEngine(max_reads=0) as engine:
        names, lens = engine.synth_reference(lengths, seed=17)
        reference = engine.reference_host(names, lens)
        assert reference.lengths == lengths and set(np.unique(reference.sequences[0])) == set(b"ACGT")
        dev = engine.synth_batch(120_000, seed=5, length=(40, 130), mix=(7, 1, 1, 1), with_qual=False,
                                 sorted_positions=sorted_positions)
        engine.count_resident(dev)
        got = engine.tables()
        host = engine.download(dev)
    if sorted_positions:
        key = host.tid.astype(np.int64) * (1 << 32) + host.pos
        assert np.all(np.diff(key) >= 0)
    want = oracle.count(host, reference, lg_bins=8192, threads=4)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
