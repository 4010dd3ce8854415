"""Host-side mirror of the reference's counting pass (``main.py:126-231``).

The reference builds three accumulators (``main.py:147-155``), feeds them one
read at a time (``main.py:165-217``) and writes ``misincorporation.txt``,
``dnacomp.txt`` and ``lgdistribution.txt`` (``main.py:229-231``).  Here the
loop body is the CUDA counting pass: alignments are decoded on the host into
SoA batches, streamed through :class:`~mapdamage_b200.engine.DamageEngine`
(double-buffered ``mdg_count_submit``), and the accumulators are filled from
the device's count slabs.  The three classes keep the reference's constructor
arguments, ``.data`` layout and ``.write()`` bytes, so everything downstream
(the R plotting / Bayesian stage) reads the same files.

Input is BAM (decoded natively on host threads, ``bamio.BamReader``: pysam/htslib are absent from this
image) or SAM text (per-record Python, for small files and tests).
"""
import logging
import os
from pathlib import Path

import numpy as np

from . import downsample as _downsample
from . import statistics
from .batch import BAMError, FILTERED_FLAGS, BatchBuilder
from .engine import DamageEngine
from .refgenome import Reference
from .samtext import iter_sam

TABLE_FILES = ("misincorporation.txt", "dnacomp.txt", "lgdistribution.txt")


def count_alignments(filename, ref, length=70, around=10, min_basequal=0, merge_libraries=False,
                     folder=None, batch_reads=1 << 20, device=0, lg_bins=1 << 16, engine=None,
                     downsample=None, downsample_seed=None):
    """Counting pass over one alignment file.

    ``ref`` is a FASTA path or a :class:`Reference`.  Returns
    ``(misincorp, dnacomp, lgdistrib)`` -- mirrors of the reference's
    ``MisincorporationRates`` / ``DNAComposition`` / ``FragmentLengths`` -- and,
    when ``folder`` is given, writes the three tables into it the way
    ``main.py:229-231`` does.  Raises :class:`~mapdamage_b200.batch.BAMError`
    where the reference does (read without a known read group, ``reader.py:63-81``).

    ``downsample`` / ``downsample_seed`` are the reference's ``-n`` / ``--downsample-seed`` (``reader.py:84-96``):
    a fraction below 1 or a number of reads, drawn with the reference's own sequence of random numbers
    (:mod:`~mapdamage_b200.downsample`).
    """
    log = logging.getLogger(__name__)
    filename, is_bam, is_stream = input_kind(filename)
    sampler = _downsample.sampler_for(downsample, downsample_seed)
    if is_stream and isinstance(sampler, _downsample.ReservoirSampler):
        # the reservoir is drawn in a first pass over the flags; the reference holds the sampled records in memory
        # instead (reader.py:144-164), which a pipe allows and this does not
        raise ValueError("down-sampling to a fixed number of reads needs a file, not a pipe")
    if is_bam:
        return _count_bam(filename, ref, length, around, min_basequal, merge_libraries, folder, batch_reads, device,
                          lg_bins, engine, sampler)
    if isinstance(sampler, _downsample.ReservoirSampler):
        # the reservoir is known once the whole stream has been walked: first pass over the flags only
        sampler.feed(sum(1 for record in iter_sam(filename)[1] if not record.flag & FILTERED_FLAGS))
        sampler = _downsample.Selection(sampler.selected())
    header, records = iter_sam(filename)
    if sampler is not None:
        records = _drawn(records, sampler)
    reference = ref if isinstance(ref, Reference) else Reference.from_fasta(ref)
    reference = reference.reordered(header.references, header.lengths)
    builder = BatchBuilder(readgroups=None if merge_libraries else header.libraries(),
                           merge_libraries=merge_libraries, apply_filter=True)
    libraries = builder.libraries
    own_engine = engine is None
    if own_engine:
        engine = DamageEngine(length=length, around=around, min_qual=min_basequal,
                              n_libraries=max(1, len(libraries)), lg_bins=lg_bins, device=device,
                              max_reads=batch_reads, max_cigar_ops=8 * batch_reads, max_bases=512 * batch_reads)
    try:
        engine.set_reference(reference)
        n_kept = 0
        for record in records:
            if builder.add(record):
                n_kept += 1
            if n_kept and n_kept % batch_reads == 0:
                _submit(engine, builder)
        _submit(engine, builder)
        mis, comp, lg = engine.tables()
        overflow = engine.lg_overflow()
    finally:
        if own_engine:
            engine.close()
    log.debug("Counted %d of %d alignments", n_kept, builder.n_seen)
    return _finish(libraries, length, around, mis, comp, lg, overflow, folder)


def input_kind(filename):
    """``(path, is_bam, is_stream)``.  The reference hands any path, ``-`` or pipe to ``pysam.AlignmentFile``
    (``reader.py:34-38``), which looks at the content; so does this for files (BGZF magic), while a pipe is taken for
    BAM unless it is named ``*.sam``."""
    if str(filename) == "-":
        return Path("/dev/stdin"), True, True
    path = Path(filename)
    if path.is_fifo() or path.is_char_device():
        return path, path.suffix.lower() != ".sam", True
    with open(path, "rb") as handle:
        magic = handle.read(2)
    return path, magic == b"\x1f\x8b", False


def _drawn(records, sampler, chunk=4096):
    """Records the flag filter drops pass through (the builder counts and drops them); of the others, those drawn."""
    keep, k = sampler.mask(chunk), 0
    for record in records:
        if record.flag & FILTERED_FLAGS:
            yield record
            continue
        if k == chunk:
            keep, k = sampler.mask(chunk), 0
        if keep[k]:
            yield record
        k += 1


def _count_bam(filename, ref, length, around, min_basequal, merge_libraries, folder, batch_reads, device, lg_bins, engine,
               sampler=None):
    """BAM input: batches come straight out of the native decoder (``bamio.BamReader``) into pinned buffers."""
    from .bamio import BamReader

    log = logging.getLogger(__name__)
    if isinstance(sampler, _downsample.ReservoirSampler):
        with BamReader(filename, merge_libraries=True, apply_filter=True) as reader:  # flags only: no library lookup
            buffers = reader.buffers(batch_reads, with_qual=False)
            while True:
                batch = reader.read_batch(buffers=buffers)
                if batch is None:
                    break
                sampler.feed(batch.n)
        sampler = _downsample.Selection(sampler.selected())
    if sampler is None and not Path(filename).is_fifo() and not Path(filename).is_char_device() \
            and os.environ.get("MDG_BAM_HOST") != "1":
        return _count_bam_on_device(filename, ref, length, around, min_basequal, merge_libraries, folder, device, lg_bins,
                                    engine)
    # the host decoder: pipes, down-sampling (the draws need every kept read's turn on the host), MDG_BAM_HOST=1
    with BamReader(filename, merge_libraries=merge_libraries, apply_filter=True,
                   lenient_libraries=sampler is not None) as reader:
        reference = ref if isinstance(ref, Reference) else Reference.from_fasta(ref)
        reference = reference.reordered(reader.header.references, reader.header.lengths)
        libraries = reader.libraries
        own_engine = engine is None
        if own_engine:
            engine = DamageEngine(length=length, around=around, min_qual=min_basequal,
                                  n_libraries=max(1, len(libraries)), lg_bins=lg_bins, device=device,
                                  max_reads=batch_reads)
        try:
            engine.set_reference(reference)
            # one more set of host buffers than staging slots: a set is reused only after its copy has drained
            sets = [reader.buffers(engine.max_reads, engine.max_cigar_ops, engine.max_bases,
                                   with_qual=min_basequal > 0, empty=engine.arena.empty) for _ in range(3)]
            n_kept = turn = 0
            while True:
                batch = reader.read_batch(buffers=sets[turn % 3])
                if batch is None:
                    break
                if sampler is not None:
                    # every read of the batch passed the flag filter; those not drawn get a filtered flag
                    drawn = sampler.mask(batch.n)
                    _downsample.apply_mask(batch, np.ones(batch.n, dtype=np.bool_), drawn)
                    # the reference looks up the library of the reads it yields only (reader.py:134-164, main.py:165-170)
                    for index, message in reader.library_failures():
                        if drawn[index]:
                            raise BAMError(message)
                        batch.lib[index] = 0
                engine.count(batch, compact=False)
                n_kept += batch.n
                turn += 1
            mis, comp, lg = engine.tables()
            overflow = engine.lg_overflow()
            n_seen = reader.records_seen
        finally:
            if own_engine:
                engine.close()
    log.debug("Counted %d of %d alignments", n_kept, n_seen)
    return _finish(libraries, length, around, mis, comp, lg, overflow, folder)


def _count_bam_on_device(filename, ref, length, around, min_basequal, merge_libraries, folder, device, lg_bins, engine):
    """BAM input decoded on the GPU (``bamio.DeviceBamStream``): the host reads the file and nothing else; slabs are
    inflated, cut into records, scattered into the batch layout and counted in HBM."""
    from .bamio import BamReader, DeviceBamStream

    import time

    log = logging.getLogger(__name__)
    clock = [("start", time.perf_counter())]
    with BamReader(filename, threads=2, merge_libraries=merge_libraries, apply_filter=True) as reader:
        header, libraries = reader.header, reader.libraries
    clock.append(("header", time.perf_counter()))
    reference = ref if isinstance(ref, Reference) else Reference.from_fasta(ref)
    reference = reference.reordered(header.references, header.lengths)
    clock.append(("fasta", time.perf_counter()))
    own_engine = engine is None
    if own_engine:
        engine = DamageEngine(length=length, around=around, min_qual=min_basequal, n_libraries=max(1, len(libraries)),
                              lg_bins=lg_bins, device=device, max_reads=0)
    try:
        engine.set_reference(reference)
        clock.append(("engine + genome upload", time.perf_counter()))
        n_kept = 0
        with DeviceBamStream(engine, filename, merge_libraries=merge_libraries, apply_filter=True,
                             with_qual=min_basequal > 0) as stream:
            clock.append(("stream open", time.perf_counter()))
            for batch in stream:
                engine.count_resident(batch)
                n_kept += batch.n
            engine.sync()
            clock.append(("decode + count", time.perf_counter()))
            n_seen = stream.stats()["records_seen"]
        clock.append(("stream close", time.perf_counter()))
        mis, comp, lg = engine.tables()
        overflow = engine.lg_overflow()
    finally:
        if own_engine:
            engine.close()
    clock.append(("tables + engine close", time.perf_counter()))
    if os.environ.get("MDG_TIMING"):
        import sys

        print("count_alignments (device): " + ", ".join("%s %.3f s" % (name, t - clock[i][1]) for i, (name, t) in enumerate(clock[1:])),
              file=sys.stderr)
    log.debug("Counted %d of %d alignments", n_kept, n_seen)
    return _finish(libraries, length, around, mis, comp, lg, overflow, folder)


def _finish(libraries, length, around, mis, comp, lg, overflow, folder):
    misincorp = statistics.MisincorporationRates(libraries, length).load(mis)
    dnacomp = statistics.DNAComposition(libraries, around, length).load(comp)
    lgdistrib = statistics.FragmentLengths(libraries).load(lg, overflow)
    if folder is not None:
        folder = Path(folder)
        folder.mkdir(parents=True, exist_ok=True)
        misincorp.write(folder / TABLE_FILES[0])
        dnacomp.write(folder / TABLE_FILES[1])
        lgdistrib.write(folder / TABLE_FILES[2])
    return misincorp, dnacomp, lgdistrib


def _submit(engine, builder):
    batch = builder.finish()
    if not batch.n:
        return
    if not engine.fits(batch):
        # long reads / long CIGARs: halve until the staging slot takes it
        for part in batch.split(2):
            _submit_batch(engine, part)
        return
    engine.count(batch)


def _submit_batch(engine, batch):
    if not batch.n:
        return
    if engine.fits(batch) or batch.n == 1:
        engine.count(batch)
    else:
        for part in batch.split(2):
            _submit_batch(engine, part)
