"""Host-side mirror of the reference's rescale pass (``rescale.py:285-383``).

``rescale_qual(ref, options)`` keeps the reference's signature, log messages
and return code; the per-read work (``_rescale_qual_read``, ``rescale.py:195-282``)
runs in ``rescale_kernel``.  Every record of the input is written to the
output in input order (``rescale.py:300,344``); rescaled reads get their new
qualities and the ``MR:f`` tag.  The substitution summary the reference logs
(``_record_subs`` / ``_qual_summary_subs`` / ``_print_subs``, ``rescale.py:106-192``)
is rebuilt on the host from integer histograms the kernel accumulates.

Input / output is BAM (native decode and encode, ``bamio``) or, for small files and tests, SAM text.
"""
import logging
import struct
import time
from pathlib import Path

import numpy as np

from .batch import BatchBuilder
from .engine import DamageEngine
from .refgenome import Reference
from .rescale_model import RescaleError, RescaleModel, get_corr_prob
from .samtext import format_record, iter_sam

__all__ = ["rescale_qual", "RescaleError", "SubstitutionSummary"]


def _phred_pval(qual):
    """``_phred_char_to_pval`` (``rescale.py:18-20``) of Phred score ``qual``."""
    return 10 ** (-(float(qual + 33) - float(33)) / 10)


class SubstitutionSummary:
    """``subs`` of the reference (``rescale.py:82-105``) from the device histograms.

    ``sub[type][slot][Q]`` counts rescaled C>T / G>A columns, ``rev[type][Q]``
    the T>C / A>G columns, ``ref_count`` the reference bases seen.  The float
    sums are formed per histogram cell instead of per base, so they can differ
    from the reference's running sums in the last bits; they are only ever
    printed with four decimals.
    """

    def __init__(self, model, sub, rev, ref_count):
        self.data = {}
        names = ("CT", "GA")
        for t, name in enumerate(names):
            before = np.zeros(130, dtype=np.int64)
            after = np.zeros(130, dtype=np.int64)
            pvals = pvals_before = 0.0
            for slot in range(model.n_slots):
                corr = model.prob[t, slot]
                for qual in np.flatnonzero(sub[t, slot]):
                    count = int(sub[t, slot, qual])
                    before[qual] += count
                    after[int(model.lut[t, slot, qual])] += count
                    pseq = 1 - _phred_pval(int(qual))
                    pvals += count * ((1 - corr) * pseq)
                    pvals_before += count * pseq
            self.data[name + "-before"] = before
            self.data[name + "-after"] = after
            self.data[name + "-pvals"] = pvals
            self.data[name + "-pvals_before"] = pvals_before
        for t, name in enumerate(("TC", "AG")):
            hist = np.zeros(130, dtype=np.int64)
            hist[:94] = rev[t]
            self.data[name + "-before"] = hist
            self.data[name + "-after"] = hist.copy()
            self.data[name + "-pvals"] = float(sum(int(rev[t, q]) * (1 - _phred_pval(int(q)))
                                                   for q in np.flatnonzero(rev[t])))
        for code, base in enumerate("ACGT"):
            self.data[base] = int(ref_count[code])

    def at_least(self, key, level):
        """``subs[key + "-Q%d"]`` of ``_qual_summary_subs`` (``rescale.py:142-161``)."""
        return int(self.data[key][level:].sum())

    def log_lines(self):
        """The lines ``_print_subs`` logs (``rescale.py:164-192``)."""
        lines = ["Expected substition frequencies before and after rescaling:"]
        for sub in ("CT", "TC", "GA", "AG"):
            base_count = self.data[sub[0]]
            if base_count:
                pvals = self.data[sub + "-pvals"]
                pvals_before = self.data.get(sub + "-pvals_before", pvals)
                lines.append("    %s>%s    %.4f    %.4f"
                             % (sub[0], sub[1], pvals_before / base_count, pvals / base_count))
            else:
                lines.append("\t%s\tNA\t\tNA" % sub)
        lines.append("Quality metrics before and after scaling:")
        for sub in ("CT", "GA"):
            for qual in (0, 10, 20, 30, 40):
                lines.append("    %s-Q%02i% 10i% 10i" % (sub, qual, self.at_least(sub + "-before", qual),
                                                           self.at_least(sub + "-after", qual)))
        return lines


def _walk_ends_in_deletion(record):
    """True when the 5'->3' walk of the alignment ends on deletion columns (``rescale.py:255-261``)."""
    columns = [op for op, n in record.cigar if op in (0, 1, 2, 7, 8) and n > 0]
    if not columns:
        return False
    return (columns[0] if record.flag & 0x10 else columns[-1]) == 2


def _mr_text(value):
    """``MR:f`` as a SAM writer prints the float32 a BAM ``f`` tag would store."""
    return "MR:f:%r" % struct.unpack("<f", struct.pack("<f", float(value)))[0]


def _rescale_qual_core(ref, options, engine=None, batch_reads=1 << 18):
    log = logging.getLogger(__name__)
    corr_prob = get_corr_prob(Path(options.folder) / "Stats_out_MCMC_correct_prob.csv",
                              rescale_length_5p=options.rescale_length_5p,
                              rescale_length_3p=options.rescale_length_3p)
    model = RescaleModel(corr_prob, options.rescale_length_5p, options.rescale_length_3p)
    from .counting import input_kind

    filename, is_bam, _ = input_kind(options.filename)  # by content, like pysam (rescale.py:298)
    if is_bam:
        return _rescale_bam(ref, options, model, engine, batch_reads, log)
    header, records = iter_sam(filename)
    reference = ref if isinstance(ref, Reference) else Reference.from_fasta(ref)
    reference = reference.reordered(header.references, header.lengths)

    own_engine = engine is None
    if own_engine:
        engine = DamageEngine(max_reads=batch_reads, max_cigar_ops=8 * batch_reads, max_bases=512 * batch_reads,
                              device=getattr(options, "device", 0))
    try:
        engine.set_reference(reference)
        engine.set_rescale_model(model)
        with open(options.rescale_out, "wt") as out:
            for line in header.lines:
                out.write(line + "\n")
            chunk = []
            for record in records:
                if "MR" in record.tags and not (record.flag & 0x4) and record.qual is not None \
                        and _would_rescale(record):
                    # rescale.py:277-278
                    raise SystemExit("Read: %s already has a MR tag, can't rescale"
                                     % format_record(record, header))
                chunk.append(record)
                if len(chunk) == batch_reads:
                    _rescale_chunk(engine, header, chunk, out, log)
                    chunk = []
            _rescale_chunk(engine, header, chunk, out, log)
        stats = engine.rescale_stats()
        summary = SubstitutionSummary(model, *engine.rescale_hist(model.n_slots))
    finally:
        if own_engine:
            engine.close()

    return _report(log, stats, summary)


def _rescale_bam_on_device(ref, options, model, engine, log):
    """BAM in, BAM out, on the GPU: the host reads one file and writes the other; inflate, record scatter, the
    rescale kernels, record re-emission (new qualities, ``MR:f``) and BGZF deflate all run in HBM
    (``bamio.DeviceBamStream``).  Every record is written, in input order, under the input's header
    (``rescale.py:298-300,344``)."""
    import os

    from .bamio import BamReader, BamWriter, DeviceBamStream
    from .counting import input_kind

    path = input_kind(options.filename)[0]
    with BamReader(path, threads=2, merge_libraries=True, apply_filter=False) as reader:
        header = reader.header
    reference = ref if isinstance(ref, Reference) else Reference.from_fasta(ref)
    reference = reference.reordered(header.references, header.lengths)
    own_engine = engine is None
    if own_engine:
        engine = DamageEngine(max_reads=0, device=getattr(options, "device", 0))
    timings = getattr(options, "timings", None)
    try:
        engine.set_reference(reference)
        engine.set_rescale_model(model)
        too_long, first = 0, 0
        with DeviceBamStream(engine, path, merge_libraries=True, apply_filter=False, with_qual=True, want_mr=True,
                             slab_bytes=int(os.environ.get("MDG_RESCALE_SLAB", "0"))) as stream, \
                BamWriter(options.rescale_out, header) as writer:
            for batch in stream:
                _, status = engine.rescale_resident(batch, want_results=True)
                engine.sync()  # raises the "quality and sequence mismatch" data error (rescale.py:266-273)
                clash = np.flatnonzero(stream.has_mr(batch) & (status & 1))
                if clash.size:  # rescale.py:277-278
                    raise SystemExit("Read: %s already has a MR tag, can't rescale"
                                     % _names_at(path, [first + int(clash[0])])[0])
                now = engine.rescale_stats()["alignment_longer_than_read"]
                if now != too_long:  # rescale.py:255-261; rare, so the names are dug out of the file again
                    host = engine.download(batch)
                    hits = []
                    for i in np.flatnonzero(status & 1):
                        columns = [op for op, n in host.cigar_of(int(i)) if op in (0, 1, 2, 7, 8) and n > 0]
                        if columns and (columns[0] if host.flag[i] & 0x10 else columns[-1]) == 2:
                            hits.append(first + int(i))
                    for name in _names_at(path, hits):
                        log.warning("The aligment of the read is longer than the actual read %s", name)
                    too_long = now
                stream.encode(batch, writer)
                first += batch.n
            flushed = stream.flush()
            if timings is not None:
                timings.update(stream.stats())
                timings.update(zip(("bytes_uncompressed", "bytes_compressed", "encode_s", "write_wait_s"), flushed))
        stats = engine.rescale_stats()
        summary = SubstitutionSummary(model, *engine.rescale_hist(model.n_slots))
    finally:
        if own_engine:
            engine.close()
    return _report(log, stats, summary)


def _names_at(path, indices):
    """Names of the records at the given positions of a BAM file (error / warning texts only)."""
    from .bamio import BamReader

    wanted, names, at = sorted(set(indices)), {}, 0
    with BamReader(path, merge_libraries=True, apply_filter=False) as reader:
        while wanted:
            batch = reader.read_batch(max_reads=1 << 18, keep_raw=True)
            if batch is None:
                break
            while wanted and wanted[0] < at + batch.n:
                index = wanted.pop(0)
                names[index] = _record_name(batch, index - at)
            at += batch.n
    return [names.get(i, "?") for i in indices]


def _rescale_bam(ref, options, model, engine, batch_reads, log):
    """BAM in, BAM out: batches from the native decoder, records re-emitted by the native encoder."""
    import os

    from .bamio import BamReader, BamWriter

    from .counting import input_kind

    if os.environ.get("MDG_BAM_HOST") != "1" and not input_kind(options.filename)[2]:
        return _rescale_bam_on_device(ref, options, model, engine, log)

    with BamReader(input_kind(options.filename)[0], merge_libraries=True, apply_filter=False) as reader:
        reference = ref if isinstance(ref, Reference) else Reference.from_fasta(ref)
        reference = reference.reordered(reader.header.references, reader.header.lengths)
        own_engine = engine is None
        if own_engine:
            engine = DamageEngine(max_reads=batch_reads, device=getattr(options, "device", 0))
        try:
            engine.set_reference(reference)
            engine.set_rescale_model(model)
            buffers = reader.buffers(engine.max_reads, engine.max_cigar_ops, engine.max_bases, with_qual=True,
                                     empty=engine.arena.empty)
            too_long = 0
            with BamWriter(options.rescale_out, reader.header) as writer:
                while True:
                    batch = reader.read_batch(buffers=buffers, keep_raw=True)
                    if batch is None:
                        break
                    qual, mr, status = engine.rescale(batch, compact=False)
                    engine.sync()
                    clash = np.flatnonzero(batch.has_mr & (status & 1))
                    if clash.size:  # rescale.py:277-278
                        raise SystemExit("Read: %s already has a MR tag, can't rescale" % _record_name(batch, clash[0]))
                    now = engine.rescale_stats()["alignment_longer_than_read"]
                    if now != too_long:  # rescale.py:255-261; rare, so the names are dug out per record
                        for i in np.flatnonzero(status & 1):
                            cigar = batch.cigar_of(int(i))
                            columns = [op for op, n in cigar if op in (0, 1, 2, 7, 8) and n > 0]
                            if columns and (columns[0] if batch.flag[i] & 0x10 else columns[-1]) == 2:
                                log.warning("The aligment of the read is longer than the actual read %s",
                                            _record_name(batch, int(i)))
                        too_long = now
                    writer.write(batch, status=status, qual=qual, mr=mr)
            stats = engine.rescale_stats()
            summary = SubstitutionSummary(model, *engine.rescale_hist(model.n_slots))
        finally:
            if own_engine:
                engine.close()
    return _report(log, stats, summary)


def _record_name(batch, i):
    start = int(batch.raw_off[i])
    l_name = int(batch.raw[start + 12])
    return batch.raw[start + 36:start + 36 + l_name - 1].tobytes().decode("latin-1")


def _report(log, stats, summary):
    if stats["pairs"]:
        log.warning(
            "Processed %i paired reads, assumed to be non-overlapping, facing inwards "
            "and correctly paired; %i of these were excluded as improperly paired.",
            stats["pairs"], stats["improper_pairs"])
    if stats["without_quals"]:
        log.warning("Skipped %i reads without quality scores", stats["without_quals"])
    if not (np.array_equal(summary.data["TC-before"], summary.data["TC-after"])
            and np.array_equal(summary.data["AG-before"], summary.data["AG-after"])):
        raise RescaleError("Qualities for T.C and A.G transitions should not change in the rescaling. "
                           "Please file a bug on github.")
    for line in summary.log_lines():
        log.info("%s", line)
    return summary


def _would_rescale(record):
    """The pairing rule of ``rescale.py:305-342``: does the reference rescale this mapped read?"""
    if not (record.flag & 0x1):
        return True
    reverse, mate_reverse = bool(record.flag & 0x10), bool(record.flag & 0x20)
    same = record.tid == record.mtid
    return ((not reverse and mate_reverse and record.mpos > record.pos and same)
            or (reverse and not mate_reverse and record.mpos < record.pos and same))


def _rescale_chunk(engine, header, chunk, out, log):
    if not chunk:
        return
    builder = BatchBuilder(merge_libraries=True, apply_filter=False)
    for record in chunk:
        builder.add(record)
    batch = builder.finish(with_qual=True)
    qual, mr, status = engine.rescale(batch)
    engine.sync()  # raises the "quality and sequence mismatch" data error (rescale.py:266-273)
    for i, record in enumerate(chunk):
        if status[i] & 1:
            if _walk_ends_in_deletion(record):
                log.warning("The aligment of the read is longer than the actual read %s", record.qname)
            off, n = int(batch.base_off[i]), int(batch.l_seq[i])
            text = (qual[off:off + n] + 33).astype(np.uint8).tobytes().decode("latin-1")
            out.write(format_record(record, header, qual=text, extra_tags=(_mr_text(mr[i]),)) + "\n")
        else:
            out.write(format_record(record, header) + "\n")


def rescale_qual(ref, options):
    """``rescale.rescale_qual`` (``rescale.py:368-383``): 0 on success, 1 on failure."""
    from ._native import NativeError

    log = logging.getLogger(__name__)
    log.info("Rescaling BAM: '%s' -> '%s'", options.filename, options.rescale_out)
    start_time = time.time()
    try:
        _rescale_qual_core(ref, options)
    except RescaleError as error:
        log.error("%s", error)
        return 1
    except NativeError as error:
        # the device reports what the reference raises as exceptions inside its loop
        log.error("Unhandled exception: %s", error.message)
        return 1
    except Exception as error:
        log.error("Unhandled exception: %s", error)
        return 1
    log.debug("Rescaling completed in %f seconds", time.time() - start_time)
    return 0
