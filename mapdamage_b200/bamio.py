"""BAM input / output through the native decoder (``csrc/mdg_bamio.cpp``).

The reference iterates ``pysam.AlignmentFile`` objects (``reader.py:38,121-132``;
``rescale.py:298-300``) and writes one (``rescale.py:299,344``).  Here a
:class:`BamReader` yields whole struct-of-arrays batches -- BGZF inflate and the
record copies run on host threads inside the library -- and a :class:`BamWriter`
re-emits the records of a batch with rescaled qualities and ``MR`` tags.  Only
the header is handled in Python.
"""
import ctypes as C

import numpy as np

from . import _native
from .batch import BAMError, FILTERED_FLAGS, ReadBatch
from .samtext import SamHeader


class BamReader:
    """Iterates a BAM file as :class:`ReadBatch` objects.

    ``merge_libraries`` / read groups follow ``reader.BAMReader`` (``reader.py:44-50,63-81``):
    ``libraries`` is the sorted ``(sample, library)`` list indexing the count slabs.
    """

    def __init__(self, path, threads=0, merge_libraries=False, apply_filter=True, lenient_libraries=False):
        """``lenient_libraries``: a read without a usable read group does not fail its batch; it gets library 0xFFFF and
        :meth:`library_failures` lists such reads (the caller down-samples first, then decides: ``reader.py:134-164``)."""
        self._lib = _native.load()
        self._reader = C.c_void_p()
        code = self._lib.mdg_bam_open(str(path).encode(), threads, C.byref(self._reader))
        if code < 0:
            raise BAMError((self._lib.mdg_bam_error(None) or b"").decode())
        text = C.create_string_buffer(int(self._lib.mdg_bam_header_text(self._reader, None, 0)) + 1)
        self._lib.mdg_bam_header_text(self._reader, text, len(text))
        self.header = SamHeader()
        for line in text.value.decode("utf-8", "replace").splitlines():
            if line:
                self.header.add(line)
        # the binary reference list is authoritative (a BAM may lack @SQ lines)
        names, lengths = [], []
        for i in range(self._lib.mdg_bam_n_references(self._reader)):
            name = C.create_string_buffer(1024)
            length = C.c_uint32()
            self._lib.mdg_bam_reference(self._reader, i, name, len(name), C.byref(length))
            names.append(name.value.decode())
            lengths.append(int(length.value))
        self.header.set_references(names, lengths)
        self.apply_filter = apply_filter
        self.merge_libraries = merge_libraries
        if lenient_libraries:
            self._lib.mdg_bam_lenient_libraries(self._reader, 1)
        if merge_libraries:
            self.libraries = [("*", "*")]
            self._lib.mdg_bam_set_libraries(self._reader, None, None, 0)
        else:
            groups = self.header.libraries()
            self.libraries = sorted(set(groups.values()))
            index = {key: i for i, key in enumerate(self.libraries)}
            ids = (C.c_char_p * max(1, len(groups)))(*[rg.encode() for rg in groups])
            libs = np.array([index[groups[rg]] for rg in groups], dtype=np.uint16)
            if groups:
                self._lib.mdg_bam_set_libraries(self._reader, ids, libs.ctypes.data, len(groups))
            else:
                # no read groups in the header: every read fails, as in the reference (reader.py:67-73)
                ids = (C.c_char_p * 1)(b"\x00")
                self._lib.mdg_bam_set_libraries(self._reader, ids, np.zeros(1, np.uint16).ctypes.data, 1)

    def close(self):
        if self._reader:
            self._lib.mdg_bam_close(self._reader)
            self._reader = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def library_failures(self):
        """``[(index in the last batch, BAMError text)]`` of the reads without a usable read group."""
        out, buf = [], C.create_string_buffer(1024)
        while True:
            index = self._lib.mdg_bam_library_failure(self._reader, len(out), buf, len(buf))
            if index < 0:
                return out
            out.append((int(index), buf.value.decode("utf-8", "replace")))

    @property
    def records_seen(self):
        return int(self._lib.mdg_bam_records_seen(self._reader))

    @staticmethod
    def buffers(max_reads=1 << 20, max_cigar=None, max_bases=None, with_qual=True, empty=np.empty):
        """Arrays one batch is decoded into; ``empty(shape, dtype)`` allocates (pinned: ``engine.arena.empty``).
        Reusing a set of buffers for several ``read_batch`` calls saves the allocations."""
        max_cigar = max_cigar or 4 * max_reads
        max_bases = max_bases or 160 * max_reads
        arrays = {name: empty(max_reads, dtype) for name, dtype in ReadBatch.FIELDS if name != "cigar_off"}
        arrays["cigar_off"] = empty(max_reads + 1, np.uint32)
        arrays["cigar"] = empty(max_cigar, np.uint32)
        arrays["seq4"] = empty(max_bases // 2, np.uint8)
        arrays["qual"] = empty(max_bases, np.uint8) if with_qual else None
        return arrays

    def read_batch(self, max_reads=1 << 20, max_cigar=None, max_bases=None, with_qual=True, keep_raw=False,
                   buffers=None):
        """Next batch, or ``None`` at the end of the file.  ``keep_raw`` attaches ``raw`` / ``raw_off`` /
        ``has_mr`` (what :class:`BamWriter` needs).  ``buffers`` (from :meth:`buffers`) are decoded into in place:
        the batch returned is a view of them."""
        arrays = buffers if buffers is not None else self.buffers(max_reads, max_cigar, max_bases, with_qual)
        max_reads = arrays["flag"].shape[0]
        max_cigar = arrays["cigar"].shape[0]
        max_bases = arrays["seq4"].shape[0] * 2
        s = _native.Batch()
        for name, array in arrays.items():
            setattr(s, name, None if array is None else array.ctypes.data)
        raw = raw_off = has_mr = None
        raw_cap = 0
        if keep_raw:
            raw_cap = max_bases * 2 + 256 * max_reads
            raw = np.empty(raw_cap, dtype=np.uint8)
            raw_off = np.empty(max_reads + 1, dtype=np.uint64)
            has_mr = np.empty(max_reads, dtype=np.uint8)
        n_cigar, n_bases = C.c_int64(), C.c_int64()
        n = self._lib.mdg_bam_read_batch(
            self._reader, C.byref(s), max_reads, max_cigar, max_bases, FILTERED_FLAGS if self.apply_filter else 0,
            None if raw is None else raw.ctypes.data, raw_cap, None if raw_off is None else raw_off.ctypes.data,
            None if has_mr is None else has_mr.ctypes.data, C.byref(n_cigar), C.byref(n_bases))
        if n < 0:
            raise BAMError((self._lib.mdg_bam_error(self._reader) or b"").decode("utf-8", "replace"))
        if n == 0:
            return None
        trimmed = {name: arrays[name][:n] for name, _ in ReadBatch.FIELDS if name != "cigar_off"}
        trimmed["cigar_off"] = arrays["cigar_off"][:n + 1]
        trimmed["cigar"] = arrays["cigar"][:n_cigar.value]
        trimmed["seq4"] = arrays["seq4"][:n_bases.value // 2]
        trimmed["qual"] = None if arrays["qual"] is None else arrays["qual"][:n_bases.value]
        batch = ReadBatch(**trimmed)
        if keep_raw:
            batch.raw, batch.raw_off, batch.has_mr = raw, raw_off[:n + 1], has_mr[:n]
        return batch

    def __iter__(self):
        while True:
            batch = self.read_batch()
            if batch is None:
                return
            yield batch


class DeviceBamStream:
    """A BAM file decoded on the GPU of ``engine`` (``mdg_bam_stream_*``, ``csrc/mdg_bamdev.cuh``).

    The host only reads the file; BGZF inflate, CRC check, record boundaries, the struct-of-arrays scatter and the read
    group -> library lookup run on the device, one slab ahead of the caller.  Iterating yields
    :class:`~mapdamage_b200.engine.DeviceBatch` objects resident in HBM (valid until the next one is asked for): feed
    them to ``engine.count_resident`` / ``engine.rescale_resident``.  The header comes from a host :class:`BamReader`.
    """

    def __init__(self, engine, path, merge_libraries=False, apply_filter=True, with_qual=True, want_mr=False,
                 slab_bytes=0):
        from .engine import DeviceBatch

        self._DeviceBatch = DeviceBatch
        self._lib = _native.load()
        self.engine = engine
        with BamReader(path, threads=2, merge_libraries=merge_libraries, apply_filter=apply_filter) as reader:
            self.header = reader.header
            self.libraries = reader.libraries
            data_start = int(self._lib.mdg_bam_data_start(reader._reader))
            groups = {} if merge_libraries else self.header.libraries()
        self.path = path
        self._stream = C.c_void_p()
        code = self._lib.mdg_bam_stream_open(engine._ctx, str(path).encode(), data_start, len(self.header.references),
                                             int(slab_bytes), C.byref(self._stream))
        if code < 0:
            raise BAMError((self._lib.mdg_bam_stream_error(None) or b"").decode())
        if not merge_libraries:
            index = {key: i for i, key in enumerate(self.libraries)}
            if groups:
                ids = (C.c_char_p * len(groups))(*[rg.encode() for rg in groups])
                libs = np.array([index[groups[rg]] for rg in groups], dtype=np.uint16)
                self._lib.mdg_bam_stream_set_libraries(self._stream, ids, libs.ctypes.data, len(groups))
            else:
                # no read groups in the header: every read fails, as in the reference (reader.py:67-73)
                ids = (C.c_char_p * 1)(b"\x00")
                self._lib.mdg_bam_stream_set_libraries(self._stream, ids, np.zeros(1, np.uint16).ctypes.data, 1)
        self._drop = FILTERED_FLAGS if apply_filter else 0
        self._with_qual, self._want_mr = bool(with_qual), bool(want_mr)

    def next_batch(self):
        """The next slab's records as a resident batch, or ``None`` at the end of the file."""
        handle = C.c_void_p()
        n = self._lib.mdg_bam_stream_next(self._stream, self._drop, int(self._with_qual), int(self._want_mr),
                                          C.byref(handle))
        if n < 0:
            raise BAMError((self._lib.mdg_bam_stream_error(self._stream) or b"").decode("utf-8", "replace"))
        if n == 0:
            return None
        batch = self._DeviceBatch(self.engine, handle, int(n), has_qual=self._with_qual)
        batch.free = lambda: None  # owned by the stream
        return batch

    def __iter__(self):
        while True:
            batch = self.next_batch()
            if batch is None:
                return
            yield batch

    def has_mr(self, batch):
        flags = np.empty(batch.n, dtype=np.uint8)
        if self._lib.mdg_bam_stream_has_mr(self._stream, flags.ctypes.data, batch.n) < 0:
            raise BAMError((self._lib.mdg_bam_stream_error(self._stream) or b"").decode())
        return flags

    def encode(self, batch, writer):
        """Re-emits the records of ``batch`` (the one handed out last) through ``writer`` (``mdg_bam_encode_batch``)."""
        if self._lib.mdg_bam_encode_batch(self._stream, batch.handle, writer._writer) < 0:
            raise OSError((self._lib.mdg_bam_stream_error(self._stream) or b"").decode())

    def flush(self):
        """Waits for the last file write; returns ``(uncompressed bytes, compressed bytes, encode s, write-wait s)``."""
        a, b = C.c_int64(), C.c_int64()
        t = (C.c_double * 2)()
        if self._lib.mdg_bam_encode_flush(self._stream, C.byref(a), C.byref(b), t) < 0:
            raise OSError((self._lib.mdg_bam_stream_error(self._stream) or b"").decode())
        return a.value, b.value, t[0], t[1]

    def stats(self):
        seen, dev, host, wrong = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        t = (C.c_double * 3)()
        self._lib.mdg_bam_stream_stats(self._stream, C.byref(seen), C.byref(dev), C.byref(host), C.byref(wrong), t)
        return {"records_seen": seen.value, "blocks_on_device": dev.value, "blocks_redone_on_host": host.value,
                "segment_guesses_corrected": wrong.value, "read_s": t[0], "decode_s": t[1], "wait_s": t[2]}

    def close(self):
        if self._stream:
            self._lib.mdg_bam_stream_close(self._stream)
            self._stream = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class BamWriter:
    """Writes the records of batches read with ``keep_raw`` (``rescale.py:299,344``)."""

    def __init__(self, path, header, threads=0, level=1):
        self._lib = _native.load()
        self._writer = C.c_void_p()
        text = "".join(line + "\n" for line in header.lines).encode()
        names = (C.c_char_p * max(1, len(header.references)))(*[n.encode() for n in header.references])
        lengths = np.array(header.lengths, dtype=np.uint32)
        code = self._lib.mdg_bam_create(str(path).encode(), text, names, lengths.ctypes.data, len(header.references),
                                        threads, level, C.byref(self._writer))
        if code < 0:
            raise OSError((self._lib.mdg_bam_writer_error(None) or b"").decode())

    def write(self, batch, status=None, qual=None, mr=None):
        """Appends every record of ``batch``; where ``status`` is set the record gets ``qual`` and ``MR``."""
        code = self._lib.mdg_bam_write_batch(
            self._writer, batch.raw.ctypes.data, batch.raw_off.ctypes.data, batch.n,
            None if status is None else np.ascontiguousarray(status, np.uint8).ctypes.data,
            None if qual is None else qual.ctypes.data, batch.base_off.ctypes.data,
            None if mr is None else np.ascontiguousarray(mr, np.float32).ctypes.data)
        if code < 0:
            raise OSError((self._lib.mdg_bam_writer_error(self._writer) or b"").decode())

    def write_soa(self, batch, first_index=0, name_prefix="r", read_groups=None):
        """Encodes ``batch`` (a :class:`ReadBatch`) as records; ``read_groups[lib]`` becomes the ``RG`` tag."""
        from .engine import batch_struct

        s = batch_struct(batch)
        groups = None
        if read_groups:
            groups = (C.c_char_p * len(read_groups))(*[g.encode() for g in read_groups])
        code = self._lib.mdg_bam_write_soa(self._writer, C.byref(s), first_index, name_prefix.encode(), groups,
                                           len(read_groups) if read_groups else 0)
        if code < 0:
            raise OSError((self._lib.mdg_bam_writer_error(self._writer) or b"").decode())

    def close(self):
        if self._writer:
            code = self._lib.mdg_bam_finish(self._writer)
            message = (self._lib.mdg_bam_writer_error(self._writer) or b"").decode()
            self._lib.mdg_bam_writer_free(self._writer)
            self._writer = C.c_void_p()
            if code < 0:
                raise OSError(message)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
