"""Python face of the CUDA engine: one :class:`DamageEngine` per GPU.

This is the object the host-side mirrors of the reference's loops drive
(:mod:`mapdamage_b200.counting` for ``main.py:165-231`` and
:mod:`mapdamage_b200.rescale` for ``rescale.py:285-365``).  It only marshals
numpy arrays across the C ABI of ``include/mapdamage_b200.h``; all per-read
work happens in the sm_100a kernels.  There is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _native
from .batch import ReadBatch

_FIELDS = ("flag", "tid", "pos", "lib", "l_seq", "base_off", "cigar_off", "cigar", "seq4",
           "tlen", "mtid", "mpos")


def batch_struct(batch, compact=False):
    """``mdg_batch`` view of a :class:`ReadBatch` (no copies).  ``compact`` passes NULL for the
    optional arrays that only hold their default (``ReadBatch.droppable``)."""
    n_bases = batch.total_bases
    if batch.seq4.shape[0] * 2 < n_bases:
        raise ValueError("seq4 shorter than the batch's base span")
    qual = batch.qual
    if qual is not None and qual.shape[0] < n_bases:
        # the library copies n_bases quality bytes; pad the odd tail slot
        qual = np.concatenate([qual, np.full(n_bases - qual.shape[0], 0xFF, np.uint8)])
        batch.qual = qual
    s = _native.Batch()
    s.n_reads = batch.n
    s.n_cigar = int(batch.cigar.shape[0])
    s.n_bases = n_bases
    drop = batch.droppable() if compact else ()
    for name in _FIELDS:
        setattr(s, name, None if name in drop and name != "cigar" else getattr(batch, name).ctypes.data)
    if "cigar" in drop:
        s.n_cigar = 1  # every read has the CIGAR word cigar[0] (mdg_batch: cigar_off NULL, n_cigar == 1)
    s.qual = None if qual is None else qual.ctypes.data
    return s


def bind_host_to_device(device):
    """Pins the calling process to the CPUs of the NUMA node GPU ``device`` is attached to, so that pinned
    batches (first touch) and the copy engine's reads stay on that node.  Returns the CPU set, or ``None``
    when the topology cannot be read.  Call before allocating pinned memory; one process per GPU."""
    import os

    buf = C.create_string_buffer(64)
    if _native.load().mdg_device_pci_bus_id(device, buf, len(buf)) < 0:
        return None
    path = "/sys/bus/pci/devices/%s/local_cpulist" % buf.value.decode().lower()
    try:
        text = open(path).read().strip()
        cpus = set()
        for part in text.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except (OSError, ValueError):
        return None


class PinnedArena:
    """numpy arrays carved out of page-locked host memory (``mdg_host_alloc``)."""

    def __init__(self, lib):
        self._lib = lib
        self._blocks = []

    def empty(self, shape, dtype):
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = max(1, count * dtype.itemsize)
        ptr = self._lib.mdg_host_alloc(nbytes)
        if not ptr:
            raise MemoryError("cudaHostAlloc(%d bytes) failed" % nbytes)
        self._blocks.append(ptr)
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def copy(self, array):
        out = self.empty(array.shape, array.dtype)
        out[...] = array
        return out

    def batch(self, batch):
        """Copy of ``batch`` whose arrays live in pinned memory."""
        arrays = {name: self.copy(getattr(batch, name)) for name in _FIELDS}
        arrays["qual"] = None if batch.qual is None else self.copy(batch.qual)
        return ReadBatch(**arrays)

    def close(self):
        for ptr in self._blocks:
            self._lib.mdg_host_free(ptr)
        self._blocks = []


class DeviceBatch:
    """A batch resident in HBM (``mdg_batch_upload``)."""

    def __init__(self, engine, handle, n_reads, has_qual=True):
        self.engine, self.handle, self.n, self.has_qual = engine, handle, n_reads, has_qual

    def free(self):
        if self.handle:
            self.engine._lib.mdg_batch_free(self.engine._ctx, self.handle)
            self.handle = None


class DamageEngine:
    def __init__(self, length=70, around=10, min_qual=0, n_libraries=1, lg_bins=8192, device=0,
                 n_slots=2, max_reads=1 << 20, max_cigar_ops=None, max_bases=None):
        self._lib = _native.load()
        self._ctx = C.c_void_p()
        self.length, self.around, self.min_qual = length, around, min_qual
        self.n_libraries, self.lg_bins, self.device = n_libraries, lg_bins, device
        self.max_reads = max_reads
        self.max_cigar_ops = max_cigar_ops if max_cigar_ops is not None else 4 * max_reads
        self.max_bases = max_bases if max_bases is not None else 160 * max_reads
        cfg = _native.Config(device, length, around, min_qual, n_libraries, lg_bins, n_slots, 0,
                             max_reads, self.max_cigar_ops, self.max_bases)
        code = self._lib.mdg_create(C.byref(self._ctx), C.byref(cfg))
        if code < 0:
            raise _native.NativeError(code, _native.last_error(None))
        self._keepalive = []
        self._arena = None

    # -- life cycle ------------------------------------------------------
    def close(self):
        if self._ctx:
            self._lib.mdg_destroy(self._ctx)
            self._ctx = C.c_void_p()
        if self._arena is not None:
            self._arena.close()
            self._arena = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code):
        return _native.check(code, self._ctx)

    @property
    def arena(self):
        if self._arena is None:
            self._arena = PinnedArena(self._lib)
        return self._arena

    # -- inputs ----------------------------------------------------------
    def set_reference(self, reference):
        packed, offsets, lengths = reference.packed()
        self._check(self._lib.mdg_set_reference(
            self._ctx, packed.ctypes.data, packed.shape[0], offsets.ctypes.data,
            lengths.ctypes.data, len(reference.names)))

    def synth_reference(self, lengths, seed=1, names=None):
        """A random genome made on the device (``mdg_synth_reference``).  Returns the contig names and lengths."""
        lens = np.array(lengths, dtype=np.uint32)
        self._check(self._lib.mdg_synth_reference(self._ctx, lens.ctypes.data, len(lens), int(seed)))
        self._synth_lengths = [int(x) for x in lens]
        return list(names or ["chr%d" % (i + 1) for i in range(len(lens))]), self._synth_lengths

    def reference_host(self, names, lengths):
        """The genome the device holds as a host :class:`~mapdamage_b200.refgenome.Reference` (for the checker)."""
        from .refgenome import Reference

        total = sum((n + 7) // 8 * 8 for n in lengths)
        image = np.empty(total // 2, dtype=np.uint8)
        self._check(self._lib.mdg_reference_download(self._ctx, image.ctypes.data, image.shape[0]))
        letter = np.full(16, ord("N"), dtype=np.uint8)
        letter[[1, 2, 4, 8]] = np.frombuffer(b"ACGT", dtype=np.uint8)
        bases = np.empty(total, dtype=np.uint8)
        bases[0::2] = letter[image & 15]
        bases[1::2] = letter[image >> 4]
        sequences, at = [], 0
        for n in lengths:
            sequences.append(bases[at:at + n])
            at += (n + 7) // 8 * 8
        return Reference(names, sequences)

    def genome_composition(self):
        """``[A, C, G, T]`` counts over the uploaded genome (``mdg_genome_composition``)."""
        counts = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.mdg_genome_composition(self._ctx, counts.ctypes.data))
        return [int(x) for x in counts]

    def upload(self, batch):
        handle = C.c_void_p()
        s = batch_struct(batch)
        self._check(self._lib.mdg_batch_upload(self._ctx, C.byref(s), C.byref(handle)))
        return DeviceBatch(self, handle, batch.n, has_qual=batch.qual is not None)

    def synth_batch(self, n_reads, seed=1, length=(100, 100), mix=(1, 0, 0, 0), paired=False,
                    with_qual=True, n_libs=1, error_rate=0.002, read_n_rate=0.0, filtered_rate=0.0,
                    damage0=0.3, damage_decay=0.7, sorted_positions=False):
        """Seeded synthetic aDNA batch generated directly in HBM (``mdg_synth_batch``)."""
        if isinstance(length, int):
            length = (length, length)
        params = _native.SynthParams(
            seed, n_reads, length[0], length[1], (C.c_int32 * 4)(*mix), int(paired), int(with_qual),
            n_libs, int(bool(sorted_positions)), error_rate, read_n_rate, filtered_rate, damage0, damage_decay, 0.0)
        handle = C.c_void_p()
        self._check(self._lib.mdg_synth_batch(self._ctx, C.byref(params), C.byref(handle)))
        return DeviceBatch(self, handle, n_reads, has_qual=bool(with_qual))

    def download(self, device_batch, pinned=False):
        """Host :class:`ReadBatch` copy of a resident batch (pinned memory on request)."""
        n, n_cigar, n_bases = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._lib.mdg_batch_sizes(self._ctx, device_batch.handle, C.byref(n), C.byref(n_cigar),
                                              C.byref(n_bases)))
        n, n_cigar, n_bases = n.value, n_cigar.value, n_bases.value
        empty = self.arena.empty if pinned else (lambda shape, dtype: np.empty(shape, dtype))
        arrays = {name: empty(n, dtype) for name, dtype in ReadBatch.FIELDS if name != "cigar_off"}
        arrays["cigar_off"] = empty(n + 1, np.uint32)
        arrays["cigar"] = empty(n_cigar, np.uint32)
        arrays["seq4"] = empty(n_bases // 2, np.uint8)
        arrays["qual"] = empty(n_bases, np.uint8) if device_batch.has_qual else None
        s = _native.Batch()
        s.n_reads, s.n_cigar, s.n_bases = n, n_cigar, n_bases
        for name in _FIELDS:
            setattr(s, name, arrays[name].ctypes.data)
        s.qual = None if arrays["qual"] is None else arrays["qual"].ctypes.data
        self._check(self._lib.mdg_batch_download(self._ctx, device_batch.handle, C.byref(s)))
        return ReadBatch(**arrays)

    def h2d_bytes(self, batch, rescale=False, compact=True):
        """Bytes ``count`` (or ``rescale``) copies to the device for ``batch`` (see ``copy_batch``)."""
        n = batch.n
        drop = batch.droppable() if compact else ()
        total = n * (2 + 4) + (4 if "cigar" in drop else batch.cigar.nbytes) + batch.total_bases // 2  # flag, pos, cigar, seq4
        total += sum(size for name, size in (("tid", 4 * n), ("l_seq", 4 * n), ("lib", 2 * n), ("tlen", 4 * n),
                                             ("base_off", 4 * n), ("cigar_off", 4 * (n + 1))) if name not in drop)
        if rescale:
            total += sum(4 * n for name in ("mtid", "mpos") if name not in drop)
        if batch.qual is not None and (rescale or self.min_qual > 0):
            total += batch.total_bases
        return total

    def fits(self, batch):
        return (batch.n <= self.max_reads and batch.cigar.shape[0] <= self.max_cigar_ops
                and batch.total_bases <= self.max_bases)

    # -- counting pass ---------------------------------------------------
    def count(self, batch, compact=True):
        """Queues one host batch (async copy + kernels); see :meth:`sync`.  With ``compact`` the optional
        arrays that only hold their default value stay on the host (``ReadBatch.droppable``)."""
        s = batch_struct(batch, compact)
        self._keepalive.append((batch, s))
        self._check(self._lib.mdg_count_submit(self._ctx, C.byref(s)))

    def count_resident(self, device_batch):
        self._check(self._lib.mdg_count_resident(self._ctx, device_batch.handle))

    def sync(self):
        try:
            self._check(self._lib.mdg_sync(self._ctx))
        finally:
            self._keepalive = []

    def reset(self):
        self._check(self._lib.mdg_reset_tables(self._ctx))
        self._keepalive = []

    def tables(self):
        """``(misincorp, dnacomp, lghist)`` uint64 slabs (layouts: mapdamage_b200.h)."""
        L, A, nl = self.length, self.around, self.n_libraries
        mis = np.zeros((nl, 2, 2, _native.N_CLASSES, L), dtype=np.uint64)
        comp = np.zeros((nl, 2, 2, 4, L + A), dtype=np.uint64)
        lg = np.zeros((nl, 2, 2, self.lg_bins), dtype=np.uint64)
        self._check(self._lib.mdg_fetch_tables(self._ctx, mis.ctypes.data, comp.ctypes.data, lg.ctypes.data))
        self._keepalive = []
        return mis, comp, lg

    def lg_overflow(self):
        """Fragment lengths beyond ``lg_bins`` as ``(lib, kind, strand, length, count)`` rows."""
        n = self._check(self._lib.mdg_fetch_lg_overflow(self._ctx, None, 0))
        if not n:
            return []
        rows = np.zeros((n, 4), dtype=np.int32)
        self._check(self._lib.mdg_fetch_lg_overflow(self._ctx, rows.ctypes.data, n))
        uniq, counts = np.unique(rows, axis=0, return_counts=True)
        return [tuple(int(x) for x in row) + (int(c),) for row, c in zip(uniq, counts)]

    # -- rescale pass ----------------------------------------------------
    def set_rescale_model(self, model):
        lut = np.ascontiguousarray(model.lut, dtype=np.uint8)
        inc = np.ascontiguousarray(model.inc, dtype=np.float64)
        self._check(self._lib.mdg_set_rescale_model(
            self._ctx, lut.ctypes.data, inc.ctypes.data, model.len5p, model.len3p))

    def rescale(self, batch, out=None, compact=True):
        """Queues the rescale of one batch; returns ``(qual_out, mr, status)`` arrays
        that are valid after :meth:`sync`."""
        s = batch_struct(batch, compact)
        if out is None:
            out = (np.empty(max(1, s.n_bases), dtype=np.uint8), np.empty(max(1, batch.n), dtype=np.float32),
                   np.empty(max(1, batch.n), dtype=np.uint8))
        qual_out, mr, status = out
        self._keepalive.append((batch, s, out))
        self._check(self._lib.mdg_rescale_submit(
            self._ctx, C.byref(s), qual_out.ctypes.data, mr.ctypes.data, status.ctypes.data))
        return qual_out[:s.n_bases], mr[:batch.n], status[:batch.n]

    def rescale_sparse(self, batch, out=None, compact=True):
        """Queues the rescale of one batch; only what changed comes back (``mdg_rescale_submit_sparse``).
        Returns ``(mr, status, ticket)``; after :meth:`rescale_collect` ``(ticket)`` the batch's own ``qual`` array has
        been patched in place and ``mr`` / ``status`` are valid."""
        s = batch_struct(batch, compact)
        if out is None:
            out = (np.empty(max(1, batch.n), dtype=np.float32), np.empty(max(1, batch.n), dtype=np.uint8))
        mr, status = out
        ticket = C.c_int32(-1)
        self._keepalive.append((batch, s, out))
        self._check(self._lib.mdg_rescale_submit_sparse(self._ctx, C.byref(s), mr.ctypes.data, status.ctypes.data,
                                                        C.byref(ticket)))
        return mr[:batch.n], status[:batch.n], ticket.value

    def rescale_collect(self, ticket, batch, scratch=None, apply=True):
        """Waits for the sparse submit ``ticket`` and (``apply``) writes the changed quality bytes into ``batch.qual``.
        Returns the number of bytes that changed; ``scratch[0][:n]`` / ``scratch[1][:n]`` hold their indices and scores."""
        cap = batch.total_bases // 4 + 4096
        if scratch is None or scratch[0].shape[0] < cap:
            scratch = (np.empty(cap, dtype=np.uint32), np.empty(cap, dtype=np.uint8))
        at, q = scratch
        n = self._check(self._lib.mdg_rescale_collect(self._ctx, ticket, at.ctypes.data, q.ctypes.data, at.shape[0],
                                                      batch.qual.ctypes.data if apply else None))
        return int(n)

    def rescale_resident(self, device_batch, want_results=False):
        """Rescales a resident batch in place (``mdg_rescale_resident``); with ``want_results`` returns ``(mr, status)``
        host arrays that are valid after :meth:`sync`."""
        handle = device_batch.handle
        if want_results:
            mr = np.empty(max(1, device_batch.n), dtype=np.float32)
            status = np.empty(max(1, device_batch.n), dtype=np.uint8)
            self._keepalive.append((mr, status))
            self._check(self._lib.mdg_rescale_resident(self._ctx, handle, mr.ctypes.data, status.ctypes.data))
            return mr[:device_batch.n], status[:device_batch.n]
        self._check(self._lib.mdg_rescale_resident(self._ctx, handle, None, None))
        return None

    def rescale_stats(self):
        stats = np.zeros(8, dtype=np.uint64)
        self._check(self._lib.mdg_fetch_rescale_stats(self._ctx, stats.ctypes.data))
        keys = ("pairs", "improper_pairs", "without_quals", "rescaled", "alignment_longer_than_read")
        return {k: int(v) for k, v in zip(keys, stats)}

    def rescale_hist(self, n_slots):
        """``(sub[2][n_slots][94], rev[2][94], ref_count[4])`` -- integer substitution bookkeeping."""
        sub = np.zeros((2, n_slots, 94), dtype=np.uint64)
        rev = np.zeros((2, 94), dtype=np.uint64)
        ref_count = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.mdg_fetch_rescale_hist(self._ctx, sub.ctypes.data, rev.ctypes.data,
                                                     ref_count.ctypes.data))
        return sub, rev, ref_count

    # -- multi-GPU -------------------------------------------------------
    @staticmethod
    def nccl_unique_id():
        buf = (C.c_uint8 * 128)()
        code = _native.load().mdg_nccl_unique_id(buf)
        if code < 0:
            raise _native.NativeError(code, _native.last_error(None))
        return bytes(buf)

    def nccl_init(self, unique_id, rank, n_ranks):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._lib.mdg_nccl_init(self._ctx, buf, rank, n_ranks))

    def allreduce_tables(self):
        self._check(self._lib.mdg_allreduce_tables(self._ctx))

    # -- measurement -----------------------------------------------------
    def event_record(self, which):
        self._check(self._lib.mdg_event_record(self._ctx, which))

    def event_elapsed_ms(self):
        ms = C.c_float()
        self._check(self._lib.mdg_event_elapsed_ms(self._ctx, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self._lib.mdg_launch_count(self._ctx))

    def kernel_ms(self):
        """Summed device time of the kernels launched since the last call."""
        ms = C.c_float()
        self._check(self._lib.mdg_last_kernel_ms(self._ctx, C.byref(ms)))
        return ms.value
