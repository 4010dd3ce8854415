"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (mdg_oracle.c).

May be imported by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs only.  The product package
(``mapdamage_b200``) never imports it.
"""
import ctypes as C
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

N_CLASSES = 30


class _Batch(C.Structure):
    _fields_ = [
        ("n_reads", C.c_int64), ("flag", C.c_void_p), ("tid", C.c_void_p), ("pos", C.c_void_p),
        ("lib", C.c_void_p), ("l_seq", C.c_void_p), ("base_off", C.c_void_p),
        ("cigar_off", C.c_void_p), ("cigar", C.c_void_p), ("seq4", C.c_void_p),
        ("qual", C.c_void_p), ("tlen", C.c_void_p), ("mtid", C.c_void_p), ("mpos", C.c_void_p),
    ]


class _Ref(C.Structure):
    _fields_ = [("n_contigs", C.c_int32), ("seq", C.POINTER(C.c_void_p)), ("len", C.c_void_p)]


class _Corr(C.Structure):
    _fields_ = [("max_pos", C.c_int32), ("ct", C.c_void_p), ("ga", C.c_void_p)]


class Subs(C.Structure):
    _fields_ = [
        ("hist", C.c_uint64 * 130 * 2 * 4), ("ref_count", C.c_uint64 * 4),
        ("pvals", C.c_double * 6), ("n_pairs", C.c_uint64), ("n_improper", C.c_uint64),
        ("n_without_quals", C.c_uint64), ("n_rescaled", C.c_uint64), ("n_too_long", C.c_uint64),
    ]


def build(force=False):
    """Compiles ``libmdg_oracle.so`` next to the source (gcc; no reference sources involved)."""
    so = _HERE / "libmdg_oracle.so"
    src = _HERE / "mdg_oracle.c"
    if force or not so.is_file() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(
            ["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", str(so), str(src), "-lm"],
            check=True,
        )
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.mdo_count.restype = C.c_int
        _LIB.mdo_rescale.restype = C.c_int
        _LIB.mdo_rescaled_phred.restype = C.c_int
        _LIB.mdo_rescaled_phred.argtypes = [C.c_int, C.c_double]
        _LIB.mdo_sizeof_subs.restype = C.c_size_t
        assert _LIB.mdo_sizeof_subs() == C.sizeof(Subs)
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data


def _batch_struct(batch):
    return _Batch(
        batch.n, _ptr(batch.flag), _ptr(batch.tid), _ptr(batch.pos), _ptr(batch.lib),
        _ptr(batch.l_seq), _ptr(batch.base_off), _ptr(batch.cigar_off), _ptr(batch.cigar),
        _ptr(batch.seq4), _ptr(batch.qual), _ptr(batch.tlen), _ptr(batch.mtid), _ptr(batch.mpos),
    )


def _ref_struct(reference):
    """ASCII contigs straight from the FASTA -- not the product's packed image."""
    n = len(reference.sequences)
    seqs = [np.ascontiguousarray(s, dtype=np.uint8) for s in reference.sequences]
    ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in seqs])
    lens = np.array([s.shape[0] for s in seqs], dtype=np.int64)
    ref = _Ref(n, ptrs, lens.ctypes.data)
    ref._keep = (seqs, ptrs, lens)
    return ref


def count(batch, reference, length=70, around=10, minqual=0, n_lib=1, lg_bins=4096, threads=1):
    """Counting pass (reference ``main.py:165-217``) -> (misincorp, dnacomp, lghist) slabs."""
    L, A = length, around
    ranges = np.linspace(0, batch.n, max(1, threads) + 1).astype(np.int64)
    bs, rs = _batch_struct(batch), _ref_struct(reference)

    def work(k):
        mis = np.zeros((n_lib, 2, 2, N_CLASSES, L), dtype=np.uint64)
        comp = np.zeros((n_lib, 2, 2, 4, L + A), dtype=np.uint64)
        lg = np.zeros((n_lib, 2, 2, lg_bins), dtype=np.uint64)
        rc = lib().mdo_count(
            C.byref(bs), C.byref(rs), C.c_int64(int(ranges[k])), C.c_int64(int(ranges[k + 1])),
            C.c_int(L), C.c_int(A), C.c_int(minqual), C.c_int(n_lib), C.c_int(lg_bins),
            C.c_void_p(mis.ctypes.data), C.c_void_p(comp.ctypes.data), C.c_void_p(lg.ctypes.data),
        )
        if rc:
            raise RuntimeError("oracle mdo_count failed: rc=%d" % rc)
        return mis, comp, lg

    if threads <= 1:
        return work(0)
    with ThreadPoolExecutor(threads) as pool:
        parts = list(pool.map(work, range(threads)))
    return tuple(sum(p[i] for p in parts) for i in range(3))


def corr_arrays(corr_prob, max_pos):
    """``{(ref, read, pos): p}`` (``rescale.py:23-46``) -> dense ct/ga arrays."""
    ct = np.zeros(2 * max_pos + 1, dtype=np.float64)
    ga = np.zeros(2 * max_pos + 1, dtype=np.float64)
    for (nt_ref, nt_seq, pos), value in corr_prob.items():
        if abs(pos) > max_pos:
            raise ValueError("position outside table")
        if (nt_ref, nt_seq) == ("C", "T"):
            ct[pos + max_pos] = value
        elif (nt_ref, nt_seq) == ("G", "A"):
            ga[pos + max_pos] = value
    return ct, ga


def rescale(batch, reference, corr_prob, max_pos=64):
    """Rescale pass (``rescale.py:285-365``) -> (qual_out, mr, status, Subs, rc)."""
    ct, ga = corr_arrays(corr_prob, max_pos)
    corr = _Corr(max_pos, ct.ctypes.data, ga.ctypes.data)
    bs, rs = _batch_struct(batch), _ref_struct(reference)
    qual_out = np.zeros_like(batch.qual) if batch.qual is not None else np.zeros(0, np.uint8)
    mr = np.zeros(batch.n, dtype=np.float32)
    status = np.zeros(batch.n, dtype=np.uint8)
    subs = Subs()
    rc = lib().mdo_rescale(
        C.byref(bs), C.byref(rs), C.c_int64(0), C.c_int64(batch.n), C.byref(corr),
        C.c_void_p(qual_out.ctypes.data), C.c_void_p(mr.ctypes.data),
        C.c_void_p(status.ctypes.data), C.byref(subs),
    )
    return qual_out, mr, status, subs, rc


def rescaled_phred(q, corr):
    return lib().mdo_rescaled_phred(int(q), float(corr))
