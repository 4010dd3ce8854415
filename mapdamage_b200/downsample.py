"""Down-sampling of the kept reads, the way the reference's reader does it (``reader.py:84-96,134-164``).

``-n X`` with ``0 < X < 1`` keeps each read that passed the flag filter with probability ``X`` -- one
``random.Random(seed).random()`` per such read, in file order; ``X >= 1`` keeps ``int(X)`` reads by reservoir sampling
with ``randint(0, index)``.  The tables are sums over reads, so the order the reference emits the sample in
(sorted by position, ``reader.py:162``) does not matter; which reads are drawn does, and that is reproduced exactly:
the generator is seeded by CPython itself (``random.Random(seed).getstate()``) and its stream of draws is continued
natively (``mdg_sample_fraction`` / ``mdg_sample_reservoir``, ``csrc/mdg_sampler.cpp``).

The selection is applied to a batch by setting the QC-fail flag bit of the reads left out: the counting kernels'
own flag filter (``reader.py:121-132``) then skips them, nothing is copied.
"""
import random

import numpy as np

from . import _native

DROP_BIT = 0x200  # "failed QC": one of the bits reader.py:121-132 filters on


def _state_of(seed):
    version, words, _ = random.Random(seed).getstate()
    if version != 3 or len(words) != 625:
        raise RuntimeError("unexpected random.Random state layout")
    return np.array(words, dtype=np.uint32)


class FractionSampler:
    """``BAMReader._downsample_to_fraction`` (``reader.py:134-142``)."""

    def __init__(self, fraction, seed=None):
        if not (0 <= fraction < 1):
            raise ValueError(fraction)
        self.fraction = float(fraction)
        self._state = _state_of(seed)
        self._lib = _native.load()

    def mask(self, n):
        """Keep flags for the next ``n`` reads that passed the flag filter."""
        keep = np.empty(int(n), dtype=np.uint8)
        rc = self._lib.mdg_sample_fraction(self._state.ctypes.data, self.fraction, int(n), keep.ctypes.data)
        if rc:
            raise ValueError("mdg_sample_fraction failed (%d)" % rc)
        return keep.view(np.bool_)


class ReservoirSampler:
    """``BAMReader._downsample_to_fixed_number`` (``reader.py:144-164``): feed the stream, then ask what is left."""

    def __init__(self, count, seed=None):
        if count < 1:
            raise ValueError(count)
        self.count = int(count)
        self._state = _state_of(seed)
        self._slots = np.full(self.count, -1, dtype=np.int64)
        self._seen = 0
        self._lib = _native.load()

    def feed(self, n):
        """Walks the next ``n`` reads that passed the flag filter."""
        rc = self._lib.mdg_sample_reservoir(self._state.ctypes.data, self._seen, int(n), self.count,
                                            self._slots.ctypes.data)
        if rc:
            raise ValueError("mdg_sample_reservoir failed (%d)" % rc)
        self._seen += int(n)

    def selected(self):
        """Sorted stream indices of the reads in the reservoir."""
        return np.sort(self._slots[self._slots >= 0])


class Selection:
    """Keep flags by stream index, handed out batch by batch."""

    def __init__(self, indices):
        self._indices = np.asarray(indices, dtype=np.int64)
        self._seen = 0

    def mask(self, n):
        lo = np.searchsorted(self._indices, self._seen)
        hi = np.searchsorted(self._indices, self._seen + n)
        keep = np.zeros(int(n), dtype=np.bool_)
        keep[self._indices[lo:hi] - self._seen] = True
        self._seen += int(n)
        return keep


def sampler_for(downsample, seed=None):
    """``None``, a :class:`FractionSampler` or a :class:`ReservoirSampler`, by the rule of ``config.py:396-400`` /
    ``reader.py:84-96``."""
    if downsample is None:
        return None
    if downsample <= 0:
        raise ValueError("-n/--downsample must be a positive value")
    if downsample < 1:
        return FractionSampler(downsample, seed)
    return ReservoirSampler(int(downsample), seed)


def apply_mask(batch, passed, keep):
    """Marks the reads of ``batch`` that passed the filter (boolean ``passed``) but were not drawn (``keep`` over them)."""
    rows = np.flatnonzero(passed)
    batch.flag[rows[~keep]] |= DROP_BIT
    batch.invalidate()
