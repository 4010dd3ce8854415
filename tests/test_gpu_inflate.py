"""BGZF / DEFLATE blocks inflated on the GPU (``mdg_inflate_blocks``, one thread per block) against zlib."""
import ctypes as C
import random
import zlib

import numpy as np
import pytest

from mapdamage_b200 import _native

pytestmark = pytest.mark.gpu


def deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return c.compress(data) + c.flush()


def payloads():
    rng = random.Random(11)
    for n in (0, 1, 7, 300, 5000, 65280):
        yield bytes(rng.getrandbits(8) for _ in range(n))
        yield bytes(rng.choice(b"ACGT") for _ in range(n))
        yield b"A" * n
        yield (b"abcdefg" * (n // 7 + 1))[:n]
        yield bytes(min(255, int(rng.expovariate(0.05))) for _ in range(n))


class Inflater:
    def __init__(self, device=0):
        self.lib = _native.load()
        self.handle = C.c_void_p()
        rc = self.lib.mdg_inflater_create(device, C.byref(self.handle))
        assert rc == 0, rc

    def close(self):
        self.lib.mdg_inflater_free(self.handle)

    def run(self, streams, sizes):
        n = len(streams)
        in_len = np.array([len(s) for s in streams], dtype=np.uint32)
        in_off = np.zeros(n, dtype=np.uint64)
        in_off[1:] = np.cumsum(in_len[:-1].astype(np.uint64))
        isize = np.array(sizes, dtype=np.uint32)
        out_off = np.zeros(n, dtype=np.uint64)
        out_off[1:] = np.cumsum(isize[:-1].astype(np.uint64))
        blob = np.frombuffer(b"".join(streams) + b"\0", dtype=np.uint8).copy()
        out = np.zeros(int(isize.sum()) + 1, dtype=np.uint8)
        status = np.full(n, -1, dtype=np.int32)
        rc = self.lib.mdg_inflate_blocks(self.handle, blob.ctypes.data, len(blob) - 1, in_off.ctypes.data, in_len.ctypes.data,
                                         out.ctypes.data, len(out) - 1, out_off.ctypes.data, isize.ctypes.data, n,
                                         status.ctypes.data)
        assert rc == 0, self.lib.mdg_inflater_error(self.handle)
        return [bytes(out[int(o):int(o) + int(s)]) for o, s in zip(out_off, isize)], status


def test_blocks_come_out_like_zlib():
    datas, streams = [], []
    for data in payloads():
        for level, strategy in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY),
                                (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            datas.append(data)
            streams.append(deflate(data, level, strategy))
    inflater = Inflater()
    try:
        outs, status = inflater.run(streams, [len(d) for d in datas])
        assert not status.any()
        assert outs == datas
        # a second call with fewer blocks reuses the buffers
        outs, status = inflater.run(streams[:5], [len(d) for d in datas[:5]])
        assert not status.any() and outs == datas[:5]
    finally:
        inflater.close()


def test_damaged_blocks_are_reported_not_fatal():
    rng = random.Random(12)
    data = bytes(rng.choice(b"ACGTN") for _ in range(40_000))
    good = deflate(data, 6)
    streams, sizes = [], []
    for k in range(200):
        hurt = bytearray(good)
        if k % 2:
            hurt[rng.randrange(len(hurt))] ^= 1 << rng.randrange(8)
        streams.append(bytes(hurt))
        sizes.append(len(data))
    streams.append(good[:len(good) // 2])  # truncated
    sizes.append(len(data))
    streams.append(good)                   # claims one byte less than it holds
    sizes.append(len(data) - 1)
    inflater = Inflater()
    try:
        outs, status = inflater.run(streams, sizes)
    finally:
        inflater.close()
    for k in range(0, 200, 2):
        assert status[k] == 0 and outs[k] == data
    assert status[200] != 0 and status[201] != 0
    # a damaged stream either fails or yields bytes the CRC check of the caller will look at; it never hangs or crashes
    assert all(s in (0, 1) for s in status)
