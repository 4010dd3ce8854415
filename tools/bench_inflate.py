#!/usr/bin/env python
"""BGZF-sized DEFLATE blocks of BAM-like data: GPU (mdg_inflate_blocks, host buffers in and out) against the host
decoders on all cores."""
import argparse
import ctypes as C
import json
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from mapdamage_b200 import _native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=4000)
    ap.add_argument("--distinct", type=int, default=200)
    args = ap.parse_args()
    lib = _native.load()
    rng = np.random.default_rng(5)
    # BAM-like payload: per 220-byte record a 36-byte compressible head, 50 random bytes (packed bases), 100 qualities
    # uniform in 2..40 and a short name
    def payload():
        n = 65280 // 220
        rec = np.zeros((n, 220), dtype=np.uint8)
        rec[:, :36] = rng.integers(0, 4, (n, 36), dtype=np.uint8)
        rec[:, 36:70] = rng.integers(48, 58, (n, 34), dtype=np.uint8)
        rec[:, 70:120] = rng.integers(0, 256, (n, 50), dtype=np.uint8)
        rec[:, 120:] = rng.integers(2, 41, (n, 100), dtype=np.uint8)
        return rec.tobytes()
    distinct = [payload() for _ in range(args.distinct)]
    packed = [zlib.compressobj(1, zlib.DEFLATED, -15).compress(d) for d in distinct]
    packed = []
    for d in distinct:
        c = zlib.compressobj(1, zlib.DEFLATED, -15)
        packed.append(c.compress(d) + c.flush())
    order = rng.integers(0, args.distinct, args.blocks)
    streams = [packed[i] for i in order]
    sizes = np.array([len(distinct[i]) for i in order], dtype=np.uint32)
    in_len = np.array([len(s) for s in streams], dtype=np.uint32)
    in_off = np.zeros(args.blocks, dtype=np.uint64)
    in_off[1:] = np.cumsum(in_len[:-1].astype(np.uint64))
    out_off = np.zeros(args.blocks, dtype=np.uint64)
    out_off[1:] = np.cumsum(sizes[:-1].astype(np.uint64))
    blob = np.frombuffer(b"".join(streams), dtype=np.uint8).copy()
    out = np.zeros(int(sizes.sum()) + 16, dtype=np.uint8)
    status = np.zeros(args.blocks, dtype=np.int32)
    handle = C.c_void_p()
    assert lib.mdg_inflater_create(0, C.byref(handle)) == 0
    times = []
    for _ in range(4):
        t0 = time.perf_counter()
        rc = lib.mdg_inflate_blocks(handle, blob.ctypes.data, len(blob), in_off.ctypes.data, in_len.ctypes.data,
                                    out.ctypes.data, int(sizes.sum()), out_off.ctypes.data, sizes.ctypes.data,
                                    args.blocks, status.ctypes.data)
        times.append(time.perf_counter() - t0)
        assert rc == 0 and not status.any()
    k = int(order[7])
    assert bytes(out[int(out_off[7]):int(out_off[7]) + int(sizes[7])]) == distinct[k]
    lib.mdg_inflater_free(handle)

    def host(i):
        buf = (C.c_uint8 * int(sizes[i])).from_buffer(out, int(out_off[i]))
        return lib.mdg_inflate_raw(streams[i], len(streams[i]), buf, int(sizes[i]))
    threads = os.cpu_count()
    with ThreadPoolExecutor(threads) as pool:
        t0 = time.perf_counter()
        got = list(pool.map(host, range(args.blocks), chunksize=16))
        t_host = time.perf_counter() - t0
    assert all(g == s for g, s in zip(got, sizes))
    total = float(sizes.sum())
    print(json.dumps({"blocks": args.blocks, "inflated_MB": total / 1e6, "compressed_MB": len(blob) / 1e6,
                      "gpu_call_s": times, "gpu_GBps_best": total / min(times) / 1e9,
                      "host_threads": threads, "host_s": t_host, "host_GBps": total / t_host / 1e9}))


if __name__ == "__main__":
    main()
