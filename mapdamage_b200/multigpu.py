"""Sharding of the counting pass over GPUs.

Reads are independent and the tables are sums (``main.py:165-217``), so rank r
of P counts a contiguous range of the reads and one ``ncclAllReduce(sum, u64)``
over the count slabs ends the pass (``mdg_allreduce_tables``).  The launcher
(``torch.distributed`` here) only carries the 128-byte NCCL unique id from rank
0 to the others; the collective itself runs inside the library, on the
engine's compute stream.
"""
import numpy as np


def shard_bounds(n_reads, rank, world):
    """``[start, stop)`` of rank ``rank``: contiguous, sizes differ by at most one read."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(n_reads), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    """The records of ``batch`` rank ``rank`` counts."""
    start, stop = shard_bounds(batch.n, rank, world)
    return batch.slice(start, stop)


def connect(engine, dist, rank=None, world=None):
    """Joins ``engine`` to the NCCL communicator of an initialised ``torch.distributed`` group."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    box = [type(engine).nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.nccl_init(box[0], rank, world)


def sum_tables_host(dist, tables):
    """All-reduce of host tables through ``torch.distributed`` (any backend): the reference
    semantics of ``mdg_allreduce_tables`` for tests that run without NCCL."""
    import torch

    on_gpu = dist.get_backend() == "nccl"  # NCCL moves device tensors only
    out = []
    for table in tables:
        t = torch.from_numpy(np.ascontiguousarray(table).astype(np.int64))
        if on_gpu:
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out.append(t.cpu().numpy().astype(np.uint64))
    return tuple(out)
