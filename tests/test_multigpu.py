"""Multi-GPU path: sharding + table reduction.  CPU (gloo, world size 2): the host logic against the
oracle.  GPU (needs 2 devices): two engines, one NCCL all-reduce inside the library."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

from mapdamage_b200 import multigpu

ROOT = Path(__file__).resolve().parent.parent


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000, 12345):
        for world in (1, 2, 3, 8):
            bounds = [multigpu.shard_bounds(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multigpu.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_ranks(script, world, tmp_path, extra_env=None):
    path = tmp_path / "rank_script.py"
    path.write_text(textwrap.dedent(script))
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PYTHONPATH=os.pathsep.join([str(ROOT), str(ROOT / "oracle"), os.environ.get("PYTHONPATH", "")]))
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, str(path), str(tmp_path)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outputs = [p.communicate(timeout=600)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outputs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out)
    return outputs


def test_sharded_count_sums_to_the_whole_gloo(tmp_path):
    """world_size 2 over gloo: per-rank oracle tables of the shards, summed, equal the oracle's tables
    of the whole batch -- the contract mdg_allreduce_tables implements on the device."""
    script = """
        import sys
        import numpy as np
        import torch.distributed as dist
        import oracle
        from mapdamage_b200 import multigpu, synth

        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        reference = synth.make_reference([200_000, 50_000], seed=3)
        batch = synth.simulate_reads(reference, 30_001, seed=4, length=(40, 120), mix=(6, 1, 1, 2), paired=False,
                                     n_libs=2, threads=1)
        mine = multigpu.shard_batch(batch, rank, world)
        part = oracle.count(mine, reference, n_lib=2, lg_bins=512)
        total = multigpu.sum_tables_host(dist, part)
        whole = oracle.count(batch, reference, n_lib=2, lg_bins=512)
        for a, b in zip(total, whole):
            assert np.array_equal(a, b)
        assert sum(int(t.sum()) for t in part) < sum(int(t.sum()) for t in whole)
        dist.destroy_process_group()
        print("rank", rank, "ok", mine.n)
    """
    outputs = _run_ranks(script, 2, tmp_path)
    assert all("ok" in out for out in outputs)


@pytest.mark.gpu
def test_two_gpu_allreduce_nccl(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = """
        import os
        import numpy as np
        import torch
        import torch.distributed as dist
        import oracle
        from mapdamage_b200 import multigpu, synth
        from mapdamage_b200.engine import DamageEngine

        rank = int(os.environ["RANK"])
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
        world = dist.get_world_size()
        reference = synth.make_reference([200_000, 50_000], seed=3)
        batch = synth.simulate_reads(reference, 60_001, seed=4, length=(40, 120), mix=(6, 1, 1, 2), threads=1)
        mine = multigpu.shard_batch(batch, rank, world)
        with DamageEngine(device=rank, max_reads=mine.n, max_cigar_ops=mine.cigar.shape[0],
                          max_bases=mine.total_bases) as engine:
            engine.set_reference(reference)
            multigpu.connect(engine, dist)
            half = mine.n // 2
            engine.count(mine.slice(0, half))
            local = engine.tables()
            engine.allreduce_tables()
            first = engine.tables()
            engine.allreduce_tables()  # reducing twice must not add the sums to themselves
            again = engine.tables()
            for a, b in zip(first, again):
                assert np.array_equal(a, b)
            # the reduced tables are the sum of the per-rank tables taken before the collective
            for a, b in zip(first, multigpu.sum_tables_host(dist, local)):
                assert np.array_equal(a, b)
            engine.count(mine.slice(half, mine.n))  # count on, reduce again: sums over reads, not sums of sums
            engine.allreduce_tables()
            got = engine.tables()
        want = oracle.count(batch, reference, lg_bins=8192, threads=2)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        dist.barrier()
        dist.destroy_process_group()
        print("rank", rank, "ok")
    """
    outputs = _run_ranks(script, 2, tmp_path)
    assert all("ok" in out for out in outputs)
