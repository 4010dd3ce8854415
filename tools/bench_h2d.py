#!/usr/bin/env python
"""Pinned host -> device copy bandwidth per rank with 1, 2, 4 or 8 ranks copying at once (where does
the end-to-end scaling go beyond two GPUs?).  Launch like bench.py:

    python -m torch.distributed.run --nproc-per-node N tools/bench_h2d.py [--bind 0|1] [--gb 2]

Every rank page-locks a buffer, optionally binds itself to the CPUs of its GPU's NUMA node first (the same
``bind_host_to_device`` bench.py uses), and times ``cudaMemcpyAsync`` H2D with CUDA events while all ranks copy.
Rank 0 prints one JSON line: per-rank GB/s, the aggregate, and what the node looks like (NUMA nodes, GPU affinity).
"""
import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", type=int, default=1)
    ap.add_argument("--gb", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpus = None
    if args.bind:
        from mapdamage_b200.engine import bind_host_to_device

        cpus = bind_host_to_device(local)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(args.gb * (1 << 30))
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.fill_(7)  # first touch on the bound node
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    back = torch.empty(n // 8, dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()

    def timed(fn):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    t_h2d = timed(lambda: dev.copy_(host, non_blocking=True))
    t_d2h = timed(lambda: back.copy_(dev[:n // 8], non_blocking=True))
    mine = torch.tensor([n * args.reps / t_h2d / 1e9, (n // 8) * args.reps / t_d2h / 1e9], dtype=torch.float64, device="cuda")
    rows = [torch.zeros_like(mine) for _ in range(world)]
    if dist is not None:
        dist.all_gather(rows, mine)
    else:
        rows = [mine]
    if rank == 0:
        def sh(cmd):
            try:
                return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
            except Exception as error:  # noqa: BLE001
                return "unavailable: %s" % error

        numa = {}
        for i in range(torch.cuda.device_count()):
            bus = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
            numa[i] = sh("nvidia-smi -i %d --query-gpu=pci.bus_id --format=csv,noheader" % i)
        h2d = [float(r[0]) for r in rows]
        print(json.dumps({
            "tool": "tools/bench_h2d.py", "ranks": world, "bound_to_gpu_numa_node": bool(args.bind),
            "cpus_of_rank0": None if cpus is None else len(cpus), "buffer_gb": args.gb, "reps": args.reps,
            "h2d_gbps_per_rank": h2d, "h2d_gbps_aggregate": sum(h2d), "h2d_gbps_min": min(h2d),
            "d2h_gbps_per_rank": [float(r[1]) for r in rows],
            "host_cpus": os.cpu_count(), "numa_nodes": sh("ls -d /sys/devices/system/node/node* | wc -l"),
            "gpu_numa": sh("for d in /sys/bus/pci/devices/*; do if [ -e $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null && "
                           "grep -q 0x0302 $d/class 2>/dev/null; then echo $(basename $d) $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done"),
            "topo": sh("nvidia-smi topo -m | head -14"), "mem": sh("free -g | head -2"),
        }))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
