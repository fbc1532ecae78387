#!/usr/bin/env python
"""Raw pinned-memory -> device copy rate per rank, alone and with all ranks copying at once; with and without binding
each rank to its GPU's NUMA node.  The end-to-end path is H2D-bound, so this is its ceiling at N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/probe_h2d.py [MiB]

Rank 0 prints one JSON line: per-rank GB/s alone / concurrent, aggregate, topology of every rank."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
import torch.distributed as dist
from gtok import numa

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
MIB = int(sys.argv[1]) if len(sys.argv) > 1 else 1024


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rate(host, devbuf, reps=5):
    devbuf.copy_(host, non_blocking=True)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        devbuf.copy_(host, non_blocking=True)
    b.record()
    barrier()
    return host.numel() * reps / (a.elapsed_time(b) / 1e3) / 1e9


def gather(v):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world == 1:
        return [v]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x) for x in out]


res = {"world": world, "MiB": MIB}
devbuf = torch.empty(MIB << 20, dtype=torch.uint8, device=dev)
for phase in ("default_affinity", "bound_to_gpu_node"):
    topo = numa.topology(local) if phase == "default_affinity" else numa.bind_to_gpu_node(local)
    host = torch.empty(MIB << 20, dtype=torch.uint8).pin_memory()
    host.fill_(1)                                             # first touch under the current affinity
    conc = rate(host, devbuf)
    alone = []
    for r in range(world):                                    # one rank at a time
        v = rate(host, devbuf) if r == rank else 0.0
        barrier()
        alone.append(v)
    res[phase] = {"concurrent_GBps_per_rank": gather(conc), "alone_GBps_per_rank": gather(alone[rank]),
                  "aggregate_concurrent_GBps": sum(gather(conc))}
    tl = [None] * world
    if world > 1:
        dist.all_gather_object(tl, topo)
    else:
        tl = [topo]
    res[phase]["topology"] = tl
    del host
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
