"""N>1 host logic on CPU: world_size-2 gloo processes shard a dataset by index, exchange the one
optional (graphs, tokens) pair and agree on global offsets; shards tile the dataset exactly."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, q):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gtok import shard, synth
    gb = synth.er_graph_token(1001, seed=7)                  # same dataset on every rank
    sub, g0 = shard.local_shard(gb, rank, world)
    n, e = sub.num_nodes(), sub.num_entries()
    tokens = int(torch.clamp(3 * e + n + 5, max=600).sum())  # closed-form IBTT row lengths (host formula)
    info = shard.global_offsets(sub.num_graphs, tokens)
    total = int(torch.clamp(3 * gb.num_entries() + gb.num_nodes() + 5, max=600).sum())
    q.put((rank, g0, sub.num_graphs, tokens, info, total, int(sub.src.numel()), int(gb.edge_off[g0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_offsets():
    from gtok import shard
    assert shard.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, g00, n0, t0, i0, total, e0, eb0), (r1, g01, n1, t1, i1, _, e1, eb1) = res
    assert (g00, g01) == (0, 501) and n0 + n1 == 1001
    assert i0["graphs"] == i1["graphs"] == 1001 and i0["tokens"] == i1["tokens"] == t0 + t1 == total
    assert i0["token_base"] == 0 and i1["token_base"] == t0 and i1["graph_base"] == n0
    assert eb0 == 0 and eb1 == e0                            # shards are contiguous slices of the flat arrays
