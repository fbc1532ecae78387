"""Multi-GPU plumbing: graphs are independent units, so a dataset shards by index with NO
collective on the data path.  One process per GPU (torchrun); rank r owns the
contiguous range [r*ceil(G/R), min(G,(r+1)*ceil(G/R))).  SENT randomness is keyed by the
GLOBAL graph id (GtokChoice.graph_id_base), so the token streams do not depend on R.

The only optional exchange is an all_gather of one (n_graphs, n_tokens) pair per rank, from
which every rank derives its base offset into the logically concatenated token stream.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .batch import GraphBatch


def shard_bounds(num_graphs: int, world_size: int) -> List[Tuple[int, int]]:
    per = -(-num_graphs // world_size) if world_size > 0 else num_graphs
    return [(min(num_graphs, r * per), min(num_graphs, (r + 1) * per)) for r in range(world_size)]


def local_shard(batch: GraphBatch, rank: int, world_size: int) -> Tuple[GraphBatch, int]:
    """-> (rank's sub-batch, global id of its first graph)."""
    g0, g1 = shard_bounds(batch.num_graphs, world_size)[rank]
    return batch.slice(g0, g1), g0


def global_offsets(n_graphs: int, n_tokens: int, group=None, device=None):
    """all_gather of (n_graphs, n_tokens) -> {'graph_base','token_base','graphs','tokens','per_rank'}.
    Works on any backend (nccl on GPUs, gloo in the CPU tests); single-process falls through."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return {"graph_base": 0, "token_base": 0, "graphs": n_graphs, "tokens": n_tokens,
                "per_rank": [(n_graphs, n_tokens)]}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.tensor([n_graphs, n_tokens], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    per = [(int(t[0]), int(t[1])) for t in allc]
    return {"graph_base": sum(p[0] for p in per[:rank]), "token_base": sum(p[1] for p in per[:rank]),
            "graphs": sum(p[0] for p in per), "tokens": sum(p[1] for p in per), "per_rank": per}
