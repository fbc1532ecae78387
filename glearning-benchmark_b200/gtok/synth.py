"""Synthetic graph workloads of the shapes BASELINE.json names.

No datasets are reachable offline, so every benchmark / parity input is synthetic:

* ``er_graph_token``  - graph-token-shaped Erdos-Renyi graphs: size bucket
  small[10,20) / medium[20,40) / large[40,50) uniform, N uniform in bucket,
  p ~ U(0.1, 0.2) (reference docs/synthetic_data.md:93-94, :133-137); each
  undirected edge listed ONCE with u < v, u-major ascending (the order the
  graph-token JSON holds and graph_token_dataset_autograph.py:378-382 keeps).
* ``zinc_shaped``     - molecule-shaped graphs: N ~ clip(round(N(23.2, 4.3)), 9, 37),
  random tree + 1..3 ring closures, edge_index in BOTH directions sorted by
  (src, dst) as PyG's ZINC stores it, atom types in [0, 28) skewed to C/N/O (ids
  >= 9 exercise the 'X' path, zinc_dataset_indexbase.py:104), bond types mostly
  1..3 with rare 4 and rare out-of-range (the 'unknown' path, :138-139).
* ``er_sparse``       - fixed-N graphs with mean degree ~8 for the large-graph sweep.

Generators are vectorised torch code and run on whichever device they are asked
to (input synthesis is plumbing, not the hot path).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .batch import GraphBatch


def _offsets(counts: torch.Tensor) -> torch.Tensor:
    off = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    off[1:] = counts.to(torch.int64).cumsum(0)
    return off


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def er_sizes(G: int, gen: torch.Generator, device) -> Tuple[torch.Tensor, torch.Tensor]:
    bucket = torch.randint(0, 3, (G,), generator=gen, device=device)
    lo = torch.tensor([10, 20, 40], device=device)[bucket]
    hi = torch.tensor([20, 40, 50], device=device)[bucket]
    u = torch.rand(G, generator=gen, device=device)
    n = (lo + (u * (hi - lo)).floor().to(torch.int64)).clamp_max(49)
    p = 0.1 + 0.1 * torch.rand(G, generator=gen, device=device)
    return n, p


def er_graph_token(G: int, seed: int = 0, device="cpu", task: str = "cycle_check",
                   chunk: int = 65536) -> GraphBatch:
    gen = _gen(seed, device)
    n, p = er_sizes(G, gen, device)
    NM = 49
    ar = torch.arange(NM, device=device)
    upper = (ar[:, None] < ar[None, :])
    gs, us, vs = [], [], []
    for c0 in range(0, G, chunk):
        nc, pc = n[c0:c0 + chunk], p[c0:c0 + chunk]
        r = torch.rand(nc.numel(), NM, NM, generator=gen, device=device)
        keep = (r < pc[:, None, None]) & upper[None] & (ar[None, None, :] < nc[:, None, None])
        nz = keep.nonzero()                       # lexicographic (g, u, v): u-major ascending
        gs.append(nz[:, 0] + c0)
        us.append(nz[:, 1])
        vs.append(nz[:, 2])
        del r, keep
    g_of_e = torch.cat(gs) if gs else torch.zeros(0, dtype=torch.int64, device=device)
    src = torch.cat(us).to(torch.int32) if us else torch.zeros(0, dtype=torch.int32, device=device)
    dst = torch.cat(vs).to(torch.int32) if vs else torch.zeros(0, dtype=torch.int32, device=device)
    ecount = torch.bincount(g_of_e, minlength=G)
    qu = qv = None
    if task == "shortest_path":
        qu = (torch.rand(G, generator=gen, device=device) * n).floor().to(torch.int32)
        qv = (torch.rand(G, generator=gen, device=device) * n).floor().to(torch.int32)
    y = torch.randint(0, 2, (G,), generator=gen, device=device)
    return GraphBatch(_offsets(n), _offsets(ecount), src, dst, None, None, qu, qv, y)


def er_sparse(G: int, num_nodes: int, mean_degree: float = 8.0, seed: int = 0, device="cpu",
              both_directions: bool = False) -> GraphBatch:
    """G graphs of exactly ``num_nodes`` nodes, ~num_nodes*mean_degree/2 unique edges."""
    gen = _gen(seed, device)
    N = int(num_nodes)
    m = int(N * mean_degree / 2)
    u = torch.randint(0, N, (G, m), generator=gen, device=device)
    v = torch.randint(0, N, (G, m), generator=gen, device=device)
    a, b = torch.minimum(u, v), torch.maximum(u, v)
    g = torch.arange(G, device=device)[:, None].expand(G, m)
    ok = a != b
    key = (g[ok] * N + a[ok]) * N + b[ok]
    key = torch.unique(key)                        # sorted (g, a, b)
    gg = key // (N * N)
    a = (key // N) % N
    b = key % N
    if both_directions:
        k2 = torch.cat([key, (gg * N + b) * N + a]).sort().values
        gg = k2 // (N * N)
        a = (k2 // N) % N
        b = k2 % N
    ecount = torch.bincount(gg, minlength=G)
    n = torch.full((G,), N, dtype=torch.int64, device=device)
    return GraphBatch(_offsets(n), _offsets(ecount), a.to(torch.int32), b.to(torch.int32))


def zinc_shaped(G: int, seed: int = 0, device="cpu", nmin: int = 9, nmax: int = 37) -> GraphBatch:
    gen = _gen(seed, device)
    n = (23.2 + 4.3 * torch.randn(G, generator=gen, device=device)).round().clamp(nmin, nmax).to(torch.int64)
    node_off = _offsets(n)
    Ntot = int(node_off[-1])
    gid = torch.repeat_interleave(torch.arange(G, device=device), n)
    local = torch.arange(Ntot, device=device) - node_off[gid]
    # random tree: parent(i) uniform in [0, i)
    par = (torch.rand(Ntot, generator=gen, device=device) * local).floor().to(torch.int64)
    par = torch.minimum(par, (local - 1).clamp_min(0))
    tsel = local > 0
    tg, ta, tb = gid[tsel], par[tsel], local[tsel]
    # 1..3 ring closures per graph
    k = torch.randint(1, 4, (G,), generator=gen, device=device)
    rg = torch.repeat_interleave(torch.arange(G, device=device), k)
    ru = (torch.rand(rg.numel(), generator=gen, device=device) * n[rg]).floor().to(torch.int64)
    rv = (torch.rand(rg.numel(), generator=gen, device=device) * n[rg]).floor().to(torch.int64)
    ok = ru != rv
    rg, ru, rv = rg[ok], ru[ok], rv[ok]
    gg = torch.cat([tg, rg])
    a = torch.cat([ta, torch.minimum(ru, rv)])
    b = torch.cat([tb, torch.maximum(ru, rv)])
    M = 64
    key = torch.unique((gg * M + a) * M + b)       # unique undirected edges
    r = torch.rand(key.numel(), generator=gen, device=device)
    attr = torch.ones_like(key)
    attr[r > 0.70] = 2
    attr[r > 0.90] = 3
    attr[r > 0.98] = 4
    attr[r > 0.99] = 0                             # out of range -> 'unknown'
    attr[r > 0.995] = 7
    gg = key // (M * M)
    a = (key // M) % M
    b = key % M
    k2 = torch.cat([key, (gg * M + b) * M + a])
    k2, perm = k2.sort()
    attr2 = torch.cat([attr, attr])[perm]
    g2 = k2 // (M * M)
    src = ((k2 // M) % M).to(torch.int32)
    dst = (k2 % M).to(torch.int32)
    ecount = torch.bincount(g2, minlength=G)
    rx = torch.rand(Ntot, generator=gen, device=device)
    x = torch.zeros(Ntot, dtype=torch.int64, device=device)
    x[rx > 0.60] = 1
    x[rx > 0.75] = 2
    hi = rx > 0.90
    x[hi] = (3 + (rx[hi] - 0.90) * 250).floor().to(torch.int64).clamp_max(27)
    y = 2.0 * torch.randn(G, generator=gen, device=device)
    return GraphBatch(node_off, _offsets(ecount), src, dst, x.to(torch.int32), attr2.to(torch.int32),
                      None, None, y.to(torch.float32))


# ----------------------------------------------------------------------------
# SURVEY.md 8d "Config 1": the CPU-runnable fixture, numpy.random.default_rng(0)
# ----------------------------------------------------------------------------
def er_config1_numpy(G: int = 1000, seed: int = 0, task: str = "cycle_check"):
    """-> list of dicts {n, edges[(u,v)], label(str), query}.  Plain loops; small G."""
    rng = np.random.default_rng(seed)
    ranges = [(10, 20), (20, 40), (40, 50)]
    out = []
    for _ in range(G):
        lo, hi = ranges[int(rng.integers(0, 3))]
        n = int(rng.integers(lo, hi))
        p = float(rng.uniform(0.1, 0.2))
        m = rng.random((n, n)) < p
        edges = [(u, v) for u in range(n) for v in range(u + 1, n) if m[u, v]]
        parent = list(range(n))

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a
        cyc = False
        for u, v in edges:
            ru, rv = find(u), find(v)
            if ru == rv:
                cyc = True
            else:
                parent[ru] = rv
        rec = {"n": n, "edges": edges}
        if task == "cycle_check":
            rec["label"] = "yes" if cyc else "no"
            rec["query"] = None
        else:
            u = int(rng.integers(0, n))
            v = (u + 1 + int(rng.integers(0, n - 1))) % n
            adj = [[] for _ in range(n)]
            for a, b in edges:
                adj[a].append(b)
                adj[b].append(a)
            dist = {u: 0}
            fr = [u]
            while fr:
                nf = []
                for a in fr:
                    for b in adj[a]:
                        if b not in dist:
                            dist[b] = dist[a] + 1
                            nf.append(b)
                fr = nf
            d = dist.get(v)
            rec["label"] = "INF" if d is None else f"len{d}"
            rec["query"] = (u, v)
        out.append(rec)
    return out
