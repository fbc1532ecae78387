"""CPU oracle for the steps AROUND the tokenization path.  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs - never by the product path.

* has_cycle_dfs           restates the reference's own label oracle, test/train_performance_test.py:27-67
                          (parent-tracking DFS on the UNDIRECTED multigraph built from an edge list that names
                          every edge once).  That module imports torch_geometric at import time (:22) and the
                          pre-refactor dataset module (:23), so it cannot be imported here; the
                          function is 30 lines of list logic and is restated statement for statement.
* shortest_distance       BFS hop count; the label the graph-token task generator (external alip67/graph-token,
                          not vendored) writes as `lenK` / `INF` and graph_data_loader/data_loader.py:19-41 parses
                          back (class = K - 1, INF dropped).
* first_token_pos         IBTT/AGTT `extract_query_nodes`: position of the FIRST occurrence of the <q> id in a
                          padded row (trainer/train_ibtt.py:89-103, trainer/train_agtt.py:101-112), -1 when absent.
* er_graph                the engine's own on-device Erdos-Renyi generator, defined HERE (it is synthetic data; the
                          reference's generator is networkx inside the external graph-token repo, driven by Python's
                          `random`, and is not reproducible from any seed we could share).  Distribution follows
                          docs/synthetic_data.md:93-94,130-138: size bucket uniform over {[10,20),[20,40),[40,50)},
                          N uniform in the bucket, p ~ U(0.1, 0.2), pairs u < v in ascending order.

Parity status: has_cycle_dfs and first_token_pos are PINNED to outputs of the reference itself
(tests/golden/tasks_labels.json, produced by tests/golden/make_tasks_golden.py, which lifts has_cycle_dfs out of the
reference file with `ast` and runs it unmodified on 400 seeded multigraph edge lists, and evaluates the trainer's
first-<q> expression with torch).  shortest_distance has no in-repo reference implementation (checked against
networkx in tests/test_oracle_tasks.py); er_graph is engine-defined - the oracle is the specification.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .sent_oracle import philox4x32_10

_MASK = 0xFFFFFFFF
ER_STREAM = 1          # Philox counter word 3: keeps the generator's stream apart from the SENT walk's (stream 0)
ER_PAIR_WORD0 = 8      # pair k of a graph reads word ER_PAIR_WORD0 + k of the graph's stream


def has_cycle_dfs(num_nodes: int, edge_list: Sequence[Tuple[int, int]]) -> bool:
    """test/train_performance_test.py:27-67, recursion unrolled into an explicit stack (same visiting order)."""
    if len(edge_list) == 0:
        return False
    adj: List[List[int]] = [[] for _ in range(num_nodes)]
    for s, t in edge_list:
        adj[s].append(t)
        adj[t].append(s)
    visited = [False] * num_nodes
    for root in range(num_nodes):
        if visited[root]:
            continue
        visited[root] = True
        stack = [(root, -1, 0)]
        while stack:
            node, parent, i = stack.pop()
            if i < len(adj[node]):
                stack.append((node, parent, i + 1))
                nb = adj[node][i]
                if not visited[nb]:
                    visited[nb] = True
                    stack.append((nb, node, 0))
                elif nb != parent:
                    return True
    return False


def shortest_distance(num_nodes: int, edge_list: Sequence[Tuple[int, int]], u: int, v: int) -> int:
    """Hops from u to v over the undirected graph; -1 when unreachable or when u / v is not a node."""
    if not (0 <= u < num_nodes and 0 <= v < num_nodes):
        return -1
    adj: List[List[int]] = [[] for _ in range(num_nodes)]
    for s, t in edge_list:
        adj[s].append(t)
        adj[t].append(s)
    dist = {u: 0}
    frontier = [u]
    while frontier:
        nxt = []
        for a in frontier:
            for b in adj[a]:
                if b not in dist:
                    dist[b] = dist[a] + 1
                    nxt.append(b)
        frontier = nxt
    return dist.get(v, -1)


def distance_label_token(d: int) -> str:
    return "INF" if d < 0 else f"len{d}"


def first_token_pos(row: Sequence[int], token: int) -> int:
    for i, t in enumerate(row):
        if int(t) == token:
            return i
    return -1


# --------------------------------------------------------------------------- ER generator (engine-defined)
class _Words:
    def __init__(self, seed: int, gid: int):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.gid = gid
        self._cache = {}

    def __call__(self, d: int) -> int:
        b = d >> 2
        if b not in self._cache:
            self._cache[b] = philox4x32_10((self.gid & _MASK, (self.gid >> 32) & _MASK, b, ER_STREAM), self.key)
        return self._cache[b][d & 3]


def prob_threshold(p: float) -> int:
    """P(edge) = thr / 2^32, thr = floor(p * 2^32) clipped to u32 (the host computes this once for p_min / p_max)."""
    return max(0, min(_MASK, int(np.floor(float(p) * 4294967296.0))))


def _sfn_edges(w, n: int, m: int, tri_q16: int):
    """Holme-Kim scale-free graphs with tunable clustering (docs/synthetic_data.md:15 "sfn"; networkx.powerlaw_cluster_graph):
    the Barabasi-Albert process of _ba_edges where, after a new node's first preferential-attachment edge, every further
    edge is with probability tri_q16 / 65536 a TRIAD step: a uniformly chosen neighbour of the last preferential target that
    the new node is not adjacent to yet (ascending node id; none -> an ordinary preferential step).  Preferential targets are
    the m distinct draws of _ba_edges, used in ascending node id, skipping any the new node became adjacent to through a
    triad.  Words: the m-subset draws as in _ba_edges, then per further edge one decision word (triad iff its top 16 bits
    < tri_q16) and, for a triad with candidates, one pick word (mulhi(word, #candidates)).  Always m * (n - m) edges."""
    adj = [set() for _ in range(n)]
    deg = [0] * n

    def add(a, b):
        adj[a].add(b)
        adj[b].add(a)
        deg[a] += 1
        deg[b] += 1
    for i in range(1, m + 1):
        add(0, i)
    e2, d = 2 * m, 0
    for source in range(m + 1, n):
        targets = []
        while len(targets) < m:
            r = (w(ER_PAIR_WORD0 + d) * e2) >> 32
            d += 1
            x, acc = 0, deg[0]
            while acc <= r:
                x += 1
                acc += deg[x]
            if x not in targets:
                targets.append(x)
        rest = sorted(targets)
        target = rest.pop(0)
        add(target, source)
        count = 1
        while count < m:
            a = w(ER_PAIR_WORD0 + d)
            d += 1
            done = False
            if (a >> 16) < tri_q16:
                cand = sorted(v for v in adj[target] if v != source and v not in adj[source])
                if cand:
                    b = w(ER_PAIR_WORD0 + d)
                    d += 1
                    add(cand[(b * len(cand)) >> 32], source)
                    done = True
            if not done:
                rest = [v for v in rest if v not in adj[source]]
                target = rest.pop(0)
                add(target, source)
            count += 1
        e2 += 2 * m
    return sorted((u, v) for u in range(n) for v in adj[u] if u < v)


def _ba_edges(w, n: int, m: int):
    """Barabasi-Albert after networkx.barabasi_albert_graph: a star on nodes 0..m, then every new node attaches to m
    DISTINCT existing nodes drawn with probability proportional to their degree (degrees as they were before this
    node's own edges).  Draw d reads word 8 + d: r = mulhi(word, sum of degrees), target = first node whose
    cumulative degree exceeds r; repeats are redrawn.  Edges returned sorted (u < v ascending)."""
    deg = [0] * n
    es = set()

    def add(a, b):
        es.add((min(a, b), max(a, b)))
        deg[a] += 1
        deg[b] += 1
    for i in range(1, m + 1):
        add(0, i)
    e2, d = 2 * m, 0
    for source in range(m + 1, n):
        targets = []
        while len(targets) < m:
            r = (w(ER_PAIR_WORD0 + d) * e2) >> 32
            d += 1
            x, acc = 0, deg[0]
            while acc <= r:
                x += 1
                acc += deg[x]
            if x not in targets:
                targets.append(x)
        for x in targets:
            add(x, source)
        e2 += 2 * m
    return sorted(es)


def er_graph(seed: int, gid: int, buckets: Sequence[Tuple[int, int]] = ((10, 20), (20, 40), (40, 50)),
             p_min: float = 0.1, p_max: float = 0.2, kind: str = "er", num_blocks: int = 2, out_ratio: float = 0.25):
    """-> (n, edges [(u, v)] with u < v ascending, (query_u, query_v)).
    word 0: bucket = mulhi(w, len(buckets));  word 1: n = lo + mulhi(w, hi - lo);
    word 2: thr = t_lo + mulhi(w, t_hi - t_lo);  word 3: query_u = mulhi(w, n);
    word 4: query_v = (query_u + 1 + mulhi(w, n - 1)) % n  (n == 1: query_v = query_u);
    pair k (row-major over u < v) is an edge iff word(8 + k) < thr (kind "er"); "sbm": < thr inside a block,
    < thr * out_ratio across blocks; "complete" / "star" (centre 0) / "path" (i - i+1): the deterministic shapes."""
    w = _Words(seed, gid)
    lo, hi = buckets[(w(0) * len(buckets)) >> 32]
    n = lo + ((w(1) * (hi - lo)) >> 32)
    t_lo, t_hi = prob_threshold(p_min), prob_threshold(p_max)
    thr = t_lo + ((w(2) * (t_hi - t_lo)) >> 32)
    qu = (w(3) * n) >> 32
    qv = (qu + 1 + ((w(4) * (n - 1)) >> 32)) % n if n > 1 else qu
    if kind == "ba":
        return n, _ba_edges(w, n, min(1 + ((w(2) * num_blocks) >> 32), n - 1)), (qu, qv)
    if kind == "sfn":      # out_ratio = probability of a triad step
        return n, _sfn_edges(w, n, min(1 + ((w(2) * num_blocks) >> 32), n - 1), max(0, min(65536, int(round(out_ratio * 65536))))), (qu, qv)
    thr_out = (thr * max(0, min(65536, int(round(out_ratio * 65536))))) >> 16
    edges = []
    k = 0
    for u in range(n):
        for v in range(u + 1, n):
            word = w(ER_PAIR_WORD0 + k)
            if kind == "er":
                keep = word < thr
            elif kind == "sbm":        # contiguous blocks: block(i) = i * num_blocks // n
                keep = word < (thr if (u * num_blocks) // n == (v * num_blocks) // n else thr_out)
            elif kind == "complete":
                keep = True
            elif kind == "star":
                keep = u == 0
            elif kind == "path":
                keep = v == u + 1
            else:
                raise ValueError(kind)
            if keep:
                edges.append((u, v))
            k += 1
    return n, edges, (qu, qv)
