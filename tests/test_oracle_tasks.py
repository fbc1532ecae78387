"""tasks oracle self-checks: the restated DFS label oracle against brute force and against the
"listed edges > N - components" form the kernels use, BFS distances against networkx, the first-<q> scan against
the reference's own expression, and the engine-defined ER stream (shape, determinism, distribution)."""
import itertools

import numpy as np
import pytest
import torch

from oracle import ibtt_oracle as O
from oracle import tasks_oracle as T


def _components(n, edges):
    parent = list(range(n))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for u, v in edges:
        parent[find(u)] = find(v)
    return len({find(a) for a in range(n)})


def test_has_cycle_dfs_known_cases():
    assert T.has_cycle_dfs(3, []) is False
    assert T.has_cycle_dfs(3, [(0, 1), (1, 2)]) is False
    assert T.has_cycle_dfs(3, [(0, 1), (1, 2), (0, 2)]) is True
    assert T.has_cycle_dfs(4, [(0, 1), (2, 3)]) is False
    assert T.has_cycle_dfs(2, [(0, 1), (0, 1)]) is True          # a repeated edge closes a cycle for that DFS
    assert T.has_cycle_dfs(2, [(0, 1), (1, 0)]) is True
    assert T.has_cycle_dfs(3, [(1, 1)]) is True                  # and so does a self loop
    assert T.has_cycle_dfs(5, [(0, 1), (1, 2), (3, 4), (4, 3)]) is True


def test_has_cycle_equals_edge_count_form():
    """The kernels compute the label as M > N - C; exhaustive over multigraph edge lists on <= 4 nodes, random beyond."""
    for n in (1, 2, 3, 4):
        pairs = [(u, v) for u in range(n) for v in range(n)]
        for m in range(0, 4):
            for edges in itertools.product(pairs, repeat=m):
                assert T.has_cycle_dfs(n, list(edges)) == (len(edges) > n - _components(n, edges)), (n, edges)
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(1, 30))
        m = int(rng.integers(0, 2 * n))
        edges = [(int(a), int(b)) for a, b in zip(rng.integers(0, n, m), rng.integers(0, n, m))]
        assert T.has_cycle_dfs(n, edges) == (len(edges) > n - _components(n, edges))


def test_shortest_distance_vs_networkx():
    nx = pytest.importorskip("networkx")
    rng = np.random.default_rng(1)
    for _ in range(100):
        n = int(rng.integers(1, 40))
        g = nx.gnp_random_graph(n, float(rng.uniform(0.02, 0.3)), seed=int(rng.integers(0, 1 << 30)))
        u, v = int(rng.integers(0, n)), int(rng.integers(0, n))
        try:
            want = nx.shortest_path_length(g, u, v)
        except nx.NetworkXNoPath:
            want = -1
        assert T.shortest_distance(n, list(g.edges()), u, v) == want
    assert T.shortest_distance(3, [(0, 1)], 0, 5) == -1
    assert T.distance_label_token(-1) == "INF" and T.distance_label_token(3) == "len3"
    # the label the reference parses back: lenK -> class K - 1, INF -> dropped (data_loader.py:19-41)
    assert O.distance_label("<q> shortest_distance 0 1 <p> len3 <eos>") == 2
    assert O.distance_label("<q> shortest_distance 0 1 <p> INF <eos>") is None


def test_first_token_pos_is_the_reference_expression():
    """(x[b] == q).nonzero()[0] of trainer/train_ibtt.py:91-96."""
    rng = np.random.default_rng(2)
    for _ in range(50):
        row = rng.integers(0, 6, int(rng.integers(1, 40)))
        q = int(rng.integers(0, 7))
        nz = (torch.from_numpy(row) == q).nonzero(as_tuple=True)[0]
        assert T.first_token_pos(row, q) == (int(nz[0]) if len(nz) else -1)


def test_er_graph_stream():
    seen_b = set()
    dens = []
    for gid in range(400):
        n, edges, (qu, qv) = T.er_graph(7, gid)
        assert 10 <= n < 50 and 0 <= qu < n and 0 <= qv < n and qu != qv
        assert edges == sorted(set(edges)) and all(u < v < n for u, v in edges)
        seen_b.add(0 if n < 20 else (1 if n < 40 else 2))
        dens.append(len(edges) / (n * (n - 1) / 2))
        if gid < 5:
            assert T.er_graph(7, gid) == (n, edges, (qu, qv))
            assert T.er_graph(8, gid) != (n, edges, (qu, qv))
    assert seen_b == {0, 1, 2}
    assert 0.13 < float(np.mean(dens)) < 0.17            # p ~ U(0.1, 0.2)
    n, edges, q = T.er_graph(0, 0, buckets=((1, 2),))
    assert n == 1 and edges == [] and q == (0, 0)
    assert T.er_graph(0, 3, buckets=((5, 6),), p_min=1.0, p_max=1.0)[1] == [(u, v) for u in range(5) for v in range(u + 1, 5)] \
        or T.prob_threshold(1.0) == 0xFFFFFFFF           # p = 1 saturates at (2^32 - 1) / 2^32
    assert T.er_graph(0, 3, buckets=((5, 6),), p_min=0.0, p_max=0.0)[1] == []


def test_oracle_matches_reference_golden():
    """tests/golden/tasks_labels.json holds outputs of the reference's own has_cycle_dfs (lifted out of
    test/train_performance_test.py by tests/golden/make_tasks_golden.py) and of the trainer's first-<q> expression."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tasks_labels.json")))
    assert len(gold["cycle_cases"]) == 400
    for c in gold["cycle_cases"]:
        edges = [tuple(e) for e in c["edges"]]
        assert T.has_cycle_dfs(c["n"], edges) == c["has_cycle"]
        assert (len(edges) > c["n"] - _components(c["n"], edges)) == c["has_cycle"]
    for r in gold["query_rows"]:
        assert T.first_token_pos(r["row"], r["q"]) == r["pos"]


def test_generator_families():
    n, e, _ = T.er_graph(3, 1, kind="complete")
    assert e == [(u, v) for u in range(n) for v in range(u + 1, n)]
    n, e, _ = T.er_graph(3, 2, kind="star")
    assert e == [(0, v) for v in range(1, n)] and not T.has_cycle_dfs(n, e)
    n, e, _ = T.er_graph(3, 3, kind="path")
    assert e == [(i, i + 1) for i in range(n - 1)] and T.shortest_distance(n, e, 0, n - 1) == n - 1
    inside = across = pin = pac = 0
    for gid in range(300):
        n, e, _ = T.er_graph(5, gid, kind="sbm", num_blocks=3, out_ratio=0.1, p_min=0.4, p_max=0.4)
        blk = lambda i: (i * 3) // n
        es = set(e)
        for u in range(n):
            for v in range(u + 1, n):
                if blk(u) == blk(v):
                    pin += 1; inside += (u, v) in es
                else:
                    pac += 1; across += (u, v) in es
    assert 0.37 < inside / pin < 0.43 and 0.03 < across / pac < 0.05
    # same header stream for every family: sizes and queries do not depend on the kind
    assert T.er_graph(9, 4, kind="sbm")[0] == T.er_graph(9, 4)[0] and T.er_graph(9, 4, kind="path")[2] == T.er_graph(9, 4)[2]


def test_ba_generator_shape():
    """Barabasi-Albert family: m * (n - m) edges, connected, every node past the initial star has degree >= m, and the
    early nodes collect the high degrees (preferential attachment)."""
    hub = tail = 0
    for gid in range(300):
        n, e, _ = T.er_graph(2, gid, kind="ba", num_blocks=3)
        m = min(1 + ((T._Words(2, gid)(2) * 3) >> 32), n - 1)
        assert len(e) == m * (n - m) and e == sorted(set(e))
        deg = [0] * n
        for u, v in e:
            deg[u] += 1
            deg[v] += 1
        assert all(deg[i] >= m for i in range(m + 1, n)) and T.shortest_distance(n, e, 0, n - 1) > 0
        hub += deg[0]
        tail += deg[n - 1]
    assert hub > 3 * tail


def test_sfn_generator_shape():
    """Holme-Kim family (docs/synthetic_data.md:15 "sfn", the test algorithm of configs/ibtt_graph_token.yaml:5): m * (n - m)
    edges like ba, connected, triad steps close triangles (clustering rises with the triad probability), hubs stay hubs, and
    with triad probability 0 the graph is the ba graph of the same words as long as no decision word is consumed (m = 1)."""
    def triangles(n, e):
        es = set(e)
        adj = [set() for _ in range(n)]
        for u, v in e:
            adj[u].add(v)
            adj[v].add(u)
        return sum(1 for u, v in es for w in adj[u] & adj[v] if w > v)
    tri = {0.0: 0, 0.9: 0}
    hub = tail = 0
    for gid in range(200):
        for p in tri:
            n, e, _ = T.er_graph(2, gid, kind="sfn", num_blocks=4, out_ratio=p)
            m = min(1 + ((T._Words(2, gid)(2) * 4) >> 32), n - 1)
            assert len(e) == m * (n - m) and e == sorted(set(e)) and all(u < v for u, v in e)
            assert T.shortest_distance(n, e, 0, n - 1) > 0
            tri[p] += triangles(n, e)
            if m == 1:
                assert e == T.er_graph(2, gid, kind="ba", num_blocks=4)[1]
        deg = [0] * n
        for u, v in e:
            deg[u] += 1
            deg[v] += 1
        hub += deg[0]
        tail += deg[n - 1]
    assert tri[0.9] > 1.3 * tri[0.0] and hub > 3 * tail
