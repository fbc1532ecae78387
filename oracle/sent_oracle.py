"""CPU oracle for the SENT trail tokenizer (AGTT half of the path).

TEST INFRASTRUCTURE ONLY (see oracle/ibtt_oracle.py header for the rule).

PARITY UNPINNED.  The reference calls ``Graph2TrailTokenizer.__call__`` from the
third-party repo BorgwardtLab/AutoGraph (module
``autograph/datamodules/data/tokenizer.py``; call sites trainer/train_agtt.py:21,
:250, :286, :514-542, :586).  That repo is not vendored under /root/reference, has
no pinned version (no submodule, no requirements entry; docs/setup.md:24-32 says
"pip install -e ."), is not installed here, and no reference test holds a golden
vector for it.  This file therefore *restates the published algorithm* (Chen et
al. 2025, "Flatten Graphs as Sequences: Transformers are Scalable Graph
Generators", SENT = Segmented Eulerian Neighborhood Trail) and is the NORMATIVE
SPEC for this repo.  What the reference itself fixes, and what this file obeys:

* special ids SOS=0 RESET=1 LADJ=2 RADJ=3 EOS=4 PAD=5 (train_agtt.py:176,195-206;
  zinc_vocab.py:197,215), ``idx_offset = 6 + len(dataset_names)`` (=6 at the call
  site, train_agtt.py:514-530 with ``dataset_names=[]``; vocab_size formula :586),
* ``node_idx_offset = idx_offset + max_num_nodes``, ``edge_idx_offset =
  node_idx_offset + num_node_types`` (train_agtt.py:177-179, 538-542),
* output is a 1-D int64 sequence starting with SOS and ending with EOS, at most
  ``max_length`` long (train_agtt.py:250, 516-517),
* the doc illustration docs/models/embed_gtt.md:17-19 (4-cycle -> ``<sos> 0 1 2 3
  ...``: node tokens are first-visit indices).

SPEC (decisions documented in DESIGN.md "SENT specification"):

 graph      canonical undirected simple graph: edge_index symmetrised, duplicate
            pairs merged (label of {u,v} = edge_attr of its first directed entry
            in caller order), self loops dropped, neighbours sorted by node id.
 state      vis[v] = first-visit index or -1; prefix property: when a node is
            first visited, ALL its edges to earlier-visited nodes are emitted at
            once (trail edge from the predecessor + bracketed neighbourhood set),
            so every edge between two visited nodes is already covered and a
            trail step always goes to an unvisited neighbour.
 emit       SOS
            start = pick(N) over nodes in id order
            visit(cur): first visit  -> tok(idx_offset + vis[cur])
                                       [labeled: tok(node_idx_offset + x[cur])]
                                       B = visited neighbours of cur except the
                                           trail predecessor, ascending vis[]
                                       if B: LADJ ([labeled: edge label] tok(idx_offset+vis[u]))* RADJ
                        revisit (only as the head of a new segment) -> tok(idx_offset + vis[cur])
            step:  C = unvisited neighbours of cur, ascending node id
                   if C: nxt = C[pick(|C|)]; [labeled: edge label of (cur,nxt)]; cur = nxt
                   else: R = visited nodes that still have an unvisited neighbour
                             (ascending node id); if none, R = unvisited nodes
                             (ascending node id); if none -> stop
                         RESET; cur = R[pick(|R|)]
            EOS;  sequence = sequence[:max_length]
 labels     node label token = node_idx_offset + x[v]; edge label token =
            edge_idx_offset + edge_attr + edge_label_shift (default 0).
 pick(c)    consumes ONE 32-bit word r per call (also when c == 1) and returns
            (r * c) >> 32.  Word d of graph g is either ``stream[g][d]`` (explicit
            choice stream; words past the end read as 0) or
            ``philox4x32_10(counter=(g_lo, g_hi, d >> 2, 0), key=(seed_lo, seed_hi))[d & 3]``.

Every edge is covered exactly once and ``decode`` recovers the input graph up to
the visit-order relabelling, which is what tests/test_oracle_sent.py checks.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

SOS, RESET, LADJ, RADJ, EOS, PAD = 0, 1, 2, 3, 4, 5
NUM_SPECIAL = 6

_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(counter: Sequence[int], key: Sequence[int]) -> Tuple[int, int, int, int]:
    """Philox-4x32-10 (Salmon et al., SC'11); same function cuRAND/Random123 ship."""
    c0, c1, c2, c3 = (int(c) & _MASK for c in counter)
    k0, k1 = (int(k) & _MASK for k in key)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


@dataclass
class SentConfig:
    idx_offset: int = 6
    node_idx_offset: int = 0       # only used when labeled
    edge_idx_offset: int = 0
    max_length: int = 600
    labeled: bool = False
    edge_label_shift: int = 0


class PhiloxWords:
    """word d of graph ``gid`` under ``seed``."""

    def __init__(self, seed: int, gid: int):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.gid = gid
        self.d = 0
        self._blk = None

    def __call__(self) -> int:
        d = self.d
        if d & 3 == 0 or self._blk is None:
            self._blk = philox4x32_10((self.gid & _MASK, (self.gid >> 32) & _MASK, d >> 2, 0), self.key)
        self.d += 1
        return self._blk[d & 3]


class StreamWords:
    def __init__(self, words: Sequence[int]):
        self.words = words
        self.d = 0

    def __call__(self) -> int:
        r = int(self.words[self.d]) if self.d < len(self.words) else 0
        self.d += 1
        return r


def canonical_csr(num_nodes: int, src: Sequence[int], dst: Sequence[int],
                  edge_attr: Optional[Sequence[int]] = None):
    """-> (rowptr int32[N+1], col int32[nnz], eattr int32[nnz] | None)."""
    first = {}
    for i, (u, v) in enumerate(zip(src, dst)):
        u, v = int(u), int(v)
        if u == v:
            continue
        key = (u, v) if u < v else (v, u)
        if key not in first:
            first[key] = int(edge_attr[i]) if edge_attr is not None else 0
    nbr: List[List[Tuple[int, int]]] = [[] for _ in range(num_nodes)]
    for (u, v), a in first.items():
        nbr[u].append((v, a))
        nbr[v].append((u, a))
    rowptr = np.zeros(num_nodes + 1, dtype=np.int32)
    col, ea = [], []
    for r in range(num_nodes):
        nbr[r].sort()
        rowptr[r + 1] = rowptr[r] + len(nbr[r])
        col += [c for c, _ in nbr[r]]
        ea += [a for _, a in nbr[r]]
    return rowptr, np.asarray(col, dtype=np.int32), (np.asarray(ea, dtype=np.int32) if edge_attr is not None else None)


def sent_tokenize(num_nodes: int, rowptr: np.ndarray, col: np.ndarray, cfg: SentConfig,
                  word: Callable[[], int], x: Optional[Sequence[int]] = None,
                  eattr: Optional[Sequence[int]] = None) -> np.ndarray:
    """One graph -> int64 token sequence, per the SPEC in the module docstring."""
    def pick(c: int) -> int:
        return (word() * c) >> 32

    out: List[int] = [SOS]
    n = int(num_nodes)
    if n > 0:
        vis = [-1] * n
        nvis = 0
        cur = pick(n)
        prev = -1
        while True:
            lo, hi = int(rowptr[cur]), int(rowptr[cur + 1])
            if vis[cur] < 0:
                vis[cur] = nvis
                nvis += 1
                out.append(cfg.idx_offset + vis[cur])
                if cfg.labeled:
                    out.append(cfg.node_idx_offset + int(x[cur]))
                B = sorted((vis[int(col[e])], e) for e in range(lo, hi)
                           if vis[int(col[e])] >= 0 and int(col[e]) != prev and int(col[e]) != cur)
                if B:
                    out.append(LADJ)
                    for k, e in B:
                        if cfg.labeled:
                            out.append(cfg.edge_idx_offset + int(eattr[e]) + cfg.edge_label_shift)
                        out.append(cfg.idx_offset + k)
                    out.append(RADJ)
            else:
                out.append(cfg.idx_offset + vis[cur])
            C = [e for e in range(lo, hi) if vis[int(col[e])] < 0]
            if C:
                e = C[pick(len(C))]
                if cfg.labeled:
                    out.append(cfg.edge_idx_offset + int(eattr[e]) + cfg.edge_label_shift)
                prev, cur = cur, int(col[e])
                continue
            R = [v for v in range(n) if vis[v] >= 0 and
                 any(vis[int(col[e])] < 0 for e in range(int(rowptr[v]), int(rowptr[v + 1])))]
            if not R:
                R = [v for v in range(n) if vis[v] < 0]
            if not R:
                break
            out.append(RESET)
            cur = R[pick(len(R))]
            prev = -1
    out.append(EOS)
    return np.asarray(out[: cfg.max_length], dtype=np.int64)


def sent_decode(tokens: Sequence[int], cfg: SentConfig):
    """Inverse of sent_tokenize on an untruncated sequence.
    -> (num_nodes, {(a,b): label} with a<b in visit-index space, [node labels])."""
    toks = [int(t) for t in tokens]
    assert toks[0] == SOS and toks[-1] == EOS, "not a complete SENT"
    i = 1
    edges = {}
    nlab: List[int] = []
    nseen = 0
    prev = -1
    pend_label = 0

    def is_node(t):
        hi = cfg.node_idx_offset if cfg.labeled else 1 << 62
        return cfg.idx_offset <= t < hi

    def is_elab(t):
        return cfg.labeled and t >= cfg.edge_idx_offset

    while toks[i] != EOS:
        t = toks[i]
        if t == RESET:
            prev = -1
            i += 1
            continue
        if is_elab(t):
            pend_label = t - cfg.edge_idx_offset - cfg.edge_label_shift
            i += 1
            continue
        assert is_node(t), f"unexpected token {t} at {i}"
        k = t - cfg.idx_offset
        i += 1
        if k == nseen:                               # first visit
            nseen += 1
            if cfg.labeled:
                nlab.append(toks[i] - cfg.node_idx_offset)
                i += 1
            if prev >= 0:
                key = (min(prev, k), max(prev, k))
                assert key not in edges, "edge covered twice"
                edges[key] = pend_label
            if toks[i] == LADJ:
                i += 1
                while toks[i] != RADJ:
                    lab = 0
                    if cfg.labeled:
                        lab = toks[i] - cfg.edge_idx_offset - cfg.edge_label_shift
                        i += 1
                    j = toks[i] - cfg.idx_offset
                    i += 1
                    assert 0 <= j < k
                    key = (j, k)
                    assert key not in edges, "edge covered twice"
                    edges[key] = lab
                i += 1
        else:
            assert k < nseen and prev < 0, "revisit only allowed at a segment head"
        prev = k
    return nseen, edges, nlab


def visit_order(num_nodes, rowptr, col, word) -> List[int]:
    """vis[] of the walk sent_tokenize would take with the same words (test helper
    to map decoded visit indices back to node ids)."""
    def pick(c):
        return (word() * c) >> 32
    n = int(num_nodes)
    vis = [-1] * n
    if n == 0:
        return vis
    nvis = 0
    cur = pick(n)
    while True:
        if vis[cur] < 0:
            vis[cur] = nvis
            nvis += 1
        C = [int(col[e]) for e in range(int(rowptr[cur]), int(rowptr[cur + 1])) if vis[int(col[e])] < 0]
        if C:
            cur = C[pick(len(C))]
            continue
        R = [v for v in range(n) if vis[v] >= 0 and
             any(vis[int(col[e])] < 0 for e in range(int(rowptr[v]), int(rowptr[v + 1])))]
        if not R:
            R = [v for v in range(n) if vis[v] < 0]
        if not R:
            return vis
        cur = R[pick(len(R))]
