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
    append_eos: bool = True        # Graph2TrailTokenizer(append_eos=...) (train_agtt.py:514-530 leaves the default)


@dataclass(frozen=True)
class SentSpec:
    """The decisions the reference leaves open (AutoGraph's source is not in the checkout): every field is a switch,
    the defaults are this repo's normative choice.  tools/pin_autograph.py runs the real tokenizer when it can be
    imported, searches this space for the one setting that reproduces ALL its outputs (fit_spec) and writes
    tests/golden/sent_autograph.json; GtokSentCfg carries the same switches for the kernels (`restart`,
    `bracket_order`, `edge_label_shift`; the others are oracle-only until a pin says they are needed).

    restart          where a new segment starts when the walk is stuck
                       "open_first"      visited nodes that still have an unvisited neighbour, else unvisited nodes
                       "unvisited"       any unvisited node
                       "frontier_first"  unvisited nodes with a visited neighbour, else any unvisited node
    bracket_order    order of the neighbourhood set: "visit_asc" | "visit_desc" | "node_asc"
    bracket_prev     the trail predecessor is listed in the bracket as well
    node_label_first node-type token BEFORE the node index token (labeled mode)
    edge_label_after edge-type token AFTER the node index it leads to
    revisit_closes   docs/models/embed_gtt.md:17-19,64 reading ("<sos> 0 1 2 3 0"): no brackets at all - the trail
                     follows ANY uncovered edge and re-emits the index of a visited node it returns to
    """
    restart: str = "open_first"
    bracket_order: str = "visit_asc"
    bracket_prev: bool = False
    node_label_first: bool = False
    edge_label_after: bool = False
    revisit_closes: bool = False


DEFAULT_SPEC = SentSpec()


class NeedChoice(Exception):
    """Raised by ScriptedWords when the script is exhausted: `count` options at this decision."""

    def __init__(self, count: int):
        super().__init__(count)
        self.count = count


def word_for(k: int, c: int) -> int:
    """Smallest 32-bit word r with (r * c) >> 32 == k: the explicit-choice-stream encoding of 'option k of c'."""
    return ((k << 32) + c - 1) // c


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
                  eattr: Optional[Sequence[int]] = None, spec: SentSpec = DEFAULT_SPEC) -> np.ndarray:
    """One graph -> int64 token sequence, per the SPEC in the module docstring (spec = the open decisions)."""
    out: List[int] = []
    _sent_emit(out, num_nodes, rowptr, col, cfg, lambda c: (word() * c) >> 32, x, eattr, spec)
    return np.asarray(out[: cfg.max_length], dtype=np.int64)


def _sent_emit(out: List[int], num_nodes: int, rowptr, col, cfg: SentConfig, pick: Callable[[int], int],
               x, eattr, spec: SentSpec) -> None:
    """Appends the untruncated sequence to `out` (the caller can read what was emitted so far when `pick` raises)."""
    out.append(SOS)
    n = int(num_nodes)
    nbrs = lambda v: range(int(rowptr[v]), int(rowptr[v + 1]))
    elab = lambda e: cfg.edge_idx_offset + int(eattr[e]) + cfg.edge_label_shift
    if n > 0 and spec.revisit_closes:
        _sent_revisit(n, rowptr, col, cfg, pick, x, eattr, spec, out)
    elif n > 0:
        vis = [-1] * n
        nvis = 0
        cur = pick(n)
        prev = -1
        pend = None                                   # edge label owed after the next node token (edge_label_after)
        while True:
            if vis[cur] < 0:
                vis[cur] = nvis
                nvis += 1
                if cfg.labeled and spec.node_label_first:
                    out.append(cfg.node_idx_offset + int(x[cur]))
                out.append(cfg.idx_offset + vis[cur])
                if cfg.labeled and not spec.node_label_first:
                    out.append(cfg.node_idx_offset + int(x[cur]))
                if pend is not None:
                    out.append(pend)
                    pend = None
                B = [(vis[int(col[e])], int(col[e]), e) for e in nbrs(cur)
                     if vis[int(col[e])] >= 0 and int(col[e]) != cur and (spec.bracket_prev or int(col[e]) != prev)]
                if spec.bracket_order == "visit_asc":
                    B.sort()
                elif spec.bracket_order == "visit_desc":
                    B.sort(reverse=True)
                elif spec.bracket_order != "node_asc":
                    raise ValueError(spec.bracket_order)
                if B:
                    out.append(LADJ)
                    for k, _, e in B:
                        if cfg.labeled and not spec.edge_label_after:
                            out.append(elab(e))
                        out.append(cfg.idx_offset + k)
                        if cfg.labeled and spec.edge_label_after:
                            out.append(elab(e))
                    out.append(RADJ)
            else:
                out.append(cfg.idx_offset + vis[cur])
            C = [e for e in nbrs(cur) if vis[int(col[e])] < 0]
            if C:
                e = C[pick(len(C))]
                if cfg.labeled:
                    if spec.edge_label_after:
                        pend = elab(e)
                    else:
                        out.append(elab(e))
                prev, cur = cur, int(col[e])
                continue
            has_unvis = lambda v: any(vis[int(col[e])] < 0 for e in nbrs(v))
            has_vis = lambda v: any(vis[int(col[e])] >= 0 for e in nbrs(v))
            if spec.restart == "open_first":
                R = [v for v in range(n) if vis[v] >= 0 and has_unvis(v)]
            elif spec.restart == "frontier_first":
                R = [v for v in range(n) if vis[v] < 0 and has_vis(v)]
            elif spec.restart == "unvisited":
                R = []
            else:
                raise ValueError(spec.restart)
            if not R:
                R = [v for v in range(n) if vis[v] < 0]
            if not R:
                break
            out.append(RESET)
            cur = R[pick(len(R))]
            prev = -1
    if cfg.append_eos:
        out.append(EOS)


def _sent_revisit(n, rowptr, col, cfg, pick, x, eattr, spec, out) -> None:
    """SentSpec.revisit_closes: segments of a plain random trail decomposition - every step follows a uniformly chosen
    uncovered incident edge (to a visited node as well, whose index is re-emitted); stuck -> RESET at a visited node
    with an uncovered edge, else at an unvisited node."""
    vis = [-1] * n
    nvis = 0
    covered = set()
    key = lambda v, e: (min(v, int(col[e])), max(v, int(col[e])))
    unc = lambda v: [e for e in range(int(rowptr[v]), int(rowptr[v + 1])) if key(v, e) not in covered]
    cur = pick(n)
    while True:
        first = vis[cur] < 0
        if first:
            vis[cur] = nvis
            nvis += 1
        if cfg.labeled and first and spec.node_label_first:
            out.append(cfg.node_idx_offset + int(x[cur]))
        out.append(cfg.idx_offset + vis[cur])
        if cfg.labeled and first and not spec.node_label_first:
            out.append(cfg.node_idx_offset + int(x[cur]))
        C = unc(cur)
        if C:
            e = C[pick(len(C))]
            covered.add(key(cur, e))
            if cfg.labeled:
                out.append(cfg.edge_idx_offset + int(eattr[e]) + cfg.edge_label_shift)
            cur = int(col[e])
            continue
        R = [v for v in range(n) if vis[v] >= 0 and unc(v)]
        if not R:
            R = [v for v in range(n) if vis[v] < 0]
        if not R:
            return
        out.append(RESET)
        cur = R[pick(len(R))]


def fit_choice_stream(num_nodes: int, rowptr, col, cfg: SentConfig, target: Sequence[int], x=None, eattr=None,
                      spec: SentSpec = DEFAULT_SPEC, max_runs: int = 100_000) -> Optional[List[int]]:
    """Is `target` a sequence this specification can produce?  Depth-first search over the decisions (start node,
    k-th unvisited neighbour, restart node): a partial script of options is extended one decision at a time and
    abandoned as soon as the tokens it forces leave the target.  -> the explicit choice stream (32-bit words, GtokChoice
    format) that replays `target` through sent_tokenize / the CUDA kernels, or None if no stream exists.  Needs nothing
    from the implementation that produced `target` but its output - which is what makes an AutoGraph pin possible
    without touching its random number generator (tools/pin_autograph.py)."""
    target = [int(t) for t in target]
    truncated = len(target) >= cfg.max_length
    runs = 0

    def run(script: List[int]):
        out: List[int] = []
        counts: List[int] = []
        it = iter(script)

        def pick(c: int) -> int:
            k = next(it, None)
            if k is None:
                raise NeedChoice(c)
            counts.append(c)
            return k
        try:
            _sent_emit(out, num_nodes, rowptr, col, cfg, pick, x, eattr, spec)
            return out, None, counts
        except NeedChoice as nc:
            return out, nc.count, counts

    stack: List[List[int]] = [[]]
    while stack:
        script = stack.pop()
        runs += 1
        if runs > max_runs:
            raise RuntimeError("fit_choice_stream: search budget exhausted")
        out, need, counts = run(script)
        m = min(len(out), len(target))
        if out[:m] != target[:m]:
            continue
        if need is None:
            if out[: cfg.max_length] != target:
                continue
        elif not (truncated and len(out) >= len(target)):     # (a truncated target is settled once it is covered:
            if len(out) > len(target):                          #  words past the script read as 0, any option will do)
                continue
            stack.extend(script + [k] for k in range(need - 1, -1, -1))
            continue
        return [word_for(k, c) for k, c in zip(script, counts)]
    return None


def fit_spec(samples: Sequence[dict], cfg_of: Callable[[dict], SentConfig], candidates: Optional[Sequence[SentSpec]] = None):
    """samples: dicts with n, rowptr, col, tokens (+ x, eattr).  -> [(spec, streams)] for every candidate specification
    under which EVERY sample has a choice stream (streams[i] replays samples[i])."""
    if candidates is None:
        candidates = [SentSpec(r, b, p, nf, ea, rv)
                      for rv in (False, True) for r in ("open_first", "unvisited", "frontier_first")
                      for b in ("visit_asc", "visit_desc", "node_asc") for p in (False, True)
                      for nf in (False, True) for ea in (False, True)
                      if not (rv and (r != "open_first" or b != "visit_asc" or p or ea))]
    good = []
    for spec in candidates:
        streams = []
        for smp in samples:
            st = fit_choice_stream(smp["n"], smp["rowptr"], smp["col"], cfg_of(smp), smp["tokens"], smp.get("x"),
                                   smp.get("eattr"), spec)
            if st is None:
                break
            streams.append(st)
        else:
            good.append((spec, streams))
    return good


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
