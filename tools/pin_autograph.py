#!/usr/bin/env python
"""Pin the SENT specification to the real AutoGraph tokenizer - one command on a box that has it:

    python tools/pin_autograph.py [--autograph /path/to/AutoGraph] [--out tests/golden/sent_autograph.json]

The reference calls `autograph.datamodules.data.tokenizer.Graph2TrailTokenizer` (trainer/train_agtt.py:21, :250,
:514-542) from an un-vendored checkout; this container has neither the package nor a network, so SENT parity is
UNPINNED (oracle/sent_oracle.py header).  This tool closes the gap without touching AutoGraph's random number
generator:

 1. it builds the fixture graphs below (unlabeled graph-token shapes and labeled ZINC shapes: paths, cycles, stars,
    cliques, trees with ring closures, disconnected graphs, isolated nodes) as PyG-style Data objects and calls the
    REAL tokenizer on each, several times, with exactly the constructor / setter calls of train_agtt.py:514-542;
 2. for every candidate setting of the open decisions (oracle.sent_oracle.SentSpec: restart policy, bracket order,
    predecessor in the bracket, label placement, the doc's revisit reading, edge_label_shift 0 / -1 / +1) it searches
    for an explicit choice stream under which the oracle emits the same tokens (fit_choice_stream: a depth-first
    search over the walk's decisions - it needs the tokenizer's OUTPUT only);
 3. it writes the samples, the settings that reproduce ALL of them and the recovered streams (GtokChoice format) to
    tests/golden/sent_autograph.json.  tests/test_sent_pin.py then replays every stream through the oracle and, on a
    GPU, through the CUDA kernels, and compares token for token; a chi-square test over repeated samples of small
    graphs checks that AutoGraph's choice distribution is the uniform one the engine draws from.

Without AutoGraph the tool can be exercised end to end with `--self-test` (the oracle under a non-default SentSpec
plays AutoGraph; used by tests/test_sent_pin.py on the CPU).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "glearning-benchmark_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from oracle import sent_oracle as S  # noqa: E402


def fixture_graphs(seed: int = 0):
    """[(name, n, [(u, v)], x | None, attr | None)] - small enough for the search, varied enough to separate the specs."""
    rng = np.random.default_rng(seed)
    gs = [("single", 1, []), ("pair", 2, [(0, 1)]), ("path5", 5, [(i, i + 1) for i in range(4)]),
          ("cycle4", 4, [(0, 1), (1, 2), (2, 3), (0, 3)]), ("cycle6", 6, [(i, (i + 1) % 6) for i in range(6)]),
          ("star6", 6, [(0, i) for i in range(1, 6)]), ("k4", 4, [(a, b) for a in range(4) for b in range(a + 1, 4)]),
          ("k5", 5, [(a, b) for a in range(5) for b in range(a + 1, 5)]),
          ("two_triangles", 6, [(0, 1), (1, 2), (0, 2), (3, 4), (4, 5), (3, 5)]),
          ("isolated", 5, [(0, 1), (1, 2)]), ("bowtie", 5, [(0, 1), (1, 2), (0, 2), (2, 3), (3, 4), (2, 4)]),
          ("ladder", 8, [(i, i + 1) for i in range(3)] + [(i + 4, i + 5) for i in range(3)] + [(i, i + 4) for i in range(4)])]
    for i in range(12):                                          # ER graphs of the graph-token size class "small"
        n = int(rng.integers(6, 13))
        e = [(a, b) for a in range(n) for b in range(a + 1, n) if rng.random() < 0.3]
        gs.append((f"er{i}", n, e))
    out = []
    for name, n, e in gs:
        out.append((name, n, e, None, None))
    for i in range(12):                                          # ZINC-shaped: random tree + ring closures, labels
        n = int(rng.integers(5, 12))
        e = [(int(rng.integers(0, k)), k) for k in range(1, n)]
        for _ in range(int(rng.integers(0, 3))):
            a, b = sorted(rng.choice(n, 2, replace=False).tolist())
            if (a, b) not in e:
                e.append((a, b))
        x = rng.integers(0, 9, n).tolist()
        attr = rng.integers(1, 4, len(e)).tolist()
        out.append((f"mol{i}", n, e, x, attr))
    return out


class _Data:
    """duck-typed torch_geometric.data.Data (used when PyG is absent, e.g. --self-test)"""


def to_data(n, edges, x, attr):
    import torch
    try:
        from torch_geometric.data import Data
    except Exception:                                            # noqa: BLE001
        Data = None
    ei = torch.tensor([[u for u, v in edges] + [v for u, v in edges], [v for u, v in edges] + [u for u, v in edges]],
                      dtype=torch.long).reshape(2, -1)
    kw = {"edge_index": ei, "num_nodes": n}
    if x is not None:
        kw["x"] = torch.tensor(x, dtype=torch.long)
        kw["edge_attr"] = torch.tensor(list(attr) + list(attr), dtype=torch.long)
    if Data is not None:
        return Data(**kw)
    d = _Data()
    d.__dict__.update(kw)
    return d


def real_tokenizer(labeled: bool, max_nodes: int, max_len: int, autograph_root=None):
    """Exactly the calls of trainer/train_agtt.py:514-542."""
    if autograph_root:
        sys.path.insert(0, autograph_root)
    from autograph.datamodules.data.tokenizer import Graph2TrailTokenizer      # the REAL one, or ImportError
    if "glearning-benchmark_b200" in (getattr(sys.modules[Graph2TrailTokenizer.__module__], "__file__", "") or ""):
        raise ImportError("autograph resolves to this repo's stand-in, not to BorgwardtLab/AutoGraph")
    tok = Graph2TrailTokenizer(dataset_names=[], max_length=max_len, truncation_length=max_len, labeled_graph=labeled,
                               undirected=True)
    tok.set_num_nodes(max_nodes)
    if labeled:
        tok.set_num_node_and_edge_types(9, 4)
    return tok


def collect(tokenize_of, repeats: int, max_nodes: int, max_len: int):
    """-> samples: one dict per (graph, repeat) with the producer's tokens."""
    samples = []
    for name, n, edges, x, attr in fixture_graphs():
        labeled = x is not None
        tok, cfg = tokenize_of(labeled)
        rp, col, ea = S.canonical_csr(n, [u for u, v in edges], [v for u, v in edges], attr)
        for r in range(repeats):
            t = [int(v) for v in tok(n, edges, x, attr, r)]
            samples.append({"name": name, "repeat": r, "n": n, "edges": [list(e) for e in edges], "x": x, "attr": attr,
                            "labeled": labeled, "tokens": t, "rowptr": rp, "col": col, "eattr": ea,
                            "cfg": {"idx_offset": cfg.idx_offset, "node_idx_offset": cfg.node_idx_offset,
                                    "edge_idx_offset": cfg.edge_idx_offset, "max_length": cfg.max_length,
                                    "labeled": cfg.labeled}})
    return samples


def fit(samples):
    """-> [{spec, edge_label_shift, streams}] for every setting that reproduces every sample."""
    fits = []
    for shift in (0, -1, 1):
        def cfg_of(s_):
            c = s_["cfg"]
            return S.SentConfig(c["idx_offset"], c["node_idx_offset"], c["edge_idx_offset"], c["max_length"], c["labeled"], shift)
        for spec, streams in S.fit_spec(samples, cfg_of):
            fits.append({"spec": spec.__dict__ if hasattr(spec, "__dict__") else
                         {f: getattr(spec, f) for f in S.SentSpec.__dataclass_fields__},
                         "edge_label_shift": shift, "streams": streams})
    return fits


def distribution_samples(tokenize_of, draws: int):
    """Repeated samples of two tiny graphs (star6: 6 x 5! orderings are too many - K4 and path5 are used): the
    frequencies of whole sequences, for the chi-square test against the uniform-choice oracle."""
    out = {}
    for name, n, edges, x, attr in fixture_graphs():
        if name not in ("k4", "path5", "bowtie"):
            continue
        tok, _ = tokenize_of(False)
        freq = {}
        for r in range(draws):
            key = " ".join(str(int(v)) for v in tok(n, edges, None, None, 1000 + r))
            freq[key] = freq.get(key, 0) + 1
        out[name] = freq
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--autograph", default=os.environ.get("AUTOGRAPH_ROOT"), help="path of the BorgwardtLab/AutoGraph checkout")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "sent_autograph.json"))
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--draws", type=int, default=3000)
    ap.add_argument("--self-test", action="store_true", help="the oracle under a non-default SentSpec plays AutoGraph")
    args = ap.parse_args()
    max_nodes, max_len = 64, 512

    if args.self_test:
        truth = S.SentSpec(restart="frontier_first", bracket_order="visit_desc")
        source = {"producer": "self-test (oracle under " + repr(truth) + ")", "pinned": False}

        def tokenize_of(labeled):
            cfg = S.SentConfig(6, 6 + max_nodes, 6 + max_nodes + 9, max_len, labeled)

            def tok(n, edges, x, attr, r):
                rp, col, ea = S.canonical_csr(n, [u for u, v in edges], [v for u, v in edges], attr)
                return S.sent_tokenize(n, rp, col, cfg, S.PhiloxWords(77, 1000 * r + n), x, ea, truth)
            return tok, cfg
    else:
        try:
            real_tokenizer(False, max_nodes, max_len, args.autograph)
        except ImportError as ex:
            print(f"pin_autograph: the real tokenizer is not importable here ({ex}).\n"
                  "Clone BorgwardtLab/AutoGraph, install its environment (reference docs/setup.md:24-32) and re-run with "
                  "--autograph <checkout>.  SENT parity stays UNPINNED until then.", file=sys.stderr)
            return 2
        import autograph
        source = {"producer": "BorgwardtLab/AutoGraph Graph2TrailTokenizer", "autograph_file": getattr(autograph, "__file__", None),
                  "pinned": True}

        def tokenize_of(labeled):
            t = real_tokenizer(labeled, max_nodes, max_len, args.autograph)
            cfg = S.SentConfig(int(t.idx_offset), int(getattr(t, "node_idx_offset", 0) or 0),
                               int(getattr(t, "edge_idx_offset", 0) or 0), max_len, labeled)

            def tok(n, edges, x, attr, r):
                return t(to_data(n, edges, x, attr)).tolist()
            return tok, cfg

    samples = collect(tokenize_of, args.repeats, max_nodes, max_len)
    fits = fit(samples)
    dist = distribution_samples(tokenize_of, args.draws)
    for s_ in samples:
        for k in ("rowptr", "col", "eattr"):
            s_.pop(k)
    json.dump({"source": source, "samples": samples, "fits": fits, "distribution": dist}, open(args.out, "w"))
    print(f"pin_autograph: {len(samples)} samples, {len(fits)} specification setting(s) reproduce all of them -> {args.out}")
    for f in fits:
        print("   ", f["spec"], "edge_label_shift", f["edge_label_shift"])
    if not fits:
        print("    NO candidate reproduces every sample: extend oracle.sent_oracle.SentSpec (the samples are in the fixture)")
    return 0 if fits else 1


if __name__ == "__main__":
    sys.exit(main())
