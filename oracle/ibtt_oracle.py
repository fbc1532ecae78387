"""CPU oracle for the index-based (IBTT) half of the graph -> token path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and only as the checker /
timed CPU baseline.

Every function restates one function of the reference
(KevinBian107/GLearning-Benchmark, read-only at /root/reference) and cites the
file:line it follows.  Two families live here:

* ``*_text`` / ``encode_text`` / ``vocab_from_texts`` follow the reference's own
  *string* pipeline (serialize to text, split, dict lookup).  They are pinned
  against outputs of the reference itself (``tests/golden/*.json`` made by
  ``tests/golden/make_golden.py``, which imports /root/reference).
* ``*_ids`` are the closed-form graph -> id restatements the CUDA kernels mirror
  (SURVEY.md Appendix A).  ``tests/test_oracle_*.py`` prove them equal to the
  string pipeline on every fixture, so the kernels are checked against something
  that is itself pinned to the reference.

Pure Python / numpy on purpose: the reference is pure Python too.
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

# graph_data_loader/data_loader.py:10
SPECIAL = ["<pad>", "<bos>", "<e>", "<n>", "<q>", "<p>", "<eos>", "yes", "no"]

# graph_data_loader/zinc_vocab.py:20-28
ZINC_ATOMS = ["C", "N", "O", "F", "P", "S", "Cl", "Br", "I"]
ZINC_BONDS = ["single", "double", "triple", "aromatic"]
ZINC_SPECIALS = ["<bos>", "<eos>", "<pad>", "<unk>", "<q>", "<p>", "<atom>", "<bond>"]


# --------------------------------------------------------------------------
# a1: graph-token serializer (external repo alip67/graph-token, NOT vendored by
# the reference; grammar restated from docs/synthetic_data.md:46-68 and
# docs/models.md:22).  Parity to the real generator's text is unpinned; the
# in-repo inverse parsers (graph_token_dataset_autograph.py:14-54) pin the
# grammar and tests round-trip through a restatement of them.
# --------------------------------------------------------------------------
def graph_token_text(num_nodes: int, edges: Sequence[Tuple[int, int]], task: str,
                     label: Optional[str], query: Optional[Tuple[int, int]] = None) -> str:
    """``<bos> u v <e> ... <n> 0 1 ... <q> TASK <p> LABEL <eos>``.

    task 'cycle_check': TASK = ``has_cycle``, LABEL in {yes,no} (docs/models.md:22)
    task 'shortest_path': TASK = ``shortest_distance u v``, LABEL = ``lenK`` / ``INF``
    """
    out = ["<bos>"]
    for u, v in edges:
        out += [str(int(u)), str(int(v)), "<e>"]
    out.append("<n>")
    out += [str(k) for k in range(num_nodes)]
    out.append("<q>")
    if task == "cycle_check":
        out.append("has_cycle")
    elif task == "shortest_path":
        assert query is not None
        out += ["shortest_distance", str(int(query[0])), str(int(query[1]))]
    else:
        raise ValueError(f"unknown task {task!r}")
    out.append("<p>")
    if label is not None:
        out.append(label)
    out.append("<eos>")
    return " ".join(out)


def parse_graph_text(text: str) -> Tuple[List[int], List[Tuple[int, int]]]:
    """Inverse of a1; restates graph_token_dataset_autograph.py:14-54 (the same
    parser is copied at graph_token_dataset_nativegraph.py:54-96 and
    test/graph_seq_test.py:24-67)."""
    toks = text.split()
    edges: List[Tuple[int, int]] = []
    nodes: List[int] = []
    i, n = 0, len(toks)
    while i < n:
        if i + 2 < n and toks[i + 2] == "<e>":
            try:
                edges.append((int(toks[i]), int(toks[i + 1])))
                i += 3
            except ValueError:
                i += 1
        elif toks[i] == "<n>" and i + 1 < n:
            i += 1
            while i < n and toks[i] not in ("<q>", "<p>", "<eos>"):
                try:
                    nodes.append(int(toks[i]))
                except ValueError:
                    break
                i += 1
            break
        else:
            i += 1
    return nodes, edges


def yes_no_label(text: str) -> Optional[int]:
    """data_loader.py:12-17."""
    for t in reversed(text.split()):
        t = t.lower()
        if t == "yes":
            return 1
        if t == "no":
            return 0
    return None


def distance_label(text: str) -> Optional[int]:
    """data_loader.py:19-41: ``<p> lenK`` -> K-1, INF/INFINITY/<EOS> -> None."""
    toks = text.split()
    for i, tok in enumerate(toks):
        if tok == "<p>" and i + 1 < len(toks):
            lab = toks[i + 1].upper()
            if lab in ("INF", "INFINITY", "<EOS>"):
                return None
            if lab.startswith("LEN"):
                try:
                    return int(lab[3:]) - 1
                except ValueError:
                    pass
    return None


# --------------------------------------------------------------------------
# a2: build_vocab_from_texts (data_loader.py:451-463)
# --------------------------------------------------------------------------
def vocab_from_texts(texts: Iterable[str], min_freq: int = 1,
                     max_tokens: Optional[int] = None) -> Tuple[Dict[str, int], Dict[int, str]]:
    """SPECIAL ids 0..8 first, then tokens by descending count, ties by first
    occurrence (Counter.most_common is a stable sort over insertion order);
    stop as soon as the next id would reach ``max_tokens``."""
    counts: Counter = Counter()
    for t in texts:
        counts.update(t.split())
    vocab = {tok: i for i, tok in enumerate(SPECIAL)}
    nxt = len(vocab)
    order = sorted(counts.items(), key=lambda kv: -kv[1])  # stable => first-seen ties
    for tok, c in order:
        if tok in vocab:
            continue
        if c < min_freq:
            break
        vocab[tok] = nxt
        nxt += 1
        if max_tokens and nxt >= max_tokens:
            break
    return vocab, {i: t for t, i in vocab.items()}


# --------------------------------------------------------------------------
# a3: TokenDataset (data_loader.py:465-486)
# --------------------------------------------------------------------------
def encode_text(text: str, vocab: Dict[str, int], max_len: int, strip_label: bool = True) -> List[int]:
    toks = text.split()
    if strip_label and "<p>" in toks:
        toks = toks[: toks.index("<p>") + 1]          # :479-481
    pad = vocab["<pad>"]
    return [vocab.get(t, pad) for t in toks][:max_len]  # :482 (OOV -> <pad> id)


def token_dataset(examples: Sequence[dict], vocab: Dict[str, int], max_len: int = 512,
                  strip_label: bool = True, require_label: bool = True):
    """Returns (seqs, labels, kept) where ``kept`` are the indices of the
    examples that survive the label filter (data_loader.py:474-477)."""
    seqs, labels, kept = [], [], []
    for i, ex in enumerate(examples):
        text = ex["text"]
        label = ex.get("label")
        if label is None and isinstance(text, str):
            label = yes_no_label(text)
        if require_label and label is None:
            continue
        seqs.append(np.asarray(encode_text(text, vocab, max_len, strip_label), dtype=np.int64))
        labels.append(int(label) if label is not None else 0)   # :484 int() truncates floats
        kept.append(i)
    return seqs, np.asarray(labels, dtype=np.int64), kept


# --------------------------------------------------------------------------
# a4: collate (data_loader.py:488-497)
# --------------------------------------------------------------------------
def collate(seqs: Sequence[np.ndarray], labels: Sequence[int], pad_id: int):
    lmax = max(len(s) for s in seqs)
    X = np.full((len(seqs), lmax), pad_id, dtype=np.int64)
    attn = np.zeros((len(seqs), lmax), dtype=np.bool_)
    for i, s in enumerate(seqs):
        X[i, : len(s)] = s
        attn[i, : len(s)] = True
    return X, attn, np.asarray(labels, dtype=np.int64)


# --------------------------------------------------------------------------
# closed form of a1+a3 - what kernel K1 computes
# --------------------------------------------------------------------------
def ibtt_length(num_nodes: int, num_edges: int, task: str, max_len: int) -> int:
    q = 5 if task == "cycle_check" else 7
    return min(max_len, 3 * num_edges + num_nodes + q)


def ibtt_ids(num_nodes: int, src: np.ndarray, dst: np.ndarray, task: str,
             vocab: Dict[str, int], max_len: int,
             query: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """ids of ``encode_text(graph_token_text(...))`` without building the string.
    node ids go through the data-dependent ``vocab[str(k)]`` table, OOV -> <pad>."""
    pad = vocab["<pad>"]
    def node(k: int) -> int:
        return vocab.get(str(int(k)), pad)
    ids = [vocab.get("<bos>", pad)]
    for u, v in zip(src, dst):
        ids += [node(u), node(v), vocab.get("<e>", pad)]
    ids.append(vocab.get("<n>", pad))
    ids += [node(k) for k in range(num_nodes)]
    ids.append(vocab.get("<q>", pad))
    if task == "cycle_check":
        ids.append(vocab.get("has_cycle", pad))
    else:
        ids += [vocab.get("shortest_distance", pad), node(query[0]), node(query[1])]
    ids.append(vocab.get("<p>", pad))
    return np.asarray(ids[:max_len], dtype=np.int64)


# --------------------------------------------------------------------------
# a5/a6: ZINCTokenizationDataset (zinc_dataset_indexbase.py:86-227)
# --------------------------------------------------------------------------
def zinc_label_token(label: float) -> str:
    """zinc_dataset_indexbase.py:192."""
    return f"val_{label:.2f}".replace(".", "_").replace("-", "neg")


def zinc_text(x: Sequence[int], src: Sequence[int], dst: Sequence[int],
              edge_attr: Sequence[int], label: float) -> str:
    """tokenize_molecule (zinc_dataset_indexbase.py:143-195): all atoms, then one
    ``<bond> TYPE u v`` per first-seen undirected pair in edge_index column order
    (u, v in the orientation of that first column)."""
    out = ["<bos>"]
    for a in x:
        a = int(a)
        out += ["<atom>", ZINC_ATOMS[a] if 0 <= a < len(ZINC_ATOMS) else "X"]      # :104
    seen = set()
    for i, (u, v) in enumerate(zip(src, dst)):
        u, v = int(u), int(v)
        key = (u, v) if u <= v else (v, u)                                          # :178
        if key in seen:
            continue
        seen.add(key)
        b = int(edge_attr[i]) if i < len(edge_attr) else 0
        out += ["<bond>", ZINC_BONDS[b - 1] if 1 <= b <= 4 else "unknown", str(u), str(v)]  # :128-139,183
    out += ["<q>", "regression", "<p>", zinc_label_token(label), "<eos>"]
    return " ".join(out)


def zinc_item_text(x, src, dst, edge_attr, label: float, max_len: int) -> str:
    """__getitem__ truncation (zinc_dataset_indexbase.py:216-221)."""
    toks = zinc_text(x, src, dst, edge_attr, label).split()
    if len(toks) > max_len:
        toks = toks[: max_len - 1] + ["<eos>"]
    return " ".join(toks)


# a7: zinc_vocab.py:35-78, :154-179
def zinc_fixed_vocab() -> Dict[str, int]:
    v: Dict[str, int] = {}
    for t in ZINC_SPECIALS + ZINC_ATOMS + ZINC_BONDS + ["regression"]:
        v[t] = len(v)
    return v


def zinc_extend_vocab(base: Dict[str, int], dynamic: Sequence[str]) -> Dict[str, int]:
    v = dict(base)
    nxt = max(v.values()) + 1
    for t in dynamic:
        if t not in v:
            v[t] = nxt
            nxt += 1
    return v


def zinc_length(num_nodes: int, num_unique_edges: int, item_max_len: int, ds_max_len: int) -> int:
    """Length after a6 (text truncation at item_max_len) then a3 (cut after <p>,
    ``[:ds_max_len]``).  Full text has T = 2N + 4E + 6 tokens, <p> at T-3."""
    T = 2 * num_nodes + 4 * num_unique_edges + 6
    if T <= item_max_len:
        L = T - 2                       # ... <p>
    elif T == item_max_len + 1:
        L = item_max_len - 1            # <p> survives at index item_max_len-2, <eos> cut by strip
    else:
        L = item_max_len                # first item_max_len-1 tokens + <eos>, no <p> left
    return min(L, ds_max_len)


def zinc_ids(x, src, dst, edge_attr, vocab: Dict[str, int], item_max_len: int,
             ds_max_len: int) -> np.ndarray:
    """Closed form of encode_text(zinc_item_text(...)) - what kernel K2 computes."""
    pad = vocab["<pad>"]
    g = lambda t: vocab.get(t, pad)
    ids = [g("<bos>")]
    for a in x:
        a = int(a)
        ids += [g("<atom>"), g(ZINC_ATOMS[a]) if 0 <= a < 9 else g("X")]
    seen = set()
    for i, (u, v) in enumerate(zip(src, dst)):
        u, v = int(u), int(v)
        key = (min(u, v), max(u, v))
        if key in seen:
            continue
        seen.add(key)
        b = int(edge_attr[i])
        ids += [g("<bond>"), g(ZINC_BONDS[b - 1]) if 1 <= b <= 4 else g("unknown"), g(str(u)), g(str(v))]
    ids += [g("<q>"), g("regression"), g("<p>")]
    T = len(ids) + 2
    if T > item_max_len:
        if T == item_max_len + 1:
            ids = ids[: item_max_len - 1]
        else:
            ids = ids[: item_max_len - 1] + [g("<eos>")]
    return np.asarray(ids[:ds_max_len], dtype=np.int64)


# --------------------------------------------------------------------------
# a11-a13: AGTT post-processing (trainer/train_agtt.py:171-302)
# --------------------------------------------------------------------------
def remap_zinc_token(t: int, idx_offset: int, node_idx_offset: int, edge_idx_offset: int) -> int:
    """train_agtt.py:193-242, branch order preserved."""
    if t == 0:
        return 0                        # SOS -> <bos>
    if t in (1, 2, 3, 5):
        return 2                        # RESET/LADJ/RADJ/PAD -> <pad>
    if t == 4:
        return 1                        # EOS -> <eos>
    if node_idx_offset <= t < edge_idx_offset:
        a = t - node_idx_offset
        return 8 + a if 0 <= a < 9 else 22 + t
    if t >= edge_idx_offset:
        b = t - edge_idx_offset + 1
        return 17 + (b - 1) if 1 <= b <= 4 else 22 + t
    if idx_offset <= t < node_idx_offset:
        return 22 + (t - idx_offset)
    return 22 + t


def remap_zinc(tokens: Sequence[int], idx_offset: int, node_idx_offset: int,
               edge_idx_offset: int) -> np.ndarray:
    return np.asarray([remap_zinc_token(int(t), idx_offset, node_idx_offset, edge_idx_offset)
                       for t in tokens], dtype=np.int64)


def append_query(tokens: np.ndarray, idx_offset: int, num_nodes: int, u: int, v: int) -> np.ndarray:
    """train_agtt.py:257-267: ``[idx_offset+num_nodes, idx_offset+u, idx_offset+v]``."""
    return np.concatenate([np.asarray(tokens, dtype=np.int64),
                           np.asarray([idx_offset + num_nodes, idx_offset + u, idx_offset + v], dtype=np.int64)])


def collate_agtt(seqs: Sequence[np.ndarray], labels: Sequence, pad_id: int = 5):
    """train_agtt.py:276-302: pad with Graph2TrailTokenizer.pad (=5) always; labels
    float32 if python floats else int64."""
    X, attn, _ = collate(seqs, [0] * len(seqs), pad_id)
    if isinstance(labels[0], float):
        Y = np.asarray(labels, dtype=np.float32)
    else:
        Y = np.asarray(labels, dtype=np.int64)
    return X, attn, Y
