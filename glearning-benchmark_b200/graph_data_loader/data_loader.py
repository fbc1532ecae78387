"""Drop-in for the reference's graph_data_loader/data_loader.py on the IBTT path.

Same public names, argument meaning and error behaviour; the per-graph Python hot
loops are replaced by the gtok CUDA kernels:

  TokenDataset.__init__   (reference data_loader.py:465-486)  -> K1 / K2 on the GPU
  collate                 (reference data_loader.py:488-497)  -> K4 on the GPU

Host-only helpers (label parsers, JSON loading, vocab building, class counting) keep
the reference's semantics, restated here; they are not on the hot path.

There is NO CPU fallback: examples whose text is not the graph-token grammar and that
carry no graph handle raise ValueError; calling the GPU collate from a forked
DataLoader worker raises RuntimeError (use num_workers=0 - collate now costs
microseconds on the device).
"""
from __future__ import annotations

import json
import os
from collections import Counter
from glob import glob
from typing import Any, Dict, List, Optional, Tuple

import weakref

import torch
from torch.utils.data import Dataset

SPECIAL = ["<pad>", "<bos>", "<e>", "<n>", "<q>", "<p>", "<eos>", "yes", "no"]    # reference :10


# --------------------------------------------------------------------------- label / query parsers
def parse_yes_no_from_text(text: str) -> Optional[int]:
    """Last yes/no token of the text (reference :12-17)."""
    for tok in reversed(text.split()):
        low = tok.lower()
        if low == "yes":
            return 1
        if low == "no":
            return 0
    return None


def parse_distance_label_from_text(text: str) -> Optional[int]:
    """'<p> lenK' -> K-1; INF / INFINITY / <EOS> -> None (reference :19-41)."""
    toks = text.split()
    for i, tok in enumerate(toks[:-1]):
        if tok != "<p>":
            continue
        lab = toks[i + 1].upper()
        if lab in ("INF", "INFINITY", "<EOS>"):
            return None
        if lab.startswith("LEN"):
            try:
                return int(lab[3:]) - 1
            except ValueError:
                pass
    return None


def parse_query_nodes_from_text(text: str) -> Optional[Tuple[int, int]]:
    """'<q> shortest_distance u v' -> (u, v) (reference :43-55)."""
    toks = text.split()
    for i, tok in enumerate(toks):
        if tok == "<q>" and i + 3 < len(toks) and toks[i + 1] == "shortest_distance":
            try:
                return int(toks[i + 2]), int(toks[i + 3])
            except ValueError:
                pass
    return None


_YES = ("yes", "true", "connected", "reachable")
_NO = ("no", "false", "disconnected", "unreachable")


def _extract_text_and_label(rec: Any, task: str = "cycle_check"):
    """(text, label, query_nodes) of one JSON record (reference :57-110)."""
    def from_text(t):
        if task == "shortest_path":
            return t, parse_distance_label_from_text(t), parse_query_nodes_from_text(t)
        return t, parse_yes_no_from_text(t), None

    if isinstance(rec, str):
        return from_text(rec.strip())
    if isinstance(rec, dict):
        text = rec.get("text") or rec.get("sequence")
        if text is None and isinstance(rec.get("tokens"), (list, tuple)):
            text = " ".join(map(str, rec["tokens"]))
        lab, query = rec.get("label"), None
        if task == "shortest_path":
            if not isinstance(lab, int) and isinstance(text, str):
                lab = parse_distance_label_from_text(text)
                query = parse_query_nodes_from_text(text)
        else:
            if isinstance(lab, str):
                low = lab.lower().strip()
                lab = 1 if low in _YES else (0 if low in _NO else None)
            elif isinstance(lab, (int, bool)):
                lab = int(bool(lab))
            if isinstance(text, str) and lab is None:
                lab = parse_yes_no_from_text(text)
        return (text.strip() if isinstance(text, str) else None), lab, query
    if isinstance(rec, list):
        if all(isinstance(x, (str, int)) for x in rec):
            return from_text(" ".join(map(str, rec)))
    return None, None, None


# --------------------------------------------------------------------------- JSON loading (host ingest)
def _records_of_file(fp: str):
    with open(fp, "r") as f:
        raw = f.read().strip()
    if not raw:
        return
    try:
        obj = json.loads(raw)
        yield from (obj if isinstance(obj, list) else [obj])
        return
    except json.JSONDecodeError:
        pass
    for line in raw.splitlines():
        line = line.strip()
        if not line:
            continue
        try:
            obj = json.loads(line)
            yield from (obj if isinstance(obj, list) else [obj])
        except json.JSONDecodeError:
            yield line


def load_examples(path_glob: str, task: str = "cycle_check", data_fraction: float = 1.0, seed: int = 0,
                  num_graphs: Optional[int] = None, num_pairs_per_graph: Optional[int] = None) -> List[Dict[str, Any]]:
    """graph-token JSON files -> [{'text','label'[,'query_u','query_v']}] with the reference's
    sampling semantics (reference :112-245): sorted glob, random.Random(seed).sample of files,
    per-file pair sampling for shortest_path, deprecated data_fraction."""
    import random
    files = sorted(glob(path_glob))
    if num_graphs is not None and len(files) > num_graphs:
        files = sorted(random.Random(seed).sample(files, num_graphs))
        print(f"[load_examples] Sampled {num_graphs}/{len(sorted(glob(path_glob)))} graph files")
    out: List[Dict[str, Any]] = []
    per_file = task == "shortest_path" and num_pairs_per_graph is not None
    rng = random.Random(seed)
    for fp in files:
        bucket = []
        for rec in _records_of_file(fp):
            text, lab, query = _extract_text_and_label(rec, task=task)
            if not text:
                continue
            if per_file and query is None:
                continue
            ex = {"text": text, "label": lab}
            if query is not None:
                ex["query_u"], ex["query_v"] = query
            bucket.append(ex)
        if per_file and len(bucket) > num_pairs_per_graph:
            bucket = rng.sample(bucket, num_pairs_per_graph)
        out.extend(bucket)
    if per_file:
        print(f"[load_examples] Loaded {len(out)} pairs from {len(files)} graphs (target: {num_pairs_per_graph} pairs/graph)")
        return out
    if num_graphs is None and data_fraction < 1.0 and out:
        out = random.Random(seed).sample(out, max(1, int(len(out) * data_fraction)))
    return out


def resolve_split_globs(root: str, task: str, algorithm: str, use_split_tasks_dirs: bool = True):
    """(train_glob, val_glob, test_glob) for one algorithm (reference :499-523)."""
    tr = os.path.join(root, "tasks_train", task, algorithm)
    te = os.path.join(root, "tasks_test", task, algorithm)
    a = (os.path.join(tr, "train", "*.json"), os.path.join(te, "val", "*.json"), os.path.join(te, "test", "*.json"))
    base = os.path.join(root, "tasks", task, algorithm)
    bb = tuple(os.path.join(base, s, "*.json") for s in ("train", "val", "test"))
    has = lambda g: len(glob(g)) > 0
    if use_split_tasks_dirs and has(a[0]):
        train_glob, val_glob, test_glob = a
    elif has(bb[0]):
        train_glob, val_glob, test_glob = bb
    else:
        train_glob, val_glob, test_glob = a
    if not has(val_glob):
        val_glob = a[2] if use_split_tasks_dirs else bb[2]
    return train_glob, val_glob, test_glob


def resolve_multi_algorithm_globs(root: str, task: str, train_algorithms: List[str], test_algorithm: str,
                                  use_split_tasks_dirs: bool = True):
    """(train_globs, val_globs, test_glob) (reference :526-588)."""
    has = lambda g: len(glob(g)) > 0
    train_globs, val_globs = [], []
    for algo in train_algorithms:
        tr = os.path.join(root, "tasks_train", task, algo)
        te = os.path.join(root, "tasks_test", task, algo)
        base = os.path.join(root, "tasks", task, algo)
        ta, va = os.path.join(tr, "train", "*.json"), os.path.join(te, "val", "*.json")
        tb, vb = os.path.join(base, "train", "*.json"), os.path.join(base, "val", "*.json")
        if use_split_tasks_dirs and has(ta):
            train_globs.append(ta); val = va
        elif has(tb):
            train_globs.append(tb); val = vb
        else:
            train_globs.append(ta); val = va
        if not has(val):
            val = os.path.join(te, "test", "*.json") if use_split_tasks_dirs else os.path.join(base, "test", "*.json")
        val_globs.append(val)
    ta = os.path.join(root, "tasks_test", task, test_algorithm, "test", "*.json")
    tb = os.path.join(root, "tasks", task, test_algorithm, "test", "*.json")
    test_glob = ta if (use_split_tasks_dirs and has(ta)) else (tb if has(tb) else ta)
    return train_globs, val_globs, test_glob


def load_examples_multi_algorithm(root: str, task: str, algorithms: List[str], split: str,
                                  use_split_tasks_dirs: bool = True, seed: int = 0,
                                  num_graphs: Optional[int] = None,
                                  num_pairs_per_graph: Optional[int] = None) -> List[Dict[str, Any]]:
    """Concatenate load_examples over algorithms, seed salted with hash(algo) % 10000 exactly
    as the reference does (reference :590-633; SURVEY C6: fix PYTHONHASHSEED for repeatability)."""
    all_examples: List[Dict[str, Any]] = []
    for algo in algorithms:
        if use_split_tasks_dirs:
            base = os.path.join(root, "tasks_test" if split in ("val", "test") else "tasks_train", task, algo)
        else:
            base = os.path.join(root, "tasks", task, algo)
        split_dir = os.path.join(base, split)
        if split == "val" and len(glob(os.path.join(split_dir, "*.json"))) == 0:
            split_dir = os.path.join(base, "test")
        ex = load_examples(os.path.join(split_dir, "*.json"), task=task, seed=seed + hash(algo) % 10000,
                           num_graphs=num_graphs, num_pairs_per_graph=num_pairs_per_graph)
        print(f"  [{algo}] Loaded {len(ex)} examples")
        all_examples.extend(ex)
    print(f"[load_examples_multi_algorithm] Total: {len(all_examples)} examples from {len(algorithms)} algorithms")
    return all_examples


def _target_size(sizes, strategy):
    import numpy as np
    if strategy == "undersample":
        return min(sizes)
    if strategy == "median":
        return int(np.median(sizes))
    if strategy == "oversample":
        return max(sizes)
    if strategy == "soft_oversample":
        return int(np.mean(sizes))
    raise ValueError(f"Unknown balancing strategy: {strategy}. Choose from: undersample, median, oversample, soft_oversample")


def balance_classes(examples: List[Dict[str, Any]], strategy: str = "undersample", seed: int = 0):
    """Per-class down/up-sampling to a common size, then shuffle (reference :248-330)."""
    import random
    groups: Dict[Any, list] = {}
    for ex in examples:
        if ex.get("label") is not None:
            groups.setdefault(ex["label"], []).append(ex)
    if not groups:
        return examples
    target = _target_size([len(v) for v in groups.values()], strategy)
    rng = random.Random(seed)
    out = []
    for label in sorted(groups):
        exs = groups[label]
        if len(exs) > target:
            out.extend(rng.sample(exs, target))
        elif len(exs) < target:                       # any strategy: smaller classes are topped up with replacement
            out.extend(exs + rng.choices(exs, k=target - len(exs)))
        else:
            out.extend(exs)
    rng.shuffle(out)
    return out


def get_balanced_indices(dataset, strategy: str = "undersample", seed: int = 0) -> List[int]:
    """Index version of balance_classes for PyG-style datasets (reference :333-405)."""
    import random
    groups: Dict[Any, list] = {}
    for i in range(len(dataset)):
        y = dataset[i].y
        groups.setdefault(y.item() if hasattr(y, "item") else y, []).append(i)
    if not groups:
        return list(range(len(dataset)))
    if strategy not in ("undersample", "median"):
        raise ValueError(f"Unknown balancing strategy: {strategy}")
    target = _target_size([len(v) for v in groups.values()], strategy)
    rng = random.Random(seed)
    out: List[int] = []
    for label in sorted(groups):
        idx = groups[label]
        out.extend(idx if len(idx) <= target else rng.sample(idx, target))
    rng.shuffle(out)
    return out


def determine_num_classes(examples: List[Dict[str, Any]], task: str) -> int:
    """2 for cycle_check, 1 for zinc, max int label + 1 otherwise (reference :636-688)."""
    if task == "cycle_check":
        return 2
    if task == "zinc":
        return 1
    best = -1
    for ex in examples:
        lab = ex.get("label")
        if lab is not None and isinstance(lab, int):
            best = max(best, lab)
    return best + 1


def determine_num_classes_pyg(dataset, task: str) -> int:
    """Same, from data.y of a PyG-style dataset (reference :691-738)."""
    if task == "cycle_check":
        return 2
    if task == "zinc":
        return 1
    best = -1
    for i in range(len(dataset)):
        d = dataset[i]
        if hasattr(d, "y") and d.y is not None:
            best = max(best, d.y.item())
    return best + 1


# --------------------------------------------------------------------------- vocab (host)
def build_vocab_from_texts(texts: List[str], min_freq: int = 1, max_tokens: Optional[int] = None):
    """SPECIAL first, then descending count with first-seen tie-break, capped at max_tokens
    (reference :451-463).  Counter.most_common is a stable sort over insertion order."""
    counts: Counter = Counter()
    for t in texts:
        counts.update(t.split())
    vocab = {tok: i for i, tok in enumerate(SPECIAL)}
    nxt = len(vocab)
    for tok, c in counts.most_common():
        if tok in vocab:
            continue
        if c < min_freq:
            break
        vocab[tok] = nxt
        nxt += 1
        if max_tokens and nxt >= max_tokens:
            break
    return vocab, {i: t for t, i in vocab.items()}


# --------------------------------------------------------------------------- graph-token text -> structure (host ingest)
def graph_of_text(text: str):
    """Inverse of the graph-token serializer for ONE text:
    -> (num_nodes, [(u,v),...], task, query) or None when the text is not of the grammar
    '<bos> (u v <e>)* <n> 0 1 .. N-1 <q> (has_cycle | shortest_distance u v) <p> ...'."""
    t = text.split()
    n = len(t)
    if n < 5 or t[0] != "<bos>":
        return None
    i, edges = 1, []
    try:
        while i + 2 < n and t[i + 2] == "<e>":
            edges.append((int(t[i]), int(t[i + 1])))
            i += 3
        if t[i] != "<n>":
            return None
        i += 1
        k = 0
        while i < n and t[i] != "<q>":
            if int(t[i]) != k:
                return None
            k += 1
            i += 1
        if i >= n:
            return None
        i += 1
        if t[i] == "has_cycle" and t[i + 1] == "<p>":
            return k, edges, "cycle_check", None
        if t[i] == "shortest_distance" and t[i + 3] == "<p>":
            return k, edges, "shortest_path", (int(t[i + 1]), int(t[i + 2]))
    except (ValueError, IndexError):
        return None
    return None


# --------------------------------------------------------------------------- TokenDataset / collate
# host storage ptr -> dataset (lets collate find the device rows).  Weak: a dataset (pinned copy + resident device rows)
# lives exactly as long as the trainer holds it, and a recycled pointer can never resolve to a dead dataset.
_REGISTRY: "weakref.WeakValueDictionary[int, TokenDataset]" = weakref.WeakValueDictionary()


class TokenDataset(Dataset):
    """Reference signature and item contract (reference :465-486):
        ds.seqs[i] LongTensor[L_i], ds.labels[i] LongTensor[], ds[i] -> (seq, label)
    Rows are produced by the CUDA tokenizers.  `seqs` are views into ONE pinned host copy of the
    ragged buffer (cheap to hand to the unchanged trainer); the device copy stays resident in
    `self.ragged` and is what collate() pads from."""

    def __init__(self, examples, vocab, max_len=512, strip_label=True, require_label=True, device=None):
        from gtok import engine
        from gtok.batch import GraphBatch
        if not torch.cuda.is_available():
            raise RuntimeError("TokenDataset needs a CUDA device: the gtok engine has no CPU fallback")
        self.vocab, self.max_len = vocab, max_len
        self.device = torch.device(device or "cuda")
        kept, labels = [], []
        for ex in examples:                                   # label filter, reference :471-477
            text, label = ex["text"], ex.get("label")
            if label is None and isinstance(text, str):
                label = parse_yes_no_from_text(text)
            if require_label and label is None:
                continue
            kept.append(ex)
            labels.append(int(label) if label is not None else 0)        # int(): floats truncate toward 0 (C4)
        self.labels = [torch.tensor(l, dtype=torch.long) for l in labels]
        zinc = [ex.get("_zinc") for ex in kept]
        if kept and all(z is not None for z in zinc):
            src_ds = zinc[0][0]
            if any(z[0] is not src_ds for z in zinc):
                raise ValueError("examples mix several ZINC datasets; build one TokenDataset per dataset")
            idx = torch.as_tensor([z[1] for z in zinc], dtype=torch.int64)
            gb = src_ds.graph_batch.select(idx).to(self.device)
            self.ragged = engine.tokenize_zinc(gb, vocab, max_len, item_max_len=src_ds.max_len)
        else:
            if kept and all(ex.get("_graph") is None for ex in kept):
                # plain strings (what load_examples returns): the texts are parsed on the GPU too
                gb, task = engine.graphs_from_texts([ex["text"] for ex in kept], self.device)
            else:
                parsed = [ex.get("_graph") or graph_of_text(ex["text"]) for ex in kept]
                if any(p is None for p in parsed):
                    bad = next(ex["text"] for ex, p in zip(kept, parsed) if p is None)
                    raise ValueError("text is not graph-token grammar and carries no graph handle; the CUDA "
                                     f"tokenizers cannot encode it (no CPU fallback): {bad[:80]!r}")
                tasks = {p[2] for p in parsed}
                if len(tasks) > 1:
                    raise ValueError(f"mixed tasks in one TokenDataset: {sorted(tasks)}")
                task = tasks.pop() if tasks else "cycle_check"
                gb = GraphBatch.from_lists([p[0] for p in parsed], [p[1] for p in parsed],
                                           queries=[p[3] for p in parsed] if task == "shortest_path" else None).to(self.device)
            self.ragged = engine.tokenize_ibtt(gb, vocab, task, max_len)
        if not strip_label and kept:
            # reference :478-482 without the strip: the tokens after <p> (label, <eos>) stay in the row.  The kernels emit
            # the stripped rows; the tails - two or three ids per text - are looked up here and re-attached on the device
            pad = vocab["<pad>"]
            tails = []
            for ex in kept:
                t = ex["text"]
                i = t.find("<p>")
                tails.append([vocab.get(w, pad) for w in t[i + 3:].split()] if i >= 0 else [])
            K = max(1, max(len(t) for t in tails))
            tail = torch.tensor([t + [pad] * (K - len(t)) for t in tails], dtype=torch.int64, device=self.device)
            tlen = torch.tensor([len(t) for t in tails], dtype=torch.int32, device=self.device)
            self.ragged = engine.ragged_append(self.ragged, tail, tlen, max_len)
        host = torch.empty(self.ragged.tokens.numel(), dtype=torch.int64).pin_memory()
        host.copy_(self.ragged.tokens)
        lens = self.ragged.lengths.cpu().tolist()
        self.seqs = list(torch.split(host, lens)) if lens else []
        self._host = host
        _REGISTRY[host.data_ptr()] = self

    def __len__(self):
        return len(self.seqs)

    def __getitem__(self, idx):
        return self.seqs[idx], self.labels[idx]


class _DeferredBatch:
    """What collate() hands back from INSIDE a DataLoader worker (configs/ibtt_graph_token.yaml:27 and ibtt_zinc.yaml:23 set
    num_workers: 2): a worker must not touch CUDA, so it ships only the row handles (which dataset, where each row starts,
    how long it is) and the batch is padded by K4 in the trainer's process at its first `.to(device)`
    (trainer/train_ibtt.py:144,164).  No token is copied or padded on the host."""

    def __init__(self, base: int, starts: torch.Tensor, lens: torch.Tensor, pad_id: int):
        self.base, self.starts, self.lens, self.pad_id = int(base), starts, lens, int(pad_id)
        self._out = None

    def materialize(self):
        if self._out is None:
            from gtok import engine
            ds = _REGISTRY.get(self.base)
            if ds is None:
                raise RuntimeError("deferred batch: its TokenDataset is gone (the dataset must outlive its DataLoader)")
            dev = ds.device
            self._out = engine.pad_mask(ds.ragged.tokens, self.starts.to(dev), self.lens.to(dev), self.pad_id,
                                        lmax=int(self.lens.max()))
        return self._out


class _DeferredTensor:
    """One member (0 = X, 1 = attn) of a _DeferredBatch; `.to(device)` returns the CUDA tensor."""

    def __init__(self, batch: _DeferredBatch, which: int):
        self.batch, self.which = batch, which

    def to(self, *args, **kwargs):
        return self.batch.materialize()[self.which].to(*args, **kwargs)

    def cuda(self, *args, **kwargs):
        return self.batch.materialize()[self.which]


def collate(batch, pad_id: int):
    """X[b,:L_b] = row, X[b,L_b:] = pad_id, attn[b,:L_b] = True, Y long - computed by K4 on the
    GPU and returned as CUDA tensors (the trainer's .to(device) becomes a no-op).  Reference :488-497.
    Inside a DataLoader worker (num_workers > 0) the rows' handles are returned instead and K4 runs at the trainer's
    `.to(device)` (_DeferredBatch)."""
    from gtok import engine
    xs, ys = zip(*batch)
    if torch.utils.data.get_worker_info() is not None:
        base = xs[0].untyped_storage().data_ptr() if not xs[0].is_cuda else None
        if base is None or base not in _REGISTRY or any(x.untyped_storage().data_ptr() != base for x in xs):
            raise RuntimeError("gtok collate inside a DataLoader worker needs rows of ONE gtok TokenDataset (their padding is "
                               "deferred to the trainer's process); foreign rows need num_workers=0")
        db = _DeferredBatch(base, torch.tensor([x.storage_offset() for x in xs], dtype=torch.int64),
                            torch.tensor([x.numel() for x in xs], dtype=torch.int32), pad_id)
        return _DeferredTensor(db, 0), _DeferredTensor(db, 1), torch.tensor([int(y) for y in ys], dtype=torch.long)
    ds = None
    base = xs[0].untyped_storage().data_ptr() if not xs[0].is_cuda else None
    if base is not None and base in _REGISTRY and all(x.untyped_storage().data_ptr() == base for x in xs):
        ds = _REGISTRY[base]
    dev = ds.device if ds is not None else torch.device("cuda")
    lens = torch.tensor([x.numel() for x in xs], dtype=torch.int32)
    if ds is not None:      # rows already resident in HBM: gather by row start, no copy
        starts = torch.tensor([x.storage_offset() for x in xs], dtype=torch.int64)
        tok = ds.ragged.tokens
    else:                   # foreign rows: one H2D of the concatenation
        starts = torch.zeros(len(xs), dtype=torch.int64)
        starts[1:] = lens[:-1].to(torch.int64).cumsum(0)
        tok = torch.cat([x.reshape(-1) for x in xs]).to(dev, non_blocking=True)
    X, attn = engine.pad_mask(tok, starts.to(dev), lens.to(dev), int(pad_id), lmax=int(lens.max()))
    Y = torch.tensor([int(y) for y in ys], dtype=torch.long, device=dev)
    return X, attn, Y
