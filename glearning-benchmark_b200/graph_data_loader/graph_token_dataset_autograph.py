"""Drop-in for the reference's graph_data_loader/graph_token_dataset_autograph.py.

graph-token JSON -> graph structure for the SENT tokenizer.  The reference builds a PyG
``InMemoryDataset`` of ``Data(edge_index, y, num_nodes[, query_u, query_v])`` and caches it
as ``processed/<key>/data.pt`` (reference :161-408).  Here the same files are parsed with the
same selection / sampling rules into ONE flat GraphBatch (the layout the CUDA kernels
ingest); ``ds[i]`` still returns a per-graph object with the reference's attributes, so the
unchanged trainer can read ``num_nodes``, ``y``, ``query_u/v`` and ``edge_index``.  PyG is not
required.
"""
from __future__ import annotations

import json
import os
from glob import glob
from typing import List, Optional, Tuple

import torch
from torch.utils.data import Dataset


def parse_graph_from_text(text: str) -> Tuple[List[int], List[Tuple[int, int]]]:
    """'<bos> u v <e> ... <n> 0 1 2 ...' -> (nodes, edges): every 'int int <e>' triple is an edge,
    the ints after '<n>' up to '<q>'/'<p>'/'<eos>' are the nodes (reference :14-54)."""
    t = text.split()
    n, i = len(t), 0
    nodes: List[int] = []
    edges: List[Tuple[int, int]] = []
    while i < n:
        if i + 2 < n and t[i + 2] == "<e>":
            try:
                edges.append((int(t[i]), int(t[i + 1])))
                i += 3
            except ValueError:
                i += 1
            continue
        if t[i] == "<n>" and i + 1 < n:
            for tok in t[i + 1:]:
                if tok in ("<q>", "<p>", "<eos>"):
                    break
                try:
                    nodes.append(int(tok))
                except ValueError:
                    break
            break
        i += 1
    return nodes, edges


def parse_query_nodes_from_text(text: str) -> Optional[Tuple[int, int]]:
    from .data_loader import parse_query_nodes_from_text as _p
    return _p(text)


def parse_label_from_text(text: str, task: str = 'cycle_check') -> Optional[int]:
    """'<p> yes|no' -> 1|0, '<p> lenK' -> K-1, INF -> None (reference :80-113)."""
    t = text.split()
    for i, tok in enumerate(t[:-1]):
        if tok != "<p>":
            continue
        lab = t[i + 1].upper()
        if lab in ("YES", "NO"):
            return 1 if lab == "YES" else 0
        if lab.startswith("LEN"):
            try:
                return int(lab[3:]) - 1
            except ValueError:
                pass
        if lab in ("INF", "INFINITY"):
            return None
    return None


def parse_graph_from_json(record: dict, task: str = 'cycle_check'):
    """-> (edges, num_nodes, label); 'nodes'/'edges'/'label' fields win over the text (reference :116-158)."""
    nodes = record.get('nodes', [])
    edges = record.get('edges', [])
    if not edges and record.get('text'):
        nodes, edges = parse_graph_from_text(record['text'])
    label = record.get('label')
    if label is None and record.get('text'):
        label = parse_label_from_text(record['text'], task=task)
    if nodes:
        num_nodes = max(nodes) + 1
    elif edges:
        num_nodes = max(max(s, t) for s, t in edges) + 1
    else:
        num_nodes = 0
    return edges, num_nodes, label


class GraphData:
    """Per-graph view with the attributes the reference's Data objects expose."""
    __slots__ = ("edge_index", "y", "num_nodes", "query_u", "query_v", "_gtok")

    def __init__(self, edge_index, y, num_nodes, query=None, handle=None):
        self.edge_index, self.y, self.num_nodes = edge_index, y, num_nodes
        self._gtok = handle
        if query is not None:
            self.query_u, self.query_v = query

    def to(self, *a, **k):
        return self


class GraphTokenDatasetForAutoGraph(Dataset):
    """Constructor and semantics of the reference class (:168-181, :259-408): algorithms may be a
    string or a list, files are sorted and sampled with random.Random(seed + hash(algo) % 10000),
    val falls back to test, shortest_path pairs are sampled per file with random.Random(seed),
    empty graphs / unlabeled records are skipped, RuntimeError when nothing is found."""

    def __init__(self, root: str, task: str = 'cycle_check', algorithm=None, split: str = 'train',
                 use_split_tasks_dirs: bool = True, seed: int = 0, num_graphs: Optional[int] = None,
                 num_pairs_per_graph: Optional[int] = None, transform=None, pre_transform=None, pre_filter=None):
        self.task = task
        self.algorithms = [algorithm] if isinstance(algorithm, str) else (['er'] if algorithm is None else list(algorithm))
        self.algorithm = self.algorithms[0]
        self.split, self.use_split_tasks_dirs, self.seed = split, use_split_tasks_dirs, seed
        self.num_graphs, self.num_pairs_per_graph = num_graphs, num_pairs_per_graph
        self._root, self.transform, self.pre_transform, self.pre_filter = root, transform, pre_transform, pre_filter
        self._records = self._load_or_process()
        self._gb = None

    # ------------------------------------------------------------------
    def _split_dir(self, algo: str) -> str:
        if self.use_split_tasks_dirs:
            base = os.path.join(self._root, 'tasks_test' if self.split in ('val', 'test') else 'tasks_train', self.task, algo)
        else:
            base = os.path.join(self._root, 'tasks', self.task, algo)
        d = os.path.join(base, self.split)
        if self.split == 'val' and self.use_split_tasks_dirs and len(glob(os.path.join(d, '*.json'))) == 0:
            d = os.path.join(base, 'test')
        return d

    @property
    def raw_dir(self) -> str:
        return self._split_dir(self.algorithm)

    @property
    def processed_dir(self) -> str:
        name = f"autograph_{self.task}_{'+'.join(sorted(self.algorithms))}_{self.split}"
        if self.use_split_tasks_dirs:
            name += "_split"
        if self.num_graphs is not None:
            name += f"_ng{self.num_graphs}"
        if self.num_pairs_per_graph is not None:
            name += f"_np{self.num_pairs_per_graph}"
        return os.path.join(self._root, 'processed', name)

    def _load_or_process(self):
        """The reference is an InMemoryDataset: process() runs once, later constructions load processed/<key>/data.pt
        (reference :233-253, :406-408).  Same contract here: the parsed records (node count, edge list, label, query) are
        cached under the same directory name (file `gtok_records.pt`: the reference's own `data.pt` holds collated PyG
        tensors and must stay readable by the reference)."""
        path = os.path.join(self.processed_dir, 'gtok_records.pt')
        if os.path.exists(path):
            try:
                return torch.load(path, weights_only=False)
            except Exception:                       # unreadable / truncated cache: rebuild it
                pass
        recs = self._process()
        try:
            os.makedirs(self.processed_dir, exist_ok=True)
            tmp = path + f".{os.getpid()}.tmp"
            torch.save(recs, tmp)
            os.replace(tmp, path)
        except OSError:                             # read-only dataset root: keep working without the cache
            pass
        return recs

    def _process(self):
        import random
        files: List[str] = []
        for algo in self.algorithms:
            fs = sorted(glob(os.path.join(self._split_dir(algo), '*.json')))
            if self.num_graphs is not None and len(fs) > self.num_graphs:
                fs = sorted(random.Random(self.seed + hash(algo) % 10000).sample(fs, self.num_graphs))
            files.extend(fs)
        if not files:
            raise RuntimeError(f"No JSON files found for algorithms {self.algorithms}. "
                               f"Did you run the graph-token task generator?")
        rng = random.Random(self.seed)
        out = []
        per_file = self.task == 'shortest_path' and self.num_pairs_per_graph is not None
        for fp in files:
            with open(fp, 'r') as f:
                content = json.load(f)
            records = content if isinstance(content, list) else [content]
            bucket = []
            for rec in records:
                edges, n, label = parse_graph_from_json(rec, task=self.task)
                if n == 0 or label is None:
                    continue
                query = None
                if self.task == 'shortest_path' and rec.get('text'):
                    query = parse_query_nodes_from_text(rec['text'])
                if per_file and query is None:
                    continue
                if self.pre_filter is not None:       # reference :350-351 / :396-397: the filter sees the Data object
                    el = [tuple(e) for e in edges]
                    ei = torch.tensor(el, dtype=torch.long).t().contiguous() if el else torch.empty((2, 0), dtype=torch.long)
                    if not self.pre_filter(GraphData(ei, torch.tensor([int(label)], dtype=torch.long), n, query)):
                        continue
                bucket.append((n, [tuple(e) for e in edges], int(label), query))
            if per_file and len(bucket) > self.num_pairs_per_graph:
                bucket = rng.sample(bucket, self.num_pairs_per_graph)
            out.extend(bucket)
        return out

    # ------------------------------------------------------------------
    @property
    def graph_batch(self):
        """The whole split as one flat GraphBatch (host side)."""
        if self._gb is None:
            from gtok.batch import GraphBatch
            r = self._records
            has_q = bool(r) and all(x[3] is not None for x in r)
            self._gb = GraphBatch.from_lists([x[0] for x in r], [x[1] for x in r],
                                             queries=[x[3] for x in r] if has_q else None,
                                             y=[x[2] for x in r])
        return self._gb

    def __len__(self):
        return len(self._records)

    def __getitem__(self, idx):
        n, edges, label, query = self._records[idx]
        ei = torch.tensor(edges, dtype=torch.long).t().contiguous() if edges else torch.empty((2, 0), dtype=torch.long)
        d = GraphData(ei, torch.tensor([label], dtype=torch.long), n, query, (self, int(idx)))
        if self.pre_transform is not None:
            d = self.pre_transform(d)
        if self.transform is not None:
            d = self.transform(d)
        return d
