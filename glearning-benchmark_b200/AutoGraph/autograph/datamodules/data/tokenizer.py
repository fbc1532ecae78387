"""Graph2TrailTokenizer - B200-native stand-in for AutoGraph's SENT tokenizer.

The reference imports this class from the un-vendored BorgwardtLab/AutoGraph checkout it
expects at ``<repo>/AutoGraph`` (trainer/train_agtt.py:17-21) and uses exactly this surface:
ctor kwargs (:514-530), ``set_num_nodes`` (:535), ``set_num_node_and_edge_types`` (:540),
``idx_offset / node_idx_offset / edge_idx_offset`` (:189-191), the class attribute ``pad``
(:286) and ``__call__(data) -> 1-D LongTensor`` (:250).  Put ``glearning-benchmark_b200/AutoGraph``
on sys.path (the reference already inserts ``<repo>/AutoGraph``) and the trainer runs unchanged.

Tokenization follows the spec in oracle/sent_oracle.py (parity to AutoGraph itself is UNPINNED,
see DESIGN.md) and runs in the K3 CUDA kernel.  The reference calls the tokenizer once per
item per epoch from a Python loop; to keep the GPU busy this class tokenizes the WHOLE dataset
an item belongs to in one launch and serves rows from that buffer; a new epoch (new Philox seed,
all rows re-sampled) starts on new_epoch() or when a row is asked for again after at least half
of the dataset has been served - the reference's "fresh random trail every epoch" behaviour,
without re-tokenizing everything for an early peek at one item.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class Graph2TrailTokenizer:
    sos, reset, ladj, radj, eos, pad = 0, 1, 2, 3, 4, 5
    special_toks = ["sos", "reset", "ladj", "radj", "eos", "pad"]

    def __init__(self, dataset_names=None, max_length: int = -1, truncation_length: Optional[int] = None,
                 labeled_graph: bool = False, undirected: bool = True, append_eos: bool = True, seed: int = 0,
                 device=None, **kwargs):
        if not undirected:
            raise NotImplementedError("gtok's SENT kernel covers undirected graphs (the only mode the reference uses)")
        self.dataset_names = list(dataset_names or [])
        self.max_length = max_length
        self.truncation_length = truncation_length if truncation_length is not None else max_length
        self.labeled_graph = labeled_graph
        self.undirected = undirected
        self.append_eos = bool(append_eos)
        self.idx_offset = len(self.special_toks) + len(self.dataset_names)
        self.max_num_nodes: Optional[int] = None
        self.num_node_types = self.num_edge_types = 0
        self.node_idx_offset = self.edge_idx_offset = None
        self.seed = int(seed)
        self.edge_label_shift = 0
        # the open decisions of the specification the kernels implement (oracle.sent_oracle.SentSpec; a pin produced by
        # tools/pin_autograph.py says which setting reproduces the real tokenizer)
        self.restart = kwargs.pop("restart", "open_first")
        self.bracket_order = kwargs.pop("bracket_order", "visit_asc")
        self.device = torch.device(device) if device is not None else None
        self._cache: Dict[int, dict] = {}

    # ---- configuration ---------------------------------------------------------------------
    def set_num_nodes(self, max_num_nodes: int) -> None:
        if self.max_num_nodes is None or self.max_num_nodes < max_num_nodes:
            self.max_num_nodes = int(max_num_nodes)
            if self.num_node_types:
                self.set_num_node_and_edge_types(self.num_node_types, self.num_edge_types)
            self._cache.clear()

    def set_num_node_and_edge_types(self, num_node_types: int, num_edge_types: int) -> None:
        if self.max_num_nodes is None:
            raise ValueError("call set_num_nodes() before set_num_node_and_edge_types()")
        self.num_node_types, self.num_edge_types = int(num_node_types), int(num_edge_types)
        self.node_idx_offset = self.idx_offset + self.max_num_nodes
        self.edge_idx_offset = self.node_idx_offset + self.num_node_types
        self._cache.clear()

    def __len__(self) -> int:
        n = self.idx_offset + (self.max_num_nodes or 0)
        return n + self.num_node_types + self.num_edge_types if self.labeled_graph else n

    # ---- engine ------------------------------------------------------------------------------
    def _limit(self) -> int:
        lim = self.truncation_length if self.truncation_length and self.truncation_length > 0 else self.max_length
        if lim is None or lim <= 0:
            raise ValueError("gtok needs a positive max_length / truncation_length")
        return int(lim)

    def _config(self):
        from gtok import engine
        if self.labeled_graph and self.node_idx_offset is None:
            raise ValueError("labeled_graph=True needs set_num_nodes() and set_num_node_and_edge_types() first")
        return engine.SentConfig(idx_offset=self.idx_offset, node_idx_offset=self.node_idx_offset or 0,
                                 edge_idx_offset=self.edge_idx_offset or 0, max_length=self._limit(),
                                 labeled=self.labeled_graph, edge_label_shift=self.edge_label_shift,
                                 append_eos=self.append_eos, restart=self.restart, bracket_order=self.bracket_order)

    def tokenize_batch(self, graph_batch, seed: Optional[int] = None, graph_id_base: int = 0, csr=None):
        """Native entry point: GraphBatch (CUDA) -> (tokens int64[G, stride], lengths int32[G])."""
        from gtok import engine
        if csr is None:
            csr = engine.build_csr(graph_batch)
        return engine.tokenize_sent(csr, self._config(), seed=self.seed if seed is None else seed,
                                    graph_id_base=graph_id_base)

    def _device(self):
        if self.device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("Graph2TrailTokenizer needs a CUDA device: the gtok engine has no CPU fallback")
            self.device = torch.device("cuda")
        return self.device

    def _epoch(self, ds) -> dict:
        from gtok import engine
        st = self._cache.get(id(ds))
        if st is None:
            gb = ds.graph_batch.to(self._device())
            st = {"csr": engine.build_csr(gb), "epoch": -1, "served": None, "ds": ds}
            self._cache[id(ds)] = st
        return st

    def _sample(self, st) -> None:
        from gtok import engine
        st["epoch"] += 1
        seed = (self.seed * 0x9E3779B97F4A7C15 + st["epoch"]) & (2 ** 64 - 1)
        tok, ln = engine.tokenize_sent(st["csr"], self._config(), seed=seed)
        # only the valid tokens cross PCIe: rows are compacted on the device (stride = max_length, mean row ~10 % of it)
        flat, off = engine.compact_rows(tok, ln)
        st["flat"], st["off"] = flat.cpu(), off.cpu().tolist()
        st["served"] = bytearray(int(ln.numel()))
        st["nserved"] = 0

    def new_epoch(self, dataset=None) -> None:
        """Explicit epoch boundary: the next item served from `dataset` (default: every dataset seen so far) comes from
        freshly sampled trails.  Trainers that can be edited call this once per epoch; the unchanged reference trainer
        relies on the automatic rule in __call__."""
        for st in self._cache.values():
            if dataset is None or st["ds"] is dataset:
                st["served"] = None

    def __call__(self, data) -> torch.Tensor:
        from torch.utils.data import get_worker_info
        if get_worker_info() is not None:
            raise RuntimeError("Graph2TrailTokenizer (gtok) runs CUDA kernels and cannot be called from a DataLoader worker "
                               "process: set num_workers: 0 (the reference's AGTT configs already do, configs/agtt_*.yaml)")
        handle = getattr(data, "_gtok", None)
        if handle is not None:
            ds, idx = handle
            st = self._epoch(ds)
            # Epoch rule: the reference draws a fresh trail whenever an item is read (train_agtt.py:246-250), i.e. every row
            # is new in every epoch.  A row asked for a second time starts a new epoch (all rows re-sampled in one launch)
            # once the current one has served at least half of the dataset; an early re-read (a peek such as
            # log_graph_examples, a sanity batch) is answered from the current buffer instead of re-tokenizing everything.
            if st["served"] is None:
                self._sample(st)
            elif st["served"][idx] and 2 * st["nserved"] >= len(st["served"]):
                self._sample(st)
            if not st["served"][idx]:
                st["served"][idx] = 1
                st["nserved"] += 1
            return st["flat"][st["off"][idx]:st["off"][idx + 1]].clone()
        # stand-alone object (no dataset handle): one-graph launch
        from gtok.batch import GraphBatch
        gb = GraphBatch.from_data_list([data]).to(self._device())
        tok, ln = self.tokenize_batch(gb, seed=self.seed)
        self.seed += 1
        return tok[0, :int(ln[0])].cpu()

    tokenize = __call__
