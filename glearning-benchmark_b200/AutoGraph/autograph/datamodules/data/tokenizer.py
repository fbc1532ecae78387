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
an item belongs to in one launch and serves rows from that buffer; asking for an item that was
already served starts a new epoch (new Philox seed, all rows re-sampled), which reproduces the
reference's "fresh random trail every epoch" behaviour.
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
        if not append_eos:
            raise NotImplementedError("append_eos=False is not supported")
        self.dataset_names = list(dataset_names or [])
        self.max_length = max_length
        self.truncation_length = truncation_length if truncation_length is not None else max_length
        self.labeled_graph = labeled_graph
        self.undirected = undirected
        self.idx_offset = len(self.special_toks) + len(self.dataset_names)
        self.max_num_nodes: Optional[int] = None
        self.num_node_types = self.num_edge_types = 0
        self.node_idx_offset = self.edge_idx_offset = None
        self.seed = int(seed)
        self.edge_label_shift = 0
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
                                 labeled=self.labeled_graph, edge_label_shift=self.edge_label_shift)

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
        keep = torch.arange(tok.shape[1], device=tok.device)[None, :] < ln[:, None]
        flat = tok[keep].cpu()
        lens = ln.cpu().to(torch.int64)
        off = torch.zeros(lens.numel() + 1, dtype=torch.int64)
        off[1:] = lens.cumsum(0)
        st["flat"], st["off"] = flat, off.tolist()
        st["served"] = bytearray(lens.numel())

    def __call__(self, data) -> torch.Tensor:
        handle = getattr(data, "_gtok", None)
        if handle is not None:
            ds, idx = handle
            st = self._epoch(ds)
            if st["served"] is None or st["served"][idx]:
                self._sample(st)                       # first use, or this row was already served: new epoch
            st["served"][idx] = 1
            return st["flat"][st["off"][idx]:st["off"][idx + 1]].clone()
        # stand-alone object (no dataset handle): one-graph launch
        from gtok.batch import GraphBatch
        gb = GraphBatch.from_data_list([data]).to(self._device())
        tok, ln = self.tokenize_batch(gb, seed=self.seed)
        self.seed += 1
        return tok[0, :int(ln[0])].cpu()

    tokenize = __call__
