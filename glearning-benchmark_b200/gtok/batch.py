"""GraphBatch: the flat edge-index layout every kernel ingests (DESIGN.md "HBM layout").

A batch is the disjoint union of G graphs, stored as struct-of-arrays so that a
warp reading one graph touches contiguous int32 runs:

  node_off int64[G+1]   node k of graph g is global node node_off[g]+k
  edge_off int64[G+1]   directed entries of graph g are [edge_off[g], edge_off[g+1])
  src,dst  int32[E]     LOCAL node ids, in the caller's edge_index column order
  x        int32[N]     node type (ZINC atom index)            optional
  edge_attr int32[E]    edge type (ZINC bond index, 1-based)   optional
  query_u,query_v int32[G]                                     optional (shortest_path)
  y                     labels, never touched by a kernel

Mirrors what the reference holds per graph as PyG ``Data(edge_index, x, edge_attr,
y, num_nodes, query_u, query_v)`` (graph_data_loader/graph_token_dataset_autograph.py:
385-394, graph_data_loader/zinc_dataset_autograph.py:51-73).
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Optional, Sequence

import torch


@dataclass
class GraphBatch:
    node_off: torch.Tensor
    edge_off: torch.Tensor
    src: torch.Tensor
    dst: torch.Tensor
    x: Optional[torch.Tensor] = None
    edge_attr: Optional[torch.Tensor] = None
    query_u: Optional[torch.Tensor] = None
    query_v: Optional[torch.Tensor] = None
    y: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------
    @property
    def num_graphs(self) -> int:
        return int(self.node_off.numel()) - 1

    @property
    def device(self) -> torch.device:
        return self.src.device

    def num_nodes(self) -> torch.Tensor:
        return (self.node_off[1:] - self.node_off[:-1])

    def num_entries(self) -> torch.Tensor:
        return (self.edge_off[1:] - self.edge_off[:-1])

    def to(self, device, non_blocking: bool = False) -> "GraphBatch":
        kw = {}
        for f in fields(self):
            v = getattr(self, f.name)
            kw[f.name] = None if v is None else v.to(device, non_blocking=non_blocking)
        return GraphBatch(**kw)

    def pin_memory(self) -> "GraphBatch":
        kw = {}
        for f in fields(self):
            v = getattr(self, f.name)
            kw[f.name] = None if v is None else v.pin_memory()
        return GraphBatch(**kw)

    def input_bytes(self) -> int:
        return sum(getattr(self, f.name).numel() * getattr(self, f.name).element_size()
                   for f in fields(self) if getattr(self, f.name) is not None and f.name != "y")

    def validate(self) -> None:
        """Host-side shape/dtype contract check (the kernels trust it)."""
        G = self.num_graphs
        if self.node_off.dtype != torch.int64 or self.edge_off.dtype != torch.int64:
            raise TypeError("node_off/edge_off must be int64")
        if self.edge_off.numel() != G + 1:
            raise ValueError("edge_off must have G+1 entries")
        for name in ("src", "dst", "x", "edge_attr", "query_u", "query_v"):
            v = getattr(self, name)
            if v is not None and v.dtype != torch.int32:
                raise TypeError(f"{name} must be int32")
        if self.src.numel() != self.dst.numel():
            raise ValueError("src/dst length mismatch")
        if self.edge_attr is not None and self.edge_attr.numel() != self.src.numel():
            raise ValueError("edge_attr must have one entry per edge_index column")

    # ------------------------------------------------------------------
    def graph(self, g: int):
        """Per-graph views: (num_nodes, src, dst, x|None, edge_attr|None)."""
        n0, n1 = int(self.node_off[g]), int(self.node_off[g + 1])
        e0, e1 = int(self.edge_off[g]), int(self.edge_off[g + 1])
        return (n1 - n0, self.src[e0:e1], self.dst[e0:e1],
                None if self.x is None else self.x[n0:n1],
                None if self.edge_attr is None else self.edge_attr[e0:e1])

    def slice(self, g0: int, g1: int) -> "GraphBatch":
        """Contiguous shard [g0, g1) - how ranks partition a dataset."""
        n0, n1 = int(self.node_off[g0]), int(self.node_off[g1])
        e0, e1 = int(self.edge_off[g0]), int(self.edge_off[g1])
        opt = lambda t, a, b: None if t is None else t[a:b]
        return GraphBatch(self.node_off[g0:g1 + 1] - n0, self.edge_off[g0:g1 + 1] - e0,
                          self.src[e0:e1], self.dst[e0:e1], opt(self.x, n0, n1),
                          opt(self.edge_attr, e0, e1), opt(self.query_u, g0, g1),
                          opt(self.query_v, g0, g1), opt(self.y, g0, g1))

    def select(self, idx: torch.Tensor) -> "GraphBatch":
        """Gather the graphs idx[0], idx[1], ... into a new batch (any order, repeats allowed).  CUDA batches go through
        the library kernels (engine.gather_graphs); host batches (tests, dataset construction) use tensor ops."""
        if self.node_off.is_cuda:
            from . import engine
            return engine.gather_graphs(self, idx.to(self.node_off.device))
        idx = idx.to(self.node_off.device, torch.int64)
        n = (self.node_off[idx + 1] - self.node_off[idx])
        e = (self.edge_off[idx + 1] - self.edge_off[idx])
        node_off = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=idx.device)
        node_off[1:] = n.cumsum(0)
        edge_off = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=idx.device)
        edge_off[1:] = e.cumsum(0)

        def gather(off_src, off_dst, cnt, arr):
            if arr is None:
                return None
            total = int(off_dst[-1])
            seg = torch.repeat_interleave(torch.arange(idx.numel(), device=idx.device), cnt)
            pos = torch.arange(total, device=idx.device) - off_dst[seg] + off_src[idx][seg]
            return arr[pos]
        opt = lambda t: None if t is None else t[idx]
        return GraphBatch(node_off, edge_off, gather(self.edge_off, edge_off, e, self.src),
                          gather(self.edge_off, edge_off, e, self.dst), gather(self.node_off, node_off, n, self.x),
                          gather(self.edge_off, edge_off, e, self.edge_attr), opt(self.query_u), opt(self.query_v),
                          opt(self.y))

    # ------------------------------------------------------------------
    @staticmethod
    def from_lists(num_nodes: Sequence[int], edge_lists: Sequence, x_lists=None, attr_lists=None,
                   queries=None, y=None) -> "GraphBatch":
        """edge_lists[g] = iterable of (u, v) (directed entries, caller order)."""
        G = len(num_nodes)
        node_off = torch.zeros(G + 1, dtype=torch.int64)
        node_off[1:] = torch.as_tensor(list(num_nodes), dtype=torch.int64).cumsum(0)
        cnt = [len(e) for e in edge_lists]
        edge_off = torch.zeros(G + 1, dtype=torch.int64)
        edge_off[1:] = torch.as_tensor(cnt, dtype=torch.int64).cumsum(0)
        flat = [p for e in edge_lists for p in e]
        if flat:
            ed = torch.as_tensor(flat, dtype=torch.int32).reshape(-1, 2)
            src, dst = ed[:, 0].contiguous(), ed[:, 1].contiguous()
        else:
            src = torch.zeros(0, dtype=torch.int32)
            dst = torch.zeros(0, dtype=torch.int32)
        cat = lambda ls: torch.as_tensor([v for l in ls for v in l], dtype=torch.int32)
        xs = None if x_lists is None else cat(x_lists)
        at = None if attr_lists is None else cat(attr_lists)
        qu = qv = None
        if queries is not None:
            qu = torch.as_tensor([q[0] for q in queries], dtype=torch.int32)
            qv = torch.as_tensor([q[1] for q in queries], dtype=torch.int32)
        yy = None if y is None else torch.as_tensor(y)
        return GraphBatch(node_off, edge_off, src, dst, xs, at, qu, qv, yy)

    @staticmethod
    def from_collated(data, slices, indices: Optional[torch.Tensor] = None) -> "GraphBatch":
        """From the storage of a PyG InMemoryDataset: `data` holds every graph's tensors concatenated
        (x [N_total, 1] or [N_total], edge_index [2, E_total] with LOCAL node ids - InMemoryDataset.collate does not
        increment them -, edge_attr [E_total(, 1)], y [G]) and `slices[key]` the G+1 boundaries per key.  No per-graph
        Python: the offsets ARE the slices.  `indices` (dataset._indices of a shuffled / subset view) gathers afterwards."""
        node_off = torch.as_tensor(slices["x"], dtype=torch.int64).clone()
        edge_off = torch.as_tensor(slices["edge_index"], dtype=torch.int64).clone()
        if node_off.numel() != edge_off.numel():
            raise ValueError("slices['x'] and slices['edge_index'] describe different graph counts")
        x = data.x
        if x.dim() == 2:
            if x.shape[1] != 1:
                raise ValueError("one integer feature per node expected (ZINC atom types)")
            x = x[:, 0]
        ei = data.edge_index
        if int(node_off[-1]) != x.shape[0] or int(edge_off[-1]) != ei.shape[1]:
            raise ValueError("slices do not cover the stored tensors")
        ea = getattr(data, "edge_attr", None)
        y = getattr(data, "y", None)
        gb = GraphBatch(node_off, edge_off, ei[0].to(torch.int32).contiguous(), ei[1].to(torch.int32).contiguous(),
                        x.to(torch.int32).contiguous(), None if ea is None else ea.reshape(-1).to(torch.int32).contiguous(),
                        None, None, None if y is None else y.reshape(-1)[: node_off.numel() - 1].clone())
        return gb if indices is None else gb.select(indices)

    @staticmethod
    def from_data_list(data_list: Sequence) -> "GraphBatch":
        """From PyG-style ``Data`` objects (duck-typed: edge_index[2,E], num_nodes,
        optional x, edge_attr, query_u, query_v, y)."""
        G = len(data_list)
        nn = [int(d.num_nodes) for d in data_list]
        node_off = torch.zeros(G + 1, dtype=torch.int64)
        node_off[1:] = torch.as_tensor(nn, dtype=torch.int64).cumsum(0)
        ne = [int(d.edge_index.shape[1]) for d in data_list]
        edge_off = torch.zeros(G + 1, dtype=torch.int64)
        edge_off[1:] = torch.as_tensor(ne, dtype=torch.int64).cumsum(0)
        if G:
            ei = torch.cat([d.edge_index.reshape(2, -1) for d in data_list], dim=1).to(torch.int32)
        else:
            ei = torch.zeros(2, 0, dtype=torch.int32)
        has = lambda a: G > 0 and getattr(data_list[0], a, None) is not None
        x = torch.cat([d.x.reshape(-1) for d in data_list]).to(torch.int32) if has("x") else None
        ea = torch.cat([d.edge_attr.reshape(-1) for d in data_list]).to(torch.int32) if has("edge_attr") else None
        qu = qv = None
        if has("query_u") and has("query_v"):
            qu = torch.as_tensor([int(d.query_u) for d in data_list], dtype=torch.int32)
            qv = torch.as_tensor([int(d.query_v) for d in data_list], dtype=torch.int32)
        y = torch.cat([d.y.reshape(-1)[:1] for d in data_list]) if has("y") else None
        return GraphBatch(node_off, edge_off, ei[0].contiguous(), ei[1].contiguous(), x, ea, qu, qv, y)


def graph_batch_from_dataset(ds) -> "GraphBatch":
    """All graphs of a PyG-style dataset as one GraphBatch.  A PyG InMemoryDataset already holds its graphs
    concatenated (`ds._data` / `ds.data` + `ds.slices`, optional `ds._indices`): that storage converts with a
    handful of tensor ops (GraphBatch.from_collated); anything else goes item by item (from_data_list)."""
    store = getattr(ds, "_data", None)
    if store is None:
        store = getattr(ds, "data", None)
    slices = getattr(ds, "slices", None)
    if store is not None and isinstance(slices, dict) and "x" in slices and "edge_index" in slices:
        idx = getattr(ds, "_indices", None)
        return GraphBatch.from_collated(store, slices, None if idx is None else torch.as_tensor(list(idx), dtype=torch.int64))

    class _V:
        pass
    items = []
    for i in range(len(ds)):
        d = ds[i]
        v = _V()
        v.x, v.edge_index, v.y = d.x, d.edge_index, getattr(d, "y", None)
        v.edge_attr = None if getattr(d, "edge_attr", None) is None else d.edge_attr.reshape(-1)
        v.num_nodes = int(d.x.shape[0])
        items.append(v)
    return GraphBatch.from_data_list(items)


def _narrowest(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Lossless narrowing of a non-negative integer array: uint8 if every value fits, else int16-as-uint16
    (stored in torch.int16 storage and reinterpreted on the device), else int32."""
    if t is None:
        return None
    if t.numel() == 0:
        return t.to(torch.uint8)
    lo, hi = int(t.min()), int(t.max())
    if lo >= 0 and hi < 256:
        return t.to(torch.uint8)
    if lo >= 0 and hi < 65536:
        return (t.to(torch.int32) & 0xFFFF).to(torch.int16)       # bit pattern of the uint16 value
    return t.to(torch.int32)


@dataclass
class PackedGraphBatch:
    """Compact host-side wire format of a GraphBatch: per-graph COUNTS instead of int64 offsets and the
    narrowest lossless unsigned type per array.  gtok_widen_i32 / gtok_counts_to_offsets (include/gtok.h)
    restore the resident int32 layout on the device.  For ZINC-shaped molecules this is 0.17 KB per graph
    over PCIe instead of 0.68 KB (the reference's own int64 tensors would be 1.3 KB)."""
    node_cnt: torch.Tensor
    entry_cnt: torch.Tensor
    src: torch.Tensor
    dst: torch.Tensor
    x: Optional[torch.Tensor] = None
    edge_attr: Optional[torch.Tensor] = None
    query_u: Optional[torch.Tensor] = None
    query_v: Optional[torch.Tensor] = None
    num_nodes_total: int = 0
    num_entries_total: int = 0
    col16: Optional[torch.Tensor] = None      # src | dst << 6 | edge_attr << 12 per column (int16 storage); replaces src/dst/edge_attr
    max_id: int = 0                           # largest local node id (the kernels size their pair bitmaps from it)
    unique_columns: bool = False              # col16 holds each undirected edge ONCE, in first-seen order (see from_batch)

    ARRAYS = ("node_cnt", "entry_cnt", "src", "dst", "x", "edge_attr", "query_u", "query_v", "col16")

    @property
    def num_graphs(self) -> int:
        return int(self.node_cnt.numel())

    @staticmethod
    def _canonical_half(b: GraphBatch) -> Optional[torch.Tensor]:
        """Mask of the src < dst columns if `b` is a canonical PyG undirected edge_index - per graph: both directions
        of every edge, no self loops, no repeats, columns sorted by (src, dst), edge_attr equal on both directions -
        else None.  In such a list the undirected pair {u, v}, u < v, is FIRST seen at column (u, v) (it sorts before
        (v, u)), so the src < dst columns in order are exactly tokenize_molecule's first-seen edges with the
        orientation and bond type it would use (zinc_dataset_indexbase.py:170-184): the other half carries nothing."""
        E = int(b.src.numel())
        if E == 0 or E % 2 or b.src.is_cuda:
            return None
        s_, d_ = b.src.to(torch.int64), b.dst.to(torch.int64)
        if bool((s_ == d_).any()):
            return None
        g = torch.repeat_interleave(torch.arange(b.num_graphs, dtype=torch.int64), b.num_entries())
        key, rkey = (g * 64 + s_) * 64 + d_, (g * 64 + d_) * 64 + s_
        if not bool((key[1:] > key[:-1]).all()):
            return None
        pos = torch.searchsorted(key, rkey).clamp_(max=E - 1)
        if not bool((key[pos] == rkey).all()):
            return None
        if b.edge_attr is not None and not bool((b.edge_attr[pos] == b.edge_attr).all()):
            return None
        return s_ < d_

    @staticmethod
    def from_batch(b: GraphBatch, fuse_columns: bool = False, drop_mirror_columns: bool = False) -> "PackedGraphBatch":
        """fuse_columns: molecule-sized graphs (local ids < 64, edge_attr in 0..15, atom types < 256) travel as ONE
        16-bit word per edge_index column (GtokPackedMolBatch, include/gtok.h) - the layout the thread-per-molecule
        ZINC kernel stages anyway; falls back to separate arrays when a value does not fit.
        drop_mirror_columns (with fuse_columns): if the batch is a canonical undirected PyG edge_index (verified
        here, see _canonical_half) only its src < dst columns travel, flagged unique_columns: half the column bytes
        over PCIe and no dedupe on the device.  Any other batch keeps all its columns."""
        if fuse_columns and b.num_graphs and b.src.numel():
            mid = max(int(b.src.max()), int(b.dst.max()))
            lo = min(int(b.src.min()), int(b.dst.min()))
            ea = b.edge_attr
            ok = lo >= 0 and mid < 64 and (ea is None or (int(ea.min()) >= 0 and int(ea.max()) < 16))
            ok = ok and b.x is not None and int(b.x.min()) >= 0 and int(b.x.max()) < 256
            if ok:
                w = b.src.to(torch.int32) | (b.dst.to(torch.int32) << 6)
                if ea is not None:
                    w = w | (ea.to(torch.int32) << 12)
                col16 = (w & 0xFFFF).to(torch.int16)
                entries = b.num_entries()
                keep = PackedGraphBatch._canonical_half(b) if drop_mirror_columns else None
                if keep is not None:
                    col16, entries = col16[keep].contiguous(), entries // 2
                return PackedGraphBatch(_narrowest(b.num_nodes()), _narrowest(entries), None, None,
                                        b.x.to(torch.uint8), None, _narrowest(b.query_u), _narrowest(b.query_v),
                                        int(b.node_off[-1]), int(col16.numel()), col16, mid, keep is not None)
        return PackedGraphBatch(_narrowest(b.num_nodes()), _narrowest(b.num_entries()), _narrowest(b.src),
                                _narrowest(b.dst), _narrowest(b.x), _narrowest(b.edge_attr), _narrowest(b.query_u),
                                _narrowest(b.query_v), int(b.node_off[-1]) if b.num_graphs else 0,
                                int(b.edge_off[-1]) if b.num_graphs else 0)

    def pin_memory(self) -> "PackedGraphBatch":
        kw = {n: (None if getattr(self, n) is None else getattr(self, n).pin_memory()) for n in self.ARRAYS}
        return PackedGraphBatch(num_nodes_total=self.num_nodes_total, num_entries_total=self.num_entries_total,
                                max_id=self.max_id, unique_columns=self.unique_columns, **kw)

    def nbytes(self) -> int:
        return sum(getattr(self, n).numel() * getattr(self, n).element_size() for n in self.ARRAYS
                   if getattr(self, n) is not None)


@dataclass
class PygStore:
    """The reference's OWN host storage of a molecule dataset: what a PyG InMemoryDataset keeps (`ds._data`, `ds.slices`)
    and what ZINCDatasetForAutoGraph / ZINCTokenizationDataset read per item (zinc_dataset_autograph.py:51-73,
    zinc_dataset_indexbase.py:63-70) - int64 tensors, LOCAL node ids, the slices are the per-graph boundaries.
    ZincIbttPipeline.run_pyg() takes this object as it is: no host-side packing, narrowing or checking."""
    edge_index: torch.Tensor       # int64 [2, E_total]
    x: torch.Tensor                # int64 [N_total] (a [N_total, 1] tensor is viewed flat)
    edge_attr: Optional[torch.Tensor]   # int64 [E_total]
    node_slices: torch.Tensor      # int64 [G+1]
    edge_slices: torch.Tensor      # int64 [G+1]

    @property
    def num_graphs(self) -> int:
        return int(self.node_slices.numel()) - 1

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.edge_index, self.x, self.edge_attr, self.node_slices, self.edge_slices) if t is not None)

    def pin_memory(self) -> "PygStore":
        p = lambda t: None if t is None else t.pin_memory()
        return PygStore(p(self.edge_index), p(self.x), p(self.edge_attr), p(self.node_slices), p(self.edge_slices))

    @staticmethod
    def from_dataset(ds) -> "PygStore":
        """The storage of a PyG InMemoryDataset, untouched (views, no copies)."""
        store = getattr(ds, "_data", None)
        if store is None:
            store = getattr(ds, "data", None)
        slices = getattr(ds, "slices", None)
        if store is None or not isinstance(slices, dict):
            raise TypeError("an InMemoryDataset (with ._data / .slices) is expected")
        if getattr(ds, "_indices", None) is not None:
            raise ValueError("the dataset is an index view; materialise it (ds.copy()) or use GraphBatch.from_collated")
        x = store.x
        ea = getattr(store, "edge_attr", None)
        return PygStore(store.edge_index.contiguous(), x.reshape(-1) if x.dim() == 2 and x.shape[1] == 1 else x,
                        None if ea is None else ea.reshape(-1), torch.as_tensor(slices["x"], dtype=torch.int64),
                        torch.as_tensor(slices["edge_index"], dtype=torch.int64))

    @staticmethod
    def from_batch(b: GraphBatch) -> "PygStore":
        """The int64 tensors PyG would hold for the graphs of `b` (synthetic benchmarks / tests)."""
        ei = torch.stack([b.src.to(torch.int64), b.dst.to(torch.int64)]).contiguous()
        return PygStore(ei, b.x.to(torch.int64), None if b.edge_attr is None else b.edge_attr.to(torch.int64),
                        b.node_off.to(torch.int64).clone(), b.edge_off.to(torch.int64).clone())
