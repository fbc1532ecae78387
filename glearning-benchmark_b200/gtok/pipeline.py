"""Whole-dataset pipelines: graphs -> ragged int64 tokens -> per-batch padded X/attn,
with every buffer preallocated so a pass is nothing but kernel launches (no host
sync, no allocation).  This is the call a user makes to tokenize an epoch:

    pipe = ZincIbttPipeline(capacity_from=host_batch, vocab=vocab, max_len=1024, batch_size=128)
    res  = pipe.run(host_batch)          # H2D + 6 kernels (+ D2H of lengths / batch extents)

Stands in for, on the reference side (trainer/train_ibtt.py:231-233, :395-402):
    [ds[i] for i in range(len(ds))] -> TokenDataset(...) -> DataLoader(collate_fn=collate)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib, engine
from ._lib import lib, check
from .batch import GraphBatch, PackedGraphBatch, PygStore
from .engine import _ptr, _stream


@dataclass
class EpochResult:
    tokens: torch.Tensor      # int64 ragged (device, capacity-sized; valid up to offsets[-1])
    offsets: torch.Tensor     # int64 [G+1]
    lengths: torch.Tensor     # int32 [G]
    padded: engine.PaddedEpoch
    host_lengths: Optional[torch.Tensor] = None   # pinned copies (e2e mode)
    host_lmax: Optional[torch.Tensor] = None
    host_xoff: Optional[torch.Tensor] = None


class _PipelineBase:
    """Shared buffer management: device copies of the graph arrays + outputs."""

    def __init__(self, G: int, n_total: int, e_total: int, tok_capacity: int, pad_capacity: int,
                 batch_size: int, device, fields: List[str]):
        self.G, self.B = int(G), int(batch_size)
        self.device = torch.device(device)
        d = self.device
        self.nb = (self.G + self.B - 1) // self.B
        i32, i64 = torch.int32, torch.int64
        self.dev_in = {"node_off": torch.empty(G + 1, dtype=i64, device=d),
                       "edge_off": torch.empty(G + 1, dtype=i64, device=d),
                       "src": torch.empty(e_total, dtype=i32, device=d),
                       "dst": torch.empty(e_total, dtype=i32, device=d)}
        if "x" in fields:
            self.dev_in["x"] = torch.empty(n_total, dtype=i32, device=d)
        if "edge_attr" in fields:
            self.dev_in["edge_attr"] = torch.empty(e_total, dtype=i32, device=d)
        if "query" in fields:
            self.dev_in["query_u"] = torch.empty(G, dtype=i32, device=d)
            self.dev_in["query_v"] = torch.empty(G, dtype=i32, device=d)
        self.lengths = torch.empty(G, dtype=i32, device=d)
        self.offsets = torch.empty(G + 1, dtype=i64, device=d)
        self.tokens = torch.empty(tok_capacity, dtype=i64, device=d)
        self.lmax = torch.empty(self.nb, dtype=i32, device=d)
        self.cells = torch.empty(self.nb, dtype=i32, device=d)
        self.xoff = torch.empty(self.nb + 1, dtype=i64, device=d)
        self.X = torch.empty(pad_capacity, dtype=i64, device=d)
        self.attn = torch.empty(pad_capacity, dtype=torch.bool, device=d)
        self.ws_bytes = lib.gtok_scan_workspace_bytes(max(G, self.nb))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=d)
        self.h_lengths = torch.empty(G, dtype=i32).pin_memory()
        self.h_lmax = torch.empty(self.nb, dtype=i32).pin_memory()
        self.h_xoff = torch.empty(self.nb + 1, dtype=i64).pin_memory()
        self.kernel_launches_per_pass = 6
        self.timers = None      # optional list to collect (name, start_event, end_event)
        self._n_cap, self._e_cap = int(n_total), int(e_total)

    def _release_wire(self) -> None:
        """Called after the kernels that read a wire slot were enqueued: the slot may be refilled once they ran."""
        slot = getattr(self, "_wire_last", None)
        if slot is not None:
            evt = torch.cuda.Event()
            evt.record(torch.cuda.current_stream())
            self._wire_free[slot] = evt
            self._wire_last = None

    # -- helpers ---------------------------------------------------------
    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.dev_in.values())

    def d2h_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.h_lengths, self.h_lmax, self.h_xoff))

    def upload(self, host: GraphBatch) -> None:
        """pinned host arrays -> the preallocated device arrays (async on the current stream)."""
        src = {"node_off": host.node_off, "edge_off": host.edge_off, "src": host.src, "dst": host.dst,
               "x": host.x, "edge_attr": host.edge_attr, "query_u": host.query_u, "query_v": host.query_v}
        for k, dst in self.dev_in.items():
            dst.copy_(src[k], non_blocking=True)

    # -- compact wire format, double buffered -----------------------------------------------------------------
    def _wire_slot(self, host: PackedGraphBatch, slot: int):
        key = tuple((n, getattr(host, n).dtype, getattr(host, n).numel()) for n in host.ARRAYS if getattr(host, n) is not None)
        if getattr(self, "_wire", None) is None:
            self._wire, self._wire_key, self._wire_evt, self._wire_src = [None, None], [None, None], [None, None], [None, None]
            self._wire_free, self._wire_last = [None, None], None
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._cnt_tmp = torch.empty(max(1, self.G), dtype=torch.int32, device=self.device)
        if self._wire_key[slot] != key:
            self._wire[slot] = {n: torch.empty_like(getattr(host, n), device=self.device) for n in host.ARRAYS
                                if getattr(host, n) is not None}
            self._wire_key[slot] = key
        return self._wire[slot]

    def _copy_packed(self, host: PackedGraphBatch, slot: int, stream) -> None:
        """pinned host arrays -> the slot's narrow device staging buffers, async on `stream`."""
        w = self._wire_slot(host, slot)
        with torch.cuda.stream(stream):
            for n, d in w.items():
                d.copy_(getattr(host, n), non_blocking=True)
            evt = torch.cuda.Event()
            evt.record(stream)
        self._wire_evt[slot], self._wire_src[slot] = evt, host

    def upload_packed(self, host: PackedGraphBatch, prefetch: Optional[PackedGraphBatch] = None) -> None:
        """compact pinned host arrays -> narrow device staging -> gtok_widen_i32 / gtok_counts_to_offsets into the
        resident int32 layout.  `prefetch` (the NEXT call's input) starts travelling on a copy stream into the other
        staging slot right away, so its PCIe time hides behind this call's kernels; a later call with that object
        finds it already on the device."""
        main = torch.cuda.current_stream()
        for h in (host, prefetch):
            if h is None:
                continue
            if h.num_graphs != self.G:
                raise ValueError(f"this pipeline was built for shards of {self.G} graphs, got {h.num_graphs}")
            if h.num_nodes_total > self._n_cap or h.num_entries_total > self._e_cap:
                raise ValueError(f"shard of {h.num_nodes_total} nodes / {h.num_entries_total} entries exceeds the pipeline's "
                                 f"capacity ({self._n_cap} / {self._e_cap}); raise node_capacity / entry_capacity")
        self._wire_slot(host, 0)
        slot = next((s_ for s_ in (0, 1) if self._wire_src[s_] is host), None)
        if slot is None:                                  # not prefetched: copy now, on the main stream
            slot = 0 if self._wire_src[0] is None or self._wire_src[1] is not None else 1
            if self._wire_evt[slot] is not None:          # an unconsumed prefetch into this slot may still be in flight
                main.wait_event(self._wire_evt[slot])
            self._copy_packed(host, slot, main)
        if prefetch is not None:
            other = 1 - slot
            # the other slot's last readers (an earlier call's widen / collate kernels) were enqueued on the main stream
            if self._wire_free[other] is not None:
                self._copy_stream.wait_event(self._wire_free[other])
            self._copy_packed(prefetch, other, self._copy_stream)
        main.wait_event(self._wire_evt[slot])
        self._wire_last = slot
        self._h2d_packed = host.nbytes()
        w = self._wire[slot]
        st = _stream()
        edge_off = self.dev_in["edge_off"]
        if host.unique_columns:
            # offsets of the de-mirrored column list live in their own buffer: dev_in keeps describing the int32 layout
            if getattr(self, "_uniq_edge_off", None) is None:
                self._uniq_edge_off = torch.empty(self.G + 1, dtype=torch.int64, device=self.device)
            edge_off = self._uniq_edge_off
        for cnt, off in (("node_cnt", self.dev_in["node_off"]), ("entry_cnt", edge_off)):
            check(lib.gtok_counts_to_offsets(_ptr(w[cnt]), w[cnt].element_size(), self.G, _ptr(off),
                                             _ptr(self._cnt_tmp), _ptr(self.ws), self.ws_bytes, st))
        self._packed_dev = None
        if "col16" in w and getattr(self, "mode", None) == "collate" and w["x"].dtype == torch.uint8:
            # fused 16-bit columns: the collate kernel reads the wire buffers directly (gtok_zinc_collate_packed)
            self._packed_dev = (w["col16"], w["x"], host.max_id, bool(host.unique_columns), edge_off)
        else:
            if "col16" in w:
                raise ValueError("fused 16-bit columns are only consumed by the ZINC pipeline in collate mode")
            for n in ("src", "dst", "x", "edge_attr", "query_u", "query_v"):
                if n in self.dev_in and n in w:
                    if w[n].numel() > self.dev_in[n].numel():
                        raise ValueError(f"{n}: {w[n].numel()} elements exceed the pipeline's capacity {self.dev_in[n].numel()}")
                    check(lib.gtok_widen_i32(_ptr(w[n]), w[n].element_size(), w[n].numel(), _ptr(self.dev_in[n]), st))
        self._wire_src[slot] = None                       # consumed (a slot is reused only after this call returned)

    def use_device_batch(self, dev_batch: GraphBatch) -> None:
        """Point the pipeline at arrays already resident in HBM (no copy)."""
        m = {"node_off": dev_batch.node_off, "edge_off": dev_batch.edge_off, "src": dev_batch.src,
             "dst": dev_batch.dst, "x": dev_batch.x, "edge_attr": dev_batch.edge_attr,
             "query_u": dev_batch.query_u, "query_v": dev_batch.query_v}
        for k in list(self.dev_in):
            self.dev_in[k] = m[k]

    def download(self) -> None:
        self.h_lengths.copy_(self.lengths, non_blocking=True)
        self.h_lmax.copy_(self.lmax, non_blocking=True)
        self.h_xoff.copy_(self.xoff, non_blocking=True)

    def _timed(self, name, fn):
        if self.timers is None:
            fn()
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        self.timers.append((name, a, b))

    def _scan(self, lens, n, off):
        check(lib.gtok_scan_offsets(_ptr(lens), n, _ptr(off), _ptr(self.ws), self.ws_bytes, _stream()))

    def _collate(self, pad_id: int):
        st = _stream()
        self._timed("batch_extents", lambda: check(lib.gtok_batch_extents(
            _ptr(self.lengths), None, self.G, self.B, _ptr(self.lmax), _ptr(self.cells), st)))
        self._timed("scan_batches", lambda: self._scan(self.cells, self.nb, self.xoff))
        self._timed("pad_mask_batched", lambda: check(lib.gtok_pad_mask_batched(
            _ptr(self.tokens), _ptr(self.offsets), _ptr(self.lengths), None, self.G, self.B, _ptr(self.lmax),
            _ptr(self.xoff), pad_id, _ptr(self.X), _ptr(self.attn), st)))

    def _result(self, e2e: bool) -> EpochResult:
        pe = engine.PaddedEpoch(self.X, self.attn, self.xoff, self.lmax, self.B, self.G)
        if e2e:
            pe._host = None
            return EpochResult(self.tokens, self.offsets, self.lengths, pe, self.h_lengths, self.h_lmax, self.h_xoff)
        return EpochResult(self.tokens, self.offsets, self.lengths, pe)


class ZincIbttPipeline(_PipelineBase):
    """BASELINE config 2: IBTT tokenization of ZINC-shaped molecules + collate."""

    def __init__(self, host_batch: GraphBatch, vocab: Dict[str, int], max_len: int, batch_size: int, device,
                 fused: Optional[bool] = None, mode: str = "auto", max_nodes: Optional[int] = None,
                 max_entries: Optional[int] = None, node_capacity: Optional[int] = None,
                 entry_capacity: Optional[int] = None):
        """`host_batch` sizes the buffers (graph count is fixed; node / entry totals and the per-molecule bounds can be
        raised with max_nodes / max_entries / node_capacity / entry_capacity when a stream of different shards of
        G graphs goes through the same pipeline).
           mode: "collate"  one kernel, molecules -> padded batches (gtok_zinc_collate; the ragged stream is
                            never materialised), the default when the shapes allow it;
                 "ragged"   gtok_zinc_tokenize (ragged ids, kept for TokenDataset.seqs) + the K4 collate kernels;
                 "auto"     "collate" if batch_size is a multiple of 32 <= 256 and molecules are <= 64 atoms /
                            <= 256 columns, else "ragged".  `fused=False` forces the three-kernel tokenizer."""
        G = host_batch.num_graphs
        n = host_batch.num_nodes()
        e = host_batch.num_entries()
        self.max_nodes = max(int(n.max()) if G else 1, max_nodes or 0)
        self.max_entries = max(int(e.max()) if G else 1, max_entries or 0)
        # per-warp staging bounds (aligned groups of 32 molecules) - only when this batch is all the pipeline will see
        stream_of_shards = any(v is not None for v in (max_nodes, max_entries, node_capacity, entry_capacity))
        self.group_nodes = 0 if stream_of_shards else engine.group_bounds(host_batch.node_off)
        self.group_entries = 0 if stream_of_shards else engine.group_bounds(host_batch.edge_off)
        n_total = max(int(host_batch.node_off[-1]), node_capacity or 0)
        e_total = max(int(host_batch.edge_off[-1]), entry_capacity or 0)
        row_bound = min(max_len, 2 * self.max_nodes + 4 * self.max_entries + 4)
        tok_cap = min(G * row_bound, 2 * n_total + 4 * e_total + 4 * G)
        super().__init__(G, n_total, e_total, tok_cap + 2, G * row_bound + 2, batch_size, device, ["x", "edge_attr"])
        self.vocab = vocab
        self.pad_id = int(vocab["<pad>"])
        self.zv, self._tbl = engine.zinc_vocab_c(vocab, max_len, max_len, self.max_nodes, self.device)
        small = self.max_nodes <= 64 and self.max_entries <= 256
        self.fused = small if fused is None else bool(fused)
        can_collate = small and self.fused and batch_size % 32 == 0 and 32 <= batch_size <= 256
        if mode == "auto":
            mode = "collate" if can_collate else "ragged"
        if mode == "collate" and not can_collate:
            raise ValueError("fused tokenize+collate needs molecule shapes and a batch size that is a multiple of 32 <= 256")
        self.mode = mode
        if self.fused:
            self.kernel_launches_per_pass = 1 if mode == "collate" else 4
            self.ws_bytes = max(self.ws_bytes, lib.gtok_zinc_tokenize_workspace_bytes(G),
                                lib.gtok_zinc_collate_workspace_bytes(G, batch_size))
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)

    def _batch_c(self):
        d = self.dev_in
        return _lib.GraphBatchC(self.G, _ptr(d["node_off"]), _ptr(d["edge_off"]), _ptr(d["src"]), _ptr(d["dst"]),
                                _ptr(d["x"]), _ptr(d["edge_attr"]), None, None, max(1, self.max_nodes),
                                max(1, self.max_entries), self.group_nodes, self.group_entries)

    def launch(self) -> None:
        """The hot path on whatever dev_in points at: 4 kernels (fused tokenizer) or 6, no sync."""
        bc = self._batch_c()
        st = _stream()
        if self.mode == "collate" and getattr(self, "_packed_dev", None) is not None:
            col16, x8, max_id, unique, edge_off = self._packed_dev
            d = self.dev_in
            pk = _lib.PackedMolBatchC(self.G, _ptr(d["node_off"]), _ptr(edge_off), _ptr(col16), _ptr(x8),
                                      max(1, self.max_nodes, max_id + 1), max(1, self.max_entries), int(unique),
                                      self.group_nodes, 0 if unique else self.group_entries, 0)
            self._timed("zinc_collate_packed", lambda: check(lib.gtok_zinc_collate_packed(
                C.byref(pk), C.byref(self.zv), self.B, self.pad_id, _ptr(self.lengths), _ptr(self.lmax), _ptr(self.xoff),
                _ptr(self.X), _ptr(self.attn), min(self.X.numel(), self.attn.numel()), _ptr(self.ws), self.ws_bytes, st)))
            return
        if self.mode == "collate":
            self._timed("zinc_collate", lambda: check(lib.gtok_zinc_collate(
                C.byref(bc), C.byref(self.zv), self.B, self.pad_id, _ptr(self.lengths), _ptr(self.lmax), _ptr(self.xoff),
                _ptr(self.X), _ptr(self.attn), min(self.X.numel(), self.attn.numel()), _ptr(self.ws), self.ws_bytes, st)))
            return
        if self.fused:
            self._timed("zinc_tokenize", lambda: check(lib.gtok_zinc_tokenize(
                C.byref(bc), C.byref(self.zv), _ptr(self.lengths), _ptr(self.offsets), _ptr(self.tokens),
                self.tokens.numel(), _ptr(self.ws), self.ws_bytes, st)))
            self._collate(self.pad_id)
            return
        self._timed("zinc_lengths", lambda: check(lib.gtok_zinc_lengths(C.byref(bc), C.byref(self.zv), _ptr(self.lengths), st)))
        self._timed("scan_rows", lambda: self._scan(self.lengths, self.G, self.offsets))
        self._timed("zinc_serialize", lambda: check(lib.gtok_zinc_serialize(
            C.byref(bc), C.byref(self.zv), _ptr(self.offsets), _ptr(self.tokens), st)))
        self._collate(self.pad_id)

    def run(self, host_batch, prefetch=None) -> EpochResult:
        """Public end-to-end call: pinned host graphs in (GraphBatch, or the compact PackedGraphBatch wire
        format), device batches + host row metadata out.  `prefetch`: the PackedGraphBatch of the NEXT call
        (its upload overlaps this call's kernels)."""
        if isinstance(host_batch, PackedGraphBatch):
            self.upload_packed(host_batch, prefetch)
        else:
            self._packed_dev = None
            self.upload(host_batch)
        self.launch()
        self._release_wire()
        self._packed_dev = None                  # launch() on resident int32 arrays again unless the next run() is packed
        self.download()
        torch.cuda.current_stream().synchronize()
        return self._result(True)

    # -- end to end from the reference's own storage ------------------------------------------------------------
    def _pyg_buffers(self):
        if getattr(self, "_pyg", None) is None:
            d, i64 = self.device, torch.int64
            e, n = self._e_cap + 16, self._n_cap + 16
            self._pyg = {"src64": torch.empty(e, dtype=i64, device=d), "dst64": torch.empty(e, dtype=i64, device=d),
                         "attr64": torch.empty(e, dtype=i64, device=d), "x64": torch.empty(n, dtype=i64, device=d),
                         "col16": torch.empty(e, dtype=torch.int16, device=d), "x8": torch.empty(n, dtype=torch.uint8, device=d),
                         "node_off": torch.empty(self.G + 1, dtype=i64, device=d),
                         "edge_off": torch.empty(self.G + 1, dtype=i64, device=d),
                         "status": torch.zeros(1, dtype=torch.int32, device=d)}
            self._pyg_hstatus = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._pyg_free = None
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
        return self._pyg

    def run_pyg(self, store: PygStore, chunks: int = 8) -> EpochResult:
        """Public end-to-end call from the REFERENCE's storage: pinned int64 PyG tensors (`PygStore`: edge_index [2, E],
        x, edge_attr, slices - zinc_dataset_autograph.py:51-73) -> device batches + host row metadata, every conversion
        inside the call and on the device.  The int64 arrays cross PCIe as they are, in `chunks` pieces on a copy
        stream; each piece is narrowed as soon as it has landed (gtok_pyg_pack_mol: 16-bit columns + byte atoms, or
        gtok_pyg_narrow into the int32 layout when a value does not fit / the pipeline is not in collate mode); the
        slices are the offsets.  Stands in for `[ds[i] for i in range(len(ds))] -> TokenDataset -> DataLoader(collate)`
        (train_ibtt.py:231-233, :395-402) on the dataset object the reference already holds."""
        if store.num_graphs != self.G:
            raise ValueError(f"this pipeline was built for shards of {self.G} graphs, got {store.num_graphs}")
        E, N = int(store.edge_index.shape[1]), int(store.x.numel())
        if E > self._e_cap or N > self._n_cap:
            raise ValueError(f"store of {N} nodes / {E} entries exceeds the pipeline's capacity ({self._n_cap} / {self._e_cap})")
        w = self._pyg_buffers()
        main, cs = torch.cuda.current_stream(), self._copy_stream
        packed = self.mode == "collate"
        has_attr = store.edge_attr is not None
        w["status"].zero_()
        if self._pyg_free is not None:
            cs.wait_event(self._pyg_free)                 # the previous call's kernels have drained the staging arrays
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            w["node_off"].copy_(store.node_slices, non_blocking=True)
            w["edge_off"].copy_(store.edge_slices, non_blocking=True)
        if not packed:
            i32 = getattr(self, "_pyg_i32", None)
            if i32 is None:
                i32 = self._pyg_i32 = {k: torch.empty(self._e_cap + 16 if k != "x" else self._n_cap + 16, dtype=torch.int32,
                                                       device=self.device) for k in ("src", "dst", "edge_attr", "x")}
        st = _stream()
        nch = max(1, int(chunks))
        cut = lambda total, c: min(total, ((total * c // nch) + 15) & ~15)
        x_flat = store.x.reshape(-1)
        for c in range(nch):
            e0, e1, n0, n1 = cut(E, c), cut(E, c + 1), cut(N, c), cut(N, c + 1)
            if c == nch - 1:
                e1, n1 = E, N
            with torch.cuda.stream(cs):
                if e1 > e0:
                    w["src64"][e0:e1].copy_(store.edge_index[0, e0:e1], non_blocking=True)
                    w["dst64"][e0:e1].copy_(store.edge_index[1, e0:e1], non_blocking=True)
                    if has_attr:
                        w["attr64"][e0:e1].copy_(store.edge_attr[e0:e1], non_blocking=True)
                if n1 > n0:
                    w["x64"][n0:n1].copy_(x_flat[n0:n1], non_blocking=True)
                evt = torch.cuda.Event()
                evt.record(cs)
            main.wait_event(evt)
            at = lambda t, o: t.data_ptr() + o * t.element_size()
            if packed:
                check(lib.gtok_pyg_pack_mol(at(w["src64"], e0), at(w["dst64"], e0), at(w["attr64"], e0) if has_attr else None,
                                            e1 - e0, at(w["x64"], n0), n1 - n0, at(w["col16"], e0), at(w["x8"], n0),
                                            _ptr(w["status"]), st))
            else:
                for k64, k32, a, b in (("src64", "src", e0, e1), ("dst64", "dst", e0, e1), ("x64", "x", n0, n1)) + \
                        ((("attr64", "edge_attr", e0, e1),) if has_attr else ()):
                    check(lib.gtok_pyg_narrow(at(w[k64], a), b - a, at(i32[k32], a), _ptr(w["status"]), st))
        self._h2d_pyg = store.nbytes()
        if packed:
            self._packed_dev = (w["col16"], w["x8"], max(0, self.max_nodes - 1), False, w["edge_off"])
            keep = self.dev_in["node_off"]
            self.dev_in["node_off"] = w["node_off"]
            self.launch()
            self.dev_in["node_off"] = keep
            self._packed_dev = None
        else:
            keep = dict(self.dev_in)
            self.dev_in.update({"node_off": w["node_off"], "edge_off": w["edge_off"], "src": i32["src"][:E], "dst": i32["dst"][:E],
                                "x": i32["x"][:N], "edge_attr": i32["edge_attr"][:E] if has_attr else None})
            self.launch()
            self.dev_in = keep
        self._pyg_free = torch.cuda.Event()
        self._pyg_free.record(main)
        self.download()
        self._pyg_hstatus.copy_(w["status"], non_blocking=True)
        main.synchronize()
        flags = int(self._pyg_hstatus[0])
        if flags:
            what = [n for b, n in ((_lib.PYG_BAD_NODE_ID, "a local node id outside 0..63"),
                                   (_lib.PYG_BAD_EDGE_ATTR, "an edge_attr outside 0..15"),
                                   (_lib.PYG_BAD_NODE_TYPE, "an atom type outside 0..255"),
                                   (_lib.PYG_BAD_RANGE, "a value outside int32")) if flags & b]
            raise ValueError("run_pyg: the store holds " + ", ".join(what) + "; build the pipeline with mode='ragged' "
                             "(int32 layout) for graphs that are not molecule sized")
        return self._result(True)

    def algorithmic_bytes(self, total_tokens: int, padded_cells: int) -> Dict[str, int]:
        """SURVEY.md 8d per-graph figures x the graphs of one launch."""
        G = self.G
        n_total = self.dev_in["x"].numel()
        e_total = self.dev_in["src"].numel()
        return {
            "zinc_collate": 12 * e_total + 4 * n_total + 16 * G + 9 * padded_cells + 4 * G,
            "zinc_tokenize": 12 * e_total + 4 * n_total + 16 * G + 8 * total_tokens,
            "zinc_lengths": 8 * e_total + 16 * G + 4 * G,
            "scan_rows": 4 * G + 8 * (G + 1),
            "zinc_serialize": 12 * e_total + 4 * n_total + 16 * G + 8 * total_tokens,
            "batch_extents": 4 * G + 8 * self.nb,
            "scan_batches": 4 * self.nb + 8 * (self.nb + 1),
            "pad_mask_batched": 8 * total_tokens + 9 * padded_cells,
        }


class GraphTokenIbttPipeline(_PipelineBase):
    """BASELINE config 1/5: IBTT tokenization of graph-token graphs + collate."""

    def __init__(self, host_batch: GraphBatch, vocab: Dict[str, int], task: str, max_len: int, batch_size: int, device,
                 mode: str = "collate"):
        """mode "collate": lengths -> batch extents -> scan -> gtok_ibtt_collate (graphs -> padded batches, the ragged
        stream never exists); "ragged": lengths -> scan -> serialize (ragged ids kept) -> K4 collate kernels."""
        G = host_batch.num_graphs
        self.mode = mode
        n = host_batch.num_nodes()
        e = host_batch.num_entries()
        self.max_nodes = int(n.max()) if G else 1
        q = 5 if task == "cycle_check" else 7
        n_total, e_total = int(host_batch.node_off[-1]), int(host_batch.edge_off[-1])
        row_bound = min(max_len, int((3 * e + n).max()) + q) if G else 1
        tok_cap = min(G * row_bound, 3 * e_total + n_total + q * G)
        fields = ["query"] if task == "shortest_path" else []
        super().__init__(G, n_total, e_total, tok_cap + 2, G * row_bound + 2, batch_size, device, fields)
        self.task = task
        self.kernel_launches_per_pass = 4 if mode == "collate" else 6
        self.pad_id = int(vocab["<pad>"])
        self.rules, self._tbl = engine.ibtt_rules(vocab, task, max_len, self.max_nodes, self.device)

    def _batch_c(self):
        d = self.dev_in
        return _lib.GraphBatchC(self.G, _ptr(d["node_off"]), _ptr(d["edge_off"]), _ptr(d["src"]), _ptr(d["dst"]),
                                None, None, _ptr(d.get("query_u")), _ptr(d.get("query_v")), max(1, self.max_nodes), 1)

    def launch(self) -> None:
        bc = self._batch_c()
        st = _stream()
        self._timed("ibtt_lengths", lambda: check(lib.gtok_ibtt_lengths(C.byref(bc), C.byref(self.rules), _ptr(self.lengths), st)))
        if self.mode == "collate":
            self._timed("batch_extents", lambda: check(lib.gtok_batch_extents(
                _ptr(self.lengths), None, self.G, self.B, _ptr(self.lmax), _ptr(self.cells), st)))
            self._timed("scan_batches", lambda: self._scan(self.cells, self.nb, self.xoff))
            self._timed("ibtt_collate", lambda: check(lib.gtok_ibtt_collate(
                C.byref(bc), C.byref(self.rules), None, self.G, self.B, _ptr(self.lmax), _ptr(self.xoff), self.pad_id,
                _ptr(self.X), _ptr(self.attn), st)))
            return
        self._timed("scan_rows", lambda: self._scan(self.lengths, self.G, self.offsets))
        self._timed("ibtt_serialize", lambda: check(lib.gtok_ibtt_serialize(
            C.byref(bc), C.byref(self.rules), _ptr(self.offsets), _ptr(self.tokens), st)))
        self._collate(self.pad_id)

    def run(self, host_batch, prefetch=None) -> EpochResult:
        if isinstance(host_batch, PackedGraphBatch):
            self.upload_packed(host_batch, prefetch)
        else:
            self.upload(host_batch)
        self.launch()
        self._release_wire()
        self.download()
        torch.cuda.current_stream().synchronize()
        return self._result(True)

    def algorithmic_bytes(self, total_tokens: int, padded_cells: int) -> Dict[str, int]:
        G = self.G
        e_total = self.dev_in["src"].numel()
        qb = 8 * G if self.task == "shortest_path" else 0
        return {
            "ibtt_collate": 8 * e_total + 32 * G + qb + 9 * padded_cells,
            "ibtt_lengths": 32 * G + 4 * G,
            "scan_rows": 4 * G + 8 * (G + 1),
            "ibtt_serialize": 8 * e_total + 16 * G + qb + 8 * total_tokens,
            "batch_extents": 4 * G + 8 * self.nb,
            "scan_batches": 4 * self.nb + 8 * (self.nb + 1),
            "pad_mask_batched": 8 * total_tokens + 9 * padded_cells,
        }
