"""Host side of the engine: torch CUDA tensors in, torch CUDA tensors out, every
byte of arithmetic done by libgtok.so kernels (no torch op computes a token).

PyTorch is plumbing here: device memory, the current stream, and (in bench.py)
torch.distributed.  Each function names the reference code it stands in for.
"""
from __future__ import annotations

import math

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from ._lib import lib, check
from .batch import GraphBatch

ZINC_ATOMS = ["C", "N", "O", "F", "P", "S", "Cl", "Br", "I"]       # zinc_vocab.py:21
ZINC_BONDS = ["single", "double", "triple", "aromatic"]           # zinc_vocab.py:25


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("gtok kernels need CUDA tensors (there is no CPU fallback)")


@dataclass
class Ragged:
    """Ragged int64 token rows: row g = tokens[offsets[g]:offsets[g+1]]."""
    tokens: torch.Tensor      # int64 [total]
    offsets: torch.Tensor     # int64 [G+1]
    lengths: torch.Tensor     # int32 [G]

    @property
    def num_rows(self) -> int:
        return int(self.lengths.numel())

    def row(self, g: int) -> torch.Tensor:
        o = self.offsets[g:g + 2].tolist()
        return self.tokens[o[0]:o[1]]

    def rows_cpu(self):
        t, o = self.tokens.cpu(), self.offsets.cpu().tolist()
        return [t[o[i]:o[i + 1]] for i in range(len(o) - 1)]


def _batch_c(b: GraphBatch, max_nodes: Optional[int], max_entries: Optional[int]) -> _lib.GraphBatchC:
    G = b.num_graphs
    if max_nodes is None:
        max_nodes = int(b.num_nodes().max()) if G else 0
    if max_entries is None:
        max_entries = int(b.num_entries().max()) if G else 0
    return _lib.GraphBatchC(G, _ptr(b.node_off), _ptr(b.edge_off), _ptr(b.src), _ptr(b.dst), _ptr(b.x),
                            _ptr(b.edge_attr), _ptr(b.query_u), _ptr(b.query_v), max(1, max_nodes), max(1, max_entries))


def group_bounds(off: torch.Tensor, group: int = 32) -> int:
    """max over aligned groups of `group` consecutive graphs of their item total (off = node_off / edge_off):
    the staging bound GtokGraphBatch.max_nodes_per_32 / max_entries_per_32 of the thread-per-molecule kernels."""
    G = off.numel() - 1
    if G <= 0:
        return 0
    idx = torch.arange(0, G + group, group, device=off.device).clamp_(max=G)
    b = off[idx]
    return int((b[1:] - b[:-1]).max())


def node_token_table(vocab: Dict[str, int], size: int, device) -> torch.Tensor:
    """node_tok[k] = vocab.get(str(k), vocab['<pad>']) - the lookup the reference does
    per token with a dict (data_loader.py:482)."""
    pad = vocab["<pad>"]
    return torch.tensor([vocab.get(str(k), pad) for k in range(size)], dtype=torch.int32, device=device)


def scan_offsets(lengths: torch.Tensor) -> torch.Tensor:
    """K0: int32[G] -> int64[G+1] exclusive offsets."""
    _require_cuda(lengths)
    G = lengths.numel()
    off = torch.empty(G + 1, dtype=torch.int64, device=lengths.device)
    nbytes = lib.gtok_scan_workspace_bytes(G)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=lengths.device)
    check(lib.gtok_scan_offsets(_ptr(lengths), G, _ptr(off), _ptr(ws), nbytes, _stream()))
    return off


# ---------------------------------------------------------------------------
# K1: graph-token IBTT  (external serializer + data_loader.py:465-486)
# ---------------------------------------------------------------------------
def ibtt_rules(vocab: Dict[str, int], task: str, max_len: int, table_size: int, device):
    pad = vocab["<pad>"]
    g = lambda t: vocab.get(t, pad)
    tbl = node_token_table(vocab, table_size, device)
    if task == "cycle_check":
        tcode, ttok = _lib.TASK_CYCLE_CHECK, g("has_cycle")
    elif task == "shortest_path":
        tcode, ttok = _lib.TASK_SHORTEST_PATH, g("shortest_distance")
    else:
        raise ValueError(f"unsupported graph-token task {task!r}")
    r = _lib.IbttRulesC(tcode, int(max_len), g("<bos>"), g("<e>"), g("<n>"), g("<q>"), g("<p>"), ttok, pad,
                        tbl.numel(), _ptr(tbl), max(vocab.values()), 0)
    return r, tbl      # keep tbl alive as long as r is used


def tokenize_ibtt(batch: GraphBatch, vocab: Dict[str, int], task: str, max_len: int,
                  max_nodes: Optional[int] = None, out: Optional[torch.Tensor] = None) -> Ragged:
    _require_cuda(batch.src)
    dev = batch.device
    G = batch.num_graphs
    if max_nodes is None:
        max_nodes = int(batch.num_nodes().max()) if G else 0
    bc = _batch_c(batch, max_nodes, 1)
    rules, tbl = ibtt_rules(vocab, task, max_len, max(1, max_nodes), dev)
    lengths = torch.empty(G, dtype=torch.int32, device=dev)
    check(lib.gtok_ibtt_lengths(C.byref(bc), C.byref(rules), _ptr(lengths), _stream()))
    off = scan_offsets(lengths)
    total = int(off[-1]) if G else 0
    if out is None:
        out = torch.empty(total, dtype=torch.int64, device=dev)
    elif out.numel() < total:
        raise ValueError("output buffer too small")
    check(lib.gtok_ibtt_serialize(C.byref(bc), C.byref(rules), _ptr(off), _ptr(out), _stream()))
    return Ragged(out[:total], off, lengths)


SPECIAL = ["<pad>", "<bos>", "<e>", "<n>", "<q>", "<p>", "<eos>", "yes", "no"]      # data_loader.py:10


def graphs_from_texts(texts: Sequence[str], device="cuda"):
    """Ingest: graph-token texts -> (GraphBatch on `device`, task).  The inverse of the external serializer, run
    per text on the host by the reference (graph_token_dataset_autograph.py:14-54 parse_graph_from_text +
    data_loader.py:60-74 parse_query_nodes_from_text); here two kernels over the concatenated bytes
    (gtok_text_scan / gtok_text_edges).  Raises ValueError on a text that is not of the grammar or on mixed
    tasks - the CUDA tokenizers cannot encode such a text and there is no CPU fallback."""
    dev = torch.device(device)
    G = len(texts)
    if G == 0:
        z = torch.zeros(1, dtype=torch.int64, device=dev)
        e = torch.zeros(0, dtype=torch.int32, device=dev)
        return GraphBatch(z, z.clone(), e, e.clone()), "cycle_check"
    try:
        blob = "".join(texts).encode("ascii")
    except UnicodeEncodeError as ex:
        raise ValueError(f"text is not graph-token grammar (non-ASCII character): {ex}") from None
    lens = torch.tensor([len(t) for t in texts], dtype=torch.int64)
    off = torch.zeros(G + 1, dtype=torch.int64)
    off[1:] = lens.cumsum(0)
    assert int(off[-1]) == len(blob)
    text_d = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    off_d = off.to(dev)
    nn = torch.empty(G, dtype=torch.int32, device=dev)
    ne = torch.empty(G, dtype=torch.int32, device=dev)
    task = torch.empty(G, dtype=torch.int32, device=dev)
    qu = torch.empty(G, dtype=torch.int32, device=dev)
    qv = torch.empty(G, dtype=torch.int32, device=dev)
    check(lib.gtok_text_scan(_ptr(text_d), _ptr(off_d), G, _ptr(nn), _ptr(ne), _ptr(task), _ptr(qu), _ptr(qv), _stream()))
    node_off, edge_off = scan_offsets(nn), scan_offsets(ne)
    E = int(edge_off[-1])
    src = torch.empty(E, dtype=torch.int32, device=dev)
    dst = torch.empty(E, dtype=torch.int32, device=dev)
    status = torch.empty(G, dtype=torch.int32, device=dev)
    check(lib.gtok_text_edges(_ptr(text_d), _ptr(off_d), G, _ptr(nn), _ptr(task), _ptr(edge_off), _ptr(src), _ptr(dst),
                              _ptr(status), _stream()))
    bad = (status == 0).nonzero()
    if bad.numel():
        i = int(bad[0])
        raise ValueError("text is not graph-token grammar and carries no graph handle; the CUDA tokenizers cannot "
                         f"encode it (no CPU fallback): {texts[i][:80]!r}")
    t0 = int(task[0])
    if bool((task != t0).any()):
        raise ValueError("mixed tasks in one TokenDataset: ['cycle_check', 'shortest_path']")
    name = "cycle_check" if t0 == _lib.TASK_CYCLE_CHECK else "shortest_path"
    sp = name == "shortest_path"
    return GraphBatch(node_off, edge_off, src, dst, None, None, qu if sp else None, qv if sp else None), name


def vocab_from_graphs(batch: GraphBatch, task: str, label_tokens: Sequence[str], min_freq: int = 1,
                      max_tokens: Optional[int] = None):
    """build_vocab_from_texts (data_loader.py:451-463) without the texts: identical (vocab, itos) to running the
    reference on the graph-token serialization of `batch` with label token label_tokens[g] ("yes"/"no",
    "len3", "INF", ...).  The O(total tokens) Counter pass runs on the GPU (K5 gtok_vocab_histogram); the host
    sorts <= max_nodes + |labels| + 1 entries by (-count, first occurrence) like Counter.most_common()."""
    _require_cuda(batch.src)
    G = batch.num_graphs
    if len(label_tokens) != G:
        raise ValueError("one label token per graph expected")
    if task == "cycle_check":
        tcode, tword, qtok = _lib.TASK_CYCLE_CHECK, "has_cycle", 3
    elif task == "shortest_path":
        tcode, tword, qtok = _lib.TASK_SHORTEST_PATH, "shortest_distance", 5
    else:
        raise ValueError(f"unsupported graph-token task {task!r}")
    dev = batch.device
    bins = int(batch.num_nodes().max()) if G else 0
    if task == "shortest_path" and G:
        bins = max(bins, int(batch.query_u.max()) + 1, int(batch.query_v.max()) + 1)
    if G:
        bins = max(bins, int(batch.src.max()) + 1 if batch.src.numel() else 0, int(batch.dst.max()) + 1 if batch.dst.numel() else 0)
    counts = torch.empty(bins, dtype=torch.int64, device=dev)
    first = torch.empty(bins, dtype=torch.int64, device=dev)
    bc = _batch_c(batch, max(1, bins), 1)
    check(lib.gtok_vocab_histogram(C.byref(bc), tcode, bins, _ptr(counts), _ptr(first), _stream()))
    counts_h, first_h = counts.cpu().tolist(), first.cpu().tolist()
    n_h, e_h = batch.num_nodes().cpu(), batch.num_entries().cpu()
    items = []                                                    # (token, count, first-occurrence key)
    for k in range(bins):
        if counts_h[k] > 0:
            items.append((str(k), counts_h[k], first_h[k] & 0xFFFFFFFFFFFFFFFF))
    if G:
        pos_q = (3 * e_h + n_h + 2)                               # position of <q> in text g
        items.append((tword, G, (0 << 32) | int(pos_q[0] + 1)))
        seen = {}
        for g, lt in enumerate(label_tokens):                     # O(G) strings, not O(tokens)
            c = seen.get(lt)
            if c is None:
                seen[lt] = [1, (g << 32) | int(pos_q[g] + qtok)]
            else:
                c[0] += 1
        items += [(t, c[0], c[1]) for t, c in seen.items()]
    items.sort(key=lambda it: (-it[1], it[2]))
    vocab = {tok: i for i, tok in enumerate(SPECIAL)}
    nxt = len(vocab)
    for tok, c, _ in items:
        if tok in vocab:
            continue
        if c < min_freq:
            break
        vocab[tok] = nxt
        nxt += 1
        if max_tokens and nxt >= max_tokens:
            break
    return vocab, {i: t for t, i in vocab.items()}


def ibtt_collate_epoch(batch: GraphBatch, vocab: Dict[str, int], task: str, max_len: int, batch_size: int,
                       batch_idx: Optional[torch.Tensor] = None, max_nodes: Optional[int] = None,
                       X: Optional[torch.Tensor] = None, attn: Optional[torch.Tensor] = None):
    """K1+K4 in one pass: graph-token graphs -> every padded batch of an epoch (serializer + TokenDataset +
    collate); batch_idx = the epoch's row order (DataLoader shuffle), None = dataset order.
    -> (PaddedEpoch, lengths int32[G] in dataset order)."""
    _require_cuda(batch.src, batch_idx)
    dev = batch.device
    G = batch.num_graphs
    rows = int(batch_idx.numel()) if batch_idx is not None else G
    B = int(batch_size)
    nb = (rows + B - 1) // B
    if max_nodes is None:
        max_nodes = int(batch.num_nodes().max()) if G else 0
    bc = _batch_c(batch, max_nodes, 1)
    rules, tbl = ibtt_rules(vocab, task, max_len, max(1, max_nodes), dev)
    lengths = torch.empty(G, dtype=torch.int32, device=dev)
    check(lib.gtok_ibtt_lengths(C.byref(bc), C.byref(rules), _ptr(lengths), _stream()))
    lmax = torch.empty(nb, dtype=torch.int32, device=dev)
    cells = torch.empty(nb, dtype=torch.int32, device=dev)
    check(lib.gtok_batch_extents(_ptr(lengths), _ptr(batch_idx), rows, B, _ptr(lmax), _ptr(cells), _stream()))
    xoff = scan_offsets(cells)
    if X is None or attn is None:
        total = int(xoff[-1]) if nb else 0
        X = torch.empty(total, dtype=torch.int64, device=dev)
        attn = torch.empty(total, dtype=torch.bool, device=dev)
    check(lib.gtok_ibtt_collate(C.byref(bc), C.byref(rules), _ptr(batch_idx), rows, B, _ptr(lmax), _ptr(xoff),
                                int(vocab["<pad>"]), _ptr(X), _ptr(attn), _stream()))
    return PaddedEpoch(X, attn, xoff, lmax, B, rows), lengths


# ---------------------------------------------------------------------------
# K2: ZINC IBTT  (zinc_dataset_indexbase.py:143-227 + data_loader.py:465-486)
# ---------------------------------------------------------------------------
def zinc_vocab_c(vocab: Dict[str, int], item_max_len: int, ds_max_len: int, table_size: int, device):
    pad = vocab["<pad>"]
    g = lambda t: vocab.get(t, pad)
    tbl = node_token_table(vocab, table_size, device)
    v = _lib.ZincVocabC()
    v.item_max_len, v.ds_max_len = int(item_max_len), int(ds_max_len)
    v.tok_bos, v.tok_eos, v.tok_atom, v.tok_bond = g("<bos>"), g("<eos>"), g("<atom>"), g("<bond>")
    v.tok_q, v.tok_regression, v.tok_p = g("<q>"), g("regression"), g("<p>")
    for i, a in enumerate(ZINC_ATOMS):
        v.atom_tok[i] = g(a)
    for i, b in enumerate(ZINC_BONDS):
        v.bond_tok[i] = g(b)
    v.atom_oov, v.bond_oov, v.oov = g("X"), g("unknown"), pad
    v.node_tok_len, v.node_tok = tbl.numel(), _ptr(tbl)
    v.max_token_id = max(vocab.values())
    return v, tbl


def tokenize_zinc(batch: GraphBatch, vocab: Dict[str, int], max_len: int, item_max_len: Optional[int] = None,
                  max_nodes: Optional[int] = None, max_entries: Optional[int] = None,
                  out: Optional[torch.Tensor] = None, fused: Optional[bool] = None) -> Ragged:
    """ZINC molecules -> ragged ids.  Molecule-shaped batches (<= 64 atoms, <= 256 columns) take the
    single-pass fused kernel; anything else takes lengths -> scan -> serialize."""
    _require_cuda(batch.src)
    if batch.x is None:
        raise ValueError("ZINC tokenization needs batch.x (atom types)")
    dev = batch.device
    G = batch.num_graphs
    bc = _batch_c(batch, max_nodes, max_entries)
    zv, tbl = zinc_vocab_c(vocab, max_len if item_max_len is None else item_max_len, max_len, bc.max_nodes, dev)
    lengths = torch.empty(G, dtype=torch.int32, device=dev)
    can_fuse = bc.max_nodes <= 64 and bc.max_entries <= 256
    if fused is None:
        fused = can_fuse
    if fused:
        if not can_fuse:
            raise ValueError("fused ZINC path needs <= 64 atoms and <= 256 edge_index columns per molecule")
        n_total, e_total = batch.x.numel(), batch.src.numel()
        bound = min(G * min(max_len, 2 * bc.max_nodes + 4 * bc.max_entries + 4), 2 * n_total + 4 * e_total + 4 * G)
        if out is None:
            out = torch.empty(bound + 2, dtype=torch.int64, device=dev)
        off = torch.empty(G + 1, dtype=torch.int64, device=dev)
        nbytes = lib.gtok_zinc_tokenize_workspace_bytes(G)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib.gtok_zinc_tokenize(C.byref(bc), C.byref(zv), _ptr(lengths), _ptr(off), _ptr(out), out.numel(),
                                     _ptr(ws), nbytes, _stream()))
        total = int(off[-1]) if G else 0
        if total > out.numel():
            raise ValueError("output buffer too small")
        return Ragged(out[:total], off, lengths)
    check(lib.gtok_zinc_lengths(C.byref(bc), C.byref(zv), _ptr(lengths), _stream()))
    off = scan_offsets(lengths)
    total = int(off[-1]) if G else 0
    if out is None:
        out = torch.empty(total, dtype=torch.int64, device=dev)
    elif out.numel() < total:
        raise ValueError("output buffer too small")
    check(lib.gtok_zinc_serialize(C.byref(bc), C.byref(zv), _ptr(off), _ptr(out), _stream()))
    return Ragged(out[:total], off, lengths)


# ---------------------------------------------------------------------------
# K4: collate  (data_loader.py:488-497, train_agtt.py:276-302)
# ---------------------------------------------------------------------------
def pad_mask(tokens: torch.Tensor, row_start: torch.Tensor, row_len: torch.Tensor, pad_id: int,
             batch_idx: Optional[torch.Tensor] = None, lmax: Optional[int] = None):
    """One batch: -> (X int64[B,Lmax], attn bool[B,Lmax])."""
    _require_cuda(tokens, row_start, row_len, batch_idx)
    B = int(batch_idx.numel() if batch_idx is not None else row_len.numel())
    if lmax is None:
        sel = row_len if batch_idx is None else row_len[batch_idx]
        lmax = int(sel.max()) if B else 0
    X = torch.empty((B, lmax), dtype=torch.int64, device=tokens.device)
    attn = torch.empty((B, lmax), dtype=torch.bool, device=tokens.device)
    check(lib.gtok_pad_mask(_ptr(tokens), _ptr(row_start), _ptr(row_len), _ptr(batch_idx), B, lmax, int(pad_id),
                            _ptr(X), _ptr(attn), _stream()))
    return X, attn


@dataclass
class PaddedEpoch:
    """All batches of one pass, padded per batch exactly as collate() would."""
    X: torch.Tensor           # int64 flat
    attn: torch.Tensor        # bool flat
    xoff: torch.Tensor        # int64 [nb+1] (device)
    lmax: torch.Tensor        # int32 [nb]   (device)
    batch_size: int
    num_rows: int
    _host: Optional[tuple] = None

    @property
    def num_batches(self) -> int:
        return int(self.lmax.numel())

    def _h(self):
        if self._host is None:
            self._host = (self.xoff.cpu().tolist(), self.lmax.cpu().tolist())
        return self._host

    def batch(self, k: int):
        xoff, lmax = self._h()
        rows = min(self.batch_size, self.num_rows - k * self.batch_size)
        sl = slice(xoff[k], xoff[k + 1])
        return self.X[sl].view(rows, lmax[k]), self.attn[sl].view(rows, lmax[k])


def pad_mask_epoch(tokens: torch.Tensor, row_start: torch.Tensor, row_len: torch.Tensor, pad_id: int,
                   batch_size: int, batch_idx: Optional[torch.Tensor] = None,
                   X: Optional[torch.Tensor] = None, attn: Optional[torch.Tensor] = None) -> PaddedEpoch:
    _require_cuda(tokens, row_start, row_len, batch_idx)
    dev = tokens.device
    G = int(batch_idx.numel() if batch_idx is not None else row_len.numel())
    B = int(batch_size)
    nb = (G + B - 1) // B
    lmax = torch.empty(nb, dtype=torch.int32, device=dev)
    cells = torch.empty(nb, dtype=torch.int32, device=dev)
    check(lib.gtok_batch_extents(_ptr(row_len), _ptr(batch_idx), G, B, _ptr(lmax), _ptr(cells), _stream()))
    xoff = scan_offsets(cells)
    if X is None or attn is None:
        total = int(xoff[-1]) if nb else 0
        X = torch.empty(total, dtype=torch.int64, device=dev)
        attn = torch.empty(total, dtype=torch.bool, device=dev)
    check(lib.gtok_pad_mask_batched(_ptr(tokens), _ptr(row_start), _ptr(row_len), _ptr(batch_idx), G, B, _ptr(lmax),
                                    _ptr(xoff), int(pad_id), _ptr(X), _ptr(attn), _stream()))
    return PaddedEpoch(X, attn, xoff, lmax, B, G)


def zinc_collate_epoch(batch: GraphBatch, vocab: Dict[str, int], max_len: int, batch_size: int,
                       item_max_len: Optional[int] = None, max_nodes: Optional[int] = None,
                       max_entries: Optional[int] = None, X: Optional[torch.Tensor] = None,
                       attn: Optional[torch.Tensor] = None):
    """K2+K4 fused: molecules -> every padded batch of the epoch (tokenize_molecule + TokenDataset +
    collate, zinc_dataset_indexbase.py:143-227, data_loader.py:465-497); the ragged stream never exists.
    -> (PaddedEpoch, lengths int32[G]).  Batches of 32..256 (multiples of 32) molecule-sized graphs (<= 64 atoms, <= 256
    edge_index columns) take the single-pass kernel (gtok_zinc_collate, 0.58 ms per 1M molecules).  Other batch sizes of
    molecule-sized graphs: single-pass tokenizer + K4 (1.1 ms; measured faster than the route below, whose lengths / serialize
    kernels are the warp-per-molecule ones).  Bigger molecules: lengths -> extents -> scan -> pad cells + masks + row offsets
    -> tokens written straight at their padded position (gtok_pad_tail_batched + gtok_zinc_serialize_rows): no ragged stream."""
    _require_cuda(batch.src)
    if batch.x is None:
        raise ValueError("ZINC tokenization needs batch.x (atom types)")
    dev = batch.device
    G, B = batch.num_graphs, int(batch_size)
    nb = (G + B - 1) // B
    bc = _batch_c(batch, max_nodes, max_entries)
    if max_nodes is None and max_entries is None:
        bc.max_nodes_per_32, bc.max_entries_per_32 = group_bounds(batch.node_off), group_bounds(batch.edge_off)
    zv, tbl = zinc_vocab_c(vocab, max_len if item_max_len is None else item_max_len, max_len, bc.max_nodes, dev)
    lengths = torch.empty(G, dtype=torch.int32, device=dev)
    lmax = torch.empty(nb, dtype=torch.int32, device=dev)
    mol_sized = bc.max_nodes <= 64 and bc.max_entries <= 256
    single_pass = 32 <= B <= 256 and B % 32 == 0 and mol_sized
    if not single_pass and mol_sized and G > 0:
        rag = tokenize_zinc(batch, vocab, max_len, item_max_len=item_max_len, max_nodes=max_nodes, max_entries=max_entries)
        ep = pad_mask_epoch(rag.tokens, rag.offsets, rag.lengths, int(vocab["<pad>"]), B, X=X, attn=attn)
        return ep, rag.lengths
    if not single_pass:
        if G == 0:
            z = torch.zeros(1, dtype=torch.int64, device=dev)
            return PaddedEpoch(torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.bool, device=dev), z, lmax, B, 0), lengths
        check(lib.gtok_zinc_lengths(C.byref(bc), C.byref(zv), _ptr(lengths), _stream()))
        cells = torch.empty(nb, dtype=torch.int32, device=dev)
        check(lib.gtok_batch_extents(_ptr(lengths), None, G, B, _ptr(lmax), _ptr(cells), _stream()))
        xoff = scan_offsets(cells)
        total = int(xoff[-1])
        if X is None or attn is None:
            X = torch.empty(total, dtype=torch.int64, device=dev)
            attn = torch.empty(total, dtype=torch.bool, device=dev)
        elif total > min(X.numel(), attn.numel()):
            raise ValueError("padded output buffers too small")
        row_off = torch.empty(G + 1, dtype=torch.int64, device=dev)
        check(lib.gtok_pad_tail_batched(_ptr(lengths), G, B, _ptr(lmax), _ptr(xoff), int(vocab["<pad>"]), _ptr(X), _ptr(attn),
                                        _ptr(row_off), _stream()))
        check(lib.gtok_zinc_serialize_rows(C.byref(bc), C.byref(zv), _ptr(row_off), _ptr(lengths), _ptr(X), _stream()))
        return PaddedEpoch(X[:total], attn[:total], xoff, lmax, B, G), lengths
    xoff = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
    if X is None or attn is None:
        cap = G * min(max_len, 2 * bc.max_nodes + 4 * bc.max_entries + 4) + 2
        X = torch.empty(cap, dtype=torch.int64, device=dev)
        attn = torch.empty(cap, dtype=torch.bool, device=dev)
    nbytes = lib.gtok_zinc_collate_workspace_bytes(G, B)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib.gtok_zinc_collate(C.byref(bc), C.byref(zv), B, int(vocab["<pad>"]), _ptr(lengths), _ptr(lmax), _ptr(xoff),
                                _ptr(X), _ptr(attn), min(X.numel(), attn.numel()), _ptr(ws), nbytes, _stream()))
    total = int(xoff[-1]) if G else 0
    if total > min(X.numel(), attn.numel()):
        raise ValueError("padded output buffers too small")
    return PaddedEpoch(X[:total], attn[:total], xoff, lmax, B, G), lengths


# ---------------------------------------------------------------------------
# K3: SENT  (AutoGraph Graph2TrailTokenizer + train_agtt.py:171-273)
# ---------------------------------------------------------------------------
@dataclass
class CsrBatch:
    """Canonical undirected CSR of a GraphBatch (see include/gtok.h GtokCsrBatch)."""
    node_off: torch.Tensor
    rowptr: torch.Tensor
    col: torch.Tensor
    eattr: Optional[torch.Tensor]
    x: Optional[torch.Tensor]
    query_u: Optional[torch.Tensor]
    query_v: Optional[torch.Tensor]
    max_nodes: int
    max_nnz: int
    max_label: int = 0          # labels lie in [0, max_label]; 0 = unknown
    adj64: Optional[torch.Tensor] = None   # int64[N_total * adj_words] adjacency bit rows (graphs of <= 256 nodes)
    adj_words: int = 1                     # 64-bit words per adjacency row (1: <= 64 nodes, 2: <= 128, 4: <= 256)
    order: Optional[torch.Tensor] = None   # int32[G] processing order (graphs sorted by work)
    scratch: Optional[torch.Tensor] = None # walk-state workspace of the large-graph SENT kernel (lazily allocated)

    @property
    def num_graphs(self) -> int:
        return int(self.node_off.numel()) - 1

    def input_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.node_off, self.rowptr, self.col, self.eattr, self.x) if t is not None)

    def sent_input_bytes(self, labeled: bool) -> int:
        """Bytes the SENT kernel has to read per pass (bit rows replace rowptr+col when present)."""
        b = self.node_off.numel() * 8
        if self.adj64 is not None:
            b += self.adj64.numel() * 8
            if labeled:
                b += self.rowptr.numel() * 4
        else:
            b += self.rowptr.numel() * 4 + self.col.numel() * 4
        if labeled:
            b += self.x.numel() * 4 + self.eattr.numel() * 4
        return b


def _finish_csr(batch: GraphBatch, rowptr, col_local, eattr, adj64, adj_words, balance_window: int) -> CsrBatch:
    G = batch.num_graphs
    dev = batch.device
    n = batch.num_nodes()
    nnz_g = (rowptr[batch.node_off[1:]] - rowptr[batch.node_off[:-1]]).to(torch.int64)
    max_label = 0
    if batch.x is not None and eattr is not None and batch.x.numel() and eattr.numel():
        lo = min(int(batch.x.min()), int(eattr.min()))
        hi = max(int(batch.x.max()), int(eattr.max()))
        max_label = hi if lo >= 0 else 0
    order = None
    if adj64 is not None and G and balance_window > 1:
        # one thread walks one graph: give the 32 lanes of a warp graphs of similar work, but only reorder
        # inside windows of consecutive graphs - a global sort scatters a warp's reads and rows over the whole
        # buffers and was measured 10-30 % SLOWER than natural order
        if rowptr.is_cuda:
            order = torch.empty(G, dtype=torch.int32, device=dev)
            check(lib.gtok_balance_order(_ptr(batch.node_off), _ptr(rowptr.to(torch.int32)), G, min(int(balance_window), 4096),
                                         _ptr(order), _stream()))
        else:                                             # host tensors (CPU tests of the sort-based reference build)
            w = torch.arange(G, device=dev) // int(balance_window)
            order = torch.argsort(w * (1 << 40) + n * (1 << 20) + nnz_g.clamp(max=(1 << 20) - 1), stable=True).to(torch.int32)
    return CsrBatch(batch.node_off, rowptr.to(torch.int32), col_local, eattr, batch.x, batch.query_u, batch.query_v,
                    int(n.max()) if G else 1, int(nnz_g.max()) if G else 0, max_label, adj64, adj_words, order)


def _build_csr_native(batch: GraphBatch, balance_window: int) -> CsrBatch:
    """K7 (csrc/gtok_csr.cu): graphs of at most 256 nodes - adjacency bit rows by atomicOr, row pointers by popcount +
    K0 scan, neighbour lists by bit enumeration, labels from the first column of every pair (atomicMin)."""
    dev = batch.device
    G = batch.num_graphs
    Ntot = int(batch.node_off[-1]) if G else 0
    max_nodes = int(batch.num_nodes().max()) if G else 1
    W = 1 if max_nodes <= 64 else (2 if max_nodes <= 128 else 4)
    bc = _batch_c(batch, max(1, max_nodes), 1)
    adj64 = torch.empty(Ntot * W, dtype=torch.int64, device=dev)
    degree = torch.empty(Ntot, dtype=torch.int32, device=dev)
    check(lib.gtok_csr_adjacency(C.byref(bc), W, Ntot, _ptr(adj64), _ptr(degree), _stream()))
    rowptr64 = scan_offsets(degree)
    nnz = int(rowptr64[-1]) if Ntot else 0
    rowptr = torch.empty(Ntot + 1, dtype=torch.int32, device=dev)
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    labeled = batch.edge_attr is not None
    eattr = torch.empty(nnz, dtype=torch.int32, device=dev) if labeled else None
    ws = torch.empty(nnz, dtype=torch.int64, device=dev) if labeled else None
    check(lib.gtok_csr_fill(C.byref(bc), W, Ntot, _ptr(adj64), _ptr(rowptr64), nnz, _ptr(rowptr), _ptr(col), _ptr(eattr),
                            _ptr(ws), _stream()))
    return _finish_csr(batch, rowptr, col, eattr, adj64, W, balance_window)


def _build_csr_large(batch: GraphBatch, balance_window: int) -> CsrBatch:
    """K7 for any graph size (csrc/gtok_csr_large.cu): degree upper bounds -> K0 scan -> scatter of (neighbour, column)
    keys -> per-row sort + first-key-per-neighbour -> K0 scan -> col / eattr.  No adjacency bit rows (graphs > 256 nodes)."""
    dev = batch.device
    G = batch.num_graphs
    Ntot = int(batch.node_off[-1]) if G else 0
    max_nodes = int(batch.num_nodes().max()) if G else 1
    bc = _batch_c(batch, max(1, max_nodes), 1)
    deg_ub = torch.empty(Ntot, dtype=torch.int32, device=dev)
    check(lib.gtok_csrl_count(C.byref(bc), Ntot, _ptr(deg_ub), _stream()))
    rowptr_ub = scan_offsets(deg_ub)
    total_ub = int(rowptr_ub[-1]) if Ntot else 0
    entries = torch.empty(max(1, total_ub), dtype=torch.int64, device=dev)
    cursor = torch.empty(max(1, Ntot), dtype=torch.int32, device=dev)
    degree = torch.empty(Ntot, dtype=torch.int32, device=dev)
    tbytes = lib.gtok_csrl_table_bytes(max(1, max_nodes))
    table = torch.empty(tbytes, dtype=torch.uint8, device=dev)
    check(lib.gtok_csrl_sort(C.byref(bc), Ntot, _ptr(rowptr_ub), _ptr(cursor), _ptr(entries), _ptr(degree), _ptr(table), tbytes,
                             _stream()))
    rowptr64 = scan_offsets(degree)
    nnz = int(rowptr64[-1]) if Ntot else 0
    if nnz >= 2 ** 31:
        raise ValueError("CSR with >= 2^31 entries: shard the dataset")
    rowptr = torch.empty(Ntot + 1, dtype=torch.int32, device=dev)
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    eattr = torch.empty(nnz, dtype=torch.int32, device=dev) if batch.edge_attr is not None else None
    check(lib.gtok_csrl_emit(C.byref(bc), Ntot, _ptr(rowptr_ub), _ptr(entries), _ptr(rowptr64), nnz, _ptr(rowptr), _ptr(col),
                             _ptr(eattr), _stream()))
    return _finish_csr(batch, rowptr, col, eattr, None, 1, balance_window)


def build_csr(batch: GraphBatch, balance_window: int = 0, native: Optional[bool] = None) -> CsrBatch:
    """edge_index (any orientation, duplicates allowed) -> canonical undirected CSR (+ adjacency bit rows up to 256
    nodes per graph).  One-off ingest per dataset (what to_undirected/coalesce does inside AutoGraph per call).  The
    label of {u,v} is the edge_attr of its FIRST directed entry in caller order (oracle: canonical_csr).
    Library kernels on the device (K7: adjacency bit rows up to 256 nodes per graph, csrc/gtok_csr.cu; per-row sort of
    (neighbour, column) keys beyond, csrc/gtok_csr_large.cu).  native=False is the sort-based reference build with
    torch ops that the tests compare against (and what host tensors get)."""
    dev = batch.device
    G = batch.num_graphs
    max_nodes_ = int(batch.num_nodes().max()) if G else 1
    if native is None:
        native = batch.node_off.is_cuda                              # (host tensors: the torch build, used by CPU tests)
    if native:
        _require_cuda(batch.node_off, batch.src)
        if max_nodes_ > 256:
            return _build_csr_large(batch, balance_window)
        return _build_csr_native(batch, balance_window)
    E = batch.src.numel()
    n = batch.num_nodes()
    Ntot = int(batch.node_off[-1]) if G else 0
    gid = torch.repeat_interleave(torch.arange(G, device=dev), batch.num_entries())
    u = batch.src.to(torch.int64) + batch.node_off[gid]
    v = batch.dst.to(torch.int64) + batch.node_off[gid]
    keep = u != v
    idx = torch.arange(E, device=dev)[keep]
    a, b = torch.minimum(u, v)[keep], torch.maximum(u, v)[keep]
    key = a * max(Ntot, 1) + b
    order = torch.argsort(key, stable=True)           # stable: first directed entry first
    ks = key[order]
    first = torch.ones_like(ks, dtype=torch.bool)
    first[1:] = ks[1:] != ks[:-1]
    sel = order[first]
    ua, ub = a[sel], b[sel]
    attr = None if batch.edge_attr is None else batch.edge_attr[idx[sel]]
    rows = torch.cat([ua, ub])
    cols = torch.cat([ub, ua])
    o2 = torch.argsort(rows * max(Ntot, 1) + cols)
    rows, cols = rows[o2], cols[o2]
    deg = torch.bincount(rows, minlength=Ntot)
    rowptr = torch.zeros(Ntot + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = deg.cumsum(0)
    if rowptr[-1] >= 2 ** 31:
        raise ValueError("CSR with >= 2^31 entries: shard the dataset")
    gnode = torch.repeat_interleave(torch.arange(G, device=dev), n)
    col_local = (cols - batch.node_off[gnode[rows]]).to(torch.int32)
    eattr = None if attr is None else torch.cat([attr, attr])[o2].to(torch.int32)
    max_nodes = int(n.max()) if G else 1
    adj64, adj_words = None, 1
    if max_nodes <= 256:
        # neighbours are unique, so the sum of distinct powers of two is the OR (bit 63 wraps to the sign bit)
        adj_words = 1 if max_nodes <= 64 else (2 if max_nodes <= 128 else 4)
        cl = col_local.to(torch.int64)
        adj64 = torch.zeros(Ntot * adj_words, dtype=torch.int64, device=dev)
        adj64.scatter_add_(0, rows * adj_words + (cl >> 6), torch.ones_like(rows) << (cl & 63))
    return _finish_csr(batch, rowptr, col_local, eattr, adj64, adj_words, balance_window)


def _csr_c(csr: CsrBatch) -> _lib.CsrBatchC:
    return _lib.CsrBatchC(csr.num_graphs, _ptr(csr.node_off), _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.eattr), _ptr(csr.x),
                          _ptr(csr.query_u), _ptr(csr.query_v), max(1, csr.max_nodes), max(0, csr.max_nnz),
                          max(0, csr.max_label), int(csr.adj_words), _ptr(csr.adj64), _ptr(csr.order))


@dataclass
class SentConfig:
    idx_offset: int = 6
    node_idx_offset: int = 0
    edge_idx_offset: int = 0
    max_length: int = 600
    labeled: bool = False
    edge_label_shift: int = 0
    remap_zinc: bool = False
    append_query: bool = False
    append_eos: bool = True              # Graph2TrailTokenizer(append_eos=...)
    restart: str = "open_first"          # oracle.sent_oracle.SentSpec.restart: "open_first" | "unvisited"
    bracket_order: str = "visit_asc"     # SentSpec.bracket_order: "visit_asc" | "visit_desc"

    RESTART = {"open_first": 0, "unvisited": 1}
    BRACKET = {"visit_asc": 0, "visit_desc": 1}


def compact_rows(tok: torch.Tensor, lengths: torch.Tensor):
    """Strided rows [G, stride] + lengths -> (ragged int64 stream, offsets int64[G+1]) with library kernels (K0 scan +
    gtok_compact_rows) - what boolean-mask indexing would do with torch ops."""
    _require_cuda(tok, lengths)
    G, stride = int(tok.shape[0]), int(tok.shape[1])
    off = scan_offsets(lengths)
    total = int(off[-1]) if G else 0
    out = torch.empty(total, dtype=torch.int64, device=tok.device)
    if total:
        check(lib.gtok_compact_rows(_ptr(tok), stride, _ptr(lengths), _ptr(off), G, _ptr(out), _stream()))
    return out, off


def render_texts(rag: "Ragged", words: Sequence[str]):
    """Ragged ids -> texts, rendered on the GPU (gtok_detok_lengths -> K0 scan -> gtok_detok_render): one ASCII byte
    buffer holding word(tok) + ' ' for every token, and the char offset of every ROW boundary.
    -> (bytes, row_char_off list[G+1]); text of row g = buf[off[g]:off[g+1]-1]."""
    _require_cuda(rag.tokens)
    dev = rag.tokens.device
    n = int(rag.tokens.numel())
    enc = [w.encode("ascii") for w in words]
    woff = torch.zeros(len(enc) + 1, dtype=torch.int32)
    woff[1:] = torch.tensor([len(b) for b in enc], dtype=torch.int32).cumsum(0)
    wbytes = torch.frombuffer(bytearray(b"".join(enc) or b" "), dtype=torch.uint8).to(dev)
    woff_d = woff.to(dev)
    chars = torch.empty(n, dtype=torch.int32, device=dev)
    check(lib.gtok_detok_lengths(_ptr(rag.tokens), n, _ptr(woff_d), len(enc), _ptr(chars), _stream()))
    coff = scan_offsets(chars)
    total = int(coff[-1]) if n else 0
    text = torch.empty(max(1, total), dtype=torch.uint8, device=dev)
    check(lib.gtok_detok_render(_ptr(rag.tokens), n, _ptr(wbytes), _ptr(woff_d), len(enc), _ptr(coff), _ptr(text), _stream()))
    row_off = coff[rag.offsets.clamp(max=n)].cpu().tolist()
    return text[:total].cpu().numpy().tobytes(), row_off


def first_occurrence(tokens: torch.Tensor, num_ids: int) -> torch.Tensor:
    """first[id] = smallest index i with tokens[i] == id, -1 when the id never occurs (int64[num_ids], device)."""
    _require_cuda(tokens)
    first = torch.empty(num_ids, dtype=torch.int64, device=tokens.device)
    check(lib.gtok_first_occurrence(_ptr(tokens), int(tokens.numel()), int(num_ids), _ptr(first), _stream()))
    return first                                            # UINT64_MAX reads as -1


def ragged_append(rag: "Ragged", tail: torch.Tensor, tail_len: torch.Tensor, max_len: int) -> "Ragged":
    """row r = (row r ++ tail[r, :tail_len[r]])[:max_len] - TokenDataset(strip_label=False): the ids of the tokens after
    <p> (label, <eos>) re-attached to the stripped rows the tokenizers produce."""
    _require_cuda(rag.tokens, tail, tail_len)
    G = rag.num_rows
    dev = rag.tokens.device
    new_len = torch.empty(G, dtype=torch.int32, device=dev)
    check(lib.gtok_ragged_append_lengths(_ptr(rag.lengths), _ptr(tail_len), G, int(max_len), _ptr(new_len), _stream()))
    new_off = scan_offsets(new_len)
    total = int(new_off[-1]) if G else 0
    out = torch.empty(total, dtype=torch.int64, device=dev)
    if total:
        check(lib.gtok_ragged_append(_ptr(rag.tokens), _ptr(rag.offsets), _ptr(rag.lengths), _ptr(tail), int(tail.shape[1]),
                                     _ptr(new_off), G, _ptr(out), _stream()))
    return Ragged(out, new_off, new_len)


def gather_graphs(batch: GraphBatch, idx: torch.Tensor) -> GraphBatch:
    """The graphs idx[0], idx[1], ... of a CUDA batch as a new batch (any order, repeats allowed): gtok_segment_counts
    -> K0 scan -> gtok_gather_segments per array.  Raises IndexError on an index outside the batch."""
    _require_cuda(batch.src, idx)
    dev = batch.device
    idx = idx.to(torch.int64).contiguous()
    M, G = int(idx.numel()), batch.num_graphs
    status = torch.zeros(1, dtype=torch.int32, device=dev)

    def offsets(off):
        cnt = torch.empty(M, dtype=torch.int32, device=dev)
        check(lib.gtok_segment_counts(_ptr(off), G, _ptr(idx), M, _ptr(cnt), _ptr(status), _stream()))
        return scan_offsets(cnt)
    node_off, edge_off = offsets(batch.node_off), offsets(batch.edge_off)
    if int(status):
        raise IndexError("gather_graphs: index outside the batch")
    one = torch.ones(M + 1, dtype=torch.int64, device=dev).cumsum(0) - 1          # offsets of one item per graph

    def gather(arr, src_off, dst_off):
        if arr is None:
            return None
        out = torch.empty(int(dst_off[-1]) if M else 0, dtype=torch.int32, device=dev)
        check(lib.gtok_gather_segments(_ptr(arr), _ptr(src_off), G, _ptr(idx), _ptr(dst_off), M, _ptr(out), _stream()))
        return out
    per_graph = lambda a: gather(a, one_src, one)
    one_src = torch.ones(G + 1, dtype=torch.int64, device=dev).cumsum(0) - 1
    y = None
    if batch.y is not None:
        y = batch.y.to(dev)[idx]                          # labels never enter a kernel (plain indexing of G values)
    return GraphBatch(node_off, edge_off, gather(batch.src, batch.edge_off, edge_off), gather(batch.dst, batch.edge_off, edge_off),
                      gather(batch.x, batch.node_off, node_off), gather(batch.edge_attr, batch.edge_off, edge_off),
                      per_graph(batch.query_u), per_graph(batch.query_v), y)


def _sent_switch(table: Dict[str, int], value: str) -> int:
    if value not in table:
        raise ValueError(f"the CUDA SENT kernels implement {sorted(table)} (got {value!r}); the other settings of "
                         "oracle.sent_oracle.SentSpec are oracle-only")
    return table[value]


def tokenize_sent(csr: CsrBatch, cfg: SentConfig, seed: int = 0, graph_id_base: int = 0,
                  choice_words: Optional[torch.Tensor] = None, choice_off: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None, stride: Optional[int] = None, use_workspace: Optional[bool] = None):
    """-> (tokens int64[G, stride], lengths int32[G]); row g valid up to lengths[g].
    use_workspace: graphs above 256 nodes can walk with their per-node state in an HBM scratch array (8 bytes per node,
    allocated once per CsrBatch; gtok_sent_tokenize_ws) instead of shared memory (gtok_sent_tokenize).  None = by
    measurement: the shared-memory kernel wins while its per-graph slice leaves >= ~16 graphs per SM (12 KB; 14.6 vs
    15.8 ms per 32M nodes at 2048 nodes), the workspace kernel beyond (18.2 vs 28.7 ms at 4096 nodes)."""
    _require_cuda(csr.node_off, csr.rowptr, csr.col)
    dev = csr.node_off.device
    G = csr.num_graphs
    need = cfg.max_length + (3 if cfg.append_query else 0)
    if stride is None:
        stride = need + (need & 1)          # even stride keeps every row 16-byte aligned
    if out is None:
        out = torch.empty((G, stride), dtype=torch.int64, device=dev)
    lengths = torch.empty(G, dtype=torch.int32, device=dev)
    cc = _csr_c(csr)
    sc = _lib.SentCfgC(cfg.idx_offset, cfg.node_idx_offset, cfg.edge_idx_offset, cfg.max_length, int(cfg.labeled),
                       cfg.edge_label_shift, int(cfg.remap_zinc), int(cfg.append_query), _sent_switch(cfg.RESTART, cfg.restart),
                       _sent_switch(cfg.BRACKET, cfg.bracket_order), int(not cfg.append_eos), 0)
    ch = _lib.ChoiceC(int(seed) & (2 ** 64 - 1), int(graph_id_base), _ptr(choice_words), _ptr(choice_off))
    n_total = int(csr.rowptr.numel()) - 1
    if use_workspace is None:
        # <= 64 nodes: the workspace is the walk kernel's item log (walk + emission kernels, csrc/gtok_sent_walk.cu)
        use_workspace = csr.max_nodes <= 256 or csr.max_nodes * (4 * (2 if cfg.labeled else 1) + 0.375) > 12288
    nbytes = lib.gtok_sent_workspace_bytes(n_total, max(1, csr.max_nodes)) if use_workspace else 0
    if nbytes:
        if csr.scratch is None or csr.scratch.numel() < nbytes:
            csr.scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)      # kept with the CSR: every epoch reuses it
        check(lib.gtok_sent_tokenize_ws(C.byref(cc), C.byref(sc), C.byref(ch), _ptr(lengths), _ptr(out), int(stride), n_total,
                                        _ptr(csr.scratch), nbytes, _stream()))
    else:
        check(lib.gtok_sent_tokenize(C.byref(cc), C.byref(sc), C.byref(ch), _ptr(lengths), _ptr(out), int(stride), _stream()))
    return out, lengths


def _sent_c(cfg: SentConfig, seed: int, graph_id_base: int, choice_words, choice_off):
    sc = _lib.SentCfgC(cfg.idx_offset, cfg.node_idx_offset, cfg.edge_idx_offset, cfg.max_length, int(cfg.labeled),
                       cfg.edge_label_shift, int(cfg.remap_zinc), int(cfg.append_query), _sent_switch(cfg.RESTART, cfg.restart),
                       _sent_switch(cfg.BRACKET, cfg.bracket_order), int(not cfg.append_eos), 0)
    ch = _lib.ChoiceC(int(seed) & (2 ** 64 - 1), int(graph_id_base), _ptr(choice_words), _ptr(choice_off))
    return sc, ch


def sent_collate_supported(csr: CsrBatch, cfg: SentConfig) -> bool:
    """True when the walk / emission kernels take this configuration (graphs up to 64 nodes with adjacency bit rows,
    tokens below 65536): sent_collate_epoch then writes the padded batches without the strided int64 rows."""
    sc, _ = _sent_c(cfg, 0, 0, None, None)
    cc = _csr_c(csr)
    return lib.gtok_sent_log_bytes(C.byref(cc), C.byref(sc), int(csr.rowptr.numel()) - 1) > 0


def sent_collate_epoch(csr: CsrBatch, cfg: SentConfig, pad_id: int, batch_size: int, seed: int = 0, graph_id_base: int = 0,
                       batch_idx: Optional[torch.Tensor] = None, choice_words: Optional[torch.Tensor] = None,
                       choice_off: Optional[torch.Tensor] = None, X: Optional[torch.Tensor] = None,
                       attn: Optional[torch.Tensor] = None):
    """K3+K4 fused: Graph2TrailTokenizer.__call__ + remap / query append + collate_fn of every batch of an epoch
    (train_agtt.py:246-302).  The walk kernel leaves an item log + row lengths, K4b's extents / scan give the batch
    geometry, the emission kernel expands every row straight into its padded batch.  -> (PaddedEpoch, lengths int32[G]);
    identical to tokenize_sent + pad_mask_epoch."""
    _require_cuda(csr.node_off, csr.rowptr, batch_idx)
    dev = csr.node_off.device
    G = csr.num_graphs
    sc, ch = _sent_c(cfg, seed, graph_id_base, choice_words, choice_off)
    cc = _csr_c(csr)
    n_total = int(csr.rowptr.numel()) - 1
    nbytes = lib.gtok_sent_log_bytes(C.byref(cc), C.byref(sc), n_total)
    if not nbytes:
        raise ValueError("sent_collate_epoch: configuration not supported by the walk kernels (see sent_collate_supported)")
    if csr.scratch is None or csr.scratch.numel() < nbytes:
        csr.scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)          # kept with the CSR: every epoch reuses it
    lengths = torch.empty(G, dtype=torch.int32, device=dev)
    check(lib.gtok_sent_walk(C.byref(cc), C.byref(sc), C.byref(ch), n_total, _ptr(csr.scratch), nbytes, _ptr(lengths), _stream()))
    R = int(batch_idx.numel() if batch_idx is not None else G)
    B = int(batch_size)
    nb = (R + B - 1) // B
    lmax = torch.empty(nb, dtype=torch.int32, device=dev)
    cells = torch.empty(nb, dtype=torch.int32, device=dev)
    check(lib.gtok_batch_extents(_ptr(lengths), _ptr(batch_idx), R, B, _ptr(lmax), _ptr(cells), _stream()))
    xoff = scan_offsets(cells)
    if X is None or attn is None:
        total = int(xoff[-1]) if nb else 0
        X = torch.empty(total, dtype=torch.int64, device=dev)
        attn = torch.empty(total, dtype=torch.bool, device=dev)
    check(lib.gtok_sent_emit_collate(C.byref(cc), C.byref(sc), n_total, _ptr(csr.scratch), _ptr(lengths), _ptr(batch_idx), R, B,
                                     _ptr(lmax), _ptr(xoff), int(pad_id), _ptr(X), _ptr(attn), _stream()))
    return PaddedEpoch(X, attn, xoff, lmax, B, R), lengths


# ---------------------------------------------------------------------------
# K6: around the path: query positions, task labels, on-device ER graphs
# ---------------------------------------------------------------------------
def query_positions(ep: PaddedEpoch, token: int = 0, token_per_batch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """First position of the <q> id in every row of a padded epoch (-1 when absent): what extract_query_nodes scans
    for on the host, row by row (trainer/train_ibtt.py:89-103; trainer/train_agtt.py:101-112 with the per-batch id
    idx_offset + num_nodes of the batch's first graph, :131).  -> int32[num_rows] in the epoch's row order."""
    _require_cuda(ep.X, token_per_batch)
    pos = torch.empty(ep.num_rows, dtype=torch.int32, device=ep.X.device)
    if token_per_batch is not None:
        if token_per_batch.dtype != torch.int64 or token_per_batch.numel() != ep.num_batches:
            raise ValueError("token_per_batch: one int64 id per batch expected")
    check(lib.gtok_find_token(_ptr(ep.X), _ptr(ep.xoff), _ptr(ep.lmax), ep.num_rows, ep.batch_size, _ptr(token_per_batch),
                              int(token), _ptr(pos), _stream()))
    return pos


def graph_labels(csr: CsrBatch, edge_off: Optional[torch.Tensor] = None, cycle: bool = True, distance: bool = False):
    """Task labels from the canonical CSR: has_cycle int32[G] (the reference's DFS oracle, test/train_performance_test.py:27-67,
    on an edge list of edge_off[g+1]-edge_off[g] entries; None = simple graph) and the BFS hop count query_u -> query_v
    int32[G] (-1 = INF).  -> (has_cycle | None, dist | None)."""
    _require_cuda(csr.node_off, csr.rowptr, edge_off)
    dev = csr.node_off.device
    G = csr.num_graphs
    if distance and (csr.query_u is None or csr.query_v is None):
        raise ValueError("distance labels need query_u / query_v")
    cyc = torch.empty(G, dtype=torch.int32, device=dev) if cycle else None
    dist = torch.empty(G, dtype=torch.int32, device=dev) if distance else None
    cc = _csr_c(csr)
    check(lib.gtok_graph_labels(C.byref(cc), _ptr(edge_off), _ptr(cyc), _ptr(dist), _stream()))
    return cyc, dist


ER_BUCKETS = ((10, 20), (20, 40), (40, 50))        # docs/synthetic_data.md:130-138


GEN_KINDS = {"er": 0, "sbm": 1, "complete": 2, "star": 3, "path": 4, "ba": 5, "sfn": 6}      # include/gtok.h GTOK_GEN_*


def er_spec(seed: int, graph_id_base: int = 0, buckets=ER_BUCKETS, p_min: float = 0.1, p_max: float = 0.2,
            kind: str = "er", num_blocks: int = 2, out_ratio: float = 0.25) -> _lib.ErSpecC:
    if not 1 <= len(buckets) <= 4:
        raise ValueError("1..4 size buckets expected")
    if kind not in GEN_KINDS:
        raise ValueError(f"unknown generator {kind!r} (have {sorted(GEN_KINDS)})")
    thr = lambda p: max(0, min(0xFFFFFFFF, int(math.floor(float(p) * 4294967296.0))))
    s = _lib.ErSpecC()
    s.num_buckets = len(buckets)
    for i, (lo, hi) in enumerate(buckets):
        s.n_lo[i], s.n_hi[i] = int(lo), int(hi)
    s.thr_lo, s.thr_hi = thr(p_min), thr(p_max)
    s.kind, s.num_blocks, s.out_ratio_q16 = GEN_KINDS[kind], int(num_blocks), max(0, min(65536, int(round(out_ratio * 65536))))
    s.seed, s.graph_id_base = int(seed) & (2 ** 64 - 1), int(graph_id_base)
    return s


def generate_er(num_graphs: int, seed: int = 0, graph_id_base: int = 0, buckets=ER_BUCKETS, p_min: float = 0.1,
                p_max: float = 0.2, device="cuda", kind: str = "er", num_blocks: int = 2, out_ratio: float = 0.25) -> GraphBatch:
    """Graph-token-shaped graphs generated ON the device (docs/synthetic_data.md:93-94,130-138 distribution; stream =
    oracle/tasks_oracle.py er_graph): every undirected edge once, u < v ascending, plus a query pair per graph.
    kind: "er" (default), "sbm" (num_blocks contiguous blocks, p inside / p * out_ratio across), "complete", "star",
    "path", "ba" (Barabasi-Albert, m ~ U{1..num_blocks} attachments per new node, graphs of at most 64 nodes), "sfn" (Holme-Kim
    scale-free graphs: ba + triad steps with probability out_ratio - the test algorithm of configs/ibtt_graph_token.yaml:5).  Keyed by the global graph id, so shards of a larger job generate their own slice."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("gtok has no CPU path: a CUDA device is required")
    G = int(num_graphs)
    spec = er_spec(seed, graph_id_base, buckets, p_min, p_max, kind, num_blocks, out_ratio)
    nn = torch.empty(G, dtype=torch.int32, device=dev)
    ne = torch.empty(G, dtype=torch.int32, device=dev)
    qu = torch.empty(G, dtype=torch.int32, device=dev)
    qv = torch.empty(G, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.gtok_er_sizes(C.byref(spec), G, _ptr(nn), _ptr(ne), _ptr(qu), _ptr(qv), _stream()))
        node_off, edge_off = scan_offsets(nn), scan_offsets(ne)
        E = int(edge_off[-1]) if G else 0
        src = torch.empty(E, dtype=torch.int32, device=dev)
        dst = torch.empty(E, dtype=torch.int32, device=dev)
        check(lib.gtok_er_edges(C.byref(spec), G, _ptr(edge_off), _ptr(src), _ptr(dst), _stream()))
    return GraphBatch(node_off, edge_off, src, dst, query_u=qu, query_v=qv)
