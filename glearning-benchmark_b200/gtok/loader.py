"""GPU-resident batch iterators: one epoch = a handful of kernel launches, batches are views.

Stand in for the reference's per-batch Python path
    DataLoader(TokenDataset, batch_size, shuffle, collate_fn=lambda b: collate(b, pad_id))      trainer/train_ibtt.py:395-402
    DataLoader(TokenizedGraphDataset, batch_size, shuffle, collate_fn=collate_fn)               trainer/train_agtt.py:276-302, :600-640
(SURVEY.md 8f rank 1).  The reference tokenizes / pads one item / one batch at a time in the training process;
here the epoch's row order (the DataLoader's shuffle) goes to the device as a permutation and every batch of the
epoch is produced by the fused kernels at once.  Iterating yields exactly what the reference's collate functions
return - `(X int64[B, Lmax_b], attn bool[B, Lmax_b], Y)` with each batch padded to ITS OWN longest row - as CUDA
tensors (the trainer's `.to(device)` is a no-op).

    loader = IbttEpochLoader.from_graphs(graph_batch, vocab, "cycle_check", max_len=600, labels=y, batch_size=128, shuffle=True)
    for X, attn, Y in loader: ...

    loader = AgttEpochLoader(csr, sent_config, labels=y, batch_size=32, shuffle=True)        # fresh trails every epoch
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional, Tuple

import torch

from . import engine
from .batch import GraphBatch


class _EpochLoaderBase:
    def __init__(self, num_rows: int, labels, batch_size: int, shuffle: bool, seed: int, drop_last: bool, device):
        self.G, self.B = int(num_rows), int(batch_size)
        self.shuffle, self.drop_last = bool(shuffle), bool(drop_last)
        self.device = torch.device(device)
        self.epoch = 0
        self._gen = torch.Generator(device="cpu")
        self._gen.manual_seed(int(seed))
        if labels is None:
            labels = torch.zeros(self.G, dtype=torch.int64)
        self.labels = torch.as_tensor(labels).to(self.device)
        if self.labels.numel() != self.G:
            raise ValueError("one label per row expected")

    def __len__(self) -> int:
        return self.G // self.B if self.drop_last else (self.G + self.B - 1) // self.B

    def _order(self) -> Optional[torch.Tensor]:
        """The epoch's row order on the device (None = dataset order)."""
        rows = len(self) * self.B if self.drop_last else self.G
        if not self.shuffle:
            return None if rows == self.G else torch.arange(rows, device=self.device)
        perm = torch.randperm(self.G, generator=self._gen)[:rows]       # torch's DataLoader draws its permutation the same way
        return perm.to(self.device)

    def _yield(self, ep: engine.PaddedEpoch, order: Optional[torch.Tensor],
               q_pos: Optional[torch.Tensor] = None) -> Iterator[Tuple[torch.Tensor, ...]]:
        xoff = ep.xoff.cpu().tolist()
        lmax = ep.lmax.cpu().tolist()
        Y = self.labels if order is None else self.labels[order]
        for k in range(ep.num_batches):
            rows = min(self.B, ep.num_rows - k * self.B)
            sl = slice(xoff[k], xoff[k + 1])
            out = (ep.X[sl].view(rows, lmax[k]), ep.attn[sl].view(rows, lmax[k]), Y[k * self.B:k * self.B + rows])
            yield out if q_pos is None else out + (q_pos[k * self.B:k * self.B + rows],)


class IbttEpochLoader(_EpochLoaderBase):
    """IBTT batches of one epoch.  Two sources:
    * `from_graphs`  graph-token graphs resident as a GraphBatch: every epoch is ONE pass of gtok_ibtt_collate
                     over the (permuted) graphs - rows are re-serialized straight into the padded batches, no
                     ragged stream is kept (K1+K4 fused);
    * `from_ragged`  rows already tokenized (any tokenizer: ZINC, graph-token): gtok_pad_mask_batched with the
                     permutation as batch_idx (K4)."""

    def __init__(self, make_epoch, num_rows, labels, batch_size, shuffle=False, seed=0, drop_last=False, device="cuda",
                 query_token: Optional[int] = None):
        super().__init__(num_rows, labels, batch_size, shuffle, seed, drop_last, device)
        self._make_epoch = make_epoch
        self.query_token = query_token

    @classmethod
    def from_graphs(cls, batch: GraphBatch, vocab: Dict[str, int], task: str, max_len: int, labels=None, batch_size: int = 128,
                    shuffle: bool = False, seed: int = 0, drop_last: bool = False, with_q_pos: bool = False):
        """with_q_pos: batches come as (X, attn, Y, q_pos) - q_pos int32[rows] = first position of <q> in each row
        (-1 when truncation cut it off), what extract_query_nodes looks up per row (train_ibtt.py:89-103)."""
        max_nodes = int(batch.num_nodes().max()) if batch.num_graphs else 0

        def make(order):
            return engine.ibtt_collate_epoch(batch, vocab, task, max_len, batch_size, batch_idx=order, max_nodes=max_nodes)[0]
        return cls(make, batch.num_graphs, labels, batch_size, shuffle, seed, drop_last, batch.device,
                   query_token=int(vocab["<q>"]) if with_q_pos else None)

    @classmethod
    def from_ragged(cls, ragged: engine.Ragged, pad_id: int, labels=None, batch_size: int = 128, shuffle: bool = False,
                    seed: int = 0, drop_last: bool = False):
        def make(order):
            return engine.pad_mask_epoch(ragged.tokens, ragged.offsets, ragged.lengths, pad_id, batch_size, batch_idx=order)
        return cls(make, ragged.num_rows, labels, batch_size, shuffle, seed, drop_last, ragged.tokens.device)

    def __iter__(self):
        order = self._order()
        self.epoch += 1
        if order is not None and order.numel() == 0:
            return iter(())
        ep = self._make_epoch(order)
        return self._yield(ep, order, None if self.query_token is None else engine.query_positions(ep, self.query_token))


class AgttEpochLoader(_EpochLoaderBase):
    """AGTT batches of one epoch: fresh SENT trails for every graph (Philox seed = f(seed, epoch), as the reference
    re-samples a trail per item per epoch, train_agtt.py:246-273), fused remap / query append per the config, then
    collate_fn with pad id 5 (:285-286) under the epoch's row order."""

    PAD = 5

    def __init__(self, csr: engine.CsrBatch, config: engine.SentConfig, labels=None, batch_size: int = 32, shuffle: bool = False,
                 seed: int = 0, drop_last: bool = False, graph_id_base: int = 0, with_q_pos: bool = False):
        """with_q_pos: batches come as (X, attn, Y, q_pos); the id searched in batch k is the reference's
        idx_offset + num_nodes of the batch's FIRST graph (train_agtt.py:131), first hit per row (:101-112)."""
        super().__init__(csr.num_graphs, labels, batch_size, shuffle, seed, drop_last, csr.node_off.device)
        self.with_q_pos = bool(with_q_pos)
        self._num_nodes = csr.node_off[1:] - csr.node_off[:-1]
        self.csr, self.config, self.seed, self.graph_id_base = csr, config, int(seed), int(graph_id_base)
        need = config.max_length + (3 if config.append_query else 0)
        self.stride = need + (need & 1)
        self._fused = engine.sent_collate_supported(csr, config)
        if not self._fused:  # strided int64 rows + K4 (graphs above 64 nodes, tokens beyond 16 bits)
            self._buf = torch.empty((self.G, self.stride), dtype=torch.int64, device=self.device)
            self._row_start = torch.arange(self.G, device=self.device, dtype=torch.int64) * self.stride

    def __iter__(self):
        order = self._order()
        seed = (self.seed * 0x9E3779B97F4A7C15 + self.epoch) & (2 ** 64 - 1)
        self.epoch += 1
        if order is not None and order.numel() == 0:
            return iter(())
        if self._fused:      # walk -> item log -> emission straight into the padded batches (csrc/gtok_sent_walk.cu)
            ep, _ = engine.sent_collate_epoch(self.csr, self.config, self.PAD, self.B, seed=seed, graph_id_base=self.graph_id_base,
                                              batch_idx=order)
        else:
            _, ln = engine.tokenize_sent(self.csr, self.config, seed=seed, graph_id_base=self.graph_id_base, out=self._buf,
                                         stride=self.stride)
            ep = engine.pad_mask_epoch(self._buf.view(-1), self._row_start, ln, self.PAD, self.B, batch_idx=order)
        q_pos = None
        if self.with_q_pos:
            first = torch.arange(0, ep.num_rows, self.B, device=self.device)
            first = first if order is None else order[first]
            q_pos = engine.query_positions(ep, token_per_batch=self.config.idx_offset + self._num_nodes[first])
        return self._yield(ep, order, q_pos)
