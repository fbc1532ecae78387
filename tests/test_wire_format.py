"""Host-side wire format (gtok/batch.py PackedGraphBatch): lossless narrowing, and the proof obligation of the
de-mirrored column list - only a canonical undirected PyG edge_index may drop its dst < src half, and what remains
must be exactly the first-seen edges tokenize_molecule would emit (checked against the ZINC oracle's text)."""
import dataclasses

import numpy as np
import torch

from gtok import synth
from gtok.batch import GraphBatch, PackedGraphBatch
from oracle import ibtt_oracle as O


def _mols(rng, G, canonical=True):
    mols = []
    for _ in range(G):
        n = int(rng.integers(2, 30))
        pairs = {(int(min(a, b)), int(max(a, b))) for a, b in zip(rng.integers(0, n, 2 * n), rng.integers(0, n, 2 * n)) if a != b}
        cols = sorted([(u, v) for u, v in pairs] + [(v, u) for u, v in pairs])
        attr = {p: int(rng.integers(1, 5)) for p in pairs}
        mols.append({"n": n, "src": [c[0] for c in cols], "dst": [c[1] for c in cols],
                     "attr": [attr[(min(c), max(c))] for c in cols], "x": rng.integers(0, 28, n).tolist()})
    return mols


def _batch(mols):
    return GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                                 x_lists=[m["x"] for m in mols], attr_lists=[m["attr"] for m in mols])


def test_canonical_batches_drop_the_mirror_half_and_keep_first_seen_edges():
    rng = np.random.default_rng(0)
    mols = _mols(rng, 200)
    gb = _batch(mols)
    pk = PackedGraphBatch.from_batch(gb, fuse_columns=True, drop_mirror_columns=True)
    full = PackedGraphBatch.from_batch(gb, fuse_columns=True)
    assert pk.unique_columns and not full.unique_columns
    assert pk.col16.numel() * 2 == full.col16.numel() and pk.num_entries_total == pk.col16.numel()
    assert pk.entry_cnt.to(torch.int64).tolist() == [len(m["src"]) // 2 for m in mols]
    # the kept columns, decoded, are the oracle's first-seen edges (orientation and bond type of the first column)
    vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in range(30)] + ["X", "unknown"])
    itos = {i: t for t, i in vocab.items()}
    w = pk.col16.to(torch.int32) & 0xFFFF
    off = 0
    for m in mols:
        k = len(m["src"]) // 2
        cols = [(int(c) & 63, (int(c) >> 6) & 63, int(c) >> 12) for c in w[off:off + k]]
        off += k
        toks = [itos[i] for i in O.zinc_ids(m["x"], m["src"], m["dst"], m["attr"], vocab, 4096, 4096).tolist()]
        bonds = [(int(toks[i + 2]), int(toks[i + 3]), O.ZINC_BONDS.index(toks[i + 1]) + 1)
                 for i, t in enumerate(toks) if t == "<bond>"]
        assert cols == bonds
    # the synthetic BASELINE workload is canonical too
    assert PackedGraphBatch.from_batch(synth.zinc_shaped(500, seed=1), True, True).unique_columns


def test_non_canonical_batches_keep_every_column():
    rng = np.random.default_rng(1)
    mols = _mols(rng, 50)
    gb = _batch(mols)
    E = gb.src.numel()

    def packed(**kw):
        return PackedGraphBatch.from_batch(dataclasses.replace(gb, **kw), fuse_columns=True, drop_mirror_columns=True)

    assert packed().unique_columns
    # two columns of one graph swapped: still symmetric, no longer sorted
    e0 = int(gb.edge_off[3])
    perm = torch.arange(E)
    perm[e0], perm[e0 + 1] = e0 + 1, e0
    assert not packed(src=gb.src[perm], dst=gb.dst[perm], edge_attr=gb.edge_attr[perm]).unique_columns
    # bond type differs between the two directions of one edge
    ea = gb.edge_attr.clone()
    ea[e0] = ea[e0] % 4 + 1
    assert not packed(edge_attr=ea).unique_columns
    # one direction replaced by a self loop / a repeated column / a column of another pair
    for new_dst in (int(gb.src[e0]), None):
        d = gb.dst.clone()
        if new_dst is None:
            s = gb.src.clone()
            s[e0 + 1], d[e0 + 1] = s[e0], d[e0]
            assert not packed(src=s, dst=d).unique_columns
        else:
            d[e0] = new_dst
            assert not packed(dst=d).unique_columns
    # an edge listed in one direction only (odd column count in one graph, even total)
    m2 = [dict(m) for m in mols]
    for g in (0, 1):
        m2[g] = {k: (v[:-1] if k in ("src", "dst", "attr") else v) for k, v in m2[g].items()}
    assert not PackedGraphBatch.from_batch(_batch(m2), True, True).unique_columns
    # every such batch still packs losslessly with all its columns
    pk = packed(edge_attr=ea)
    assert pk.col16.numel() == E and pk.entry_cnt.to(torch.int64).tolist() == gb.num_entries().tolist()
    # without the flag nothing is dropped, whatever the batch looks like
    assert not PackedGraphBatch.from_batch(gb, fuse_columns=True).unique_columns


def test_inmemory_dataset_storage_converts_without_a_python_loop():
    """a PyG InMemoryDataset keeps its graphs concatenated (`_data` + `slices`, local node ids in
    edge_index); GraphBatch.from_collated must equal the item-by-item conversion, with and without `_indices`."""
    import types
    from gtok.batch import graph_batch_from_dataset
    rng = np.random.default_rng(4)
    mols = _mols(rng, 60)
    per_item = _batch(mols)
    per_item.y = torch.arange(60, dtype=torch.float32) * 0.5

    class FakeInMemory:
        def __init__(self, indices=None):
            self._data = types.SimpleNamespace(
                x=per_item.x.to(torch.int64).reshape(-1, 1),
                edge_index=torch.stack([per_item.src, per_item.dst]).to(torch.int64),
                edge_attr=per_item.edge_attr.to(torch.int64), y=per_item.y)
            self.slices = {"x": per_item.node_off.clone(), "edge_index": per_item.edge_off.clone(),
                           "edge_attr": per_item.edge_off.clone(), "y": torch.arange(61)}
            self._indices = indices

        def __len__(self):
            return 60 if self._indices is None else len(self._indices)

        def __getitem__(self, i):
            raise AssertionError("the collated storage must be used, not per-item access")

    def same(a, b):
        for f in ("node_off", "edge_off", "src", "dst", "x", "edge_attr", "y"):
            assert torch.equal(getattr(a, f), getattr(b, f)), f
        assert a.src.dtype == torch.int32 and a.x.dtype == torch.int32 and a.node_off.dtype == torch.int64

    same(graph_batch_from_dataset(FakeInMemory()), per_item)
    idx = [5, 3, 59, 0, 3]
    same(graph_batch_from_dataset(FakeInMemory(idx)), per_item.select(torch.tensor(idx)))
    # a plain sequence of Data-like items still converts item by item
    items = []
    for g in range(60):
        n, s_, d_, x_, a_ = per_item.graph(g)
        items.append(types.SimpleNamespace(x=x_.reshape(-1, 1), edge_index=torch.stack([s_, d_]), edge_attr=a_.reshape(-1, 1),
                                           y=per_item.y[g:g + 1]))
    same(graph_batch_from_dataset(items), per_item)
    bad = FakeInMemory()
    bad.slices["x"] = bad.slices["x"][:-1]
    try:
        graph_batch_from_dataset(bad)
        assert False
    except ValueError:
        pass
