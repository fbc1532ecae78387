"""End to end from the reference's OWN storage (int64 PyG tensors) on the GPU: the device-side ingest kernels
(gtok_pyg_pack_mol / gtok_pyg_narrow) and ZincIbttPipeline.run_pyg against the resident-layout path and the oracle."""
import numpy as np
import pytest
import torch

from oracle import ibtt_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _vocab():
    return O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in range(40)] + ["X", "unknown"])


@pytest.mark.parametrize("E,N", [(0, 0), (1, 1), (7, 15), (8, 16), (1000003, 400001)])
def test_pyg_pack_kernels_bitexact(dev, E, N):
    from gtok import _lib
    from gtok._lib import lib, check
    from gtok.engine import _ptr, _stream
    rng = np.random.default_rng(E + 1)
    s = torch.from_numpy(rng.integers(0, 64, E)).to(dev)
    d = torch.from_numpy(rng.integers(0, 64, E)).to(dev)
    a = torch.from_numpy(rng.integers(0, 16, E)).to(dev)
    x = torch.from_numpy(rng.integers(0, 256, N)).to(dev)
    col16 = torch.full((E + 16,), -1, dtype=torch.int16, device=dev)
    x8 = torch.full((N + 16,), 255, dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.gtok_pyg_pack_mol(_ptr(s), _ptr(d), _ptr(a), E, _ptr(x), N, _ptr(col16), _ptr(x8), _ptr(status), _stream()))
    want = (s | (d << 6) | (a << 12))
    assert torch.equal(col16[:E].to(torch.int64) & 0xFFFF, want) and bool((col16[E:] == -1).all())
    assert torch.equal(x8[:N].to(torch.int64), x) and bool((x8[N:] == 255).all())
    assert int(status) == 0
    # without edge_attr the attr field is zero
    check(lib.gtok_pyg_pack_mol(_ptr(s), _ptr(d), None, E, _ptr(x), N, _ptr(col16), _ptr(x8), _ptr(status), _stream()))
    assert torch.equal(col16[:E].to(torch.int64) & 0xFFFF, s | (d << 6)) and int(status) == 0
    # int64 -> int32
    v = torch.from_numpy(rng.integers(-2 ** 31, 2 ** 31, E)).to(dev)
    out = torch.full((E + 8,), -7, dtype=torch.int32, device=dev)
    check(lib.gtok_pyg_narrow(_ptr(v), E, _ptr(out), _ptr(status), _stream()))
    assert torch.equal(out[:E].to(torch.int64), v) and bool((out[E:] == -7).all()) and int(status) == 0
    if E >= 8:
        # every out-of-range class is reported, wherever it sits (vector body and tail)
        for pos in (0, E - 1, E // 2):
            for arr, bit in ((s, _lib.PYG_BAD_NODE_ID), (d, _lib.PYG_BAD_NODE_ID), (a, _lib.PYG_BAD_EDGE_ATTR)):
                keep = int(arr[pos])
                for badv in (64 if arr is not a else 16, -1, 1 << 40):
                    arr[pos] = badv
                    status.zero_()
                    check(lib.gtok_pyg_pack_mol(_ptr(s), _ptr(d), _ptr(a), E, _ptr(x), N, _ptr(col16), _ptr(x8), _ptr(status), _stream()))
                    assert int(status) == bit, (pos, badv)
                arr[pos] = keep
        x[N - 1] = 256
        status.zero_()
        check(lib.gtok_pyg_pack_mol(_ptr(s), _ptr(d), _ptr(a), E, _ptr(x), N, _ptr(col16), _ptr(x8), _ptr(status), _stream()))
        assert int(status) == _lib.PYG_BAD_NODE_TYPE
        v[E // 3] = 1 << 31
        status.zero_()
        check(lib.gtok_pyg_narrow(_ptr(v), E, _ptr(out), _ptr(status), _stream()))
        assert int(status) == _lib.PYG_BAD_RANGE


@pytest.mark.parametrize("mode", ["collate", "ragged"])
@pytest.mark.parametrize("chunks", [1, 3, 8])
def test_run_pyg_equals_resident_path_and_oracle(dev, mode, chunks):
    from gtok import engine, synth
    from gtok.batch import PygStore
    from gtok.pipeline import ZincIbttPipeline
    G = 5000
    gb = synth.zinc_shaped(G, seed=21)
    vocab = _vocab()
    store = PygStore.from_batch(gb).pin_memory()
    assert store.edge_index.dtype == torch.int64 and store.edge_index.shape[0] == 2
    pipe = ZincIbttPipeline(gb, vocab, 1024, 128, dev, mode=mode)
    for _ in range(2):                                    # twice: staging reuse across calls
        r = pipe.run_pyg(store, chunks=chunks)
    ref, lens = engine.zinc_collate_epoch(gb.to(dev), vocab, 1024, 128)
    tot = int(ref.xoff[-1])
    assert int(r.host_xoff[-1]) == tot and torch.equal(r.host_lengths, lens.cpu())
    assert torch.equal(r.padded.X[:tot], ref.X[:tot]) and torch.equal(r.padded.attn[:tot], ref.attn[:tot])
    xo = r.host_xoff.tolist()
    lm = r.host_lmax.tolist()
    for g in range(0, G, 331):
        nn, s, d, x, a = gb.graph(g)
        want = O.zinc_ids(x.tolist(), s.tolist(), d.tolist(), a.tolist(), vocab, 1024, 1024).tolist()
        k, rrow = divmod(g, 128)
        row = r.padded.X[xo[k] + rrow * lm[k]: xo[k] + (rrow + 1) * lm[k]].cpu().tolist()
        assert row[:len(want)] == want and all(t == vocab["<pad>"] for t in row[len(want):])


def test_run_pyg_from_inmemory_dataset_storage_and_errors(dev):
    """PygStore.from_dataset takes an InMemoryDataset's storage untouched ([N,1] x, [E] edge_attr, dict slices)."""
    from gtok import synth
    from gtok.batch import PygStore
    from gtok.pipeline import ZincIbttPipeline
    gb = synth.zinc_shaped(640, seed=3)

    class _Data:
        pass

    class _DS:
        _indices = None
    ds = _DS()
    ds._data = _Data()
    ds._data.x = gb.x.to(torch.int64).reshape(-1, 1)
    ds._data.edge_index = torch.stack([gb.src.to(torch.int64), gb.dst.to(torch.int64)])
    ds._data.edge_attr = gb.edge_attr.to(torch.int64)
    ds.slices = {"x": gb.node_off.clone(), "edge_index": gb.edge_off.clone(), "edge_attr": gb.edge_off.clone()}
    store = PygStore.from_dataset(ds).pin_memory()
    vocab = _vocab()
    pipe = ZincIbttPipeline(gb, vocab, 1024, 128, dev)
    r = pipe.run_pyg(store)
    a = pipe.run(gb.pin_memory())
    tot = int(a.host_xoff[-1])
    assert int(r.host_xoff[-1]) == tot
    Xa = a.padded.X[:tot].clone()
    r = pipe.run_pyg(store)
    assert torch.equal(r.padded.X[:tot], Xa)
    bad = PygStore.from_batch(gb)
    bad.edge_attr[5] = 99
    with pytest.raises(ValueError, match="edge_attr outside"):
        pipe.run_pyg(bad.pin_memory())
    with pytest.raises(ValueError, match="built for shards"):
        pipe.run_pyg(PygStore.from_batch(synth.zinc_shaped(100, seed=1)).pin_memory())
