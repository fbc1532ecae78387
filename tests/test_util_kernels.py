"""Plumbing kernels (csrc/gtok_util.cu) against plain tensor expressions: compact rows, gather graphs, balance order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("G,stride", [(0, 8), (1, 1), (1000, 37), (50000, 64)])
def test_compact_rows(dev, G, stride):
    from gtok import engine
    g = torch.Generator().manual_seed(G)
    tok = torch.randint(0, 1 << 40, (G, stride), generator=g, dtype=torch.int64)
    ln = torch.randint(0, stride + 1, (G,), generator=g, dtype=torch.int32)
    flat, off = engine.compact_rows(tok.to(dev), ln.to(dev))
    keep = torch.arange(stride)[None, :] < ln[:, None]
    assert torch.equal(flat.cpu(), tok[keep])
    want_off = torch.zeros(G + 1, dtype=torch.int64)
    want_off[1:] = ln.to(torch.int64).cumsum(0)
    assert torch.equal(off.cpu(), want_off)


def test_gather_graphs_matches_host_select(dev):
    from gtok import synth
    gb = synth.zinc_shaped(3000, seed=2)
    gb.query_u = torch.arange(3000, dtype=torch.int32)
    gb.query_v = torch.arange(3000, dtype=torch.int32) * 2
    rng = np.random.default_rng(0)
    for idx in (torch.from_numpy(rng.permutation(3000)), torch.from_numpy(rng.integers(0, 3000, 777)),
                torch.zeros(0, dtype=torch.int64), torch.tensor([2999, 0, 0, 5])):
        want = gb.select(idx)                                  # host tensor ops
        got = gb.to(dev).select(idx.to(dev))                   # library kernels
        for f in ("node_off", "edge_off", "src", "dst", "x", "edge_attr", "query_u", "query_v", "y"):
            assert torch.equal(getattr(got, f).cpu(), getattr(want, f)), f
    with pytest.raises(IndexError):
        gb.to(dev).select(torch.tensor([3000], device=dev))


@pytest.mark.parametrize("window", [2, 100, 2048, 4096])
def test_balance_order_is_a_windowed_stable_sort(dev, window):
    from gtok import engine, synth
    G = 10000
    gb = synth.er_graph_token(G, seed=3, device=dev)
    csr = engine.build_csr(gb, balance_window=window)
    order = csr.order.cpu().to(torch.int64)
    n = gb.num_nodes().cpu()
    nnz = (csr.rowptr[gb.node_off[1:]] - csr.rowptr[gb.node_off[:-1]]).cpu().to(torch.int64)
    w = torch.arange(G) // window
    want = torch.argsort(w * (1 << 40) + n * (1 << 20) + nnz, stable=True)
    assert torch.equal(order, want)
    assert sorted(order.tolist()) == list(range(G))
