"""Drop-in surface on the GPU: the reference's own call sequences (trainer/train_ibtt.py:231-233,
:361-402; trainer/train_agtt.py:246-302, :514-542) executed against the gtok package, compared
with the golden vectors the reference produced and with the oracle."""
import json

import numpy as np
import pytest
import torch

from helpers import gt_examples, sha_int64
from oracle import ibtt_oracle as O
from oracle import sent_oracle as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


class _Mol:
    pass


def _fake_zinc(mols):
    out = []
    for m in mols:
        d = _Mol()
        d.x = torch.tensor(m["x"], dtype=torch.long).reshape(-1, 1)
        d.edge_index = torch.tensor([m["src"], m["dst"]], dtype=torch.long)
        d.edge_attr = torch.tensor(m["attr"], dtype=torch.long)
        d.y = torch.tensor([m["y"]], dtype=torch.float32)
        out.append(d)
    return out


@pytest.mark.parametrize("task,G,max_len,max_vocab", [("cycle_check", 1000, 600, 600), ("shortest_path", 200, 600, 60)])
def test_ibtt_trainer_sequence_graph_token(golden_gt, task, G, max_len, max_vocab):
    import graph_data_loader as g
    gold = golden_gt[f"{task}_G{G}_L{max_len}_V{max_vocab}"]
    _, texts, ex = gt_examples(task, G)
    vocab, _ = g.build_vocab_from_texts([e["text"] for e in ex], max_tokens=max_vocab)      # train_ibtt.py:391
    assert vocab == gold["vocab"]
    ds = g.TokenDataset(ex, vocab, max_len)                                                  # :395
    assert len(ds) == gold["num_kept"]
    assert sha_int64([s.numpy() for s in ds.seqs]) == gold["ids_sha256"]
    assert [int(l) for l in ds.labels] == gold["labels"]
    seq, lab = ds[3]
    assert seq.dtype == torch.long and lab.dtype == torch.long and not seq.is_cuda
    pad_id = vocab["<pad>"]
    for k, b0 in enumerate(range(0, len(ds), 128)):                                          # collate lambda :399
        X, A, Y = g.collate([ds[i] for i in range(b0, min(b0 + 128, len(ds)))], pad_id)
        assert X.is_cuda and A.dtype == torch.bool and Y.dtype == torch.long
        assert sha_int64([X.cpu().numpy()]) == gold["batch_sha"][k][1]
        assert sha_int64([Y.cpu().numpy()]) == gold["batch_sha"][k][3]
    # shuffled batch + foreign (non-registered) rows take the upload path
    idx = np.random.default_rng(0).permutation(len(ds))[:37]
    X, A, Y = g.collate([ds[int(i)] for i in idx], pad_id)
    Xw, Aw, _ = O.collate([ds.seqs[int(i)].numpy() for i in idx], [0] * 37, pad_id)
    assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
    X2, A2, _ = g.collate([(ds.seqs[int(i)].clone(), ds.labels[int(i)]) for i in idx], pad_id)
    assert torch.equal(X2, X) and torch.equal(A2, A)
    with pytest.raises(ValueError, match="no CPU fallback"):
        g.TokenDataset([{"text": "not a graph at all", "label": 1}], vocab, 64)


@pytest.mark.parametrize("max_len", [1024, 120])
def test_ibtt_trainer_sequence_zinc(golden_zinc, max_len):
    import graph_data_loader as g
    gold = golden_zinc[f"L{max_len}"]
    zds = g.ZINCTokenizationDataset(split="train", max_vocab=2000, max_len=max_len, zinc_dataset=_fake_zinc(golden_zinc["mols"]))
    ex = [zds[i] for i in range(len(zds))]                                                   # train_ibtt.py:231
    assert [e["text"] for e in ex[:3]] == gold["first_texts"]
    vocab, _ = g.build_fixed_zinc_vocab()
    dyn = []
    for e in ex:
        for t in e["text"].split():
            if t not in vocab and t not in dyn:
                dyn.append(t)
    vocab = g.extend_vocab_with_dynamic_tokens(vocab, dyn)                                   # :372
    assert vocab == gold["vocab"]
    ds = g.TokenDataset(ex, vocab, max_len)
    assert sha_int64([s.numpy() for s in ds.seqs]) == gold["ids_sha256"]
    assert [int(l) for l in ds.labels] == gold["labels"]                                     # int(label) truncation, C4
    X, A, Y = g.collate([ds[i] for i in range(8)], vocab["<pad>"])
    assert X.cpu().tolist() == gold["collate8"]["X"] and Y.cpu().tolist() == gold["collate8"]["Y"]


def test_agtt_tokenizer_shim_zinc(golden_zinc):
    import graph_data_loader as g
    from autograph.datamodules.data.tokenizer import Graph2TrailTokenizer
    mols = golden_zinc["mols"][:64]
    pyg = g.ZINCDatasetForAutoGraph(split="train", zinc_dataset=_fake_zinc(mols))
    nt, et = g.get_zinc_num_types()
    tok = Graph2TrailTokenizer(dataset_names=[], max_length=1024, truncation_length=1024, labeled_graph=True,
                               undirected=True)                                              # train_agtt.py:514-520
    nmax = max(pyg[i].num_nodes for i in range(len(pyg)))
    tok.set_num_nodes(nmax)
    tok.set_num_node_and_edge_types(nt, et)
    cfg = S.SentConfig(6, tok.node_idx_offset, tok.edge_idx_offset, 1024, True)
    first = []
    for i, m in enumerate(mols):
        t = tok(pyg[i])                                                                      # :250
        assert t.dtype == torch.long and t.dim() == 1 and t[0] == 0 and t[-1] == 4
        rp, col, ea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"])
        want = S.sent_tokenize(m["n"], rp, col, cfg, S.PhiloxWords(0, i), x=m["x"], eattr=ea)
        assert t.tolist() == want.tolist()
        first.append(t.tolist())
    again = [tok(pyg[i]).tolist() for i in range(len(mols))]                                 # next epoch: fresh trails
    assert again != first
    for i, m in enumerate(mols[:8]):
        rp, col, ea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"])
        assert again[i] == S.sent_tokenize(m["n"], rp, col, cfg, S.PhiloxWords(1, i), x=m["x"], eattr=ea).tolist()
    assert Graph2TrailTokenizer.pad == 5


def test_agtt_tokenizer_shim_graph_token(tmp_path):
    import graph_data_loader as g
    from autograph.datamodules.data.tokenizer import Graph2TrailTokenizer
    recs, texts, _ = gt_examples("shortest_path", 40, seed=9)
    d = tmp_path / "tasks_train" / "shortest_path" / "er" / "train"
    d.mkdir(parents=True)
    for i, t in enumerate(texts):
        (d / f"{i:03d}.json").write_text(json.dumps({"text": t}))
    ds = g.GraphTokenDatasetForAutoGraph(str(tmp_path), task="shortest_path", algorithm=["er"], split="train", seed=0)
    tok = Graph2TrailTokenizer(dataset_names=[], max_length=600, truncation_length=600, labeled_graph=False, undirected=True)
    tok.set_num_nodes(max(ds[i].num_nodes for i in range(len(ds))))
    cfg = S.SentConfig(idx_offset=6, max_length=600)
    kept = [r for r in recs if r["label"] != "INF"]
    assert len(ds) == len(kept)
    for i, r in enumerate(kept):
        data = ds[i]
        assert (data.query_u, data.query_v) == tuple(r["query"]) and data.num_nodes == r["n"]
        rp, col, _ = S.canonical_csr(r["n"], [e[0] for e in r["edges"]], [e[1] for e in r["edges"]])
        want = S.sent_tokenize(r["n"], rp, col, cfg, S.PhiloxWords(0, i))
        assert tok(data).tolist() == want.tolist()


def test_epoch_loaders_match_reference_collate():
    """gtok.loader: one epoch of GPU-resident batches == the reference's per-batch collate over the same row order
    (IBTT from graphs, IBTT from tokenized rows, AGTT with fresh trails per epoch), incl. shuffle, drop_last and a
    ragged last batch."""
    import numpy as np
    from gtok import engine
    from gtok.batch import GraphBatch
    from gtok.loader import AgttEpochLoader, IbttEpochLoader
    from helpers import gt_examples
    from oracle import ibtt_oracle as O
    from oracle import sent_oracle as S
    dev = torch.device("cuda:0")
    recs, texts, ex = gt_examples("cycle_check", 300, seed=4)
    vocab, _ = O.vocab_from_texts(texts, max_tokens=600)
    gb = GraphBatch.from_lists([r["n"] for r in recs], [r["edges"] for r in recs]).to(dev)
    labels = torch.tensor([e["label"] for e in ex])
    want = [np.asarray(O.encode_text(t, vocab, 600), dtype=np.int64) for t in texts]

    def check(loader, rows_of, pad, seed):
        for epoch in range(2):
            gen = torch.Generator(device="cpu")
            gen.manual_seed(seed)
            perms = [torch.randperm(300, generator=gen) for _ in range(epoch + 1)]
            order = perms[-1].tolist() if loader.shuffle else list(range(300))
            if loader.drop_last:
                order = order[:len(loader) * loader.B]
            got = list(loader)
            assert len(got) == len(loader)
            for k, (X, A, Y) in enumerate(got):
                idx = order[k * loader.B:(k + 1) * loader.B]
                Xw, Aw, _ = O.collate([rows_of(epoch, i) for i in idx], [0] * len(idx), pad)
                assert X.is_cuda and np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
                assert Y.cpu().tolist() == [int(labels[i]) for i in idx]

    for shuffle, drop_last, B in ((False, False, 128), (True, False, 64), (True, True, 128)):
        check(IbttEpochLoader.from_graphs(gb, vocab, "cycle_check", 600, labels, B, shuffle, seed=7, drop_last=drop_last),
              lambda ep, i: want[i], vocab["<pad>"], 7)
        rag = engine.tokenize_ibtt(gb, vocab, "cycle_check", 600)
        check(IbttEpochLoader.from_ragged(rag, vocab["<pad>"], labels, B, shuffle, seed=7, drop_last=drop_last),
              lambda ep, i: want[i], vocab["<pad>"], 7)
    # AGTT: trails differ per epoch, rows must equal the oracle's walk under the epoch's Philox seed
    csr = engine.build_csr(gb)
    cfg = engine.SentConfig(idx_offset=6, max_length=600)
    ocfg = S.SentConfig(idx_offset=6, max_length=600)
    csrs = [S.canonical_csr(r["n"], [e[0] for e in r["edges"]], [e[1] for e in r["edges"]]) for r in recs]
    cache = {}

    def sent_row(ep, i):
        if (ep, i) not in cache:
            seed = (11 * 0x9E3779B97F4A7C15 + ep) & (2 ** 64 - 1)
            rp, col, _ = csrs[i]
            cache[(ep, i)] = S.sent_tokenize(recs[i]["n"], rp, col, ocfg, S.PhiloxWords(seed, i))
        return cache[(ep, i)]
    check(AgttEpochLoader(csr, cfg, labels, 32, shuffle=True, seed=11), sent_row, 5, 11)


def test_epoch_loaders_query_positions():
    """with_q_pos: every batch carries the position extract_query_nodes would find by scanning the row on the host
    (train_ibtt.py:91-96 with vocab['<q>']; train_agtt.py:101-112,131 with idx_offset + num_nodes of the batch's
    first graph - which can collide with a node-position token of a larger graph, quirk C9, reproduced)."""
    from gtok import engine
    from gtok.batch import GraphBatch
    from gtok.loader import AgttEpochLoader, IbttEpochLoader
    from helpers import gt_examples
    from oracle import ibtt_oracle as O
    dev = torch.device("cuda:0")
    recs, texts, ex = gt_examples("shortest_path", 200, seed=6)
    vocab, _ = O.vocab_from_texts(texts, max_tokens=600)
    gb = GraphBatch.from_lists([r["n"] for r in recs], [r["edges"] for r in recs], queries=[r["query"] for r in recs]).to(dev)
    for max_len in (600, 120):
        n_rows = 0
        for X, A, Y, qp in IbttEpochLoader.from_graphs(gb, vocab, "shortest_path", max_len, None, 32, shuffle=True, seed=1, with_q_pos=True):
            assert qp.dtype == torch.int32 and qp.numel() == X.shape[0]
            for b in range(X.shape[0]):
                nz = (X[b] == vocab["<q>"]).nonzero(as_tuple=True)[0]
                assert int(qp[b]) == (int(nz[0]) if len(nz) else -1)
            n_rows += X.shape[0]
        assert n_rows == 200
    csr = engine.build_csr(gb)
    cfg = engine.SentConfig(idx_offset=6, max_length=600, append_query=True)
    loader = AgttEpochLoader(csr, cfg, None, 16, shuffle=True, seed=3, with_q_pos=True)
    gen = torch.Generator(device="cpu")
    gen.manual_seed(3)
    order = torch.randperm(200, generator=gen).tolist()
    found = 0
    for k, (X, A, Y, qp) in enumerate(loader):
        q_id = 6 + recs[order[k * 16]]["n"]                      # data_list[0].num_nodes
        for b in range(X.shape[0]):
            nz = (X[b] == q_id).nonzero(as_tuple=True)[0]
            assert int(qp[b]) == (int(nz[0]) if len(nz) else -1)
            found += len(nz) > 0
    assert found > 0


@pytest.mark.parametrize("max_len", [600, 97, 64])
def test_token_dataset_keeps_label_tokens_when_asked(golden_zinc, max_len):
    """TokenDataset(strip_label=False) (reference data_loader.py:478-482 without the strip): the whole text, label and <eos>
    included, cut at max_len - graph-token texts and ZINC texts (truncated items keep their own tail)."""
    import graph_data_loader as g
    for task in ("cycle_check", "shortest_path"):
        _, texts, ex = gt_examples(task, 200)
        vocab, _ = g.build_vocab_from_texts(texts, max_tokens=600)
        ds = g.TokenDataset(ex, vocab, max_len, strip_label=False)
        kept = [e for e in ex if e["label"] is not None]
        assert len(ds) == len(kept)
        for s, e in zip(ds.seqs, kept):
            assert s.tolist() == O.encode_text(e["text"], vocab, max_len, strip_label=False)
    zds = g.ZINCTokenizationDataset(split="train", max_vocab=2000, max_len=max_len, zinc_dataset=_fake_zinc(golden_zinc["mols"]))
    ex = [zds[i] for i in range(len(zds))]
    vocab = golden_zinc["L1024"]["vocab"]
    ds = g.TokenDataset(ex, vocab, max_len, strip_label=False)
    for s, e in zip(ds.seqs, ex):
        assert s.tolist() == O.encode_text(e["text"], vocab, max_len, strip_label=False)
    X, A, Y = g.collate([ds[i] for i in range(16)], vocab["<pad>"])
    Xw, Aw, _ = O.collate([s.numpy() for s in ds.seqs[:16]], [0] * 16, vocab["<pad>"])
    assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)


def test_token_dataset_registry_does_not_keep_datasets_alive():
    import gc
    import weakref
    import graph_data_loader as g
    from graph_data_loader import data_loader as dl
    _, texts, ex = gt_examples("cycle_check", 50)
    vocab, _ = g.build_vocab_from_texts(texts, max_tokens=600)
    ds = g.TokenDataset(ex, vocab, 600)
    key = ds._host.data_ptr()
    assert dl._REGISTRY[key] is ds
    ref = weakref.ref(ds)
    del ds
    gc.collect()
    assert ref() is None and key not in dl._REGISTRY


def test_sent_shim_epoch_rule_and_worker_guard(golden_zinc):
    """an early re-read of an item is served from the current epoch's trails; a re-read after half of the dataset has been
    served starts a new epoch; new_epoch() forces one; DataLoader workers are refused with a clear message."""
    import graph_data_loader as g
    from autograph.datamodules.data.tokenizer import Graph2TrailTokenizer
    zds = g.ZINCDatasetForAutoGraph(split="train", zinc_dataset=_fake_zinc(golden_zinc["mols"][:64]))
    tok = Graph2TrailTokenizer(dataset_names=[], max_length=512, truncation_length=512, labeled_graph=True, undirected=True)
    tok.set_num_nodes(40)
    tok.set_num_node_and_edge_types(9, 4)
    a0 = tok(zds[0]).tolist()
    assert tok(zds[0]).tolist() == a0                        # peek: same epoch, same trail
    first = [tok(zds[i]).tolist() for i in range(40)]        # more than half served
    assert first[0] == a0
    b0 = tok(zds[0]).tolist()                                # served before and > half done: new epoch, fresh trails
    again = [tok(zds[i]).tolist() for i in range(1, 40)]
    assert any(x != y for x, y in zip(first[1:], again)) or b0 != a0
    tok.new_epoch()
    c = [tok(zds[i]).tolist() for i in range(40)]
    assert any(x != y for x, y in zip([b0] + again, c))
    import torch.utils.data as tud
    real = tud.get_worker_info
    tud.get_worker_info = lambda: object()
    try:
        with pytest.raises(RuntimeError, match="num_workers: 0"):
            tok(zds[0])
    finally:
        tud.get_worker_info = real


@pytest.mark.parametrize("max_len", [1024, 120, 64])
def test_zinc_items_rendered_on_the_gpu_match_host_rendering(golden_zinc, max_len):
    """every item's text comes from one GPU-rendered byte buffer and equals the per-item host rendering (the
    restatement of tokenize_molecule + __getitem__ truncation, pinned to the reference's own texts in the golden file);
    dynamic_tokens() reproduces the first-seen dynamic vocabulary of the golden vocab without reading any text."""
    import graph_data_loader as g
    mols = _fake_zinc(golden_zinc["mols"])
    zds = g.ZINCTokenizationDataset(split="train", max_vocab=2000, max_len=max_len, zinc_dataset=mols)
    ex = [zds[i] for i in range(len(zds))]
    gold = golden_zinc.get(f"L{max_len}")
    if gold is not None:
        assert [e["text"] for e in ex[:3]] == gold["first_texts"]
    for i, e in enumerate(ex):
        d = mols[i]
        lab = d.y.item()
        toks = zds.tokenize_molecule(d, lab).split()
        if len(toks) > max_len:
            toks = toks[:max_len - 1] + ["<eos>"]
        assert e["text"] == " ".join(toks) and e["label"] == lab, i
    if max_len == 1024:
        fixed, _ = g.build_fixed_zinc_vocab()
        vocab = g.extend_vocab_with_dynamic_tokens(fixed, zds.dynamic_tokens())
        assert vocab == gold["vocab"]
        # and it is what the trainer's own loop collects (train_ibtt.py:364-369), in first-seen order
        dyn = []
        for e in ex:
            for t in e["text"].split():
                if t not in fixed and t not in dyn:
                    dyn.append(t)
        assert zds.dynamic_tokens() == dyn


def test_zinc_bulk_text_rendering_speed():
    """[ds[i] for i in range(len(ds))] over 100k molecules: well under a second per 100k once the buffer exists."""
    import time
    import graph_data_loader as g
    from gtok import synth
    gb = synth.zinc_shaped(100_000, seed=5)

    class _DS:
        _indices = None

    class _D:
        pass
    ds = _DS()
    ds._data = _D()
    ds._data.x = gb.x.to(torch.int64).reshape(-1, 1)
    ds._data.edge_index = torch.stack([gb.src.to(torch.int64), gb.dst.to(torch.int64)])
    ds._data.edge_attr = gb.edge_attr.to(torch.int64)
    ds._data.y = gb.y
    ds.slices = {"x": gb.node_off, "edge_index": gb.edge_off}
    ds.__class__.__len__ = lambda self: 100_000
    zds = g.ZINCTokenizationDataset(split="train", max_len=1024, zinc_dataset=ds)
    t0 = time.perf_counter()
    ex = [zds[i] for i in range(len(zds))]
    dt = time.perf_counter() - t0
    assert len(ex) == 100_000 and ex[7]["text"].startswith("<bos> <atom> ") and ex[7]["text"].endswith(" <eos>")
    assert dt < 5.0, dt                                    # (reference: ~18 s per 100k molecules on one core)
    print(f"100k ZINC items in {dt:.2f} s")


def test_ibtt_dataloader_with_workers_as_in_the_reference_yaml(golden_gt):
    """configs/ibtt_graph_token.yaml:27 / ibtt_zinc.yaml:23 set num_workers: 2 and the trainer builds its loaders with it
    (trainer/train_ibtt.py:399-402).  Workers cannot touch CUDA: collate() hands back row handles there and K4 pads the batch
    at the trainer's `.to(device)` (:144).  Batches must equal the num_workers=0 ones (and the reference's golden sha)."""
    import graph_data_loader as g
    from torch.utils.data import DataLoader
    gold = golden_gt["cycle_check_G1000_L600_V600"]
    _, texts, ex = gt_examples("cycle_check", 1000)
    vocab, _ = g.build_vocab_from_texts([e["text"] for e in ex], max_tokens=600)
    ds = g.TokenDataset(ex, vocab, 600)
    pad_id = vocab["<pad>"]
    coll = lambda b: g.collate(b, pad_id)                                                    # train_ibtt.py:399
    device = torch.device("cuda")
    dl = DataLoader(ds, batch_size=128, shuffle=False, num_workers=2, collate_fn=coll)       # :401
    seen = 0
    for k, (X, A, Y) in enumerate(dl):
        X, A, Y = X.to(device), A.to(device), Y.to(device)                                   # :144
        assert X.is_cuda and X.dtype == torch.long and A.dtype == torch.bool and Y.dtype == torch.long
        assert sha_int64([X.cpu().numpy()]) == gold["batch_sha"][k][1]
        assert sha_int64([Y.cpu().numpy()]) == gold["batch_sha"][k][3]
        seen += X.shape[0]
    assert seen == len(ds)
    # shuffled epochs deliver every row exactly once
    dl = DataLoader(ds, batch_size=64, shuffle=True, num_workers=2, collate_fn=coll)
    total = sum(int(A.to(device).sum()) for X, A, Y in dl)
    assert total == sum(int(s.numel()) for s in ds.seqs)
