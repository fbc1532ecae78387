"""GPU parity tests: CUDA kernels (through the C-ABI, via gtok.engine) vs the oracle
and vs the golden fixtures made from the reference itself.  Bit-exact everywhere
(integer / index work).  Run with: pytest -m gpu
"""
import hashlib

import numpy as np
import pytest
import torch

from helpers import gt_examples, sha_int64
from oracle import ibtt_oracle as O
from oracle import sent_oracle as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _gb_from_recs(recs, with_query):
    from gtok.batch import GraphBatch
    return GraphBatch.from_lists([r["n"] for r in recs], [r["edges"] for r in recs],
                                 queries=[r["query"] for r in recs] if with_query else None)


# --------------------------------------------------------------------------- K0
@pytest.mark.parametrize("n", [0, 1, 5, 2047, 2048, 2049, 100003, 1 << 20, 3_000_017])
def test_scan_offsets(dev, n):
    from gtok import engine
    g = torch.Generator().manual_seed(n)
    lens = torch.randint(0, 1025, (n,), generator=g, dtype=torch.int32)
    off = engine.scan_offsets(lens.to(dev)).cpu()
    want = torch.zeros(n + 1, dtype=torch.int64)
    want[1:] = lens.to(torch.int64).cumsum(0)
    assert torch.equal(off, want)


def test_scan_offsets_beyond_int32(dev):
    from gtok import engine
    n = 3_000_000
    lens = torch.full((n,), 1024, dtype=torch.int32, device=dev)
    off = engine.scan_offsets(lens)
    assert int(off[-1]) == n * 1024 > 2 ** 31
    assert int(off[n // 2]) == (n // 2) * 1024


# --------------------------------------------------------------------------- K1
GT_CASES = [("cycle_check", 1000, 600, 600), ("shortest_path", 200, 600, 60), ("cycle_check", 64, 97, 30)]


@pytest.mark.parametrize("task,G,max_len,max_vocab", GT_CASES)
def test_ibtt_graph_token_matches_reference_golden(dev, golden_gt, task, G, max_len, max_vocab):
    from gtok import engine
    gold = golden_gt[f"{task}_G{G}_L{max_len}_V{max_vocab}"]
    vocab = gold["vocab"]
    recs, texts, ex = gt_examples(task, G)
    kept = [r for r, e in zip(recs, ex) if e["label"] is not None]       # C3 filter happens on the host
    gb = _gb_from_recs(kept, task == "shortest_path").to(dev)
    rag = engine.tokenize_ibtt(gb, vocab, task, max_len)
    rows = rag.rows_cpu()
    assert [len(r) for r in rows] == gold["lengths"]
    assert sha_int64([r.numpy() for r in rows]) == gold["ids_sha256"]
    assert [r.tolist() for r in rows[:24]] == gold["first_seqs"]
    # K4 on every batch of 128, against the reference's own collate output
    ep = engine.pad_mask_epoch(rag.tokens, rag.offsets, rag.lengths, vocab["<pad>"], 128)
    assert ep.num_batches == len(gold["batch_sha"])
    for k, (shp, xs, as_, _) in enumerate(gold["batch_sha"]):
        X, A = ep.batch(k)
        assert list(X.shape) == shp and A.dtype == torch.bool
        assert sha_int64([X.cpu().numpy()]) == xs
        assert hashlib.sha256(A.cpu().numpy().tobytes()).hexdigest() == as_
    X, A = engine.pad_mask(rag.tokens, rag.offsets, rag.lengths, vocab["<pad>"],
                           batch_idx=torch.arange(8, device=dev))
    assert X.cpu().tolist() == gold["collate8"]["X"] and A.cpu().to(torch.int64).tolist() == gold["collate8"]["attn"]


def test_ibtt_random_shapes_vs_oracle(dev):
    """empty edge lists, single nodes, OOV node ids (vocab cap), odd/even row starts, both tasks."""
    from gtok import engine
    from gtok.batch import GraphBatch
    rng = np.random.default_rng(5)
    for task in ("cycle_check", "shortest_path"):
        recs = []
        for i in range(300):
            n = int(rng.integers(1, 70))
            m = int(rng.integers(0, 3 * n)) if i % 7 else 0
            edges = [(int(rng.integers(0, n)), int(rng.integers(0, n))) for _ in range(m)]
            recs.append({"n": n, "edges": edges, "query": (int(rng.integers(0, n)), int(rng.integers(0, n)))})
        vocab = {t: i for i, t in enumerate(O.SPECIAL)}
        for k in rng.permutation(40):                      # only 40 of up to 70 node ids known, shuffled ids
            vocab[str(int(k))] = len(vocab)
        vocab["has_cycle"] = len(vocab)
        vocab["shortest_distance"] = len(vocab)
        for max_len in (600, 33, 1):
            gb = GraphBatch.from_lists([r["n"] for r in recs], [r["edges"] for r in recs],
                                       queries=[r["query"] for r in recs]).to(dev)
            rows = engine.tokenize_ibtt(gb, vocab, task, max_len).rows_cpu()
            for r, got in zip(recs, rows):
                src = np.asarray([a for a, _ in r["edges"]], dtype=np.int32)
                dst = np.asarray([b for _, b in r["edges"]], dtype=np.int32)
                want = O.ibtt_ids(r["n"], src, dst, task, vocab, max_len, r["query"])
                assert got.tolist() == want.tolist()


@pytest.mark.parametrize("task", ["cycle_check", "shortest_path"])
def test_ibtt_fused_collate_vs_oracle(dev, task):
    """graphs -> padded batches in one kernel == reference-style encode + collate, in dataset order and under a
    shuffled epoch order (batch_idx), with truncation, OOV node ids and a vocabulary beyond 16 bits."""
    from gtok import engine
    recs, texts, ex = gt_examples(task, 700, seed=5)
    vocab, _ = O.vocab_from_texts(texts, max_tokens=40)          # small cap: many node ids fall out of the vocab
    wide = {t: (i if i < 9 else i + 70000) for t, i in vocab.items()}
    gb = _gb_from_recs(recs, task == "shortest_path").to(dev)
    rng = np.random.default_rng(3)
    perm = torch.from_numpy(rng.permutation(len(recs))[:613]).to(dev)
    for voc, max_len, B, idx in ((vocab, 600, 128, None), (vocab, 100, 37, perm), (wide, 600, 64, perm), (vocab, 3000, 1, None)):
        want = [np.asarray(O.encode_text(t, voc, max_len), dtype=np.int64) for t in texts]
        ep, lens = engine.ibtt_collate_epoch(gb, voc, task, max_len, B, batch_idx=idx)
        assert lens.cpu().tolist() == [len(w_) for w_ in want]
        order = list(range(len(recs))) if idx is None else idx.cpu().tolist()
        for k in range(ep.num_batches):
            rows = [want[i] for i in order[k * B:(k + 1) * B]]
            Xw, Aw, _ = O.collate(rows, [0] * len(rows), voc["<pad>"])
            X, A = ep.batch(k)
            assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw), (max_len, B, k)


def test_ibtt_fused_collate_equals_two_step_large(dev):
    """200k ER graphs + three 4096-node graphs (rows longer than the shared-memory row buffer: closed-form path):
    byte-identical to gtok_ibtt_serialize followed by gtok_pad_mask_batched."""
    from gtok import engine, synth
    from gtok.batch import GraphBatch
    vocab = {t: i for i, t in enumerate(["<pad>", "<bos>", "<e>", "<n>", "<q>", "<p>", "<eos>", "yes", "no"])}
    for k in range(4096):
        vocab[str(k)] = len(vocab)
    vocab["has_cycle"] = len(vocab)
    gb = synth.er_graph_token(200_000, seed=2, device=dev)
    rag = engine.tokenize_ibtt(gb, vocab, "cycle_check", 600)
    ep2 = engine.pad_mask_epoch(rag.tokens, rag.offsets, rag.lengths, vocab["<pad>"], 128)
    ep1, lens = engine.ibtt_collate_epoch(gb, vocab, "cycle_check", 600, 128)
    assert torch.equal(lens, rag.lengths) and torch.equal(ep1.xoff, ep2.xoff) and torch.equal(ep1.lmax, ep2.lmax)
    assert torch.equal(ep1.X, ep2.X) and torch.equal(ep1.attn, ep2.attn)
    rng = np.random.default_rng(0)
    big = GraphBatch.from_lists([4096, 17, 4096, 4096], [list(map(tuple, rng.integers(0, 4096, (m, 2)).tolist())) if n > 17 else
                                                        [(0, 1), (3, 16)] for n, m in ((4096, 9000), (17, 2), (4096, 1), (4096, 16000))]).to(dev)
    rag = engine.tokenize_ibtt(big, vocab, "cycle_check", 100000)
    ep2 = engine.pad_mask_epoch(rag.tokens, rag.offsets, rag.lengths, vocab["<pad>"], 2)
    ep1, lens = engine.ibtt_collate_epoch(big, vocab, "cycle_check", 100000, 2)
    assert torch.equal(lens, rag.lengths) and torch.equal(ep1.xoff, ep2.xoff)
    assert torch.equal(ep1.X, ep2.X) and torch.equal(ep1.attn, ep2.attn)


@pytest.mark.parametrize("task", ["cycle_check", "shortest_path"])
def test_vocab_from_graphs_matches_reference_order(dev, golden_gt, task):
    """K5: the vocabulary built from the graphs (GPU histogram + first-occurrence keys) is the dict
    build_vocab_from_texts returns on their texts - same ids, same tie-breaks, same max_tokens / min_freq cuts."""
    from gtok import engine
    for G, seed in ((1000, 0), (300, 7)):
        recs, texts, ex = gt_examples(task, G, seed=seed)
        gb = _gb_from_recs(recs, task == "shortest_path").to(dev)
        labs = [t.split()[t.split().index("<p>") + 1] for t in texts]
        for max_tokens, min_freq in ((None, 1), (600, 1), (40, 1), (25, 1), (None, 900), (2000, 50)):
            want, witos = O.vocab_from_texts(texts, min_freq=min_freq, max_tokens=max_tokens)
            got, gitos = engine.vocab_from_graphs(gb, task, labs, min_freq=min_freq, max_tokens=max_tokens)
            assert got == want and gitos == witos, (G, max_tokens, min_freq)
    for t_, G, max_len, max_vocab in GT_CASES:                    # the reference's own build_vocab_from_texts output
        if t_ != task:
            continue
        gold = golden_gt[f"{t_}_G{G}_L{max_len}_V{max_vocab}"]
        recs, texts, ex = gt_examples(task, G)
        gb = _gb_from_recs(recs, task == "shortest_path").to(dev)
        labs = [t.split()[t.split().index("<p>") + 1] for t in texts]
        got, _ = engine.vocab_from_graphs(gb, task, labs, max_tokens=max_vocab)
        assert got == gold["vocab"]


def test_ibtt_empty_batch(dev):
    from gtok import engine
    from gtok.batch import GraphBatch
    gb = GraphBatch.from_lists([], []).to(dev)
    rag = engine.tokenize_ibtt(gb, {t: i for i, t in enumerate(O.SPECIAL)}, "cycle_check", 64, max_nodes=1)
    assert rag.tokens.numel() == 0 and rag.offsets.cpu().tolist() == [0]


@pytest.mark.parametrize("task", ["cycle_check", "shortest_path"])
def test_text_ingest_matches_host_parser(dev, task):
    """graph-token texts -> graph arrays on the GPU == the host inverse parser (graph_of_text, itself checked
    against the reference's parse_graph_from_text in the host tests), incl. odd whitespace; every malformed text
    is rejected with ValueError like the host path."""
    from gtok import engine
    from graph_data_loader.data_loader import graph_of_text
    recs, texts, ex = gt_examples(task, 500, seed=13)
    texts = list(texts)
    texts[3] = "  " + texts[3].replace(" ", "  ") + " \n"            # runs of blanks, leading / trailing whitespace
    texts[4] = texts[4].replace(" ", "\t", 5)
    gb, got_task = engine.graphs_from_texts(texts, dev)
    assert got_task == task and gb.num_graphs == len(texts)
    cpu = gb.to("cpu")
    for g, t in enumerate(texts):
        n, edges, tk, q = graph_of_text(t)
        nn, s_, d_, _, _ = cpu.graph(g)
        assert nn == n and list(zip(s_.tolist(), d_.tolist())) == edges and tk == task
        if task == "shortest_path":
            assert (int(cpu.query_u[g]), int(cpu.query_v[g])) == q
    good = texts[0]
    toks = good.split()
    bad_cases = [
        " ".join(toks[1:]),                                           # no <bos>
        good.replace("<n>", "<m>", 1),                                # no <n>
        good.replace("<p>", "<x>", 1),                                # no <p>
        " ".join(toks[:1] + ["x7"] + toks[2:]),                       # edge endpoint is not an integer
        " ".join(toks[:3] + ["<f>"] + toks[4:]) if toks[3] == "<e>" else good.replace("<e>", "<f>", 1),
        good.replace("<n> 0 1", "<n> 0 2", 1),                        # node list is not 0 1 2 ..
        good.replace("has_cycle", "has_cycles").replace("shortest_distance", "shortest_path"),
        "",
        good + " \u00e9",
    ]
    for b_ in bad_cases:
        assert graph_of_text(b_) is None if b_.isascii() else True
        with pytest.raises(ValueError):
            engine.graphs_from_texts([good, b_, good], dev)
    other = gt_examples("shortest_path" if task == "cycle_check" else "cycle_check", 4, seed=1)[1]
    with pytest.raises(ValueError):
        engine.graphs_from_texts([good, other[0]], dev)               # mixed tasks


def test_fused_entry_points_edge_cases(dev):
    """empty inputs, ragged last batches, graphs without edges / single nodes, and the error codes of the fused
    tokenize+collate entry points (no silent fallback: unsupported shapes raise)."""
    import ctypes as C
    from gtok import engine, _lib
    from gtok._lib import lib, GtokError
    from gtok.batch import GraphBatch
    from gtok.engine import _ptr, _stream
    gvoc = {t: i for i, t in enumerate(O.SPECIAL)}
    gvoc.update({str(k): 9 + k for k in range(8)})
    gvoc["has_cycle"] = 17
    zvoc = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in range(8)] + ["X", "unknown"])
    empty = GraphBatch.from_lists([], [], x_lists=[], attr_lists=[]).to(dev)
    ep, lens = engine.ibtt_collate_epoch(empty, gvoc, "cycle_check", 64, 128, max_nodes=1)
    assert ep.num_batches == 0 and lens.numel() == 0 and ep.X.numel() == 0
    ep, lens = engine.zinc_collate_epoch(empty, zvoc, 64, 128, max_nodes=1, max_entries=1)
    assert ep.num_batches == 0 and lens.numel() == 0 and int(ep.xoff[-1]) == 0
    # 33 tiny graphs: no edges, single nodes, a 2-cycle listed twice; batch 32 -> ragged last batch of one row
    nn = [1, 2, 3] * 11
    el = [[], [(0, 1), (1, 0), (0, 1)], [(2, 0)]] * 11
    gb = GraphBatch.from_lists(nn, el, x_lists=[[0] * n for n in nn], attr_lists=[[1] * len(e) for e in el]).to(dev)
    ep, lens = engine.zinc_collate_epoch(gb, zvoc, 64, 32)
    want = [O.zinc_ids([0] * n, [a for a, _ in e], [b for _, b in e], [1] * len(e), zvoc, 64, 64) for n, e in zip(nn, el)]
    assert ep.num_batches == 2 and lens.cpu().tolist() == [len(w_) for w_ in want]
    for k in range(2):
        Xw, Aw, _ = O.collate(want[k * 32:(k + 1) * 32], [0] * len(want[k * 32:(k + 1) * 32]), zvoc["<pad>"])
        X, A = ep.batch(k)
        assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
    ep, lens = engine.ibtt_collate_epoch(gb, gvoc, "cycle_check", 64, 32)
    wantg = [np.asarray(O.ibtt_ids(n, np.asarray([a for a, _ in e], dtype=np.int64), np.asarray([b for _, b in e], dtype=np.int64),
                                   "cycle_check", gvoc, 64)) for n, e in zip(nn, el)]
    for k in range(2):
        Xw, Aw, _ = O.collate(wantg[k * 32:(k + 1) * 32], [0] * len(wantg[k * 32:(k + 1) * 32]), 0)
        X, A = ep.batch(k)
        assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
    # bounds that are too small for the data must not corrupt memory: the affected warps emit header-only rows
    mols = [{"n": 20, "src": list(range(19)) * 2, "dst": list(range(1, 20)) * 2, "x": [0] * 20, "attr": [1] * 38}] * 70
    gbb = _zinc_batch(mols, dev)
    ep, lens = engine.zinc_collate_epoch(gbb, zvoc, 1024, 32, max_nodes=20, max_entries=2)
    torch.cuda.synchronize()
    assert lens.cpu().tolist() == [4] * 70 and ep.batch(0)[0][0].cpu().tolist() == [zvoc["<bos>"], zvoc["<q>"], zvoc["regression"], zvoc["<p>"]]
    rag = engine.tokenize_zinc(gbb, zvoc, 1024, max_nodes=20, max_entries=2)
    assert rag.lengths.cpu().tolist() == [4] * 70
    # the single-pass kernel reports shapes it does not cover (GTOK_EUNSUPPORTED, never a silent fallback inside the library) ...
    with pytest.raises(GtokError) as ei:
        bc48 = engine._batch_c(gb, 3, 3)
        zv48, _t48 = engine.zinc_vocab_c(zvoc, 64, 64, 3, dev)
        lm48, xo48 = torch.empty(1, dtype=torch.int32, device=dev), torch.zeros(2, dtype=torch.int64, device=dev)
        X48, A48 = torch.empty(4096, dtype=torch.int64, device=dev), torch.empty(4096, dtype=torch.bool, device=dev)
        ws48 = torch.empty(max(16, lib.gtok_zinc_collate_workspace_bytes(33, 48)), dtype=torch.uint8, device=dev)
        engine.check(lib.gtok_zinc_collate(C.byref(bc48), C.byref(zv48), 48, 2, _ptr(lens), _ptr(lm48), _ptr(xo48), _ptr(X48), _ptr(A48),
                                           4096, _ptr(ws48), ws48.numel(), _stream()))
    assert ei.value.code == _lib.GTOK_EUNSUPPORTED
    # ... and the engine routes them through lengths -> extents -> pad cells / masks / row offsets -> tokens at their padded
    # position (gtok_pad_tail_batched + gtok_zinc_serialize_rows): any batch size, any molecule size the tokenizer takes
    for Bx, kw in ((48, {}), (1, {}), (7, {}), (500, {}), (32, {"max_nodes": 100})):
        ep, lens = engine.zinc_collate_epoch(gb, zvoc, 64, Bx, **kw)
        assert lens.cpu().tolist() == [len(w_) for w_ in want] and ep.num_batches == (33 + Bx - 1) // Bx
        for k in range(ep.num_batches):
            Xw, Aw, _ = O.collate(want[k * Bx:(k + 1) * Bx], [0] * len(want[k * Bx:(k + 1) * Bx]), zvoc["<pad>"])
            X, A = ep.batch(k)
            assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw), (Bx, k)
    with pytest.raises(GtokError) as ei:                                   # null pointers are EINVAL, not a crash
        bc = engine._batch_c(gb, 3, 3)
        zv, _t = engine.zinc_vocab_c(zvoc, 64, 64, 3, dev)
        engine.check(lib.gtok_zinc_collate(C.byref(bc), C.byref(zv), 32, 2, None, None, _ptr(lens), None, None, 0, None, 0, _stream()))
    assert ei.value.code == _lib.GTOK_EINVAL


def test_ibtt_roundtrip_property_large(dev):
    """Size-independent property at scale: detokenised rows parse back (reference inverse
    parser grammar, graph_token_dataset_autograph.py:14-54) to the input edge lists, and
    the total length equals the closed form."""
    from gtok import engine, synth
    G = 200_000
    gb = synth.er_graph_token(G, seed=11, device=dev)
    vocab = {t: i for i, t in enumerate(O.SPECIAL)}
    for k in range(49):
        vocab[str(k)] = len(vocab)
    vocab["has_cycle"] = len(vocab)
    itos = {i: t for t, i in vocab.items()}
    rag = engine.tokenize_ibtt(gb, vocab, "cycle_check", 600)
    n, e = gb.num_nodes(), gb.num_entries()
    want_len = torch.clamp(3 * e + n + 5, max=600).to(torch.int32)
    assert torch.equal(rag.lengths, want_len)
    assert int(rag.offsets[-1]) == int(want_len.sum())
    cpu = gb.to("cpu")
    for g in list(range(0, G, G // 50)):
        row = rag.row(g).cpu().tolist()
        nn, s, d, _, _ = cpu.graph(g)
        if len(row) < 600:
            nodes, edges = O.parse_graph_text(" ".join(itos[t] for t in row))
            assert nodes == list(range(nn)) and edges == list(zip(s.tolist(), d.tolist()))


# --------------------------------------------------------------------------- K2
def _zinc_batch(mols, dev):
    from gtok.batch import GraphBatch
    return GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                                 x_lists=[m["x"] for m in mols], attr_lists=[m["attr"] for m in mols]).to(dev)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("max_len", [1024, 120, 121, 122, 64])
def test_zinc_matches_reference_golden(dev, golden_zinc, max_len, fused):
    """fused = single-pass tile kernel; not fused = lengths -> scan -> serialize (bitmap path)."""
    from gtok import engine
    gold = golden_zinc[f"L{max_len}"]
    vocab = gold["vocab"]
    gb = _zinc_batch(golden_zinc["mols"], dev)
    rag = engine.tokenize_zinc(gb, vocab, max_len, fused=fused)
    rows = rag.rows_cpu()
    assert [len(r) for r in rows] == gold["lengths"]
    assert sha_int64([r.numpy() for r in rows]) == gold["ids_sha256"]
    assert [r.tolist() for r in rows[:32]] == gold["first_seqs"]
    X, A = engine.pad_mask(rag.tokens, rag.offsets, rag.lengths, vocab["<pad>"], batch_idx=torch.arange(8, device=dev))
    assert X.cpu().tolist() == gold["collate8"]["X"] and A.cpu().to(torch.int64).tolist() == gold["collate8"]["attn"]


def test_zinc_adversarial_vs_oracle(dev):
    """duplicates, self loops, reversed-first orientation, no edges, single atom, OOV atoms/bonds,
    ids missing from the vocab, many columns (second kernel variant)."""
    from gtok import engine
    rng = np.random.default_rng(9)
    vocab = O.zinc_fixed_vocab()
    vocab = O.zinc_extend_vocab(vocab, [str(k) for k in rng.permutation(30)] + ["X"])     # no 'unknown', ids >= 30 OOV
    wide = {t: (i if i < 22 else i + 70000) for t, i in vocab.items()}                  # ids >= 65536: 32-bit staging path
    for big, nmax, fused in ((False, 45, True), (True, 45, True), (False, 45, False), (True, 64, False),
                             (False, 100, None), (True, 300, None)):      # > 64 atoms: hash-table path
        mols = []
        for i in range(200):
            n = int(rng.integers(1, nmax + 1))
            m = int(rng.integers(0, 200 if big else 100)) if i % 9 else 0
            s = rng.integers(0, n, m)
            d = rng.integers(0, n, m)
            mols.append({"n": n, "src": s.tolist(), "dst": d.tolist(), "x": rng.integers(-1, 12, n).tolist(),
                         "attr": rng.integers(0, 7, m).tolist()})
        gb = _zinc_batch(mols, dev)
        for max_len, voc in ((1024, vocab), (77, vocab), (1024, wide), (50, wide)):
            rows = engine.tokenize_zinc(gb, voc, max_len, fused=fused).rows_cpu()
            for mo, got in zip(mols, rows):
                want = O.zinc_ids(mo["x"], mo["src"], mo["dst"], mo["attr"], voc, max_len, max_len)
                assert got.tolist() == want.tolist()


@pytest.mark.parametrize("B", [32, 128, 256])
def test_zinc_fused_collate_vs_oracle(dev, golden_zinc, B):
    """molecules -> padded batches in one kernel == oracle tokenize + oracle collate, batch by batch;
    golden molecules (every truncation case) and adversarial shapes, 16-bit and 32-bit staging."""
    from gtok import engine
    rng = np.random.default_rng(11)
    sets = [(golden_zinc["mols"], golden_zinc["L1024"]["vocab"])]
    vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in rng.permutation(40)] + ["X"])
    mols = []
    for i in range(700):
        n = int(rng.integers(1, 46))
        m = int(rng.integers(0, 120)) if i % 9 else 0
        mols.append({"n": n, "src": rng.integers(0, n, m).tolist(), "dst": rng.integers(0, n, m).tolist(),
                     "x": rng.integers(-1, 12, n).tolist(), "attr": rng.integers(0, 7, m).tolist()})
    sets.append((mols, vocab))
    sets.append((mols[:333], {t: (i if i < 22 else i + 70000) for t, i in vocab.items()}))
    for ms, voc in sets:
        gb = _zinc_batch(ms, dev)
        for max_len in (1024, 121, 64):
            ep, lens = engine.zinc_collate_epoch(gb, voc, max_len, B)
            want = [O.zinc_ids(m["x"], m["src"], m["dst"], m["attr"], voc, max_len, max_len) for m in ms]
            assert lens.cpu().tolist() == [len(w) for w in want]
            assert ep.num_batches == (len(ms) + B - 1) // B
            for k in range(ep.num_batches):
                Xw, Aw, _ = O.collate(want[k * B:(k + 1) * B], [0] * len(want[k * B:(k + 1) * B]), voc["<pad>"])
                X, A = ep.batch(k)
                assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw), (max_len, k)


def test_zinc_collate_packed_adversarial(dev, golden_zinc):
    """gtok_zinc_collate_packed (16-bit columns, byte atoms) == gtok_zinc_collate on shuffled / duplicated / self-loop
    columns, out-of-range bond types 0 and 5..15, atom types up to 255, truncation; values that do not fit fall back."""
    from gtok import engine
    from gtok.batch import PackedGraphBatch
    from gtok.pipeline import ZincIbttPipeline
    rng = np.random.default_rng(21)
    vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in rng.permutation(50)] + ["X", "unknown"])
    mols = []
    for i in range(900):
        n = int(rng.integers(1, 64))
        m = int(rng.integers(0, 120)) if i % 7 else 0
        mols.append({"n": n, "src": rng.integers(0, n, m).tolist(), "dst": rng.integers(0, n, m).tolist(),
                     "x": rng.integers(0, 256, n).tolist(), "attr": rng.integers(0, 16, m).tolist()})
    gb = _zinc_batch(mols, "cpu")
    for max_len in (1024, 90):
        pipe = ZincIbttPipeline(gb, vocab, max_len, 64, dev, mode="collate")
        ref = pipe.run(gb.pin_memory())
        tot = int(ref.host_xoff[-1])
        Xa, Aa, la = ref.padded.X[:tot].clone(), ref.padded.attn[:tot].clone(), ref.host_lengths.clone()
        pf = PackedGraphBatch.from_batch(gb, fuse_columns=True)
        assert pf.col16 is not None
        got = pipe.run(pf.pin_memory())
        assert int(got.host_xoff[-1]) == tot and torch.equal(got.host_lengths, la)
        assert torch.equal(got.padded.X[:tot], Xa) and torch.equal(got.padded.attn[:tot], Aa)
        want = [O.zinc_ids(m["x"], m["src"], m["dst"], m["attr"], vocab, max_len, max_len) for m in mols[:64]]
        Xw, Aw, _ = O.collate(want, [0] * 64, vocab["<pad>"])
        X0, A0 = got.padded.batch(0)
        assert np.array_equal(X0.cpu().numpy(), Xw) and np.array_equal(A0.cpu().numpy(), Aw)
    big = _zinc_batch([{"n": 3, "src": [0, 1], "dst": [1, 2], "x": [0, 300, 1], "attr": [1, 17]}], "cpu")
    assert PackedGraphBatch.from_batch(big, fuse_columns=True).col16 is None       # 300 / 17 do not fit: separate arrays


def test_zinc_fused_collate_equals_two_step_large(dev):
    """200k ZINC-shaped molecules: the fused kernel's X/attn/extents are byte-identical to
    gtok_zinc_tokenize followed by gtok_pad_mask_batched."""
    from gtok import engine, synth
    G = 200_000
    gb = synth.zinc_shaped(G, seed=3, device=dev)
    vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in range(37)] + ["X", "unknown"])
    rag = engine.tokenize_zinc(gb, vocab, 1024)
    ep2 = engine.pad_mask_epoch(rag.tokens, rag.offsets, rag.lengths, vocab["<pad>"], 128)
    ep1, lens = engine.zinc_collate_epoch(gb, vocab, 1024, 128)
    assert torch.equal(lens, rag.lengths) and torch.equal(ep1.xoff, ep2.xoff) and torch.equal(ep1.lmax, ep2.lmax)
    tot = int(ep2.xoff[-1])
    assert torch.equal(ep1.X[:tot], ep2.X[:tot]) and torch.equal(ep1.attn[:tot], ep2.attn[:tot])


def test_zinc_properties_large(dev):
    """1M ZINC-shaped molecules (BASELINE config 2 size): closed-form lengths, checksum of a
    strided sample against the oracle, structural invariants on every row start."""
    from gtok import engine, synth
    G = 1_000_000
    gb = synth.zinc_shaped(G, seed=0, device=dev)
    vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in range(37)] + ["X", "unknown"])
    rag = engine.tokenize_zinc(gb, vocab, 1024)
    n, ed = gb.num_nodes(), gb.num_entries()
    assert torch.equal(rag.lengths.to(torch.int64), 2 * n + 4 * (ed // 2) + 4)      # both directions present
    assert int(rag.offsets[-1]) == int(rag.lengths.sum())
    assert bool((rag.tokens[rag.offsets[:-1]] == vocab["<bos>"]).all())
    assert bool((rag.tokens[rag.offsets[1:] - 1] == vocab["<p>"]).all())
    cpu = gb.to("cpu")
    for g in range(0, G, G // 64):
        nn, s, d, x, a = cpu.graph(g)
        want = O.zinc_ids(x.tolist(), s.tolist(), d.tolist(), a.tolist(), vocab, 1024, 1024)
        assert rag.row(g).cpu().tolist() == want.tolist()


def test_ingest_widen_and_packed_pipeline(dev):
    """compact wire format: widen kernels restore the int32 layout bit for bit, and the pipeline fed with a
    PackedGraphBatch produces the same batches as with the int32 GraphBatch."""
    from gtok import engine, synth
    from gtok._lib import lib, check
    from gtok.batch import PackedGraphBatch
    from gtok.engine import _ptr, _stream
    from gtok.pipeline import ZincIbttPipeline
    rng = np.random.default_rng(5)
    for n, hi, dt in ((0, 10, torch.uint8), (1, 255, torch.uint8), (1000003, 255, torch.uint8), (77, 65535, torch.int16),
                      (524288 + 5, 65535, torch.int16), (1234, 1 << 30, torch.int32)):
        v = torch.from_numpy(rng.integers(0, hi + 1, n).astype(np.int64))
        src = ((v & 0xFFFF).to(torch.int16) if dt == torch.int16 else v.to(dt)).to(dev)
        out = torch.full((n + 8,), -7, dtype=torch.int32, device=dev)
        check(lib.gtok_widen_i32(_ptr(src), src.element_size(), n, _ptr(out), _stream()))
        assert torch.equal(out[:n].cpu().to(torch.int64), v) and bool((out[n:] == -7).all())
    gb = synth.zinc_shaped(5000, seed=4)
    vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in range(37)] + ["X", "unknown"])
    pk = PackedGraphBatch.from_batch(gb)
    assert pk.src.dtype == torch.uint8 and pk.node_cnt.dtype == torch.uint8 and pk.nbytes() * 3 < gb.input_bytes()
    for mode in ("collate", "ragged"):
        pipe = ZincIbttPipeline(gb, vocab, 1024, 128, dev, mode=mode)
        a = pipe.run(gb.pin_memory())
        Xa, Aa, la = a.padded.X.clone(), a.padded.attn.clone(), a.host_lengths.clone()
        tot = int(a.host_xoff[-1])
        pipe.dev_in["src"].zero_()
        pipe.dev_in["node_off"].zero_()
        b = pipe.run(pk.pin_memory())
        assert int(b.host_xoff[-1]) == tot and torch.equal(b.host_lengths, la)
        assert torch.equal(b.padded.X[:tot], Xa[:tot]) and torch.equal(b.padded.attn[:tot], Aa[:tot])
        assert torch.equal(pipe.dev_in["node_off"].cpu(), gb.node_off) and torch.equal(pipe.dev_in["edge_off"].cpu(), gb.edge_off)
        if mode == "collate":
            # fused 16-bit columns + byte atoms: consumed by gtok_zinc_collate_packed without widening
            pf = PackedGraphBatch.from_batch(gb, fuse_columns=True)
            assert pf.col16 is not None and pf.src is None and pf.nbytes() < pk.nbytes()
            c_ = pipe.run(pf.pin_memory())
            assert int(c_.host_xoff[-1]) == tot and torch.equal(c_.host_lengths, la)
            assert torch.equal(c_.padded.X[:tot], Xa[:tot]) and torch.equal(c_.padded.attn[:tot], Aa[:tot])
        # double-buffered uploads: a stream of three different shards, each prefetched during its predecessor
        shards = [synth.zinc_shaped(5000, seed=s_) for s_ in (11, 12, 13)]
        pks = [PackedGraphBatch.from_batch(s_).pin_memory() for s_ in shards]
        p2 = ZincIbttPipeline(shards[0], vocab, 1024, 128, dev, mode=mode, max_nodes=37, max_entries=80,
                              node_capacity=max(int(s_.node_off[-1]) for s_ in shards),
                              entry_capacity=max(int(s_.edge_off[-1]) for s_ in shards))
        for i, (sh, pk_) in enumerate(zip(shards, pks)):
            r = p2.run(pk_, prefetch=pks[i + 1] if i + 1 < len(pks) else None)
            ref = engine.zinc_collate_epoch(sh.to(dev), vocab, 1024, 128)[0] if mode == "collate" else None
            want_len = [len(O.zinc_ids(*[t.tolist() for t in (sh.graph(g)[3], sh.graph(g)[1], sh.graph(g)[2], sh.graph(g)[4])], vocab, 1024, 1024))
                        for g in range(0, 5000, 617)]
            assert [int(r.host_lengths[g]) for g in range(0, 5000, 617)] == want_len
            if ref is not None:
                tot = int(ref.xoff[-1])
                assert int(r.host_xoff[-1]) == tot and torch.equal(r.padded.X[:tot], ref.X[:tot])


# --------------------------------------------------------------------------- K4
def test_collate_permuted_batches(dev):
    from gtok import engine
    rng = np.random.default_rng(2)
    lens = rng.integers(0, 50, 1000)
    seqs = [rng.integers(0, 999, l).astype(np.int64) for l in lens]
    tok = torch.from_numpy(np.concatenate(seqs)).to(dev)
    rl = torch.from_numpy(lens.astype(np.int32)).to(dev)
    rs = torch.zeros(1001, dtype=torch.int64)
    rs[1:] = torch.from_numpy(lens).cumsum(0)
    rs = rs.to(dev)
    perm = rng.permutation(1000)
    ep = engine.pad_mask_epoch(tok, rs, rl, 7, 37, batch_idx=torch.from_numpy(perm).to(dev))
    for k in range(ep.num_batches):
        idx = perm[k * 37:(k + 1) * 37]
        if max(len(seqs[i]) for i in idx) == 0:
            continue
        Xw, Aw, _ = O.collate([seqs[i] for i in idx], [0] * len(idx), 7)
        X, A = ep.batch(k)
        assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
        X1, A1 = engine.pad_mask(tok, rs, rl, 7, batch_idx=torch.from_numpy(idx).to(dev))
        assert np.array_equal(X1.cpu().numpy(), Xw) and np.array_equal(A1.cpu().numpy(), Aw)


# --------------------------------------------------------------------------- K3
def _sent_inputs(recs, dev, labeled_mols=None):
    from gtok import engine
    from gtok.batch import GraphBatch
    if labeled_mols is None:
        gb = GraphBatch.from_lists([r["n"] for r in recs], [r["edges"] for r in recs],
                                   queries=[r["query"] for r in recs] if recs and recs[0].get("query") else None)
    else:
        gb = GraphBatch.from_lists([m["n"] for m in labeled_mols],
                                   [list(zip(m["src"], m["dst"])) for m in labeled_mols],
                                   x_lists=[m["x"] for m in labeled_mols], attr_lists=[m["attr"] for m in labeled_mols])
    return gb, engine.build_csr(gb.to(dev))


def test_sent_unlabeled_seeded_vs_oracle(dev):
    from gtok import engine, synth
    recs = synth.er_config1_numpy(300, seed=4, task="cycle_check")
    recs += [{"n": 1, "edges": [], "query": None}, {"n": 5, "edges": [], "query": None},
             {"n": 2, "edges": [(0, 1)], "query": None}, {"n": 6, "edges": [(0, 1), (2, 3), (4, 5)], "query": None}]
    gb, csr = _sent_inputs(recs, dev)
    for max_length, seed, base in ((600, 0, 0), (600, 12345678901234, 1000), (40, 3, 7)):
        cfg = engine.SentConfig(idx_offset=6, max_length=max_length)
        tok, ln = engine.tokenize_sent(csr, cfg, seed=seed, graph_id_base=base)
        tok, ln = tok.cpu(), ln.cpu().tolist()
        ocfg = S.SentConfig(idx_offset=6, max_length=max_length)
        for g, r in enumerate(recs):
            rp, col, _ = S.canonical_csr(r["n"], [e[0] for e in r["edges"]], [e[1] for e in r["edges"]])
            want = S.sent_tokenize(r["n"], rp, col, ocfg, S.PhiloxWords(seed, base + g))
            assert tok[g, :ln[g]].tolist() == want.tolist(), (g, max_length, seed)


def test_sent_labeled_remap_vs_oracle(dev, golden_zinc):
    from gtok import engine
    mols = golden_zinc["mols"][:128]
    gb, csr = _sent_inputs(None, dev, labeled_mols=mols)
    nmax = max(m["n"] for m in mols)
    for remap in (False, True):
        cfg = engine.SentConfig(idx_offset=6, node_idx_offset=6 + nmax, edge_idx_offset=6 + nmax + 9, max_length=1024,
                                labeled=True, remap_zinc=remap)
        tok, ln = engine.tokenize_sent(csr, cfg, seed=99)
        tok, ln = tok.cpu(), ln.cpu().tolist()
        ocfg = S.SentConfig(6, 6 + nmax, 6 + nmax + 9, 1024, True)
        for g, m in enumerate(mols):
            rp, col, ea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"])
            want = S.sent_tokenize(m["n"], rp, col, ocfg, S.PhiloxWords(99, g), x=m["x"], eattr=ea)
            if remap:
                want = O.remap_zinc(want, 6, 6 + nmax, 6 + nmax + 9)
            assert tok[g, :ln[g]].tolist() == want.tolist(), g


def test_sent_choice_stream_replay_and_query(dev):
    """Explicit choice words (the deterministic-replay mode) + fused <q> u v append."""
    from gtok import engine, synth
    recs = synth.er_config1_numpy(64, seed=8, task="shortest_path")
    gb, csr = _sent_inputs(recs, dev)
    rng = np.random.default_rng(0)
    words = [rng.integers(0, 2 ** 32, 2 * r["n"] + 2, dtype=np.uint64).astype(np.uint32) for r in recs]
    words[3] = words[3][:2]                                   # short stream: words past the end read as 0
    woff = np.zeros(len(recs) + 1, dtype=np.int64)
    woff[1:] = np.cumsum([len(w) for w in words])
    cw = torch.from_numpy(np.concatenate(words).view(np.int32)).to(dev)
    cfg = engine.SentConfig(idx_offset=6, max_length=600, append_query=True)
    tok, ln = engine.tokenize_sent(csr, cfg, choice_words=cw, choice_off=torch.from_numpy(woff).to(dev))
    tok, ln = tok.cpu(), ln.cpu().tolist()
    ocfg = S.SentConfig(idx_offset=6, max_length=600)
    for g, r in enumerate(recs):
        rp, col, _ = S.canonical_csr(r["n"], [e[0] for e in r["edges"]], [e[1] for e in r["edges"]])
        want = S.sent_tokenize(r["n"], rp, col, ocfg, S.StreamWords(words[g]))
        want = O.append_query(want, 6, r["n"], r["query"][0], r["query"][1])
        assert tok[g, :ln[g]].tolist() == want.tolist(), g


def test_sent_validity_and_distribution_large(dev):
    """Free-running mode at scale: every sampled row decodes to the input graph (each edge covered
    exactly once); trail-length and token-frequency statistics match the oracle run with OTHER seeds."""
    from gtok import engine, synth
    G = 100_000
    gb = synth.er_graph_token(G, seed=21, device=dev)
    csr = engine.build_csr(gb)
    cfg = engine.SentConfig(idx_offset=6, max_length=2000)
    tok, ln = engine.tokenize_sent(csr, cfg, seed=777)
    cpu = gb.to("cpu")
    ocfg = S.SentConfig(idx_offset=6, max_length=2000)
    lens_gpu, lens_cpu, resets_gpu, resets_cpu = [], [], 0, 0
    sample = list(range(0, G, G // 400))
    tk = tok[torch.tensor(sample, device=dev)].cpu()
    lns = ln[torch.tensor(sample, device=dev)].cpu().tolist()
    for j, g in enumerate(sample):
        n, s, d, _, _ = cpu.graph(g)
        row = tk[j, :lns[j]].tolist()
        nn, edges, _ = S.sent_decode(row, ocfg)
        rp, col, _ = S.canonical_csr(n, s.tolist(), d.tolist())
        assert nn == n and len(edges) == len(col) // 2
        vis = S.visit_order(n, rp, col, S.PhiloxWords(777, g))
        want_edges = {(min(vis[a], vis[b]), max(vis[a], vis[b])) for a, b in zip(s.tolist(), d.tolist())}
        assert set(edges) == want_edges
        lens_gpu.append(len(row))
        resets_gpu += row.count(S.RESET)
        o = S.sent_tokenize(n, rp, col, ocfg, S.PhiloxWords(4242, g))
        lens_cpu.append(len(o))
        resets_cpu += o.tolist().count(S.RESET)
    # same graphs, independent randomness: means within 3 sigma of each other
    a, b = np.asarray(lens_gpu, float), np.asarray(lens_cpu, float)
    diff = a - b
    assert abs(diff.mean()) < 4 * diff.std() / np.sqrt(len(diff)) + 1e-9
    assert abs(resets_gpu - resets_cpu) < 6 * np.sqrt(resets_gpu + resets_cpu + 1)


def test_sent_free_running_token_frequencies_chi_square(dev):
    """north_star: "trail-length and token-frequency distribution check for the free-running stochastic mode".  The same
    300 graphs are sampled 40 times by the CUDA kernels (seeds 1..40) and 40 times by the oracle (seeds 1001..1040):
    the two token-id histograms must be draws of one distribution (chi-square test of homogeneity), and so must the two
    trail-length samples (two-sample Kolmogorov-Smirnov).  Thresholds at p ~ 1e-4: a wrong neighbour-selection or
    restart distribution (e.g. always picking the lowest neighbour) fails both by orders of magnitude."""
    from scipy import stats
    from gtok import engine, synth
    recs = synth.er_config1_numpy(300, seed=9, task="cycle_check")
    gb, csr = _sent_inputs(recs, dev)
    cfg = engine.SentConfig(idx_offset=6, max_length=2000)
    ocfg = S.SentConfig(idx_offset=6, max_length=2000)
    nbins = 6 + 50
    h_gpu, h_cpu = np.zeros(nbins), np.zeros(nbins)
    len_gpu, len_cpu = [], []
    csrs = [S.canonical_csr(r["n"], [e[0] for e in r["edges"]], [e[1] for e in r["edges"]]) for r in recs]
    for rep in range(40):
        tok, ln = engine.tokenize_sent(csr, cfg, seed=1 + rep)
        keep = torch.arange(tok.shape[1], device=dev)[None, :] < ln[:, None]
        h_gpu += np.bincount(tok[keep].cpu().numpy(), minlength=nbins)[:nbins]
        len_gpu += ln.cpu().tolist()
        for g, r in enumerate(recs):
            o = S.sent_tokenize(r["n"], csrs[g][0], csrs[g][1], ocfg, S.PhiloxWords(1001 + rep, g))
            h_cpu += np.bincount(o, minlength=nbins)[:nbins]
            len_cpu.append(len(o))
    use = (h_gpu + h_cpu) >= 20                                # (SOS / EOS counts are deterministic and equal: harmless)
    chi2, p, dof, _ = stats.chi2_contingency(np.stack([h_gpu[use], h_cpu[use]]))
    assert p > 1e-4, (chi2, dof, p)
    ks = stats.ks_2samp(len_gpu, len_cpu)
    assert ks.pvalue > 1e-4, ks
    # the test has teeth: a biased walk (always the first unvisited neighbour, oracle with a constant-zero stream) is rejected
    h_bad = np.zeros(nbins)
    for rep in range(40):
        for g, r in enumerate(recs):
            h_bad += np.bincount(S.sent_tokenize(r["n"], csrs[g][0], csrs[g][1], ocfg, S.StreamWords([])), minlength=nbins)[:nbins]
    useb = (h_gpu + h_bad) >= 20
    assert stats.chi2_contingency(np.stack([h_gpu[useb], h_bad[useb]]))[1] < 1e-6


def test_sent_sharding_invariance(dev):
    """Rank shards see the same rows as the single-GPU run (seeds keyed by GLOBAL graph id)."""
    from gtok import engine, synth
    gb = synth.er_graph_token(4096, seed=1, device="cpu")
    cfg = engine.SentConfig(idx_offset=6, max_length=600)
    full, fl = engine.tokenize_sent(engine.build_csr(gb.to(dev)), cfg, seed=5)
    for r in range(4):
        sh = gb.slice(r * 1024, (r + 1) * 1024).to(dev)
        t, l = engine.tokenize_sent(engine.build_csr(sh), cfg, seed=5, graph_id_base=r * 1024)
        assert torch.equal(l, fl[r * 1024:(r + 1) * 1024])
        m = torch.arange(t.shape[1], device=dev)[None, :] < l[:, None]
        assert torch.equal(t[m], full[r * 1024:(r + 1) * 1024][m])


# --------------------------------------------------------------------------- large graphs (BASELINE config 5)
@pytest.mark.parametrize("N", [64, 65, 128, 129, 200, 256, 257, 512, 4096])
def test_large_graph_sweep_ibtt_sent_collate(dev, N):
    """64..4096-node graphs, mean degree ~8, untruncated: IBTT rows, SENT rows (general warp-per-graph
    kernel beyond 64 nodes) and the padded batch, all against the oracle."""
    from gtok import engine, synth
    G = 6 if N >= 512 else 24
    gb = synth.er_sparse(G, N, mean_degree=8.0, seed=N)
    vocab = {t: i for i, t in enumerate(O.SPECIAL)}
    for k in range(N):
        vocab[str((k * 7919) % N)] = len(vocab)            # scrambled, data-dependent ids
    vocab["has_cycle"] = len(vocab)
    E = int(gb.num_entries().max())
    max_len = 3 * E + N + 5
    rag = engine.tokenize_ibtt(gb.to(dev), vocab, "cycle_check", max_len)
    rows = rag.rows_cpu()
    seqs = []
    for g in range(G):
        n, s, d, _, _ = gb.graph(g)
        want = O.ibtt_ids(n, s.numpy(), d.numpy(), "cycle_check", vocab, max_len)
        assert rows[g].tolist() == want.tolist()
        seqs.append(want)
    X, A = engine.pad_mask(rag.tokens, rag.offsets, rag.lengths, 0)
    Xw, Aw, _ = O.collate(seqs, [0] * G, 0)
    assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
    # SENT, untruncated (bound: 2 + nodes + resets + members + brackets)
    bound = 2 + 4 * N + E + 8
    cfg = engine.SentConfig(idx_offset=6, max_length=bound)
    csr = engine.build_csr(gb.to(dev))
    tok, ln = engine.tokenize_sent(csr, cfg, seed=N)
    tok, ln = tok.cpu(), ln.cpu().tolist()
    ocfg = S.SentConfig(idx_offset=6, max_length=bound)
    for g in range(G if N < 4096 else 2):
        n, s, d, _, _ = gb.graph(g)
        rp, col, _ = S.canonical_csr(n, s.tolist(), d.tolist())
        want = S.sent_tokenize(n, rp, col, ocfg, S.PhiloxWords(N, g))
        assert tok[g, :ln[g]].tolist() == want.tolist(), (N, g)


@pytest.mark.parametrize("N", [70, 128, 256, 300, 700])
def test_sent_mid_labeled_truncated_replay(dev, N):
    """65..256-node graphs on the multi-word thread-per-graph kernel, larger ones on the CSR thread-per-graph kernel: labeled mode with the ZINC remap, truncation
    at max_length, dense graphs (brackets longer than the token ring), explicit choice stream + query append."""
    from gtok import engine
    from gtok.batch import GraphBatch
    rng = np.random.default_rng(N)
    mols = []
    for i in range(40):
        n = int(rng.integers(65, N + 1)) if N > 65 else 70
        m = int(rng.integers(n, 6 * n)) if i % 5 else int(min(n * (n - 1) // 3, 20000))      # every fifth graph is dense
        s_, d_ = rng.integers(0, n, m), rng.integers(0, n, m)
        mols.append({"n": n, "src": s_.tolist(), "dst": d_.tolist(), "x": rng.integers(0, 12, n).tolist(),
                     "attr": rng.integers(1, 6, m).tolist()})
    gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                               x_lists=[m["x"] for m in mols], attr_lists=[m["attr"] for m in mols],
                               queries=[(int(rng.integers(0, m["n"])), int(rng.integers(0, m["n"]))) for m in mols]).to(dev)
    csr = engine.build_csr(gb)
    if N <= 256:
        assert csr.adj_words == (2 if N <= 128 else 4)
    else:
        assert csr.adj64 is None
    nmax = max(m["n"] for m in mols)
    for max_length, remap, append in ((20000, False, False), (301, True, False), (1000, True, True)):
        cfg = engine.SentConfig(idx_offset=6, node_idx_offset=6 + nmax, edge_idx_offset=6 + nmax + 12, max_length=max_length,
                                labeled=True, remap_zinc=remap, append_query=append)
        tok, ln = engine.tokenize_sent(csr, cfg, seed=5)
        tok, ln = tok.cpu(), ln.cpu().tolist()
        ocfg = S.SentConfig(6, 6 + nmax, 6 + nmax + 12, max_length, True)
        for g, m in enumerate(mols):
            rp, col, ea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"])
            want = S.sent_tokenize(m["n"], rp, col, ocfg, S.PhiloxWords(5, g), x=m["x"], eattr=ea)
            if remap:
                want = O.remap_zinc(want, 6, 6 + nmax, 6 + nmax + 12)
            if append:
                want = O.append_query(want, 6, m["n"], int(gb.query_u[g]), int(gb.query_v[g]))
            assert tok[g, :ln[g]].tolist() == np.asarray(want).tolist(), (N, max_length, g)


def test_sent_small_labeled_dense_long_events(dev):
    """<= 64-node labeled graphs whose single events add up to 26 tokens (vertex + label + 10 bracketed (edge label,
    vertex) pairs + edge label): the per-lane token ring must drain fast enough (found by tools/fuzz_gpu.py seed 2)."""
    from gtok import engine
    from gtok.batch import GraphBatch
    rng = np.random.default_rng(11)
    mols = []
    for i in range(160):
        n = int(rng.integers(12, 65))
        m = int(rng.integers(4 * n, 12 * n))
        s_, d_ = rng.integers(0, n, m), rng.integers(0, n, m)
        mols.append({"n": n, "src": s_.tolist(), "dst": d_.tolist(), "x": rng.integers(0, 12, n).tolist(),
                     "attr": rng.integers(1, 6, m).tolist()})
    gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                               x_lists=[m["x"] for m in mols], attr_lists=[m["attr"] for m in mols]).to(dev)
    csr = engine.build_csr(gb)
    assert csr.adj64 is not None and csr.adj_words == 1
    for max_length in (30000, 777):
        cfg = engine.SentConfig(idx_offset=6, node_idx_offset=70, edge_idx_offset=82, max_length=max_length, labeled=True)
        tok, ln = engine.tokenize_sent(csr, cfg, seed=21)
        tok, ln = tok.cpu(), ln.cpu().tolist()
        ocfg = S.SentConfig(6, 70, 82, max_length, True)
        for g, m in enumerate(mols):
            rp, col, ea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"])
            want = S.sent_tokenize(m["n"], rp, col, ocfg, S.PhiloxWords(21, g), x=m["x"], eattr=ea)
            assert tok[g, :ln[g]].tolist() == np.asarray(want).tolist(), (max_length, g)


def test_batches_without_any_edge(dev):
    """Every graph of the batch has zero edges: the edge arrays are empty tensors (data_ptr 0) and every entry point
    must still produce the reference rows (found by tools/fuzz_gpu.py seed 3)."""
    from gtok import engine
    from gtok.batch import GraphBatch, PackedGraphBatch
    ns = [1, 3, 7, 2, 5] * 9
    qs = [(0, n - 1) for n in ns]
    gb = GraphBatch.from_lists(ns, [[] for _ in ns], queries=qs).to(dev)
    for task in ("cycle_check", "shortest_path"):
        vocab = engine.vocab_from_graphs(gb, task, ["no"] * len(ns))[0]
        texts = [O.graph_token_text(n, [], task, "no", q) for n, q in zip(ns, qs)]
        assert vocab == O.vocab_from_texts(texts)[0]
        want = [O.ibtt_ids(n, np.zeros(0, np.int32), np.zeros(0, np.int32), task, vocab, 600, q) for n, q in zip(ns, qs)]
        rows = engine.tokenize_ibtt(gb, vocab, task, 600).rows_cpu()
        assert [r.tolist() for r in rows] == [w.tolist() for w in want]
        ep, lens = engine.ibtt_collate_epoch(gb, vocab, task, 600, 16)
        assert lens.cpu().tolist() == [len(w) for w in want]
        for k in range(ep.num_batches):
            Xw, Aw, _ = O.collate(want[k * 16:(k + 1) * 16], [0] * len(want[k * 16:(k + 1) * 16]), vocab["<pad>"])
            X, A = ep.batch(k)
            assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
    # ZINC: atoms only
    xs = [list(range(n)) for n in ns]
    zb = GraphBatch.from_lists(ns, [[] for _ in ns], x_lists=xs, attr_lists=[[] for _ in ns])
    zv = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(i) for i in range(28)])
    want = [O.zinc_ids(x, [], [], [], zv, 64, 64) for x in xs]
    assert [r.tolist() for r in engine.tokenize_zinc(zb.to(dev), zv, 64).rows_cpu()] == [w.tolist() for w in want]
    ep, lens = engine.zinc_collate_epoch(zb.to(dev), zv, 64, 32)
    assert lens.cpu().tolist() == [len(w) for w in want]
    from gtok.pipeline import ZincIbttPipeline
    pipe = ZincIbttPipeline(zb, zv, 64, 32, dev, mode="collate")
    res = pipe.run(PackedGraphBatch.from_batch(zb, fuse_columns=True).pin_memory())
    assert res.host_lengths.tolist() == [len(w) for w in want]
    ep2 = res.padded
    for k in range(ep.num_batches):
        Xw, Aw, _ = O.collate(want[k * 32:(k + 1) * 32], [0] * len(want[k * 32:(k + 1) * 32]), zv["<pad>"])
        for e in (ep, ep2):
            X, A = e.batch(k)
            assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
    # SENT on edgeless graphs
    csr = engine.build_csr(gb)
    tok, ln = engine.tokenize_sent(csr, engine.SentConfig(idx_offset=6, max_length=100), seed=1)
    for g, n in enumerate(ns):
        rp, col, _ = S.canonical_csr(n, [], [])
        w = S.sent_tokenize(n, rp, col, S.SentConfig(idx_offset=6, max_length=100), S.PhiloxWords(1, g))
        assert tok[g, :int(ln[g])].cpu().tolist() == w.tolist()


# --------------------------------------------------------------------------- K6
def test_graph_labels_match_reference_golden(dev):
    """has_cycle from the kernels == the reference's own DFS label oracle (golden fixture), on both kernels
    (thread-per-graph bit rows; warp-per-graph bitmaps) with repeated edges and self loops in the lists."""
    import json
    import os
    from gtok import engine
    from gtok.batch import GraphBatch
    from oracle import tasks_oracle as T
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tasks_labels.json")))["cycle_cases"]
    rng = np.random.default_rng(0)
    qs = [(int(rng.integers(0, c["n"])), int(rng.integers(0, c["n"]))) for c in gold]
    gb = GraphBatch.from_lists([c["n"] for c in gold], [[tuple(e) for e in c["edges"]] for c in gold], queries=qs).to(dev)
    csr = engine.build_csr(gb)
    want_c = [int(c["has_cycle"]) for c in gold]
    want_d = [T.shortest_distance(c["n"], [tuple(e) for e in c["edges"]], u, v) for c, (u, v) in zip(gold, qs)]
    assert csr.adj64 is not None and csr.adj_words == 1
    cyc, dist = engine.graph_labels(csr, edge_off=gb.edge_off, distance=True)
    assert cyc.cpu().tolist() == want_c and dist.cpu().tolist() == want_d
    import dataclasses
    general = dataclasses.replace(csr, adj64=None)                        # force the warp-per-graph kernel
    cyc2, dist2 = engine.graph_labels(general, edge_off=gb.edge_off, distance=True)
    assert cyc2.cpu().tolist() == want_c and dist2.cpu().tolist() == want_d
    only_c, none_d = engine.graph_labels(csr, edge_off=gb.edge_off)
    assert none_d is None and only_c.cpu().tolist() == want_c


@pytest.mark.parametrize("N", [64, 65, 300, 2500])
def test_graph_labels_larger_graphs_vs_oracle(dev, N):
    from gtok import engine
    from gtok.batch import GraphBatch
    from oracle import tasks_oracle as T
    rng = np.random.default_rng(N)
    recs = []
    for i in range(24):
        n = int(rng.integers(max(2, N // 2), N + 1))
        if i % 3 == 0:      # forest with a few components
            edges = [(int(rng.integers(max(0, v - 40), v)), v) for v in range(1, n) if rng.random() < 0.97]
        elif i % 3 == 1:    # long path (deep BFS) + chords
            edges = [(v - 1, v) for v in range(1, n)] + [(int(a), int(b)) for a, b in zip(rng.integers(0, n, 3), rng.integers(0, n, 3)) if a != b]
        else:               # sparse ER-like, mean degree ~2: many components, unreachable pairs
            m = n
            edges = list({(int(min(a, b)), int(max(a, b))) for a, b in zip(rng.integers(0, n, m), rng.integers(0, n, m)) if a != b})
        recs.append((n, edges, (int(rng.integers(0, n)), int(rng.integers(0, n)))))
    gb = GraphBatch.from_lists([r[0] for r in recs], [r[1] for r in recs], queries=[r[2] for r in recs]).to(dev)
    csr = engine.build_csr(gb)
    cyc, dist = engine.graph_labels(csr, edge_off=gb.edge_off, distance=True)
    assert cyc.cpu().tolist() == [int(T.has_cycle_dfs(n, e)) for n, e, _ in recs]
    assert dist.cpu().tolist() == [T.shortest_distance(n, e, u, v) for n, e, (u, v) in recs]
    # queries that are not nodes of their graph: distance -1 (INF), on both kernels; graphs without nodes / edges
    import dataclasses
    bad_q = dataclasses.replace(csr, query_u=csr.query_u.clone().fill_(-3), query_v=csr.query_v + 100000)
    assert engine.graph_labels(bad_q, distance=True)[1].cpu().tolist() == [-1] * len(recs)
    assert engine.graph_labels(dataclasses.replace(bad_q, adj64=None), distance=True)[1].cpu().tolist() == [-1] * len(recs)
    empty = engine.build_csr(GraphBatch.from_lists([0, 1, 3], [[], [], []], queries=[(0, 0), (0, 0), (2, 2)]).to(dev))
    c0, d0 = engine.graph_labels(empty, distance=True)
    assert c0.cpu().tolist() == [0, 0, 0] and d0.cpu().tolist() == [-1, 0, 0]
    simple, _ = engine.graph_labels(csr)                                   # edge_off None: M = CSR entries / 2
    assert simple.cpu().tolist() == [int(T.has_cycle_dfs(n, sorted(set(e)))) for n, e, _ in recs]


def test_query_positions_vs_reference_expression(dev, golden_gt):
    """gtok_find_token on real padded epochs == (x[b] == q).nonzero()[0] row by row (train_ibtt.py:91-96), for the
    IBTT <q> id, for a per-batch id (AGTT, train_agtt.py:131) and for an id that is absent / cut off by truncation."""
    from gtok import engine
    recs, texts, ex = gt_examples("shortest_path", 300, seed=9)
    vocab, _ = O.vocab_from_texts(texts)
    gb = _gb_from_recs(recs, True).to(dev)
    for max_len, B in ((600, 32), (90, 7)):
        ep, _ = engine.ibtt_collate_epoch(gb, vocab, "shortest_path", max_len, B)
        per_batch = torch.tensor([vocab["<q>"] if k % 3 else vocab.get("5", 0) for k in range(ep.num_batches)], dtype=torch.int64, device=dev)
        for tok, tpb in ((vocab["<q>"], None), (10 ** 9, None), (0, per_batch)):
            pos = engine.query_positions(ep, tok, tpb).cpu().tolist()
            r = 0
            for k in range(ep.num_batches):
                X, _ = ep.batch(k)
                t = int(per_batch[k]) if tpb is not None else tok
                for b in range(X.shape[0]):
                    nz = (X[b] == t).nonzero(as_tuple=True)[0]
                    assert pos[r] == (int(nz[0]) if len(nz) else -1), (max_len, k, b)
                    r += 1
            assert r == ep.num_rows


def test_generate_er_matches_oracle_and_shards(dev):
    """On-device ER graphs == oracle/tasks_oracle.py er_graph bit for bit; a shard generated with graph_id_base is
    the same slice of the full job; labels from the generated graphs agree with the DFS / BFS oracles; the generated
    batch tokenizes through K1."""
    from gtok import engine
    from oracle import tasks_oracle as T
    G = 700
    gb = engine.generate_er(G, seed=5, device=dev)
    n = gb.num_nodes().cpu().tolist()
    eo = gb.edge_off.cpu().tolist()
    src, dst = gb.src.cpu().tolist(), gb.dst.cpu().tolist()
    qu, qv = gb.query_u.cpu().tolist(), gb.query_v.cpu().tolist()
    recs = []
    for g in range(G):
        wn, we, (wu, wv) = T.er_graph(5, g)
        assert n[g] == wn and (qu[g], qv[g]) == (wu, wv), g
        assert list(zip(src[eo[g]:eo[g + 1]], dst[eo[g]:eo[g + 1]])) == we, g
        recs.append((wn, we, (wu, wv)))
    part = engine.generate_er(200, seed=5, graph_id_base=300, device=dev)
    assert part.num_nodes().cpu().tolist() == n[300:500]
    assert part.src.cpu().tolist() == src[eo[300]:eo[500]] and part.query_v.cpu().tolist() == qv[300:500]
    csr = engine.build_csr(gb)
    cyc, dist = engine.graph_labels(csr, edge_off=gb.edge_off, distance=True)
    assert cyc.cpu().tolist() == [int(T.has_cycle_dfs(a, e)) for a, e, _ in recs]
    assert dist.cpu().tolist() == [T.shortest_distance(a, e, u, v) for a, e, (u, v) in recs]
    labels = [T.distance_label_token(d) for d in dist.cpu().tolist()]
    vocab, _ = engine.vocab_from_graphs(gb, "shortest_path", labels)
    texts = [O.graph_token_text(a, e, "shortest_path", lt, q) for (a, e, q), lt in zip(recs, labels)]
    assert vocab == O.vocab_from_texts(texts)[0]
    rows = engine.tokenize_ibtt(gb, vocab, "shortest_path", 600).rows_cpu()
    for t, got in zip(texts[:100], rows[:100]):
        assert got.tolist() == O.encode_text(t, vocab, 600)
    # other bucket shapes: one node (no pair), larger graphs (several Philox trips per lane), p = 0
    for buckets, p0, p1, gcount in ((((1, 3),), 0.5, 0.5, 64), (((150, 260), (2, 9)), 0.01, 0.08, 40), (((5, 30),), 0.0, 0.0, 33)):
        b2 = engine.generate_er(gcount, seed=11, buckets=buckets, p_min=p0, p_max=p1, device=dev)
        n2, eo2 = b2.num_nodes().cpu().tolist(), b2.edge_off.cpu().tolist()
        s2, d2 = b2.src.cpu().tolist(), b2.dst.cpu().tolist()
        for g in range(gcount):
            wn, we, wq = T.er_graph(11, g, buckets=buckets, p_min=p0, p_max=p1)
            assert n2[g] == wn and list(zip(s2[eo2[g]:eo2[g + 1]], d2[eo2[g]:eo2[g + 1]])) == we, (buckets, g)
            assert (int(b2.query_u[g]), int(b2.query_v[g])) == wq
    assert engine.generate_er(0, device=dev).num_graphs == 0


@pytest.mark.parametrize("labeled", [False, True])
def test_sent_workspace_kernel_equals_shared_memory_kernel(dev, labeled):
    """> 256 nodes: the walk with its per-node state in the HBM workspace (gtok_sent_tokenize_ws) and the one with the
    state in shared memory (gtok_sent_tokenize) are the same function of (graph, choice words), rows longer than a
    warp included; both equal the oracle on a sample.  Also: truncation, query append, explicit choice stream."""
    from gtok import engine
    from gtok.batch import GraphBatch
    rng = np.random.default_rng(17 + labeled)
    mols = []
    for i in range(36):
        n = int(rng.integers(257, 1400))
        if i % 6 == 0:
            m = int(min(n * 40, 60000))                   # hubs: rows of 80+ entries (chunked path)
        elif i % 6 == 1:
            m = n // 3                                    # many components, isolated nodes, restarts at unvisited nodes
        else:
            m = 4 * n
        s_, d_ = rng.integers(0, n, m), rng.integers(0, n, m)
        if i % 6 == 2:
            d_[: m // 4] = 0                              # one node of very high degree
        mols.append({"n": n, "src": s_.tolist(), "dst": d_.tolist(), "x": rng.integers(0, 12, n).tolist(),
                     "attr": rng.integers(1, 6, m).tolist()})
    gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                               x_lists=[m["x"] for m in mols], attr_lists=[m["attr"] for m in mols],
                               queries=[(int(rng.integers(0, m["n"])), int(rng.integers(0, m["n"]))) for m in mols]).to(dev)
    csr = engine.build_csr(gb)
    assert csr.adj64 is None
    nmax = max(m["n"] for m in mols)
    words = [rng.integers(0, 2 ** 32, 3 * m["n"], dtype=np.uint64).astype(np.uint32) for m in mols]
    woff = np.zeros(len(mols) + 1, dtype=np.int64)
    woff[1:] = np.cumsum([len(w) for w in words])
    cw = torch.from_numpy(np.concatenate(words).view(np.int32)).to(dev)
    coff = torch.from_numpy(woff).to(dev)
    for max_length, append, stream in ((60000, False, False), (2000, True, False), (777, False, True), (60000, True, True)):
        cfg = engine.SentConfig(idx_offset=6, node_idx_offset=6 + nmax, edge_idx_offset=6 + nmax + 12, max_length=max_length,
                                labeled=labeled, remap_zinc=labeled and max_length == 2000, append_query=append)
        kw = dict(choice_words=cw, choice_off=coff) if stream else dict(seed=9, graph_id_base=5)
        ta, la = engine.tokenize_sent(csr, cfg, use_workspace=True, **kw)
        tb, lb = engine.tokenize_sent(csr, cfg, use_workspace=False, **kw)
        assert csr.scratch is not None
        assert torch.equal(la, lb)
        la_h = la.cpu().tolist()
        ta, tb = ta.cpu(), tb.cpu()
        for g in range(len(mols)):
            assert torch.equal(ta[g, :la_h[g]], tb[g, :la_h[g]]), (max_length, g)
        ocfg = S.SentConfig(6, 6 + nmax, 6 + nmax + 12, max_length, labeled)
        for g in (0, 1, 2, 7, 35):
            m = mols[g]
            rp, col, ea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"])
            wsrc = S.StreamWords(words[g]) if stream else S.PhiloxWords(9, 5 + g)
            want = S.sent_tokenize(m["n"], rp, col, ocfg, wsrc, x=m["x"] if labeled else None, eattr=ea if labeled else None)
            if cfg.remap_zinc:
                want = O.remap_zinc(want, 6, 6 + nmax, 6 + nmax + 12)
            if append:
                want = O.append_query(want, 6, m["n"], int(gb.query_u[g]), int(gb.query_v[g]))
            assert ta[g, :la_h[g]].tolist() == np.asarray(want).tolist(), (max_length, g)


def test_zinc_collate_packed_unique_columns(dev, golden_zinc):
    """De-mirrored wire format: a canonical undirected edge_index travels as its src < dst half, the kernel skips the
    dedupe, and the batches equal the oracle's / the int32 path's bit for bit; a non-canonical shard through the same
    pipeline keeps all its columns (and the dedupe)."""
    from gtok import engine, synth
    from gtok.batch import PackedGraphBatch
    from gtok.pipeline import ZincIbttPipeline
    vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in range(37)] + ["X", "unknown"])
    shards = [synth.zinc_shaped(3000, seed=s_) for s_ in (21, 22)]
    rng = np.random.default_rng(5)
    perm = torch.from_numpy(rng.permutation(int(shards[1].src.numel())))
    # shard 3: shard 2 with its columns shuffled inside each graph (not canonical any more)
    g_of = torch.repeat_interleave(torch.arange(3000), shards[1].num_entries())
    order = torch.argsort(g_of[perm], stable=True)
    idx = perm[order]
    import dataclasses
    shards.append(dataclasses.replace(shards[1], src=shards[1].src[idx], dst=shards[1].dst[idx], edge_attr=shards[1].edge_attr[idx]))
    for max_len in (1024, 90):
        pipe = ZincIbttPipeline(shards[0], vocab, max_len, 128, dev, mode="collate", max_nodes=37, max_entries=90,
                                node_capacity=max(int(s_.node_off[-1]) for s_ in shards),
                                entry_capacity=max(int(s_.edge_off[-1]) for s_ in shards))
        pks = [PackedGraphBatch.from_batch(s_, fuse_columns=True, drop_mirror_columns=True).pin_memory() for s_ in shards]
        assert [p.unique_columns for p in pks] == [True, True, False]
        assert pks[1].nbytes() < 0.66 * pks[2].nbytes()
        for i, (sh, pk_) in enumerate(zip(shards, pks)):
            r = pipe.run(pk_, prefetch=pks[i + 1] if i + 1 < len(pks) else None)
            want = [O.zinc_ids(*[t.tolist() for t in (sh.graph(g)[3], sh.graph(g)[1], sh.graph(g)[2], sh.graph(g)[4])], vocab, max_len, max_len)
                    for g in range(3000)]
            assert r.host_lengths.tolist() == [len(w_) for w_ in want]
            ep = r.padded
            for k in (0, 7, ep.num_batches - 1):
                Xw, Aw, _ = O.collate(want[k * 128:(k + 1) * 128], [0] * len(want[k * 128:(k + 1) * 128]), vocab["<pad>"])
                X, A = ep.batch(k)
                assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw), (max_len, i, k)
            ref = engine.zinc_collate_epoch(sh.to(dev), vocab, max_len, 128)[0]
            tot = int(ref.xoff[-1])
            assert int(r.host_xoff[-1]) == tot and torch.equal(r.padded.X[:tot], ref.X[:tot]) and torch.equal(r.padded.attn[:tot], ref.attn[:tot])
        # the int32 layout of the pipeline was not disturbed by the de-mirrored offsets
        r = pipe.run(shards[0].pin_memory())
        assert int(r.host_xoff[-1]) == int(engine.zinc_collate_epoch(shards[0].to(dev), vocab, max_len, 128)[0].xoff[-1])


def test_sent_workspace_kernel_above_4096_nodes(dev):
    """> 4096 nodes per graph the workspace kernel falls back to 32-bit state words + rowptr reads: same rows as the
    shared-memory kernel and the oracle."""
    from gtok import engine
    from gtok.batch import GraphBatch
    rng = np.random.default_rng(3)
    recs = []
    for n in (4097, 5000, 6100, 300):
        m = 3 * n
        s_, d_ = rng.integers(0, n, m), rng.integers(0, n, m)
        recs.append((n, list(zip(s_.tolist(), d_.tolist()))))
    gb = GraphBatch.from_lists([r[0] for r in recs], [r[1] for r in recs]).to(dev)
    csr = engine.build_csr(gb)
    cfg = engine.SentConfig(idx_offset=6, max_length=40000)
    ta, la = engine.tokenize_sent(csr, cfg, seed=2, use_workspace=True)
    tb, lb = engine.tokenize_sent(csr, cfg, seed=2, use_workspace=False)
    assert torch.equal(la, lb)
    for g, (n, edges) in enumerate(recs):
        L = int(la[g])
        assert torch.equal(ta[g, :L], tb[g, :L])
        if g in (0, 3):
            rp, col, _ = S.canonical_csr(n, [e[0] for e in edges], [e[1] for e in edges])
            want = S.sent_tokenize(n, rp, col, S.SentConfig(idx_offset=6, max_length=40000), S.PhiloxWords(2, g))
            assert ta[g, :L].cpu().tolist() == want.tolist()
    # automatic choice: workspace kernel for these sizes, shared-memory kernel for 1000-node graphs
    tc, lc = engine.tokenize_sent(csr, cfg, seed=2)
    assert torch.equal(lc, la) and torch.equal(tc[0, :int(la[0])], ta[0, :int(la[0])])


@pytest.mark.parametrize("kind", ["sbm", "complete", "star", "path", "ba", "sfn"])
def test_generator_families_match_oracle(dev, kind):
    """The other generator families of docs/synthetic_data.md:93-94 on the device == oracle/tasks_oracle.py er_graph,
    and their labels follow from the shape (star / path: no cycle; complete on >= 3 nodes: cycle)."""
    from gtok import engine
    from oracle import tasks_oracle as T
    kw = dict(kind=kind, num_blocks=3, out_ratio=0.2, p_min=0.15, p_max=0.5, buckets=((1, 6), (10, 50), (60, 90)))
    if kind in ("ba", "sfn"):
        kw.update(num_blocks=5, buckets=((1, 6), (10, 50), (60, 65)))       # attachments m ~ U{1..5}, at most 64 nodes
    if kind == "sfn":
        kw.update(out_ratio=0.6)                                            # triad probability
    okw = dict(kw)
    gb = engine.generate_er(150, seed=8, graph_id_base=40, device=dev, **kw)
    n, eo = gb.num_nodes().cpu().tolist(), gb.edge_off.cpu().tolist()
    src, dst = gb.src.cpu().tolist(), gb.dst.cpu().tolist()
    recs = []
    for g in range(150):
        wn, we, wq = T.er_graph(8, 40 + g, **okw)
        assert n[g] == wn and list(zip(src[eo[g]:eo[g + 1]], dst[eo[g]:eo[g + 1]])) == we, (kind, g)
        assert (int(gb.query_u[g]), int(gb.query_v[g])) == wq
        recs.append((wn, we, wq))
    cyc, dist = engine.graph_labels(engine.build_csr(gb), edge_off=gb.edge_off, distance=True)
    assert cyc.cpu().tolist() == [int(T.has_cycle_dfs(a, e)) for a, e, _ in recs]
    assert dist.cpu().tolist() == [T.shortest_distance(a, e, u, v) for a, e, (u, v) in recs]
    if kind in ("star", "path"):
        assert int(cyc.sum()) == 0 and int((dist < 0).sum()) == 0
    if kind == "complete":
        assert cyc.cpu().tolist() == [int(a >= 3) for a, _, _ in recs]
    if kind in ("ba", "sfn"):
        assert gb.num_entries().cpu().tolist() == [len(e) for _, e, _ in recs] and int((dist < 0).sum()) == 0   # connected
        from gtok._lib import GtokError
        with pytest.raises(GtokError):
            engine.generate_er(4, kind=kind, buckets=((10, 80),), device=dev)                                   # > 64 nodes
    with pytest.raises(ValueError):
        engine.generate_er(4, kind="smallworld", device=dev)


@pytest.mark.parametrize("nmax", [1, 40, 64, 65, 128, 200, 256])
def test_native_csr_build_equals_sort_based_build(dev, nmax):
    """K7: the library's CSR builder (bit rows by atomicOr, popcount row pointers, first-column labels by atomicMin)
    produces exactly the canonical CSR of the sort-based torch build and of the oracle, with repeated columns in
    both directions, self loops, isolated nodes, empty graphs and conflicting labels on repeats."""
    from gtok import engine
    from gtok.batch import GraphBatch
    rng = np.random.default_rng(100 + nmax)
    mols = []
    for i in range(90):
        n = int(rng.integers(0 if i % 17 == 0 else 1, nmax + 1))
        m = 0 if (n < 1 or i % 11 == 0) else int(rng.integers(0, 6 * n + 1))
        s_ = rng.integers(0, max(n, 1), m)
        d_ = rng.integers(0, max(n, 1), m)
        if m > 4:                                            # repeats in both directions with different labels
            s_[1], d_[1] = d_[0], s_[0]
            s_[3], d_[3] = s_[2], d_[2]
        mols.append({"n": n, "src": s_.tolist(), "dst": d_.tolist(), "x": rng.integers(0, 12, n).tolist(),
                     "attr": rng.integers(1, 9, m).tolist()})
    for labeled in (True, False):
        gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                                   x_lists=[m["x"] for m in mols] if labeled else None,
                                   attr_lists=[m["attr"] for m in mols] if labeled else None).to(dev)
        a = engine.build_csr(gb, native=True)
        b = engine.build_csr(gb, native=False)
        assert a.adj_words == b.adj_words == (1 if max(1, max(m["n"] for m in mols)) <= 64 else (2 if max(m["n"] for m in mols) <= 128 else 4))
        for f in ("rowptr", "col", "adj64"):
            assert torch.equal(getattr(a, f), getattr(b, f)), (nmax, labeled, f)
        assert (a.eattr is None) == (not labeled) and (a.eattr is None or torch.equal(a.eattr, b.eattr))
        assert (a.max_nodes, a.max_nnz, a.max_label) == (b.max_nodes, b.max_nnz, b.max_label)
        rp, col = a.rowptr.cpu().numpy(), a.col.cpu().numpy()
        ea = None if a.eattr is None else a.eattr.cpu().numpy()
        n0 = 0
        for m in mols[:25]:
            wrp, wcol, wea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"] if labeled else None)
            lo, hi = rp[n0], rp[n0 + m["n"]]
            assert (rp[n0:n0 + m["n"] + 1] - lo).tolist() == wrp.tolist() and col[lo:hi].tolist() == wcol.tolist()
            if labeled:
                assert ea[lo:hi].tolist() == wea.tolist()
            n0 += m["n"]
    assert engine.build_csr(GraphBatch.from_lists([], []).to(dev)).num_graphs == 0


@pytest.mark.parametrize("shape", ["sparse_300", "mixed_1000", "star_9000", "dense_700", "tiny_and_huge"])
def test_large_csr_build_equals_sort_based_build(dev, shape):
    """K7 beyond 256 nodes (csrc/gtok_csr_large.cu): rows sorted in registers (<= 32 entries), in shared memory (<= 4096)
    and through the neighbour table (longer) - all three against the sort-based torch build; repeats in both directions,
    self loops, conflicting labels, isolated nodes."""
    from gtok import engine
    from gtok.batch import GraphBatch
    rng = np.random.default_rng(len(shape))
    mols = []

    def add(n, m, hub=None):
        s_ = rng.integers(0, n, m)
        d_ = rng.integers(0, n, m)
        if hub is not None:
            s_[: m // 2] = hub                                 # a hub row far beyond 4096 entries (with repeats)
        if m > 4:
            s_[1], d_[1] = d_[0], s_[0]
            s_[3], d_[3] = s_[2], d_[2]
        mols.append({"n": n, "src": s_.tolist(), "dst": d_.tolist(), "x": rng.integers(0, 12, n).tolist(),
                     "attr": rng.integers(1, 9, m).tolist()})
    if shape == "sparse_300":
        for _ in range(40):
            n = int(rng.integers(257, 301))
            add(n, int(rng.integers(0, 3 * n)))
    elif shape == "mixed_1000":
        for _ in range(12):
            n = int(rng.integers(300, 1001))
            add(n, int(rng.integers(n, 12 * n)))
    elif shape == "star_9000":
        add(9000, 30000, hub=7)
        add(300, 100)
    elif shape == "dense_700":
        add(700, 120000)                                       # rows of ~340 entries before dedupe: shared-memory sort
        add(700, 700 * 20, hub=3)                              # hub row of 7000 entries: table path
    else:
        add(1, 0); add(2, 1); add(5000, 4000); add(3, 9); add(260, 0)
    for labeled in (True, False):
        gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                                   x_lists=[m["x"] for m in mols] if labeled else None,
                                   attr_lists=[m["attr"] for m in mols] if labeled else None).to(dev)
        a = engine.build_csr(gb, native=True)
        b = engine.build_csr(gb, native=False)
        assert a.adj64 is None and b.adj64 is None
        for f in ("rowptr", "col"):
            assert torch.equal(getattr(a, f), getattr(b, f)), (shape, labeled, f)
        assert (a.eattr is None) == (not labeled) and (a.eattr is None or torch.equal(a.eattr, b.eattr))
        assert (a.max_nodes, a.max_nnz, a.max_label) == (b.max_nodes, b.max_nnz, b.max_label)
    m = mols[0]
    wrp, wcol, wea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"])
    gb = GraphBatch.from_lists([m["n"]], [list(zip(m["src"], m["dst"]))], x_lists=[m["x"]], attr_lists=[m["attr"]]).to(dev)
    a = engine.build_csr(gb)
    assert a.rowptr.cpu().tolist() == wrp.tolist() and a.col.cpu().tolist() == wcol.tolist() and a.eattr.cpu().tolist() == wea.tolist()


# --------------------------------------------------------------------------- K3 as walk + emission kernels (<= 64 nodes)
def _mixed_small_graphs(seed, count, labeled, nmax=64):
    rng = np.random.default_rng(seed)
    mols = []
    for i in range(count):
        kind = i % 6
        n = [0, 1, int(rng.integers(2, 12)), int(rng.integers(12, 33)), int(rng.integers(33, nmax + 1)), nmax][kind]
        dens = [0, 0, 1.5, 2.5, 3.0, 6.0][kind]
        m = int(n * dens * rng.random() * 2) if n > 1 else 0
        s_, d_ = (rng.integers(0, n, m), rng.integers(0, n, m)) if m else (np.zeros(0, np.int64), np.zeros(0, np.int64))
        mols.append({"n": n, "src": s_.tolist(), "dst": d_.tolist(), "x": rng.integers(0, 9, n).tolist(),
                     "attr": rng.integers(1, 5, m).tolist(), "query": (int(rng.integers(0, max(n, 1))), int(rng.integers(0, max(n, 1))))})
    return mols


@pytest.mark.gpu
@pytest.mark.parametrize("nmax", [64, 200, 256])
@pytest.mark.parametrize("labeled", [False, True])
def test_sent_walk_emit_kernels_vs_single_kernel_and_oracle(dev, labeled, nmax):
    """Graphs of 0..64 nodes (empty, single node, 33+ nodes = two emission rounds / 64-bit masks, dense 64-node graphs)
    through the walk + emission kernels (workspace = item log), the single-kernel path (use_workspace=False) and the oracle:
    every max_length (early stops), both restart policies, both bracket orders, query append, remap."""
    from gtok import engine
    from gtok.batch import GraphBatch
    mols = _mixed_small_graphs(5, 198 if nmax == 64 else 90, labeled, nmax)
    gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                               x_lists=[m["x"] for m in mols] if labeled else None,
                               attr_lists=[m["attr"] for m in mols] if labeled else None,
                               queries=[m["query"] for m in mols]).to(dev)
    csr = engine.build_csr(gb, balance_window=64)
    assert csr.adj64 is not None and csr.adj_words == (1 if nmax == 64 else 2 if nmax <= 128 else 4)
    no, eo = 6 + nmax, 6 + nmax + 9
    for max_length, restart, order, remap, append in ((600, "open_first", "visit_asc", False, False),
                                                      (1, "open_first", "visit_asc", False, False),
                                                      (2, "open_first", "visit_asc", False, True),
                                                      (37, "unvisited", "visit_desc", labeled, True),
                                                      (100, "open_first", "visit_desc", labeled, False),
                                                      (4000, "unvisited", "visit_asc", False, True)):
        cfg = engine.SentConfig(6, no, eo, max_length, labeled, remap_zinc=remap, append_query=append, restart=restart,
                                bracket_order=order)
        assert engine.sent_collate_supported(csr, cfg)
        t1, l1 = engine.tokenize_sent(csr, cfg, seed=17, graph_id_base=5)
        t0, l0 = engine.tokenize_sent(csr, cfg, seed=17, graph_id_base=5, use_workspace=False)
        assert torch.equal(l1, l0)
        if nmax > 64:       # a CSR without adjacency bit rows: the neighbour-list walk kernel instead of the bit-row walk
            import dataclasses
            bare = dataclasses.replace(csr, adj64=None, adj_words=0, scratch=None, order=None)
            t2, l2 = engine.tokenize_sent(bare, cfg, seed=17, graph_id_base=5)
            m2 = torch.arange(t2.shape[1], device=dev)[None, :] < l2[:, None]
            assert torch.equal(l2, l1) and torch.equal(t2[m2], t1[m2])
        t1, t0, ln = t1.cpu(), t0.cpu(), l1.cpu().tolist()
        spec = S.SentSpec(restart=restart, bracket_order=order)
        ocfg = S.SentConfig(6, no, eo, max_length, labeled)
        for g, m in enumerate(mols):
            assert t1[g, :ln[g]].tolist() == t0[g, :ln[g]].tolist(), (max_length, g)
            if g % 3 == 0:
                rp, col, ea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"] if labeled else None)
                want = S.sent_tokenize(m["n"], rp, col, ocfg, S.PhiloxWords(17, 5 + g), x=m["x"] if labeled else None, eattr=ea,
                                       spec=spec)
                if remap:
                    want = O.remap_zinc(want, 6, no, eo)
                if append:
                    want = O.append_query(want, 6, m["n"], m["query"][0], m["query"][1])
                assert t1[g, :ln[g]].tolist() == np.asarray(want).tolist(), (max_length, g)


@pytest.mark.gpu
@pytest.mark.parametrize("nmax", [64, 256])
@pytest.mark.parametrize("labeled", [False, True])
@pytest.mark.parametrize("B", [1, 37, 128])
def test_sent_collate_epoch_equals_two_step(dev, labeled, B, nmax):
    """gtok_sent_walk + gtok_batch_extents + gtok_sent_emit_collate == gtok_sent_tokenize + gtok_pad_mask_batched, byte for
    byte, under a shuffled row order, for explicit choice streams too; one batch is checked against the oracle's collate_fn."""
    from gtok import engine
    from gtok.batch import GraphBatch
    mols = _mixed_small_graphs(9, 300 if nmax == 64 else 120, labeled, nmax)
    no, eo = 6 + nmax, 6 + nmax + 9
    gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                               x_lists=[m["x"] for m in mols] if labeled else None,
                               attr_lists=[m["attr"] for m in mols] if labeled else None,
                               queries=[m["query"] for m in mols]).to(dev)
    csr = engine.build_csr(gb)
    rng = np.random.default_rng(B)
    perm = torch.from_numpy(rng.permutation(len(mols))).to(dev)
    for max_length, append, order in ((600, False, None), (50, True, perm), (600, True, perm[:len(mols) * 2 // 3])):
        cfg = engine.SentConfig(6, no, eo, max_length, labeled, remap_zinc=labeled, append_query=append)
        tok, ln = engine.tokenize_sent(csr, cfg, seed=3, graph_id_base=11)
        stride = tok.shape[1]
        rs = torch.arange(len(mols), device=dev, dtype=torch.int64) * stride
        ep2 = engine.pad_mask_epoch(tok.view(-1), rs, ln, 5, B, batch_idx=order)
        ep1, ln1 = engine.sent_collate_epoch(csr, cfg, 5, B, seed=3, graph_id_base=11, batch_idx=order)
        assert torch.equal(ln1, ln) and torch.equal(ep1.lmax, ep2.lmax) and torch.equal(ep1.xoff, ep2.xoff)
        assert torch.equal(ep1.X, ep2.X) and torch.equal(ep1.attn, ep2.attn)
    # oracle collate_fn on the first batch of the last configuration
    idx = perm[:len(mols) * 2 // 3][:B].cpu().tolist()
    rows = [tok[i, :int(ln[i])].cpu().numpy() for i in idx]
    Xw, Aw, _ = O.collate(rows, [0] * len(rows), 5)
    X, A = ep1.batch(0)
    assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
    # explicit choice streams
    words = [rng.integers(0, 2 ** 32, 2 * m["n"] + 2, dtype=np.uint64).astype(np.uint32) for m in mols]
    woff = np.zeros(len(mols) + 1, dtype=np.int64)
    woff[1:] = np.cumsum([len(w) for w in words])
    cw, co = torch.from_numpy(np.concatenate(words).view(np.int32)).to(dev), torch.from_numpy(woff).to(dev)
    cfg = engine.SentConfig(6, no, eo, 600, labeled)
    tok, ln = engine.tokenize_sent(csr, cfg, choice_words=cw, choice_off=co)
    ep2 = engine.pad_mask_epoch(tok.view(-1), torch.arange(len(mols), device=dev, dtype=torch.int64) * tok.shape[1], ln, 5, B)
    ep1, ln1 = engine.sent_collate_epoch(csr, cfg, 5, B, choice_words=cw, choice_off=co)
    assert torch.equal(ln1, ln) and torch.equal(ep1.X, ep2.X) and torch.equal(ep1.attn, ep2.attn)


@pytest.mark.gpu
@pytest.mark.parametrize("N,count,use_ws", [(40, 64, True), (40, 64, False), (150, 24, None), (400, 8, None)])
def test_sent_append_eos_false_every_kernel_family(dev, N, count, use_ws):
    """Graph2TrailTokenizer(append_eos=False): the row ends without EOS - walk + emission kernels, the single-kernel
    thread-per-graph path, the 65..256-node kernel and the warp-per-graph kernel, against the oracle (with a max_length that
    cuts some rows exactly where EOS would have been)."""
    from gtok import engine, synth
    gb = synth.er_sparse(count, N, mean_degree=5.0, seed=N, device="cpu")
    csr = engine.build_csr(gb.to(dev))
    src, dst = gb.src.numpy(), gb.dst.numpy()
    eo = gb.edge_off.numpy()
    full = engine.tokenize_sent(csr, engine.SentConfig(idx_offset=6, max_length=100000, append_eos=False), seed=5,
                                use_workspace=use_ws)[1].cpu().tolist()
    for max_length in (100000, full[0], full[1] + 1):
        cfg = engine.SentConfig(idx_offset=6, max_length=max_length, append_eos=False)
        tok, ln = engine.tokenize_sent(csr, cfg, seed=5, use_workspace=use_ws)
        tok, ln = tok.cpu(), ln.cpu().tolist()
        ocfg = S.SentConfig(idx_offset=6, max_length=max_length, append_eos=False)
        for g in range(count):
            rp, col, _ = S.canonical_csr(N, src[eo[g]:eo[g + 1]], dst[eo[g]:eo[g + 1]])
            want = S.sent_tokenize(N, rp, col, ocfg, S.PhiloxWords(5, g))
            assert tok[g, :ln[g]].tolist() == want.tolist(), (N, max_length, g)
            assert S.EOS not in want.tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 100, 128, 333, 1000])
def test_zinc_collate_any_batch_size_and_big_molecules(dev, golden_zinc, B):
    """K2+K4 outside the single-pass kernel's shapes: every batch size (the golden molecules vs the reference's own collate
    through the two-step path) and molecules beyond 64 atoms (vs the oracle), incl. truncating max_len values."""
    from gtok import engine
    mols = golden_zinc["mols"]
    gb = _zinc_batch(mols, dev)
    vocab = golden_zinc["L1024"]["vocab"]
    for max_len in (1024, 64):
        ep, lens = engine.zinc_collate_epoch(gb, vocab, max_len, B)
        rag = engine.tokenize_zinc(gb, vocab, max_len)
        ep2 = engine.pad_mask_epoch(rag.tokens, rag.offsets, rag.lengths, vocab["<pad>"], B)
        assert torch.equal(lens, rag.lengths) and torch.equal(ep.xoff, ep2.xoff) and torch.equal(ep.lmax, ep2.lmax)
        assert torch.equal(ep.X, ep2.X) and torch.equal(ep.attn, ep2.attn)
    # molecules of 65..150 atoms (general tokenizer: hash dedupe), odd batch size
    rng = np.random.default_rng(B)
    big = []
    for i in range(90):
        n = int(rng.integers(65, 151))
        m = int(rng.integers(1, 120))
        big.append({"n": n, "src": rng.integers(0, n, m).tolist(), "dst": rng.integers(0, n, m).tolist(),
                    "x": rng.integers(0, 12, n).tolist(), "attr": rng.integers(0, 6, m).tolist()})
    gbb = _zinc_batch(big, dev)
    voc = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in range(150)] + ["X", "unknown"])
    Bb = min(B, 37)
    ep, lens = engine.zinc_collate_epoch(gbb, voc, 1024, Bb)
    want = [O.zinc_ids(m["x"], m["src"], m["dst"], m["attr"], voc, 1024, 1024) for m in big]
    assert lens.cpu().tolist() == [len(w_) for w_ in want]
    for k in range(ep.num_batches):
        Xw, Aw, _ = O.collate(want[k * Bb:(k + 1) * Bb], [0] * len(want[k * Bb:(k + 1) * Bb]), voc["<pad>"])
        X, A = ep.batch(k)
        assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw), k
