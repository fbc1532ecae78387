"""Pin oracle/ibtt_oracle.py to outputs of the reference itself (tests/golden/*.json,
made by tests/golden/make_golden.py which imports /root/reference)."""
import numpy as np
import pytest

from helpers import gt_examples, sha_int64
from oracle import ibtt_oracle as O

GT_CASES = [("cycle_check", 1000, 600, 600), ("shortest_path", 200, 600, 60), ("cycle_check", 64, 97, 30)]


@pytest.mark.parametrize("task,G,max_len,max_vocab", GT_CASES)
def test_graph_token_pipeline_matches_reference(golden_gt, task, G, max_len, max_vocab):
    gold = golden_gt[f"{task}_G{G}_L{max_len}_V{max_vocab}"]
    recs, texts, ex = gt_examples(task, G)
    assert texts[:4] == gold["first_texts"]
    vocab, itos = O.vocab_from_texts(texts, max_tokens=max_vocab)
    assert vocab == gold["vocab"]                      # a2: ids, order, cap
    assert list(vocab) == list(gold["vocab"])
    seqs, labels, kept = O.token_dataset(ex, vocab, max_len)
    assert len(seqs) == gold["num_kept"]               # C3: unlabeled (INF) examples dropped
    assert [len(s) for s in seqs] == gold["lengths"]
    assert labels.tolist() == gold["labels"]
    assert sha_int64(seqs) == gold["ids_sha256"]       # a3 bit-exact over all sequences
    assert [s.tolist() for s in seqs[:24]] == gold["first_seqs"]
    # a4 on every batch of 128
    for bi, b0 in enumerate(range(0, len(seqs), 128)):
        X, A, Y = O.collate(seqs[b0:b0 + 128], labels[b0:b0 + 128], vocab["<pad>"])
        shp, xs, as_, ys = gold["batch_sha"][bi]
        assert list(X.shape) == shp and sha_int64([X]) == xs and sha_int64([Y]) == ys
        import hashlib
        assert hashlib.sha256(A.tobytes()).hexdigest() == as_
    X, A, Y = O.collate(seqs[:8], labels[:8], vocab["<pad>"])
    assert X.tolist() == gold["collate8"]["X"] and A.astype(np.int64).tolist() == gold["collate8"]["attn"]
    assert Y.tolist() == gold["collate8"]["Y"]


@pytest.mark.parametrize("task,G,max_len,max_vocab", GT_CASES)
def test_closed_form_ids_equal_text_pipeline(golden_gt, task, G, max_len, max_vocab):
    """ibtt_ids (what kernel K1 mirrors) == encode_text(graph_token_text(...))."""
    gold = golden_gt[f"{task}_G{G}_L{max_len}_V{max_vocab}"]
    vocab = gold["vocab"]
    recs, texts, ex = gt_examples(task, G)
    out = []
    for r, e in zip(recs, ex):
        if e["label"] is None:
            continue
        src = np.asarray([a for a, _ in r["edges"]], dtype=np.int32)
        dst = np.asarray([b for _, b in r["edges"]], dtype=np.int32)
        ids = O.ibtt_ids(r["n"], src, dst, task, vocab, max_len, r["query"])
        assert len(ids) == O.ibtt_length(r["n"], len(src), task, max_len)
        out.append(ids)
    assert sha_int64(out) == gold["ids_sha256"]


def test_inverse_parser_matches_reference(golden_gt):
    gold = golden_gt["cycle_check_G1000_L600_V600"]
    recs, texts, _ = gt_examples("cycle_check", 16)
    for r, t, (pn, pe) in zip(recs, texts, gold["parsed_nodes_edges"]):
        nodes, edges = O.parse_graph_text(t)
        assert nodes == pn and [list(e) for e in edges] == pe
        assert nodes == list(range(r["n"])) and edges == [tuple(e) for e in r["edges"]]   # round trip a1 -> a8


def test_doc_kats(golden_gt):
    k = golden_gt["doc_kats"]
    vocab, _ = O.vocab_from_texts(k["texts"])
    assert vocab == k["vocab"]
    assert [O.encode_text(t, vocab, 512) for t in k["texts"]] == k["ids"]
    assert O.vocab_from_texts(["b a c a b c d"])[0] == k["tie_break"]             # C1 ties: first seen
    assert O.vocab_from_texts(["b a c a b c d"], max_tokens=11)[0] == k["cap11"]  # C1 cap
    # docs/models/embed_gtt.md:55-60: <q> of the 4-cycle example sits at position 18
    ids = O.encode_text(k["texts"][3], vocab, 512)
    assert ids.index(vocab["<q>"]) == 18
    assert O.SPECIAL == ["<pad>", "<bos>", "<e>", "<n>", "<q>", "<p>", "<eos>", "yes", "no"]


@pytest.mark.parametrize("max_len", [1024, 120, 121, 122, 64])
def test_zinc_pipeline_matches_reference(golden_zinc, max_len):
    gold = golden_zinc[f"L{max_len}"]
    vocab = gold["vocab"]
    mols = golden_zinc["mols"]
    seqs_text, seqs_cf = [], []
    for i, m in enumerate(mols):
        text = O.zinc_item_text(m["x"], m["src"], m["dst"], m["attr"], m["y"], max_len)
        if i < 3:
            assert text == gold["first_texts"][i]
        seqs_text.append(O.encode_text(text, vocab, max_len))
        ids = O.zinc_ids(m["x"], m["src"], m["dst"], m["attr"], vocab, max_len, max_len)
        seqs_cf.append(ids)
        nuniq = len({(min(a, b), max(a, b)) for a, b in zip(m["src"], m["dst"])})
        assert len(ids) == O.zinc_length(m["n"], nuniq, max_len, max_len)
    assert [len(s) for s in seqs_text] == gold["lengths"]
    assert sha_int64(seqs_text) == gold["ids_sha256"]
    assert sha_int64(seqs_cf) == gold["ids_sha256"]
    assert [list(map(int, s)) for s in seqs_cf[:32]] == gold["first_seqs"]
    labels = [int(m["y"]) for m in mols]               # C4: int() truncation toward zero
    assert labels == gold["labels"]
    X, A, Y = O.collate(seqs_cf[:8], labels[:8], vocab["<pad>"])
    assert X.tolist() == gold["collate8"]["X"] and A.astype(np.int64).tolist() == gold["collate8"]["attn"]
    assert Y.tolist() == gold["collate8"]["Y"]


def test_zinc_vocab_kats(golden_zinc):
    v = O.zinc_fixed_vocab()
    assert v == golden_zinc["fixed_vocab"] and list(v) == list(golden_zinc["fixed_vocab"])
    assert [v[a] for a in O.ZINC_ATOMS] == golden_zinc["atom_ids"] == list(range(8, 17))
    assert [v[b] for b in O.ZINC_BONDS] == golden_zinc["bond_ids"] == [17, 18, 19, 20]
    assert v["regression"] == 21 and v["<pad>"] == 2
    assert O.zinc_label_token(4.23) == "val_4_23" and O.zinc_label_token(-2.1) == "val_neg2_10"
    e = O.zinc_extend_vocab(v, ["0", "C", "1", "0", "X"])
    assert (e["0"], e["1"], e["X"]) == (22, 23, 24) and e["C"] == 8


def test_agtt_post_matches_reference(golden_agtt):
    z = golden_agtt["zinc"]
    for raw, want in zip(z["raw"], z["remapped"]):
        got = O.remap_zinc(raw, z["idx_offset"], z["node_idx_offset"], z["edge_idx_offset"])
        assert got.tolist() == want
    seqs = [np.asarray(s) for s in z["remapped"][:8]]
    X, A, Y = O.collate_agtt(seqs, [0.5] * 8)
    assert X.tolist() == z["collate8"]["X"] and A.astype(np.int64).tolist() == z["collate8"]["attn"]
    assert z["collate8"]["Y_dtype"] == "torch.float32" and Y.dtype == np.float32
    sp = golden_agtt["shortest_path"]
    for raw, n, (u, v), want in zip(sp["raw"], sp["num_nodes"], sp["query"], sp["out"]):
        assert O.append_query(np.asarray(raw), sp["idx_offset"], n, u, v).tolist() == want
    X, A, Y = O.collate_agtt([np.asarray(s) for s in sp["out"][:8]], [1] * 8)
    assert X.tolist() == sp["collate8"]["X"] and sp["collate8"]["Y_dtype"] == "torch.int64" and Y.dtype == np.int64
