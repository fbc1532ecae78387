"""Host half of the drop-in graph_data_loader package: names, parsers, vocab, loaders.
Where the reference checkout is present (build container) the same inputs are also run through
the reference's own functions (differential test); on the GPU box only the golden vectors are used."""
import json
import os
import sys

import pytest

REF = "/root/reference"


def test_export_surface_matches_reference_init():
    import graph_data_loader as g
    hot = ['SPECIAL', 'load_examples', 'load_examples_multi_algorithm', 'determine_num_classes',
           'determine_num_classes_pyg', 'build_vocab_from_texts', 'TokenDataset', 'collate',
           'build_fixed_zinc_vocab', 'extend_vocab_with_dynamic_tokens', 'ZINCTokenizationDataset',
           'GraphTokenDatasetForAutoGraph', 'ZINCDatasetForAutoGraph', 'get_zinc_num_types',
           'get_atom_type_id', 'get_bond_type_id']           # trainer/train_ibtt.py:11-20, train_agtt.py:21-30
    for n in hot:
        assert hasattr(g, n), n
    with pytest.raises(ImportError):
        g.GraphTokenDataset                                  # MPNN/GPS input class: out of scope, loud
    from autograph.datamodules.data.tokenizer import Graph2TrailTokenizer as T
    assert (T.sos, T.reset, T.ladj, T.radj, T.eos, T.pad) == (0, 1, 2, 3, 4, 5)
    t = T(dataset_names=[], max_length=600, truncation_length=600, labeled_graph=True, undirected=True)
    t.set_num_nodes(37)
    t.set_num_node_and_edge_types(9, 4)
    assert (t.idx_offset, t.node_idx_offset, t.edge_idx_offset) == (6, 43, 52)      # train_agtt.py:177-179,538-542


def test_vocab_and_parsers_against_golden(golden_gt, golden_zinc):
    import graph_data_loader as g
    k = golden_gt["doc_kats"]
    assert g.build_vocab_from_texts(k["texts"])[0] == k["vocab"]
    assert g.build_vocab_from_texts(["b a c a b c d"])[0] == k["tie_break"]
    assert g.build_vocab_from_texts(["b a c a b c d"], max_tokens=11)[0] == k["cap11"]
    v, itos = g.build_fixed_zinc_vocab()
    assert v == golden_zinc["fixed_vocab"] and itos[8] == "C"
    assert [g.get_atom_type_id(i) for i in range(9)] == golden_zinc["atom_ids"]
    assert [g.get_bond_type_id(i) for i in range(1, 5)] == golden_zinc["bond_ids"]
    with pytest.raises(ValueError):
        g.get_atom_type_id(9)
    with pytest.raises(ValueError):
        g.get_bond_type_id(0)
    gold = golden_gt["cycle_check_G1000_L600_V600"]
    from graph_data_loader.graph_token_dataset_autograph import parse_graph_from_text, parse_graph_from_json
    for text, (pn, pe), pj in zip(gold["first_texts"], gold["parsed_nodes_edges"], gold["parsed_json"]):
        nodes, edges = parse_graph_from_text(text)
        assert nodes == pn and [list(e) for e in edges] == pe
        e2, n2, lab = parse_graph_from_json({"text": text}, task="cycle_check")
        assert [[list(e) for e in e2], n2, lab] == pj
    assert g.parse_distance_label_from_text("x <p> len3 <eos>") == 2
    assert g.parse_distance_label_from_text("x <p> INF <eos>") is None
    assert g.parse_query_nodes_from_text("a <q> shortest_distance 4 9 <p> len1") == (4, 9)


def _write_tree(tmp_path, task):
    from helpers import gt_examples
    recs, texts, ex = gt_examples(task, 24, seed=5)
    for algo, lo in (("er", 0), ("ba", 12)):
        for split, base in (("train", "tasks_train"), ("test", "tasks_test")):
            d = tmp_path / base / task / algo / split
            d.mkdir(parents=True)
            for i in range(lo, lo + 12):
                rec = [{"text": texts[i]}, {"text": texts[(i + 1) % 24]}] if task == "shortest_path" else {"text": texts[i]}
                (d / f"g{i:03d}.json").write_text(json.dumps(rec))
    return texts


@pytest.mark.parametrize("task", ["cycle_check", "shortest_path"])
def test_loaders_on_generated_tree(tmp_path, task):
    import graph_data_loader as g
    texts = _write_tree(tmp_path, task)
    kw = dict(num_pairs_per_graph=1) if task == "shortest_path" else {}
    ex = g.load_examples_multi_algorithm(str(tmp_path), task, ["er", "ba"], "train", seed=3, num_graphs=8, **kw)
    assert len(ex) == 16 and all(e["text"] in texts for e in ex)
    val = g.load_examples_multi_algorithm(str(tmp_path), task, ["er"], "val", seed=3)       # val falls back to test
    assert len(val) == (24 if task == "shortest_path" else 12)
    ds = g.GraphTokenDatasetForAutoGraph(str(tmp_path), task=task, algorithm=["er", "ba"], split="train", seed=3,
                                         num_graphs=8, **kw)
    assert len(ds) <= 16 and ds[0].edge_index.shape[0] == 2 and ds[0].y.dtype.is_floating_point is False
    gb = ds.graph_batch
    assert gb.num_graphs == len(ds) and int(gb.num_nodes()[0]) == ds[0].num_nodes
    assert g.determine_num_classes(ex, "cycle_check") == 2 and g.determine_num_classes_pyg(ds, "zinc") == 1
    with pytest.raises(RuntimeError):
        g.GraphTokenDatasetForAutoGraph(str(tmp_path), task=task, algorithm="nope", split="train")
    if os.path.isdir(REF):          # differential test against the reference's own loader (same process => same hash salt)
        sys.path.insert(0, REF)
        for m in [m for m in sys.modules if m == "graph_data_loader" or m.startswith("graph_data_loader.")]:
            sys.modules.pop(m)
        sys.path.remove(REF); sys.path.insert(0, REF)
        import importlib
        R = importlib.import_module("graph_data_loader.data_loader")
        try:
            assert "reference" in R.__file__
            rex = R.load_examples_multi_algorithm(str(tmp_path), task, ["er", "ba"], "train", seed=3, num_graphs=8, **kw)
            assert rex == ex
        finally:
            sys.path.remove(REF)
            for m in [m for m in sys.modules if m == "graph_data_loader" or m.startswith("graph_data_loader.")]:
                sys.modules.pop(m)


def test_graph_of_text_roundtrip(golden_gt):
    from graph_data_loader.data_loader import graph_of_text
    from helpers import gt_examples
    for task in ("cycle_check", "shortest_path"):
        recs, texts, _ = gt_examples(task, 32, seed=2)
        for r, t in zip(recs, texts):
            n, edges, tk, q = graph_of_text(t)
            assert (n, edges, tk) == (r["n"], [tuple(e) for e in r["edges"]], task)
            assert q == (tuple(r["query"]) if r["query"] else None)
    assert graph_of_text("<bos> hello world") is None


def test_balance_helpers():
    import graph_data_loader as g
    ex = [{"text": "t", "label": i % 3 == 0} for i in range(30)]
    out = g.balance_classes(ex, "undersample", seed=1)
    assert sum(e["label"] for e in out) == len(out) - sum(e["label"] for e in out) == 10
    assert len(g.balance_classes(ex, "oversample", seed=1)) == 40
    with pytest.raises(ValueError):
        g.balance_classes(ex, "bogus")


def test_autograph_dataset_processed_cache_and_pre_filter(tmp_path):
    """InMemoryDataset contract of the reference class (graph_token_dataset_autograph.py:233-253, :350-351, :406-408): the
    parse runs once per (task, algorithms, split, sampling) key and later constructions load the processed file - here even
    after the raw JSON files are gone; pre_filter sees Data-like objects and drops graphs before they are stored."""
    import shutil
    import graph_data_loader as g
    _write_tree(tmp_path, "cycle_check")
    ds = g.GraphTokenDatasetForAutoGraph(str(tmp_path), task="cycle_check", algorithm=["er", "ba"], split="train", seed=3)
    first = [(d.num_nodes, d.edge_index.tolist(), int(d.y)) for d in ds]
    assert os.path.exists(os.path.join(ds.processed_dir, "gtok_records.pt"))
    shutil.rmtree(tmp_path / "tasks_train")
    ds2 = g.GraphTokenDatasetForAutoGraph(str(tmp_path), task="cycle_check", algorithm=["er", "ba"], split="train", seed=3)
    assert [(d.num_nodes, d.edge_index.tolist(), int(d.y)) for d in ds2] == first
    # another key -> another cache entry; its raw files are still there
    big = g.GraphTokenDatasetForAutoGraph(str(tmp_path), task="cycle_check", algorithm="er", split="test", seed=3,
                                          pre_filter=lambda d: d.num_nodes >= 10 and int(d.y) in (0, 1))
    allg = g.GraphTokenDatasetForAutoGraph(str(tmp_path / "."), task="cycle_check", algorithm="er", split="test", seed=3,
                                           num_graphs=1000)
    assert 0 < len(big) <= len(allg) and all(d.num_nodes >= 10 for d in big)
