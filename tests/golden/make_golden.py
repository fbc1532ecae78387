#!/usr/bin/env python
"""Generate tests/golden/*.json by running the REFERENCE ITSELF (read-only at
/root/reference) on seeded synthetic inputs.  Run in the build container only;
the GPU box has no /root/reference, it consumes the committed fixtures.

    python tests/golden/make_golden.py

What is executed from the reference (unmodified, imported from where it lies):
  graph_data_loader.build_vocab_from_texts / TokenDataset / collate   (a2, a3, a4)
  graph_data_loader.zinc_dataset_indexbase.ZINCTokenizationDataset
      .tokenize_molecule / .__getitem__                               (a5, a6)
  graph_data_loader.zinc_vocab.*                                      (a7)
  graph_data_loader.graph_token_dataset_autograph.parse_graph_from_text/json (a8)
  trainer.train_agtt.TokenizedGraphDataset.remap_zinc_tokens /
      .__getitem__ / collate_fn                                       (a11-a13)
torch_geometric, autograph, matplotlib and seaborn are absent offline; they are
replaced by empty stub modules ONLY so the reference files import (none of their
code runs on these paths; the AutoGraph tokenizer is replaced by a fake that
returns token sequences produced by oracle/sent_oracle.py).
The graph-token *serializer* (a1) is external to the reference, so the input
texts come from oracle.ibtt_oracle.graph_token_text (grammar per
docs/synthetic_data.md:46-68); everything downstream of the text is reference code.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "glearning-benchmark_b200"))
REF = "/root/reference"


def _stub_modules():
    class _Any:
        def __init__(self, *a, **k):
            self.__dict__.update(k)

    def _ga(k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any

    tg = types.ModuleType("torch_geometric")
    tgd = types.ModuleType("torch_geometric.datasets")
    tgdata = types.ModuleType("torch_geometric.data")
    tgd.ZINC = _Any
    tgdata.Data = _Any
    tgdata.InMemoryDataset = type("InMemoryDataset", (), {})
    tgdata.Batch = _Any
    tg.datasets, tg.data = tgd, tgdata
    sys.modules.update({"torch_geometric": tg, "torch_geometric.datasets": tgd, "torch_geometric.data": tgdata})
    for name in ("torch_geometric.loader", "torch_geometric.nn", "torch_geometric.utils",
                 "torch_geometric.transforms"):
        m = types.ModuleType(name)
        m.__getattr__ = _ga                       # type: ignore
        sys.modules[name] = m
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    ag = [types.ModuleType(n) for n in ("autograph", "autograph.datamodules", "autograph.datamodules.data",
                                        "autograph.datamodules.data.tokenizer")]

    class Graph2TrailTokenizer:                    # only the class attr is read (train_agtt.py:286)
        pad = 5
    ag[3].Graph2TrailTokenizer = Graph2TrailTokenizer
    for m in ag:
        sys.modules[m.__name__] = m
    os.environ.setdefault("WANDB_MODE", "disabled")


def sha(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(np.asarray(a, dtype=np.int64)).tobytes())
    return h.hexdigest()


def main():
    _stub_modules()
    sys.path.insert(0, REF)
    import graph_data_loader as R
    from graph_data_loader import zinc_dataset_indexbase as RZ
    from graph_data_loader import graph_token_dataset_autograph as RA
    from oracle import ibtt_oracle as O
    from oracle import sent_oracle as S
    from gtok import synth

    # ------------------------------------------------------------------ IBTT graph-token
    out = {}
    for task, G, max_len, max_vocab in (("cycle_check", 1000, 600, 600), ("shortest_path", 200, 600, 60),
                                        ("cycle_check", 64, 97, 30)):
        recs = synth.er_config1_numpy(G, seed=0, task=task)
        texts = [O.graph_token_text(r["n"], r["edges"], task, r["label"], r["query"]) for r in recs]
        if task == "shortest_path":
            ex = []
            for t in texts:
                lab = R.parse_distance_label_from_text(t)
                q = R.parse_query_nodes_from_text(t)
                e = {"text": t, "label": lab}
                if q:
                    e["query_u"], e["query_v"] = q
                ex.append(e)
        else:
            ex = [{"text": t, "label": R.data_loader.parse_yes_no_from_text(t)} for t in texts]
        vocab, _ = R.build_vocab_from_texts(texts, max_tokens=max_vocab)
        ds = R.TokenDataset(ex, vocab, max_len)
        seqs = [s.tolist() for s in ds.seqs]
        labels = [int(l) for l in ds.labels]
        B = 128
        bsha = []
        for b0 in range(0, len(ds), B):
            X, A, Y = R.collate([ds[i] for i in range(b0, min(b0 + B, len(ds)))], vocab["<pad>"])
            bsha.append([list(X.shape), sha([X.numpy()]), hashlib.sha256(A.numpy().tobytes()).hexdigest(),
                         sha([Y.numpy()])])
        X0, A0, Y0 = R.collate([ds[i] for i in range(min(8, len(ds)))], vocab["<pad>"])
        # inverse parser (a8) on the first texts
        parsed = [RA.parse_graph_from_text(t) for t in texts[:16]]
        pj = [RA.parse_graph_from_json({"text": t}, task=task) for t in texts[:16]]
        key = f"{task}_G{G}_L{max_len}_V{max_vocab}"
        out[key] = {
            "task": task, "G": G, "max_len": max_len, "max_vocab": max_vocab, "seed": 0,
            "vocab": vocab, "num_kept": len(ds), "lengths": [len(s) for s in seqs],
            "labels": labels, "ids_sha256": sha(seqs), "first_seqs": seqs[:24],
            "first_texts": texts[:4], "batch_sha": bsha,
            "collate8": {"X": X0.tolist(), "attn": A0.to(torch.int64).tolist(), "Y": Y0.tolist()},
            "parsed_nodes_edges": [[p[0], [list(e) for e in p[1]]] for p in parsed],
            "parsed_json": [[[list(e) for e in p[0]], p[1], p[2]] for p in pj],
        }
    # doc KAT strings (docs/synthetic_data.md:66-68, docs/models/embed_gtt.md:55-60)
    kats = ["<bos> 0 1 <e> 1 2 <e> <n> 0 1 2 <q> 0 <p> d2 <eos>",
            "<bos> 0 1 <e> 1 2 <e> <n> 0 1 2 <q> 0 2 <p> yes <eos>",
            "<bos> 0 1 <e> 1 2 <e> <n> 0 1 2 <q> 0 3 <p> no <eos>",
            "<bos> 0 1 <e> 1 2 <e> 2 3 <e> 0 3 <e> <n> 0 1 2 3 <q> shortest_distance 0 2 <p> len2 <eos>"]
    kv, _ = R.build_vocab_from_texts(kats)
    kds = R.TokenDataset([{"text": t, "label": 0} for t in kats], kv, 512)
    out["doc_kats"] = {"texts": kats, "vocab": kv, "ids": [s.tolist() for s in kds.seqs],
                       "tie_break": R.build_vocab_from_texts(["b a c a b c d"])[0],
                       "cap11": R.build_vocab_from_texts(["b a c a b c d"], max_tokens=11)[0]}
    json.dump(out, open(os.path.join(HERE, "ibtt_graph_token.json"), "w"))

    # ------------------------------------------------------------------ IBTT ZINC
    zout = {}
    gb = synth.zinc_shaped(256, seed=0)
    # a few adversarial molecules: shuffled column order, duplicated columns, reverse-first orientation
    rs = np.random.default_rng(1)
    mols = []
    for g in range(gb.num_graphs):
        n, s, d, x, a = gb.graph(g)
        s, d, x, a = s.numpy().copy(), d.numpy().copy(), x.numpy().copy(), a.numpy().copy()
        if g % 4 == 1:
            p = rs.permutation(len(s)); s, d, a = s[p], d[p], a[p]
        if g % 4 == 2:
            k = rs.integers(0, len(s), 5)
            s, d, a = np.concatenate([s, s[k]]), np.concatenate([d, d[k]]), np.concatenate([a, (a[k] % 4) + 1])
        if g % 4 == 3:
            s, d = d.copy(), s.copy()
            o = np.argsort(-s, kind="stable"); s, d, a = s[o], d[o], a[o]
        mols.append((n, s, d, x, a, float(gb.y[g])))

    class D:
        pass

    def mk(m):
        n, s, d, x, a, y = m
        o = D()
        o.x = torch.tensor(x, dtype=torch.long).reshape(-1, 1)
        o.edge_index = torch.tensor(np.stack([s, d]), dtype=torch.long)
        o.edge_attr = torch.tensor(a, dtype=torch.long)
        o.y = torch.tensor([y], dtype=torch.float32)
        return o
    for max_len in (1024, 120, 121, 122, 64):
        ds = RZ.ZINCTokenizationDataset.__new__(RZ.ZINCTokenizationDataset)
        ds.split, ds.max_len, ds.zinc_dataset = "train", max_len, [mk(m) for m in mols]
        ex = [ds[i] for i in range(len(mols))]
        vocab, _ = R.build_fixed_zinc_vocab()
        dyn = []
        for e in ex:                     # deterministic order (the reference uses set order, C5)
            for t in e["text"].split():
                if t not in vocab and t not in dyn:
                    dyn.append(t)
        vocab = R.extend_vocab_with_dynamic_tokens(vocab, dyn)
        td = R.TokenDataset(ex, vocab, max_len)
        seqs = [s.tolist() for s in td.seqs]
        X0, A0, Y0 = R.collate([td[i] for i in range(8)], vocab["<pad>"])
        zout[f"L{max_len}"] = {"max_len": max_len, "vocab": vocab, "lengths": [len(s) for s in seqs],
                               "labels": [int(l) for l in td.labels], "ids_sha256": sha(seqs),
                               "first_seqs": seqs[:32], "first_texts": [e["text"] for e in ex[:3]],
                               "collate8": {"X": X0.tolist(), "attn": A0.to(torch.int64).tolist(), "Y": Y0.tolist()}}
    zout["mols"] = [{"n": int(m[0]), "src": m[1].tolist(), "dst": m[2].tolist(), "x": m[3].tolist(),
                     "attr": m[4].tolist(), "y": m[5]} for m in mols]
    fv, _ = R.build_fixed_zinc_vocab()
    zout["fixed_vocab"] = fv
    zout["atom_ids"] = [R.get_atom_type_id(i) for i in range(9)]
    zout["bond_ids"] = [R.get_bond_type_id(i) for i in range(1, 5)]
    json.dump(zout, open(os.path.join(HERE, "ibtt_zinc.json"), "w"))

    # ------------------------------------------------------------------ AGTT post-processing
    sys.path.insert(0, REF)
    from trainer import train_agtt as T

    class FakeTok:
        pad = 5

        def __init__(self, idx_offset, max_nodes, ntypes):
            self.idx_offset = idx_offset
            self.node_idx_offset = idx_offset + max_nodes
            self.edge_idx_offset = self.node_idx_offset + ntypes
            self.seqs = {}

        def __call__(self, data):
            return torch.tensor(self.seqs[data.gid], dtype=torch.long)

    aout = {}
    max_nodes = 37
    tok = FakeTok(6, max_nodes, 9)
    cfg = S.SentConfig(idx_offset=6, node_idx_offset=tok.node_idx_offset, edge_idx_offset=tok.edge_idx_offset,
                       max_length=1024, labeled=True)
    datas = []
    for g in range(48):
        n, s, d, x, a, y = mols[g]
        rp, col, ea = S.canonical_csr(n, s, d, a)
        t = S.sent_tokenize(n, rp, col, cfg, S.PhiloxWords(0, g), x=x, eattr=ea)
        o = D(); o.gid = g; o.num_nodes = n; o.y = torch.tensor([y], dtype=torch.float32)
        tok.seqs[g] = t.tolist()
        datas.append(o)
    # raw token sweep hits every remap branch, incl. out-of-range atom/bond types
    sweep = list(range(0, tok.edge_idx_offset + 12))
    o = D(); o.gid = "sweep"; o.num_nodes = 5; o.y = torch.tensor([0.5]); tok.seqs["sweep"] = sweep
    datas.append(o)
    tds = T.TokenizedGraphDataset(datas, tok, task="zinc", remap_to_fixed_vocab=True)
    items = [tds[i] for i in range(len(datas))]
    Xp, Ap, Yp, _ = T.collate_fn(items[:8])
    aout["zinc"] = {"idx_offset": 6, "node_idx_offset": tok.node_idx_offset, "edge_idx_offset": tok.edge_idx_offset,
                    "raw": [tok.seqs[d.gid] for d in datas], "remapped": [it[0].tolist() for it in items],
                    "collate8": {"X": Xp.tolist(), "attn": Ap.to(torch.int64).tolist(), "Y": Yp.tolist(),
                                 "Y_dtype": str(Yp.dtype)}}
    # shortest_path query append (a12) on unlabeled graphs
    tok2 = FakeTok(6, 49, 0)
    cfg2 = S.SentConfig(idx_offset=6, max_length=600)
    recs = synth.er_config1_numpy(32, seed=3, task="shortest_path")
    d2 = []
    for g, r in enumerate(recs):
        src = [e[0] for e in r["edges"]]; dst = [e[1] for e in r["edges"]]
        rp, col, _ = S.canonical_csr(r["n"], src, dst)
        t = S.sent_tokenize(r["n"], rp, col, cfg2, S.PhiloxWords(7, g))
        o = D(); o.gid = g; o.num_nodes = r["n"]; o.query_u, o.query_v = r["query"]
        o.y = torch.tensor([g % 5], dtype=torch.long)
        tok2.seqs[g] = t.tolist()
        d2.append(o)
    tds2 = T.TokenizedGraphDataset(d2, tok2, task="shortest_path", remap_to_fixed_vocab=False)
    it2 = [tds2[i] for i in range(len(d2))]
    X2, A2, Y2, _ = T.collate_fn(it2[:8])
    aout["shortest_path"] = {"idx_offset": 6, "raw": [tok2.seqs[d.gid] for d in d2],
                             "num_nodes": [d.num_nodes for d in d2],
                             "query": [[d.query_u, d.query_v] for d in d2],
                             "out": [it[0].tolist() for it in it2],
                             "collate8": {"X": X2.tolist(), "attn": A2.to(torch.int64).tolist(), "Y": Y2.tolist(),
                                          "Y_dtype": str(Y2.dtype)}}
    json.dump(aout, open(os.path.join(HERE, "agtt_post.json"), "w"))
    for f in ("ibtt_graph_token.json", "ibtt_zinc.json", "agtt_post.json"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
