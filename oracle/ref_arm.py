"""Times the reference's OWN implementation of the hot path (oracle/_ref, see oracle/make_ref.py) on host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: used by bench.py's `--impl reference` arm and `cpu_baseline` legs.

What runs, unmodified, per shard of synthetic graphs (the same generators as the GPU arm):
  ibtt_zinc          ZINCTokenizationDataset.__getitem__ (tokenize_molecule + truncation,
                     zinc_dataset_indexbase.py:143-227) -> TokenDataset (data_loader.py:465-486) -> collate per
                     batch (data_loader.py:488-497); the per-molecule PyG Data objects are prebuilt int64 tensors
                     (what `ZINC(...)[idx]` hands out), the vocab is the fixed ZINC vocab + dynamic tokens
                     (train_ibtt.py:361-372);
  ibtt_graph_token   build_vocab_from_texts + TokenDataset + collate over graph-token texts (the external
                     serializer's output; texts are produced before the clock starts, as the reference reads
                     them from disk).
The AutoGraph tokenizer is not part of the reference checkout, so SENT has no reference arm (oracle port only).
"""
from __future__ import annotations

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ROOT = os.path.dirname(HERE)


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "graph_data_loader", "data_loader.py"))


def _import_ref():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)          # shadows this repo's drop-in package of the same name
    for m in [k for k in sys.modules if k == "graph_data_loader" or k.startswith("graph_data_loader.")]:
        if not (getattr(sys.modules[m], "__file__", "") or "").startswith(REF_DIR):
            del sys.modules[m]
    import graph_data_loader as R
    from graph_data_loader import zinc_dataset_indexbase as RZ
    assert R.__file__.startswith(REF_DIR), R.__file__
    return R, RZ


def zinc_shard(seed: int, count: int, max_len: int = 1024, batch: int = 128):
    """-> (graphs, tokens, seconds) of the reference's ZINC IBTT path on `count` synthetic molecules."""
    import torch
    torch.set_num_threads(1)
    for p in (ROOT, os.path.join(ROOT, "glearning-benchmark_b200")):
        if p not in sys.path:
            sys.path.append(p)
    from gtok import synth
    R, RZ = _import_ref()
    gb = synth.zinc_shaped(count, seed=seed)

    class D:
        pass
    items = []
    x64, s64, d64, a64 = gb.x.to(torch.int64), gb.src.to(torch.int64), gb.dst.to(torch.int64), gb.edge_attr.to(torch.int64)
    no, eo = gb.node_off.tolist(), gb.edge_off.tolist()
    for g in range(count):
        o = D()
        o.x = x64[no[g]:no[g + 1]].reshape(-1, 1)
        o.edge_index = torch.stack([s64[eo[g]:eo[g + 1]], d64[eo[g]:eo[g + 1]]])
        o.edge_attr = a64[eo[g]:eo[g + 1]]
        o.y = gb.y[g:g + 1].to(torch.float32)
        items.append(o)
    ds = RZ.ZINCTokenizationDataset.__new__(RZ.ZINCTokenizationDataset)
    ds.split, ds.max_len, ds.zinc_dataset = "train", max_len, items
    vocab, _ = R.build_fixed_zinc_vocab()
    vocab = R.extend_vocab_with_dynamic_tokens(vocab, [str(k) for k in range(64)] + ["X", "unknown"])
    t0 = time.perf_counter()
    ex = [ds[i] for i in range(count)]                       # train_ibtt.py:231-233
    td = R.TokenDataset(ex, vocab, max_len)                  # train_ibtt.py:395-397
    ntok = 0
    for b0 in range(0, len(td), batch):                      # DataLoader(collate_fn=collate), train_ibtt.py:399-402
        X, A, Y = R.collate([td[i] for i in range(b0, min(b0 + batch, len(td)))], vocab["<pad>"])
        ntok += int(A.sum())
    return count, ntok, time.perf_counter() - t0


def graph_token_shard(seed: int, count: int, max_len: int = 600, max_vocab: int = 600, batch: int = 128):
    """-> (graphs, tokens, seconds) of build_vocab_from_texts + TokenDataset + collate on `count` ER graph-token texts."""
    import torch
    torch.set_num_threads(1)
    for p in (ROOT, os.path.join(ROOT, "glearning-benchmark_b200")):
        if p not in sys.path:
            sys.path.append(p)
    from gtok import synth
    from oracle import ibtt_oracle as O
    R, _ = _import_ref()
    recs = synth.er_config1_numpy(count, seed=seed)
    texts = [O.graph_token_text(r["n"], r["edges"], "cycle_check", r["label"]) for r in recs]
    ex = [{"text": t, "label": R.data_loader.parse_yes_no_from_text(t)} for t in texts]
    t0 = time.perf_counter()
    vocab, _ = R.build_vocab_from_texts(texts, max_tokens=max_vocab)
    td = R.TokenDataset(ex, vocab, max_len)
    ntok = 0
    for b0 in range(0, len(td), batch):
        X, A, Y = R.collate([td[i] for i in range(b0, min(b0 + batch, len(td)))], vocab["<pad>"])
        ntok += int(A.sum())
    return count, ntok, time.perf_counter() - t0
