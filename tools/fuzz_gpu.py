#!/usr/bin/env python
"""Randomised CUDA-vs-oracle campaign beyond the unit tests (bit-exactness on odd shapes):
   python tools/fuzz_gpu.py [rounds] [seed]
ZINC: molecules up to 64 atoms / 256 columns, shuffled / duplicated / self-loop columns, out-of-range atom and bond
types, random max_len (all truncation branches), batch sizes 32..256, graph counts around tile multiples; ragged,
padded and packed-layout kernels.  Graph-token: random tasks / max_len / OOV vocabularies, fused collate with shuffled
row order.  SENT: graphs up to 900 nodes (thread-per-graph word counts, both warp-per-graph kernels), labeled / remap / query /
truncation.  Around the path: all device generator families, task labels, query positions."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import numpy as np
import torch
from gtok import engine
from gtok.batch import GraphBatch, PackedGraphBatch
from gtok.pipeline import ZincIbttPipeline
from oracle import ibtt_oracle as O
from oracle import sent_oracle as S

dev = torch.device("cuda:0")
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0


def zinc_round():
    global bad
    nmax = int(rng.choice([9, 33, 37, 64]))
    G = int(rng.choice([1, 31, 127, 128, 129, 257, 1000, 3001]))
    dense = rng.random() < 0.3
    mols = []
    canonical = rng.random() < 0.35            # a PyG-style undirected edge_index: both directions, sorted, symmetric types
    for i in range(G):
        n = int(rng.integers(1, nmax + 1))
        m = 0 if rng.random() < 0.05 else int(rng.integers(1, 257 if dense else 90))
        s = rng.integers(0, n, m)
        d = rng.integers(0, n, m) if rng.random() < 0.7 else (s + 1) % n
        if canonical:
            pairs = sorted({(int(min(a, b)), int(max(a, b))) for a, b in zip(s, d) if a != b})[:128]
            at = {p_: int(rng.integers(0, 16)) for p_ in pairs}
            cols = sorted(pairs + [(b, a) for a, b in pairs])
            if cols and i % 50 == 49 and rng.random() < 0.5:
                cols = cols[::-1]              # one graph breaks the order: the whole batch must keep all its columns
            mols.append({"n": n, "src": [c[0] for c in cols], "dst": [c[1] for c in cols], "x": rng.integers(-2, 40, n).tolist(),
                         "attr": [at[(min(c), max(c))] for c in cols]})
            continue
        mols.append({"n": n, "src": s.tolist(), "dst": d.tolist(), "x": rng.integers(-2, 40, n).tolist(),
                     "attr": rng.integers(-1, 16, m).tolist()})
    names = [str(k) for k in rng.permutation(70)[: int(rng.integers(1, 70))]] + (["X"] if rng.random() < 0.7 else []) + \
            (["unknown"] if rng.random() < 0.7 else [])
    vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), names)
    if rng.random() < 0.3:
        vocab = {t: (i if i < 22 else i + 70000) for t, i in vocab.items()}
    max_len = int(rng.choice([1024, 2, 3, 7, 40, 97, 200]))
    B = int(rng.choice([32, 64, 128, 256]))
    gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                               x_lists=[m["x"] for m in mols], attr_lists=[m["attr"] for m in mols])
    want = [O.zinc_ids(m["x"], m["src"], m["dst"], m["attr"], vocab, max_len, max_len) for m in mols]
    rag = engine.tokenize_zinc(gb.to(dev), vocab, max_len)
    rows = rag.rows_cpu()
    ok = all(r.tolist() == w.tolist() for r, w in zip(rows, want))
    ep, lens = engine.zinc_collate_epoch(gb.to(dev), vocab, max_len, B)
    ok = ok and lens.cpu().tolist() == [len(w) for w in want]
    for k in range(ep.num_batches):
        Xw, Aw, _ = O.collate(want[k * B:(k + 1) * B], [0] * len(want[k * B:(k + 1) * B]), vocab["<pad>"])
        X, A = ep.batch(k)
        ok = ok and np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
    # packed layout needs values in range: clip a copy and compare against the int32 path on the same clipped data
    gc = GraphBatch(gb.node_off, gb.edge_off, gb.src, gb.dst, gb.x.clamp(0, 255), gb.edge_attr.clamp(0, 15))
    pf = PackedGraphBatch.from_batch(gc, fuse_columns=True)
    if pf.col16 is not None:
        pipe = ZincIbttPipeline(gc, vocab, max_len, B, dev, mode="collate")
        a = pipe.run(gc.pin_memory())
        tot = int(a.host_xoff[-1])
        Xa, Aa, la = a.padded.X[:tot].clone(), a.padded.attn[:tot].clone(), a.host_lengths.clone()
        for pk_ in (pf, PackedGraphBatch.from_batch(gc, fuse_columns=True, drop_mirror_columns=True)):
            b = pipe.run(pk_.pin_memory())
            ok = ok and int(b.host_xoff[-1]) == tot and torch.equal(b.host_lengths, la) and torch.equal(b.padded.X[:tot], Xa) \
                and torch.equal(b.padded.attn[:tot], Aa)
    if not ok:
        bad += 1
    print(f"zinc   G={G:5d} nmax={nmax:2d} dense={int(dense)} canonical={int(canonical)} max_len={max_len:4d} B={B:3d}: {'ok' if ok else 'MISMATCH'}")


def gt_round():
    global bad
    task = str(rng.choice(["cycle_check", "shortest_path"]))
    G = int(rng.choice([1, 7, 255, 256, 1000]))
    nn, el, qs = [], [], []
    for i in range(G):
        n = int(rng.integers(1, 60))
        m = int(rng.integers(0, 4 * n))
        nn.append(n)
        el.append(list(zip(rng.integers(0, n, m).tolist(), rng.integers(0, n, m).tolist())))
        qs.append((int(rng.integers(0, n)), int(rng.integers(0, n))))
    vocab = {t: i for i, t in enumerate(O.SPECIAL)}
    for k in rng.permutation(60)[: int(rng.integers(1, 60))]:
        vocab[str(int(k))] = len(vocab) + (70000 if rng.random() < 0.1 else 0)
    vocab["has_cycle"] = max(vocab.values()) + 1
    vocab["shortest_distance"] = max(vocab.values()) + 1
    max_len = int(rng.choice([600, 5, 33, 250]))
    B = int(rng.choice([1, 37, 128]))
    gb = GraphBatch.from_lists(nn, el, queries=qs if task == "shortest_path" else None).to(dev)
    want = [np.asarray(O.ibtt_ids(n, np.asarray([a for a, _ in e], dtype=np.int64), np.asarray([b for _, b in e], dtype=np.int64),
                                  task, vocab, max_len, q if task == "shortest_path" else None)) for n, e, q in zip(nn, el, qs)]
    rows = engine.tokenize_ibtt(gb, vocab, task, max_len).rows_cpu()
    ok = all(r.tolist() == w.tolist() for r, w in zip(rows, want))
    perm = torch.from_numpy(rng.permutation(G)).to(dev)
    ep, _ = engine.ibtt_collate_epoch(gb, vocab, task, max_len, B, batch_idx=perm)
    order = perm.cpu().tolist()
    for k in range(ep.num_batches):
        idx = order[k * B:(k + 1) * B]
        Xw, Aw, _ = O.collate([want[i] for i in idx], [0] * len(idx), vocab["<pad>"])
        X, A = ep.batch(k)
        ok = ok and np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(A.cpu().numpy(), Aw)
    if not ok:
        bad += 1
    print(f"gtoken G={G:5d} task={task:13s} max_len={max_len:4d} B={B:3d}: {'ok' if ok else 'MISMATCH'}")


def sent_round():
    global bad
    nmax = int(rng.choice([5, 32, 64, 100, 128, 200, 256, 300, 900]))
    G = int(rng.choice([1, 33, 120]))
    labeled = rng.random() < 0.5
    mols = []
    for i in range(G):
        n = int(rng.integers(1, nmax + 1))
        m = int(rng.integers(0, 5 * n)) if rng.random() < 0.9 else int(min(n * n // 3, 6000))
        mols.append({"n": n, "src": rng.integers(0, n, m).tolist(), "dst": rng.integers(0, n, m).tolist(),
                     "x": rng.integers(0, 12, n).tolist(), "attr": rng.integers(1, 6, m).tolist()})
    gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                               x_lists=[m["x"] for m in mols], attr_lists=[m["attr"] for m in mols],
                               queries=[(int(rng.integers(0, m["n"])), int(rng.integers(0, m["n"]))) for m in mols]).to(dev)
    csr = engine.build_csr(gb, balance_window=int(rng.choice([0, 16])))
    mx = max(m["n"] for m in mols)
    csr_ok = True
    if mx <= 256:                                               # library CSR builder == sort-based torch build
        ref_csr = engine.build_csr(gb, native=False)
        csr_ok = all(torch.equal(getattr(csr, f), getattr(ref_csr, f)) for f in ("rowptr", "col", "eattr", "adj64"))
    max_length = int(rng.choice([1, 2, 17, 200, 30000]))
    remap, append = bool(labeled and rng.random() < 0.5), bool(rng.random() < 0.4)
    seed = int(rng.integers(0, 2 ** 40))
    restart, order = str(rng.choice(["open_first", "unvisited"])), str(rng.choice(["visit_asc", "visit_desc"]))
    eos = bool(rng.random() < 0.8)
    cfg = engine.SentConfig(idx_offset=6, node_idx_offset=6 + mx, edge_idx_offset=6 + mx + 12, max_length=max_length, labeled=labeled,
                            remap_zinc=remap, append_query=append, append_eos=eos, restart=restart, bracket_order=order)
    use_ws = [None, True, False][int(rng.integers(0, 3))]       # above 256 nodes: HBM-workspace or shared-memory walk state
    tok, ln = engine.tokenize_sent(csr, cfg, seed=seed, graph_id_base=5, use_workspace=use_ws)
    tok, ln = tok.cpu(), ln.cpu().tolist()
    ocfg = S.SentConfig(6, 6 + mx, 6 + mx + 12, max_length, labeled, append_eos=eos)
    spec = S.SentSpec(restart=restart, bracket_order=order)
    ok = csr_ok
    if engine.sent_collate_supported(csr, cfg):                 # walk -> item log -> emission into the padded batches
        B = int(rng.choice([1, 7, 32, 128]))
        perm = torch.from_numpy(rng.permutation(G)).to(dev) if rng.random() < 0.5 else None
        t2, l2 = engine.tokenize_sent(csr, cfg, seed=seed, graph_id_base=5)
        ep2 = engine.pad_mask_epoch(t2.view(-1), torch.arange(G, device=dev, dtype=torch.int64) * t2.shape[1], l2, 5, B, batch_idx=perm)
        ep1, l1 = engine.sent_collate_epoch(csr, cfg, 5, B, seed=seed, graph_id_base=5, batch_idx=perm)
        ok = ok and torch.equal(l1, l2) and torch.equal(ep1.X, ep2.X) and torch.equal(ep1.attn, ep2.attn)
    for g, m in enumerate(mols):
        rp, col, ea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"])
        w = S.sent_tokenize(m["n"], rp, col, ocfg, S.PhiloxWords(seed, 5 + g), x=m["x"], eattr=ea, spec=spec)
        if remap:
            w = O.remap_zinc(w, 6, 6 + mx, 6 + mx + 12)
        if append:
            w = O.append_query(w, 6, m["n"], int(gb.query_u[g]), int(gb.query_v[g]))
        ok = ok and tok[g, :ln[g]].tolist() == np.asarray(w).tolist()
    if not ok:
        bad += 1
    print(f"sent   G={G:5d} nmax={nmax:3d} labeled={int(labeled)} remap={int(remap)} query={int(append)} max_length={max_length:5d} {restart}/{order} eos={int(eos)} ws={use_ws}: "
          f"{'ok' if ok else 'MISMATCH'}")


def tasks_round():
    """device generators (all families) + task labels + query positions vs oracle/tasks_oracle.py"""
    global bad
    from oracle import tasks_oracle as T
    kind = str(rng.choice(["er", "sbm", "complete", "star", "path", "ba", "sfn"]))
    nb = int(rng.integers(1, 5))
    lows = sorted(int(v) for v in rng.integers(1, 60, nb))
    buckets = tuple((lo, lo + int(rng.integers(1, 6))) for lo in lows)            # at most 64 nodes (ba)
    p0 = float(rng.uniform(0, 0.4))
    kw = dict(kind=kind, buckets=buckets, p_min=p0, p_max=p0 + float(rng.uniform(0, 0.5)),
              num_blocks=int(rng.integers(1, 9)), out_ratio=float(rng.uniform(0, 1)))
    G, seed, base = int(rng.choice([1, 33, 257])), int(rng.integers(0, 2 ** 62)), int(rng.integers(0, 2 ** 40))
    gb = engine.generate_er(G, seed=seed, graph_id_base=base, device=dev, **kw)
    n, eo = gb.num_nodes().cpu().tolist(), gb.edge_off.cpu().tolist()
    src, dst = gb.src.cpu().tolist(), gb.dst.cpu().tolist()
    ok, recs = True, []
    for g in range(G):
        wn, we, wq = T.er_graph(seed, base + g, **kw)
        ok = ok and n[g] == wn and list(zip(src[eo[g]:eo[g + 1]], dst[eo[g]:eo[g + 1]])) == we \
            and (int(gb.query_u[g]), int(gb.query_v[g])) == wq
        recs.append((wn, we, wq))
    cyc, dist = engine.graph_labels(engine.build_csr(gb), edge_off=gb.edge_off, distance=True)
    ok = ok and cyc.cpu().tolist() == [int(T.has_cycle_dfs(a, e)) for a, e, _ in recs]
    ok = ok and dist.cpu().tolist() == [T.shortest_distance(a, e, u, v) for a, e, (u, v) in recs]
    vocab = {t: i for i, t in enumerate(O.SPECIAL)}
    for k in range(70):
        vocab[str(k)] = len(vocab)
    vocab["shortest_distance"] = len(vocab)
    max_len = int(rng.choice([600, 40, 4000]))
    ep, _ = engine.ibtt_collate_epoch(gb, vocab, "shortest_path", max_len, 16)
    qp = engine.query_positions(ep, vocab["<q>"]).cpu().tolist()
    want = [T.first_token_pos(O.ibtt_ids(a, np.asarray([u for u, _ in e], dtype=np.int64), np.asarray([v for _, v in e], dtype=np.int64),
                                         "shortest_path", vocab, max_len, q).tolist(), vocab["<q>"]) for a, e, q in recs]
    ok = ok and qp == want
    if not ok:
        bad += 1
    print(f"tasks  G={G:5d} kind={kind:8s} buckets={buckets} max_len={max_len:4d}: {'ok' if ok else 'MISMATCH'}")


only = sys.argv[3] if len(sys.argv) > 3 else "all"          # "sent": SENT rounds only
for r in range(rounds):
    if only in ("all", "zinc"):
        zinc_round()
    if only in ("all", "gt"):
        gt_round()
    if only in ("all", "sent"):
        sent_round()
    if only in ("all", "tasks"):
        tasks_round()
print("mismatching rounds:", bad)
sys.exit(1 if bad else 0)
