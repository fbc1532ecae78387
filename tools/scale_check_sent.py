#!/usr/bin/env python
"""BASELINE configs[3] at scale, one launch vs shards: the fused SENT epoch (walk -> item log -> emission into the padded batches)
over G ZINC-shaped graphs in ONE call must equal, byte for byte, the same epoch computed shard by shard with graph_id_base =
the shard's first graph (what `bench.py` / a multi-GPU job does), and its rows must carry the structural invariants of a SENT
(SOS first, EOS last, every node index exactly once outside brackets + heads, balanced brackets).
   python tools/scale_check_sent.py [G] [shards]"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth

dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4_096_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B = 128
assert G % (S * B) == 0
gb = synth.zinc_shaped(G, seed=11, device=dev)
nmax = int(gb.num_nodes().max())
cfg = engine.SentConfig(idx_offset=6, node_idx_offset=6 + nmax, edge_idx_offset=6 + nmax + 9, max_length=384, labeled=True)
csr = engine.build_csr(gb, balance_window=2048)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ep, ln = engine.sent_collate_epoch(csr, cfg, 5, B, seed=7)
a.record()
ep, ln = engine.sent_collate_epoch(csr, cfg, 5, B, seed=7, X=ep.X, attn=ep.attn)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
# shard by shard
per = G // S
ok = True
for s in range(S):
    idx = torch.arange(s * per, (s + 1) * per, device=dev)
    sub = gb.select(idx)
    scsr = engine.build_csr(sub, balance_window=2048)
    sep, sln = engine.sent_collate_epoch(scsr, cfg, 5, B, seed=7, graph_id_base=s * per)
    nb = per // B
    lo, hi = int(ep.xoff[s * nb]), int(ep.xoff[(s + 1) * nb])
    ok = ok and torch.equal(sln, ln[s * per:(s + 1) * per]) and torch.equal(sep.X, ep.X[lo:hi]) and torch.equal(sep.attn, ep.attn[lo:hi])
    del sub, scsr, sep, sln
# invariants on the unpadded rows (labeled, unremapped: ids 0..5 special, 6.. node indices, 6+nmax.. labels)
X, A = ep.X, ep.attn
tot = int(ln.sum())
assert int(A.sum()) == tot
n = gb.num_nodes()
first_ok = bool((X[ep.xoff[:-1]] == 0).all())                                  # SOS opens the first row of every batch
idx_tok = (X >= 6) & (X < 6 + nmax) & A
ladj, radj = int(((X == 2) & A).sum()), int(((X == 3) & A).sum())
eos = int(((X == 4) & A).sum())
res = {"graphs": G, "shards": S, "batch": B, "ms_one_launch": ms, "graphs_per_s": G / ms * 1e3, "tokens": tot,
       "shards_identical_to_one_launch": bool(ok), "sos_first": first_ok, "eos_count_equals_graphs": eos == G,
       "brackets_balanced": ladj == radj, "index_tokens_at_least_nodes": int(idx_tok.sum()) >= int(n.sum())}
print(json.dumps(res))
sys.exit(0 if all(v for k, v in res.items() if isinstance(v, bool)) else 1)
