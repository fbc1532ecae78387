#!/usr/bin/env python
"""BASELINE configs[4]: large-graph sweep, N in {64..4096}, mean degree ~8, IBTT (graphs -> padded batches) and
SENT (+ collate of the trails), batch size x graph size.  CUDA events, inputs resident.  One JSON line per point.
   python tools/sweep_large.py [total_nodes_per_point]"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth

dev = torch.device("cuda:0")
NODES = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000       # nodes per sweep point (graphs = NODES / N)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


vocab = {t: i for i, t in enumerate(["<pad>", "<bos>", "<e>", "<n>", "<q>", "<p>", "<eos>", "yes", "no"])}
for k in range(4096):
    vocab[str(k)] = len(vocab)
vocab["has_cycle"] = len(vocab)

for N in (64, 128, 256, 512, 1024, 2048, 4096):
    G = max(64, NODES // N)
    gb = synth.er_sparse(G, N, mean_degree=8.0, seed=N, device=dev)
    E = gb.src.numel()
    for B in (8, 128, 1024):
        if B > G:
            continue
        max_len = 3 * int(gb.num_entries().max()) + N + 5
        ep, lens = engine.ibtt_collate_epoch(gb, vocab, "cycle_check", max_len, B)
        cells = int(ep.xoff[-1])
        X, A = ep.X, ep.attn
        ms = timeit(lambda: engine.ibtt_collate_epoch(gb, vocab, "cycle_check", max_len, B, max_nodes=N, X=X, attn=A))
        byt = 8 * E + 32 * G + 9 * cells
        print(json.dumps({"kind": "ibtt_collate", "N": N, "G": G, "B": B, "ms": round(ms, 4), "graphs_per_s": G / ms * 1e3,
                          "tokens_per_s": int(lens.sum()) / ms * 1e3, "GBps": byt / ms / 1e6, "frac": byt / ms / 1e6 / PEAK}))
        del ep, X, A
    csr = engine.build_csr(gb)
    Lcap = 2 + 2 * N + 2 * (E // G + N) + 64
    Lcap = min(Lcap + (Lcap & 1), 65534)
    cfg = engine.SentConfig(idx_offset=6, max_length=Lcap)
    out = torch.empty((G, Lcap), dtype=torch.int64, device=dev)
    ms = timeit(lambda: engine.tokenize_sent(csr, cfg, seed=1, out=out, stride=Lcap), steps=3, warm=1)
    _, ln = engine.tokenize_sent(csr, cfg, seed=1, out=out, stride=Lcap)
    tot = int(ln.sum())
    byt = csr.sent_input_bytes(False) + 8 * tot + 4 * G
    print(json.dumps({"kind": "sent", "N": N, "G": G, "ms": round(ms, 4), "graphs_per_s": G / ms * 1e3, "tokens_per_s": tot / ms * 1e3,
                      "mean_len": tot / G, "truncated": int((ln >= Lcap).sum()), "GBps": byt / ms / 1e6, "frac": byt / ms / 1e6 / PEAK}))
    off = torch.arange(G, device=dev, dtype=torch.int64) * Lcap
    ms = timeit(lambda: engine.pad_mask_epoch(out.view(-1), off, ln, 5, 128 if G >= 128 else G), steps=3, warm=1)
    ep = engine.pad_mask_epoch(out.view(-1), off, ln, 5, 128 if G >= 128 else G)
    cells = int(ep.xoff[-1])
    byt = 8 * tot + 9 * cells
    print(json.dumps({"kind": "sent_collate", "N": N, "G": G, "ms": round(ms, 4), "GBps": byt / ms / 1e6, "frac": byt / ms / 1e6 / PEAK}))
    if engine.sent_collate_supported(csr, cfg):      # walk -> item log -> emission into the padded batches (<= 256 nodes)
        Bc = 128 if G >= 128 else G
        X, A = ep.X, ep.attn
        ep1, l1 = engine.sent_collate_epoch(csr, cfg, 5, Bc, seed=1, X=X.clone(), attn=A.clone())
        same = bool(torch.equal(l1, ln) and torch.equal(ep1.X, ep.X) and torch.equal(ep1.attn, ep.attn))
        ms = timeit(lambda: engine.sent_collate_epoch(csr, cfg, 5, Bc, seed=1, X=X, attn=A), steps=3, warm=1)
        byt = csr.sent_input_bytes(False) + 9 * cells + 4 * G
        print(json.dumps({"kind": "sent_epoch_fused", "N": N, "G": G, "ms": round(ms, 4), "identical_to_two_step": same,
                          "GBps": byt / ms / 1e6, "frac": byt / ms / 1e6 / PEAK}))
        del ep1, X, A
    del out, csr, gb, ep
    torch.cuda.empty_cache()
