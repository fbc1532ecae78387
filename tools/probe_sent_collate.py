#!/usr/bin/env python
"""GPU probe: SENT + collate_fn of an epoch, two-step (tokenize_sent + pad_mask_epoch) vs fused (sent_collate_epoch: walk kernel ->
item log -> emission straight into the padded batches).  python tools/probe_sent_collate.py [G] [er|zinc|both] [batch]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth

dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
WHICH = sys.argv[2] if len(sys.argv) > 2 else "both"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
PEAK = 6444.4


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


def run(name, batch, cfg, stride, labeled, window, pad=5):
    csr = engine.build_csr(batch, balance_window=window)
    out = torch.empty((G, stride), dtype=torch.int64, device=dev)
    row_start = torch.arange(G, device=dev, dtype=torch.int64) * stride
    _, ln = engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=stride)
    ep2 = engine.pad_mask_epoch(out.view(-1), row_start, ln, pad, B)
    X2, A2 = ep2.X.clone(), ep2.attn.clone()
    X = torch.empty_like(X2)
    A = torch.empty_like(A2)
    ep1, ln1 = engine.sent_collate_epoch(csr, cfg, pad, B, seed=0, X=X, attn=A)
    assert torch.equal(ln, ln1) and torch.equal(ep1.xoff, ep2.xoff) and torch.equal(ep1.lmax, ep2.lmax)
    assert torch.equal(X, X2) and torch.equal(A, A2), "fused != two-step"
    t_sent = timeit(lambda: engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=stride))
    t_two = timeit(lambda: engine.pad_mask_epoch(engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=stride)[0].view(-1), row_start,
                                                 ln, pad, B, X=X2, attn=A2))
    t_fused = timeit(lambda: engine.sent_collate_epoch(csr, cfg, pad, B, seed=0, X=X, attn=A))
    tot = int(ln.sum())
    cells = int(ep2.xoff[-1])
    b_sent = csr.sent_input_bytes(labeled) + 8 * tot + 4 * G
    b_fused = csr.sent_input_bytes(labeled) + 9 * cells + 4 * G
    print(f"{name}: G={G} B={B} avgL={tot / G:.1f} cells/G={cells / G:.1f}  sent {t_sent:.3f} ms (frac {b_sent / t_sent / 1e6 / PEAK:.3f})  "
          f"sent+collate two-step {t_two:.3f} ms  fused {t_fused:.3f} ms (frac {b_fused / t_fused / 1e6 / PEAK:.3f})  identical=True")


if WHICH in ("both", "er"):
    gb = synth.er_graph_token(G, seed=0, device=dev)
    run("ER unlabeled", gb, engine.SentConfig(idx_offset=6, max_length=600), 600, False, 2048)
    del gb
if WHICH in ("both", "zinc"):
    zb = synth.zinc_shaped(G, seed=0, device=dev)
    nmax = int(zb.num_nodes().max())
    cfgz = engine.SentConfig(idx_offset=6, node_idx_offset=6 + nmax, edge_idx_offset=6 + nmax + 9, max_length=384, labeled=True,
                             remap_zinc=True)
    run("ZINC labeled+remap", zb, cfgz, 384, True, 2048)
