#!/usr/bin/env python
"""GPU probe: SENT kernel on 1M ER graph-token graphs (unlabeled) and 1M ZINC-shaped graphs (labeled+remap).
   python tools/probe_sent.py [G] [balance_window]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth

dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
BW = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
WHICH = sys.argv[3] if len(sys.argv) > 3 else "both"
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


def run(name, batch, cfg, stride, labeled, window):
    csr = engine.build_csr(batch, balance_window=window)
    out = torch.empty((G, stride), dtype=torch.int64, device=dev)
    ms = timeit(lambda: engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=stride))
    _, ln = engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=stride)
    tot = int(ln.sum())
    byt = csr.sent_input_bytes(labeled) + 8 * tot + 4 * G
    print(f"{name}: G={G} avgL={tot / G:.1f} maxL={int(ln.max())} {ms:.3f} ms  {G / ms / 1e3:.0f} Mgraphs/s  "
          f"{tot / ms / 1e6:.1f} Gtok/s  {byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / PEAK:.3f}")


if WHICH in ("both", "er"):
    gb = synth.er_graph_token(G, seed=0, device=dev)
    run("sent ER unlabeled (window %d)" % BW, gb, engine.SentConfig(idx_offset=6, max_length=600), 600, False, BW)
    del gb
if WHICH in ("both", "zinc", "zincb"):
    zb = synth.zinc_shaped(G, seed=0, device=dev)
    nmax = int(zb.num_nodes().max())
    cfgz = engine.SentConfig(idx_offset=6, node_idx_offset=6 + nmax, edge_idx_offset=6 + nmax + 9, max_length=384, labeled=True,
                             remap_zinc=True)
    run("sent ZINC labeled+remap", zb, cfgz, 384, True, BW if WHICH == "zincb" else 0)
