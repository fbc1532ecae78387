#!/usr/bin/env python
"""GPU probe: the two SENT kernels timed apart (walk alone through gtok_sent_walk; emission = total - walk).
   python tools/probe_sent_split.py [G] [er|zinc|both] [window]"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth
from gtok.engine import lib, _ptr, _stream, _csr_c, _sent_c, check

dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
WHICH = sys.argv[2] if len(sys.argv) > 2 else "both"
BW = int(sys.argv[3]) if len(sys.argv) > 3 else 2048


def timeit(fn, steps=10, warm=3):
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


def run(name, batch, cfg, stride, window):
    csr = engine.build_csr(batch, balance_window=window)
    out = torch.empty((G, stride), dtype=torch.int64, device=dev)
    engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=stride)
    sc, ch = _sent_c(cfg, 0, 0, None, None)
    cc = _csr_c(csr)
    n_total = int(csr.rowptr.numel()) - 1
    nbytes = lib.gtok_sent_log_bytes(C.byref(cc), C.byref(sc), n_total)
    ln = torch.empty(G, dtype=torch.int32, device=dev)
    walk = lambda: check(lib.gtok_sent_walk(C.byref(cc), C.byref(sc), C.byref(ch), n_total, _ptr(csr.scratch), nbytes, _ptr(ln), _stream()))
    t_walk = timeit(walk)
    t_all = timeit(lambda: engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=stride))
    print(f"{name} (window {window}): walk {t_walk:.3f} ms  emit {t_all - t_walk:.3f} ms  total {t_all:.3f} ms")


if WHICH in ("both", "er"):
    gb = synth.er_graph_token(G, seed=0, device=dev)
    run("ER", gb, engine.SentConfig(idx_offset=6, max_length=600), 600, BW)
    del gb
if WHICH in ("both", "zinc"):
    zb = synth.zinc_shaped(G, seed=0, device=dev)
    nmax = int(zb.num_nodes().max())
    cfgz = engine.SentConfig(idx_offset=6, node_idx_offset=6 + nmax, edge_idx_offset=6 + nmax + 9, max_length=384, labeled=True,
                             remap_zinc=True)
    run("ZINC", zb, cfgz, 384, 0)
    run("ZINC", zb, cfgz, 384, BW)
if WHICH == "sparse":                                      # large-graph sweep point: python probe_sent_split.py NODES sparse N
    N = BW
    Gs = max(64, G // N)
    G = Gs
    gb = synth.er_sparse(Gs, N, mean_degree=8.0, seed=N, device=dev)
    E = gb.src.numel()
    Lcap = 2 + 2 * N + 2 * (E // Gs + N) + 64
    Lcap = min(Lcap + (Lcap & 1), 65534)
    run(f"sparse N={N} G={Gs}", gb, engine.SentConfig(idx_offset=6, max_length=Lcap), Lcap, 0)
