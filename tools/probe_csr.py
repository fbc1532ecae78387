#!/usr/bin/env python
"""GPU probe: CSR build (K7 library kernels vs the sort-based torch build) on 1M graphs.  python tools/probe_csr.py [G]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth

dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
for name, gb in (("ER graph-token", synth.er_graph_token(G, seed=0, device=dev)), ("ZINC-shaped (labeled)", synth.zinc_shaped(G, seed=0, device=dev))):
    for native in (True, False):
        for _ in range(2):
            engine.build_csr(gb, native=native)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            c = engine.build_csr(gb, native=native)
        torch.cuda.synchronize()
        print(f"{name}: {'K7 kernels' if native else 'torch sort build'}: {(time.perf_counter() - t) / 3 * 1e3:.2f} ms per {G} graphs "
              f"({c.col.numel()} CSR entries)")
