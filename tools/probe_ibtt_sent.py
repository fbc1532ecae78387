#!/usr/bin/env python
"""GPU probe: throughput of K1 (graph-token IBTT) and K3 (SENT) on 1M-graph workloads."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth
from gtok.pipeline import GraphTokenIbttPipeline

dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
BW = int(sys.argv[2]) if len(sys.argv) > 2 else 0
PEAK = 6444.4

def timeit(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps

# ---- K1
gb = synth.er_graph_token(G, seed=0, device=dev)
vocab = {t: i for i, t in enumerate(["<pad>", "<bos>", "<e>", "<n>", "<q>", "<p>", "<eos>", "yes", "no"])}
for k in range(49): vocab[str(k)] = len(vocab)
vocab["has_cycle"] = len(vocab)
host = gb.to("cpu")
pipe = GraphTokenIbttPipeline(host, vocab, "cycle_check", 600, 128, dev)
pipe.use_device_batch(gb)
pipe.timers = []
ms = timeit(pipe.launch)
k = {}
for n, a, b in pipe.timers: k.setdefault(n, []).append(a.elapsed_time(b))
tot, cells = int(pipe.lengths.sum()), int(pipe.xoff[-1])
algo = pipe.algorithmic_bytes(tot, cells)
print(f"K1 ibtt graph-token G={G} tokens={tot} ms/step={ms:.3f} graphs/s={G/ms*1e3/1e6:.1f}M tokens/s={tot/ms*1e3/1e9:.1f}G")
for n, v in k.items():
    v = v[3:]; m = sum(v) / len(v)
    print(f"   {n:18s} {m:.4f} ms {algo[n]/m/1e6:7.0f} GB/s frac {algo[n]/m/1e6/PEAK:.3f}")
pipe.timers = None

# ---- K3 unlabeled (config 3)
csr = engine.build_csr(gb, balance_window=BW)
cfg = engine.SentConfig(idx_offset=6, max_length=600)
out = torch.empty((G, 600), dtype=torch.int64, device=dev)
ms = timeit(lambda: engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=600), steps=5, warm=2)
_, ln = engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=600)
tot = int(ln.sum())
byt = 4 * (csr.rowptr.numel() + G) + 4 * csr.col.numel() + 8 * tot
print(f"K3 sent ER unlabeled G={G} tokens={tot} avgL={tot/G:.1f} ms={ms:.3f} graphs/s={G/ms*1e3/1e6:.1f}M tokens/s={tot/ms*1e3/1e9:.2f}G  {byt/ms/1e6:.0f} GB/s frac {byt/ms/1e6/PEAK:.3f}")
del out
# ---- K3 labeled zinc (config 4 shape, 1M per GPU)
zb = synth.zinc_shaped(G, seed=0, device=dev)
zc = engine.build_csr(zb, balance_window=BW)
cfgz = engine.SentConfig(idx_offset=6, node_idx_offset=6 + 37, edge_idx_offset=6 + 37 + 9, max_length=1024, labeled=True, remap_zinc=True)
stride = 2 + 3 * 37 + 4 * 45 + 40
stride = 384
cfgz.max_length = 384
out = torch.empty((G, stride), dtype=torch.int64, device=dev)
ms = timeit(lambda: engine.tokenize_sent(zc, cfgz, seed=0, out=out, stride=stride), steps=5, warm=2)
_, ln = engine.tokenize_sent(zc, cfgz, seed=0, out=out, stride=stride)
tot = int(ln.sum())
byt = 4 * (zc.rowptr.numel() + G) + 8 * zc.col.numel() + 4 * zc.x.numel() + 8 * tot
print(f"K3 sent ZINC labeled+remap G={G} tokens={tot} avgL={tot/G:.1f} maxL={int(ln.max())} ms={ms:.3f} graphs/s={G/ms*1e3/1e6:.1f}M tokens/s={tot/ms*1e3/1e9:.2f}G  {byt/ms/1e6:.0f} GB/s frac {byt/ms/1e6/PEAK:.3f}")
