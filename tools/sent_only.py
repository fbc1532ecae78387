"""One SENT workload for ncu: python tools/sent_only.py [er|zinc] [fused]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth
dev = torch.device("cuda:0")
G = 1_000_000
which = sys.argv[1] if len(sys.argv) > 1 else "er"
fused = len(sys.argv) > 2 and sys.argv[2] == "fused"
if which == "er":
    gb = synth.er_graph_token(G, seed=0, device=dev)
    cfg = engine.SentConfig(idx_offset=6, max_length=600)
    stride = 600
else:
    gb = synth.zinc_shaped(G, seed=0, device=dev)
    nmax = int(gb.num_nodes().max())
    cfg = engine.SentConfig(idx_offset=6, node_idx_offset=6 + nmax, edge_idx_offset=6 + nmax + 9, max_length=384, labeled=True,
                            remap_zinc=True)
    stride = 384
csr = engine.build_csr(gb, balance_window=2048)
out = torch.empty((G, stride), dtype=torch.int64, device=dev)
X = A = None
for _ in range(3):
    if fused:
        ep, _ = engine.sent_collate_epoch(csr, cfg, 5, 128, seed=0, X=X, attn=A)
        X, A = ep.X, ep.attn
    else:
        engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=stride)
torch.cuda.synchronize()
print("ok")
