"""One large-graph sweep point for ncu: python tools/sent_sparse_only.py N [total_nodes]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth
dev = torch.device("cuda:0")
N = int(sys.argv[1])
NODES = int(sys.argv[2]) if len(sys.argv) > 2 else 32_000_000
G = max(64, NODES // N)
gb = synth.er_sparse(G, N, mean_degree=8.0, seed=N, device=dev)
E = gb.src.numel()
Lcap = 2 + 2 * N + 2 * (E // G + N) + 64
Lcap = min(Lcap + (Lcap & 1), 65534)
csr = engine.build_csr(gb)
cfg = engine.SentConfig(idx_offset=6, max_length=Lcap)
out = torch.empty((G, Lcap), dtype=torch.int64, device=dev)
for _ in range(3):
    engine.tokenize_sent(csr, cfg, seed=1, out=out, stride=Lcap)
torch.cuda.synchronize()
print("ok")
