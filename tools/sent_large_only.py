"""one SENT launch on large sparse graphs (ncu target): python tools/sent_large_only.py [N] [G] [smem]
(third argument "smem": keep the walk state in shared memory instead of the HBM workspace)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
G = int(sys.argv[2]) if len(sys.argv) > 2 else 15625
gb = synth.er_sparse(G, N, mean_degree=8.0, seed=N, device=dev)
csr = engine.build_csr(gb)
E = gb.src.numel()
Lcap = 2 + 2 * N + 2 * (E // G + N) + 64
Lcap = min(Lcap + (Lcap & 1), 65534)
cfg = engine.SentConfig(idx_offset=6, max_length=Lcap)
WS = not (len(sys.argv) > 3 and sys.argv[3] == 'smem')
out = torch.empty((G, Lcap), dtype=torch.int64, device=dev)
for _ in range(3):
    engine.tokenize_sent(csr, cfg, seed=1, out=out, stride=Lcap, use_workspace=WS)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
engine.tokenize_sent(csr, cfg, seed=1, out=out, stride=Lcap, use_workspace=WS)
b.record()
torch.cuda.synchronize()
print(f"N={N} G={G} {'workspace' if WS else 'smem'}: {a.elapsed_time(b):.3f} ms")
