import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth
dev = torch.device("cuda:0")
G = 1_000_000
gb = synth.er_graph_token(G, seed=0, device=dev)
csr = engine.build_csr(gb)
cfg = engine.SentConfig(idx_offset=6, max_length=600)
out = torch.empty((G, 600), dtype=torch.int64, device=dev)
for _ in range(3):
    engine.tokenize_sent(csr, cfg, seed=0, out=out, stride=600)
torch.cuda.synchronize()
print("ok")
