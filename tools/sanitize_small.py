"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): smoke() plus the SENT size
classes, the CSR builders, the text parser and the generators at sizes that finish under the tool."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
import __graft_entry__ as E
from gtok import engine, synth

E.smoke()
dev = torch.device("cuda:0")
for N, G in ((40, 96), (150, 40), (400, 12), (3000, 3)):                 # walk+emit / single kernel, mid, warp, workspace kernels
    gb = synth.er_sparse(G, N, mean_degree=6.0, seed=N, device=dev)
    csr = engine.build_csr(gb, balance_window=32)
    cfg = engine.SentConfig(idx_offset=6, max_length=4 * N + 64)
    a, la = engine.tokenize_sent(csr, cfg, seed=1)
    if N <= 64:
        b, lb = engine.tokenize_sent(csr, cfg, seed=1, use_workspace=False)
        assert torch.equal(la, lb)
        engine.sent_collate_epoch(csr, cfg, 5, 16, seed=1)
    if N > 256:
        b, lb = engine.tokenize_sent(csr, cfg, seed=1, use_workspace=True)
        assert torch.equal(la, lb)
zb = synth.zinc_shaped(300, seed=3, device=dev)
zc = engine.build_csr(zb)
nmax = int(zb.num_nodes().max())
zcfg = engine.SentConfig(6, 6 + nmax, 6 + nmax + 9, 384, True, remap_zinc=True)
engine.sent_collate_epoch(zc, zcfg, 5, 32, seed=2)
for kind in ("er", "sbm", "ba", "sfn", "path"):
    eb = engine.generate_er(200, seed=4, kind=kind, device=dev, **({"buckets": ((5, 40),), "num_blocks": 3} if kind in ("ba", "sfn") else {}))
    engine.graph_labels(engine.build_csr(eb), edge_off=eb.edge_off, distance=True)
torch.cuda.synchronize()
print("sanitize_small: ok")
