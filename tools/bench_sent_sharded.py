#!/usr/bin/env python
"""BASELINE configs[3]: AGTT SENT trail tokenization of ZINC-shaped graphs, 10M graphs sharded by index across the
GPUs of one box (labeled walk + fused remap_zinc_tokens, then collate_fn with pad id 5).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sent_sharded.py [total_graphs]

No collective on the data path; rank r owns graphs [r*G/N, (r+1)*G/N) and keys its Philox words by the GLOBAL graph id,
so the trails do not depend on N.  Prints one JSON line (rank 0): aggregate graphs/s and tokens/s, max over ranks."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
import torch.distributed as dist
from gtok import engine, synth

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
TOTAL = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
G = TOTAL // world
CHUNK = 1_250_000                                   # graphs per launch (bounds the [G, stride] trail buffer: 3.8 GB)
STRIDE, BATCH = 384, 128
steps = 3

chunks = []
for c0 in range(0, G, CHUNK):
    n = min(CHUNK, G - c0)
    gb = synth.zinc_shaped(n, seed=1000 * rank + c0 // CHUNK, device=dev)
    chunks.append((c0, n, engine.build_csr(gb)))
    del gb
nmax = max(c.max_nodes for _, _, c in chunks)
cfg = engine.SentConfig(idx_offset=6, node_idx_offset=6 + nmax, edge_idx_offset=6 + nmax + 9, max_length=STRIDE, labeled=True,
                        remap_zinc=True)
buf = torch.empty((min(CHUNK, G), STRIDE), dtype=torch.int64, device=dev)
row_start = torch.arange(min(CHUNK, G), device=dev, dtype=torch.int64) * STRIDE
Xcap = torch.empty(min(CHUNK, G) * 200, dtype=torch.int64, device=dev)
Acap = torch.empty(min(CHUNK, G) * 200, dtype=torch.bool, device=dev)


def epoch(seed):
    toks = 0
    for c0, n, csr in chunks:
        _, ln = engine.tokenize_sent(csr, cfg, seed=seed, graph_id_base=rank * G + c0, out=buf[:n], stride=STRIDE)
        engine.pad_mask_epoch(buf.view(-1), row_start[:n], ln, 5, BATCH, X=Xcap, attn=Acap)
        toks = toks + ln.sum()
    return toks


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


epoch(0)
barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for s in range(steps):
    toks = epoch(1 + s)
b.record()
barrier()
t = torch.tensor([a.elapsed_time(b) / steps, float(toks)], dtype=torch.float64, device=dev)
if world > 1:
    tm = t.clone()
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms, tokens = float(tm[0]), float(t[1])
else:
    ms, tokens = float(t[0]), float(t[1])
if rank == 0:
    print(json.dumps({"workload": f"AGTT SENT (labeled + remap) + collate_fn, {G * world} ZINC-shaped graphs over {world} GPU(s)",
                      "n_gpus": world, "graphs_per_gpu": G, "ms_per_epoch": ms, "graphs_per_s": G * world / ms * 1e3,
                      "tokens_per_s": tokens / ms * 1e3, "scaling": "strong"}))
if world > 1:
    dist.destroy_process_group()
