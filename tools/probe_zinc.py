#!/usr/bin/env python
"""GPU probe: ZINC tokenizer kernels on 1M molecules (CUDA events, inputs resident).
   python tools/probe_zinc.py [G] ; env GTOK_ZINC_TILE / GTOK_ZINC_PREFETCH tune the tile size / L2 prefetch lead,
   GTOK_ZINC_CTAS_PER_SM sizes the thread-per-molecule staging."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200")]
import torch
from gtok import engine, synth, _lib
from gtok._lib import lib, check
from gtok.engine import _ptr, _stream

dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
B = 128
PEAK = 6444.4


def timeit(fn, steps=20, warm=5):
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


gb = synth.zinc_shaped(G, seed=0, device=dev)
vocab = {t: i for i, t in enumerate(["<bos>", "<eos>", "<pad>", "<unk>", "<q>", "<p>", "<atom>", "<bond>", "C", "N", "O", "F", "P",
                                     "S", "Cl", "Br", "I", "single", "double", "triple", "aromatic", "regression"] +
                                    [str(k) for k in range(37)] + ["X", "unknown"])}
mn, me = int(gb.num_nodes().max()), int(gb.num_entries().max())
bc = engine._batch_c(gb, mn, me)
if os.environ.get("GTOK_GROUP_BOUNDS", "1") != "0":
    bc.max_nodes_per_32, bc.max_entries_per_32 = engine.group_bounds(gb.node_off), engine.group_bounds(gb.edge_off)
    print("group bounds", bc.max_nodes_per_32, bc.max_entries_per_32)
zv, tbl = engine.zinc_vocab_c(vocab, 1024, 1024, mn, dev)
n_total, e_total = gb.x.numel(), gb.src.numel()
cap = 2 * n_total + 4 * e_total + 4 * G + 2
lengths = torch.empty(G, dtype=torch.int32, device=dev)
off = torch.empty(G + 1, dtype=torch.int64, device=dev)
tok = torch.empty(cap, dtype=torch.int64, device=dev)
nbytes = lib.gtok_zinc_tokenize_workspace_bytes(G)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)


def ragged():
    check(lib.gtok_zinc_tokenize(C.byref(bc), C.byref(zv), _ptr(lengths), _ptr(off), _ptr(tok), tok.numel(), _ptr(ws), nbytes, _stream()))


ms = timeit(ragged)
tot = int(off[-1])
byt = 12 * e_total + 4 * n_total + 16 * G + 8 * tot
print(f"zinc_tokenize (ragged) tile={os.environ.get('GTOK_ZINC_TILE', '128')}: "
      f"{ms:.4f} ms  {G / ms / 1e3:.0f} Mgraphs/s  {byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / PEAK:.3f}")

nb = (G + B - 1) // B
xcap = G * min(1024, 2 * mn + 4 * me + 4) + 2
X = torch.empty(xcap, dtype=torch.int64, device=dev)
A = torch.empty(xcap, dtype=torch.bool, device=dev)
lmax = torch.empty(nb, dtype=torch.int32, device=dev)
xoff = torch.empty(nb + 1, dtype=torch.int64, device=dev)
nb2 = lib.gtok_zinc_collate_workspace_bytes(G, B)
ws2 = torch.empty(nb2, dtype=torch.uint8, device=dev)


def fused():
    check(lib.gtok_zinc_collate(C.byref(bc), C.byref(zv), B, 2, _ptr(lengths), _ptr(lmax), _ptr(xoff), _ptr(X), _ptr(A), xcap,
                                _ptr(ws2), nb2, _stream()))


ms = timeit(fused)
cells = int(xoff[-1])
byt = 12 * e_total + 4 * n_total + 16 * G + 9 * cells
print(f"zinc_collate (fused tokenize+pad/mask, B={B}): {ms:.4f} ms  {G / ms / 1e3:.0f} Mgraphs/s  cells={cells}  {byt / ms / 1e6:.0f} GB/s  "
      f"frac {byt / ms / 1e6 / PEAK:.3f}")

# ---- compact resident layout: 16-bit columns + byte atoms, consumed directly
from gtok.batch import PackedGraphBatch
pkb = PackedGraphBatch.from_batch(gb.to("cpu"), fuse_columns=True)
col16, x8 = pkb.col16.to(dev), pkb.x.to(dev)
pk = _lib.PackedMolBatchC(G, _ptr(gb.node_off), _ptr(gb.edge_off), _ptr(col16), _ptr(x8), max(mn, pkb.max_id + 1), me, 0,
                          bc.max_nodes_per_32, bc.max_entries_per_32, 0)


def fused_packed():
    check(lib.gtok_zinc_collate_packed(C.byref(pk), C.byref(zv), B, 2, _ptr(lengths), _ptr(lmax), _ptr(xoff), _ptr(X), _ptr(A), xcap,
                                       _ptr(ws2), nb2, _stream()))


Xref = X[:cells].clone()
ms = timeit(fused_packed)
assert int(xoff[-1]) == cells and torch.equal(X[:cells], Xref)
byt = 2 * e_total + n_total + 16 * G + 9 * cells
print(f"zinc_collate_packed (16-bit columns resident, B={B}): {ms:.4f} ms  {G / ms / 1e3:.0f} Mgraphs/s  {byt / ms / 1e6:.0f} GB/s of its own "
      f"bytes ({byt / 1e9:.2f} GB)  frac {byt / ms / 1e6 / PEAK:.3f}")
