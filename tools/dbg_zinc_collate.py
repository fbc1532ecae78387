import sys, os, json
sys.path[:0] = ["/root/repo", "/root/repo/glearning-benchmark_b200", "/root/repo/tests"]
import numpy as np, torch
from gtok import engine
from gtok.batch import GraphBatch
from oracle import ibtt_oracle as O
dev = torch.device("cuda:0")
gz = json.load(open("/root/repo/tests/golden/ibtt_zinc.json"))
def zb(mols):
    return GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                                 x_lists=[m["x"] for m in mols], attr_lists=[m["attr"] for m in mols]).to(dev)
rng = np.random.default_rng(11)
sets = [(gz["mols"], gz["L1024"]["vocab"])]
vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in rng.permutation(40)] + ["X"])
mols = []
for i in range(700):
    n = int(rng.integers(1, 46)); m = int(rng.integers(0, 120)) if i % 9 else 0
    mols.append({"n": n, "src": rng.integers(0, n, m).tolist(), "dst": rng.integers(0, n, m).tolist(),
                 "x": rng.integers(-1, 12, n).tolist(), "attr": rng.integers(0, 7, m).tolist()})
sets.append((mols, vocab))
for B in (32, 128, 256):
  for si, (ms, voc) in enumerate(sets):
    gb = zb(ms)
    for max_len in (1024, 121, 64):
        ep, lens = engine.zinc_collate_epoch(gb, voc, max_len, B)
        want = [O.zinc_ids(m["x"], m["src"], m["dst"], m["attr"], voc, max_len, max_len) for m in ms]
        okl = lens.cpu().tolist() == [len(w) for w in want]
        bad = []
        for k in range(ep.num_batches):
            Xw, Aw, _ = O.collate(want[k * B:(k + 1) * B], [0] * len(want[k * B:(k + 1) * B]), voc["<pad>"])
            X, A = ep.batch(k)
            X, A = X.cpu().numpy(), A.cpu().numpy()
            if X.shape != Xw.shape or not np.array_equal(X, Xw) or not np.array_equal(A, Aw):
                bad.append(k)
                if len(bad) == 1:
                    print("  shape", X.shape, Xw.shape)
                    if X.shape == Xw.shape:
                        d = np.argwhere(X != Xw); da = np.argwhere(A != Aw)
                        print("  X diffs", len(d), d[:5].tolist(), [ (int(X[i,j]), int(Xw[i,j])) for i,j in d[:5]], "A diffs", len(da), da[:5].tolist())
        print(f"B={B} set={si} max_len={max_len} lens_ok={okl} bad_batches={bad[:10]} of {ep.num_batches}")
