import os, sys, subprocess, pickle
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "glearning-benchmark_b200"), os.path.join(ROOT, "tests")]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from gtok import engine
    from oracle import ibtt_oracle as O
    from gtok.batch import GraphBatch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    vocab = O.zinc_extend_vocab(O.zinc_fixed_vocab(), [str(k) for k in rng.permutation(40)] + ["X"])
    mols = []
    for i in range(700):
        n = int(rng.integers(1, 46))
        m = int(rng.integers(0, 120)) if i % 9 else 0
        mols.append({"n": n, "src": rng.integers(0, n, m).tolist(), "dst": rng.integers(0, n, m).tolist(),
                     "x": rng.integers(-1, 12, n).tolist(), "attr": rng.integers(0, 7, m).tolist()})
    gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                               [m["x"] for m in mols], [m["attr"] for m in mols]).to(dev)
    out = {}
    for B in (32, 128):
        for max_len in (1024, 121, 64):
            ep, lens = engine.zinc_collate_epoch(gb, vocab, max_len, B)
            out[(B, max_len)] = (ep.X.cpu(), ep.attn.cpu(), ep.xoff.cpu(), ep.lmax.cpu(), lens.cpu())
    pickle.dump(out, open(sys.argv[2], "wb"))
    sys.exit(0)
res = {}
for t in ("0", "1"):
    env = dict(os.environ, GTOK_ZINC_TMA=t)
    subprocess.run([sys.executable, __file__, "child", f"/tmp/dbg_{t}.pkl"], env=env, check=True)
    res[t] = pickle.load(open(f"/tmp/dbg_{t}.pkl", "rb"))
import torch
for key in res["0"]:
    a, b = res["0"][key], res["1"][key]
    okx, oka = torch.equal(a[0], b[0]), torch.equal(a[1], b[1])
    print(key, "X equal", okx, "attn equal", oka, "xoff equal", torch.equal(a[2], b[2]), "lens", torch.equal(a[4], b[4]))
    for name, u, v in (("X", a[0], b[0]), ("attn", a[1], b[1])):
        if not torch.equal(u, v):
            bad = (u != v).nonzero().flatten()
            xoff, lmax = a[2].tolist(), a[3].tolist()
            print("  ", name, "mismatches", bad.numel(), "first", bad[:10].tolist())
            c = int(bad[0])
            import bisect
            k = bisect.bisect_right(xoff, c) - 1
            r, p = divmod(c - xoff[k], lmax[k])
            print("   batch", k, "lmax", lmax[k], "row", r, "pos", p, "len", int(a[4][k * key[0] + r]), "want", u[c - 3:c + 5].tolist(), "got", v[c - 3:c + 5].tolist())
            rows = sorted(set(((int(i) - xoff[bisect.bisect_right(xoff, int(i)) - 1]) // lmax[bisect.bisect_right(xoff, int(i)) - 1], bisect.bisect_right(xoff, int(i)) - 1) for i in bad[:2000].tolist()))
            print("   (row,batch) with mismatches:", rows[:30])
