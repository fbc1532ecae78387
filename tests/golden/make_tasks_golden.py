#!/usr/bin/env python
"""Generate tests/golden/tasks_labels.json by running the REFERENCE's own label oracle and query scan.
Build container only (the GPU box has no /root/reference; it consumes the committed fixture).

    python tests/golden/make_tasks_golden.py

Executed from the reference, unmodified:
  test/train_performance_test.py  has_cycle_dfs (:27-67) - the module imports torch_geometric and a pre-refactor
      dataset module at import time, so the FUNCTION is lifted out of the file with `ast` and compiled as it stands;
  the first-<q> scan of trainer/train_ibtt.py:91-96, `(x[b] == q_token_id).nonzero(as_tuple=True)[0]`, evaluated
      with torch on random rows (an expression, not a function: quoted, not imported).
Inputs: seeded random multigraph edge lists (repeated edges and self loops included) and token rows.
"""
import ast
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/test/train_performance_test.py"


def reference_has_cycle_dfs():
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "has_cycle_dfs")
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["has_cycle_dfs"]


def main():
    f = reference_has_cycle_dfs()
    rng = np.random.default_rng(20251018)
    cases = []
    for i in range(400):
        n = int(rng.integers(1, 48))
        kind = i % 4
        if kind == 0:                                   # forest-ish: random parent links, a few dropped
            edges = [(int(rng.integers(0, v)), v) for v in range(1, n) if rng.random() < 0.85]
        elif kind == 1:                                 # ER with p ~ U(0, 0.2), each edge once
            p = rng.uniform(0, 0.2)
            edges = [(u, v) for u in range(n) for v in range(u + 1, n) if rng.random() < p]
        elif kind == 2:                                 # forest + one repeated edge or one self loop
            edges = [(int(rng.integers(0, v)), v) for v in range(1, n)]
            if edges and rng.random() < 0.5:
                a, b = edges[int(rng.integers(0, len(edges)))]
                edges.append((b, a) if rng.random() < 0.5 else (a, b))
            else:
                k = int(rng.integers(0, n))
                edges.append((k, k))
        else:                                           # anything
            m = int(rng.integers(0, 2 * n))
            edges = [(int(a), int(b)) for a, b in zip(rng.integers(0, n, m), rng.integers(0, n, m))]
        order = rng.permutation(len(edges))
        edges = [edges[j] for j in order]
        cases.append({"n": n, "edges": edges, "has_cycle": bool(f(n, edges))})
    rows = []
    for _ in range(100):
        row = rng.integers(0, 8, int(rng.integers(1, 70))).tolist()
        q = int(rng.integers(0, 9))
        nz = (torch.tensor(row) == q).nonzero(as_tuple=True)[0]
        rows.append({"row": row, "q": q, "pos": int(nz[0]) if len(nz) else -1})
    out = os.path.join(HERE, "tasks_labels.json")
    json.dump({"source": "test/train_performance_test.py:27-67 has_cycle_dfs; trainer/train_ibtt.py:91-96",
               "cycle_cases": cases, "query_rows": rows}, open(out, "w"), separators=(",", ":"))
    print(out, os.path.getsize(out), "bytes;", sum(c["has_cycle"] for c in cases), "of", len(cases), "cyclic")


if __name__ == "__main__":
    main()
