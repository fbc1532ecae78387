"""Shared helpers for the test-suite (no product code here)."""
import hashlib

import numpy as np


def sha_int64(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(np.asarray(a, dtype=np.int64)).tobytes())
    return h.hexdigest()


def gt_examples(task, G, seed=0):
    """Config-1 records + their graph-token texts (oracle a1) + reference-style examples."""
    from gtok import synth
    from oracle import ibtt_oracle as O
    recs = synth.er_config1_numpy(G, seed=seed, task=task)
    texts = [O.graph_token_text(r["n"], r["edges"], task, r["label"], r["query"]) for r in recs]
    ex = []
    for r, t in zip(recs, texts):
        if task == "shortest_path":
            e = {"text": t, "label": O.distance_label(t), "query_u": r["query"][0], "query_v": r["query"][1]}
        else:
            e = {"text": t, "label": O.yes_no_label(t)}
        ex.append(e)
    return recs, texts, ex
