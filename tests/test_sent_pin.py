"""SENT specification switches and the AutoGraph pin.

* CPU: tools/pin_autograph.py end to end in --self-test mode (the oracle under a non-default SentSpec plays AutoGraph): the
  search has to find the producing specification from outputs alone and hand back replayable choice streams.
* CPU / GPU, only when tests/golden/sent_autograph.json exists (written by tools/pin_autograph.py on a box that has
  BorgwardtLab/AutoGraph): every recorded AutoGraph sequence is reproduced token for token by the oracle - and, on a GPU, by
  the CUDA kernels - from its recovered choice stream.  Until that file exists SENT parity is UNPINNED.
* GPU: the switches the kernels implement (restart policy, bracket order) against the oracle, on all four kernel families.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import sent_oracle as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIN = os.path.join(ROOT, "tests", "golden", "sent_autograph.json")


def test_pin_tool_self_test(tmp_path):
    out = tmp_path / "pin.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pin_autograph.py"), "--self-test", "--out", str(out),
                        "--draws", "50"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = json.load(open(out))
    assert d["source"]["pinned"] is False and len(d["samples"]) >= 100
    specs = [(f["spec"]["restart"], f["spec"]["bracket_order"], f["edge_label_shift"]) for f in d["fits"]]
    assert ("frontier_first", "visit_desc", 0) in specs            # the producing specification is among the survivors
    assert all(f["spec"]["bracket_order"] == "visit_desc" and not f["spec"]["revisit_closes"] for f in d["fits"])
    _replay_fixture(d, None)


def test_pin_tool_without_autograph_says_unpinned():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pin_autograph.py"), "--out", os.devnull],
                       capture_output=True, text=True, env=dict(os.environ, AUTOGRAPH_ROOT=""))
    if r.returncode == 0:
        pytest.skip("AutoGraph is importable here")
    assert r.returncode == 2 and "UNPINNED" in r.stderr


def _replay_fixture(d, dev):
    """every sample of a pin fixture from its recovered stream: oracle always, CUDA when `dev` is given and the kernels
    implement the fitted specification."""
    assert d["fits"], "no specification setting reproduces the recorded samples"
    fit = d["fits"][0]
    spec = S.SentSpec(**fit["spec"])
    native = (spec.restart in ("open_first", "unvisited") and spec.bracket_order in ("visit_asc", "visit_desc") and
              not (spec.bracket_prev or spec.node_label_first or spec.edge_label_after or spec.revisit_closes))
    for smp, stream in zip(d["samples"], fit["streams"]):
        c = smp["cfg"]
        cfg = S.SentConfig(c["idx_offset"], c["node_idx_offset"], c["edge_idx_offset"], c["max_length"], c["labeled"],
                           fit["edge_label_shift"])
        rp, col, ea = S.canonical_csr(smp["n"], [e[0] for e in smp["edges"]], [e[1] for e in smp["edges"]], smp["attr"])
        got = S.sent_tokenize(smp["n"], rp, col, cfg, S.StreamWords(stream), smp["x"], ea, spec)
        assert got.tolist() == smp["tokens"], smp["name"]
    if dev is None or not native:
        return native
    from gtok import engine
    from gtok.batch import GraphBatch
    for labeled in (False, True):
        idx = [i for i, s_ in enumerate(d["samples"]) if s_["labeled"] == labeled]
        if not idx:
            continue
        ss = [d["samples"][i] for i in idx]
        gb = GraphBatch.from_lists([s_["n"] for s_ in ss], [[tuple(e) for e in s_["edges"]] for s_ in ss],
                                   x_lists=[s_["x"] for s_ in ss] if labeled else None,
                                   attr_lists=[s_["attr"] for s_ in ss] if labeled else None).to(dev)
        csr = engine.build_csr(gb)
        words = [np.asarray(fit["streams"][i], dtype=np.uint64).astype(np.uint32) for i in idx]
        woff = np.zeros(len(ss) + 1, dtype=np.int64)
        woff[1:] = np.cumsum([len(w) for w in words])
        cw = torch.from_numpy(np.concatenate(words + [np.zeros(1, np.uint32)]).view(np.int32)).to(dev)
        c = ss[0]["cfg"]
        cfg = engine.SentConfig(c["idx_offset"], c["node_idx_offset"], c["edge_idx_offset"], c["max_length"], labeled,
                                fit["edge_label_shift"], restart=spec.restart, bracket_order=spec.bracket_order)
        tok, ln = engine.tokenize_sent(csr, cfg, choice_words=cw, choice_off=torch.from_numpy(woff).to(dev))
        tok, ln = tok.cpu(), ln.cpu().tolist()
        for g, s_ in enumerate(ss):
            assert tok[g, :ln[g]].tolist() == s_["tokens"], s_["name"]
    return True


@pytest.mark.skipif(not os.path.exists(PIN), reason="tests/golden/sent_autograph.json absent: SENT parity UNPINNED "
                                                     "(run tools/pin_autograph.py on a box with BorgwardtLab/AutoGraph)")
def test_oracle_reproduces_autograph_fixture():
    _replay_fixture(json.load(open(PIN)), None)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(PIN), reason="tests/golden/sent_autograph.json absent: SENT parity UNPINNED")
def test_cuda_reproduces_autograph_fixture():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    assert _replay_fixture(json.load(open(PIN)), torch.device("cuda:0")), "the fitted specification is oracle-only"


@pytest.mark.gpu
def test_cuda_replays_self_test_fixture(tmp_path):
    """the whole pin path on the GPU with the oracle standing in for AutoGraph under a specification the kernels implement"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import pin_autograph as P
    truth = S.SentSpec(restart="unvisited", bracket_order="visit_desc")

    def tokenize_of(labeled):
        cfg = S.SentConfig(6, 6 + 64, 6 + 64 + 9, 512, labeled)

        def tok(n, edges, x, attr, r):
            rp, col, ea = S.canonical_csr(n, [u for u, v in edges], [v for u, v in edges], attr)
            return S.sent_tokenize(n, rp, col, cfg, S.PhiloxWords(5, 100 * r + n), x, ea, truth)
        return tok, cfg
    samples = P.collect(tokenize_of, 2, 64, 512)
    fits = [f for f in P.fit(samples) if f["spec"]["restart"] == "unvisited" and f["edge_label_shift"] == 0]
    assert fits
    for s_ in samples:
        for k in ("rowptr", "col", "eattr"):
            s_.pop(k)
    assert _replay_fixture({"samples": samples, "fits": fits}, torch.device("cuda:0")) is True


@pytest.mark.gpu
@pytest.mark.parametrize("N", [40, 100, 300, 5000])
@pytest.mark.parametrize("restart,order", [("unvisited", "visit_asc"), ("open_first", "visit_desc"), ("unvisited", "visit_desc")])
def test_sent_spec_switches_vs_oracle(N, restart, order):
    """restart policy / bracket order on the <= 64-node, <= 256-node, shared-memory and workspace kernels."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gtok import engine
    from gtok.batch import GraphBatch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(N)
    mols = []
    for i in range(24 if N < 1000 else 6):
        n = int(rng.integers(max(2, N // 2), N + 1))
        m = int(rng.integers(n // 2, 4 * n)) if i % 4 else int(min(n * 8, 30000))   # sparse ones restart often
        s_, d_ = rng.integers(0, n, m), rng.integers(0, n, m)
        mols.append({"n": n, "src": s_.tolist(), "dst": d_.tolist(), "x": rng.integers(0, 9, n).tolist(),
                     "attr": rng.integers(1, 5, m).tolist()})
    gb = GraphBatch.from_lists([m["n"] for m in mols], [list(zip(m["src"], m["dst"])) for m in mols],
                               x_lists=[m["x"] for m in mols], attr_lists=[m["attr"] for m in mols]).to(dev)
    csr = engine.build_csr(gb)
    nmax = max(m["n"] for m in mols)
    spec = S.SentSpec(restart=restart, bracket_order=order)
    for labeled in (False, True):
        for ws in ((None,) if N <= 256 else (False, True)):
            cfg = engine.SentConfig(6, 6 + nmax, 6 + nmax + 9, 100000, labeled, restart=restart, bracket_order=order)
            tok, ln = engine.tokenize_sent(csr, cfg, seed=17, use_workspace=ws)
            tok, ln = tok.cpu(), ln.cpu().tolist()
            ocfg = S.SentConfig(6, 6 + nmax, 6 + nmax + 9, 100000, labeled)
            for g, m in enumerate(mols):
                rp, col, ea = S.canonical_csr(m["n"], m["src"], m["dst"], m["attr"])
                want = S.sent_tokenize(m["n"], rp, col, ocfg, S.PhiloxWords(17, g), x=m["x"], eattr=ea, spec=spec)
                assert tok[g, :ln[g]].tolist() == want.tolist(), (N, labeled, ws, g)
