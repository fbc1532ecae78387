"""SENT oracle self-checks (the oracle is the normative spec; parity to AutoGraph is unpinned):
Philox known-answer vectors, validity (every edge covered exactly once, decodes to the input
graph), determinism / replay, the doc illustration, and the host-side CSR builder."""
import numpy as np
import pytest
import torch

from oracle import sent_oracle as S


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert S.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert S.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert S.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_doc_illustration_four_cycle():
    """docs/models/embed_gtt.md:17-19: the 4-cycle starts '<sos> 0 1 2 3 ...' (first-visit indices);
    the closing edge is carried by the bracketed neighbourhood set."""
    rp, col, _ = S.canonical_csr(4, [0, 1, 2, 0], [1, 2, 3, 3])
    cfg = S.SentConfig()
    for seed in range(8):
        t = S.sent_tokenize(4, rp, col, cfg, S.PhiloxWords(seed, 0)).tolist()
        assert t == [S.SOS, 6, 7, 8, 9, S.LADJ, 6, S.RADJ, S.EOS]
        n, edges, _ = S.sent_decode(t, cfg)
        assert n == 4 and set(edges) == {(0, 1), (1, 2), (2, 3), (0, 3)}


def _random_graph(rng, n, p, labeled):
    m = rng.random((n, n)) < p
    src, dst = [], []
    for u in range(n):
        for v in range(u + 1, n):
            if m[u, v]:
                src.append(u); dst.append(v)
    attr = rng.integers(1, 5, len(src)).tolist() if labeled else None
    x = rng.integers(0, 9, n).tolist() if labeled else None
    return src, dst, attr, x


@pytest.mark.parametrize("labeled", [False, True])
def test_every_edge_once_and_decodes_to_input(labeled):
    rng = np.random.default_rng(0)
    for trial in range(150):
        n = int(rng.integers(1, 40))
        src, dst, attr, x = _random_graph(rng, n, float(rng.uniform(0.02, 0.4)), labeled)
        rp, col, ea = S.canonical_csr(n, src, dst, attr)
        cfg = S.SentConfig(6, 6 + 40, 6 + 40 + 9, 10_000, labeled)
        t = S.sent_tokenize(n, rp, col, cfg, S.PhiloxWords(trial, trial), x=x, eattr=ea)
        assert t[0] == S.SOS and t[-1] == S.EOS
        nn, edges, nlab = S.sent_decode(t, cfg)
        vis = S.visit_order(n, rp, col, S.PhiloxWords(trial, trial))
        assert nn == n and sorted(vis) == list(range(n))
        want = {}
        for i, (a, b) in enumerate(zip(src, dst)):
            want[(min(vis[a], vis[b]), max(vis[a], vis[b]))] = attr[i] if labeled else 0
        assert edges == want                       # sent_decode asserts no edge is covered twice
        if labeled:
            inv = {vis[v]: v for v in range(n)}
            assert nlab == [x[inv[k]] for k in range(n)]


def test_replay_and_truncation():
    rng = np.random.default_rng(3)
    src, dst, _, _ = _random_graph(rng, 30, 0.2, False)
    rp, col, _ = S.canonical_csr(30, src, dst)
    cfg = S.SentConfig(max_length=10_000)
    rec = []
    w = S.PhiloxWords(11, 5)
    t1 = S.sent_tokenize(30, rp, col, cfg, lambda: rec.append(w()) or rec[-1])
    t2 = S.sent_tokenize(30, rp, col, cfg, S.StreamWords(rec))          # explicit choice-stream replay
    assert t1.tolist() == t2.tolist()
    t3 = S.sent_tokenize(30, rp, col, S.SentConfig(max_length=17), S.StreamWords(rec))
    assert t3.tolist() == t1[:17].tolist()
    t4 = S.sent_tokenize(30, rp, col, cfg, S.PhiloxWords(12, 5))
    assert t4.tolist() != t1.tolist()


def test_canonical_csr_rules():
    # duplicates, both directions, self loop, label of the FIRST directed entry wins
    rp, col, ea = S.canonical_csr(4, [2, 0, 1, 0, 3], [0, 2, 1, 1, 0], [5, 6, 7, 8, 9])
    assert rp.tolist() == [0, 3, 4, 5, 6]
    assert col.tolist() == [1, 2, 3, 0, 0, 0]
    assert ea.tolist() == [8, 5, 9, 8, 5, 9]


def test_engine_build_csr_matches_oracle_on_cpu():
    from gtok import engine, synth
    from gtok.batch import GraphBatch
    rng = np.random.default_rng(1)
    nn, el, al, xl = [], [], [], []
    for _ in range(60):
        n = int(rng.integers(1, 20))
        m = int(rng.integers(0, 40))
        nn.append(n)
        el.append([(int(rng.integers(0, n)), int(rng.integers(0, n))) for _ in range(m)])
        al.append(rng.integers(0, 6, m).tolist())
        xl.append(rng.integers(0, 9, n).tolist())
    gb = GraphBatch.from_lists(nn, el, x_lists=xl, attr_lists=al)
    csr = engine.build_csr(gb)
    for g in range(60):
        rp, col, ea = S.canonical_csr(nn[g], [e[0] for e in el[g]], [e[1] for e in el[g]], al[g])
        n0, n1 = int(gb.node_off[g]), int(gb.node_off[g + 1])
        r = csr.rowptr[n0:n1 + 1]
        assert (r - r[0]).tolist() == rp.tolist()
        assert csr.col[int(r[0]):int(r[-1])].tolist() == col.tolist()
        assert csr.eattr[int(r[0]):int(r[-1])].tolist() == ea.tolist()
    assert csr.max_nodes == max(nn)
    z = synth.zinc_shaped(500, seed=3)
    c2 = engine.build_csr(z)
    assert int(c2.rowptr[-1]) == z.src.numel()            # already canonical: both directions, unique
    assert torch.equal(c2.col, z.dst) and torch.equal(c2.eattr, z.edge_attr)


# ---- the open decisions as switches + the fit search behind tools/pin_autograph.py ------------------------------------
def _samples(spec, labeled, count, seed, nmax=14):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        n = int(rng.integers(1, nmax))
        src, dst, attr, x = _random_graph(rng, n, float(rng.uniform(0.1, 0.5)), labeled)
        rp, col, ea = S.canonical_csr(n, src, dst, attr)
        cfg = S.SentConfig(6, 6 + 20, 6 + 20 + 9, 10_000, labeled)
        t = S.sent_tokenize(n, rp, col, cfg, S.PhiloxWords(seed, i), x=x, eattr=ea, spec=spec)
        out.append({"n": n, "rowptr": rp, "col": col, "x": x, "eattr": ea, "tokens": t.tolist(), "cfg": cfg})
    return out


@pytest.mark.parametrize("spec", [S.SentSpec(), S.SentSpec(restart="unvisited", bracket_order="visit_desc"),
                                  S.SentSpec(restart="frontier_first", bracket_order="node_asc", bracket_prev=True,
                                             node_label_first=True, edge_label_after=True),
                                  S.SentSpec(revisit_closes=True)])
@pytest.mark.parametrize("labeled", [False, True])
def test_fit_choice_stream_recovers_a_replayable_stream(spec, labeled):
    """Whatever produced a sequence, if the specification CAN produce it the search finds an explicit choice stream
    that replays it word for word (the pin mechanism: no access to the producer's RNG needed)."""
    for smp in _samples(spec, labeled, 25, seed=3, nmax=10 if spec.revisit_closes else 14):
        st = S.fit_choice_stream(smp["n"], smp["rowptr"], smp["col"], smp["cfg"], smp["tokens"], smp["x"], smp["eattr"], spec)
        assert st is not None
        again = S.sent_tokenize(smp["n"], smp["rowptr"], smp["col"], smp["cfg"], S.StreamWords(st), smp["x"], smp["eattr"], spec)
        assert again.tolist() == smp["tokens"]
    # a sequence the specification cannot produce has no stream
    smp = _samples(spec, labeled, 1, seed=9, nmax=7)[0]
    bad = list(smp["tokens"])
    bad[1] += 1                                            # first node token is always index 0
    assert S.fit_choice_stream(smp["n"], smp["rowptr"], smp["col"], smp["cfg"], bad, smp["x"], smp["eattr"], spec) is None


def test_fit_choice_stream_truncated_target():
    spec = S.SentSpec()
    rng = np.random.default_rng(5)
    src, dst, _, _ = _random_graph(rng, 12, 0.4, False)
    rp, col, _ = S.canonical_csr(12, src, dst)
    cfg = S.SentConfig(max_length=17)
    t = S.sent_tokenize(12, rp, col, cfg, S.PhiloxWords(1, 0), spec=spec)
    assert len(t) == 17
    st = S.fit_choice_stream(12, rp, col, cfg, t, spec=spec)
    assert S.sent_tokenize(12, rp, col, cfg, S.StreamWords(st), spec=spec).tolist() == t.tolist()


def test_fit_spec_identifies_the_producing_specification():
    """Given only outputs, the candidate search keeps the specification that produced them and rejects the default
    when the producer differs from it (what tools/pin_autograph.py does with AutoGraph's outputs)."""
    truth = S.SentSpec(restart="frontier_first", bracket_order="visit_desc", edge_label_after=True)
    smp = _samples(truth, True, 30, seed=11)
    good = S.fit_spec(smp, lambda s_: s_["cfg"])
    specs = [g[0] for g in good]
    assert truth in specs and S.SentSpec() not in specs
    for spec, streams in good:
        for s_, st in zip(smp, streams):
            assert S.sent_tokenize(s_["n"], s_["rowptr"], s_["col"], s_["cfg"], S.StreamWords(st), s_["x"], s_["eattr"],
                                   spec).tolist() == s_["tokens"]


def test_revisit_mode_matches_the_reference_doc_illustration():
    """docs/models/embed_gtt.md:17-19,64 draws the 4-cycle as '<sos> 0 1 2 3 0': under revisit_closes the walk returns
    to the start node and re-emits its index (the default specification carries that edge in a bracket instead)."""
    rp, col, _ = S.canonical_csr(4, [0, 1, 2, 0], [1, 2, 3, 3])
    t = S.sent_tokenize(4, rp, col, S.SentConfig(), S.PhiloxWords(0, 0), spec=S.SentSpec(revisit_closes=True)).tolist()
    assert t == [S.SOS, 6, 7, 8, 9, 6, S.EOS]
