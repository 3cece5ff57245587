"""A model whose HMMs have five emitting states (hmm_vit_eval_5st_lr, ref: src/hmm.c:166-304):
no bundled model has them, so tests/model_variants.py stretches the en-us senone sequences to
[a, a, b, c, c] and writes seeded 5 x 6 left-to-right transition matrices.  Fixtures
(tests/golden/five_state_en-us.npz) come from the unmodified reference aligning the test
utterance with that model."""
import os

import numpy as np
import pytest

import soundswallower_b200 as ssb
import model_variants as mv
from conftest import GOLDEN, model_dir
from test_oracle_golden import sha


@pytest.fixture(scope="module")
def five(tmp_path_factory):
    d = mv.write_five_state_model(model_dir("en-us"), str(tmp_path_factory.mktemp("five") / "five"))
    return d, np.load(os.path.join(GOLDEN, "five_state_en-us.npz"))


def _cases(g0):
    feat, words = g0["feat"], g0["words"]
    rs = np.random.RandomState(5)
    noisy = feat + rs.randn(*feat.shape).astype(np.float32) * np.float32(0.2)
    return [("win", feat, words, True), ("nowin", feat, words, False), ("noisy", noisy, words, True),
            ("short", feat[:120], words[:3], False)]


def _same_on_path(r, want, words, parent, win):
    """States on the best path carry the alignment; states the path skipped (5-state matrices
    have skip arcs) keep what alignment_populate put there: the word's window, score 0
    (ref: src/ps_alignment.c:237-240, src/state_align_search.c:236-263)."""
    on = r["dur"] > 0
    assert on.sum() >= 0.6 * len(on) and not on.all()
    got = np.stack([r["start"], r["dur"], r["score"]], 1)
    assert np.array_equal(got[on], want[on, 1:4])
    wi = np.repeat(parent, 5)[~on]
    z = np.zeros(len(wi), np.int32)
    assert np.array_equal(want[~on, 1], words[wi, 1] if win else z)
    assert np.array_equal(want[~on, 2], words[wi, 2] if win else z) and not want[~on, 3].any()
    return True


def _chain(lx, words, win):
    n = len(words)
    z = np.zeros(n, np.int32)
    return lx.populate(words[:, 0], words[:, 1] if win else z, words[:, 2] if win else z)


def test_loader_and_oracle_on_five_state_model(five, golden):
    from oracle.oracle import Oracle
    d, g = five
    m = ssb.AcousticModel(d, device=-1)
    assert m.n_emit == 5 and m.arrays()["tp"].shape == (42, 5, 6)
    assert sha(m.arrays()["tp"]) == str(g["tp_sha"]) and sha(m.arrays()["sseq"]) == str(g["sseq_sha"])
    lx = ssb.Lexicon(m, hmmdir=d)
    o = Oracle(d)
    for name, feat, words, win in _cases(golden["en-us"]):
        c = _chain(lx, words, win)
        r = o.state_align(feat, c["ssid"], c["tmat"], c["sf"], c["ef"], want_tokens=True)
        want = g[name + "_states"]
        assert r["rv"] == int(g[name + "_rv"]) == 0 and r["best_score"] == int(g[name + "_best"]), name
        assert _same_on_path(r, want, words, c["parent"], win), name
        assert np.array_equal(m.arrays()["sseq"][c["ssid"]].reshape(-1), want[:, 0]), name
        assert sha(r["tokens"]) == str(g[name + "_tokens_sha"]), name


def _two_pass_inputs(golden):
    feat = golden["en-us"]["feat"]
    rs = np.random.RandomState(5)
    noisy = feat + rs.randn(*feat.shape).astype(np.float32) * np.float32(0.2)
    return [("fsg", feat), ("fsg_noisy", noisy)]


def test_oracle_grammar_search_with_five_state_hmms(five, golden):
    """fsg_search (default mode) + decoder_alignment on one decoder, 5-state HMMs throughout."""
    from oracle.oracle import Oracle
    d, g = five
    m = ssb.AcousticModel(d, device=-1)
    lx = ssb.Lexicon(m, hmmdir=d)
    o = Oracle(d)
    G = lx.align_graph("go forward ten meters")
    for name, feat in _two_pass_inputs(golden):
        p1 = o.fsg_search_active(G, feat)
        assert p1["rv"] == 0 and np.array_equal(p1["hist"], g[name + "_hist"])
        assert p1["hyp_score"] == int(g[name + "_hyp_score"]) and p1["n_sen_eval"] == int(g[name + "_n_sen_eval"])
        assert p1["n_hmm_eval"] == int(g[name + "_n_hmm_eval"]) and np.array_equal(p1["active"], g[name + "_active"])
        segs = g[name + "_segs"]
        segs = segs[segs[:, 0] >= 0]
        c = lx.populate(segs[:, 0], segs[:, 1], segs[:, 2] - segs[:, 1] + 1)
        left = [int(w * 32 + b) for w, v in enumerate(p1["active"]) for b in range(32) if (int(v) >> b) & 1]
        r = o.state_align(feat, c["ssid"], c["tmat"], c["sf"], c["ef"], init_active=left, init_topn=p1["carried"])
        on = r["dur"] > 0
        want = g[name + "_p2_states"]
        assert r["rv"] == int(g[name + "_p2_rv"]) == 0
        assert np.array_equal(np.stack([r["start"], r["dur"], r["score"]], 1)[on], want[on, 1:4])


@pytest.mark.gpu
def test_two_passes_with_five_state_hmms(five, golden):
    """K4 (active lists, hmm_step5) -> words -> chain_viterbi_kernel<5>, through ssb_fsg_batch /
    ssb_align_batch and through ssb_align_texts."""
    d, g = five
    m = ssb.AcousticModel(d)
    lx = ssb.Lexicon(m, hmmdir=d)
    text = "go forward ten meters"
    G = lx.align_graph(text)
    inputs = _two_pass_inputs(golden)
    p1 = ssb.fsg_batch(m, [f for _n, f in inputs], [G], want_hist=True, compallsen=False)
    ta = ssb.TextAlignment(m, lx, [f for _n, f in inputs], [text] * len(inputs), align_level=2)
    for u, ((name, feat), r) in enumerate(zip(inputs, p1)):
        assert r["rv"] == 0 and np.array_equal(r["hist"], g[name + "_hist"]), name
        assert r["hyp_score"] == int(g[name + "_hyp_score"]) and r["n_sen_eval"] == int(g[name + "_n_sen_eval"])
        assert r["n_hmm_eval"] == int(g[name + "_n_hmm_eval"]) and np.array_equal(r["active"], g[name + "_active"])
        assert np.array_equal(r["segs"][:, 1:], g[name + "_segs"][:, 1:])
        # dense mode searches the same graph with hmm_step5 too
        assert ssb.fsg_batch(m, [feat], [G])[0]["exit"] > 0
        want = g[name + "_p2_states"]
        st = ta.entries(u, "states")
        assert ta.status(u)[0] == 0 and ta.status(u)[1] == int(g[name + "_hyp_score"])
        assert np.array_equal(st, want), name     # skipped states keep alignment_populate's values


@pytest.mark.gpu
def test_chain_viterbi_with_five_state_hmms(five, golden):
    """chain_viterbi_kernel<5> + backtrace: state segmentations and the whole token stack."""
    d, g = five
    m = ssb.AcousticModel(d)
    lx = ssb.Lexicon(m, hmmdir=d)
    cases = _cases(golden["en-us"])
    chains = [_chain(lx, words, win) for _n, _f, words, win in cases]
    res = ssb.align_batch(m, [c[1] for c in cases], chains, want_tokens=True)
    for (name, _f, words, win), r, c in zip(cases, res, chains):
        want = g[name + "_states"]
        assert r["rv"] == 0 and r["best_score"] == int(g[name + "_best"]), name
        assert _same_on_path(r, want, words, c["parent"], win), name
        assert sha(r["tokens"]) == str(g[name + "_tokens_sha"]), name
    # the search-module drop-in and the vtable scorer work on the model as well
    a = ssb.state_align_search(m, lx, golden["en-us"]["words"][:, 0], golden["en-us"]["words"][:, 1],
                               golden["en-us"]["words"][:, 2])
    a.start(); a.forward(golden["en-us"]["feat"]); assert a.finish() == 0
    assert np.array_equal(a.alignment("states")[:, :4], g["win_states"][:, :4])   # populate's values kept
